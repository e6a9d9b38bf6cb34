"""One network over several GPUs (BASELINE.json configs[4]): the host side.

`split_network` cuts a flat network by a node -> rank map into per-rank sub-networks with ghost
nodes and duplicated cut conduits (include/swmm_b200.h: swb_partition_desc); `PartitionedSolver`
is one rank's solver plus the gather of its owned results back into the global numbering.  The
boundary exchange itself runs inside the persistent kernel (csrc/swb_engine.h: halo_exchange);
the only host-side communication is the one-off swap of the 64-byte window handles and whatever
result gather the caller wants, both through any transport (`torch.distributed` in bench.py and
the tests).

The reference has no counterpart: `dynwave_execute` (dynwave.c:224-262) loops over every link and
node of one address space.  What must hold is that each rank reproduces, for the objects it owns,
exactly the arithmetic the unpartitioned solver performs:
  * links keep their relative order, so node sums are formed in the reference's order (A.3);
  * a cut conduit is computed on both sides from identical inputs (dwflow.c:57-293 reads the two
    end-node depths and the conduit's own state only);
  * only true conduits are cut (the ordered regulator pass, A.4, stays within one rank).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi, solver

HANDLE_BYTES = 64


@dataclass
class Part:
    rank: int
    n_ranks: int
    net: abi.Network
    node_gid: np.ndarray      # local node -> global node (owned first, then ghosts)
    link_gid: np.ndarray      # local link -> global link (ascending)
    n_owned: int
    send_node: np.ndarray
    send_rank: np.ndarray
    send_slot: np.ndarray
    recv_node: np.ndarray
    link_owned: np.ndarray    # 1 = this rank reports the link (owner of its node1)

    def owned_links(self) -> np.ndarray:
        return np.nonzero(self.link_owned)[0]


def stripes(n_nodes_per_row: int, n_rows: int, n_ranks: int, extra_nodes: int = 0) -> np.ndarray:
    """Owner map for a row-major grid (node = row * n_nodes_per_row + col) cut into n_ranks
    contiguous blocks of rows; `extra_nodes` trailing nodes (the outfall) go to the last rank."""
    rows = np.arange(n_rows)
    owner_of_row = np.minimum(rows * n_ranks // n_rows, n_ranks - 1).astype(np.int32)
    owner = np.repeat(owner_of_row, n_nodes_per_row)
    return np.concatenate([owner, np.full(extra_nodes, n_ranks - 1, dtype=np.int32)])


def split_network(net: abi.Network, owner: np.ndarray, n_ranks: int) -> list[Part]:
    A = net.arrays
    nN, nL = net.n_nodes, net.n_links
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    assert owner.size == nN and owner.min() >= 0 and owner.max() < n_ranks
    n1, n2 = A["link_node1"], A["link_node2"]
    o1, o2 = owner[n1], owner[n2]
    cut = o1 != o2
    if np.any(cut & ~net.true_conduit_mask()):
        raise ValueError("a pump / regulator / dummy link crosses the partition border")
    node_arrays = [n for n, _ in abi.DESC_ARRAYS if abi.desc_array_len(n, net.scalars) == nN
                   and n.startswith(("node_", "outfall_", "storage_"))]
    link_arrays = [n for n, _ in abi.DESC_ARRAYS
                   if n not in node_arrays and not n.startswith(("curve_", "shape_", "pollut_"))]
    parts = []
    ghost_lists = []
    for r in range(n_ranks):
        links = np.nonzero((o1 == r) | (o2 == r))[0]                      # ascending global index
        owned = np.nonzero(owner == r)[0]
        ends = np.concatenate([n1[links], n2[links]])
        ghosts = np.unique(ends[owner[ends] != r])
        node_gid = np.concatenate([owned, ghosts]).astype(np.int64)
        g2l = np.full(nN, -1, dtype=np.int64)
        g2l[node_gid] = np.arange(node_gid.size)
        arrays = {}
        for name in node_arrays:
            arrays[name] = A[name][node_gid]
        for name in link_arrays:
            arrays[name] = A[name][links]
        arrays["link_node1"] = g2l[n1[links]].astype(np.int32)
        arrays["link_node2"] = g2l[n2[links]].astype(np.int32)
        for name, _ in abi.DESC_ARRAYS:
            if name not in arrays:
                arrays[name] = A[name].copy()
        sc = dict(net.scalars)
        sc["n_nodes"], sc["n_links"] = int(node_gid.size), int(links.size)
        local = abi.Network(sc, arrays, net.options)
        ghost_lists.append(ghosts)
        parts.append(Part(r, n_ranks, local, node_gid, links.astype(np.int64), int(owned.size),
                          None, None, None,
                          (owned.size + np.arange(ghosts.size)).astype(np.int32),
                          (o1[links] == r).astype(np.int32)))
    # send lists: the owner of g publishes it into slot k of every rank that lists g as ghost k
    for r in range(n_ranks):
        sn, sr, ss = [], [], []
        g2l = np.full(nN, -1, dtype=np.int64)
        g2l[parts[r].node_gid[:parts[r].n_owned]] = np.arange(parts[r].n_owned)
        for q in range(n_ranks):
            if q == r:
                continue
            slots = np.nonzero(owner[ghost_lists[q]] == r)[0]
            sn.append(g2l[ghost_lists[q][slots]])
            sr.append(np.full(slots.size, q))
            ss.append(slots)
        cat = lambda x: np.concatenate(x).astype(np.int32) if x else np.zeros(0, dtype=np.int32)
        parts[r].send_node, parts[r].send_rank, parts[r].send_slot = cat(sn), cat(sr), cat(ss)
        assert parts[r].send_node.size == 0 or parts[r].send_node.min() >= 0
    return parts


def split_state(part: Part, state: dict, n_pollut: int) -> dict:
    """Global single-member state image (field name -> array) -> this rank's image."""
    out = {}
    for k, v in state.items():
        fid = abi.FIELD[k]
        gid = part.node_gid if abi.is_node_field(fid) else part.link_gid
        w = abi.field_width(fid, n_pollut)
        a = np.asarray(v, dtype=np.float64)
        out[k] = a.reshape(-1, w)[gid].reshape(-1) if w > 1 else a[gid]
    return out


def split_inflows(part: Part, inflows: dict, n_pollut: int) -> dict:
    """Keyword arguments of Solver.set_inflows restricted to the nodes this rank owns."""
    node = np.asarray(inflows["node"])
    g2l = {int(g): i for i, g in enumerate(part.node_gid[:part.n_owned])}
    keep = [k for k, g in enumerate(node) if int(g) in g2l]
    ts_start = np.asarray(inflows["ts_start"])
    ts_t, ts_q = np.asarray(inflows["ts_t"]), np.asarray(inflows["ts_q"])
    st, tt, tq = [0], [], []
    for k in keep:
        tt.append(ts_t[ts_start[k]:ts_start[k + 1]])
        tq.append(ts_q[ts_start[k]:ts_start[k + 1]])
        st.append(st[-1] + tt[-1].size)
    out = dict(inflows)
    out["node"] = np.array([g2l[int(node[k])] for k in keep], dtype=np.int32)
    out["ts_start"] = np.array(st, dtype=np.int32)
    out["ts_t"] = np.concatenate(tt) if tt else np.zeros(0)
    out["ts_q"] = np.concatenate(tq) if tq else np.zeros(0)
    out["sfactor"] = np.asarray(inflows["sfactor"])[keep]
    out["baseline"] = np.asarray(inflows["baseline"])[keep]
    if inflows.get("concen") is not None:
        out["concen"] = np.asarray(inflows["concen"]).reshape(-1, max(n_pollut, 1))[keep].reshape(-1)
    return out


class PartitionDesc(C.Structure):
    _fields_ = [(n, t) for (n, t, _, _) in abi._parse_struct(abi._TEXT, "swb_partition_desc")]


class PartitionedSolver(solver.Solver):
    """One rank of a partitioned single-model run.  Usage on every rank:

        ps = PartitionedSolver(part, device=local_rank)
        handles = all_gather(ps.export_handle())      # any transport
        ps.connect(handles)
        ps.load_state(split_state(part, state0, nP)); ps.set_inflows(**split_inflows(...))
        ps.run_steps(n, t_end)                        # every rank, same arguments
    """

    def __init__(self, part: Part, device: int = 0, lib_path: str | None = None, timeout_s: float = 30.0):
        super().__init__(part.net, 1, device=device, lib_path=lib_path)
        self.part = part
        lib = self.lib
        lib.swb_partition_attach.argtypes = [C.c_void_p, C.POINTER(PartitionDesc)]
        lib.swb_partition_export.argtypes = [C.c_void_p, C.c_void_p]
        lib.swb_partition_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.swb_partition_exchanges.argtypes = [C.c_void_p]
        lib.swb_partition_exchanges.restype = C.c_longlong
        d = PartitionDesc()
        d.rank, d.n_ranks, d.n_owned_nodes = part.rank, part.n_ranks, part.n_owned
        d.n_send, d.n_recv = int(part.send_node.size), int(part.recv_node.size)
        d.timeout_s = float(timeout_s)
        self._part_keep = []
        for name in ("send_node", "send_rank", "send_slot", "recv_node", "link_owned"):
            a = np.ascontiguousarray(getattr(part, name), dtype=np.int32)
            self._part_keep.append(a)
            setattr(d, name, a.ctypes.data_as(C.POINTER(C.c_int)))
        self._chk(lib.swb_partition_attach(self._h, C.byref(d)))

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(HANDLE_BYTES)
        self._chk(self.lib.swb_partition_export(self._h, buf))
        return buf.raw

    def connect(self, handles):
        """handles[r] = export_handle() of rank r (own entry ignored)."""
        for r, h in enumerate(handles):
            if r == self.part.rank:
                continue
            buf = C.create_string_buffer(bytes(h), HANDLE_BYTES)
            self._chk(self.lib.swb_partition_connect(self._h, r, buf))

    def exchanges(self) -> int:
        return int(self.lib.swb_partition_exchanges(self._h))

    def owned_field(self, field) -> tuple[np.ndarray, np.ndarray]:
        """(global indices, values) of the objects this rank reports for a field."""
        fid = self._fid(field)
        w = abi.field_width(fid, self.net.n_pollut)
        a = self.get_field(fid)[0]
        if abi.is_node_field(fid):
            loc = np.arange(self.part.n_owned)
            gid = self.part.node_gid[:self.part.n_owned]
        else:
            loc = self.part.owned_links()
            gid = self.part.link_gid[loc]
        return gid, (a.reshape(-1, w)[loc] if w > 1 else a[loc])


def assemble(pieces, n_items: int, width: int = 1) -> np.ndarray:
    """Global array from the (gid, values) pieces of every rank."""
    out = np.zeros((n_items, width)) if width > 1 else np.zeros(n_items)
    for gid, v in pieces:
        out[gid] = v
    return out.reshape(-1)
