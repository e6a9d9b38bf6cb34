// swb_host.h -- host-side preparation shared by the CUDA library (swb_api.cu) and the CPU test
// emulation (tests/emul): validation of a swb_network_desc, the derived static arrays (link flags,
// end-node elevations, CSR incidence in the reference's summation order) and the field table that
// maps swb_field ids onto State members.
#ifndef SWB_HOST_H
#define SWB_HOST_H

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>
#include "swb_state.h"
#include "swb_xsect.h"
#include "swb_hds5_tables.h"

namespace swb {

struct Derived {
    std::vector<int> outfall_nodes, link_flags, adj_start, adj, adjq_start, adjq, nc_links, outfall_link, link_order,
                     outfall_slot, link_pre_node, pre_links,
                     link_kernel, node_order;
    std::vector<double> link_z1, link_z2, xs_tables, culvert_params, road_tables, xs_rcp_yfull, cond_rcp_mod_length, link_rows, link_cols_d;
    std::vector<int> link_cols_i;
    std::vector<AdjEntry> adj_packed;
    int nTrue = 0, nNonConduit = 0;
};

inline size_t desc_count(const swb_network_desc &d, char kind0, char kind1)
{
    switch (kind0) {
      case 'N': return d.n_nodes;
      case 'L': return d.n_links;
      case 'P': return d.n_pollut;
      case 'C': return kind1 == '1' ? d.n_curves + 1 : (kind1 == 'P' ? d.n_curve_pts : d.n_curves);
      case 'T': return kind1 == 'T' ? (size_t)d.n_shape_tbls * d.shape_tbl_len : d.n_shape_tbls;
    }
    return 0;
}

// returns "" when the descriptor is acceptable, else the reason
inline std::string validate_desc(const swb_network_desc &d, const swb_options &o)
{
    if (d.n_nodes <= 0 || d.n_links < 0) return "empty network";
    if (d.n_pollut > SWB_MAX_POLLUT) return "more than SWB_MAX_POLLUT pollutants";
    if (o.max_trials < 1) return "max_trials < 1";
    if (o.max_trials > SWB_MAX_TRIALS_CAP) return "max_trials > 32 is not supported (per-trial bookkeeping rows)";
#define X(T, name, kind) if (desc_count(d, #kind[0], #kind[1]) > 0 && d.name == nullptr) return "null array: " #name;
    SWB_DESC_ARRAYS(X)
#undef X
    for (int j = 0; j < d.n_links; j++) {
        int a = d.link_node1[j], b = d.link_node2[j];
        if (a < 0 || a >= d.n_nodes || b < 0 || b >= d.n_nodes) return "link end node out of range";
        if (d.link_type[j] == SWB_CONDUIT) {
            if (d.xs_culvert[j] > SWB_MAX_CULVERT_CODE) return "culvert code out of range";
            if (d.xs_type[j] == XS_DUMMY &&
                (d.node_type[a] == SWB_STORAGE || d.node_type[a] == SWB_DIVIDER))
                return "dummy conduit out of a storage / divider node";
            if (d.cond_barrels[j] < 1) return "conduit with < 1 barrel";
        }
        int t = d.xs_table[j];
        if (t >= d.n_shape_tbls) return "xs_table out of range";
    }
    return "";
}

// Ticket order of the node phase: outfalls (boundary depth = critical / normal depth solves) and
// storage nodes (curve integration) first, so their long serial latency overlaps the bulk of the
// phase instead of trailing it; the first nFirst entries are a permutation of [0, nFirst).
inline void node_order_expensive_first(const int *node_type, int nFirst, int nN, std::vector<int> &order)
{
    order.clear();
    for (int i = 0; i < nFirst; i++) if (node_type[i] == SWB_OUTFALL || node_type[i] == SWB_STORAGE) order.push_back(i);
    for (int i = 0; i < nFirst; i++) if (!(node_type[i] == SWB_OUTFALL || node_type[i] == SWB_STORAGE)) order.push_back(i);
    for (int i = nFirst; i < nN; i++) order.push_back(i);
}

inline void derive(const swb_network_desc &d, const swb_options &o, Derived &r)
{
    const int nN = d.n_nodes, nL = d.n_links;
    r.link_flags.assign(nL, 0); r.link_z1.assign(nL, 0.0); r.link_z2.assign(nL, 0.0);
    r.outfall_link.assign(nN, -1); r.link_kernel.assign(nL, LK_GENERIC);
    r.xs_rcp_yfull.assign(nL, 0.0); r.cond_rcp_mod_length.assign(nL, 0.0);
    r.nc_links.clear(); r.nTrue = 0;
    for (int j = 0; j < nL; j++) {
        int a = d.link_node1[j], b = d.link_node2[j], f = 0;
        bool trueConduit = (d.link_type[j] == SWB_CONDUIT && d.xs_type[j] != XS_DUMMY);
        if (trueConduit) { f |= LF_TRUE_CONDUIT; r.nTrue++; } else r.nc_links.push_back(j);
        if (d.link_has_flap[j]) f |= LF_HAS_FLAP;
        if (d.node_type[a] == SWB_OUTFALL) { f |= LF_N1_OUTFALL; if (d.outfall_flap[a]) f |= LF_N1_OUT_FLAP; }
        if (d.node_type[b] == SWB_OUTFALL) { f |= LF_N2_OUTFALL; if (d.outfall_flap[b]) f |= LF_N2_OUT_FLAP; }
        if (d.node_type[a] == SWB_STORAGE) f |= LF_N1_STORAGE;
        if (d.node_type[b] == SWB_STORAGE) f |= LF_N2_STORAGE;
        bool open = xs_is_open(d.xs_type[j]);
        if (open) f |= LF_OPEN_SHAPE;
        if (trueConduit && (d.link_seep_rate[j] > 0.0 || open)) f |= LF_HAS_LOSSRATE;
        r.link_flags[j] = f;
        // plain circular / closed rectangular pipes run the specialised conduit functions; culvert
        // inlets (any shape) and force mains need the extra terms only the generic one carries
        if (trueConduit && d.xs_culvert[j] <= 0) {
            if (d.xs_type[j] == XS_CIRCULAR) r.link_kernel[j] = LK_CIRCULAR;
            else if (d.xs_type[j] == XS_RECT_CLOSED) r.link_kernel[j] = LK_RECT_CLOSED;
        }
        r.xs_rcp_yfull[j] = exact_rcp(d.xs_yfull[j]);
        r.cond_rcp_mod_length[j] = exact_rcp(d.cond_mod_length[j]);
        r.link_z1[j] = d.node_invert[a] + d.link_offset1[j];
        r.link_z2[j] = d.node_invert[b] + d.link_offset2[j];
        // link_setOutfallDepth (link.c:743-753) tests node2 first; a later link overrides an
        // earlier one exactly as the reference's loop over all links does
        if (d.node_type[b] == SWB_OUTFALL) r.outfall_link[b] = j;
        else if (d.node_type[a] == SWB_OUTFALL) r.outfall_link[a] = j;
    }
    r.nNonConduit = (int)r.nc_links.size();
    r.outfall_nodes.clear();
    for (int i = 0; i < nN; i++) if (d.node_type[i] == SWB_OUTFALL) r.outfall_nodes.push_back(i);
    // the normal / critical depth an outfall needs (link.c:728-766) is a serial root search on its conduit's new
    // flow: the link phase computes it right after that conduit's update, and such conduits are drawn FIRST, so
    // the search overlaps the rest of the phase instead of stalling the node phase (and, partitioned, every rank)
    r.outfall_slot.assign(nN, -1); r.link_pre_node.assign(nL, -1);
    for (size_t k = 0; k < r.outfall_nodes.size(); k++) {
        const int i = r.outfall_nodes[k], j = r.outfall_link[i];
        if (j >= 0 && (r.link_flags[j] & LF_TRUE_CONDUIT) && r.link_pre_node[j] < 0) {
            r.link_pre_node[j] = i;
            r.outfall_slot[i] = (int)k;
        }
    }
    r.link_order.clear();
    for (int j = 0; j < nL; j++) if (r.link_flags[j] & LF_TRUE_CONDUIT) r.link_order.push_back(j);
    std::stable_sort(r.link_order.begin(), r.link_order.end(),
                     [&](int a, int b) {
                         if (r.link_kernel[a] != r.link_kernel[b]) return r.link_kernel[a] < r.link_kernel[b];
                         const bool pa = r.link_pre_node[a] >= 0, pb = r.link_pre_node[b] >= 0;
                         if (pa != pb) return pa;
                         return d.xs_type[a] < d.xs_type[b];
                     });
    r.pre_links.clear();
    for (size_t k = 0; k < r.link_order.size(); k++) if (r.link_pre_node[r.link_order[k]] >= 0) r.pre_links.push_back((int)k);
    // packed static rows of the true conduits, in ticket order (swb_dynwave.h: LR_*, CfStaged)
    r.link_rows.assign(r.link_order.size() * (size_t)LR_STRIDE, 0.0);
    for (size_t k = 0; k < r.link_order.size(); k++) {
        const int j = r.link_order[k];
        double *row = r.link_rows.data() + k * (size_t)LR_STRIDE;
        row[LR_YFULL] = d.xs_yfull[j]; row[LR_WMAX] = d.xs_wmax[j]; row[LR_YWMAX] = d.xs_ywmax[j];
        row[LR_AFULL] = d.xs_afull[j]; row[LR_RFULL] = d.xs_rfull[j]; row[LR_SFULL] = d.xs_sfull[j];
        row[LR_SMAX] = d.xs_smax[j]; row[LR_YBOT] = d.xs_ybot[j]; row[LR_ABOT] = d.xs_abot[j];
        row[LR_SBOT] = d.xs_sbot[j]; row[LR_RBOT] = d.xs_rbot[j]; row[LR_RCP_YFULL] = r.xs_rcp_yfull[j];
        row[LR_Z1] = r.link_z1[j]; row[LR_Z2] = r.link_z2[j];
        row[LR_OFFSET1] = d.link_offset1[j]; row[LR_OFFSET2] = d.link_offset2[j];
        row[LR_INV1] = d.node_invert[d.link_node1[j]]; row[LR_INV2] = d.node_invert[d.link_node2[j]];
        row[LR_LENGTH] = d.cond_length[j]; row[LR_MOD_LENGTH] = d.cond_mod_length[j];
        row[LR_RCP_MOD_LENGTH] = r.cond_rcp_mod_length[j]; row[LR_ROUGH] = d.cond_rough_factor[j];
        row[LR_BETA] = d.cond_beta[j]; row[LR_QMAX] = d.cond_q_max[j]; row[LR_QLIMIT] = d.link_q_limit[j];
        row[LR_CLOSS_IN] = d.link_closs_in[j]; row[LR_CLOSS_OUT] = d.link_closs_out[j];
        row[LR_CLOSS_AVG] = d.link_closs_avg[j]; row[LR_SEEP] = d.link_seep_rate[j];
        int iv[LRI_INTS];
        iv[LRI_FLAGS] = r.link_flags[j]; iv[LRI_XS_TYPE] = d.xs_type[j]; iv[LRI_BARRELS] = d.cond_barrels[j];
        iv[LRI_HAS_LOSSES] = d.cond_has_losses[j]; iv[LRI_DIRECTION] = d.link_direction[j];
        iv[LRI_CULVERT] = d.xs_culvert[j]; iv[LRI_NODE1] = d.link_node1[j]; iv[LRI_NODE2] = d.link_node2[j];
        static_assert(LR_DOUBLES * sizeof(double) + LRI_INTS * sizeof(int) <= LR_STRIDE * sizeof(double), "row too small");
        memcpy(row + LR_DOUBLES, iv, sizeof(iv));
    }
    {   // the same attributes column-wise (single models)
        const size_t nT = r.link_order.size();
        r.link_cols_d.assign(nT * (size_t)LR_DOUBLES, 0.0);
        r.link_cols_i.assign(nT * (size_t)LRI_INTS, 0);
        for (size_t k = 0; k < nT; k++) {
            const double *row = r.link_rows.data() + k * (size_t)LR_STRIDE;
            for (int f = 0; f < LR_DOUBLES; f++) r.link_cols_d[(size_t)f * nT + k] = row[f];
            int iv[LRI_INTS];
            memcpy(iv, row + LR_DOUBLES, sizeof(iv));
            for (int f = 0; f < LRI_INTS; f++) r.link_cols_i[(size_t)f * nT + k] = iv[f];
        }
    }
    node_order_expensive_first(d.node_type, nN, nN, r.node_order);
    // CSR incidence.  adjq: ascending link index.  adj: true conduits first, then the rest.
    std::vector<std::vector<int>> inc(nN);
    for (int j = 0; j < nL; j++) {
        inc[d.link_node1[j]].push_back((j << 1) | 0);
        inc[d.link_node2[j]].push_back((j << 1) | 1);
    }
    r.adj_start.assign(nN + 1, 0); r.adjq_start.assign(nN + 1, 0);
    r.adj.clear(); r.adjq.clear();
    for (int i = 0; i < nN; i++) {
        r.adjq_start[i] = (int)r.adjq.size();
        r.adj_start[i] = (int)r.adj.size();
        for (int e : inc[i]) r.adjq.push_back(e);
        for (int e : inc[i]) if (r.link_flags[e >> 1] & LF_TRUE_CONDUIT) r.adj.push_back(e);
        for (int e : inc[i]) if (!(r.link_flags[e >> 1] & LF_TRUE_CONDUIT)) r.adj.push_back(e);
    }
    r.adjq_start[nN] = (int)r.adjq.size();
    r.adj_start[nN] = (int)r.adj.size();
    r.adj_packed.resize(r.adj.size());
    for (size_t e = 0; e < r.adj.size(); e++) {
        const int j = r.adj[e] >> 1;
        AdjEntry a;
        a.je = r.adj[e];
        a.barrels = (d.link_type[j] == SWB_CONDUIT) ? d.cond_barrels[j] : 1;
        a.flags = r.link_flags[j];
        a.kind = d.link_type[j] | ((d.link_type[j] == SWB_PUMP ? d.pump_type[j] : 0) << 8);
        r.adj_packed[e] = a;
    }
    static const double tab[] = { SWB_XS_TABLE_DATA };
    r.xs_tables.assign(tab, tab + XT_TOTAL);
    static const double cul[] = { SWB_CULVERT_PARAM_DATA };
    r.culvert_params.assign(cul, cul + 5 * (SWB_MAX_CULVERT_CODE + 1));
    static const double road[] = { SWB_ROAD_TABLE_DATA };
    r.road_tables.assign(road, road + 2 * RT_TOTAL);
    (void)o;
}

inline void fill_net_scalars(Net &n, const swb_network_desc &d, const swb_options &o, const Derived &r)
{
    n.nN = d.n_nodes; n.nL = d.n_links; n.nP = d.n_pollut; n.nCurves = d.n_curves;
    n.nShapeTbl = d.n_shape_tbls; n.shapeTblLen = d.shape_tbl_len;
    n.nTrue = r.nTrue; n.nNonConduit = r.nNonConduit; n.nOutfallLinks = 0;
    n.nOutfallNodes = (int)r.outfall_nodes.size();
    n.lk_count[0] = n.lk_count[1] = n.lk_count[2] = 0;
    for (int j : r.link_order) n.lk_count[r.link_kernel[j]]++;
    n.nPre = (int)r.pre_links.size();
    n.anyLossRate = 0;
    for (int f : r.link_flags) if (f & LF_HAS_LOSSRATE) n.anyLossRate = 1;
    n.opt = o;
    n.crownCutoff = (o.surcharge_method == SWB_SLOT) ? 0.985257 : 0.96;   // dynwave.c:64-65,159
}

// ---- field table ---------------------------------------------------------------------------------
struct FieldInfo { int id; char kind; bool is_u8; size_t offset; };   // offset of the member in State

inline const std::vector<FieldInfo> &field_table()
{
    static std::vector<FieldInfo> t;
    if (t.empty()) {
#define X(T, name, id, kind) t.push_back(FieldInfo{id, #kind[0] == 'N' ? (#kind[1] == 'P' ? 'n' : 'N') \
                                                        : (#kind[1] == 'P' ? 'l' : 'L'), \
                                          sizeof(T) == 1, offsetof(State, name)});
        SWB_STATE_FIELDS(X)
#undef X
    }
    return t;
}
inline const FieldInfo *find_field(int id)
{
    if (id == SWB_COND_Q2) id = SWB_COND_Q1;     // q2 == q1 under dynamic wave: one array
    for (const FieldInfo &f : field_table()) if (f.id == id) return &f;
    return nullptr;
}
// items per member of a field ('N' nodes, 'L' links, 'n'/'l' x pollutants)
inline size_t field_items(const FieldInfo &f, int nN, int nL, int nP)
{
    switch (f.kind) {
      case 'N': return nN;  case 'L': return nL;
      case 'n': return (size_t)nN * nP;  default: return (size_t)nL * nP;
    }
}

} // namespace swb
#endif
