// swb_staged.cuh -- the routing step as a chain of kernels on one stream (wide ensembles).
//
// The persistent cooperative kernel (swb_route_kernel) gives every phase of a step the same register
// budget and the same number of resident warps.  The conduit update wants ~150 registers, the streaming
// phases (state swap, node gather, quality, Courant search) want as many warps in flight as the SM holds:
// fused, the conduit update spills and the streaming phases starve (profiles/ncu_r02a: 2.7 k warp
// instructions per conduit-update, long-scoreboard stalls 14 warps per issue).  With 256 or more lockstep
// members a phase lasts milliseconds, so the launch boundary costs nothing measurable and each phase gets
// a kernel of its own: the SAME phase functions of swb_engine.h that engine_run() chains between grid
// barriers, here with the kernel boundary as the barrier.
//
//   sg_control(-1)  ph_advance, ph_massbal_fold, per-step resets, ordered member list      1 CTA
//   sg_prologue     ph_prologue                                                            streaming
//   per trial k:    sg_links<LK> (one per conduit-function class, tickets)                 persistent CTAs
//                   [sg_presum, sg_regulators]   networks with regulators
//                   sg_nodes (tickets)
//                   sg_control(k)  member list of trial k + 1                              1 CTA
//   sg_epilogue, sg_qual_nodes, sg_qual_links, sg_outflows, sg_stats                       streaming
//   sg_dt_search, sg_dt_arg, sg_control(-2) = ph_nextdt_final + ph_advance_time            streaming, 1 CTA
//
// Trials are launched up to MaxTrials; a trial whose member list is empty returns at once (the list
// length lives in State::ctl).  Every kernel's first thread closes the phase timer of the kernel before
// it, so swb_get_phase_times works as with the persistent kernel.
#ifndef SWB_STAGED_CUH
#define SWB_STAGED_CUH

#ifndef SWB_SG_LINK_BLOCK
#define SWB_SG_LINK_BLOCK 256
#endif
#ifndef SWB_SG_LINK_MINB
#define SWB_SG_LINK_MINB 3          // 80 registers, 24 warps per SM: measured best with the cp.async pipeline
                                    // (profiles/README.md round 2: 128 regs x 16 warps 298 ms, 80 x 24 267, 64 x 32 274)
#endif
#ifndef SWB_SG_NODE_BLOCK
#define SWB_SG_NODE_BLOCK 256
#endif
#ifndef SWB_SG_NODE_MINB
#define SWB_SG_NODE_MINB 4
#endif
#ifndef SWB_SG_STREAM_BLOCK
#define SWB_SG_STREAM_BLOCK 256
#endif
#ifndef SWB_SG_STREAM_MINB
#define SWB_SG_STREAM_MINB 4
#endif
#define SWB_SG_CTL_BLOCK 1024
#define TP_MARK (SWB_N_PHASES - 1)  // phase_ns[TP_MARK]: globaltimer at the start of the running phase

__device__ __forceinline__ void sg_tick(const State &st, int prev)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (prev >= 0) st.phase_ns[prev] += t - st.phase_ns[TP_MARK];
        st.phase_ns[TP_MARK] = t;
    }
}

__device__ __forceinline__ CudaCtx sg_ctx(const State &st, const double *tab)
{
    CudaCtx ctx;
    ctx.alive_list = st.alive;           // global memory: written by sg_control
    ctx.warp_size = 32;
    ctx.warp_lane = threadIdx.x & 31;
    ctx.scan = nullptr;
    ctx.blk_ticket = nullptr;
    ctx.blk_round = 0;
    ctx.tile = nullptr;
    ctx.Tw = nullptr;
    ctx.tid = blockIdx.x * blockDim.x + threadIdx.x;
    ctx.G = gridDim.x * blockDim.x;
    ctx.lane = threadIdx.x;
    ctx.block_size = blockDim.x;
    ctx.T = tab;
    return ctx;
}
__device__ __forceinline__ void sg_tables(double *tab, const Net &net)
{
    for (int i = threadIdx.x; i < XT_TOTAL; i += blockDim.x) tab[i] = net.xs_tables[i];
    __syncthreads();
}

// k == -1: start of a step.  k >= 1: after the node phase of trial k.  k == -2: end of a step.
__global__ void __launch_bounds__(SWB_SG_CTL_BLOCK, 1)
sg_control(const __grid_constant__ Net net, const __grid_constant__ State st, const __grid_constant__ RunArgs args,
           int k, int prev)
{
    __shared__ int s_scan[1 + SWB_SG_CTL_BLOCK / 32];
    sg_tick(st, prev);
    const int M = st.M;
    CudaCtx ctx = sg_ctx(st, nullptr);
    ctx.scan = s_scan;
    if (k == -1) {
        bool anyLeft = true;
        if (args.phases & PH_ADVANCE) {
            for (int m = threadIdx.x; m < M; m += blockDim.x) {
                ph_advance(net, st, args, m);
                if (args.controls.active && !st.done[m]) ph_controls(net, st, args.controls, m, st.dt[m], net.xs_tables);
            }
            __syncthreads();
            anyLeft = false;
            for (int m = threadIdx.x; m < M; m += blockDim.x) anyLeft = anyLeft || !st.done[m];
            anyLeft = ctx.block_or(anyLeft);
        }
        if (threadIdx.x == 0) { st.ctl[CTL_ANY_LEFT] = anyLeft ? 1 : 0; st.ctl[CTL_N_ALIVE] = 0; }
        if (!anyLeft) return;
        const int maxTrials = engine_max_trials(net);
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            if ((args.phases & PH_MASSBAL) && member_active(st, args, m)) ph_massbal_fold(net, st, m, st.dt[m]);
            if (args.phases & PH_DYNWAVE)
                for (int kk = 0; kk < maxTrials; kk++) st.not_conv[kk * M + m] = 0;
            if (args.phases & PH_NEXTDT) ph_nextdt_reset(st, args, m);
        }
        for (int c = threadIdx.x; c < SWB_TICKETS_PER_TRIAL * SWB_MAX_TRIALS_CAP; c += blockDim.x) st.tickets[c] = 0ull;
        if (args.phases & PH_DYNWAVE) {
            const int nAlive = ctx.compact_members(M, [&](int mm) { return member_active(st, args, mm); });
            if (threadIdx.x == 0) st.ctl[CTL_N_ALIVE] = nAlive;
        }
    } else if (k == -2) {
        if (!st.ctl[CTL_ANY_LEFT]) return;
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            if (!member_active(st, args, m)) continue;
            if (args.phases & PH_NEXTDT) ph_nextdt_final(net, st, args, m);
            if (args.phases & PH_ADVANCE) ph_advance_time(st, m, st.dt[m]);
        }
    } else {
        if (!st.ctl[CTL_ANY_LEFT] || st.ctl[CTL_N_ALIVE] == 0) return;
        const int nAlive = ctx.compact_members(M, [&](int mm) { return member_iterates(st, args, mm, k); });
        if (threadIdx.x == 0) st.ctl[CTL_N_ALIVE] = nAlive;
    }
}

__global__ void sg_tick_only(const __grid_constant__ State st, int prev) { sg_tick(st, prev); }

__global__ void __launch_bounds__(SWB_SG_STREAM_BLOCK, SWB_SG_STREAM_MINB)
sg_prologue(const __grid_constant__ Net net, const __grid_constant__ State st, const __grid_constant__ RunArgs args, int prev)
{
    sg_tick(st, prev);
    if (!st.ctl[CTL_ANY_LEFT]) return;
    CudaCtx ctx = sg_ctx(st, nullptr);
    const ThreadMap tm = thread_map(ctx, st.M);
    if (member_active(st, args, tm.m)) ph_prologue(net, st, args, ctx, tm, st.dt[tm.m]);
}

// findLinkFlows pass (i) for the true conduits of one conduit-function class (dynwave.c:387-395)
template <int LK>
__global__ void __launch_bounds__(SWB_SG_LINK_BLOCK, SWB_SG_LINK_MINB)
sg_links(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    __shared__ double tab[XT_TOTAL];
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    sg_tables(tab, net);
    CudaCtx ctx = sg_ctx(st, tab);
    const int M = st.M;
    const int j0 = (LK > 0 ? net.lk_count[0] : 0) + (LK > 1 ? net.lk_count[1] : 0);
    for_tiles(ctx, net.lk_count[LK], nAlive, st.tickets + SWB_TICKETS_PER_TRIAL * k + TK_LINKS + LK, [&](int jj, int mm) {
        const int j = net.link_order[j0 + jj];
        if (k >= 2) {                      // findBypassedLinks of the previous trial (dynwave.c:335-345)
            const bool byp = st.n_converged[SWB_IX(net.link_node1[j], mm, M)] &&
                             st.n_converged[SWB_IX(net.link_node2[j], mm, M)];
            st.l_bypassed[SWB_IX(j, mm, M)] = byp ? 1 : 0;
            if (byp) return;
        }
        if (LK == LK_CIRCULAR) conduit_flow<XS_CIRCULAR>(net, st, j, mm, k, st.dt[mm], tab);
        else if (LK == LK_RECT_CLOSED) conduit_flow<XS_RECT_CLOSED>(net, st, j, mm, k, st.dt[mm], tab);
        else conduit_flow_generic(net, st, j, mm, k, st.dt[mm], tab);
        if (net.link_flags[j] & (LF_N1_OUTFALL | LF_N2_OUTFALL)) outfall_precompute(net, st, j, mm, tab);
    });
}

// The same pass as a software pipeline per warp (Blackwell / Hopper async copies, no register cost):
//   tile i + 3   ticket atomic in flight
//   tile i + 2   link / end-node / member indices being loaded (plain loads, consumed an iteration later)
//   tile i + 1   its seven input rows on their way into the warp's shared-memory stage (cp.async, LDGSTS)
//   tile i       conduit update reading its inputs from the stage (LDS)
// so neither the ticket counter's round trip, nor the index chain link_order -> link_node -> alive, nor the
// DRAM latency of the state rows sits on the critical path of a warp that has only 3 siblings per scheduler.
#ifndef SWB_SG_LINK_PF
#define SWB_SG_LINK_PF 1
#endif
struct PfIds { int j, n1, n2, mm, row; };   // mm < 0: lane has no member in this tile (or the conduit is bypassed);
                                            // row = position of the conduit in link_order (its packed static row)

__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int LK>
__global__ void __launch_bounds__(SWB_SG_LINK_BLOCK, SWB_SG_LINK_MINB)
sg_links_pf(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    __shared__ double tab[XT_TOTAL];
    __shared__ double stage_all[(SWB_SG_LINK_BLOCK / 32) * 2 * CF_FIELDS * 32];
    __shared__ __align__(16) double rows_all[(SWB_SG_LINK_BLOCK / 32) * 2 * LR_STRIDE];
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    sg_tables(tab, net);
    const int M = st.M;
    const int j0 = (LK > 0 ? net.lk_count[0] : 0) + (LK > 1 ? net.lk_count[1] : 0);
    unsigned long long *ticket = st.tickets + SWB_TICKETS_PER_TRIAL * k + TK_LINKS + LK;
    auto update = [&](int j, int mm, double dt, auto &&in) {
        if (LK == LK_CIRCULAR) conduit_flow_in<XS_CIRCULAR>(net, st, j, mm, k, dt, tab, in);
        else if (LK == LK_RECT_CLOSED) conduit_flow_in<XS_RECT_CLOSED>(net, st, j, mm, k, dt, tab, in);
        else conduit_flow_in<-1>(net, st, j, mm, k, dt, tab, in);
    };
    if (nAlive < 32) {                     // few members left: several conduits per tile, plain loads
        CudaCtx ctx = sg_ctx(st, tab);
        for_tiles(ctx, net.lk_count[LK], nAlive, ticket, [&](int jj, int mm) {
            const int j = net.link_order[j0 + jj];
            if (k >= 2) {
                const bool byp = st.n_converged[SWB_IX(net.link_node1[j], mm, M)] &&
                                 st.n_converged[SWB_IX(net.link_node2[j], mm, M)];
                st.l_bypassed[SWB_IX(j, mm, M)] = byp ? 1 : 0;
                if (byp) return;
            }
            conduit_update(net, st, j, mm, k, st.dt[mm], tab);     // (out of line: this path is rare)
            if (net.link_flags[j] & (LF_N1_OUTFALL | LF_N2_OUTFALL)) outfall_precompute(net, st, j, mm, tab);
        });
        return;
    }
    const int lane = threadIdx.x & 31;
    double *my = stage_all + (size_t)(threadIdx.x >> 5) * 2 * CF_FIELDS * 32 + lane;
    double *myRows = rows_all + (size_t)(threadIdx.x >> 5) * 2 * LR_STRIDE;
    const int nChunks = (nAlive + 31) / 32;
    const unsigned total = (unsigned)net.lk_count[LK] * (unsigned)nChunks;     // < 2^31 (checked by the launcher)
    auto issue = [&]() -> unsigned long long { unsigned long long v = 0; if (lane == 0) v = atomicAdd(ticket, 1ull); return v; };
    auto take = [&](unsigned long long v) -> unsigned {
        v = __shfl_sync(0xffffffffu, v, 0);
        return v < (unsigned long long)total ? (unsigned)v : total;
    };
    auto ids = [&](unsigned t) -> PfIds {
        PfIds d;
        const unsigned jj = t / (unsigned)nChunks;
        const int slot = (int)(t - jj * (unsigned)nChunks) * 32 + lane;
        d.row = j0 + (int)jj;
        d.j = net.link_order[d.row];
        d.n1 = net.link_node1[d.j]; d.n2 = net.link_node2[d.j];
        d.mm = slot < nAlive ? st.alive[slot] : -1;
        if (k >= 2 && d.mm >= 0) {         // findBypassedLinks of the previous trial (dynwave.c:335-345)
            const bool byp = st.n_converged[SWB_IX(d.n1, d.mm, M)] && st.n_converged[SWB_IX(d.n2, d.mm, M)];
            st.l_bypassed[SWB_IX(d.j, d.mm, M)] = byp ? 1 : 0;
            if (byp) d.mm = -1;
        }
        return d;
    };
    auto prefetch = [&](const PfIds &d, int stg, bool valid) {
        // the conduit's static row: 18 lanes x 16 bytes (warp-uniform source, every lane of the tile reads it)
        if (valid && lane < LR_STRIDE / 2)
            cp_async16(myRows + stg * LR_STRIDE + lane * 2, net.link_rows + (size_t)d.row * LR_STRIDE + lane * 2);
        if (d.mm >= 0) {
            double *b = my + stg * (CF_FIELDS * 32);
            const size_t ix = SWB_IX(d.j, d.mm, M);
            cp_async8(b + CF_QLAST * 32, &st.c_q1[ix]);
            cp_async8(b + CF_DEPTH1 * 32, &st.n_depth[SWB_IX(d.n1, d.mm, M)]);
            cp_async8(b + CF_DEPTH2 * 32, &st.n_depth[SWB_IX(d.n2, d.mm, M)]);
            cp_async8(b + CF_SETTING * 32, &st.l_setting[ix]);
            cp_async8(b + CF_AOLD * 32, &st.c_a2[ix]);
            cp_async8(b + CF_OLDFLOW * 32, &st.l_old_flow[ix]);
            cp_async8(b + CF_DT * 32, &st.dt[d.mm]);
        }
        cp_async_commit();
    };
    const PfIds none = { 0, 0, 0, -1, 0 };
    unsigned t0 = take(issue());
    unsigned t1 = t0 < total ? take(issue()) : total;
    unsigned t2 = t1 < total ? take(issue()) : total;
    PfIds i0 = t0 < total ? ids(t0) : none, i1 = t1 < total ? ids(t1) : none, i2 = t2 < total ? ids(t2) : none;
    prefetch(i0, 0, t0 < total);
    int stg = 0;
    while (t0 < total) {
        const unsigned long long pend = (t2 < total) ? issue() : (unsigned long long)total;
        __syncwarp();                      // every lane is done reading the static row the next copy overwrites
        prefetch(i1, stg ^ 1, t1 < total); // (an empty group when there is no next tile)
        cp_async_wait<1>();                // everything but the group just committed: tile t0's rows are in
        __syncwarp();                      // ... including the parts of the static row other lanes copied
        if (i0.mm >= 0) {
            const double *b = my + stg * (CF_FIELDS * 32);
            const CfStaged in = { b, myRows + stg * LR_STRIDE };
            update(i0.j, i0.mm, b[CF_DT * 32], in);
            if (in.flags() & (LF_N1_OUTFALL | LF_N2_OUTFALL)) outfall_precompute(net, st, i0.j, i0.mm, tab);
        }
        t0 = t1; i0 = i1; t1 = t2; i1 = i2; stg ^= 1;
        t2 = (t2 < total) ? take(pend) : total;
        i2 = t2 < total ? ids(t2) : none;
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(SWB_SG_NODE_BLOCK, SWB_SG_NODE_MINB)
sg_presum(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    CudaCtx ctx = sg_ctx(st, nullptr);
    for_tiles(ctx, net.nN, nAlive, st.tickets + SWB_TICKETS_PER_TRIAL * k + TK_PRESUM,
              [&](int ii, int mm) { picard_node_presum(net, st, net.node_order[ii], mm); });
}

__global__ void __launch_bounds__(128, 1)
sg_regulators(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    __shared__ double tab[XT_TOTAL];
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    sg_tables(tab, net);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nAlive) {
        const int mm = st.alive[t];
        regulator_pass(net, st, mm, k, st.dt[mm], tab);
    }
}

// findNodeDepths (dynwave.c:593-632)
__global__ void __launch_bounds__(SWB_SG_NODE_BLOCK, SWB_SG_NODE_MINB)
sg_nodes(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    __shared__ double tab[XT_TOTAL];
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    sg_tables(tab, net);
    CudaCtx ctx = sg_ctx(st, tab);
    const int M = st.M;
    for_tiles(ctx, net.nN, nAlive, st.tickets + SWB_TICKETS_PER_TRIAL * k + TK_NODES, [&](int ii, int mm) {
        if (!picard_node(net, st, net.node_order[ii], mm, k, st.dt[mm], tab)) st.not_conv[k * M + mm] = 1;
    });
}

// findNodeDepths as a software pipeline per warp, for networks without regulators (the node sums then come
// straight from the conduits, dynwave.c:528-589):
//   tile i + 3   ticket atomic in flight
//   tile i + 2   node, its incidence range and up to four packed incidence entries (plain loads)
//   tile i + 1   its rows on their way into the warp's shared-memory stage (cp.async): depth, losses, lateral
//                inflow, old depth, old net inflow of the node and flow / surface area / dqdh of each link end
//   tile i       node_init_acc / node_apply_link_end / node_set_depth on the staged values
// A tile the stage does not fit (outfall, more than four link ends, a conduit with a loss rate) runs picard_node
// as before.  Same functions, same operand values, same order of the sums: the same bits.
// MEASURED SLOWER than sg_nodes (profiles/README.md, round 2: node phase 281 -> 464 ms per 50 steps at 4 096
// members, 78 -> 122 at 1 024): the plain kernel already keeps a node's 12-17 loads in flight per thread at 32
// warps per SM, while the stage costs 17 LDGSTS + 17 LDS per thread and caps the SM at 24 warps.  Compiled out.
#ifndef SWB_SG_NODE_PF
#define SWB_SG_NODE_PF 0
#endif
#if SWB_SG_NODE_PF
#define SWB_SG_NODEPF_BLOCK 128
enum { NF_DEPTH = 0, NF_LOSSES, NF_LATFLOW, NF_OLD_DEPTH, NF_OLD_NET_INFLOW, NF_LINK0, NF_FIELDS = NF_LINK0 + 12 };
struct NfIds { int i, mm, deg, plain; };     // plain: the tile runs picard_node (nothing staged)

__global__ void __launch_bounds__(SWB_SG_NODEPF_BLOCK, 6)
sg_nodes_pf(const __grid_constant__ Net net, const __grid_constant__ State st, int k, int prev)
{
    __shared__ double stage_all[(SWB_SG_NODEPF_BLOCK / 32) * 2 * NF_FIELDS * 32];
    __shared__ __align__(16) AdjEntry adj_all[(SWB_SG_NODEPF_BLOCK / 32) * 2 * 4];
    sg_tick(st, prev);
    const int nAlive = st.ctl[CTL_N_ALIVE];
    if (!st.ctl[CTL_ANY_LEFT] || nAlive == 0) return;
    const int M = st.M;
    const double *T = net.xs_tables;                 // (global copy: only an outfall without a precomputed depth reads it)
    unsigned long long *ticket = st.tickets + SWB_TICKETS_PER_TRIAL * k + TK_NODES;
    if (nAlive < 32) {                               // few members left: several nodes per tile, plain loads
        CudaCtx ctx = sg_ctx(st, nullptr);
        ctx.T = T;
        for_tiles(ctx, net.nN, nAlive, ticket, [&](int ii, int mm) {
            if (!picard_node(net, st, net.node_order[ii], mm, k, st.dt[mm], T)) st.not_conv[k * M + mm] = 1;
        });
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *my = stage_all + (size_t)warp * 2 * NF_FIELDS * 32 + lane;
    AdjEntry *myAdj = adj_all + warp * 2 * 4;
    const int nChunks = (nAlive + 31) / 32;
    const unsigned total = (unsigned)net.nN * (unsigned)nChunks;
    auto issue = [&]() -> unsigned long long { unsigned long long v = 0; if (lane == 0) v = atomicAdd(ticket, 1ull); return v; };
    auto take = [&](unsigned long long v) -> unsigned {
        v = __shfl_sync(0xffffffffu, v, 0);
        return v < (unsigned long long)total ? (unsigned)v : total;
    };
    const AdjEntry noAdj = {0, 0, 0, -1};
    struct Fetched { NfIds d; AdjEntry a0, a1, a2, a3; };
    auto ids = [&](unsigned t) -> Fetched {
        Fetched f;
        const unsigned ii = t / (unsigned)nChunks;
        const int slot = (int)(t - ii * (unsigned)nChunks) * 32 + lane;
        const int i = net.node_order[(int)ii];
        const int e0 = net.adj_start[i], e1 = net.adj_start[i + 1];
        f.d.i = i; f.d.deg = e1 - e0;
        f.d.mm = slot < nAlive ? st.alive[slot] : -1;
        f.d.plain = (net.node_type[i] == SWB_OUTFALL || f.d.deg > 4) ? 1 : 0;
        f.a0 = f.d.deg > 0 ? net.adj_packed[e0] : noAdj;
        f.a1 = f.d.deg > 1 ? net.adj_packed[e0 + 1] : noAdj;
        f.a2 = f.d.deg > 2 ? net.adj_packed[e0 + 2] : noAdj;
        f.a3 = f.d.deg > 3 ? net.adj_packed[e0 + 3] : noAdj;
        return f;
    };
    auto lossy = [&](const AdjEntry &a) { return a.kind >= 0 && (a.kind & 0xff) == SWB_CONDUIT && (a.flags & LF_HAS_LOSSRATE); };
    auto end_rows = [&](double *b, int e, const AdjEntry &a, int mm) {
        const size_t ix = SWB_IX(a.je >> 1, mm, M);
        cp_async8(b + (NF_LINK0 + 3 * e + 0) * 32, &st.l_flow[ix]);
        cp_async8(b + (NF_LINK0 + 3 * e + 1) * 32, (a.je & 1) == 0 ? &st.l_surf_area1[ix] : &st.l_surf_area2[ix]);
        cp_async8(b + (NF_LINK0 + 3 * e + 2) * 32, &st.l_dqdh[ix]);
    };
    // returns the ids with `plain` finalised; stores the incidence entries for the compute stage
    auto prefetch = [&](Fetched &f, int stg, bool valid) -> NfIds {
        if (valid) {
            if (lossy(f.a0) || lossy(f.a1) || lossy(f.a2) || lossy(f.a3)) f.d.plain = 1;
            if (!f.d.plain) {
                if (lane == 0) { AdjEntry *a = myAdj + stg * 4; a[0] = f.a0; a[1] = f.a1; a[2] = f.a2; a[3] = f.a3; }
                if (f.d.mm >= 0) {
                    double *b = my + stg * (NF_FIELDS * 32);
                    const size_t ix = SWB_IX(f.d.i, f.d.mm, M);
                    cp_async8(b + NF_DEPTH * 32, &st.n_depth[ix]);
                    cp_async8(b + NF_LOSSES * 32, &st.n_losses[ix]);
                    cp_async8(b + NF_LATFLOW * 32, &st.n_latflow[ix]);
                    cp_async8(b + NF_OLD_DEPTH * 32, &st.n_old_depth[ix]);
                    cp_async8(b + NF_OLD_NET_INFLOW * 32, &st.n_old_net_inflow[ix]);
                    if (f.d.deg > 0) end_rows(b, 0, f.a0, f.d.mm);
                    if (f.d.deg > 1) end_rows(b, 1, f.a1, f.d.mm);
                    if (f.d.deg > 2) end_rows(b, 2, f.a2, f.d.mm);
                    if (f.d.deg > 3) end_rows(b, 3, f.a3, f.d.mm);
                }
            }
        }
        cp_async_commit();
        return f.d;
    };
    const NfIds none = { 0, -1, 0, 1 };
    Fetched fnone; fnone.d = none; fnone.a0 = fnone.a1 = fnone.a2 = fnone.a3 = noAdj;
    unsigned t0 = take(issue());
    unsigned t1 = t0 < total ? take(issue()) : total;
    unsigned t2 = t1 < total ? take(issue()) : total;
    Fetched f0 = t0 < total ? ids(t0) : fnone, f1 = t1 < total ? ids(t1) : fnone, f2 = t2 < total ? ids(t2) : fnone;
    NfIds i0 = prefetch(f0, 0, t0 < total);
    int stg = 0;
    while (t0 < total) {
        const unsigned long long pend = (t2 < total) ? issue() : (unsigned long long)total;
        __syncwarp();                      // every lane is done with the incidence entries the next tile overwrites
        const NfIds i1 = prefetch(f1, stg ^ 1, t1 < total);
        cp_async_wait<1>();                // tile t0's rows are in
        __syncwarp();
        if (i0.mm >= 0) {
            const int i = i0.i, mm = i0.mm;
            if (i0.plain) {
                if (!picard_node(net, st, i, mm, k, st.dt[mm], T)) st.not_conv[k * M + mm] = 1;
            } else {
                const double *b = my + stg * (NF_FIELDS * 32);
                const AdjEntry *a = myAdj + stg * 4;
                // node_init_acc on the staged values (dynwave.c:297-331)
                NodeAcc acc;
                const double depth = b[NF_DEPTH * 32];
                acc.surfArea = net.opt.allow_ponding ? node_ponded_area(net, i, depth) : node_surf_area(net, i, depth);
                acc.inflow = 0.0;
                acc.outflow = b[NF_LOSSES * 32];
                const double lat = b[NF_LATFLOW * 32];
                if (lat >= 0.0) acc.inflow += lat; else acc.outflow -= lat;
                acc.sumdqdh = 0.0;
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (e < i0.deg) {
                        const LinkEndData d = { b[(NF_LINK0 + 3 * e + 0) * 32], 0.0, b[(NF_LINK0 + 3 * e + 1) * 32],
                                                b[(NF_LINK0 + 3 * e + 2) * 32] };
                        node_apply_link_end(d, a[e], acc);
                    }
                const NodeOld old = { b[NF_OLD_DEPTH * 32], b[NF_OLD_NET_INFLOW * 32] };
                if (!node_set_depth(net, st, i, mm, k, st.dt[mm], acc, old)) st.not_conv[k * M + mm] = 1;
            }
        }
        t0 = t1; i0 = i1; f1 = f2; t1 = t2; stg ^= 1;
        t2 = (t2 < total) ? take(pend) : total;
        f2 = t2 < total ? ids(t2) : fnone;
    }
    cp_async_wait<0>();
}
#endif   // SWB_SG_NODE_PF

enum { SG_EPILOGUE = 0, SG_QUAL_NODES, SG_QUAL_LINKS, SG_OUTFLOWS, SG_STATS, SG_DT_SEARCH, SG_DT_ARG };

// the streaming phases after the Picard loop: one template, one instance (and register budget) per phase
template <int WHAT>
__global__ void __launch_bounds__(SWB_SG_STREAM_BLOCK, SWB_SG_STREAM_MINB)
sg_stream(const __grid_constant__ Net net, const __grid_constant__ State st, const __grid_constant__ RunArgs args, int prev)
{
    __shared__ double tab[WHAT == SG_STATS ? XT_TOTAL : 1];
    sg_tick(st, prev);
    if (!st.ctl[CTL_ANY_LEFT]) return;
    if (WHAT == SG_STATS) sg_tables(tab, net);
    CudaCtx ctx = sg_ctx(st, tab);
    const int M = st.M;
    const ThreadMap tm = thread_map(ctx, M);
    const int m = tm.m;
    const bool active = member_active(st, args, m);
    const double dt = st.dt[m];
    const bool withQual = (net.nP > 0) && !net.opt.ignore_quality;
    if (WHAT == SG_EPILOGUE) { if (active) ph_epilogue(net, st, tm); }
    else if (WHAT == SG_QUAL_NODES) { if (active) ph_qual_nodes(net, st, ctx, tm, dt, net.nN); }
    else if (WHAT == SG_QUAL_LINKS) { if (active) ph_qual_links(net, st, ctx, tm, dt); }
    else if (WHAT == SG_OUTFLOWS) { if (active) ph_outflows(net, st, args, ctx, tm, dt, withQual, net.nN); }
    else if (WHAT == SG_STATS) ph_stats(net, st, args, tm, active, dt, withQual, net.nN, tab);
    else if (WHAT == SG_DT_SEARCH) {
        const bool search = active && nextdt_variable(net, args) && st.var_step[m] != 0.0;
        const DtCand c = ph_nextdt_search(net, st, args, ctx, tm, search, net.nN);
        const size_t S = (size_t)st.dt_cand_stride;
        st.dt_cand[ctx.tid] = c.tl; st.dt_cand[S + ctx.tid] = c.tn;
        st.dt_cand[2 * S + ctx.tid] = (double)c.il; st.dt_cand[3 * S + ctx.tid] = (double)c.in;
    } else if (WHAT == SG_DT_ARG) {
        const bool search = active && nextdt_variable(net, args) && st.var_step[m] != 0.0;
        if (search) {
            const size_t S = (size_t)st.dt_cand_stride;
            DtCand c;
            c.tl = st.dt_cand[ctx.tid]; c.tn = st.dt_cand[S + ctx.tid];
            c.il = (int)st.dt_cand[2 * S + ctx.tid]; c.in = (int)st.dt_cand[3 * S + ctx.tid];
            ph_nextdt_arg(st, ctx, m, c);
        }
    }
}

// host layout [m][item][p] <-> device layout [(p, item)][m] (swb_step_host), 32 x 32 tiles through
// shared memory, both sides coalesced; the same arithmetic as CudaCtx::transpose
__global__ void __launch_bounds__(256, 4)
sg_transpose(double *dst, const double *src, int R, int C, int planes, const __grid_constant__ State st, int prev)
{
    __shared__ double tile[32][33];
    sg_tick(st, prev);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tilesC = (C + 31) / 32, total = ((R + 31) / 32) * tilesC;
    const int items = planes > 0 ? C / planes : 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int r0 = (t / tilesC) * 32, c0 = (t % tilesC) * 32;
        for (int r = ty; r < 32; r += nw)
            if (r0 + r < R && c0 + tx < C) tile[r][tx] = src[(size_t)(r0 + r) * C + c0 + tx];
        __syncthreads();
        for (int c = ty; c < 32; c += nw) {
            int cc = c0 + c;
            if (cc < C && r0 + tx < R) {
                if (planes > 0) { int item = cc / planes; cc = (cc - item * planes) * items + item; }
                dst[(size_t)cc * R + r0 + tx] = tile[tx][c];
            }
        }
        __syncthreads();
    }
}

#if SWB_SG_LINK_PF
#define SG_LINKS_FAST sg_links_pf
#else
#define SG_LINKS_FAST sg_links
#endif

namespace swb { namespace backend {

struct StagedInfo {
    bool ready;
    int link_blocks[3], node_blocks, nodepf_blocks, presum_blocks, stream_blocks[8], prologue_blocks, transpose_blocks;
};
static StagedInfo g_staged[SWB_MAX_DEVICES];

template <class K>
static int sg_occupancy(K kernel, int block, int sms)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, 0) != cudaSuccess || n < 1) n = 1;
    return n * sms;
}

static void staged_init(int device)
{
    StagedInfo &S = g_staged[device];
    if (S.ready) return;
    const int sms = g_dev[device].sms;
#if SWB_SG_LINK_PF
    S.link_blocks[0] = sg_occupancy(sg_links_pf<0>, SWB_SG_LINK_BLOCK, sms);
    S.link_blocks[1] = sg_occupancy(sg_links_pf<1>, SWB_SG_LINK_BLOCK, sms);
#else
    S.link_blocks[0] = sg_occupancy(sg_links<0>, SWB_SG_LINK_BLOCK, sms);
    S.link_blocks[1] = sg_occupancy(sg_links<1>, SWB_SG_LINK_BLOCK, sms);
#endif
    S.link_blocks[2] = sg_occupancy(sg_links<2>, SWB_SG_LINK_BLOCK, sms);
    S.node_blocks = sg_occupancy(sg_nodes, SWB_SG_NODE_BLOCK, sms);
#if SWB_SG_NODE_PF
    S.nodepf_blocks = sg_occupancy(sg_nodes_pf, SWB_SG_NODEPF_BLOCK, sms);
#endif
    S.presum_blocks = sg_occupancy(sg_presum, SWB_SG_NODE_BLOCK, sms);
    S.prologue_blocks = sg_occupancy(sg_prologue, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_EPILOGUE] = sg_occupancy(sg_stream<SG_EPILOGUE>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_QUAL_NODES] = sg_occupancy(sg_stream<SG_QUAL_NODES>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_QUAL_LINKS] = sg_occupancy(sg_stream<SG_QUAL_LINKS>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_OUTFLOWS] = sg_occupancy(sg_stream<SG_OUTFLOWS>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_STATS] = sg_occupancy(sg_stream<SG_STATS>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_DT_SEARCH] = sg_occupancy(sg_stream<SG_DT_SEARCH>, SWB_SG_STREAM_BLOCK, sms);
    S.stream_blocks[SG_DT_ARG] = S.stream_blocks[SG_DT_SEARCH];      // the two share their thread mapping
    S.transpose_blocks = sg_occupancy(sg_transpose, 256, sms);
    S.ready = true;
}

// threads of a streaming kernel: as many CTAs as fit the machine, rounded DOWN to a multiple of
// M / gcd(M, block) so that threads % M == 0 (a thread keeps one member), at least one such unit
static int sg_stream_blocks(int maxBlocks, int block, int M, long long items)
{
    int g = M, b = block;
    while (b) { int t = g % b; g = b; b = t; }
    const long long unit = M / g;
    long long work = (items * M + block - 1) / block;
    long long blocks = std::max(1LL, std::min((long long)maxBlocks, work));
    blocks = (blocks / unit) * unit;
    if (blocks < unit) blocks = unit;
    return (int)blocks;
}

// threads the per-thread Courant candidates must be sized for (State::dt_cand_stride)
static int staged_max_threads(int device)
{
    std::string err;
    if (!init(device, err)) return 0;
    return g_dev[device].sms * 2048 + SWB_MAX_MEMBERS;
}

// minimum ensemble width that takes the staged path (0 = never); SWB_STAGED_MIN_M overrides
static int g_staged_min = -1;
static int staged_min_members()
{
    if (g_staged_min < 0) {
        const char *e = getenv("SWB_STAGED_MIN_M");
        g_staged_min = e ? atoi(e) : 256;
        if (g_staged_min < 0) g_staged_min = 0;
    }
    return g_staged_min;
}
static int set_staged_min_members(int v)
{
    const int old = staged_min_members();
    g_staged_min = v < 0 ? 0 : v;
    return old;
}
static bool staged_applies(const State &st, const RunArgs &args)
{
    const int mn = staged_min_members();
    return mn > 0 && st.M >= mn && st.halo.nRanks <= 1 && args.debug == 0 && st.dt_cand != nullptr;
}

static bool launch_staged(const Net &net, const State &st, const RunArgs &args, int device, float *ms,
                          std::string &err, bool wait, int *n_kernels)
{
    if (!init(device, err)) return false;
    staged_init(device);
    const DevInfo &D = g_dev[device];
    const StagedInfo &S = g_staged[device];
    const int M = st.M, nN = net.nN, nL = net.nL, nP = net.nP;
    const int ph = args.phases;
    const bool withQual = (nP > 0) && !net.opt.ignore_quality;
    const int maxTrials = net.opt.max_trials < SWB_MAX_TRIALS_CAP ? net.opt.max_trials : SWB_MAX_TRIALS_CAP;
    const long long items = std::max(nN, nL);
    cudaStream_t q = g_stream;
    int launched = 0;
    auto sblocks = [&](int maxBlocks) { return sg_stream_blocks(maxBlocks, SWB_SG_STREAM_BLOCK, M, items); };
    const int dtBlocks = sblocks(S.stream_blocks[SG_DT_SEARCH]);
    if ((long long)dtBlocks * SWB_SG_STREAM_BLOCK > st.dt_cand_stride) { err = "dt_cand too small"; return false; }
    if ((long long)net.nTrue * ((M + 31) / 32) >= (1LL << 31)) { err = "more than 2^31 conduit tiles per trial"; return false; }
    const long long chunks = (M + 31) / 32;
    auto tblocks = [&](int maxBlocks, long long tiles) {
        return (int)std::max(1LL, std::min((long long)maxBlocks, (tiles * 32 + SWB_SG_LINK_BLOCK - 1) / SWB_SG_LINK_BLOCK));
    };
    persist_window(device, q, net.arena, net.arena_bytes);
    cudaError_t e = cudaSuccess;
    if (wait) { e = cudaEventRecord(D.ev0, q); if (e != cudaSuccess) { err = cuda_err("cudaEventRecord", e); return false; } }
    int prev = -1;
    if (args.stg_lat || args.stg_losses || args.stg_qual) {
        const int tb = S.transpose_blocks;
        if (args.stg_lat) { sg_transpose<<<tb, 256, 0, q>>>(const_cast<double *>(args.host_lat), args.stg_lat, M, nN, 1, st, prev); prev = TP_PROLOGUE; launched++; }
        if (args.stg_losses) { sg_transpose<<<tb, 256, 0, q>>>(const_cast<double *>(args.host_losses), args.stg_losses, M, nN, 1, st, prev); prev = TP_PROLOGUE; launched++; }
        if (args.stg_qual) { sg_transpose<<<tb, 256, 0, q>>>(const_cast<double *>(args.host_qual), args.stg_qual, M, nN * nP, nP, st, prev); prev = TP_PROLOGUE; launched++; }
    }
    const int chunkSteps = 32;
    int *flag = nullptr;
    for (int step = 0; step < args.n_steps; step++) {
        sg_control<<<1, SWB_SG_CTL_BLOCK, 0, q>>>(net, st, args, -1, prev); prev = TP_PROLOGUE; launched++;
        if (ph & (PH_SWAP | PH_INFLOWS | PH_QSWAP | PH_DYNWAVE | PH_HOSTIN)) {
            sg_prologue<<<sblocks(S.prologue_blocks), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); launched++;
        }
        if (ph & PH_DYNWAVE) {
            for (int k = 0; k < maxTrials; k++) {
                if (net.lk_count[0] > 0) { SG_LINKS_FAST<0><<<tblocks(S.link_blocks[0], net.lk_count[0] * chunks), SWB_SG_LINK_BLOCK, 0, q>>>(net, st, k, prev); prev = TP_LINKS; launched++; }
                if (net.lk_count[1] > 0) { SG_LINKS_FAST<1><<<tblocks(S.link_blocks[1], net.lk_count[1] * chunks), SWB_SG_LINK_BLOCK, 0, q>>>(net, st, k, prev); prev = TP_LINKS; launched++; }
                if (net.lk_count[2] > 0) { sg_links<2><<<tblocks(S.link_blocks[2], net.lk_count[2] * chunks), SWB_SG_LINK_BLOCK, 0, q>>>(net, st, k, prev); prev = TP_LINKS; launched++; }
                if (net.nNonConduit > 0) {
                    sg_presum<<<tblocks(S.presum_blocks, nN * chunks), SWB_SG_NODE_BLOCK, 0, q>>>(net, st, k, prev); prev = TP_REGULATORS; launched++;
                    sg_regulators<<<(M + 127) / 128, 128, 0, q>>>(net, st, k, prev); launched++;
                }
#if SWB_SG_NODE_PF
                if (net.nNonConduit == 0)
                    sg_nodes_pf<<<(int)std::max(1LL, std::min((long long)S.nodepf_blocks, (nN * chunks * 32 + SWB_SG_NODEPF_BLOCK - 1) / SWB_SG_NODEPF_BLOCK)), SWB_SG_NODEPF_BLOCK, 0, q>>>(net, st, k, prev);
                else
#endif
                    sg_nodes<<<tblocks(S.node_blocks, nN * chunks), SWB_SG_NODE_BLOCK, 0, q>>>(net, st, k, prev);
                prev = TP_NODES; launched++;
                if (k + 1 >= maxTrials) break;
                if (k >= 1) { sg_control<<<1, SWB_SG_CTL_BLOCK, 0, q>>>(net, st, args, k, prev); prev = TP_CONTROL; launched++; }
            }
            sg_stream<SG_EPILOGUE><<<sblocks(S.stream_blocks[SG_EPILOGUE]), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_EPILOGUE; launched++;
        }
        if (withQual && (ph & PH_QUALITY)) {
            sg_stream<SG_QUAL_NODES><<<sblocks(S.stream_blocks[SG_QUAL_NODES]), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_QUAL_NODES; launched++;
            sg_stream<SG_QUAL_LINKS><<<sblocks(S.stream_blocks[SG_QUAL_LINKS]), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_QUAL_LINKS; launched++;
        }
        if (ph & PH_MASSBAL) { sg_stream<SG_OUTFLOWS><<<sblocks(S.stream_blocks[SG_OUTFLOWS]), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_QUAL_LINKS; launched++; }
        if ((ph & PH_STATS) && st.stat_node) { sg_stream<SG_STATS><<<sblocks(S.stream_blocks[SG_STATS]), SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_QUAL_LINKS; launched++; }
        if (ph & PH_NEXTDT) {
            sg_stream<SG_DT_SEARCH><<<dtBlocks, SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); prev = TP_NEXTDT; launched++;
            sg_stream<SG_DT_ARG><<<dtBlocks, SWB_SG_STREAM_BLOCK, 0, q>>>(net, st, args, prev); launched++;
        }
        if (ph & (PH_NEXTDT | PH_ADVANCE)) { sg_control<<<1, SWB_SG_CTL_BLOCK, 0, q>>>(net, st, args, -2, prev); prev = TP_NEXTDT; launched++; }
        // a long run to t_end: look at the "some member still running" word every few steps
        if (wait && args.n_steps > chunkSteps && (step + 1) % chunkSteps == 0 && step + 1 < args.n_steps) {
            if (!flag && cudaMallocHost((void **)&flag, sizeof(int)) != cudaSuccess) { err = "cudaMallocHost"; return false; }
            // the word belongs to the step just queued: a step that found every member at t_end cleared it
            cudaMemcpyAsync(flag, st.ctl + CTL_ANY_LEFT, sizeof(int), cudaMemcpyDeviceToHost, q);
            e = cudaStreamSynchronize(q);
            if (e != cudaSuccess) { err = cuda_err("staged routing step", e); cudaFreeHost(flag); return false; }
            if (*flag == 0) break;
        }
    }
    if (flag) cudaFreeHost(flag);
    if (args.stg_depth || args.stg_flow) {
        const int tb = S.transpose_blocks;
        if (args.stg_depth) { sg_transpose<<<tb, 256, 0, q>>>(args.stg_depth, st.n_depth, nN, M, 0, st, prev); prev = TP_NEXTDT; launched++; }
        if (args.stg_flow) { sg_transpose<<<tb, 256, 0, q>>>(args.stg_flow, st.l_flow, nL, M, 0, st, prev); prev = TP_NEXTDT; launched++; }
    }
    sg_tick_only<<<1, 1, 0, q>>>(st, prev); launched++;
    if (n_kernels) *n_kernels = launched;
    e = cudaGetLastError();
    if (e != cudaSuccess) { err = cuda_err("staged kernel launch", e); return false; }
    *ms = 0.f;
    if (!wait) return true;
    e = cudaEventRecord(D.ev1, q);
    if (e != cudaSuccess) { err = cuda_err("cudaEventRecord", e); return false; }
    e = cudaEventSynchronize(D.ev1);
    if (e != cudaSuccess) { err = cuda_err("staged routing step", e); return false; }
    cudaEventElapsedTime(ms, D.ev0, D.ev1);
    return true;
}

// one entry for both execution forms: wide unpartitioned ensembles run staged, everything else persistent
static bool launch(const Net &net, const State &st, const RunArgs &args, int device, float *ms, std::string &err,
                   bool wait = true, int *n_kernels = nullptr)
{
    if (staged_applies(st, args)) return launch_staged(net, st, args, device, ms, err, wait, n_kernels);
    if (n_kernels) *n_kernels = 1;
    return launch_persistent(net, st, args, device, ms, err, wait);
}

} }
#endif
