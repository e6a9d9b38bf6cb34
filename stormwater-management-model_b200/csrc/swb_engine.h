// swb_engine.h -- K9: the persistent routing-step driver.
//
// ONE cooperative launch runs `n_steps` complete routing steps for all M lockstep members: step
// prologue, the Picard loop of dynwave_execute (dynwave.c:224-262) with its data-dependent trip
// count, capacity flags, quality routing and the Courant search for the next step.  Phases are
// separated by grid-wide barriers; there is no host round trip per iteration or per step.
//
// Work mapping: the launch has G threads with G % M == 0, so thread t always works on member
// m = t % M and walks objects  t / M, t / M + G / M, ...  A warp is 32 consecutive members of one
// object (see swb_state.h).  Convergence is tracked per member: not_conv[k][m] is raised by any
// node of member m that misses the head tolerance in trial k; a member leaves the loop after
// trial k >= 1 with not_conv[k][m] == 0, exactly like `if (Steps > 1 && converged) break`.
//
// The body is a template over an execution context so that tests/emul can run the very same
// control flow (including every barrier) on host threads.  Ctx provides:
//     int tid, G, lane, block_size;   const double *T (shape tables);
//     void grid_sync();  bool block_or(bool);  unsigned long long now_ns();
//     int *alive_list;  int compact_members(int M, Pred alive);   // ordered list, same in every CTA
//     int warp_size, warp_lane;  unsigned long long next_ticket(unsigned long long *);  // warp-uniform
//     void atomic_min_u64(unsigned long long*, unsigned long long);
//     void atomic_add_f64(double*, double);  void atomic_min_i32(int*, int);
//     unsigned long long warp_min_u64(unsigned long long);  double warp_sum_f64(double);   // full-warp reductions
#ifndef SWB_ENGINE_H
#define SWB_ENGINE_H

#include "swb_dynwave.h"
#include "swb_qual.h"
#include "swb_regulator.h"
#include "swb_controls.h"
#include "swb_stats.h"

#ifdef __CUDACC__
#define SWB_ENGINE __device__
#else
#define SWB_ENGINE
#endif

// build switches of the node phase (measured variants, profiles/README.md)
#ifndef SWB_NODE_EARLY
#define SWB_NODE_EARLY 0    // A/B switch: old depth / net inflow loaded before the gather
#endif

namespace swb {

enum {                         // RunArgs.phases
    PH_SWAP    = 1,            // old <- new, node_initFlows, overflow reset (routing.c:399-409)
    PH_INFLOWS = 2,            // evaluate lateral inflows + quality preload on the device
    PH_DYNWAVE = 4,            // dynwave_execute
    PH_QUALITY = 8,            // qualrout_execute
    PH_NEXTDT  = 16,           // dynwave_getRoutingStep -> var_step
    PH_ADVANCE = 32,           // ensemble clock: pick dt from var_step / t_end, advance sim_time
    PH_QSWAP   = 64,           // quality old <- new, new = 0 (routing.c:312-336)
    PH_HOSTIN  = 128,          // lateral inflows / losses / quality loads from host-fed staging
    PH_MASSBAL = 256,          // routing totals: removeSystemOutflows + massbal_updateRoutingTotals
    PH_STATS   = 512           // per-object statistics (swb_stats.h), only when the planes are allocated
};

struct Inflows {               // device image of swb_inflow_desc
    int n;                     // inflow nodes
    const int    *node, *ts_start;
    const double *ts_t, *ts_q, *sfactor, *baseline, *concen;
    const double *member_scale, *member_shift;
    const int    *node_slot;   // per node: index into the inflow list or -1
    double start_day, start_secs;
    // optional parts (null / 0 = absent); `general` is set when any of them is present: the step
    // prologue then books the external quality mass itself (ph_outflows only does so in the plain case)
    int general;
    const double *cfactor;     // per inflow node
    const int    *base_pattern;
    int nPatterns;
    const int    *pat_type;    // MONTHLY 0, DAILY 1, HOURLY 2, WEEKEND 3
    const double *pat_factor;  // [nPatterns][24]
    const int    *node_q_start;                      // [nN + 1] CSR: a node's pollutant inflow records
    const int    *q_pollut, *q_type, *q_series, *q_pattern;
    const double *q_cfactor, *q_sfactor, *q_baseline;
    const int    *series_start;
    const double *series_t, *series_v;
    const int    *node_dwf_start;                    // [nN + 1] CSR: a node's dry-weather records
    const int    *dwf_param;
    const double *dwf_avg;
    const int    *dwf_patterns;                      // [record][4]
    const double *pollut_dwf_concen;                 // [nP]
    // routing interface file records (null / 0 = none)
    int nIfaceNodes, nIfaceRec;
    const int    *if_slot;                           // [nN] file node of a project node or -1
    const double *if_date, *if_val;                  // [nIfaceRec], [record][file node][1 + nP]
};

enum { DBG_SKIP_LINKS = 1, DBG_SKIP_NODES = 2 };   // RunArgs.debug (profiling aid: isolate one Picard phase)

struct RunArgs {
    int    phases;
    int    n_steps;
    int    debug;
    double t_end;              // PH_ADVANCE: members stop at this simulated time (s)
    double fixed_step;         // RouteStep for PH_NEXTDT / PH_ADVANCE
    Inflows inflows;
    Controls controls;         // control rules, pump depths, timed outfall stages (PH_ADVANCE only)
    const double *host_lat, *host_losses, *host_qual;   // PH_HOSTIN images, device layout
    // host-layout landing zones of swb_step_host ([member][item(,p)]): when set, the kernel itself
    // transposes them into the images above before the first step, and the step's depths / flows
    // into stg_depth / stg_flow after the last one, so the copy engines are the only other
    // consumers of the stream (no transpose launches queued behind another block's kernel)
    const double *stg_lat, *stg_losses, *stg_qual;
    double *stg_depth, *stg_flow;
};

// Object loop of one thread: round k covers objects [k*stride, (k+1)*stride) and the thread takes
// slot (first + k) % stride of it.  The rotation matters: with a fixed slot a periodic pattern in
// the object order (e.g. alternating circular / rectangular conduits with an even stride) would
// hand all expensive objects to the same warps and every grid barrier would wait for them.
#define SWB_FOR_ITEMS(idx, n) \
    for (int _base = 0, _r = first, idx = first; _base < (n); \
         _base += stride, _r = (_r + 1 == stride ? 0 : _r + 1), idx = _base + _r) \
        if (idx < (n))

#define SWB_TICK(phase) do { if (ctx.tid == 0) { unsigned long long _t = ctx.now_ns(); \
        st.phase_ns[phase] += _t - tmark; tmark = _t; } } while (0)

#define SWB_FLOW_TOL 0.00001   // consts.h: FLOW_TOL, routing.c:455

// hydrograph value at time t (s) for inflow slot k: linear between breakpoints, 0 outside
// (table_tseriesLookup with extend = FALSE, table.c:745-806)
SWB_HD double inflow_series(const Inflows &f, int k, double t) { return series_lookup(f.ts_start, f.ts_t, f.ts_q, k, t); }

// getPatternFactor (inflow.c:456-486)
SWB_HD double pattern_factor(const Inflows &f, int p, const DateParts &d)
{
    const double *fac = f.pat_factor + (size_t)p * 24;
    switch (f.pat_type[p]) {
      case 0: if (d.month >= 0 && d.month < 12) return fac[d.month]; break;
      case 1: if (d.day >= 0 && d.day < 7) return fac[d.day]; break;
      case 2: if (d.hour >= 0 && d.hour < 24) return fac[d.hour]; break;
      case 3: if (d.day == 0 || d.day == 6) { if (d.hour >= 0 && d.hour < 24) return fac[d.hour]; } break;
    }
    return 1.0;
}
// inflow_getDwfInflow (inflow.c:361-392)
SWB_HD double dwf_value(const Inflows &f, int r, const DateParts &d)
{
    const int *pat = f.dwf_patterns + (size_t)r * 4;
    double fac = 1.0;
    if (pat[0] >= 0) fac *= pattern_factor(f, pat[0], d);
    if (pat[1] >= 0) fac *= pattern_factor(f, pat[1], d);
    const int p1 = pat[2], p2 = pat[3];
    if (p2 >= 0) {
        if (d.day == 0 || d.day == 6) fac *= pattern_factor(f, p2, d);
        else if (p1 >= 0) fac *= pattern_factor(f, p1, d);
    }
    else if (p1 >= 0) fac *= pattern_factor(f, p1, d);
    return fac * f.dwf_avg[r];
}

// iface_getNumIfaceNodes (iface.c:187-215) without the file cursor: the bracket of `date` in the record dates and
// the fraction between its ends.  rec = index of the NEW record, or -1 when the file has no data for the date.
struct IfaceAt { int rec; double frac; };
SWB_HD IfaceAt iface_bracket(const Inflows &f, double date)
{
    IfaceAt a = { -1, 0.0 };
    const int R = f.nIfaceRec;
    if (R <= 0 || f.if_date[0] > date) return a;                 // file begins after the current date
    int lo = 0, hi = R;                                          // first record with date >= current date
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (f.if_date[mid] < date) lo = mid + 1; else hi = mid; }
    if (lo >= R) return a;                                       // past the last record (NO_DATE)
    a.rec = lo;
    const double dNew = f.if_date[lo], dOld = f.if_date[lo > 0 ? lo - 1 : 0];
    double frac = (date - dOld) / (dNew - dOld);                 // (0 / 0 at the very first record: NaN -> 1 below)
    frac = SWB_MAX(0.0, frac);
    frac = SWB_MIN(frac, 1.0);
    a.frac = frac;
    return a;
}
// (1 - f) * old + f * new of column c of file node k (iface.c:233-275); before the first record "old" is zero
SWB_HD double iface_value(const Inflows &f, const IfaceAt &a, int k, int c, int nP)
{
    const size_t w = (size_t)(1 + nP), row = (size_t)f.nIfaceNodes * w;
    const double v2 = f.if_val[(size_t)a.rec * row + (size_t)k * w + c];
    const double v1 = a.rec > 0 ? f.if_val[(size_t)(a.rec - 1) * row + (size_t)k * w + c] : 0.0;
    return (1.0 - a.frac) * v1 + a.frac * v2;
}

// ---- per-(object, member) bodies of the Picard phases --------------------------------------------
SWB_FI void picard_link(const Net &net, const State &st, int j, int m, int k, double dt, const double *T)
{
    if (!(net.link_flags[j] & LF_TRUE_CONDUIT)) return;
    if (k >= 2) {                      // findBypassedLinks of the previous trial (dynwave.c:335-345)
        const int M = st.M;
        bool byp = st.n_converged[SWB_IX(net.link_node1[j], m, M)] &&
                   st.n_converged[SWB_IX(net.link_node2[j], m, M)];
        st.l_bypassed[SWB_IX(j, m, M)] = byp ? 1 : 0;
        if (byp) return;
    }
    conduit_update(net, st, j, m, k, dt, T);
    if (net.link_flags[j] & (LF_N1_OUTFALL | LF_N2_OUTFALL)) outfall_precompute(net, st, j, m, T);
}
// the same for a single model (M == 1), jj = position of the conduit in link_order: every static attribute,
// the two end nodes included, comes from the column-wise copy in that order
SWB_FI void picard_link_single(const Net &net, const State &st, int jj, int k, double dt, const double *T, bool wantPre)
{
    const int j = net.link_order[jj];
    // a conduit that serves an outfall runs on a thread of its own at the start of the phase (engine_run): its
    // update plus the outfall's normal / critical depth solve is the longest dependent chain of the phase
    const int lflags = net.link_cols_i[(size_t)LRI_FLAGS * net.nTrue + jj];
    const bool isPre = (lflags & (LF_N1_OUTFALL | LF_N2_OUTFALL)) && net.link_pre_node[j] >= 0;
    if (isPre != wantPre) return;
    const int n1 = net.link_cols_i[(size_t)LRI_NODE1 * net.nTrue + jj], n2 = net.link_cols_i[(size_t)LRI_NODE2 * net.nTrue + jj];
    if (k >= 2) {
        const bool byp = st.n_converged[n1] && st.n_converged[n2];
        st.l_bypassed[j] = byp ? 1 : 0;
        if (byp) return;
    }
    conduit_update_cols(net, st, j, jj, n1, n2, 0, k, dt, T);
    if (isPre) outfall_precompute(net, st, j, 0, T);
}
// node sums over the true conduits only, stored for the ordered regulator pass
SWB_FI void picard_node_presum(const Net &net, const State &st, int i, int m)
{
    NodeAcc acc = node_init_acc(net, st, i, m);
    int e1 = net.adj_start[i];
    while (e1 < net.adj_start[i + 1] && (net.adj_packed[e1].flags & LF_TRUE_CONDUIT)) e1++;
    node_gather(net, st, net.adj_start[i], e1, m, acc);
    size_t ix = SWB_IX(i, m, st.M);
    st.n_inflow[ix] = acc.inflow; st.n_outflow[ix] = acc.outflow;
    st.n_new_surf_area[ix] = acc.surfArea; st.n_sumdqdh[ix] = acc.sumdqdh;
}
// gather + outfall boundary / node depth update; returns the node's converged flag
SWB_FI bool picard_node(const Net &net, const State &st, int i, int m, int k, double dt, const double *T)
{
    NodeAcc acc;
    size_t ix = SWB_IX(i, m, st.M);
#if SWB_NODE_EARLY
    // old depth / old net inflow requested together with the node's first loads, not after the
    // link gather (nothing in between stores, but the compiler does not hoist loads over the loop)
    const NodeOld old = node_load_old(st, i, m);
#endif
    if (net.nNonConduit > 0) {
        acc.inflow = st.n_inflow[ix]; acc.outflow = st.n_outflow[ix];
        acc.surfArea = st.n_new_surf_area[ix]; acc.sumdqdh = st.n_sumdqdh[ix];
    } else {
        acc = node_init_acc(net, st, i, m);
        node_gather(net, st, net.adj_start[i], net.adj_start[i + 1], m, acc);
    }
    if (net.node_type[i] == SWB_OUTFALL) {
        st.n_inflow[ix] = acc.inflow; st.n_outflow[ix] = acc.outflow;
        outfall_depth(net, st, i, m, T);
        return true;
    }
#if !SWB_NODE_EARLY
    const NodeOld old = node_load_old(st, i, m);
#endif
    return node_set_depth(net, st, i, m, k, dt, acc, old);
}

// Dynamic tile loop: tile t = (object t / nChunks, chunk t % nChunks of the member list); a warp
// draws tickets until the phase is exhausted.  ctx.warp_size is 32 on the device, 1 in the host
// emulation.
struct NoPrefetch { SWB_ENGINE void operator()(int, int) const {} };

// Tickets run two tiles ahead of the work: while a warp computes tile t it already knows tile t+1
// (whose first cache lines it has asked for through `pre`) and has the atomic for tile t+2 in
// flight, so neither the ~1 us round trip of the ticket counter nor the first DRAM round trip of a
// tile sits on the warp's critical path.
#ifndef SWB_TICKETS_AHEAD
#define SWB_TICKETS_AHEAD 0
#endif
#ifndef SWB_TILE_MODE
#define SWB_TILE_MODE 0      // 0: one ticket per warp and tile; 1: static striding; 2: one ticket per CTA
#endif
#ifndef SWB_PREFETCH
#define SWB_PREFETCH 0
#endif
#ifndef SWB_TICKET_BATCH
#define SWB_TICKET_BATCH 1
#endif
#ifndef SWB_PREFETCH_CUR
#define SWB_PREFETCH_CUR 0
#endif
template <class Ctx, class Body, class Pre = NoPrefetch>
SWB_ENGINE inline void for_tiles(Ctx &ctx, int nItems, int nAlive, unsigned long long *ticket, Body body,
                                 Pre pre = Pre())
{
    const int W = ctx.warp_size;
    const bool wide = (nAlive >= W);
    // wide ensembles: a tile is one object x W consecutive list entries.  Few members (a single
    // model has one): a tile is W / nAlive objects x all list entries, so a lone network still
    // fills every lane with a different object.
    const int nChunks = wide ? (nAlive + W - 1) / W : 1;
    const int perTile = wide ? 1 : W / nAlive;
    const long long total = wide ? (long long)nItems * nChunks : ((long long)nItems + perTile - 1) / perTile;
    const int sub = wide ? 0 : ctx.warp_lane / nAlive;
    const int lslot = wide ? ctx.warp_lane : ctx.warp_lane - sub * nAlive;
    auto run = [&](long long t, auto &&fn) {
        if (wide) {
            int item = (int)(t / nChunks);
            int slot = (int)(t - (long long)item * nChunks) * W + lslot;
            if (slot < nAlive) fn(item, ctx.alive_list[slot]);
        } else {
            long long item = t * perTile + sub;
            if (sub < perTile && item < nItems) fn((int)item, ctx.alive_list[lslot]);
        }
    };
#if defined(__CUDACC__) && SWB_TILE_MODE == 1
    // A/B: static striding, no counter at all (tile = global warp id + round * warps in the grid)
    {
        const long long nW = (long long)ctx.G / W, w0 = (long long)ctx.tid / W;
        for (long long t = w0; t < total; t += nW) run(t, body);
    }
#elif defined(__CUDACC__) && SWB_TILE_MODE == 2
    // A/B: one atomic per CTA hands every warp of the block its own tile (neighbouring member chunks of
    // one object): 16x fewer round trips to the counter, and the block's warps run the same object
    for (;;) {
        const long long t0 = (long long)ctx.next_ticket_block(ticket);
        if (t0 >= total) break;
        const long long t = t0 + ctx.lane / W;
        if (t < total) run(t, body);
    }
#elif SWB_TICKETS_AHEAD == 0
    // One atomic hands out SWB_TICKET_BATCH consecutive tiles (same object, neighbouring member
    // chunks): fewer round trips to the ticket counter, and inside a batch the next tile is known
    // without an atomic, so its first cache lines can be requested while the current one computes.
    for (;;) {
        long long t = (long long)ctx.next_ticket(ticket) * SWB_TICKET_BATCH;
        if (t >= total) break;
#if SWB_PREFETCH_CUR
        run(t, pre);
#endif
#pragma unroll 1
        for (int b = 0; b < SWB_TICKET_BATCH && t + b < total; b++) {
#if SWB_PREFETCH_CUR
            if (b + 1 < SWB_TICKET_BATCH && t + b + 1 < total) run(t + b + 1, pre);
#endif
            run(t + b, body);
        }
    }
#elif SWB_TICKETS_AHEAD == 1
    long long t = (long long)ctx.next_ticket(ticket);
    while (t < total) {
        long long tn = (long long)ctx.next_ticket(ticket);
#if SWB_PREFETCH
        if (tn < total) run(tn, pre);
#endif
        run(t, body);
        t = tn;
    }
#else
    long long t = (long long)ctx.next_ticket(ticket);
    long long tn = (t < total) ? (long long)ctx.next_ticket(ticket) : total;
    while (t < total) {
        unsigned long long pending = (tn < total) ? ctx.ticket_issue(ticket) : 0ull;
#if SWB_PREFETCH
        if (tn < total) run(tn, pre);
#endif
        run(t, body);
        t = tn;
        tn = (tn < total) ? (long long)ctx.ticket_take(pending) : total;
    }
#endif
}

template <class Ctx>
SWB_ENGINE inline void qual_acc_flush(Ctx &ctx, const State &st, int p, int m, int M, double dt, const QualAcc &a)
{
    double r = a.reacted * dt, sp = a.seepage * dt, f = a.finalStorage;
    if (M == 1) {                      // one member: shuffle-sum the warp, one atomic per warp
        r = ctx.warp_sum_f64(r); sp = ctx.warp_sum_f64(sp); f = ctx.warp_sum_f64(f);
        if (ctx.warp_lane != 0) return;
    }
    if (r != 0.0) ctx.atomic_add_f64(&st.mb_reacted[p * M + m], r);
    if (sp != 0.0) ctx.atomic_add_f64(&st.mb_seepage[p * M + m], sp);
    if (f != 0.0) ctx.atomic_add_f64(&st.mb_final_storage[p * M + m], f);
    // the same three as step RATES for the routing totals (StepQualTotals, massbal.c:517-555)
    double *rate = st.mb_rate + (size_t)(MB_FLOW_TERMS + p * MB_QUAL_TERMS) * M + m;
    if (r != 0.0) ctx.atomic_add_f64(rate + (size_t)MBQ_REACTED * M, r / dt);
    if (sp != 0.0) ctx.atomic_add_f64(rate + (size_t)MBQ_SEEP * M, sp / dt);
    if (f != 0.0) ctx.atomic_add_f64(rate + (size_t)MBQ_FINAL * M, f);
}

// ---- partitioned network: one exchange = push border values + reduction operands into every
// peer's window, flag, wait for every peer's flag, pull.  Two grid barriers (one when the producing
// phase pushes its values itself, as the node phase does); rank-to-rank traffic
// is a few KB of direct stores, so the cost is latency (NVLink round trip + barriers), not bandwidth.
//   get(s, w)      value w of send entry s                      (any thread)
//   put(r, w, v)   store value w of receive slot r              (any thread)
//   red_in[]       this rank's reduction operands               (same value in every thread)
//   fold(rows)     rows[p * HALO_RED + q] = operand q of rank p (thread 0, after the wait)
template <class Ctx, class Get, class Put, class Fold>
SWB_ENGINE inline void halo_exchange(Ctx &ctx, const Halo &H, unsigned long long &epoch, int nvals,
                                     const unsigned long long *red_in, Get get, Put put, Fold fold,
                                     bool prepushed = false)
{
    epoch++;
    const int par = (int)(epoch & 1ull);
    // prepushed: the producing phase has already stored (and fenced) this epoch's border values and
    // its closing grid barrier has passed -- no push loop and no extra barrier here
    if (!prepushed) {
        for (int e = ctx.tid; e < H.nSend * nvals; e += ctx.G) {
            const int s = e / nvals, w = e - s * nvals;
            H.peer_stage[H.send_rank[s]][((size_t)H.send_slot[s] * 2 + par) * H.W + w] = get(s, w);
        }
        if (ctx.tid < H.nSend * nvals) ctx.fence_system();   // only threads that stored remotely
        ctx.grid_sync();                                     // every push of this rank has been fenced
    }
    // signal: thread p of the grid serves peer p (reduction operands, then the released epoch flag),
    // so the peers are written in parallel, not one NVLink round trip after the other
    for (int p = ctx.tid; p < H.nRanks; p += ctx.G) {
        if (p == H.rank) continue;
        for (int q = 0; q < HALO_RED; q++)
            H.peer_red[p][((size_t)par * SWB_MAX_RANKS + H.rank) * HALO_RED + q] = red_in[q];
        ctx.store_release_sys(H.peer_ctrl[p] + H.rank, epoch);
    }
    if (ctx.tid == 0)                                    // own row, folded by this same thread below
        for (int q = 0; q < HALO_RED; q++)
            H.red[((size_t)par * SWB_MAX_RANKS + H.rank) * HALO_RED + q] = red_in[q];
    // wait: every CTA polls on its own (thread p watches peer p, the block follows): no second grid
    // barrier between the flags and the pull
    const unsigned long long t0 = (ctx.tid == 0) ? ctx.now_ns() : 0ull;
    for (int p = ctx.lane; p < H.nRanks; p += ctx.block_size) {
        if (p == H.rank) continue;
        const unsigned long long tp = ctx.now_ns();
        while (ctx.load_acquire_sys(H.ctrl + p) < epoch) {
            if (ctx.load_acquire_sys(H.ctrl + HALO_ERR) != 0ull) break;
            if (ctx.now_ns() - tp > H.timeout_ns) {           // never hang the device on a lost peer
                for (int r = 0; r < H.nRanks; r++) ctx.store_release_sys(H.peer_ctrl[r] + HALO_ERR, 1ull);
                break;
            }
        }
    }
    ctx.block_sync();
    if (ctx.tid == 0) {
        H.wait_ns[0] += ctx.now_ns() - t0;              // time spent waiting for the slowest peer
        fold(H.red + (size_t)par * SWB_MAX_RANKS * HALO_RED);
    }
    for (int e = ctx.tid; e < H.nRecv * nvals; e += ctx.G) {
        const int r = e / nvals, w = e - r * nvals;
        put(r, w, ctx.load_sys_f64(&H.stage[((size_t)r * 2 + par) * H.W + w]));
    }
    ctx.grid_sync();
}

// ticket counters of one Picard trial (State::tickets[SWB_TICKETS_PER_TRIAL * k + ...]): the persistent
// kernel draws every conduit tile from TK_LINKS; the staged link kernels draw per conduit-function class
// (TK_LINKS + LK_*)
enum { TK_LINKS = 0, TK_PRESUM = 3, TK_NODES = 4 };

// member mm goes on to trial k + 1: still running and not converged in any trial 1..k (dynwave.c:249-251)
SWB_ENGINE inline bool member_iterates(const State &st, const RunArgs &args, int mm, int k)
{
    bool a = !((args.phases & PH_ADVANCE) && st.done[mm]);
    for (int kk = 1; kk <= k && a; kk++) a = (st.not_conv[kk * st.M + mm] != 0);
    return a;
}

// ---- the phases of one routing step ---------------------------------------------------------------
// Every function below is the work BETWEEN two grid-wide synchronisation points.  engine_run() chains
// them with grid barriers inside one persistent cooperative kernel (narrow ensembles, single models,
// partitioned networks, the host emulation); swb_staged.cuh launches each of them as a kernel of its
// own with its own register budget (wide ensembles), where the kernel boundary is the barrier.
//
// Streaming phases use a fixed mapping: thread t works on member t % M and walks objects
// t / M + k * (G / M) (see SWB_FOR_ITEMS).
struct ThreadMap { int m, first, stride; bool owner; };
template <class Ctx>
SWB_ENGINE inline ThreadMap thread_map(const Ctx &ctx, int M)
{
    ThreadMap t;
    t.m = ctx.tid % M; t.first = ctx.tid / M; t.stride = ctx.G / M;
    t.owner = (t.first == 0);            // the one thread that owns member m's scalars
    return t;
}
SWB_ENGINE inline int engine_max_trials(const Net &net)
{
    return net.opt.max_trials < SWB_MAX_TRIALS_CAP ? net.opt.max_trials : SWB_MAX_TRIALS_CAP;
}
SWB_ENGINE inline bool member_active(const State &st, const RunArgs &args, int m)
{
    return !((args.phases & PH_ADVANCE) && st.done[m]);
}

// execRouting (swmm5.c:528-546): step = variable step, shortened to end at t_end (owner threads)
SWB_ENGINE inline void ph_advance(const Net &net, const State &st, const RunArgs &args, int m)
{
    const double tms = st.time_ms[m];
    int done = (st.sim_time[m] >= args.t_end) ? 1 : 0;
    double dt = st.var_step[m];
    if (net.opt.courant_factor == 0.0) dt = args.fixed_step;
    else if (dt == 0.0) {      // first call of dynwave_getRoutingStep (dynwave.c:209-218)
        dt = floor(1000.0 * net.opt.min_route_step) / 1000.0;
        st.var_step[m] = dt;
    }
    // time is kept in milliseconds like NewRoutingTime (routing.c:301)
    dt = ctl_clamp_step(args.controls, m, tms, dt);
    double nextms = tms + 1000.0 * dt, endms = 1000.0 * args.t_end;
    if (!done && nextms > endms) {
        dt = (endms - tms) / 1000.0;
        dt = SWB_MAX(dt, 1. / 1000.0);
    }
    st.dt[m] = dt;
    st.done[m] = done;
}

// massbal_updateRoutingTotals (massbal.c:587-617): the previous step's rates over the second half of
// that step (deferred from its end) and over the first half of this one (owner threads)
SWB_ENGINE inline void ph_massbal_fold(const Net &net, const State &st, int m, double dt)
{
    const int M = st.M, nMb = MB_FLOW_TERMS + MB_QUAL_TERMS * net.nP;
    const double dtPrev = st.mb_dt_prev[m];
    for (int q = 0; q < nMb; q++) {
        const double r = st.mb_rate[q * M + m];
        double tot = st.mb_total[q * M + m];
        // (the mass moved to "final storage" is added unweighted by BOTH half-step
        // updates, massbal.c:614 -- kept, it is what the reference reports)
        const bool mass = q >= MB_FLOW_TERMS && (q - MB_FLOW_TERMS) % MB_QUAL_TERMS == MBQ_FINAL;
        tot += mass ? r : r * (dtPrev / 2.);
        tot += mass ? r : r * (dt / 2.);
        st.mb_total[q * M + m] = tot;
        st.mb_rate[q * M + m] = 0.0;
    }
    st.mb_dt_prev[m] = dt;
}

// old <- new, node_initFlows, overflow reset, lateral inflows + quality preload, initRoutingStep
template <class Ctx>
SWB_ENGINE inline void ph_prologue(const Net &net, const State &st, const RunArgs &args, Ctx &ctx, const ThreadMap &tm, double dt)
{
    const int M = st.M, nN = net.nN, nL = net.nL, nP = net.nP;
    const int m = tm.m, first = tm.first, stride = tm.stride;
    // getDateTime(NewRoutingTime): 1 ms after the routing time (swmm5.c:1551), in days
    const double tNow = args.inflows.start_day +
        (args.inflows.start_secs + (st.time_ms[m] + 1.0) / 1000.0) / 86400.0;
    // Every body below issues ALL of its loads before its first store: the state arrays may
    // alias as far as the compiler knows, so a load written after a store cannot be hoisted
    // above it and each load/store pair would cost a full DRAM round trip.
    const int ph = args.phases;
    const Inflows &F = args.inflows;
    const bool general = (ph & PH_INFLOWS) && F.general;
    DateParts dp = {0, 0, 0};
    if (general && F.nPatterns > 0) dp = date_parts(tNow);
    IfaceAt ifAt = { -1, 0.0 };
    if (general && F.nIfaceRec > 0) ifAt = iface_bracket(F, tNow);
    double exq[SWB_MAX_POLLUT];          // external quality mass rate booked by this thread (general inflows)
#pragma unroll
    for (int p = 0; p < SWB_MAX_POLLUT; p++) exq[p] = 0.0;
    SWB_FOR_ITEMS(i, nN) {
        const size_t ix = SWB_IX(i, m, M);
        // ---- loads
        double qv[SWB_MAX_POLLUT], hq[SWB_MAX_POLLUT];
        const bool touchQual = (ph & (PH_QSWAP | PH_INFLOWS)) || ((ph & PH_HOSTIN) && args.host_qual);
#pragma unroll
        for (int p = 0; p < SWB_MAX_POLLUT; p++) {
            qv[p] = 0.0; hq[p] = 0.0;
            if (p < nP && touchQual) {
                size_t iq = SWB_IXP(p, i, nN, m, M);
                qv[p] = st.n_qual[iq];
                if ((ph & PH_HOSTIN) && args.host_qual) hq[p] = args.host_qual[iq];
            }
        }
        double depth = 0.0, volume = 0.0, inflow = 0.0, outflow = 0.0, lat = 0.0, losses = 0.0;
        if (ph & PH_SWAP) {
            depth = st.n_depth[ix]; volume = st.n_volume[ix];
            inflow = st.n_inflow[ix]; outflow = st.n_outflow[ix];
        }
        bool newLat = false;
        int slot = -1;
        // Node.oldLatFlow = newLatFlow (routing.c:330), only where this step sets a new one
        const double prevLat = (ph & (PH_INFLOWS | PH_HOSTIN)) ? st.n_latflow[ix] : 0.0;
        double extq = 0.0, dwfq = 0.0, revq = 0.0, ifq = 0.0;
        bool ifOn = false;
        if (ph & PH_INFLOWS) {
            // addExternalInflows (routing.c:435-490)
            slot = F.node_slot[i];
            if (slot >= 0) {
                double tsv = inflow_series(F, slot, tNow - F.member_shift[m]) * (F.sfactor[slot] * F.member_scale[m]);
                double blv = F.baseline[slot];
                if (general && F.base_pattern && F.base_pattern[slot] >= 0) blv *= pattern_factor(F, F.base_pattern[slot], dp);
                lat = tsv + blv;
                if (general && F.cfactor) lat = F.cfactor[slot] * lat;
                if (fabs(lat) < SWB_FLOW_TOL) lat = 0.0;
            }
            extq = lat;
            if (general) {
                // reverse flow through an outfall joins the flow that carries CONCEN inflows (routing.c:474);
                // Node.oldNetInflow still holds the value of the step before the swap below
                if (F.node_q_start && F.node_q_start[i + 1] > F.node_q_start[i] && net.node_type[i] == SWB_OUTFALL) {
                    const double oni = st.n_old_net_inflow[ix];
                    if (oni < 0.0) revq = oni;
                }
                // addDryWeatherInflows (routing.c:499-575): the flow record of the node's list
                if (F.node_dwf_start)
                    for (int r = F.node_dwf_start[i]; r < F.node_dwf_start[i + 1]; r++)
                        if (F.dwf_param[r] < 0) {
                            dwfq = dwf_value(F, r, dp);
                            if (fabs(dwfq) < SWB_FLOW_TOL) dwfq = 0.0;
                            lat += dwfq;
                            break;
                        }
            }
            if (general && F.if_slot && F.if_slot[i] >= 0 && ifAt.rec >= 0) {
                // addIfaceInflows (routing.c:736-775)
                const double q = iface_value(F, ifAt, F.if_slot[i], 0, nP);
                if (!(fabs(q) < SWB_FLOW_TOL)) { ifq = q; lat += q; ifOn = true; }
            }
            newLat = true;
        }
        if (ph & PH_HOSTIN) {
            lat = args.host_lat[ix];
            losses = args.host_losses ? args.host_losses[ix] : 0.0;
            newLat = true;
        }
        if (!newLat && (ph & PH_SWAP)) { lat = st.n_latflow[ix]; losses = st.n_losses[ix]; }
        // ---- stores
#pragma unroll
        for (int p = 0; p < SWB_MAX_POLLUT; p++)
            if (p < nP && touchQual) {
                size_t iq = SWB_IXP(p, i, nN, m, M);
                double c = qv[p];
                if (ph & PH_QSWAP) { st.n_old_qual[iq] = c; c = 0.0; }      // routing.c:312-336
                if ((ph & PH_INFLOWS) && !general) { if (slot >= 0 && lat >= 0.0) c += F.concen[slot * nP + p] * lat; }
                else if (ph & PH_INFLOWS) {
                    double wsum = 0.0;                      // (for the routing totals; c is summed in the reference's order)
                    if (extq >= 0.0) {                      // a negative external flow takes no pollutant inflow (:466-470)
                        if (!F.node_q_start) { if (slot >= 0) { double w = F.concen[slot * nP + p] * extq; c += w; wsum += w; } }
                        else {
                            const double qq = extq - revq;
                            for (int r = F.node_q_start[i]; r < F.node_q_start[i + 1]; r++) {
                                if (F.q_pollut[r] != p) continue;
                                // inflow_getExtInflow (inflow.c:207-234)
                                double blv = F.q_baseline[r], tsv = 0.0;
                                if (F.q_pattern[r] >= 0) blv *= pattern_factor(F, F.q_pattern[r], dp);
                                if (F.q_series[r] >= 0)
                                    tsv = series_lookup(F.series_start, F.series_t, F.series_v, F.q_series[r], tNow) * F.q_sfactor[r];
                                double w = F.q_cfactor[r] * (tsv + blv);
                                if (F.q_type[r] == 1) w *= qq;          // CONCEN_INFLOW
                                c += w; wsum += w;
                            }
                        }
                    }
                    if (dwfq > 0.0) {                       // (:536-572)
                        const double dc = F.pollut_dwf_concen ? F.pollut_dwf_concen[p] : 0.0;
                        if (dc > 0.0) { double w = dwfq * dc; c += w; wsum += w; }
                        for (int r = F.node_dwf_start[i]; r < F.node_dwf_start[i + 1]; r++) {
                            if (F.dwf_param[r] != p) continue;
                            double w = dwfq * dwf_value(F, r, dp);
                            c += w; wsum += w;
                            if (dc > 0.0) { w = dwfq * dc; c -= w; wsum -= w; }
                        }
                    }
                    if (ifOn) { const double w = ifq * iface_value(F, ifAt, F.if_slot[i], 1 + p, nP); c += w; wsum += w; }
                    exq[p] += wsum;
                }
                if ((ph & PH_HOSTIN) && args.host_qual) c += hq[p];
                st.n_qual[iq] = c;
            }
        if (newLat) { st.n_latflow[ix] = lat; st.n_losses[ix] = losses; st.n_old_latflow[ix] = prevLat; }
        if (ph & PH_SWAP) {
            // node_setOldHydState, node_initFlows, flowrout.c:153-162
            st.n_old_depth[ix] = depth;
            st.n_old_volume[ix] = volume;
            st.n_old_net_inflow[ix] = inflow - outflow;
            st.n_old_inflow[ix] = inflow;              // oldFlowInflow (node.c:302)
            st.n_inflow[ix] = lat;
            st.n_outflow[ix] = losses;
            double ov = 0.0, fullVolume = net.node_full_volume[i];
            if (net.node_type[i] != SWB_STORAGE && volume > fullVolume) ov = (volume - fullVolume) / dt;
            st.n_overflow[ix] = ov;
        }
        if (ph & PH_DYNWAVE) {                 // initRoutingStep (dynwave.c:276-293)
            st.n_converged[ix] = 0;
            st.n_dydt[ix] = 0.0;
        }
    }
    if (general && (ph & PH_MASSBAL))
#pragma unroll
        for (int p = 0; p < SWB_MAX_POLLUT; p++)
            if (p < nP) {
                double v = exq[p];
                if (M == 1) v = ctx.warp_sum_f64(v);
                if (v != 0.0 && (M != 1 || ctx.warp_lane == 0))
                    ctx.atomic_add_f64(&st.mb_rate[(size_t)(MB_FLOW_TERMS + p * MB_QUAL_TERMS + MBQ_EX_INFLOW) * M + m], v);
            }
    SWB_FOR_ITEMS(j, nL) {
        const size_t ix = SWB_IX(j, m, M);
        // ---- loads
        double qv[SWB_MAX_POLLUT];
#pragma unroll
        for (int p = 0; p < SWB_MAX_POLLUT; p++) {
            qv[p] = 0.0;
            if (p < nP && (ph & PH_QSWAP)) qv[p] = st.l_qual[SWB_IXP(p, j, nL, m, M)];
        }
        double depth = 0.0, flow = 0.0, volume = 0.0, a1 = 0.0;
        const bool isConduit = (net.link_type[j] == SWB_CONDUIT);
        if (ph & PH_SWAP) { depth = st.l_depth[ix]; flow = st.l_flow[ix]; volume = st.l_volume[ix]; }
        if ((ph & PH_DYNWAVE) && isConduit) a1 = st.c_a1[ix];
        // ---- stores
#pragma unroll
        for (int p = 0; p < SWB_MAX_POLLUT; p++)
            if (p < nP && (ph & PH_QSWAP)) {
                size_t iq = SWB_IXP(p, j, nL, m, M);
                st.l_old_qual[iq] = qv[p];
                st.l_qual[iq] = 0.0;
            }
        if (ph & PH_SWAP) {                     // link_setOldHydState (link.c:564-583)
            st.l_old_depth[ix] = depth;
            st.l_old_flow[ix] = flow;
            st.l_old_volume[ix] = volume;
        }
        if (ph & PH_DYNWAVE) {
            st.l_bypassed[ix] = 0;
            if (!(net.link_flags[j] & LF_TRUE_CONDUIT)) {
                st.l_surf_area1[ix] = 0.0;
                st.l_surf_area2[ix] = 0.0;
            }
            if (isConduit) st.c_a2[ix] = a1;    // dynwave.c:292
        }
    }
}

// updateConvergenceStats, findLimitedLinks (dynwave.c:257-260, 349-378); active members only
SWB_ENGINE inline void ph_epilogue(const Net &net, const State &st, const ThreadMap &tm)
{
    const int M = st.M, nL = net.nL, m = tm.m, first = tm.first, stride = tm.stride;
    const int maxTrials = engine_max_trials(net);
    if (tm.owner) {
        // iterations used: first trial k >= 1 that ended converged, else MaxTrials
        int itersDone = maxTrials;
        for (int kk = 1; kk < maxTrials; kk++)
            if (st.not_conv[kk * M + m] == 0) { itersDone = kk + 1; break; }
        if (maxTrials == 1) itersDone = 1;
        st.iters[m] = itersDone;
        st.tot_iters[m] += itersDone;
        st.tot_steps[m] += 1;
        if (st.not_conv[(itersDone - 1) * M + m]) st.non_conv[m] += 1;
    }
    // stats_updateConvergenceStats (dynwave.c:266-272): every node that missed the
    // tolerance in a step that ended without convergence (outfalls count: their flag is
    // never raised, dynwave.c:611)
    // (node-level convergence counts: see the statistics phase below)
    SWB_FOR_ITEMS(j, nL) {
        if (!(net.link_flags[j] & LF_TRUE_CONDUIT)) continue;
        size_t ix = SWB_IX(j, m, M);
        unsigned char lim = 0;
        if (st.c_a1[ix] >= net.xs_afull[j]) {
            int n1 = net.link_node1[j], n2 = net.link_node2[j];
            double h1 = st.n_depth[SWB_IX(n1, m, M)] + net.node_invert[n1];
            double h2 = st.n_depth[SWB_IX(n2, m, M)] + net.node_invert[n2];
            // Conduit.length is the user's length (dynwave.c:374), not the true length
            if ((h1 - h2) > fabs(net.cond_slope[j]) * net.cond_user_length[j]) lim = 1;
        }
        st.c_cap_limited[ix] = lim;
    }
}

// qualrout_execute (qualrout.c:100-142), node pass then link pass; pollutant-major loops keep the
// three mass-balance partial sums in registers
template <class Ctx>
SWB_ENGINE inline void ph_qual_nodes(const Net &net, const State &st, Ctx &ctx, const ThreadMap &tm, double dt, int nNo)
{
    const int M = st.M, m = tm.m, first = tm.first, stride = tm.stride;
    // pollutants in pairs: one sweep over the nodes serves two of them (swb_qual.h)
    for (int p0 = 0; p0 < net.nP; p0 += 2) {
        if (p0 + 1 < net.nP) {
            QualAcc acc[2] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
            SWB_FOR_ITEMS(i, nNo) qual_node<2>(net, st, i, m, p0, dt, acc);
            qual_acc_flush(ctx, st, p0, m, M, dt, acc[0]);
            qual_acc_flush(ctx, st, p0 + 1, m, M, dt, acc[1]);
        } else {
            QualAcc acc[1] = {{0.0, 0.0, 0.0}};
            SWB_FOR_ITEMS(i, nNo) qual_node<1>(net, st, i, m, p0, dt, acc);
            qual_acc_flush(ctx, st, p0, m, M, dt, acc[0]);
        }
    }
}
template <class Ctx>
SWB_ENGINE inline void ph_qual_links(const Net &net, const State &st, Ctx &ctx, const ThreadMap &tm, double dt)
{
    const int M = st.M, nL = net.nL, m = tm.m, first = tm.first, stride = tm.stride;
    const Halo &H = st.halo;
    const bool part = H.nRanks > 1;
    for (int p0 = 0; p0 < net.nP; p0 += 2) {
        // the copy of a cut conduit is routed too, but only its owner reports the losses
        if (p0 + 1 < net.nP) {
            QualAcc acc[2] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}, copy[2] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
            SWB_FOR_ITEMS(j, nL) { if (part && !H.link_owned[j]) qual_link<2>(net, st, j, m, p0, dt, copy);
                                   else qual_link<2>(net, st, j, m, p0, dt, acc); }
            qual_acc_flush(ctx, st, p0, m, M, dt, acc[0]);
            qual_acc_flush(ctx, st, p0 + 1, m, M, dt, acc[1]);
        } else {
            QualAcc acc[1] = {{0.0, 0.0, 0.0}}, copy[1] = {{0.0, 0.0, 0.0}};
            SWB_FOR_ITEMS(j, nL) { if (part && !H.link_owned[j]) qual_link<1>(net, st, j, m, p0, dt, copy);
                                   else qual_link<1>(net, st, j, m, p0, dt, acc); }
            qual_acc_flush(ctx, st, p0, m, M, dt, acc[0]);
        }
    }
}

// removeSystemOutflows (routing.c:776-806, 812-925).  Step rates of everything that enters or leaves
// the system, summed per member: external inflows (routing.c:466-489), outfall discharge and flooding
// (node.c:438-493), negative lateral flows, storage and conduit losses.  Per-thread partial sums, one
// atomic per term.
template <class Ctx>
SWB_ENGINE inline void ph_outflows(const Net &net, const State &st, const RunArgs &args, Ctx &ctx, const ThreadMap &tm,
                                   double dt, bool withQual, int nNo)
{
    const int M = st.M, nN = net.nN, nL = net.nL, nP = net.nP, m = tm.m, first = tm.first, stride = tm.stride;
    const Halo &H = st.halo;
    const bool part = H.nRanks > 1;
    auto flush = [&](int q, double v) {
        if (M == 1) v = ctx.warp_sum_f64(v);
        if (v != 0.0 && (M != 1 || ctx.warp_lane == 0)) ctx.atomic_add_f64(&st.mb_rate[(size_t)q * M + m], v);
    };
    // pass 0: the flow terms and pollutants 0-1; further passes: two more pollutants each
    // (keeps the partial sums in registers)
    for (int p0 = 0; p0 == 0 || (withQual && p0 < nP); p0 += 2) {
        double fl[MB_FLOW_TERMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
        double ql[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        SWB_FOR_ITEMS(i, nNo) {
            const size_t ix = SWB_IX(i, m, M);
            const double lat = st.n_latflow[ix];
            const int type = net.node_type[i];
            double q = 0.0;                       // node_getSystemOutflow (node.c:438-493)
            bool flooded = false;
            if (type == SWB_OUTFALL) {            // (warp-uniform: a warp holds one node)
                const double inflow = st.n_inflow[ix], outflow = st.n_outflow[ix];
                if (outflow == 0.0) q = inflow;
                else if (inflow == 0.0) q = -outflow;
            } else {
                // q = overflow if newVolume <= fullVolume: the volume only matters when
                // there is an overflow at all, so junctions cost two loads here
                const double ov = st.n_overflow[ix];
                if (ov != 0.0 && st.n_volume[ix] <= net.node_full_volume[i]) q = ov;
                flooded = q > 0.0;
            }
            if (p0 == 0) {
                if (lat >= 0.0) fl[MB_EX_INFLOW] += lat; else fl[MB_OUTFLOW] -= lat;      // routing.c:466-469
                if (q > 0.0) { if (flooded) fl[MB_FLOODING] += q; else fl[MB_OUTFLOW] += q; }  // :904
                else fl[MB_EX_INFLOW] -= q;                                                // :911
                if (type == SWB_STORAGE) {                                                 // :812-836
                    fl[MB_EVAP] += st.n_evap_loss[ix] / dt;
                    fl[MB_SEEP] += st.n_exfil_loss[ix] / dt;
                }
            }
            if (withQual) {
                const int slot = (args.phases & PH_INFLOWS) ? args.inflows.node_slot[i] : -1;
#pragma unroll
                for (int pp = 0; pp < 2; pp++) {
                    const int p = p0 + pp;
                    if (p >= nP) continue;
                    const size_t iq = SWB_IXP(p, i, nN, m, M);
                    if (slot >= 0 && lat >= 0.0 && !args.inflows.general) ql[pp][MBQ_EX_INFLOW] += args.inflows.concen[slot * nP + p] * lat;
                    if ((args.phases & PH_HOSTIN) && args.host_qual) ql[pp][MBQ_EX_INFLOW] += args.host_qual[iq];
                    if (q > 0.0 || lat < 0.0) {   // the concentration only where mass leaves
                        const double c = st.n_qual[iq];
                        if (q > 0.0) ql[pp][flooded ? MBQ_FLOODING : MBQ_OUTFLOW] += q * c;  // :906-908
                        if (lat < 0.0) ql[pp][MBQ_OUTFLOW] += -lat * c;                      // :915-922
                    }
                }
            }
        }
        if (p0 == 0 && net.anyLossRate)                    // removeConduitLosses (:840-867)
            SWB_FOR_ITEMS(j, nL) {
                if (!(net.link_flags[j] & LF_HAS_LOSSRATE) || (part && !H.link_owned[j])) continue;
                const size_t ix = SWB_IX(j, m, M);
                const double barrels = (double)net.cond_barrels[j];
                fl[MB_EVAP] += st.c_evap_loss[ix] * barrels;
                fl[MB_SEEP] += st.c_seep_loss[ix] * barrels;
            }
        if (p0 == 0)
#pragma unroll
            for (int q = 0; q < MB_FLOW_TERMS; q++) flush(q, fl[q]);
        if (withQual)
#pragma unroll
            for (int pp = 0; pp < 2; pp++)
                if (p0 + pp < nP)
#pragma unroll
                    for (int t = 0; t < 3; t++) flush(MB_FLOW_TERMS + (p0 + pp) * MB_QUAL_TERMS + t, ql[pp][t]);
    }
}

// statistics (routing.c:256-260, stats.c:449-754); the caller has checked st.stat_node
SWB_ENGINE inline void ph_stats(const Net &net, const State &st, const RunArgs &args, const ThreadMap &tm, bool active,
                                double dt, bool withQual, int nNo, const double *T)
{
    const int M = st.M, nL = net.nL, m = tm.m, first = tm.first, stride = tm.stride;
    const Halo &H = st.halo;
    const bool part = H.nRanks > 1;
    const int maxTrials = engine_max_trials(net);
    if (!active) return;
    // new routing time of this step in elapsed seconds, formed like NewRoutingTime (ms)
    const double tNew = (st.time_ms[m] + 1000.0 * dt) / 1000.0;
    if (tNew >= st.stat_report_start) {
        // this thread's copy of "iterations used" (the owner's write to st.iters is not
        // ordered against other CTAs): the step ended unconverged iff its last trial did
        int itersDone = maxTrials;
        for (int kk = 1; kk < maxTrials; kk++)
            if (st.not_conv[kk * M + m] == 0) { itersDone = kk + 1; break; }
        if (maxTrials == 1) itersDone = 1;
        const bool stepFailed = (args.phases & PH_DYNWAVE) && st.not_conv[(itersDone - 1) * M + m] != 0;
        SWB_FOR_ITEMS(i, nNo) {
            stats_node(net, st, i, m, dt, tNew, withQual);
            if (stepFailed && !st.n_converged[SWB_IX(i, m, M)]) nstat(st, net, SWB_NS_NONCONV_COUNT, i, m) += 1.0;
        }
        SWB_FOR_ITEMS(j, nL) {
            if (part && !H.link_owned[j]) continue;
            stats_link(net, st, j, m, dt, tNew, T);
        }
        if (tm.owner) {
            double *sys = st.stat_sys;
            sys[SWB_SS_REPORT_STEPS * M + m] += 1.0;
            sys[SWB_SS_ROUTING_SPAN * M + m] += dt;
            // SysOutfallFlow: the outfalls' inflows summed in node order like stats.c:626 (the
            // node phase wrote them several grid barriers ago); MaxOutfallFlow = running maximum
            double sysOut = 0.0;
            for (int k = 0; k < net.nOutfallNodes; k++) {
                const int i = net.outfall_nodes[k];
                if (i < nNo) sysOut += st.n_inflow[SWB_IX(i, m, M)];
            }
            double &mx = sys[SWB_SS_MAX_OUTFALL_FLOW * M + m];
            mx = SWB_MAX(mx, sysOut);
        }
    }
    if (tm.owner) {                      // stats_updateTimeStepStats (stats.c:486-518), no steady state
        double *sys = st.stat_sys;
        if (st.sim_time[m] > 0.0) {   // OldRoutingTime > 0: the first step does not set the minimum
            double &mn = sys[SWB_SS_MIN_DT * M + m];
            mn = SWB_MIN(mn, dt);
        }
        double &mx = sys[SWB_SS_MAX_DT * M + m];
        mx = SWB_MAX(mx, dt);
        sys[SWB_SS_ROUTING_TIME * M + m] += dt;
        sys[SWB_SS_STEP_COUNT * M + m] += 1.0;
        sys[SWB_SS_TRIALS * M + m] += (double)st.iters[m];
    }
}

// dynwave_getRoutingStep (dynwave.c:195-220, 799-921).  Two-level arg-min with integer atomics only
// (deterministic): the minimum itself as the ordered bit image of the positive double, then the
// smallest object index among the objects that attain it ("first index wins", like the reference's
// strict `<` loops).  reset (owner) | barrier | search | barrier | arg | barrier | final (owner).
struct DtCand { double tl, tn; int il, in; };    // a thread's own best link / node candidate
SWB_ENGINE inline void ph_nextdt_reset(const State &st, const RunArgs &args, int m)
{
    const int M = st.M;
    unsigned long long t0 = dbits(args.fixed_step);
    st.tmin_bits[m] = t0;                 // links: min(maxStep, link candidates)
    st.tmin_bits[M + m] = t0;             // nodes
    st.crit_link[m] = 0x7fffffff; st.crit_node[m] = 0x7fffffff;
}
SWB_ENGINE inline bool nextdt_variable(const Net &net, const RunArgs &args)
{
    return !(net.opt.courant_factor == 0.0 || args.fixed_step < SWB_MINTIMESTEP);
}
template <class Ctx>
SWB_ENGINE inline DtCand ph_nextdt_search(const Net &net, const State &st, const RunArgs &args, Ctx &ctx,
                                          const ThreadMap &tm, bool search, int nNo)
{
    const int M = st.M, nL = net.nL, m = tm.m, first = tm.first, stride = tm.stride;
    DtCand c;
    c.tl = args.fixed_step; c.tn = args.fixed_step; c.il = -1; c.in = -1;
    if (search) {
        SWB_FOR_ITEMS(j, nL) {
            double t = link_step(net, st, j, m);
            if (t >= 0.0 && t < c.tl) { c.tl = t; c.il = j; }
        }
        SWB_FOR_ITEMS(i, nNo) {
            double t = node_step(net, st, i, m);
            if (t >= 0.0 && t < c.tn) { c.tn = t; c.in = i; }
        }
        unsigned long long bl = dbits(c.tl), bn = dbits(c.tn);
        if (M == 1) {          // all lanes share the member: shuffle-reduce, one atomic per warp
            bl = ctx.warp_min_u64(bl); bn = ctx.warp_min_u64(bn);
            if (ctx.warp_lane == 0) {
                if (bl < dbits(args.fixed_step)) ctx.atomic_min_u64(&st.tmin_bits[m], bl);
                if (bn < dbits(args.fixed_step)) ctx.atomic_min_u64(&st.tmin_bits[M + m], bn);
            }
        } else {
            if (c.il >= 0) ctx.atomic_min_u64(&st.tmin_bits[m], bl);
            if (c.in >= 0) ctx.atomic_min_u64(&st.tmin_bits[M + m], bn);
        }
    }
    return c;
}
template <class Ctx>
SWB_ENGINE inline void ph_nextdt_arg(const State &st, Ctx &ctx, int m, const DtCand &c)
{
    const int M = st.M;
    if (c.il >= 0 && dbits(c.tl) == st.tmin_bits[m]) ctx.atomic_min_i32(&st.crit_link[m], c.il);
    if (c.in >= 0 && dbits(c.tn) == st.tmin_bits[M + m]) ctx.atomic_min_i32(&st.crit_node[m], c.in);
}
SWB_ENGINE inline void ph_nextdt_final(const Net &net, const State &st, const RunArgs &args, int m)
{
    const int M = st.M;
    const bool part = st.halo.nRanks > 1;
    if (!nextdt_variable(net, args)) st.var_step[m] = args.fixed_step;
    else {
        double vs;
        int cl = st.crit_link[m], cn = st.crit_node[m];
        if (cl == 0x7fffffff) cl = -1;
        if (cn == 0x7fffffff) cn = -1;
        if (st.var_step[m] == 0.0) { vs = net.opt.min_route_step; cl = cn = -1; }
        else {
            // getVariableStep (dynwave.c:813-831): a node wins only when strictly below
            // the link step, and then the link is dropped
            double tLink = dfrombits(st.tmin_bits[m]), tNode = dfrombits(st.tmin_bits[M + m]);
            vs = tLink;
            // (partitioned: the critical node may live on another rank; a node
            // candidate exists iff the reduced node minimum is below the fixed step)
            const bool haveNode = cn >= 0 || (part && st.tmin_bits[M + m] < dbits(args.fixed_step));
            if (haveNode && tNode < tLink) { vs = tNode; cl = -1; } else cn = -1;
            if (vs < net.opt.min_route_step) vs = net.opt.min_route_step;
        }
        st.crit_link[m] = cl; st.crit_node[m] = cn;
        st.var_step[m] = floor(1000.0 * vs) / 1000.0;
        // stats_updateCriticalTimeCount (stats.c:522-533; called from getVariableStep)
        if (st.stat_node && (args.phases & PH_STATS)) {
            if (cn >= 0) nstat(st, net, SWB_NS_TIME_COURANT, cn, m) += 1.0;
            else if (cl >= 0) lstat(st, net, SWB_LS_TIME_COURANT, cl, m) += 1.0;
        }
    }
}
// NewRoutingTime += 1000 * routingStep (routing.c:301-302), kept in ms
SWB_ENGINE inline void ph_advance_time(const State &st, int m, double dt)
{
    st.time_ms[m] = st.time_ms[m] + 1000.0 * dt;
    st.sim_time[m] = st.time_ms[m] / 1000.0;
}

template <class Ctx>
SWB_ENGINE inline void engine_run(const Net &net, const State &st, const RunArgs &args, Ctx &ctx)
{
    const int M = st.M, nN = net.nN, nL = net.nL, nP = net.nP;
    const ThreadMap tm = thread_map(ctx, M);
    const int m = tm.m;
    const bool owner = tm.owner;
    const double *T = ctx.T;
    const int maxTrials = engine_max_trials(net);
    const bool withQual = (nP > 0) && !net.opt.ignore_quality;
    // partitioned network (M == 1): nodes [0, nNo) are this rank's, the rest are ghosts kept current
    // by halo_exchange; every decision that shapes the control flow is reduced over all ranks
    const Halo &H = st.halo;
    const bool part = H.nRanks > 1;
    const int nNo = part ? H.nOwnedN : nN;
    unsigned long long epoch = part ? H.ctrl[HALO_EPOCH] : 0ull;
    const int *nodeOrder = part ? H.node_order : net.node_order;

    unsigned long long tmark = (ctx.tid == 0) ? ctx.now_ns() : 0ull;

    if (args.stg_lat || args.stg_losses || args.stg_qual) {
        // host layout [m][item][p] -> device layout [(p, item)][m]
        if (args.stg_lat) ctx.transpose(const_cast<double *>(args.host_lat), args.stg_lat, M, nN, 1);
        if (args.stg_losses) ctx.transpose(const_cast<double *>(args.host_losses), args.stg_losses, M, nN, 1);
        if (args.stg_qual) ctx.transpose(const_cast<double *>(args.host_qual), args.stg_qual, M, nN * nP, nP);
        ctx.grid_sync();
    }
    ctx.load_tables(net.xs_tables);

    for (int step = 0; step < args.n_steps; step++) {
        // ================= step prologue =======================================================
        if (args.phases & PH_ADVANCE) {
            if (owner) {
                ph_advance(net, st, args, m);
                // evaluateControlRules (routing.c:269-308): reads the previous step's results, before
                // the prologue replaces them
                if (args.controls.active && !st.done[m]) ph_controls(net, st, args.controls, m, st.dt[m], T);
                // (before the barrier: the prologue of OTHER threads books external quality mass into the
                // rates this fold zeroes)
                if ((args.phases & PH_MASSBAL) && !st.done[m]) ph_massbal_fold(net, st, m, st.dt[m]);
            }
            ctx.grid_sync();
            // every member has reached t_end: leave the launch (grid-uniform decision)
            bool anyLeft = false;
            for (int mm = ctx.lane; mm < M; mm += ctx.block_size) anyLeft = anyLeft || !st.done[mm];
            if (!ctx.block_or(anyLeft)) break;
        }
        const bool active = member_active(st, args, m);
        const double dt = st.dt[m];
        if ((args.phases & PH_MASSBAL) && !(args.phases & PH_ADVANCE) && owner && active) ph_massbal_fold(net, st, m, dt);

        if (active && (args.phases & (PH_SWAP | PH_INFLOWS | PH_QSWAP | PH_DYNWAVE | PH_HOSTIN)))
            ph_prologue(net, st, args, ctx, tm, dt);
        if (owner && (args.phases & PH_DYNWAVE))
            for (int k = 0; k < maxTrials; k++) st.not_conv[k * M + m] = 0;
        for (int c = ctx.tid; c < SWB_TICKETS_PER_TRIAL * SWB_MAX_TRIALS_CAP; c += ctx.G) st.tickets[c] = 0ull;
        ctx.grid_sync();
        SWB_TICK(TP_PROLOGUE);

        // ================= dynwave_execute: Picard iterations ==================================
        // The members that still iterate are kept as an ordered list (all active members for
        // trials 0 and 1, then only those that have not converged, dynwave.c:249-251).  Work is cut
        // into TILES = (object, 32 consecutive list entries) and handed to warps through an atomic
        // ticket counter per phase: warps stay full when only a few members keep iterating, and a
        // warp that drew cheap tiles (dry or bypassed conduits) simply draws more instead of
        // waiting at the barrier for the warps that drew surcharged ones.
        if (args.phases & PH_DYNWAVE) {
            int nAlive = ctx.compact_members(M, [&](int mm) { return member_active(st, args, mm); });
            for (int k = 0; k < maxTrials && nAlive > 0; k++) {
                unsigned long long *tickets = st.tickets + SWB_TICKETS_PER_TRIAL * k;
                // ---- findLinkFlows, pass (i): true conduits (dynwave.c:387-395)
                if (!(args.debug & DBG_SKIP_LINKS) && M == 1 && ctx.tid < net.nPre)
                    picard_link_single(net, st, net.pre_links[ctx.tid], k, st.dt[0], T, true);
                if (!(args.debug & DBG_SKIP_LINKS))
                for_tiles(ctx, net.nTrue, nAlive, tickets + TK_LINKS, [&](int jj, int mm) {
                    if (M == 1) picard_link_single(net, st, jj, k, st.dt[0], T, false);
                    else picard_link(net, st, net.link_order[jj], mm, k, st.dt[mm], T);
                });
                ctx.grid_sync();
                SWB_TICK(TP_LINKS);
                // ---- networks with regulators / dummy links: ordered pass (A.4)
                if (net.nNonConduit > 0) {
                    for_tiles(ctx, nNo, nAlive, tickets + TK_PRESUM, [&](int ii, int mm) { picard_node_presum(net, st, nodeOrder[ii], mm); });
                    ctx.grid_sync();
                    if (ctx.tid < nAlive) {
                        int mm = ctx.alive_list[ctx.tid];
                        regulator_pass(net, st, mm, k, st.dt[mm], T);
                    }
                    ctx.grid_sync();
                    SWB_TICK(TP_REGULATORS);
                }
                // ---- findNodeDepths (dynwave.c:593-632)
                if (!(args.debug & DBG_SKIP_NODES))
                for_tiles(ctx, nNo, nAlive, tickets + TK_NODES, [&](int ii, int mm) {
                    const int i = nodeOrder[ii];
                    if (!picard_node(net, st, i, mm, k, st.dt[mm], T)) st.not_conv[k * M + mm] = 1;
                    if (part && H.send_start[i + 1] > H.send_start[i]) {
                        // a border node: its new depth and converged flag go straight into the
                        // peers' windows (parity of the exchange that follows this phase)
                        const int par = (int)((epoch + 1ull) & 1ull);
                        const double d = st.n_depth[i], c = (double)st.n_converged[i];
                        for (int e = H.send_start[i]; e < H.send_start[i + 1]; e++) {
                            double *dst = H.peer_stage[H.send_rank[e]] + ((size_t)H.send_slot[e] * 2 + par) * H.W;
                            dst[0] = d; dst[1] = c;
                        }
                        ctx.fence_system();
                    }
                });
                ctx.grid_sync();
                SWB_TICK(TP_NODES);
                if (part) {
                    // border depths + converged flags to the peers; "converged" becomes the AND over
                    // all ranks (dynwave.c:248-251 tests the whole network)
                    unsigned long long rin[HALO_RED] = { (unsigned long long)st.not_conv[k * M], 0ull, 0ull, 0ull };
                    halo_exchange(ctx, H, epoch, 2, rin,
                        [&](int sx, int w) { int i = H.send_node[sx];
                                             return w == 0 ? st.n_depth[i] : (double)st.n_converged[i]; },
                        [&](int r, int w, double v) { int i = H.recv_node[r];
                                                      if (w == 0) st.n_depth[i] = v;
                                                      else st.n_converged[i] = (v != 0.0) ? 1 : 0; },
                        [&](const unsigned long long *rows) {
                            int any = 0;
                            for (int p = 0; p < H.nRanks; p++) any |= (rows[p * HALO_RED] != 0ull);
                            st.not_conv[k * M] = any; }, true);
                    SWB_TICK(TP_HALO);
                }
                // ---- loop control: Steps++ ; if (Steps > 1 && converged) break (:248-251).
                // Every CTA rebuilds the same ordered list of members that go on to trial k + 1.
                if (k + 1 >= maxTrials) break;
                if (k >= 1)
                    nAlive = ctx.compact_members(M, [&](int mm) { return member_iterates(st, args, mm, k); });
                SWB_TICK(TP_CONTROL);
            }
            if (active) ph_epilogue(net, st, tm);
        }

        SWB_TICK(TP_EPILOGUE);
        // ================= qualrout_execute (qualrout.c:100-142) ===============================
        if (withQual && (args.phases & PH_QUALITY)) {
            if (active) ph_qual_nodes(net, st, ctx, tm, dt, nNo);
            ctx.grid_sync();
            SWB_TICK(TP_QUAL_NODES);
            if (part) {                // a cut conduit mixes with its upstream node's new quality
                unsigned long long rin[HALO_RED] = { 0ull, 0ull, 0ull, 0ull };
                halo_exchange(ctx, H, epoch, nP, rin,
                    [&](int sx, int w) { return st.n_qual[SWB_IXP(w, H.send_node[sx], nN, 0, M)]; },
                    [&](int r, int w, double v) { st.n_qual[SWB_IXP(w, H.recv_node[r], nN, 0, M)] = v; },
                    [&](const unsigned long long *) {});
                SWB_TICK(TP_HALO);
            }
            if (active) ph_qual_links(net, st, ctx, tm, dt);
        }

        if ((args.phases & PH_MASSBAL) && active) ph_outflows(net, st, args, ctx, tm, dt, withQual, nNo);
        if ((args.phases & PH_STATS) && st.stat_node) ph_stats(net, st, args, tm, active, dt, withQual, nNo, T);

        SWB_TICK(TP_QUAL_LINKS);
        // ================= dynwave_getRoutingStep (dynwave.c:195-220, 799-921) =================
        if (args.phases & PH_NEXTDT) {
            if (owner) ph_nextdt_reset(st, args, m);
            ctx.grid_sync();
            const bool search = active && nextdt_variable(net, args) && st.var_step[m] != 0.0;
            const DtCand cand = ph_nextdt_search(net, st, args, ctx, tm, search, nNo);
            ctx.grid_sync();
            if (part) {                // MIN over the ranks (ordered bit images, like the atomics above)
                unsigned long long rin[HALO_RED] = { st.tmin_bits[0], st.tmin_bits[M], 0ull, 0ull };
                halo_exchange(ctx, H, epoch, 0, rin,
                    [&](int, int) { return 0.0; }, [&](int, int, double) {},
                    [&](const unsigned long long *rows) {
                        unsigned long long a = rows[0], b = rows[1];
                        for (int p = 1; p < H.nRanks; p++) {
                            a = rows[p * HALO_RED] < a ? rows[p * HALO_RED] : a;
                            b = rows[p * HALO_RED + 1] < b ? rows[p * HALO_RED + 1] : b;
                        }
                        st.tmin_bits[0] = a; st.tmin_bits[M] = b; });
                SWB_TICK(TP_HALO);
            }
            if (search) ph_nextdt_arg(st, ctx, m, cand);
            ctx.grid_sync();
            if (owner && active) ph_nextdt_final(net, st, args, m);
        }
        if ((args.phases & PH_ADVANCE) && owner && active) ph_advance_time(st, m, dt);
        if (step + 1 < args.n_steps) ctx.grid_sync();
        SWB_TICK(TP_NEXTDT);
    }
    if (part && ctx.tid == 0) H.ctrl[HALO_EPOCH] = epoch;
    if (args.stg_depth || args.stg_flow) {
        // device layout [item][m] -> host layout [m][item]
        ctx.grid_sync();
        if (args.stg_depth) ctx.transpose(args.stg_depth, st.n_depth, nN, M, 0);
        if (args.stg_flow) ctx.transpose(args.stg_flow, st.l_flow, nL, M, 0);
    }
}

} // namespace swb
#endif
