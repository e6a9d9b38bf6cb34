// swb_state.h -- HBM layout of the static network (Net) and of the lockstep ensemble state (State).
//
// Layout rules (DESIGN.md "Data layout in HBM"):
//   * static geometry is structure-of-arrays indexed by node / link, shared by every member and
//     small enough (about 2.2 MB for the 10k-node grid) to stay L2-resident;
//   * dynamic state is structure-of-arrays with the MEMBER index fastest: element (item, m) of a
//     field lives at field[item * M + m].  A warp therefore works on 32 members of ONE link or
//     node: every dynamic load/store is a fully coalesced 256-byte transaction, every static load
//     is a warp-uniform broadcast, and shape / link-type branches never diverge.  With M == 1 the
//     same indexing degenerates to the reference's plain per-object arrays.
#ifndef SWB_STATE_H
#define SWB_STATE_H

#include "swb_common.h"
#include "../../include/swmm_b200.h"

namespace swb {

// ---- descriptor arrays mirrored on the device, name for name (X(type, name, count-kind)) --------
// count-kind: N nodes, L links, C curves, C1 curves+1, CP curve points, T shape tables,
//             TT shape tables x table length, P pollutants
#define SWB_DESC_ARRAYS(X) \
    X(int, node_type, N) X(int, node_degree, N) X(double, node_invert, N) \
    X(double, node_full_depth, N) X(double, node_sur_depth, N) X(double, node_ponded_area, N) \
    X(double, node_full_volume, N) X(double, node_crown_elev, N) X(int, outfall_type, N) \
    X(int, outfall_flap, N) X(int, storage_shape, N) X(int, storage_curve, N) \
    X(double, storage_a0, N) X(double, storage_a1, N) X(double, storage_a2, N) \
    X(int, link_type, L) X(int, link_node1, L) X(int, link_node2, L) X(int, link_direction, L) \
    X(int, link_has_flap, L) X(double, link_offset1, L) X(double, link_offset2, L) \
    X(double, link_q_limit, L) X(double, link_q_full, L) X(double, link_closs_in, L) \
    X(double, link_closs_out, L) X(double, link_closs_avg, L) X(double, link_seep_rate, L) \
    X(int, xs_type, L) X(int, xs_culvert, L) X(int, xs_table, L) X(double, xs_yfull, L) \
    X(double, xs_wmax, L) X(double, xs_ywmax, L) X(double, xs_afull, L) X(double, xs_rfull, L) \
    X(double, xs_sfull, L) X(double, xs_smax, L) X(double, xs_ybot, L) X(double, xs_abot, L) \
    X(double, xs_sbot, L) X(double, xs_rbot, L) X(int, cond_barrels, L) X(int, cond_has_losses, L) \
    X(double, cond_length, L) X(double, cond_user_length, L) X(double, cond_mod_length, L) X(double, cond_rough_factor, L) \
    X(double, cond_slope, L) X(double, cond_beta, L) X(double, cond_q_max, L) X(int, pump_type, L) \
    X(int, pump_curve, L) X(double, pump_xmin, L) X(double, pump_xmax, L) X(int, orif_type, L) \
    X(double, orif_cdisch, L) X(double, orif_length, L) X(int, weir_type, L) \
    X(int, weir_can_surcharge, L) X(int, weir_cd_curve, L) X(double, weir_cdisch1, L) \
    X(double, weir_cdisch2, L) X(double, weir_end_con, L) X(double, weir_slope, L) \
    X(double, weir_length, L) X(double, weir_road_width, L) X(int, weir_road_surface, L) \
    X(int, outlet_curve, L) X(int, outlet_curve_type, L) \
    X(double, outlet_qcoeff, L) X(double, outlet_qexpon, L) X(int, curve_start, C1) \
    X(int, curve_type, C) X(double, curve_x, CP) X(double, curve_y, CP) X(int, shape_tbl_n, T) \
    X(double, shape_area_tbl, TT) X(double, shape_hrad_tbl, TT) X(double, shape_width_tbl, TT) \
    X(double, pollut_kdecay, P)

// link_flags bits (static, derived once on the host)
enum {
    LF_TRUE_CONDUIT = 1,     // type == CONDUIT && xsect != DUMMY (dynwave.c:416-419)
    LF_HAS_FLAP     = 2,     // Link.hasFlapGate
    LF_N1_OUTFALL   = 4,   LF_N2_OUTFALL = 8,
    LF_N1_OUT_FLAP  = 16,  LF_N2_OUT_FLAP = 32,   // end node is an outfall with a flap gate
    LF_N1_STORAGE   = 64,  LF_N2_STORAGE = 128,
    LF_OPEN_SHAPE   = 256,   // xsect_isOpen
    LF_HAS_LOSSRATE = 512    // seepRate > 0 or open shape (evaporation possible)
};

enum { LK_CIRCULAR = 0, LK_RECT_CLOSED = 1, LK_GENERIC = 2 };

// one incident link end of a node, everything static the gather needs in ONE 16-byte load
// (instead of the chain adj -> link_flags -> link_type -> cond_barrels -> pump_type)
struct alignas(16) AdjEntry {
    int je;          // (link << 1) | end
    int barrels;     // Conduit.barrels, 1 for other link types
    int flags;       // link_flags of the link
    int kind;        // link_type | pump_type << 8
};

struct Net {
    int nN, nL, nP, nCurves, nShapeTbl, shapeTblLen;
    int nTrue;               // true conduits
    int lk_count[3];         // of which per conduit-function class LK_* (link_order is grouped by class)
    int nNonConduit;         // links handled by the ordered regulator pass
    int nPre;                // conduits that serve an outfall (link_pre_node >= 0); pre_links = their positions in
                             // link_order: a single model hands each to a thread of its own at the start of the phase
    int nOutfallLinks;
    int anyLossRate;         // some conduit can evaporate / seep (LF_HAS_LOSSRATE)
    swb_options opt;
    double crownCutoff;      // dynwave.c:159-160
#define X(T, name, kind) const T *name;
    SWB_DESC_ARRAYS(X)
#undef X
    // derived
    const int    *link_flags;
    const double *link_z1, *link_z2;       // Node[n].invertElev + Link.offset (dwflow.c:110-111)
    const double *xs_rcp_yfull;            // exact_rcp(xs_yfull): reciprocal for div_rcp, or 0 (swb_common.h)
    const double *cond_rcp_mod_length;     // exact_rcp(cond_mod_length)
    const int    *adj_start;               // nN+1: CSR node -> incident link ends
    const int    *adj;                     // (link << 1) | end, ordered true conduits first, then
                                           // other links, each ascending by link index (A.3)
    const AdjEntry *adj_packed;            // adj[] with the static link attributes of each entry
    const int    *adjq_start, *adjq;       // same incidence ordered by plain link index (quality)
    const int    *link_order;              // true conduits grouped by cross-section shape: tickets are
                                           // drawn in this order, so at any moment every warp of the
                                           // chip runs the same specialised conduit function (one
                                           // instruction-cache footprint instead of several)
    const int    *node_order;              // node-phase ticket order: outfalls and storage nodes first
    const int    *nc_links;                // non-true-conduit links in ascending index order
    const int    *outfall_link;            // per node: its (single) link, or -1
    const int    *outfall_slot;            // per node: row of State::o_ynorm / o_ycrit its link's update fills, or -1
    const int    *pre_links;               // [nPre] positions in link_order
    const int    *link_pre_node;           // per link: the outfall node whose normal / critical depth follows this
                                           // link's update (a true conduit that is that node's outfall_link), or -1
    const int    *outfall_nodes;           // outfall nodes in ascending index order
    int           nOutfallNodes;
    const int    *link_kernel;             // LK_*: which conduit function a true conduit runs
    const double *link_rows;               // [nTrue][LR_STRIDE]: packed static row per true conduit, in
                                           // link_order order (swb_dynwave.h: LR_*), 16-byte aligned
    const double *link_cols_d;             // [LR_DOUBLES][nTrue] / [LRI_INTS][nTrue]: the same attributes column-wise
    const int    *link_cols_i;             // in link_order order, for single models (CfCols)
    const double *culvert_params;          // [58][5] FHWA inlet-control coefficients (culvert.c:33)
    const double *road_tables;             // RT_TOTAL (x, y) pairs (roadway.c:42-69)
    const double *xs_tables;               // XT_TOTAL doubles (global copy of the shape tables)
    // every array above lives in ONE device allocation [arena, arena + arena_bytes): the launcher pins
    // that range in L2 (persisting access-policy window), so the ~100 dependent static loads of a
    // conduit / node update hit L2 instead of being evicted by the streaming ensemble state
    const void *arena;
    size_t arena_bytes;
};

// ---- dynamic fields: X(ctype, member name, field id, kind) ; kind N = per node, L = per link,
//      NP / LP = per node / link x pollutant -----------------------------------------------------
#define SWB_STATE_FIELDS(X) \
    X(double, n_depth, SWB_NODE_NEW_DEPTH, N) X(double, n_old_depth, SWB_NODE_OLD_DEPTH, N) \
    X(double, n_volume, SWB_NODE_NEW_VOLUME, N) X(double, n_old_volume, SWB_NODE_OLD_VOLUME, N) \
    X(double, n_latflow, SWB_NODE_NEW_LATFLOW, N) X(double, n_losses, SWB_NODE_LOSSES, N) \
    X(double, n_inflow, SWB_NODE_INFLOW, N) X(double, n_outflow, SWB_NODE_OUTFLOW, N) \
    X(double, n_overflow, SWB_NODE_OVERFLOW, N) X(double, n_old_net_inflow, SWB_NODE_OLD_NET_INFLOW, N) \
    X(double, n_new_surf_area, SWB_NODE_NEW_SURF_AREA, N) X(double, n_old_surf_area, SWB_NODE_OLD_SURF_AREA, N) \
    X(double, n_sumdqdh, SWB_NODE_SUMDQDH, N) X(double, n_dydt, SWB_NODE_DYDT, N) \
    X(unsigned char, n_converged, SWB_NODE_CONVERGED, N) X(double, n_stage, SWB_NODE_OUTFALL_STAGE, N) \
    X(double, n_evap_loss, SWB_NODE_STORAGE_EVAP_LOSS, N) X(double, n_exfil_loss, SWB_NODE_STORAGE_EXFIL_LOSS, N) \
    X(double, n_hrt, SWB_NODE_HRT, N) X(double, n_qual, SWB_NODE_NEW_QUAL, NP) \
    X(double, n_old_qual, SWB_NODE_OLD_QUAL, NP) \
    X(double, n_old_latflow, SWB_NODE_OLD_LATFLOW, N) X(double, n_old_inflow, SWB_NODE_OLD_INFLOW, N) \
    X(double, l_flow, SWB_LINK_NEW_FLOW, L) X(double, l_old_flow, SWB_LINK_OLD_FLOW, L) \
    X(double, l_depth, SWB_LINK_NEW_DEPTH, L) X(double, l_old_depth, SWB_LINK_OLD_DEPTH, L) \
    X(double, l_volume, SWB_LINK_NEW_VOLUME, L) X(double, l_old_volume, SWB_LINK_OLD_VOLUME, L) \
    X(double, l_setting, SWB_LINK_SETTING, L) X(double, l_target_setting, SWB_LINK_TARGET_SETTING, L) \
    X(double, l_dqdh, SWB_LINK_DQDH, L) X(double, l_froude, SWB_LINK_FROUDE, L) \
    X(unsigned char, l_flow_class, SWB_LINK_FLOW_CLASS, L) X(double, l_surf_area1, SWB_LINK_SURF_AREA1, L) \
    X(double, l_surf_area2, SWB_LINK_SURF_AREA2, L) X(unsigned char, l_bypassed, SWB_LINK_BYPASSED, L) \
    X(unsigned char, l_normal_flow, SWB_LINK_NORMAL_FLOW, L) X(unsigned char, l_inlet_control, SWB_LINK_INLET_CONTROL, L) \
    X(double, c_a1, SWB_COND_A1, L) X(double, c_a2, SWB_COND_A2, L) X(double, c_q1, SWB_COND_Q1, L) \
    X(unsigned char, c_full_state, SWB_COND_FULL_STATE, L) \
    X(unsigned char, c_cap_limited, SWB_COND_CAPACITY_LIMITED, L) X(double, c_evap_loss, SWB_COND_EVAP_LOSS, L) \
    X(double, c_seep_loss, SWB_COND_SEEP_LOSS, L) X(double, o_corif, SWB_ORIF_CORIF, L) \
    X(double, o_cweir, SWB_ORIF_CWEIR, L) X(double, o_hcrit, SWB_ORIF_HCRIT, L) \
    X(double, r_surf_area, SWB_REG_SURF_AREA, L) X(double, w_csurcharge, SWB_WEIR_CSURCHARGE, L) \
    X(double, l_qual, SWB_LINK_NEW_QUAL, LP) X(double, l_old_qual, SWB_LINK_OLD_QUAL, LP) \
    X(double, l_total_load, SWB_LINK_TOTAL_LOAD, LP)

enum { MB_EX_INFLOW = 0, MB_FLOODING, MB_OUTFLOW, MB_EVAP, MB_SEEP, MB_FLOW_TERMS,
       MBQ_EX_INFLOW = 0, MBQ_FLOODING, MBQ_OUTFLOW, MBQ_REACTED, MBQ_SEEP, MBQ_FINAL, MB_QUAL_TERMS };
enum { CTL_ANY_LEFT = 0,       // some member has not reached t_end: the step runs
       CTL_N_ALIVE,            // entries of State::alive for the trial about to run
       CTL_MARK_LO, CTL_MARK_HI,   // globaltimer at the start of the running phase (phase timers)
       SWB_CTL_WORDS = 8 };
enum { TP_PROLOGUE = 0, TP_LINKS, TP_REGULATORS, TP_NODES, TP_CONTROL, TP_EPILOGUE, TP_QUAL_NODES,
       TP_QUAL_LINKS, TP_NEXTDT, TP_HALO, TP_HALO_WAIT, SWB_N_PHASES = 12 };
#define SWB_MAX_TRIALS_CAP 32     // notConv bookkeeping rows (MaxTrials is 8 by default)
#define SWB_TICKETS_PER_TRIAL 8    // ticket counters per Picard trial (swb_engine.h: TK_*)

// One network partitioned over several GPUs (include/swmm_b200.h: swb_partition_desc).  Every rank
// owns a receive WINDOW in its own memory that the peers write into directly (peer memory over
// NVLink on the device, shared memory in the host emulation):
//   ctrl  [16] u64   [p] = last epoch rank p has published, [8] = error flag, [9] = epoch at launch end
//   red   [2][SWB_MAX_RANKS][HALO_RED] u64   per-parity reduction operands, one row per source rank
//   stage [n_recv][2][W] f64                  per-parity border values, W = max(2, pollutants)
// Values are double-buffered by the parity of the exchange epoch: a peer can only write parity e & 1
// again at epoch e + 2, i.e. after this rank has signalled e + 1, which it does after consuming e.
enum { HALO_CTRL_WORDS = 16, HALO_ERR = 8, HALO_EPOCH = 9, HALO_RED = 4 };
struct Halo {
    int rank, nRanks;            // nRanks <= 1: not partitioned
    int nOwnedN;                 // local nodes [0, nOwnedN) are updated here, the rest are ghosts
    int nSend, nRecv, W;
    const int *send_node, *send_rank, *send_slot, *recv_node, *link_owned;
    const int *node_order;       // Net::node_order restricted to the owned nodes
    const int *send_start;       // [nOwnedN + 1] CSR over the send entries (sorted by node): the node
                                 // phase publishes a border node the moment it has computed it
    unsigned long long *ctrl, *red;
    double *stage;
    unsigned long long *peer_ctrl[SWB_MAX_RANKS], *peer_red[SWB_MAX_RANKS];
    double *peer_stage[SWB_MAX_RANKS];
    unsigned long long timeout_ns;
    unsigned long long *wait_ns;  // = phase_ns + TP_HALO_WAIT: the part of TP_HALO spent spinning on peers
};

// Dynamic state is touched once per phase and a wide ensemble's state (2.5 GB for 512 members) never
// fits the 126 MB L2, so on the device every field access carries the streaming hint (ld/st.global.cs:
// first candidate for eviction).  What the hint protects is the kernel's own local-memory lines
// (register spills of the conduit update), which ARE re-read within microseconds: without it the
// state streams push them out to HBM (profiles/README.md round 2).  SPtr<T> is layout-identical to T*.
#ifndef SWB_STREAM_STATE
#define SWB_STREAM_STATE 0
#endif
#if defined(__CUDACC__) && SWB_STREAM_STATE
template <class T> struct SRef {
    T *p;
#ifdef __CUDA_ARCH__
    __host__ __device__ __forceinline__ operator T() const { return __ldcs(p); }
    __host__ __device__ __forceinline__ const SRef &operator=(T v) const { __stcs(p, v); return *this; }
#else
    __host__ __device__ __forceinline__ operator T() const { return *p; }
    __host__ __device__ __forceinline__ const SRef &operator=(T v) const { *p = v; return *this; }
#endif
    __host__ __device__ __forceinline__ const SRef &operator=(const SRef &o) const { return *this = (T)o; }
};
template <class T> struct SPtr {
    T *p;
    __host__ __device__ __forceinline__ SRef<T> operator[](size_t i) const { return SRef<T>{p + i}; }
    __host__ __device__ __forceinline__ operator T *() const { return p; }
};
#else
template <class T> using SPtr = T *;
#endif
// SWB_STREAM_STATE: 0 off, 1 every field, 2 link / conduit fields only (node fields are re-read within
// a step: quality mixing reads what the node phase has just written)
#if SWB_STREAM_STATE == 2
#define SWB_FIELD_PTR_N(T)  T *
#define SWB_FIELD_PTR_NP(T) T *
#else
#define SWB_FIELD_PTR_N(T)  SPtr<T>
#define SWB_FIELD_PTR_NP(T) SPtr<T>
#endif
#define SWB_FIELD_PTR_L(T)  SPtr<T>
#define SWB_FIELD_PTR_LP(T) SPtr<T>

struct State {
    int M;                       // members
    Halo halo;
#define X(T, name, id, kind) SWB_FIELD_PTR_##kind(T) name;
    SWB_STATE_FIELDS(X)
#undef X
    // per member
    double *dt;                  // routing step being taken
    double *var_step;            // VariableStep (dynwave.c:84); 0 before the first step
    double *o_ynorm, *o_ycrit;   // [nOutfallNodes][M] normal / critical depth of an outfall's conduit at its new flow,
                                 // written by the warp that updates that conduit (link phase), read by outfall_depth
    double *sim_time;            // elapsed simulated seconds (= time_ms / 1000, for reporting)
    double *time_ms;             // NewRoutingTime: elapsed ms, accumulated like routing.c:301-302
    double *evap_rate, *hydcon;  // per-member climate scalars
    int    *iters;               // iterations used by the last dynwave_execute
    long long *tot_iters, *tot_steps, *non_conv;
    int    *crit_node, *crit_link;
    unsigned long long *tmin_bits;  // Courant search scratch (ordered bits of a positive double)
    int    *alive;               // [M] staged kernels: ordered list of the members that iterate in the current trial
    int    *ctl;                 // [SWB_CTL_WORDS] staged kernels: CTL_* words written by the control kernels
    double *dt_cand;             // [4][threads] staged kernels: every thread's Courant candidates between the
                                 // search and the arg-min kernels (DtCand as four planes)
    int     dt_cand_stride;      // threads the planes were sized for
    int    *not_conv;            // [SWB_MAX_TRIALS_CAP][M] some node failed the head tolerance
    int    *done;                // member has reached t_end
    // mass-balance accumulators per member x pollutant (massbal.c:517-555)
    double *mb_reacted, *mb_seepage, *mb_final_storage;
    // routing totals per member (massbal.c:587-633): MB_FLOW_TERMS flow terms and, per pollutant,
    // MB_QUAL_TERMS mass terms.  mb_rate holds the rates of the step just taken (StepFlowTotals /
    // StepQualTotals, summed with atomics), mb_total the time integral; a step's rates are weighted
    // by half of its own step plus half of the NEXT one (routing.c:220,264), so the second half is
    // added by the next step's prologue (or by the host when it reads the totals).
    double *mb_rate, *mb_total;  // [(MB_FLOW_TERMS + MB_QUAL_TERMS * nP)][M]
    double *mb_dt_prev;          // [M] step whose rates are pending in mb_rate (0 = none)
    // per-object statistics (swb_stats.h); null until swb_enable_statistics
    double *stat_node;               // [(SWB_NS_PLANES + nP)][nN][M]
    double *stat_link;               // [SWB_LS_PLANES][nL][M]
    double *stat_sys;                // [SWB_SS_PLANES][M]
    double stat_report_start;        // statistics start at this elapsed simulated time (s)
    // device-side phase timers (ns, accumulated by thread 0 between grid barriers)
    unsigned long long *phase_ns;    // [SWB_N_PHASES]
    unsigned long long *tickets;     // [SWB_TICKETS_PER_TRIAL * SWB_MAX_TRIALS_CAP] work-distribution counters of one step
};

// pollutant-plane index: field[(p * nItems + item) * M + m]
#define SWB_IX(item, m, M)            ((size_t)(item) * (size_t)(M) + (size_t)(m))
#define SWB_IXP(p, item, n, m, M)     (((size_t)(p) * (size_t)(n) + (size_t)(item)) * (size_t)(M) + (size_t)(m))

} // namespace swb
#endif
