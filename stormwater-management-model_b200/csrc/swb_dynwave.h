// swb_dynwave.h -- per-(object, member) device functions of the dynamic-wave step.
//
//   K1  conduit_flow        <- dwflow_findConduitFlow + findSurfArea + getFlowClass + getWidth /
//                              getSlotWidth / getArea / getHydRad + checkNormalFlow +
//                              findLocalLosses (dwflow.c:57-686), link_getFroude (link.c:847),
//                              conduit_getLossRate (link.c:1334), link_setFlapGate (link.c:643),
//                              link_getYnorm / link_getYcrit (link.c:770-804)
//   K3  node_gather         <- initNodeStates + updateNodeFlows (dynwave.c:297-331, 528-589) as a
//                              fixed-order CSR gather (no atomics, reference summation order)
//   K4  outfall_depth       <- link_setOutfallDepth / node_setOutletDepth / outfall_setOutletDepth
//                              (link.c:728-766, node.c:532-558, 1413-1492)
//   K5  node_set_depth      <- setNodeDepth + getFloodedDepth (dynwave.c:636-795)
//   K7  link_step/node_step <- getLinkStep / getNodeStep (dynwave.c:836-921)
//
// Every function is a pure function of (static Net, State, object index, member index): the
// persistent kernel in swb_kernels.cu and the host emulation in tests/emul call exactly these.
#ifndef SWB_DYNWAVE_H
#define SWB_DYNWAVE_H

#include "swb_state.h"
#include "swb_xsect.h"
#include "swb_culvert.h"

namespace swb {

SWB_FI Xs load_xs(const Net &n, int j)
{
    Xs x;
    x.type = n.xs_type[j];
    x.yFull = n.xs_yfull[j]; x.wMax = n.xs_wmax[j]; x.ywMax = n.xs_ywmax[j];
    x.aFull = n.xs_afull[j]; x.rFull = n.xs_rfull[j]; x.sFull = n.xs_sfull[j];
    x.sMax = n.xs_smax[j];   x.yBot = n.xs_ybot[j];   x.aBot = n.xs_abot[j];
    x.sBot = n.xs_sbot[j];   x.rBot = n.xs_rbot[j];
    x.rYFull = n.xs_rcp_yfull[j];
    x.ntbl = 0; x.atbl = x.rtbl = x.wtbl = nullptr;
    int t = n.xs_table[j];
    if (t >= 0) {
        x.ntbl = n.shape_tbl_n[t];
        x.atbl = n.shape_area_tbl + (size_t)t * n.shapeTblLen;
        x.rtbl = n.shape_hrad_tbl + (size_t)t * n.shapeTblLen;
        x.wtbl = n.shape_width_tbl + (size_t)t * n.shapeTblLen;
    }
    return x;
}

// ---- dwflow.c:575-633 ---------------------------------------------------------------------------
// Sjoberg slot factor exp(-yNorm^2.4) (dwflow.c:620-633).  The caller only gets here for
// crownCutoff <= yNorm <= 1.78.  Host build: the reference's own expression.  Device: the exponent by
// a degree-16 polynomial in u = (yNorm - mid) / half on [0.98, 1.785] (tools/gen_sjoberg_poly.py)
// instead of libdevice's pow() -- 17 fused operations against ~110 plus a call, three times per
// surcharged conduit; exp(-p) is then within 7 ulp of the exact value (glibc's exp(-pow()) itself:
// 3.8 ulp, CUDA's: similar), i.e. the same class of deviation the device's libm already has.
#define SWB_SJOBERG_MID   0x1.61eb851eb851fp+0
#define SWB_SJOBERG_RHALF 0x1.3e032e1c9f019p+1
SWB_FI double dw_sjoberg_exponent_poly(double yNorm)
{
    const double u = (yNorm - SWB_SJOBERG_MID) * SWB_SJOBERG_RHALF;
    double p = -0x1.8bf5330e8f840p-41;
    p = fma(p, u, 0x1.acc159b30c145p-39);
    p = fma(p, u, -0x1.6f3b90eba18e4p-37);
    p = fma(p, u, 0x1.a1447fdb2f40bp-35);
    p = fma(p, u, -0x1.f1adf6c3c8bf8p-33);
    p = fma(p, u, 0x1.2a1cf2279644ep-30);
    p = fma(p, u, -0x1.721f57be79f3cp-28);
    p = fma(p, u, 0x1.e18d3a7a91c8ep-26);
    p = fma(p, u, -0x1.4c49fa880ecbep-23);
    p = fma(p, u, 0x1.f03c23a7c93e8p-21);
    p = fma(p, u, -0x1.9e4735d9d1908p-18);
    p = fma(p, u, 0x1.9a77c9469b612p-15);
    p = fma(p, u, -0x1.135d671fe64f8p-11);
    p = fma(p, u, 0x1.8a174c2adda9fp-7);
    p = fma(p, u, 0x1.3d41041976bc2p-2);
    p = fma(p, u, 0x1.852d91bfd53b0p+0);
    p = fma(p, u, 0x1.167ce5f9f7a7fp+1);
    return p;
}
SWB_FI double dw_sjoberg(double yNorm)
{
#if defined(__CUDA_ARCH__) && !defined(SWB_SJOBERG_LIBM)
    return exp(-dw_sjoberg_exponent_poly(yNorm));
#else
    return exp(-pow(yNorm, 2.4));
#endif
}
SWB_FI double dw_slot_width(const Net &n, const Xs &x, bool isOpen, double y)
{
    double yNorm = xs_ynorm(x, y);
    if (n.opt.surcharge_method != SWB_SLOT || isOpen || yNorm < n.crownCutoff) return 0.0;
    if (yNorm > 1.78) return 0.01 * x.wMax;
    return x.wMax * 0.5423 * dw_sjoberg(yNorm);
}
template <int S>
SWB_FI double dw_width(const Net &n, const Xs &x, bool isOpen, double y, const double *T)
{
    double wSlot = dw_slot_width(n, x, isOpen, y);
    if (wSlot > 0.0) return wSlot;
    if (xs_ynorm(x, y) >= n.crownCutoff && !isOpen) y = n.crownCutoff * x.yFull;
    return xw<S>(x, y, T);
}
// top width given the slot width already evaluated at the same depth (dwflow.c:592-605)
template <int S>
SWB_FI double dw_width_ws(const Net &n, const Xs &x, bool isOpen, double y, double wSlot, const double *T)
{
    if (wSlot > 0.0) return wSlot;
    if (xs_ynorm(x, y) >= n.crownCutoff && !isOpen) y = n.crownCutoff * x.yFull;
    return xw<S>(x, y, T);
}
template <int S>
SWB_FI double dw_area(const Xs &x, double y, double wSlot, const double *T)
{
    if (y >= x.yFull) return x.aFull + (y - x.yFull) * wSlot;
    return xa<S>(x, y, T);
}
template <int S>
SWB_FI double dw_hyd_rad(const Xs &x, double y, const double *T)
{
    if (y >= x.yFull) return x.rFull;
    return xr<S>(x, y, T);
}

// ---- link.c:847-871 (conduits only) -------------------------------------------------------------
template <int S>
SWB_FI double link_froude(const Xs &x, bool isOpen, double v, double y, const double *T)
{
    if (y <= SWB_FUDGE) return 0.0;
    if (!isOpen && x.yFull - y <= SWB_FUDGE) return 0.0;
    y = xa<S>(x, y, T) / xw<S>(x, y, T);
    return fabs(v) / sqrt(SWB_GRAVITY * y);
}

// ---- link.c:770-804 -----------------------------------------------------------------------------
SWB_NI double link_ynorm_vals(int flags, double qMax, double beta, const Xs &x, double q, const double *T)
{
    if (!(flags & LF_TRUE_CONDUIT)) return 0.0;
    q = fabs(q);
    if (q > qMax) q = qMax;
    if (q <= 0.0) return 0.0;
    double s = q / beta;
    double a = xs_a_of_s_ni(x, s, T);
    return xs_y_of_a_ni(x, a, T);
}
SWB_FI double link_ynorm(const Net &n, int j, const Xs &x, double q, const double *T)
{
    return link_ynorm_vals(n.link_flags[j], n.cond_q_max[j], n.cond_beta[j], x, q, T);
}

// ---- link.c:643-670 with the node tests folded into link_flags ----------------------------------
SWB_FI bool link_flap_closed(int flags, int direction, double q)
{
    if (flags & LF_HAS_FLAP) { if (q * (double)direction < 0.0) return true; }
    if (q < 0.0 && (flags & LF_N2_OUT_FLAP)) return true;
    if (q > 0.0 && (flags & LF_N1_OUT_FLAP)) return true;
    return false;
}

// ---- link.c:1334-1399 (DW branch) ---------------------------------------------------------------
template <int S>
SWB_FI double conduit_loss_rate(const State &s, int j, int m, const Xs &x, double length, double seepRate,
                                       bool isOpen, double dt, double evapRate, double hydcon,
                                       double &evapLoss, double &seepLoss, const double *T)
{
    size_t ix = SWB_IX(j, m, s.M);
    double depth = 0.5 * (s.l_old_depth[ix] + s.l_depth[ix]);
    double evapLossRate = 0.0, seepLossRate = 0.0, totalLossRate = 0.0;
    if (depth > SWB_FUDGE) {
        if (isOpen && evapRate > 0.0) {
            double topWidth = xw<S>(x, depth, T);
            evapLossRate = topWidth * length * evapRate;
        }
        if (seepRate > 0.0) {
            double width;
            if (x.type == XS_RECT_CLOSED) width = x.wMax;
            else {
                if (depth >= x.ywMax) depth = x.ywMax;
                width = xw<S>(x, depth, T);
            }
            seepLossRate = seepRate * width * length;
            seepLossRate *= hydcon;
        }
        totalLossRate = evapLossRate + seepLossRate;
        double q = s.l_volume[ix] / dt;
        if (totalLossRate > q) {
            evapLossRate = evapLossRate * q / totalLossRate;
            seepLossRate = seepLossRate * q / totalLossRate;
            totalLossRate = q;
        }
    }
    evapLoss = evapLossRate;
    seepLoss = seepLossRate;
    return totalLossRate;
}

// ---- dwflow.c:297-413 ---------------------------------------------------------------------------
struct FlowClassOut { int cls; double yC, yN, fasnh; };

// Normal and critical depth through the out-of-line solvers.  The cross section is taken BY VALUE
// on purpose: only this copy has its address passed to the opaque callees, so the caller's own Xs
// never escapes, stays in registers, and its compile-time shape keeps folding every geometry
// switch after the call (with a shared object the compiler had to reload x.type from the stack and
// re-instantiated all 26 shapes behind every later lookup).
template <class In>
SWB_FI void flow_class_depths(const In &in, Xs xc, double q, const double *T, double &yN, double &yC)
{
    yN = link_ynorm_vals(in.flags(), in.qMax(), in.beta(), xc, q, T);
    yC = xs_ycrit_ni(xc, q, T);
}

template <class In>
SWB_FI FlowClassOut dw_flow_class(const In &in, const Xs &x, int flags, double q,
                                         double h1, double h2, double y1, double y2,
                                         double depth1, double depth2, double yMidGuess,
                                         const double *T)
{
    FlowClassOut o;
    o.cls = SWB_SUBCRITICAL; o.fasnh = 1.0; o.yC = yMidGuess; o.yN = yMidGuess;
    double z1 = in.offset1(), z2 = in.offset2();
    if (flags & LF_N1_OUTFALL) z1 = SWB_MAX(0.0, (z1 - depth1));
    if (flags & LF_N2_OUTFALL) z2 = SWB_MAX(0.0, (z2 - depth2));

    if (y1 > SWB_FUDGE && y2 > SWB_FUDGE) {
        if (q < 0.0) {
            if (z1 > 0.0) {
                flow_class_depths(in, x, fabs(q), T, o.yN, o.yC);
                double ycMin = SWB_MIN(o.yN, o.yC);
                if (y1 < ycMin) o.cls = SWB_UP_CRITICAL;
            }
        } else {
            if (z2 > 0.0) {
                flow_class_depths(in, x, fabs(q), T, o.yN, o.yC);
                double ycMin = SWB_MIN(o.yN, o.yC);
                double ycMax = SWB_MAX(o.yN, o.yC);
                if (y2 < ycMin) o.cls = SWB_DN_CRITICAL;
                else if (y2 < ycMax) {
                    if (ycMax - ycMin < SWB_FUDGE) o.fasnh = 0.0;
                    else o.fasnh = (ycMax - y2) / (ycMax - ycMin);
                }
            }
        }
    }
    else if (y1 <= SWB_FUDGE && y2 <= SWB_FUDGE) o.cls = SWB_DRY;
    else if (y2 > SWB_FUDGE) {
        if (h2 < in.z1()) o.cls = SWB_UP_DRY;
        else if (z1 > 0.0) {
            flow_class_depths(in, x, fabs(q), T, o.yN, o.yC);
            o.cls = SWB_UP_CRITICAL;
        }
    }
    else {
        if (h1 < in.z2()) o.cls = SWB_DN_DRY;
        else if (z2 > 0.0) {
            flow_class_depths(in, x, fabs(q), T, o.yN, o.yC);
            o.cls = SWB_DN_CRITICAL;
        }
    }
    return o;
}

// ---- K1: dwflow.c:57-293 (with findSurfArea :417-550, checkNormalFlow :637-686) ------------------
// Where the per-(conduit, member) inputs of the update come from.  CfLoad reads the state arrays at the
// point of use (late loads keep the register budget of the persistent kernel); CfStaged reads this lane's
// column of a shared-memory stage that cp.async filled while the warp was computing its previous tile
// (swb_staged.cuh: sg_links_pf) -- the same values, hence the same bits.
// Static attributes of a true conduit packed into one 16-byte-aligned row (Net::link_rows, rows in
// link_order order): the staged link kernel copies a conduit's row into shared memory with one cp.async per
// lane while the warp works on its previous tile, so the update reads its ~40 static values with LDS at the
// point of use instead of ~40 dependent global loads (and keeps none of them live across the geometry).
enum { LR_YFULL = 0, LR_WMAX, LR_YWMAX, LR_AFULL, LR_RFULL, LR_SFULL, LR_SMAX, LR_YBOT, LR_ABOT, LR_SBOT, LR_RBOT,
       LR_RCP_YFULL, LR_Z1, LR_Z2, LR_OFFSET1, LR_OFFSET2, LR_INV1, LR_INV2, LR_LENGTH, LR_MOD_LENGTH,
       LR_RCP_MOD_LENGTH, LR_ROUGH, LR_BETA, LR_QMAX, LR_QLIMIT, LR_CLOSS_IN, LR_CLOSS_OUT, LR_CLOSS_AVG, LR_SEEP,
       LR_DOUBLES,                       // 29 doubles, then 10 ints two per slot
       LR_STRIDE = 36 };                 // doubles per row: 288 bytes = 18 lanes x 16 bytes
enum { LRI_FLAGS = 0, LRI_XS_TYPE, LRI_BARRELS, LRI_HAS_LOSSES, LRI_DIRECTION, LRI_CULVERT, LRI_NODE1, LRI_NODE2, LRI_INTS };
enum { CF_QLAST = 0, CF_DEPTH1, CF_DEPTH2, CF_SETTING, CF_AOLD, CF_OLDFLOW, CF_DT, CF_FIELDS };

struct CfLoad {
    static constexpr bool kStaged = false;
    const Net &n;
    const State &s;
    int j;
    size_t ix, ix1, ix2;
    SWB_FI double qLast()   const { return s.c_q1[ix]; }
    SWB_FI double depth1()  const { return s.n_depth[ix1]; }
    SWB_FI double depth2()  const { return s.n_depth[ix2]; }
    SWB_FI double setting() const { return s.l_setting[ix]; }
    SWB_FI double aOld()    const { return s.c_a2[ix]; }
    SWB_FI double oldFlow() const { return s.l_old_flow[ix]; }
    // static attributes
    SWB_FI Xs     xs()        const { return load_xs(n, j); }
    SWB_FI int    flags()     const { return n.link_flags[j]; }
    SWB_FI double inv1()      const { return n.node_invert[n.link_node1[j]]; }
    SWB_FI double inv2()      const { return n.node_invert[n.link_node2[j]]; }
    SWB_FI double z1()        const { return n.link_z1[j]; }
    SWB_FI double z2()        const { return n.link_z2[j]; }
    SWB_FI double offset1()   const { return n.link_offset1[j]; }
    SWB_FI double offset2()   const { return n.link_offset2[j]; }
    SWB_FI double length()    const { return n.cond_length[j]; }
    SWB_FI double modLength() const { return n.cond_mod_length[j]; }
    SWB_FI double rcpModLength() const { return n.cond_rcp_mod_length[j]; }
    SWB_FI double rough()     const { return n.cond_rough_factor[j]; }
    SWB_FI double beta()      const { return n.cond_beta[j]; }
    SWB_FI double qMax()      const { return n.cond_q_max[j]; }
    SWB_FI double qLimit()    const { return n.link_q_limit[j]; }
    SWB_FI double clossIn()   const { return n.link_closs_in[j]; }
    SWB_FI double clossOut()  const { return n.link_closs_out[j]; }
    SWB_FI double clossAvg()  const { return n.link_closs_avg[j]; }
    SWB_FI double seepRate()  const { return n.link_seep_rate[j]; }
    SWB_FI int    barrels()   const { return n.cond_barrels[j]; }
    SWB_FI int    hasLosses() const { return n.cond_has_losses[j]; }
    SWB_FI int    direction() const { return n.link_direction[j]; }
    SWB_FI int    culvert()   const { return n.xs_culvert[j]; }
};
// Single models (M == 1): a warp holds 32 DIFFERENT conduits, consecutive in link_order.  The same packed
// attributes stored column-wise in that order (Net::link_cols_d / link_cols_i) make every static load of the
// warp one contiguous 256-byte request; indexed by the original link number the shape-grouped ticket order
// reads the descriptor arrays with a stride of two or more and fetches twice the bytes.
struct CfCols {
    static constexpr bool kStaged = false;
    const Net &n;
    const State &s;
    int j, k;                    // link index, position in link_order
    size_t ix, ix1, ix2;
    SWB_FI double qLast()   const { return s.c_q1[ix]; }
    SWB_FI double depth1()  const { return s.n_depth[ix1]; }
    SWB_FI double depth2()  const { return s.n_depth[ix2]; }
    SWB_FI double setting() const { return s.l_setting[ix]; }
    SWB_FI double aOld()    const { return s.c_a2[ix]; }
    SWB_FI double oldFlow() const { return s.l_old_flow[ix]; }
    SWB_FI double rd(int f) const { return n.link_cols_d[(size_t)f * n.nTrue + k]; }
    SWB_FI int    ri(int f) const { return n.link_cols_i[(size_t)f * n.nTrue + k]; }
    SWB_FI Xs xs() const
    {
        Xs x;
        x.type = ri(LRI_XS_TYPE);
        x.yFull = rd(LR_YFULL); x.wMax = rd(LR_WMAX); x.ywMax = rd(LR_YWMAX);
        x.aFull = rd(LR_AFULL); x.rFull = rd(LR_RFULL); x.sFull = rd(LR_SFULL);
        x.sMax = rd(LR_SMAX);   x.yBot = rd(LR_YBOT);   x.aBot = rd(LR_ABOT);
        x.sBot = rd(LR_SBOT);   x.rBot = rd(LR_RBOT);   x.rYFull = rd(LR_RCP_YFULL);
        x.ntbl = 0; x.atbl = x.rtbl = x.wtbl = nullptr;
        return x;
    }
    SWB_FI int    flags()     const { return ri(LRI_FLAGS); }
    SWB_FI double inv1()      const { return rd(LR_INV1); }
    SWB_FI double inv2()      const { return rd(LR_INV2); }
    SWB_FI double z1()        const { return rd(LR_Z1); }
    SWB_FI double z2()        const { return rd(LR_Z2); }
    SWB_FI double offset1()   const { return rd(LR_OFFSET1); }
    SWB_FI double offset2()   const { return rd(LR_OFFSET2); }
    SWB_FI double length()    const { return rd(LR_LENGTH); }
    SWB_FI double modLength() const { return rd(LR_MOD_LENGTH); }
    SWB_FI double rcpModLength() const { return rd(LR_RCP_MOD_LENGTH); }
    SWB_FI double rough()     const { return rd(LR_ROUGH); }
    SWB_FI double beta()      const { return rd(LR_BETA); }
    SWB_FI double qMax()      const { return rd(LR_QMAX); }
    SWB_FI double qLimit()    const { return rd(LR_QLIMIT); }
    SWB_FI double clossIn()   const { return rd(LR_CLOSS_IN); }
    SWB_FI double clossOut()  const { return rd(LR_CLOSS_OUT); }
    SWB_FI double clossAvg()  const { return rd(LR_CLOSS_AVG); }
    SWB_FI double seepRate()  const { return rd(LR_SEEP); }
    SWB_FI int    barrels()   const { return ri(LRI_BARRELS); }
    SWB_FI int    hasLosses() const { return ri(LRI_HAS_LOSSES); }
    SWB_FI int    direction() const { return ri(LRI_DIRECTION); }
    SWB_FI int    culvert()   const { return ri(LRI_CULVERT); }
};
struct CfStaged {
    static constexpr bool kStaged = true;
    const double *b;             // &stage[lane]; field f of this lane is b[f * 32]
    const double *r;             // the conduit's static row in shared memory (LR_*)
    SWB_FI double qLast()   const { return b[CF_QLAST * 32]; }
    SWB_FI double depth1()  const { return b[CF_DEPTH1 * 32]; }
    SWB_FI double depth2()  const { return b[CF_DEPTH2 * 32]; }
    SWB_FI double setting() const { return b[CF_SETTING * 32]; }
    SWB_FI double aOld()    const { return b[CF_AOLD * 32]; }
    SWB_FI double oldFlow() const { return b[CF_OLDFLOW * 32]; }
    SWB_FI int ri(int k) const { return reinterpret_cast<const int *>(r + LR_DOUBLES)[k]; }
    // only the specialised, table-free shapes run staged: no per-object tables
    SWB_FI Xs xs() const
    {
        Xs x;
        x.type = ri(LRI_XS_TYPE);
        x.yFull = r[LR_YFULL]; x.wMax = r[LR_WMAX]; x.ywMax = r[LR_YWMAX];
        x.aFull = r[LR_AFULL]; x.rFull = r[LR_RFULL]; x.sFull = r[LR_SFULL];
        x.sMax = r[LR_SMAX];   x.yBot = r[LR_YBOT];   x.aBot = r[LR_ABOT];
        x.sBot = r[LR_SBOT];   x.rBot = r[LR_RBOT];   x.rYFull = r[LR_RCP_YFULL];
        x.ntbl = 0; x.atbl = x.rtbl = x.wtbl = nullptr;
        return x;
    }
    SWB_FI int    flags()     const { return ri(LRI_FLAGS); }
    SWB_FI double inv1()      const { return r[LR_INV1]; }
    SWB_FI double inv2()      const { return r[LR_INV2]; }
    SWB_FI double z1()        const { return r[LR_Z1]; }
    SWB_FI double z2()        const { return r[LR_Z2]; }
    SWB_FI double offset1()   const { return r[LR_OFFSET1]; }
    SWB_FI double offset2()   const { return r[LR_OFFSET2]; }
    SWB_FI double length()    const { return r[LR_LENGTH]; }
    SWB_FI double modLength() const { return r[LR_MOD_LENGTH]; }
    SWB_FI double rcpModLength() const { return r[LR_RCP_MOD_LENGTH]; }
    SWB_FI double rough()     const { return r[LR_ROUGH]; }
    SWB_FI double beta()      const { return r[LR_BETA]; }
    SWB_FI double qMax()      const { return r[LR_QMAX]; }
    SWB_FI double qLimit()    const { return r[LR_QLIMIT]; }
    SWB_FI double clossIn()   const { return r[LR_CLOSS_IN]; }
    SWB_FI double clossOut()  const { return r[LR_CLOSS_OUT]; }
    SWB_FI double clossAvg()  const { return r[LR_CLOSS_AVG]; }
    SWB_FI double seepRate()  const { return r[LR_SEEP]; }
    SWB_FI int    barrels()   const { return ri(LRI_BARRELS); }
    SWB_FI int    hasLosses() const { return ri(LRI_HAS_LOSSES); }
    SWB_FI int    direction() const { return ri(LRI_DIRECTION); }
    SWB_FI int    culvert()   const { return ri(LRI_CULVERT); }
};

template <int S, class In>
SWB_FI void conduit_flow_in(const Net &n, const State &s, int j, int m, int steps, double dt,
                            const double *T, const In &in)
{
    const int M = s.M;
    const size_t ix = SWB_IX(j, m, M);
    const int flags = in.flags();
    Xs x = in.xs();
    if constexpr (S >= 0) x.type = S;        // compile-time shape: every geometry switch folds
    // Staged inputs: the cross section is re-read from shared memory at the start of every block below, so
    // none of its 12 values has to stay in a register (or a spill slot) across the blocks in between.
#define SWB_XS_REFRESH() do { if constexpr (In::kStaged) { x = in.xs(); if constexpr (S >= 0) x.type = S; } } while (0)
    const bool isOpen = (flags & LF_OPEN_SHAPE) != 0;
    const bool slot = (n.opt.surcharge_method == SWB_SLOT);
    // Register budget (64 per thread at 32 warps/SM): values that are only needed late (old flow,
    // old area, barrels, true length, end-node depths for the dry-node test) are loaded late or
    // re-read instead of being kept live across the geometry evaluation.
    double qLast = in.qLast();
    double evapLoss = 0.0, seepLoss = 0.0;

    const double depth1 = in.depth1(), depth2 = in.depth2();
    const double inv1 = in.inv1(), inv2 = in.inv2();
    double z1 = in.z1(), z2 = in.z2();
    double h1 = depth1 + inv1, h2 = depth2 + inv2;
    h1 = SWB_MAX(h1, z1);
    h2 = SWB_MAX(h2, z2);
    double y1 = h1 - z1, y2 = h2 - z2;
    y1 = SWB_MAX(y1, SWB_FUDGE);
    y2 = SWB_MAX(y2, SWB_FUDGE);
    if (!slot) { y1 = SWB_MIN(y1, x.yFull); y2 = SWB_MIN(y2, x.yFull); }

    const double length = in.modLength(), rLength = in.rcpModLength();

    // --- findSurfArea (dwflow.c:417-550) on the previous iteration's flow
    int flowClass;
    double surfArea1 = 0.0, surfArea2 = 0.0;
    // Preissmann slot widths at the three depths: each is needed for a top width AND an area
    // (dwflow.c:143-152, 600), so it is evaluated once (pow + exp) and shared; < 0 = not yet
    double ws1 = -1.0, ws2 = -1.0, wsM = -1.0, widthMid = 0.0;
    bool widthMidPlain = false;          // widthMid == W(yMid) with no slot and no crown clamp
    {
        double fd1 = y1, fd2 = y2, fdMid, width1 = 0.0, width2 = 0.0;
        double normalDepth = (fd1 + fd2) / 2.0, criticalDepth = normalDepth, fasnh = 1.0;
        if (fd1 >= x.yFull && fd2 >= x.yFull) flowClass = SWB_SUBCRITICAL;
        else {
            FlowClassOut fc = dw_flow_class(in, x, flags, qLast, h1, h2, y1, y2, depth1, depth2,
                                            normalDepth, T);
            flowClass = fc.cls; criticalDepth = fc.yC; normalDepth = fc.yN; fasnh = fc.fasnh;
        }
        // depths per class (dwflow.c:465-545); the widths are pure functions of depth, so they are
        // evaluated once below instead of once per case (identical values, 3 call sites not 15)
        switch (flowClass) {
          case SWB_UP_CRITICAL:
            fd1 = criticalDepth;
            if (normalDepth < criticalDepth) fd1 = normalDepth;
            fd1 = SWB_MAX(fd1, SWB_FUDGE);
            h1 = z1 + fd1;
            break;
          case SWB_DN_CRITICAL:
            fd2 = criticalDepth;
            if (normalDepth < criticalDepth) fd2 = normalDepth;
            fd2 = SWB_MAX(fd2, SWB_FUDGE);
            h2 = z2 + fd2;
            break;
          case SWB_UP_DRY: fd1 = SWB_FUDGE; break;
          case SWB_DN_DRY: fd2 = SWB_FUDGE; break;
          default: break;
        }
        if (flowClass == SWB_DRY) {
            surfArea1 = SWB_FUDGE * length / 2.0;
            surfArea2 = surfArea1;
        } else {
            fdMid = 0.5 * (fd1 + fd2);
            if (fdMid < SWB_FUDGE) fdMid = SWB_FUDGE;
            wsM = dw_slot_width(n, x, isOpen, fdMid);
            widthMid = dw_width_ws<S>(n, x, isOpen, fdMid, wsM, T);
            widthMidPlain = !(wsM > 0.0) && !(xs_ynorm(x, fdMid) >= n.crownCutoff && !isOpen);
            if (flowClass != SWB_UP_CRITICAL) {
                ws1 = dw_slot_width(n, x, isOpen, fd1);
                width1 = dw_width_ws<S>(n, x, isOpen, fd1, ws1, T);
            }
            if (flowClass != SWB_DN_CRITICAL) {
                ws2 = dw_slot_width(n, x, isOpen, fd2);
                width2 = dw_width_ws<S>(n, x, isOpen, fd2, ws2, T);
            }
            switch (flowClass) {
              case SWB_SUBCRITICAL:
                surfArea1 = (width1 + widthMid) * length / 4.;
                surfArea2 = (widthMid + width2) * length / 4. * fasnh;
                break;
              case SWB_UP_CRITICAL:
                surfArea2 = (widthMid + width2) * length * 0.5;
                break;
              case SWB_DN_CRITICAL:
                surfArea1 = (width1 + widthMid) * length * 0.5;
                break;
              case SWB_UP_DRY:
                surfArea2 = (widthMid + width2) * length / 4.;
                if (in.offset1() <= 0.0) surfArea1 = (width1 + widthMid) * length / 4.;
                break;
              case SWB_DN_DRY:
                surfArea1 = (widthMid + width1) * length / 4.;
                if (in.offset2() <= 0.0) surfArea2 = (width2 + widthMid) * length / 4.;
                break;
            }
        }
        y1 = fd1; y2 = fd2;
    }
    s.l_surf_area1[ix] = surfArea1;
    s.l_surf_area2[ix] = surfArea2;

    // --- areas and hydraulic radii (dwflow.c:142-153)
    SWB_XS_REFRESH();
    if (ws1 < 0.0) ws1 = dw_slot_width(n, x, isOpen, y1);
    double a1 = dw_area<S>(x, y1, ws1, T);
    double r1 = dw_hyd_rad<S>(x, y1, T);
    if (ws2 < 0.0) ws2 = dw_slot_width(n, x, isOpen, y2);
    double a2 = dw_area<S>(x, y2, ws2, T);
    double yMid = 0.5 * (y1 + y2);       // == fdMid whenever widthMid was evaluated (y1, y2 >= FUDGE)
    if (wsM < 0.0) wsM = dw_slot_width(n, x, isOpen, yMid);
    double aMid = dw_area<S>(x, yMid, wsM, T);
    double rMid = dw_hyd_rad<S>(x, yMid, T);

    bool isFull = (y1 >= x.yFull && y2 >= x.yFull);

    // --- dry / closed exit (dwflow.c:165-180)
    const bool isClosed = (in.setting() == 0);
    const double barrels = (double)in.barrels();
    if (flowClass == SWB_DRY || flowClass == SWB_UP_DRY || flowClass == SWB_DN_DRY || isClosed ||
        aMid <= SWB_FUDGE) {
        const double trueLength = in.length();
        double a1s = 0.5 * (a1 + a2);
        s.c_a1[ix] = a1s;
        s.c_q1[ix] = 0.0;
        s.l_dqdh[ix] = div_by(SWB_GRAVITY * dt * aMid, length, rLength) * barrels;
        s.l_froude[ix] = 0.0;
        s.l_depth[ix] = SWB_MIN(yMid, x.yFull);
        s.l_volume[ix] = a1s * trueLength * barrels;
        s.l_flow[ix] = 0.0;
        s.l_flow_class[ix] = (unsigned char)flowClass;
        if (flags & LF_HAS_LOSSRATE) { s.c_evap_loss[ix] = 0.0; s.c_seep_loss[ix] = 0.0; }
        return;
    }

    // --- velocity, Froude number, inertial damping (dwflow.c:183-208)
    SWB_XS_REFRESH();
    double v = qLast / aMid;
    if (fabs(v) > SWB_MAXVELOCITY) v = SWB_MAXVELOCITY * SWB_SGN(qLast);
    // link_getFroude (link.c:847-871); A(yMid) and W(yMid) are reused when they are the very values
    // already computed above (depth below full, no slot, no crown clamp)
    double froude;
    if (yMid <= SWB_FUDGE) froude = 0.0;
    else if (!isOpen && x.yFull - yMid <= SWB_FUDGE) froude = 0.0;
    else {
        double aF = (yMid < x.yFull) ? aMid : xa<S>(x, yMid, T);
        double wF = widthMidPlain ? widthMid : xw<S>(x, yMid, T);
        froude = fabs(v) / sqrt(SWB_GRAVITY * (aF / wF));
    }
    if (flowClass == SWB_SUBCRITICAL && froude > 1.0) flowClass = SWB_SUPCRITICAL;
    double sigma;
    if      (froude <= 0.5) sigma = 1.0;
    else if (froude >= 1.0) sigma = 0.0;
    else    sigma = 2.0 * (1.0 - froude);
    double rho = 1.0;
    if (!isFull && qLast > 0.0 && h1 >= h2) rho = sigma;
    double aWtd = a1 + (aMid - a1) * rho;
    double rWtd = r1 + (rMid - r1) * rho;
    if      (n.opt.inert_damping == SWB_NO_DAMPING)   sigma = 1.0;
    else if (n.opt.inert_damping == SWB_FULL_DAMPING) sigma = 0.0;
    if (isFull && !isOpen) sigma = 0.0;

    // --- momentum terms (dwflow.c:210-236)
    const double trueLength = in.length();
    double aOld = in.aOld();
    aOld = SWB_MAX(aOld, SWB_FUDGE);
    const double qOld = in.oldFlow() / barrels;
    double dq1;
    if (S < 0 && x.type == XS_FORCE_MAIN && isFull)
         dq1 = dt * forcemain_fric_slope(n, x, fabs(v), rMid);
    else dq1 = dt * in.rough() / pow(rWtd, 1.33333) * fabs(v);
    double dq2 = div_by(dt * SWB_GRAVITY * aWtd * (h2 - h1), length, rLength);
    double dq3 = 0.0, dq4 = 0.0;
    if (sigma > 0.0) {
        dq3 = 2.0 * v * (aMid - aOld) * sigma;
        dq4 = div_by(dt * v * v * (a2 - a1), length, rLength) * sigma;
    }
    double dq5 = 0.0;
    if (in.hasLosses()) {
        double losses = 0.0, qa = fabs(qLast);
        if (a1 > SWB_FUDGE)   losses += in.clossIn()  * (qa / a1);
        if (a2 > SWB_FUDGE)   losses += in.clossOut() * (qa / a2);
        if (aMid > SWB_FUDGE) losses += in.clossAvg() * (qa / aMid);
        dq5 = div_by(losses / 2.0, length, rLength) * dt;
    }
    double dq6 = 0.0;
    if (flags & LF_HAS_LOSSRATE) {
        double lossRate = conduit_loss_rate<S>(s, j, m, x, trueLength, in.seepRate(), isOpen, dt, s.evap_rate[m], s.hydcon[m],
                                            evapLoss, seepLoss, T);
        dq6 = lossRate * 2.5 * dt * v / trueLength;
    }

    double denom = 1.0 + dq1 + dq5;
    double q = (qOld - dq2 + dq3 + dq4 + dq6) / denom;
    double dqdh = div_by(1.0 / denom * SWB_GRAVITY * dt * aWtd, length, rLength) * barrels;

    // --- flow limitations (dwflow.c:245-259); culvert-coded links always run the generic instance
    SWB_XS_REFRESH();
    unsigned char normalFlow = 0, inletControl = 0;
    const bool hasCulvert = (S < 0) && in.culvert() > 0;
    if (q > 0.0) {
        if (hasCulvert && !isFull) q = culvert_inflow(n, x, j, q, h1, dqdh, inletControl, T);
        else if (n.opt.normal_flow_ltd != SWB_NF_NEITHER && y1 < x.yFull &&
            (flowClass == SWB_SUBCRITICAL || flowClass == SWB_SUPCRITICAL)) {
            // checkNormalFlow (dwflow.c:637-686)
            bool check = false;
            bool hasOutfall = (flags & (LF_N1_OUTFALL | LF_N2_OUTFALL)) != 0;
            int nfl = n.opt.normal_flow_ltd;
            if (nfl == SWB_NF_SLOPE || nfl == SWB_NF_BOTH || hasOutfall) { if (y1 < y2) check = true; }
            if (!check && (nfl == SWB_NF_FROUDE || nfl == SWB_NF_BOTH) && !hasOutfall) {
                if (y1 > SWB_FUDGE && y2 > SWB_FUDGE) {
                    double f1 = link_froude<S>(x, isOpen, q / a1, y1, T);   // y1 < yFull here
                    if (f1 >= 1.0) check = true;
                }
            }
            if (check) {
                double qNorm = in.beta() * a1 * pow(r1, 2. / 3.);
                if (qNorm < q) { normalFlow = 1; q = qNorm; }
            }
        }
    }

    // --- under-relaxation, limits, flap gates, dry nodes (dwflow.c:261-281)
    if (steps > 0) {
        q = (1.0 - SWB_OMEGA) * qLast + SWB_OMEGA * q;
        if (q * qLast < 0.0) q = 0.001 * SWB_SGN(q);
    }
    double qLimit = in.qLimit();
    if (qLimit > 0.0) { if (fabs(q) > qLimit) q = SWB_SGN(q) * qLimit; }
    if (link_flap_closed(flags, in.direction(), q)) q = 0.0;
    if (q >  SWB_FUDGE && in.depth1() <= SWB_FUDGE) q =  SWB_FUDGE;
    if (q < -SWB_FUDGE && in.depth2() <= SWB_FUDGE) q = -SWB_FUDGE;

    // --- save (dwflow.c:283-292)
    s.c_a1[ix] = aMid;
    s.c_q1[ix] = q;                       // Conduit.q2 == q1 under dynamic wave (dwflow.c:285-286)
    s.l_depth[ix] = SWB_MIN(yMid, x.yFull);
    double aAvg = (a1 + a2) / 2.0;
    unsigned char fullState = 0;
    if (a1 >= x.aFull) fullState = (a2 >= x.aFull) ? SWB_ALL_FULL : SWB_UP_FULL;
    else if (a2 >= x.aFull) fullState = SWB_DN_FULL;
    s.c_full_state[ix] = fullState;
    s.l_volume[ix] = aAvg * trueLength * barrels;
    s.l_flow[ix] = q * barrels;
    s.l_dqdh[ix] = dqdh;
    s.l_froude[ix] = froude;
    s.l_flow_class[ix] = (unsigned char)flowClass;
    s.l_normal_flow[ix] = normalFlow;
    if (hasCulvert) s.l_inlet_control[ix] = inletControl;     // 0 for ever on every other link
    if (flags & LF_HAS_LOSSRATE) { s.c_evap_loss[ix] = evapLoss; s.c_seep_loss[ix] = seepLoss; }
#undef SWB_XS_REFRESH
}

template <int S>
SWB_FI void conduit_flow(const Net &n, const State &s, int j, int m, int steps, double dt,
                         const double *T)
{
    const CfLoad in = { n, s, j, SWB_IX(j, m, s.M), SWB_IX(n.link_node1[j], m, s.M), SWB_IX(n.link_node2[j], m, s.M) };
    conduit_flow_in<S>(n, s, j, m, steps, dt, T, in);
}

// Real (non-inlined) entry points: one compact function per specialised shape plus the generic one.
// A warp holds one link, so the dispatch in conduit_update is warp-uniform.
#ifdef SWB_CONDUIT_INLINE
#define SWB_CF SWB_FI
#else
#define SWB_CF SWB_NI
#endif
SWB_CF void conduit_flow_circular(const Net &n, const State &s, int j, int m, int steps, double dt, const double *T)
{ conduit_flow<XS_CIRCULAR>(n, s, j, m, steps, dt, T); }
SWB_CF void conduit_flow_rect_closed(const Net &n, const State &s, int j, int m, int steps, double dt, const double *T)
{ conduit_flow<XS_RECT_CLOSED>(n, s, j, m, steps, dt, T); }
SWB_NI void conduit_flow_generic(const Net &n, const State &s, int j, int m, int steps, double dt, const double *T)
{ conduit_flow<-1>(n, s, j, m, steps, dt, T); }
// the two table-free shapes of a single model, statics read column-wise (k = position in link_order)
template <int S>
SWB_NI void conduit_flow_cols(const Net &n, const State &s, int j, int k, int n1, int n2, int m, int steps, double dt, const double *T)
{
    const CfCols in = { n, s, j, k, SWB_IX(j, m, s.M), SWB_IX(n1, m, s.M), SWB_IX(n2, m, s.M) };
    conduit_flow_in<S>(n, s, j, m, steps, dt, T, in);
}
SWB_FI void conduit_update_cols(const Net &n, const State &s, int j, int k, int n1, int n2, int m, int steps, double dt, const double *T)
{
    // link_order is grouped by conduit-function class: the class follows from the position
    if (k < n.lk_count[0]) conduit_flow_cols<XS_CIRCULAR>(n, s, j, k, n1, n2, m, steps, dt, T);
    else if (k < n.lk_count[0] + n.lk_count[1]) conduit_flow_cols<XS_RECT_CLOSED>(n, s, j, k, n1, n2, m, steps, dt, T);
    else conduit_flow_generic(n, s, j, m, steps, dt, T);
}
SWB_FI void conduit_update(const Net &n, const State &s, int j, int m, int steps, double dt, const double *T)
{
    switch (n.link_kernel[j]) {
      case LK_CIRCULAR:    conduit_flow_circular(n, s, j, m, steps, dt, T); break;
      case LK_RECT_CLOSED: conduit_flow_rect_closed(n, s, j, m, steps, dt, T); break;
      default:             conduit_flow_generic(n, s, j, m, steps, dt, T);
    }
}

// ---- node.c:362-396, 562-585, 930-1018 -----------------------------------------------------------
SWB_HD double curve_lookup_ex(const Net &n, int c, double x)          // table.c:467-500
{
    int i0 = n.curve_start[c], i1 = n.curve_start[c + 1];
    if (i1 <= i0) return 0.0;
    double x1 = n.curve_x[i0], y1 = n.curve_y[i0], x2, y2, sl = 0.0;
    if (x <= x1) { if (x1 > 0.0) return x / x1 * y1; else return y1; }
    for (int i = i0 + 1; i < i1; i++) {
        x2 = n.curve_x[i]; y2 = n.curve_y[i];
        if (x2 != x1) sl = (y2 - y1) / (x2 - x1);
        if (x <= x2) {
            double dx = x2 - x1;
            if (fabs(dx) < 1.0e-20) return (y1 + y2) / 2.;
            return y1 + (x - x1) * (y2 - y1) / dx;
        }
        x1 = x2; y1 = y2;
    }
    if (sl < 0.0) sl = 0.0;
    return y1 + sl * (x - x1);
}

SWB_HD double tbl_interp(double x, double x1, double y1, double x2, double y2)  // table.c:51
{
    double dx = x2 - x1;
    if (fabs(dx) < 1.0e-20) return (y1 + y2) / 2.;
    return y1 + (x - x1) * (y2 - y1) / dx;
}

SWB_HD double curve_storage_volume(const Net &n, int c, double x)    // table.c:590-648
{
    int i0 = n.curve_start[c], i1 = n.curve_start[c + 1];
    if (i1 <= i0) return 0.0;
    double a, a1, x1, v = 0.0, dx = 0.0, dy = 0.0, sl;
    x1 = n.curve_x[i0]; a1 = n.curve_y[i0];
    if (x <= x1) {
        if (x1 < 1.e-6) return 0.0;
        return (a1 / x1) * x * x / 2.0;
    }
    for (int i = i0 + 1; i < i1; i++) {
        double ex = n.curve_x[i], ey = n.curve_y[i];
        if (ex >= x) {
            a = tbl_interp(x, x1, a1, ex, ey);
            return v + (a1 + a) / 2.0 * (x - x1);
        }
        dx = ex - x1; dy = ey - a1;
        v = v + (a1 + ey) / 2.0 * dx;
        x1 = ex; a1 = ey;
    }
    if (dx > 1.0e-6) {
        sl = dy / dx;
        a = a1 + sl * (x - x1);
        if (a < 0.0) v = v - a1 * a1 / sl / 2.0;
        else v = v + (a1 + a) / 2.0 * (x - x1);
    }
    return v;
}

SWB_HD double storage_surf_area(const Net &n, int i, double d)
{
    double area = 0.0;
    const double ucfL = n.opt.ucf_length;
    switch (n.storage_shape[i]) {
      case 0: { int c = n.storage_curve[i]; if (c >= 0) area = curve_lookup_ex(n, c, d * ucfL); break; }
      case 1: area = n.storage_a0[i] + n.storage_a1[i] * pow(d * ucfL, n.storage_a2[i]); break;
      case 2: case 3: case 4: case 5:
        d *= ucfL;
        area = n.storage_a0[i] + d * (n.storage_a1[i] + d * n.storage_a2[i]);
        break;
      default: return 0.0;
    }
    return area / ucfL / ucfL;
}

SWB_HD double storage_volume(const Net &n, int i, double d)
{
    if (d == 0.0) return 0.0;
    if (d >= n.node_full_depth[i] && n.node_full_volume[i] > 0.0) return n.node_full_volume[i];
    const double ucfL = n.opt.ucf_length, ucfV = n.opt.ucf_volume;
    switch (n.storage_shape[i]) {
      case 0: { int c = n.storage_curve[i];
                if (c >= 0) return curve_storage_volume(n, c, d * ucfL) / ucfV;
                return 0.0; }
      case 1: { d *= ucfL;
                double nn = n.storage_a2[i] + 1.0;
                double v = (n.storage_a0[i] * d) + n.storage_a1[i] / nn * pow(d, nn);
                return v / ucfV; }
      case 2: case 3: case 4: case 5: {
                d *= ucfL;
                double v = d * (n.storage_a0[i] + d * (n.storage_a1[i] / 2.0 + d * n.storage_a2[i] / 3.0));
                return v / ucfV; }
      default: return 0.0;
    }
}

SWB_FI double node_surf_area(const Net &n, int i, double d)
{
    return n.node_type[i] == SWB_STORAGE ? storage_surf_area(n, i, d) : 0.0;
}
SWB_FI double node_ponded_area(const Net &n, int i, double d)        // node.c:562-585
{
    if (d <= n.node_full_depth[i] || n.node_ponded_area[i] == 0.0) return node_surf_area(n, i, d);
    double a = n.node_ponded_area[i];
    if (a <= 0.0) a = node_surf_area(n, i, n.node_full_depth[i]);
    return a;
}
SWB_FI double node_volume(const Net &n, int i, double d)             // node.c:345-358
{
    if (n.node_type[i] == SWB_STORAGE) return storage_volume(n, i, d);
    if (n.node_full_depth[i] > 0.0) return n.node_full_volume[i] * (d / n.node_full_depth[i]);
    return 0.0;
}

// ---- K3: initNodeStates + updateNodeFlows over the node's incidence list ------------------------
// Accumulates in the reference's order: lateral flow / losses first, then the incident link ends in
// CSR order [first, last).  Called once for the true-conduit segment and, for networks with
// regulators, link by link from the ordered pass.
struct NodeAcc { double inflow, outflow, surfArea, sumdqdh; };

SWB_FI NodeAcc node_init_acc(const Net &n, const State &s, int i, int m)
{
    NodeAcc a;
    size_t ix = SWB_IX(i, m, s.M);
    double depth = s.n_depth[ix];
    a.surfArea = n.opt.allow_ponding ? node_ponded_area(n, i, depth) : node_surf_area(n, i, depth);
    a.inflow = 0.0;
    a.outflow = s.n_losses[ix];
    double lat = s.n_latflow[ix];
    if (lat >= 0.0) a.inflow += lat; else a.outflow -= lat;
    a.sumdqdh = 0.0;
    return a;
}

// contribution of link j (end = 0: node is node1 / upstream, 1: node2) -- dynwave.c:528-589
SWB_FI void node_add_link_end(const Net &n, const State &s, int j, int end, int m, NodeAcc &a)
{
    size_t ix = SWB_IX(j, m, s.M);
    int flags = n.link_flags[j];
    double q = s.l_flow[ix];
    if (q >= 0.0) { if (end == 0) a.outflow += q; else a.inflow += q; }
    else          { if (end == 0) a.inflow -= q;  else a.outflow -= q; }
    int barrels = 1;
    if (n.link_type[j] == SWB_CONDUIT) {
        barrels = n.cond_barrels[j];
        if (flags & LF_HAS_LOSSRATE) {
            double lossRate = (s.c_evap_loss[ix] + s.c_seep_loss[ix]) * barrels;
            if (lossRate > 0.0) {
                bool o1 = (flags & LF_N1_OUTFALL) != 0, o2 = (flags & LF_N2_OUTFALL) != 0;
                if (!o1 && !o2) lossRate /= 2.0;
                if (end == 0 ? !o1 : !o2) a.outflow += lossRate;
            }
        }
    }
    a.surfArea += (end == 0 ? s.l_surf_area1[ix] : s.l_surf_area2[ix]) * barrels;
    double dqdh = s.l_dqdh[ix];
    if (end == 0) a.sumdqdh += dqdh;
    else if (n.link_type[j] == SWB_PUMP) { if (n.pump_type[j] != 3 /*TYPE4_PUMP*/) a.sumdqdh += dqdh; }
    else a.sumdqdh += dqdh;
}

// The same contribution split into its loads and its arithmetic, so that a caller can issue the
// loads of several link ends before consuming any of them (one DRAM round trip per node instead of one
// per link end); the accumulation order, hence every bit of the sums, is unchanged.  All static
// attributes come from the packed incidence entry.
struct LinkEndData { double q, loss, sa, dqdh; };
SWB_FI LinkEndData node_load_link_end(const State &s, const AdjEntry &a, int m)
{
    LinkEndData d;
    const int j = a.je >> 1;
    const size_t ix = SWB_IX(j, m, s.M);
    d.q = s.l_flow[ix];
    d.sa = ((a.je & 1) == 0 ? s.l_surf_area1[ix] : s.l_surf_area2[ix]);
    d.dqdh = s.l_dqdh[ix];
    d.loss = 0.0;
    if ((a.kind & 0xff) == SWB_CONDUIT && (a.flags & LF_HAS_LOSSRATE)) d.loss = s.c_evap_loss[ix] + s.c_seep_loss[ix];
    return d;
}
SWB_FI void node_apply_link_end(const LinkEndData &d, const AdjEntry &a, NodeAcc &acc)
{
    const int end = a.je & 1, type = a.kind & 0xff;
    const double q = d.q;
    if (q >= 0.0) { if (end == 0) acc.outflow += q; else acc.inflow += q; }
    else          { if (end == 0) acc.inflow -= q;  else acc.outflow -= q; }
    if (type == SWB_CONDUIT && (a.flags & LF_HAS_LOSSRATE)) {
        double lossRate = d.loss * a.barrels;
        if (lossRate > 0.0) {
            bool o1 = (a.flags & LF_N1_OUTFALL) != 0, o2 = (a.flags & LF_N2_OUTFALL) != 0;
            if (!o1 && !o2) lossRate /= 2.0;
            if (end == 0 ? !o1 : !o2) acc.outflow += lossRate;
        }
    }
    acc.surfArea += d.sa * a.barrels;
    if (end == 0) acc.sumdqdh += d.dqdh;
    else if (type == SWB_PUMP) { if ((a.kind >> 8) != 3 /*TYPE4_PUMP*/) acc.sumdqdh += d.dqdh; }
    else acc.sumdqdh += d.dqdh;
}
// entries [e0, e1) of the packed incidence list, four link ends in flight at a time (named scalars,
// not arrays: conditionally initialised arrays end up in local memory)
SWB_FI void node_gather(const Net &n, const State &s, int e0, int e1, int m, NodeAcc &acc)
{
    const AdjEntry none = {0, 0, 0, -1};
    for (int e = e0; e < e1; e += 4) {
        const int cnt = e1 - e;
        const AdjEntry a0 = n.adj_packed[e];
        const AdjEntry a1 = cnt > 1 ? n.adj_packed[e + 1] : none;
        const AdjEntry a2 = cnt > 2 ? n.adj_packed[e + 2] : none;
        const AdjEntry a3 = cnt > 3 ? n.adj_packed[e + 3] : none;
        LinkEndData d0 = node_load_link_end(s, a0, m), d1 = {0.0, 0.0, 0.0, 0.0}, d2 = d1, d3 = d1;
        if (cnt > 1) d1 = node_load_link_end(s, a1, m);
        if (cnt > 2) d2 = node_load_link_end(s, a2, m);
        if (cnt > 3) d3 = node_load_link_end(s, a3, m);
        node_apply_link_end(d0, a0, acc);
        if (cnt > 1) node_apply_link_end(d1, a1, acc);
        if (cnt > 2) node_apply_link_end(d2, a2, acc);
        if (cnt > 3) node_apply_link_end(d3, a3, acc);
    }
}

// link_setOutfallDepth's normal / critical depth (link.c:728-766) for the outfall served by true conduit j, at
// the flow the update of j has just stored; called by whoever updated the conduit (flags says it ends at an outfall)
SWB_NI void outfall_precompute(const Net &n, const State &s, int j, int m, const double *T)
{
    const int i = n.link_pre_node[j];
    if (i < 0) return;
    const size_t io = (size_t)n.outfall_slot[i] * s.M + m;
    Xs x = load_xs(n, j);
    double q = fabs(s.l_flow[SWB_IX(j, m, s.M)] / n.cond_barrels[j]);
    s.o_ynorm[io] = link_ynorm(n, j, x, q, T);
    s.o_ycrit[io] = xs_ycrit_ni(x, q, T);
}

// ---- K4: outfall boundary depth (link.c:728-766, node.c:1413-1492) ------------------------------
SWB_NI void outfall_depth(const Net &n, const State &s, int i, int m, const double *T)
{
    int j = n.outfall_link[i];
    if (j < 0) return;
    size_t ixl = SWB_IX(j, m, s.M), ixn = SWB_IX(i, m, s.M);
    double z = (n.link_node2[j] == i) ? n.link_offset2[j] : n.link_offset1[j];
    double yNorm = 0.0, yCrit = 0.0;
    if (n.link_type[j] == SWB_CONDUIT) {
        const int slot = n.outfall_slot[i];
        if (slot >= 0) {             // computed by the link phase right after this conduit's update
            yNorm = s.o_ynorm[(size_t)slot * s.M + m];
            yCrit = s.o_ycrit[(size_t)slot * s.M + m];
        } else {
            Xs x = load_xs(n, j);
            double q = fabs(s.l_flow[ixl] / n.cond_barrels[j]);
            yNorm = link_ynorm(n, j, x, q, T);
            yCrit = xs_ycrit_ni(x, q, T);
        }
    }
    double yNew;
    switch (n.outfall_type[i]) {
      case SWB_FREE_OUTFALL:
        if (z > 0.0) yNew = 0.0; else yNew = SWB_MIN(yNorm, yCrit);
        break;
      case SWB_NORMAL_OUTFALL:
        if (z > 0.0) yNew = 0.0; else yNew = yNorm;
        break;
      default: {
        double stage = s.n_stage[ixn];
        double inv = n.node_invert[i];
        yCrit = SWB_MIN(yCrit, yNorm);
        if (yCrit + z + inv < stage) yNew = stage - inv;
        else if (z > 0.0) {
            if (stage < inv + z) yNew = SWB_MAX(0.0, (stage - inv));
            else yNew = z + yCrit;
        }
        else yNew = yCrit;
      }
    }
    s.n_depth[ixn] = yNew;
}

// ---- K5: setNodeDepth + getFloodedDepth (dynwave.c:636-795); returns the converged flag ---------
struct NodeOld { double yOld, oldNetInflow; };      // loaded by the caller before the link gather
SWB_FI NodeOld node_load_old(const State &s, int i, int m)
{
    const size_t ix = SWB_IX(i, m, s.M);
    NodeOld o = { s.n_old_depth[ix], s.n_old_net_inflow[ix] };
    return o;
}
SWB_FI bool node_set_depth(const Net &n, const State &s, int i, int m, int steps, double dt,
                                  const NodeAcc &acc, const NodeOld &old)
{
    const size_t ix = SWB_IX(i, m, s.M);
    const double fullDepth = n.node_full_depth[i];
    const bool canPond = (n.opt.allow_ponding && n.node_ponded_area[i] > 0.0);
    const double yLast = s.n_depth[ix];
    const bool isPonded = (canPond && yLast > fullDepth);
    const double yCrown = n.node_crown_elev[i] - n.node_invert[i];
    const double yOld = old.yOld;
    double overflow = 0.0, newVolume;
    double surfArea = acc.surfArea;
    surfArea = SWB_MAX(surfArea, n.opt.min_surf_area);
    double dQ = acc.inflow - acc.outflow;
    double dV = 0.5 * (old.oldNetInflow + dQ) * dt;
    bool isSurcharged = false;
    double yNew, dy;

    if (n.opt.surcharge_method == SWB_EXTRAN) {
        if (isPonded) isSurcharged = false;
        else if (n.node_type[i] == SWB_STORAGE)
            isSurcharged = (n.node_sur_depth[i] > 0.0 && yLast > fullDepth);
        else isSurcharged = (yCrown > 0.0 && yLast > yCrown);
    }
    if (!isSurcharged) {
        dy = dV / surfArea;
        yNew = yOld + dy;
        if (!isPonded) s.n_old_surf_area[ix] = surfArea;
        if (steps > 0) yNew = (1.0 - SWB_OMEGA) * yLast + SWB_OMEGA * yNew;
        if (isPonded && yNew < fullDepth) yNew = fullDepth - SWB_FUDGE;
    } else {
        double corr = 1.0;
        if (n.node_degree[i] < 0) corr = 0.6;
        double denom = acc.sumdqdh;
        if (yLast < 1.25 * yCrown) {
            double f = (yLast - yCrown) / yCrown;
            denom += (s.n_old_surf_area[ix] / dt - acc.sumdqdh) * exp(-15.0 * f);
        }
        if (denom == 0.0) dy = 0.0;
        else dy = corr * dQ / denom;
        yNew = yLast + dy;
        if (yNew < yCrown) yNew = yCrown - SWB_FUDGE;
        if (canPond && yNew > fullDepth) yNew = fullDepth + SWB_FUDGE;
    }
    if (yNew < 0) yNew = 0.0;
    double yMax = fullDepth;
    if (!canPond) yMax += n.node_sur_depth[i];
    if (yNew > yMax) {
        // getFloodedDepth (dynwave.c:766-795)
        double fullVolume = n.node_full_volume[i];
        if (!canPond) {
            overflow = dV / dt;
            newVolume = fullVolume;
            yNew = yMax;
        } else {
            double oldVolume = s.n_old_volume[ix];
            newVolume = SWB_MAX((oldVolume + dV), fullVolume);
            overflow = (newVolume - SWB_MAX(oldVolume, fullVolume)) / dt;
        }
        if (overflow < SWB_FUDGE) overflow = 0.0;
    }
    else newVolume = node_volume(n, i, yNew);

    s.n_dydt[ix] = fabs(yNew - yOld) / dt;
    s.n_depth[ix] = yNew;
    s.n_volume[ix] = newVolume;
    s.n_overflow[ix] = overflow;
    s.n_inflow[ix] = acc.inflow;
    s.n_outflow[ix] = acc.outflow;
    bool conv = !(fabs(yLast - yNew) > n.opt.head_tol);
    s.n_converged[ix] = conv ? 1 : 0;
    return conv;
}

// ---- K7: Courant / depth-change step candidates (dynwave.c:836-921) ------------------------------
// return the candidate step of one object, or a negative value when the object is skipped
SWB_FI double link_step(const Net &n, const State &s, int j, int m)
{
    if (n.link_type[j] != SWB_CONDUIT) return -1.0;
    size_t ix = SWB_IX(j, m, s.M);
    double barrels = (double)n.cond_barrels[j];
    // all four loads are issued before the first test: one DRAM round trip per link instead of a
    // chain of dependent ones (the early exits only skip arithmetic)
    const double flow = s.l_flow[ix], froude = s.l_froude[ix], a1 = s.c_a1[ix], volume = s.l_volume[ix];
    double q = fabs(flow) / barrels;
    if (q <= SWB_FUDGE || a1 <= SWB_FUDGE || froude <= 0.01) return -1.0;
    double t = volume / barrels / q;
    t = t * n.cond_mod_length[j] / n.cond_length[j];
    t = t * froude / (1.0 + froude) * n.opt.courant_factor;
    return t;
}
SWB_FI double node_step(const Net &n, const State &s, int i, int m)
{
    if (n.node_type[i] == SWB_OUTFALL) return -1.0;
    size_t ix = SWB_IX(i, m, s.M);
    const double depth = s.n_depth[ix], dYdT = s.n_dydt[ix];     // both loads up front
    double yCrown = n.node_crown_elev[i] - n.node_invert[i];
    if (depth <= SWB_FUDGE) return -1.0;
    if (depth + SWB_FUDGE >= yCrown) return -1.0;
    double maxDepth = yCrown * 0.25;
    if (maxDepth < SWB_FUDGE) return -1.0;
    if (dYdT < SWB_FUDGE) return -1.0;
    return maxDepth / dYdT;
}

} // namespace swb
#endif
