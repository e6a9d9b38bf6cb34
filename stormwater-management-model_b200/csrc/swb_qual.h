// swb_qual.h -- K8: completely-mixed water-quality routing, per (object, member).
//
//   qual_node  <- findLinkMassFlow gathered per node + findNodeQual / findStorageQual + updateHRT
//                 (qualrout.c:179-249, 398-494)
//   qual_link  <- findLinkQual + the link's totalLoad update (qualrout.c:253-353, 214)
//   getMixedQual / getReactedQual (qualrout.c:146-174, 498-518)
//
// The reference scatters link loads into Node.newQual in ascending link order (qualrout.c:112);
// the gather below walks the node's incidence list in the same ascending link order, so each
// node's floating-point sum is identical.  Treatment (treatmnt.c) stays on the host.
#ifndef SWB_QUAL_H
#define SWB_QUAL_H

#include "swb_state.h"

namespace swb {

struct QualAcc { double reacted, seepage, finalStorage; };   // massbal.c:517-555 contributions

SWB_FI double qual_mixed(double c, double v1, double wIn, double qIn, double tStep)
{
    if (qIn <= SWB_ZERO) return c;
    double vIn = qIn * tStep;
    double cIn = wIn * tStep / vIn;
    double cMax = SWB_MAX(c, cIn);
    c = (c * v1 + wIn * tStep) / (v1 + vIn);
    c = SWB_MIN(c, cMax);
    c = SWB_MAX(c, 0.0);
    return c;
}

SWB_FI double qual_reacted(double kDecay, double c, double v1, double tStep, double &reacted)
{
    if (kDecay == 0.0) return c;
    double c2 = c * (1.0 - kDecay * tStep);
    c2 = SWB_MAX(0.0, c2);
    double lossRate = (c - c2) * v1 / tStep;
    reacted += lossRate;
    return c2;
}

// node i, pollutants p0 .. p0 + K - 1 in ONE pass over the node's fields and incident links (K = 1 or 2: the
// hydraulic values -- link flows, node inflow, volumes, depth -- are the same for every pollutant, so a second
// pollutant costs only its own concentrations instead of a second sweep over the state).  Per pollutant the
// operations and their order are those of findLinkMassFlow / findNodeQual / findStorageQual (qualrout.c:
// 179-249, 398-494).  n_qual holds the external mass-rate preload on entry (routing.c:488 ...), the new
// concentration on exit.
template <int K>
SWB_FI void qual_node(const Net &n, const State &s, int i, int m, int p0, double tStep, QualAcc (&acc)[K])
{
    const int M = s.M;
    const size_t ix = SWB_IX(i, m, M);
    size_t ixq[K];
    double w[K];
#pragma unroll
    for (int t = 0; t < K; t++) { ixq[t] = SWB_IXP(p0 + t, i, n.nN, m, M); w[t] = s.n_qual[ixq[t]]; }
    // --- link mass flows into this node, ascending link index (qualrout.c:112, 179-217)
    for (int k = n.adjq_start[i]; k < n.adjq_start[i + 1]; k++) {
        int e = n.adjq[k], j = e >> 1, end = e & 1;
        double q = s.l_flow[SWB_IX(j, m, M)];
        bool into = (q < 0.0) ? (end == 0) : (end == 1);
        if (!into) continue;
#pragma unroll
        for (int t = 0; t < K; t++) w[t] += fabs(q) * s.l_old_qual[SWB_IXP(p0 + t, j, n.nL, m, M)];
    }
    double qIn = s.n_inflow[ix];                       // Node.qualInflow = Node.inflow (:118)
    double oldVolume = s.n_old_volume[ix];
    bool isStorage = (n.node_type[i] == SWB_STORAGE);
    if (isStorage || oldVolume > SWB_ZERO_VOLUME) {
        // findStorageQual (qualrout.c:398-474)
        double v1 = oldVolume, qExfil = 0.0, vEvap = 0.0, fEvap = 1.0;
        if (isStorage) {
            if (p0 == 0) {                             // updateHRT once per node (:478-494)
                double hrt = s.n_hrt[ix];
                if (v1 < SWB_ZERO) hrt = 0.0;
                else hrt = (hrt + tStep) * v1 / (v1 + qIn * tStep);
                s.n_hrt[ix] = SWB_MAX(hrt, 0.0);
            }
            qExfil = s.n_exfil_loss[ix] / tStep;
            vEvap = s.n_evap_loss[ix];
            if (vEvap > 0.0 && v1 > SWB_ZERO_VOLUME) fEvap += vEvap / v1;
        }
        const double newVolume = s.n_volume[ix];
        const bool emptied = (newVolume <= SWB_ZERO_VOLUME || s.n_depth[ix] <= SWB_ZERO_DEPTH) && qIn <= SWB_ZERO;
        double cOld[K];                                 // (all loads before the first store)
#pragma unroll
        for (int t = 0; t < K; t++) cOld[t] = s.n_old_qual[ixq[t]];
#pragma unroll
        for (int t = 0; t < K; t++) {
            double c1 = cOld[t];
            acc[t].seepage += qExfil * c1;
            c1 *= fEvap;
            c1 = qual_reacted(n.pollut_kdecay[p0 + t], c1, v1, tStep, acc[t].reacted);
            double c2 = qual_mixed(c1, v1, w[t], qIn, tStep);
            if (emptied) { acc[t].finalStorage += c2 * newVolume; c2 = 0.0; }
            s.n_qual[ixq[t]] = c2;
        }
    } else {
        // findNodeQual (qualrout.c:221-249)
        const bool wet = s.n_depth[ix] > SWB_ZERO_DEPTH;
#pragma unroll
        for (int t = 0; t < K; t++) {
            double c2;
            if (qIn > SWB_ZERO) c2 = w[t] / qIn;
            else if (wet) c2 = s.n_old_qual[ixq[t]];
            else c2 = 0.0;
            s.n_qual[ixq[t]] = c2;
        }
    }
}

// link j, pollutants p0 .. p0 + K - 1 (qualrout.c:253-394), one pass over the link's hydraulic values
template <int K>
SWB_FI void qual_link(const Net &n, const State &s, int j, int m, int p0, double tStep, QualAcc (&acc)[K])
{
    const int M = s.M;
    const size_t ix = SWB_IX(j, m, M);
    const double newFlow = s.l_flow[ix];
    double qAbs = fabs(newFlow);                       // totalLoad (qualrout.c:210-214)
    int up = n.link_node1[j];
    if (newFlow < 0.0) up = n.link_node2[j];
    // every load of this link is issued before its first store (the arrays may alias as far as
    // the compiler knows, so a load placed after a store waits for a second DRAM round trip)
    size_t iq[K];
    double c1[K], load0[K], cUp[K];
#pragma unroll
    for (int t = 0; t < K; t++) {
        iq[t] = SWB_IXP(p0 + t, j, n.nL, m, M);
        c1[t] = s.l_old_qual[iq[t]];
        load0[t] = s.l_total_load[iq[t]];
        cUp[t] = s.n_qual[SWB_IXP(p0 + t, up, n.nN, m, M)];
    }
    if (!(n.link_flags[j] & LF_TRUE_CONDUIT)) {
#pragma unroll
        for (int t = 0; t < K; t++) { s.l_total_load[iq[t]] = load0[t] + qAbs * c1[t] * tStep; s.l_qual[iq[t]] = cUp[t]; }
        return;
    }
    double barrels = (double)n.cond_barrels[j];
    double qIn = fabs(s.c_q1[ix]) * barrels;
    double qSeep = 0.0, vEvap = 0.0;
    if (n.link_flags[j] & LF_HAS_LOSSRATE) {
        qSeep = s.c_seep_loss[ix] * barrels;
        vEvap = s.c_evap_loss[ix] * barrels * tStep;
    }
    double v1 = s.l_old_volume[ix], v2 = s.l_volume[ix];
    const double depth = s.l_depth[ix];
#pragma unroll
    for (int t = 0; t < K; t++) s.l_total_load[iq[t]] = load0[t] + qAbs * c1[t] * tStep;
    double vLosses = qSeep * tStep + vEvap;
    double fEvap = 1.0;
    if (vEvap > 0.0 && v1 > SWB_ZERO_VOLUME) fEvap += vEvap / v1;
    qIn = qIn + (v2 + vLosses - v1) / tStep;
    qIn = SWB_MAX(qIn, 0.0);
    const bool emptied = (v2 < SWB_ZERO_VOLUME || depth <= SWB_ZERO_DEPTH);
#pragma unroll
    for (int t = 0; t < K; t++) {
        double c = c1[t];
        acc[t].seepage += qSeep * c;
        c *= fEvap;
        double c2 = qual_reacted(n.pollut_kdecay[p0 + t], c, v1, tStep, acc[t].reacted);
        double wIn = cUp[t] * qIn;
        c2 = qual_mixed(c2, v1, wIn, qIn, tStep);
        if (emptied) { acc[t].finalStorage += c2 * v2; c2 = 0.0; }
        s.l_qual[iq[t]] = c2;
    }
}

} // namespace swb
#endif
