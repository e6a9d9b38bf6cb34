// swb_stats.h -- per-object routing statistics of every ensemble member on the device (SURVEY 8f
// rank 1): stats_updateNodeStats / stats_updateLinkStats / stats_updateFlowStats (stats.c:449-754),
// stats_updateConvergenceStats (:536-540), stats_updateCriticalTimeCount (:522-533),
// stats_updateTimeStepStats (:486-518).  Run once per routing step after quality routing and the
// routing totals, exactly where routing_execute calls them (routing.c:258-259).
//
// Layout: plane-major like the pollutant planes, stat[(plane * nItems + item) * M + member]; the
// plane ids are the public enums of include/swmm_b200.h.  Dates are kept as elapsed simulated
// seconds of the step's NEW routing time (the reference stores getDateTime(NewRoutingTime)).
#ifndef SWB_STATS_H
#define SWB_STATS_H

#include "swb_dynwave.h"

namespace swb {

#define SWB_MIN_RUNOFF_FLOW 0.001      // consts.h: MIN_RUNOFF_FLOW (cfs)
#define SWB_KW_PER_HP       0.7457     // consts.h: KWperHP

SWB_FI double &nstat(const State &s, const Net &n, int plane, int i, int m)
{ return s.stat_node[SWB_IXP(plane, i, n.nN, m, s.M)]; }
SWB_FI double &lstat(const State &s, const Net &n, int plane, int j, int m)
{ return s.stat_link[SWB_IXP(plane, j, n.nL, m, s.M)]; }

// stats_updateNodeStats (stats.c:544-647)
SWB_FI void stats_node(const Net &n, const State &s, int i, int m, double tStep, double tNew, bool withQual)
{
    const size_t ix = SWB_IX(i, m, s.M);
    const int type = n.node_type[i];
    double newVolume = s.n_volume[ix];
    const double newDepth = s.n_depth[ix], overflow = s.n_overflow[ix], inflow = s.n_inflow[ix];
    const double newLat = s.n_latflow[ix], oldLat = s.n_old_latflow[ix];
    const double fullVolume = n.node_full_volume[i];
    const bool canPond = (n.opt.allow_ponding && n.node_ponded_area[i] > 0.0);

    nstat(s, n, SWB_NS_SUM_DEPTH, i, m) += newDepth;
    if (newDepth > nstat(s, n, SWB_NS_MAX_DEPTH, i, m)) {
        nstat(s, n, SWB_NS_MAX_DEPTH, i, m) = newDepth;
        nstat(s, n, SWB_NS_MAX_DEPTH_TIME, i, m) = tNew;
    }
    if (type != SWB_OUTFALL) {
        if (newVolume > fullVolume || overflow > 0.0) {
            nstat(s, n, SWB_NS_TIME_FLOODED, i, m) += tStep;
            nstat(s, n, SWB_NS_VOL_FLOODED, i, m) += overflow * tStep;
            if (canPond) {
                double &mp = nstat(s, n, SWB_NS_MAX_PONDED_VOL, i, m);
                mp = SWB_MAX(mp, (newVolume - fullVolume));
            }
        }
        if ((type != SWB_STORAGE || n.node_sur_depth[i] > 0.0) &&
            newDepth + n.node_invert[i] + SWB_FUDGE >= n.node_crown_elev[i])
            nstat(s, n, SWB_NS_TIME_SURCHARGED, i, m) += tStep;
    }
    if (type == SWB_STORAGE) {
        nstat(s, n, SWB_NS_X_SUM, i, m) += newVolume;
        nstat(s, n, SWB_NS_X_EVAP, i, m) += s.n_evap_loss[ix];
        nstat(s, n, SWB_NS_X_EXFIL, i, m) += s.n_exfil_loss[ix];
        newVolume = SWB_MIN(newVolume, fullVolume);
        if (newVolume > nstat(s, n, SWB_NS_X_MAX, i, m)) {
            nstat(s, n, SWB_NS_X_MAX, i, m) = newVolume;
            nstat(s, n, SWB_NS_X_MAX_TIME, i, m) = tNew;
        }
        double &mf = nstat(s, n, SWB_NS_X_MAX_FLOW, i, m);
        mf = SWB_MAX(mf, s.n_outflow[ix]);
    }
    if (type == SWB_OUTFALL) {
        if (inflow >= SWB_MIN_RUNOFF_FLOW) {
            nstat(s, n, SWB_NS_X_SUM, i, m) += inflow;
            double &mf = nstat(s, n, SWB_NS_X_MAX, i, m);
            mf = SWB_MAX(mf, inflow);
            nstat(s, n, SWB_NS_X_MAX_TIME, i, m) += 1.0;          // OutfallStats.totalPeriods
        }
        if (withQual)
            for (int p = 0; p < n.nP; p++)
                nstat(s, n, SWB_NS_LOAD0 + p, i, m) += inflow * s.n_qual[SWB_IXP(p, i, n.nN, m, s.M)] * tStep;
    }
    nstat(s, n, SWB_NS_TOT_LATFLOW, i, m) += ((oldLat + newLat) * 0.5 * tStep);
    if (fabs(newLat) > fabs(nstat(s, n, SWB_NS_MAX_LATFLOW, i, m))) nstat(s, n, SWB_NS_MAX_LATFLOW, i, m) = newLat;
    if (inflow > nstat(s, n, SWB_NS_MAX_INFLOW, i, m)) {
        nstat(s, n, SWB_NS_MAX_INFLOW, i, m) = inflow;
        nstat(s, n, SWB_NS_MAX_INFLOW_TIME, i, m) = tNew;
    }
    if (overflow > nstat(s, n, SWB_NS_MAX_OVERFLOW, i, m)) {
        nstat(s, n, SWB_NS_MAX_OVERFLOW, i, m) = overflow;
        nstat(s, n, SWB_NS_MAX_OVERFLOW_TIME, i, m) = tNew;
    }
}

// stats_updateLinkStats (stats.c:651-754) with link_getVelocity (link.c:821-843), link_getPower (:875)
SWB_FI void stats_link(const Net &n, const State &s, int j, int m, double tStep, double tNew, const double *T)
{
    const size_t ix = SWB_IX(j, m, s.M);
    const int type = n.link_type[j];
    const double newFlow = s.l_flow[ix], oldFlow = s.l_old_flow[ix], newDepth = s.l_depth[ix];
    const double dq = newFlow - oldFlow;
    const double q = fabs(newFlow);
    if (q > lstat(s, n, SWB_LS_MAX_FLOW, j, m)) {
        lstat(s, n, SWB_LS_MAX_FLOW, j, m) = q;
        lstat(s, n, SWB_LS_MAX_FLOW_TIME, j, m) = tNew;
    }
    // link_getVelocity: conduits only, zero below 0.01 ft of depth
    double v = 0.0;
    if (type == SWB_CONDUIT && !(newDepth <= 0.01) && n.xs_type[j] != XS_DUMMY) {
        Xs x = load_xs(n, j);
        double flow = q / (double)n.cond_barrels[j];
        double area = xs_a_of_y_ni(x, newDepth, T);
        if (area > SWB_FUDGE) v = flow / area;
    }
    if (v > lstat(s, n, SWB_LS_MAX_VELOC, j, m)) lstat(s, n, SWB_LS_MAX_VELOC, j, m) = v;
    if (newDepth > lstat(s, n, SWB_LS_MAX_DEPTH, j, m)) lstat(s, n, SWB_LS_MAX_DEPTH, j, m) = newDepth;

    if (type == SWB_PUMP) {
        if (q >= n.link_q_full[j]) lstat(s, n, SWB_LS_TIME_FULL_FLOW, j, m) += tStep;
        if (q > SWB_MIN_RUNOFF_FLOW) {
            const int n1 = n.link_node1[j], n2 = n.link_node2[j];
            double &mn = lstat(s, n, SWB_LS_PUMP_MIN_FLOW, j, m);
            mn = SWB_MIN(mn, q);            // (starts at 0 in the reference, stats.c:278, and so stays 0)
            lstat(s, n, SWB_LS_PUMP_SUM_FLOW, j, m) += q;
            lstat(s, n, SWB_LS_PUMP_VOLUME, j, m) += q * tStep;
            lstat(s, n, SWB_LS_PUMP_UTILIZED, j, m) += tStep;
            const double dh = (n.node_invert[n1] + s.n_depth[SWB_IX(n1, m, s.M)]) -
                              (n.node_invert[n2] + s.n_depth[SWB_IX(n2, m, s.M)]);
            lstat(s, n, SWB_LS_PUMP_ENERGY, j, m) += (fabs(dh) * q / 8.814 * SWB_KW_PER_HP) * tStep / 3600.0;
            const int cls = s.l_flow_class[ix];
            if (cls == SWB_DN_DRY) lstat(s, n, SWB_LS_PUMP_OFF_LOW, j, m) += tStep;
            if (cls == SWB_UP_DRY) lstat(s, n, SWB_LS_PUMP_OFF_HIGH, j, m) += tStep;
            if (oldFlow < SWB_MIN_RUNOFF_FLOW) lstat(s, n, SWB_LS_PUMP_STARTUPS, j, m) += 1.0;
            lstat(s, n, SWB_LS_PUMP_PERIODS, j, m) += 1.0;
            lstat(s, n, SWB_LS_TIME_SURCHARGED, j, m) += tStep;
            lstat(s, n, SWB_LS_TIME_FULL_UP, j, m) += tStep;
            lstat(s, n, SWB_LS_TIME_FULL_DN, j, m) += tStep;
        }
    }
    else if (type == SWB_CONDUIT) {
        if (s.l_normal_flow[ix]) lstat(s, n, SWB_LS_TIME_NORMAL, j, m) += tStep;
        if (s.l_inlet_control[ix]) lstat(s, n, SWB_LS_TIME_INLET, j, m) += tStep;
        const int k = s.l_flow_class[ix];
        if (k >= 0 && k < 7) lstat(s, n, SWB_LS_TIME_CLASS0 + k, j, m) += tStep;
        if (q >= n.link_q_full[j] * (double)n.cond_barrels[j]) lstat(s, n, SWB_LS_TIME_FULL_FLOW, j, m) += tStep;
        if (s.c_cap_limited[ix]) lstat(s, n, SWB_LS_TIME_CAP_LIMITED, j, m) += tStep;
        switch (s.c_full_state[ix]) {
          case SWB_ALL_FULL:
            lstat(s, n, SWB_LS_TIME_SURCHARGED, j, m) += tStep;
            lstat(s, n, SWB_LS_TIME_FULL_UP, j, m) += tStep;
            lstat(s, n, SWB_LS_TIME_FULL_DN, j, m) += tStep;
            break;
          case SWB_UP_FULL: lstat(s, n, SWB_LS_TIME_FULL_UP, j, m) += tStep; break;
          case SWB_DN_FULL: lstat(s, n, SWB_LS_TIME_FULL_DN, j, m) += tStep; break;
        }
    }
    const double k = lstat(s, n, SWB_LS_TURN_SIGN, j, m);
    const double sgn = (double)SWB_SGN(dq);
    lstat(s, n, SWB_LS_TURN_SIGN, j, m) = sgn;
    if (fabs(dq) > 0.001 && k * sgn < 0.0) lstat(s, n, SWB_LS_TURNS, j, m) += 1.0;
}

} // namespace swb
#endif
