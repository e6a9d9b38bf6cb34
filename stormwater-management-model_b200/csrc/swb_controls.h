// swb_controls.h -- calendar helpers, time-series lookups, and the control rules of an ensemble member.
//
// Restates evaluateControlRules (routing.c:269-308): link_setTargetSetting (link.c:604-624),
// controls_evaluate / evaluatePremise / getVariableValue / compareTimes / compareValues /
// updateActionValue / getPIDSetting / updateActionList / executeActionList (controls.c:495-552,
// 1086-1450), link_setSetting with orifice_setSetting / orifice_getWeirCoeff / weir_setSetting
// (link.c:626-639, 1729-1809, 2166-2190), and the stage of TIDAL / TIMESERIES outfalls
// (node.c:1437-1458).  One thread per member walks the rules in order: they are few, order dependent
// (ControlValue / SetPoint carry over from premise to action) and read a handful of state values.
#ifndef SWB_CONTROLS_H
#define SWB_CONTROLS_H

#include "swb_regulator.h"

namespace swb {

#define SWB_MISSING (-1.E10)     // consts.h:35

// ---- calendar (datetime.c:158-241, 439-492) ---------------------------------------------------------
struct DateParts { int month, day, hour, yday; };   // month 0-11, day of week 0 = Sunday, day of year 1-366
SWB_HD DateParts date_parts(double date)
{
    DateParts r;
    const int DateDelta = 693594, D1 = 365, D4 = 1461, D100 = 36524, D400 = 146097;
    int t = (int)floor(date) + DateDelta;
    r.day = t % 7;
    int month = 1, yday = 1;
    if (t > 0) {
        t--;
        int y = 1;
        while (t >= D400) { t -= D400; y += 400; }
        int i = t / D100, d = t - i * D100;
        if (i == 4) { i--; d += D100; }
        y += i * 100;
        i = d / D4; d = d - i * D4;
        y += i * 4;
        i = d / D1; d = d - i * D1;
        if (i == 4) { i--; d += D1; }
        y += i;
        yday = d + 1;
        const bool leap = (y % 4 == 0) && ((y % 100 != 0) || (y % 400 == 0));
        for (;;) {
            int n = (month == 2) ? (leap ? 29 : 28) : ((month == 4 || month == 6 || month == 9 || month == 11) ? 30 : 31);
            if (d < n) break;
            d -= n; month++;
        }
    }
    r.month = month - 1;
    r.yday = yday;
    const double fracDay = (date - floor(date)) * 86400.;
    int secs = (int)floor(fracDay + 0.5);
    if (secs >= 86400) secs = 86399;
    int h = (secs / 60) / 60;
    if (h > 23) h = 0;
    r.hour = h;
    return r;
}

// ---- time series (table.c:745-806): linear between breakpoints; outside the range 0 (extend = FALSE)
//      or the nearest end value (extend = TRUE).  The reference walks a cursor forward in time; with
//      monotone time that is the first bracket whose right end is >= t, which is what the scan finds.
SWB_HD double series_lookup(const int *start, const double *tt, const double *vv, int k, double t, bool extend = false)
{
    int i0 = start[k], i1 = start[k + 1];
    if (i1 <= i0) return 0.0;
    if (t < tt[i0]) return extend ? vv[i0] : 0.0;
    if (t > tt[i1 - 1]) return extend ? vv[i1 - 1] : 0.0;
    for (int i = i0 + 1; i < i1; i++) {
        if (t <= tt[i]) return tbl_interp(t, tt[i - 1], vv[i - 1], tt[i], vv[i]);
    }
    return extend ? vv[i1 - 1] : 0.0;
}

// ---- control rules ------------------------------------------------------------------------------------
enum { RA_DEPTH = 0, RA_MAXDEPTH, RA_HEAD, RA_VOLUME, RA_INFLOW, RA_FLOW, RA_FULLFLOW, RA_FULLDEPTH, RA_STATUS,
       RA_SETTING, RA_LENGTH, RA_SLOPE, RA_VELOCITY, RA_TIMEOPEN, RA_TIMECLOSED, RA_TIME, RA_DATE, RA_CLOCKTIME,
       RA_DAYOFYEAR, RA_DAY, RA_MONTH };                         // RuleAttrib (controls.c:69-73)
enum { RO_GAGE = 0, RO_NODE = 1, RO_LINK = 2, RO_SIM = 8 };     // RuleObject (controls.c:67-68) as stored
enum { RP_OR = 3 };                                              // RuleState r_OR (controls.c:65)
enum { RS_PID = 2 };                                             // RuleSetting r_PID (controls.c:75)

struct Controls {               // device image of swb_controls_desc + the per-member rule state
    int active;                 // 0: no swb_set_controls call
    int nRules, nAct, nWatch, nStage;
    double rule_step;           // RuleStep, s
    double start_datetime;      // StartDateTime
    double start_day, start_secs;
    const double *rule_priority;
    const int    *rule_prem_start, *rule_then_start, *rule_else_start, *act_then, *act_else;
    const int    *prem_type, *prem_lhs_obj, *prem_lhs_index, *prem_lhs_attr;
    const int    *prem_rhs_is_var, *prem_rhs_obj, *prem_rhs_index, *prem_rhs_attr, *prem_relation;
    const double *prem_value;
    const int    *act_rule, *act_link, *act_attr, *act_curve, *act_tseries, *act_slot;
    const double *act_kp, *act_ki, *act_kd;
    const int    *series_start;
    const double *series_t, *series_v;
    const double *pump_y_on, *pump_y_off, *orif_orate;      // per link
    const int    *watch;        // links whose setting can change (pumps, action targets), ascending
    const int    *link_watch;   // per link: index into watch or -1
    const int    *stage_node, *stage_kind, *stage_table;    // outfalls with a TIDAL curve (1) / stage series (2)
    // per-member state
    double *control_value, *set_point;       // [M]      ControlValue / SetPoint (controls.c:168-169)
    double *act_val, *act_e1, *act_e2;       // [nAct][M] TAction.value / e1 / e2
    double *time_last_set;                   // [nWatch][M] Link.timeLastSet
    int    *winner;                          // [nWatch][M] the action list: winning action per link or -1
    double *new_rule_time;                   // [M] NewRuleTime, ms
};

struct CtlClock { double date, curDate, curTime, elapsed, tStep; };   // controls_evaluate's shared variables

// getVariableValue (controls.c:1286-1386)
SWB_HD double ctl_variable(const Net &n, const State &s, const Controls &c, int m, int obj, int idx, int attr,
                           const CtlClock &ck, const double *T)
{
    const int M = s.M;
    int i = -1, j = -1;
    if (obj == RO_GAGE) return SWB_MISSING;           // (rejected when the rules are installed)
    if (obj == RO_NODE) i = idx;
    if (obj == RO_LINK) j = idx;
    const double ucfQ = n.opt.ucf_flow, ucfL = n.opt.ucf_length, ucfV = n.opt.ucf_volume;
    switch (attr) {
      case RA_TIME: return ck.elapsed;
      case RA_DATE: return ck.curDate;
      case RA_CLOCKTIME: return ck.curTime;
      case RA_DAY: return (double)(date_parts(ck.curDate).day + 1);
      case RA_MONTH: return (double)(date_parts(ck.curDate).month + 1);
      case RA_DAYOFYEAR: return (double)date_parts(ck.curDate).yday;
      case RA_STATUS:
        if (j < 0 || (n.link_type[j] != SWB_CONDUIT && n.link_type[j] != SWB_PUMP)) return SWB_MISSING;
        return s.l_setting[SWB_IX(j, m, M)];
      case RA_SETTING:
        if (j < 0 || (n.link_type[j] != SWB_PUMP && n.link_type[j] != SWB_ORIFICE && n.link_type[j] != SWB_WEIR)) return SWB_MISSING;
        return s.l_setting[SWB_IX(j, m, M)];
      case RA_FLOW:
        if (j < 0) return SWB_MISSING;
        return n.link_direction[j] * s.l_flow[SWB_IX(j, m, M)] * ucfQ;
      case RA_FULLFLOW: case RA_FULLDEPTH: case RA_VELOCITY: case RA_LENGTH: case RA_SLOPE:
        if (j < 0) return SWB_MISSING;
        if (n.link_type[j] != SWB_CONDUIT) return SWB_MISSING;
        switch (attr) {
          case RA_FULLFLOW: return n.link_q_full[j] * ucfQ;
          case RA_FULLDEPTH: return n.xs_yfull[j] * ucfL;
          case RA_VELOCITY: {                          // link_getVelocity (link.c:821-843)
            const double depth = s.l_depth[SWB_IX(j, m, M)];
            double veloc = 0.0;
            if (depth > 0.01) {
                const double flow = s.l_flow[SWB_IX(j, m, M)] / n.cond_barrels[j];
                const Xs x = load_xs(n, j);
                const double area = xs_a_of_y_ni(x, depth, T);
                if (area > SWB_FUDGE) veloc = flow / area;
            }
            return veloc * ucfL; }
          case RA_LENGTH: return n.cond_length[j] * ucfL;
          default: return n.cond_slope[j];
        }
      case RA_DEPTH:
        if (j >= 0) return s.l_depth[SWB_IX(j, m, M)] * ucfL;
        if (i >= 0) return s.n_depth[SWB_IX(i, m, M)] * ucfL;
        return SWB_MISSING;
      case RA_MAXDEPTH:
        if (i >= 0) return n.node_full_depth[i] * ucfL;
        return SWB_MISSING;
      case RA_HEAD:
        if (i < 0) return SWB_MISSING;
        return (s.n_depth[SWB_IX(i, m, M)] + n.node_invert[i]) * ucfL;
      case RA_VOLUME:
        if (i < 0) return SWB_MISSING;
        return s.n_volume[SWB_IX(i, m, M)] * ucfV;
      case RA_INFLOW:
        if (i < 0) return SWB_MISSING;
        return s.n_latflow[SWB_IX(i, m, M)] * ucfQ;
      case RA_TIMEOPEN: case RA_TIMECLOSED: {
        if (j < 0) return SWB_MISSING;
        const double setting = s.l_setting[SWB_IX(j, m, M)];
        if (attr == RA_TIMEOPEN ? (setting <= 0.0) : (setting > 0.0)) return SWB_MISSING;
        const int w = c.link_watch[j];
        // a link no rule or pump depth can switch keeps the time it was last set before the run
        const double tls = w >= 0 ? c.time_last_set[(size_t)w * M + m] : c.start_datetime;
        return ck.curDate + ck.curTime - tls; }
      default: return SWB_MISSING;
    }
}

// compareValues / compareTimes (controls.c:1405-1450)
SWB_HD bool ctl_compare_values(const Controls &c, int m, double lhs, int rel, double rhs)
{
    c.set_point[m] = rhs;
    c.control_value[m] = lhs;
    switch (rel) {
      case 0: return lhs == rhs;
      case 1: return lhs != rhs;
      case 2: return lhs <  rhs;
      case 3: return lhs <= rhs;
      case 4: return lhs >  rhs;
      case 5: return lhs >= rhs;
    }
    return false;
}
SWB_HD bool ctl_compare_times(const Controls &c, int m, double lhs, int rel, double rhs, double halfStep)
{
    if (rel == 0) return lhs >= rhs - halfStep && lhs < rhs + halfStep;
    if (rel == 1) return lhs < rhs - halfStep || lhs >= rhs + halfStep;
    return ctl_compare_values(c, m, lhs, rel, rhs);
}

// evaluatePremise (controls.c:1243-1282)
SWB_HD bool ctl_premise(const Net &n, const State &s, const Controls &c, int m, int p, const CtlClock &ck, const double *T)
{
    const double lhs = ctl_variable(n, s, c, m, c.prem_lhs_obj[p], c.prem_lhs_index[p], c.prem_lhs_attr[p], ck, T);
    const double rhs = c.prem_rhs_is_var[p]
        ? ctl_variable(n, s, c, m, c.prem_rhs_obj[p], c.prem_rhs_index[p], c.prem_rhs_attr[p], ck, T) : c.prem_value[p];
    if (lhs == SWB_MISSING || rhs == SWB_MISSING) return false;
    const int rel = c.prem_relation[p];
    switch (c.prem_lhs_attr[p]) {
      case RA_TIME: case RA_CLOCKTIME:
        return ctl_compare_times(c, m, lhs, rel, rhs, ck.tStep / 2.0);
      case RA_TIMEOPEN: case RA_TIMECLOSED: {
        const bool r = ctl_compare_times(c, m, lhs, rel, rhs, ck.tStep / 2.0);
        c.control_value[m] = lhs * 24.0;
        return r; }
      default:
        return ctl_compare_values(c, m, lhs, rel, rhs);
    }
}

// updateActionValue / getPIDSetting (controls.c:1086-1164)
SWB_HD void ctl_action_value(const Net &n, const State &s, const Controls &c, int m, int a, const CtlClock &ck)
{
    const int M = s.M;
    const size_t ia = (size_t)a * M + m;
    if (c.act_curve[a] >= 0) c.act_val[ia] = curve_lookup(n, c.act_curve[a], c.control_value[m]);
    else if (c.act_tseries[a] >= 0)
        c.act_val[ia] = series_lookup(c.series_start, c.series_t, c.series_v, c.act_tseries[a], ck.date, true);
    else if (c.act_attr[a] == RS_PID) {
        const double tolerance = 0.0001;
        const double dt = ck.tStep * 1440.0;
        const double setPoint = c.set_point[m], controlValue = c.control_value[m];
        double e0 = setPoint - controlValue;
        if (fabs(e0) > SWB_TINY) {
            if (setPoint != 0.0) e0 = e0 / setPoint;
            else                 e0 = e0 / controlValue;
        }
        double e1 = c.act_e1[ia], e2 = c.act_e2[ia];
        if (fabs(e0 - e1) < tolerance) { e2 = 0.0; e1 = 0.0; }
        const double p = (e0 - e1);
        double i;
        if (c.act_ki[a] == 0.0) i = 0.0;
        else i = e0 * dt / c.act_ki[a];
        const double d = c.act_kd[a] * (e0 - 2.0 * e1 + e2) / dt;
        double update = c.act_kp[a] * (p + i + d);
        if (fabs(update) < tolerance) update = 0.0;
        const int j = c.act_link[a];
        double setting = s.l_target_setting[SWB_IX(j, m, M)] + update;
        c.act_e2[ia] = e1;
        c.act_e1[ia] = e0;
        if (setting < 0.0) setting = 0.0;
        if (n.link_type[j] != SWB_PUMP && setting > 1.0) setting = 1.0;
        c.act_val[ia] = setting;
    }
}

// orifice_setSetting with orifice_getWeirCoeff (link.c:1729-1809), weir_setSetting (link.c:2166-2190)
SWB_HD void ctl_set_setting(const Net &n, const State &s, const Controls &c, int m, int j, double tstep, const double *T)
{
    const int M = s.M;
    const size_t ix = SWB_IX(j, m, M);
    const double target = s.l_target_setting[ix];
    const int type = n.link_type[j];
    if (type == SWB_ORIFICE) {
        double setting = s.l_setting[ix];
        const double orate = c.orif_orate[j];
        if (orate == 0.0 || tstep == 0.0) setting = target;
        else {
            const double delta = target - setting;
            const double step = tstep / orate;
            if (step + 0.001 >= fabs(delta)) setting = target;
            else setting += SWB_SGN(delta) * step;
        }
        s.l_setting[ix] = setting;
        const Xs x = load_xs(n, j);
        double h = setting * x.yFull;
        const double f = xs_a_of_y_ni(x, h, T) * sqrt(2.0 * SWB_GRAVITY);
        const double cDisch = n.orif_cdisch[j];
        s.o_corif[ix] = cDisch * f;
        if (n.orif_type[j] == 1 /* BOTTOM_ORIFICE */) {
            double aOverL;
            if (x.type == XS_CIRCULAR) aOverL = h / 4.0;
            else { const double w = x.wMax; aOverL = (h * w) / (2.0 * (h + w)); }
            h = cDisch / 0.414 * aOverL;
            s.o_hcrit[ix] = h;
        } else {
            s.o_hcrit[ix] = h;
            h = h / 2.0;
        }
        s.o_cweir[ix] = (cDisch * sqrt(h)) * f;
    } else if (type == SWB_WEIR) {
        s.l_setting[ix] = target;
        if (!n.weir_can_surcharge[j]) return;
        if (n.weir_type[j] == 4 /* ROADWAY_WEIR */) return;
        if (target == 0.0) s.w_csurcharge[ix] = 0.0;
        else {
            const Xs x = load_xs(n, j);
            double h = target * x.yFull;
            RegCtx r;
            r.j = j; r.m = m; r.n1 = n.link_node1[j]; r.n2 = n.link_node2[j]; r.flags = 0;
            r.ix = ix; r.ix1 = SWB_IX(r.n1, m, M); r.ix2 = SWB_IX(r.n2, m, M);
            r.depth1 = r.depth2 = r.inv1 = r.inv2 = 0.0;
            double q1, q2;
            weir_flow(n, s, r, x, h, 1.0, false, q1, q2, T);
            const double q = q1 + q2;
            h = h / 2.0;
            s.w_csurcharge[ix] = q / sqrt(h);
        }
    } else s.l_setting[ix] = target;
}

// evaluateControlRules for member m at the start of a routing step of length dt (routing.c:269-308); the
// state still holds the previous step's "new" values, sim_time the time before the step.  Also sets the
// stage of TIDAL / TIMESERIES outfalls for the step, which the reference reads with NewRoutingTime already
// advanced (node.c:1437-1458 is reached from the node phase).
SWB_HD void ph_controls(const Net &n, const State &s, const Controls &c, int m, double dt, const double *T)
{
    const int M = s.M;
    const double tms = s.time_ms[m];                                    // NewRoutingTime
    CtlClock ck;
    ck.date = c.start_day + (c.start_secs + (tms + 1.0) / 1000.0) / 86400.0;      // getDateTime (swmm5.c:1543-1552)
    ck.curDate = floor(ck.date);
    ck.curTime = ck.date - floor(ck.date);
    ck.elapsed = ck.date - c.start_datetime;
    ck.tStep = dt / 86400.0;
    // ---- link_setTargetSetting: only pumps do anything (link.c:604-624)
    for (int w = 0; w < c.nWatch; w++) {
        const int j = c.watch[w];
        if (n.link_type[j] != SWB_PUMP) continue;
        const size_t ix = SWB_IX(j, m, M);
        const double setting = s.l_setting[ix];
        double target = setting;
        const double yOn = c.pump_y_on[j], yOff = c.pump_y_off[j];
        const double depth = s.n_depth[SWB_IX(n.link_node1[j], m, M)];
        if (yOff > 0.0 && setting > 0.0 && depth < yOff) target = 0.0;
        if (yOn > 0.0 && setting == 0.0 && depth > yOn) target = 1.0;
        s.l_target_setting[ix] = target;
    }
    // ---- controls_evaluate (controls.c:495-552)
    if (c.nRules > 0 && (c.rule_step == 0.0 || fabs(tms - c.new_rule_time[m]) < 1.0)) {
        for (int w = 0; w < c.nWatch; w++) c.winner[(size_t)w * M + m] = -1;
        for (int r = 0; r < c.nRules; r++) {
            bool result = true;
            for (int p = c.rule_prem_start[r]; p < c.rule_prem_start[r + 1]; p++) {
                if (c.prem_type[p] == RP_OR) { if (!result) result = ctl_premise(n, s, c, m, p, ck, T); }
                else {
                    if (!result) break;
                    result = ctl_premise(n, s, c, m, p, ck, T);
                }
            }
            const int *list = result ? c.act_then : c.act_else;
            const int *lstart = result ? c.rule_then_start : c.rule_else_start;
            for (int q = lstart[r]; q < lstart[r + 1]; q++) {
                const int a = list[q];
                ctl_action_value(n, s, c, m, a, ck);
                // updateActionList (controls.c:1168-1202): one entry per link, replaced only by a rule
                // of strictly higher priority
                const size_t iw = (size_t)c.act_slot[a] * M + m;
                const int a1 = c.winner[iw];
                if (a1 < 0 || c.rule_priority[c.act_rule[a]] > c.rule_priority[c.act_rule[a1]]) c.winner[iw] = a;
            }
        }
        // executeActionList (controls.c:1206-1239)
        for (int w = 0; w < c.nWatch; w++) {
            const int a = c.winner[(size_t)w * M + m];
            if (a < 0) continue;
            const size_t ix = SWB_IX(c.act_link[a], m, M);
            const double v = c.act_val[(size_t)a * M + m];
            if (s.l_target_setting[ix] != v) s.l_target_setting[ix] = v;
        }
    }
    // ---- settings follow their targets (routing.c:286-297)
    for (int w = 0; w < c.nWatch; w++) {
        const int j = c.watch[w];
        const size_t ix = SWB_IX(j, m, M);
        const double target = s.l_target_setting[ix], setting = s.l_setting[ix];
        if (target != setting) {
            if (target * setting == 0.0) c.time_last_set[(size_t)w * M + m] = ck.date;
            ctl_set_setting(n, s, c, m, j, dt, T);
        }
    }
    // ---- NewRuleTime (routing.c:303-306)
    const double newms = tms + 1000.0 * dt;
    if (fabs(newms - (c.new_rule_time[m] + 1000.0 * c.rule_step)) < 1) c.new_rule_time[m] += 1000.0 * c.rule_step;
    // ---- outfall stages for the step (node.c:1437-1458)
    for (int k = 0; k < c.nStage; k++) {
        const int i = c.stage_node[k];
        double stage;
        if (c.stage_kind[k] == 1) {
            const int cv = c.stage_table[k];
            double x = n.curve_x[n.curve_start[cv]];
            const double currentDate = newms / 8.64e7;
            x += (currentDate - floor(currentDate)) * 24.0;
            stage = curve_lookup(n, cv, x) / n.opt.ucf_length;
        } else {
            const double currentDate = c.start_datetime + newms / 8.64e7;
            stage = series_lookup(c.series_start, c.series_t, c.series_v, c.stage_table[k], currentDate, true) / n.opt.ucf_length;
        }
        s.n_stage[SWB_IX(i, m, M)] = stage;
    }
}

// routing_getRoutingStep's rule clamp (routing.c:190-199)
SWB_HD double ctl_clamp_step(const Controls &c, int m, double tms, double dt)
{
    if (c.active && c.rule_step > 0.0) {
        const double nextRuleTime = c.new_rule_time[m] + 1000. * c.rule_step;
        const double nextRoutingTime = tms + 1000. * dt;
        if (nextRoutingTime >= nextRuleTime) dt = (nextRuleTime - tms) / 1000.0;
    }
    return dt;
}

}  // namespace swb
#endif
