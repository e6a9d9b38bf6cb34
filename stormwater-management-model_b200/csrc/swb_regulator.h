// swb_regulator.h -- K2: pumps, orifices, weirs, outlets and dummy conduits.
//
// These links are few and ORDER DEPENDENT: findNonConduitFlow runs in
// ascending link order between updateNodeFlows calls (dynwave.c:404-411) and ideal pumps, dummy
// conduits and Type 2-4 pumps read the partially accumulated node inflow / outflow.  One thread per
// member therefore walks them in index order (a warp = 32 members in lockstep on the same link).
//
// Restates: findNonConduitFlow / getModPumpFlow / findNonConduitSurfArea (dynwave.c:423-524),
// link_getInflow (link.c:543), pump_getInflow (:1548), orifice_getInflow / getFlow (:1812-2004),
// weir_getInflow / getFlow / getOrificeFlow / getOpenArea / getdqdh (:2198-2517),
// outlet_getInflow / getFlow (:2608-2692), node_getMaxOutflow (node.c:418),
// table_lookup / getSlope / intervalLookup (table.c:395-520).
// Regulator formulas are evaluated in USER units exactly like the reference (UCF factors).
#ifndef SWB_REGULATOR_H
#define SWB_REGULATOR_H

#include "swb_dynwave.h"
#include "swb_hds5_tables.h"

namespace swb {

SWB_HD double curve_lookup(const Net &n, int c, double x)             // table.c:395-426
{
    int i0 = n.curve_start[c], i1 = n.curve_start[c + 1];
    if (i1 <= i0) return 0.0;
    double x1 = n.curve_x[i0], y1 = n.curve_y[i0];
    if (x <= x1) return y1;
    for (int i = i0 + 1; i < i1; i++) {
        double x2 = n.curve_x[i], y2 = n.curve_y[i];
        if (x <= x2) return tbl_interp(x, x1, y1, x2, y2);
        x1 = x2; y1 = y2;
    }
    return y1;
}
SWB_HD double curve_slope(const Net &n, int c, double x)              // table.c:430-460
{
    int i0 = n.curve_start[c], i1 = n.curve_start[c + 1];
    if (i1 <= i0) return 0.0;
    double x1 = n.curve_x[i0], y1 = n.curve_y[i0], x2 = x1, y2 = y1;
    for (int i = i0 + 1; i < i1; i++) {
        x2 = n.curve_x[i]; y2 = n.curve_y[i];
        if (x <= x2) break;
        x1 = x2; y1 = y2;
    }
    double dx = x2 - x1;
    if (dx == 0.0) return 0.0;
    return (y2 - y1) / dx;
}
SWB_HD double curve_interval_lookup(const Net &n, int c, double x)    // table.c:504-522
{
    int i0 = n.curve_start[c], i1 = n.curve_start[c + 1];
    if (i1 <= i0) return 0.0;
    if (x < n.curve_x[i0]) return n.curve_y[i0];
    for (int i = i0 + 1; i < i1; i++) if (x < n.curve_x[i]) return n.curve_y[i];
    return n.curve_y[i1 - 1];
}

struct RegCtx {             // one link of one member, gathered once
    int j, m, n1, n2, flags;
    size_t ix, ix1, ix2;
    double depth1, depth2, inv1, inv2;
};

// ---- pump_getInflow (link.c:1548-1634) -----------------------------------------------------------
SWB_HD double pump_inflow(const Net &n, const State &s, const RegCtx &r)
{
    const int j = r.j;
    const double ucfL = n.opt.ucf_length, ucfV = n.opt.ucf_volume, ucfQ = n.opt.ucf_flow;
    int c = n.pump_curve[j];
    s.l_flow_class[r.ix] = 0;                       // NO
    double setting = s.l_target_setting[r.ix];
    s.l_setting[r.ix] = setting;
    if (setting == 0.0) return 0.0;
    double qIn, dh = 0.001, sp = 1.0;
    if (n.pump_type[j] == 5 /*IDEAL_PUMP*/) qIn = s.n_inflow[r.ix1] + s.n_overflow[r.ix1];
    else switch (n.curve_type[c]) {
      case 7: {   // PUMP1_CURVE
        double vol = s.n_volume[r.ix1] * ucfV;
        qIn = curve_interval_lookup(n, c, vol) / ucfQ;
        if (vol < n.pump_xmin[j] || vol > n.pump_xmax[j]) s.l_flow_class[r.ix] = 1;
        break; }
      case 8: {   // PUMP2_CURVE
        double depth = r.depth1 * ucfL;
        qIn = curve_interval_lookup(n, c, depth) / ucfQ;
        if (depth < n.pump_xmin[j] || depth > n.pump_xmax[j]) s.l_flow_class[r.ix] = 1;
        break; }
      case 9: case 11: {   // PUMP3_CURVE, PUMP5_CURVE
        if (n.curve_type[c] == 11) sp = setting;
        double head = ((r.depth2 + r.inv2) - (r.depth1 + r.inv1)) / sp / sp;
        head = SWB_MAX(head, 0.0) * ucfL;
        qIn = curve_lookup(n, c, head) / ucfQ;
        s.l_dqdh[r.ix] = -curve_slope(n, c, head) * ucfL / ucfQ / sp;
        if (head < n.pump_xmin[j] || head > n.pump_xmax[j]) s.l_flow_class[r.ix] = 1;
        break; }
      case 10: {  // PUMP4_CURVE
        double depth = r.depth1;
        qIn = curve_lookup(n, c, depth * ucfL) / ucfQ;
        double qIn1 = curve_lookup(n, c, (depth + dh) * ucfL) / ucfQ;
        s.l_dqdh[r.ix] = (qIn1 - qIn) / dh;
        depth *= ucfL;
        if (depth < n.pump_xmin[j]) s.l_flow_class[r.ix] = SWB_DN_DRY;
        if (depth > n.pump_xmax[j]) s.l_flow_class[r.ix] = SWB_UP_DRY;
        break; }
      default: qIn = 0.0;
    }
    if (qIn < 0.0) qIn = 0.0;
    return qIn * setting;
}

// ---- orifice_getFlow (link.c:1938-2004), one level of the flap-gate recursion unrolled ------------
SWB_HD double orifice_flow_core(const State &s, size_t ix, double head, double f)
{
    double q;
    if (head == 0.0 || f <= 0.0) { s.l_dqdh[ix] = 0.0; return 0.0; }
    else if (f < 1.0) {
        q = s.o_cweir[ix] * pow(f, 1.5);
        s.l_dqdh[ix] = 1.5 * q / (f * s.o_hcrit[ix]);
    } else {
        q = s.o_corif[ix] * sqrt(head);
        s.l_dqdh[ix] = q / (2.0 * head);
    }
    return q;
}
SWB_HD double orifice_flow(const Net &n, const State &s, const RegCtx &r, const Xs &x,
                                  double head, double f, bool hasFlapGate, const double *T)
{
    if (head == 0.0 || f <= 0.0) { s.l_dqdh[r.ix] = 0.0; return 0.0; }
    double q = orifice_flow_core(s, r.ix, head, f);
    if (hasFlapGate) {
        double area = xs_a_of_y_ni(x, s.l_setting[r.ix] * x.yFull, T);
        double veloc = q / area;
        double hLoss = (4.0 / SWB_GRAVITY) * veloc * veloc * exp(-1.15 * veloc / sqrt(head));
        if (f < 1.0) { f = f - hLoss / s.o_hcrit[r.ix]; if (f < 0.0) f = 0.0; }
        else { head = head - hLoss; if (head < 0.0) head = 0.0; }
        q = orifice_flow_core(s, r.ix, head, f);
    }
    return q;
}

// ---- orifice_getInflow (link.c:1812-1934) ---------------------------------------------------------
SWB_HD double orifice_inflow(const Net &n, const State &s, const RegCtx &r, const double *T)
{
    const int j = r.j;
    const Xs x = load_xs(n, j);
    double h1 = r.depth1 + r.inv1, h2 = r.depth2 + r.inv2, head, f, hcrest, hcrown, hmidpt;
    double dir = (h1 >= h2) ? +1.0 : -1.0;
    double y1 = r.depth1;
    if (dir < 0.0) { head = h1; h1 = h2; h2 = head; y1 = r.depth2; }
    double setting = s.l_setting[r.ix];
    if (n.orif_type[j] == 1 /*BOTTOM_ORIFICE*/) {
        hcrest = r.inv1 + n.link_offset1[j];
        if (h1 < hcrest) head = 0.0;
        else if (h2 > hcrest) head = h1 - h2;
        else head = h1 - hcrest;
        f = head / s.o_hcrit[r.ix];
        f = SWB_MIN(f, 1.0);
    } else {
        hcrest = r.inv1 + n.link_offset1[j];
        hcrown = hcrest + x.yFull * setting;
        hmidpt = (hcrest + hcrown) / 2.0;
        if (h1 < hcrown && hcrown > hcrest) f = (h1 - hcrest) / (hcrown - hcrest);
        else f = 1.0;
        if (f < 1.0)          head = h1 - hcrest;
        else if (h2 < hmidpt) head = h1 - hmidpt;
        else                  head = h1 - h2;
    }
    if (head <= SWB_FUDGE || y1 <= SWB_FUDGE || link_flap_closed(r.flags, n.link_direction[j], dir)) {
        s.l_depth[r.ix] = 0.0;
        s.l_flow_class[r.ix] = SWB_DRY;
        s.r_surf_area[r.ix] = SWB_FUDGE * n.orif_length[j];
        s.l_dqdh[r.ix] = 0.0;
        return 0.0;
    }
    int cls = SWB_SUBCRITICAL;
    if (hcrest > h2) cls = (dir == 1.0) ? SWB_DN_CRITICAL : SWB_UP_CRITICAL;
    s.l_flow_class[r.ix] = (unsigned char)cls;
    y1 = x.yFull * setting;
    if (n.orif_type[j] == 0 /*SIDE_ORIFICE*/) {
        double d = y1 * f;
        s.l_depth[r.ix] = d;
        s.r_surf_area[r.ix] = xs_w_of_y_ni(x, d, T) * n.orif_length[j];
    } else {
        s.l_depth[r.ix] = y1;
        s.r_surf_area[r.ix] = xs_a_of_y_ni(x, y1, T);
    }
    double q = dir * orifice_flow(n, s, r, x, head, f, (r.flags & LF_HAS_FLAP) != 0, T);
    if (f < 1.0 && h2 > hcrest) {
        double ratio = (h2 - hcrest) / (h1 - hcrest);
        q *= pow((1.0 - pow(ratio, 1.5)), 0.385);
    }
    return q;
}

// ---- weirs (link.c:2198-2517) ---------------------------------------------------------------------
SWB_HD double weir_open_area(const State &s, const RegCtx &r, const Xs &x, double y, const double *T)
{
    double z = (1.0 - s.l_setting[r.ix]) * x.yFull;
    double zy = z + y;
    zy = SWB_MIN(zy, x.yFull);
    return xs_a_of_y_ni(x, zy, T) - xs_a_of_y_ni(x, z, T);
}
SWB_HD double weir_dqdh(int wtype, double dir, double h, double q1, double q2)
{
    if (fabs(h) < SWB_FUDGE) return 0.0;
    double q1h = fabs(q1 / h), q2h = fabs(q2 / h);
    switch (wtype) {
      case 0: return 1.5 * q1h;
      case 1: if (dir < 0.0) return 1.5 * q1h; else return 1.67 * q1h;
      case 2: if (q2h == 0.0) return 2.5 * q1h; else return 1.5 * q1h + 2.5 * q2h;
      case 3: return 1.5 * q1h + 2.5 * q2h;
    }
    return 0.0;
}
SWB_HD void weir_flow_core(const Net &n, const State &s, const RegCtx &r, const Xs &x,
                                  double head, double dir, double &q1, double &q2, const double *T)
{
    const int j = r.j;
    q1 = 0.0; q2 = 0.0;
    s.l_dqdh[r.ix] = 0.0;
    if (head <= 0.0) return;
    const double ucfL = n.opt.ucf_length;
    double length = x.wMax * ucfL;
    double h = head * ucfL;
    double cDisch1 = n.weir_cdisch1[j];
    int cdCurve = n.weir_cd_curve[j];
    if (cdCurve >= 0) cDisch1 = curve_lookup(n, cdCurve, h);
    double setting = s.l_setting[r.ix];
    int wType = n.weir_type[j];
    if (wType == 2 && setting < 1.0) wType = 3;
    switch (wType) {
      case 0:
        length -= 0.1 * n.weir_end_con[j] * h;
        length = SWB_MAX(length, 0.0);
        q1 = cDisch1 * length * pow(h, 1.5);
        break;
      case 1:
        length -= 0.1 * n.weir_end_con[j] * h;
        length = SWB_MAX(length, 0.0);
        if (dir < 0.0) q1 = cDisch1 * length * pow(h, 1.5);
        else q1 = cDisch1 * pow(length, 0.83) * pow(h, 1.67);
        break;
      case 2:
        q1 = cDisch1 * n.weir_slope[j] * pow(h, 2.5);
        break;
      case 3: {
        double y = (1.0 - setting) * x.yFull;
        length = xs_w_of_y_ni(x, y, T) * ucfL;
        q1 = cDisch1 * length * pow(h, 1.5);
        q2 = n.weir_cdisch2[j] * n.weir_slope[j] * pow(h, 2.5);
        break; }
    }
    if (n.opt.unit_system == 1) { q1 /= SWB_M3_PER_FT3; q2 /= SWB_M3_PER_FT3; }
}
SWB_HD void weir_flow(const Net &n, const State &s, const RegCtx &r, const Xs &x, double head,
                             double dir, bool hasFlapGate, double &q1, double &q2, const double *T)
{
    q1 = 0.0; q2 = 0.0;
    s.l_dqdh[r.ix] = 0.0;
    if (head <= 0.0) return;
    weir_flow_core(n, s, r, x, head, dir, q1, q2, T);
    if (hasFlapGate) {
        double area = weir_open_area(s, r, x, head, T);
        if (area > SWB_TINY) {
            double veloc = (q1 + q2) / area;
            double hLoss = (4.0 / SWB_GRAVITY) * veloc * veloc * exp(-1.15 * veloc / sqrt(head));
            head = head - hLoss;
            if (head < 0.0) head = 0.0;
            // inner call of the recursion (hasFlapGate = FALSE): sets dqdh for the reduced head,
            // or returns with q = 0, dqdh = 0 when the head vanished
            weir_flow_core(n, s, r, x, head, dir, q1, q2, T);
            if (head > 0.0) s.l_dqdh[r.ix] = weir_dqdh(n.weir_type[r.j], dir, head, q1, q2);
        }
    }
    s.l_dqdh[r.ix] = weir_dqdh(n.weir_type[r.j], dir, head, q1, q2);
}
// ---- roadway.c:85-200: flow overtopping a roadway (FHWA HDS-5 weir with variable Cd) ------------
SWB_HD double road_curve(const double *tb, int first, int np, double x)      // getY (roadway.c:180)
{
    const double *p = tb + 2 * first;
    if (x <= p[0]) return p[1];
    if (x >= p[2 * (np - 1)]) return p[2 * (np - 1) + 1];
    for (int i = 1; i < np; i++) {
        if (x <= p[2 * i]) {
            double x1 = p[2 * i - 2], dx = p[2 * i] - x1;
            double y1 = p[2 * i - 1], dy = p[2 * i + 1] - y1;
            return y1 + (x - x1) * dy / dx;
        }
    }
    return p[2 * (np - 1) + 1];
}
SWB_HD double road_cd(const double *tb, double hWr, double ht, double roadWidth, int roadSurf)
{
    const bool paved = (roadSurf == 1);
    double kT = 1.0, cR;
    if (hWr <= 0.0) return 0.0;
    double hL = hWr / roadWidth;
    if (hL <= 0.15)
        cR = paved ? road_curve(tb, RT_CR_LOW_PAVED, RN_CR_LOW_PAVED, hWr)
                   : road_curve(tb, RT_CR_LOW_GRAVEL, RN_CR_LOW_GRAVEL, hWr);
    else
        cR = paved ? road_curve(tb, RT_CR_HIGH_PAVED, RN_CR_HIGH_PAVED, hL)
                   : road_curve(tb, RT_CR_HIGH_GRAVEL, RN_CR_HIGH_GRAVEL, hL);
    if (ht > 0.0) {
        double htH = ht / hWr;
        kT = paved ? road_curve(tb, RT_KT_PAVED, RN_KT_PAVED, htH)
                   : road_curve(tb, RT_KT_GRAVEL, RN_KT_GRAVEL, htH);
    }
    return cR * kT;
}
SWB_NI double roadway_inflow(const Net &n, const State &s, const RegCtx &r, const Xs &x, double dir,
                             double hRoad, double h1, double h2)
{
    const int j = r.j;
    double roadWidth = n.weir_road_width[j];
    int roadSurf = n.weir_road_surface[j];
    double cD = n.weir_cdisch1[j], q = 0.0, dqdh = 0.0;
    if (n.opt.unit_system == 1) cD = cD / 0.552;
    bool useVariableCd = (roadWidth > 0.0 && roadSurf >= 1);
    double hWr = h1 - hRoad, ht = h2 - hRoad;
    if (hWr > SWB_FUDGE) {
        if (useVariableCd) cD = road_cd(n.road_tables, hWr, ht, roadWidth, roadSurf);
        double length = x.wMax;
        q = cD * length * pow(hWr, 1.5);
        dqdh = 1.5 * q / hWr;
    }
    s.l_dqdh[r.ix] = dqdh;
    s.l_depth[r.ix] = SWB_MAX(h1 - hRoad, 0.0);
    int cls = SWB_SUBCRITICAL;
    if (hRoad > h2) cls = (dir == 1.0) ? SWB_DN_CRITICAL : SWB_UP_CRITICAL;
    s.l_flow_class[r.ix] = (unsigned char)cls;
    return dir * q;
}
SWB_HD double weir_inflow(const Net &n, const State &s, const RegCtx &r, const double *T)
{
    const int j = r.j;
    const Xs x = load_xs(n, j);
    double h1 = r.depth1 + r.inv1, h2 = r.depth2 + r.inv2, head, q1, q2, y;
    double dir = (h1 > h2) ? +1.0 : -1.0;
    if (dir < 0.0) { head = h1; h1 = h2; h2 = head; }
    double hcrest = r.inv1 + n.link_offset1[j];
    double hcrown = hcrest + x.yFull;
    if (n.weir_type[j] == SWB_ROADWAY_WEIR) return roadway_inflow(n, s, r, x, dir, hcrest, h1, h2);
    double setting = s.l_setting[r.ix];
    hcrest += (1.0 - setting) * x.yFull;
    head = h1 - hcrest;
    s.l_dqdh[r.ix] = 0.0;
    if (head <= SWB_FUDGE || hcrest >= hcrown || link_flap_closed(r.flags, n.link_direction[j], dir)) {
        s.l_depth[r.ix] = 0.0;
        s.l_flow_class[r.ix] = SWB_DRY;
        return 0.0;
    }
    int cls = SWB_SUBCRITICAL;
    if (hcrest > h2) cls = (dir == 1.0) ? SWB_DN_CRITICAL : SWB_UP_CRITICAL;
    s.l_flow_class[r.ix] = (unsigned char)cls;
    y = x.yFull - (hcrown - SWB_MIN(h1, hcrown));
    s.r_surf_area[r.ix] = xs_w_of_y_ni(x, y, T) * n.weir_length[j];
    bool hasFlap = (r.flags & LF_HAS_FLAP) != 0;
    if (h1 >= hcrown) {
        if (n.weir_can_surcharge[j]) {
            y = (hcrest + hcrown) / 2.0;
            if (h2 < y) head = h1 - y; else head = h1 - h2;
            y = hcrown - hcrest;
            // weir_getOrificeFlow (link.c:2442-2468)
            double cOrif = s.w_csurcharge[r.ix];
            double q = cOrif * sqrt(head);
            if (hasFlap) {
                double a = weir_open_area(s, r, x, y, T);
                if (a > 0.0) {
                    double v = q / a;
                    double hloss = (4.0 / SWB_GRAVITY) * v * v * exp(-1.15 * v / sqrt(y));
                    head -= hloss;
                    head = SWB_MAX(head, 0.0);
                    q = cOrif * sqrt(head);
                }
            }
            if (head > 0.0) s.l_dqdh[r.ix] = q / (2.0 * head); else s.l_dqdh[r.ix] = 0.0;
            s.l_depth[r.ix] = y;
            return dir * q;
        }
        else head = hcrown - hcrest;
    }
    weir_flow(n, s, r, x, head, dir, hasFlap, q1, q2, T);
    if (h2 > hcrest) {
        double ratio = (h2 - hcrest) / (h1 - hcrest);
        const double weirPower[4] = {1.5, 5. / 3., 2.5, 1.5};
        q1 *= pow((1.0 - pow(ratio, weirPower[n.weir_type[j]])), 0.385);
        if (q2 > 0.0) q2 *= pow((1.0 - pow(ratio, weirPower[2])), 0.385);
    }
    s.l_depth[r.ix] = SWB_MIN((h1 - hcrest), x.yFull);
    return dir * (q1 + q2);
}

// ---- outlet_getInflow / getFlow (link.c:2608-2692) -------------------------------------------------
SWB_HD double outlet_inflow(const Net &n, const State &s, const RegCtx &r)
{
    const int j = r.j;
    double h1 = r.depth1 + r.inv1, h2 = r.depth2 + r.inv2, head, y1;
    double dir = (h1 >= h2) ? +1.0 : -1.0;
    y1 = r.depth1;
    if (dir < 0.0) { y1 = h1; h1 = h2; h2 = y1; y1 = r.depth2; }
    double hcrest = r.inv1 + n.link_offset1[j];
    if (n.outlet_curve_type[j] == 1 /*NODE_HEAD*/) head = h1 - SWB_MAX(h2, hcrest);
    else head = h1 - hcrest;
    if (head <= SWB_FUDGE || y1 <= SWB_FUDGE || link_flap_closed(r.flags, n.link_direction[j], dir)) {
        s.l_depth[r.ix] = 0.0;
        s.l_flow_class[r.ix] = SWB_DRY;
        return 0.0;
    }
    s.l_depth[r.ix] = head;
    s.l_flow_class[r.ix] = SWB_SUBCRITICAL;
    double h = head * n.opt.ucf_length, qf;
    int c = n.outlet_curve[j];
    if (c >= 0) qf = curve_lookup(n, c, h) / n.opt.ucf_flow;
    else qf = n.outlet_qcoeff[j] * pow(h, n.outlet_qexpon[j]) / n.opt.ucf_flow;
    return dir * s.l_setting[r.ix] * qf;
}

// ---- node_getMaxOutflow (node.c:418-434) ------------------------------------------------------------
SWB_HD double node_max_outflow(const Net &n, const State &s, int i, size_t ixn, double q, double dt)
{
    if (n.node_full_volume[i] > 0.0) {
        double qMax = s.n_inflow[ixn] + s.n_old_volume[ixn] / dt;
        if (q > qMax) q = qMax;
    }
    return SWB_MAX(0.0, q);
}

// ---- findNonConduitFlow + updateNodeFlows for one member, ascending link order ----------------------
SWB_NI void regulator_pass(const Net &n, const State &s, int m, int steps, double dt, const double *T)
{
    const int M = s.M;
    for (int k = 0; k < n.nNonConduit; k++) {
        RegCtx r;
        r.j = n.nc_links[k]; r.m = m;
        const int j = r.j;
        r.n1 = n.link_node1[j]; r.n2 = n.link_node2[j]; r.flags = n.link_flags[j];
        r.ix = SWB_IX(j, m, M); r.ix1 = SWB_IX(r.n1, m, M); r.ix2 = SWB_IX(r.n2, m, M);
        r.depth1 = s.n_depth[r.ix1]; r.depth2 = s.n_depth[r.ix2];
        r.inv1 = n.node_invert[r.n1]; r.inv2 = n.node_invert[r.n2];
        const int type = n.link_type[j];

        bool bypassed = false;
        if (steps >= 2) bypassed = s.n_converged[r.ix1] && s.n_converged[r.ix2];
        s.l_bypassed[r.ix] = bypassed ? 1 : 0;
        if (!bypassed) {
            // findNonConduitFlow (dynwave.c:423-454)
            double qLast = s.l_flow[r.ix], qNew;
            s.l_dqdh[r.ix] = 0.0;
            if (s.l_setting[r.ix] == 0) qNew = 0.0;                   // link_getInflow (link.c:550)
            else switch (type) {
              case SWB_CONDUIT: {                                      // dummy conduit
                qNew = s.n_inflow[r.ix1] + s.n_overflow[r.ix1];        // node_getOutflow (node.c:400)
                double qLimit = n.link_q_limit[j];
                if (qLimit > 0.0) qNew = SWB_MIN(qNew, qLimit);
                break; }
              case SWB_PUMP:    qNew = pump_inflow(n, s, r); break;
              case SWB_ORIFICE: qNew = orifice_inflow(n, s, r, T); break;
              case SWB_WEIR:    qNew = weir_inflow(n, s, r, T); break;
              case SWB_OUTLET:  qNew = outlet_inflow(n, s, r); break;
              default:          qNew = s.n_inflow[r.ix1] + s.n_overflow[r.ix1];
            }
            if (type == SWB_PUMP && qNew != 0.0) {
                // getModPumpFlow (dynwave.c:458-499)
                if (n.node_type[r.n1] == SWB_STORAGE) qNew = node_max_outflow(n, s, r.n1, r.ix1, qNew, dt);
                else switch (n.pump_type[j]) {
                  case 0: qNew = node_max_outflow(n, s, r.n1, r.ix1, qNew, dt); break;
                  case 1: case 2: case 3: {
                    double newNetInflow = s.n_inflow[r.ix1] - s.n_outflow[r.ix1] - qNew;
                    double netFlowVolume = 0.5 * (s.n_old_net_inflow[r.ix1] + newNetInflow) * dt;
                    double y = s.n_old_depth[r.ix1] + netFlowVolume / s.n_new_surf_area[r.ix1];
                    if (y <= 0.0) qNew = s.n_inflow[r.ix1];
                    break; }
                }
            }
            // findNonConduitSurfArea (dynwave.c:503-524)
            double sa1 = 0.0;
            if (type == SWB_ORIFICE) sa1 = s.r_surf_area[r.ix] / 2.;
            double sa2 = sa1;
            int cls = s.l_flow_class[r.ix];
            if (cls == SWB_UP_CRITICAL || (r.flags & LF_N1_STORAGE)) sa1 = 0.0;
            if (cls == SWB_DN_CRITICAL || (r.flags & LF_N2_STORAGE)) sa2 = 0.0;
            s.l_surf_area1[r.ix] = sa1;
            s.l_surf_area2[r.ix] = sa2;
            if (steps > 0 && type != SWB_PUMP) {
                qNew = (1.0 - SWB_OMEGA) * qLast + SWB_OMEGA * qNew;
                if (qNew * qLast < 0.0) qNew = 0.001 * SWB_SGN(qNew);
            }
            s.l_flow[r.ix] = qNew;
        }
        // updateNodeFlows (dynwave.c:528-589) for both end nodes
        NodeAcc a1, a2;
        a1.inflow = s.n_inflow[r.ix1]; a1.outflow = s.n_outflow[r.ix1];
        a1.surfArea = s.n_new_surf_area[r.ix1]; a1.sumdqdh = s.n_sumdqdh[r.ix1];
        node_add_link_end(n, s, j, 0, m, a1);
        s.n_inflow[r.ix1] = a1.inflow; s.n_outflow[r.ix1] = a1.outflow;
        s.n_new_surf_area[r.ix1] = a1.surfArea; s.n_sumdqdh[r.ix1] = a1.sumdqdh;
        a2.inflow = s.n_inflow[r.ix2]; a2.outflow = s.n_outflow[r.ix2];
        a2.surfArea = s.n_new_surf_area[r.ix2]; a2.sumdqdh = s.n_sumdqdh[r.ix2];
        node_add_link_end(n, s, j, 1, m, a2);
        s.n_inflow[r.ix2] = a2.inflow; s.n_outflow[r.ix2] = a2.outflow;
        s.n_new_surf_area[r.ix2] = a2.surfArea; s.n_sumdqdh[r.ix2] = a2.sumdqdh;
    }
}

} // namespace swb
#endif
