// swb_culvert.h -- the two conduit special cases of the momentum equation:
//   * force mains flowing full: Hazen-Williams / Darcy-Weisbach friction slope
//     (forcmain.c:95-175, called from dwflow.c:212-213);
//   * culvert-coded conduits: FHWA HDS-5 inlet control (culvert.c:174-398, called from
//     dwflow.c:251-252), including the Ridder root search of findroot.c:91-138 for Form-1 inlets.
// Both are reached only from the generic conduit function, so the two specialised hot instances
// (plain circular / closed rectangular pipes) carry none of this code.
#ifndef SWB_CULVERT_H
#define SWB_CULVERT_H

#include "swb_state.h"
#include "swb_xsect.h"
#include "swb_hds5_tables.h"

namespace swb {

enum { SWB_FM_HAZEN_WILLIAMS = 0, SWB_FM_DARCY_WEISBACH = 1 };      // enums.h:341-343
#define SWB_BIG     1.E10                                             // consts.h:32
#define SWB_VISCOS  1.1E-5                                            // forcmain.c:19

// Swamee-Jain friction factor with the laminar branch and the 2000 < Re < 4000 blend
// (forcmain.c:130-175); the reference's one-level recursion is the turbulent formula at Re = 4000.
SWB_HD double fm_turbulent_f(double e, double hrad, double re)
{
    double f = e / 3.7 / (4.0 * hrad);
    if (re < 1.0e10) f += 5.74 / pow(re, 0.9);
    f = log10(f);
    return 0.25 / f / f;
}
SWB_HD double fm_fric_factor(double e, double hrad, double re)
{
    if (re < 10.0) re = 10.0;
    if (re <= 2000.0) return 64.0 / re;
    if (re < 4000.0) {
        double f = fm_turbulent_f(e, hrad, 4000.0);
        return 0.032 + (f - 0.032) * (re - 2000.0) / 2000.0;
    }
    return fm_turbulent_f(e, hrad, re);
}
// forcemain_getFricSlope (forcmain.c:95-118): xs.sBot / xs.rBot hold the force-main factors set
// up by link validation (link.c:1125-1128), which the flattened descriptor already carries
SWB_NI double forcemain_fric_slope(const Net &n, const Xs &x, double v, double hrad)
{
    switch (n.opt.force_main_eqn) {
      case SWB_FM_HAZEN_WILLIAMS:
        return x.sBot * pow(v, 0.852) / pow(hrad, 1.1667);
      case SWB_FM_DARCY_WEISBACH: {
        double re = 4.0 * hrad * v / SWB_VISCOS;
        double f = fm_fric_factor(x.rBot, hrad, re);
        return f * x.sBot * v / hrad;
      }
    }
    return 0.0;
}

// ---- culvert inlet control ---------------------------------------------------------------------
struct Culvert {
    const double *par;          // FORM, K, M, C, Y of this culvert code
    double yFull, scf, dQdH, qc, kk, mm, ad, hPlus;
};
enum { CP_FORM = 0, CP_K, CP_M, CP_C, CP_Y };

// residual of FHWA equation Form 1 at trial critical depth yc (culvert.c:367-398); leaves the
// matching critical flow in c.qc as the reference's callback does
SWB_HD double culvert_form1_residual(Culvert &c, const Xs &x, double yc, const double *T)
{
    double ac = xs_a_of_y_ni(x, yc, T);
    double wc = xs_w_of_y_ni(x, yc, T);
    double yh = ac / wc;
    c.qc = ac * sqrt(SWB_GRAVITY * yh);
    return c.hPlus - yc / c.yFull - yh / 2.0 / c.yFull - c.kk * pow(c.qc / c.ad, c.mm);
}
#define SWB_SIGN(a, b) ((b) >= 0.0 ? fabs(a) : -fabs(a))              // findroot.c:16
// Ridder bracketing search on [0.01h, h] to 0.001 ft (culvert.c:356, findroot.c:91-138).  Only the
// side effect matters: the flow of the LAST residual evaluation.
SWB_HD void culvert_form1_solve(Culvert &c, const Xs &x, double h, const double *T)
{
    const double x1 = 0.01 * h, x2 = h, xacc = 0.001;
    double flo = culvert_form1_residual(c, x, x1, T);
    double fhi = culvert_form1_residual(c, x, x2, T);
    if (flo == 0.0 || fhi == 0.0) return;
    if (!((flo > 0.0 && fhi < 0.0) || (flo < 0.0 && fhi > 0.0))) return;
    double ans = 0.5 * (x1 + x2), xlo = x1, xhi = x2;
    for (int it = 1; it <= 60; it++) {
        double xm = 0.5 * (xlo + xhi);
        double fm = culvert_form1_residual(c, x, xm, T);
        double sq = sqrt(fm * fm - flo * fhi);
        if (sq == 0.0) return;
        double xnew = xm + (xm - xlo) * ((flo >= fhi ? 1.0 : -1.0) * fm / sq);
        if (fabs(xnew - ans) <= xacc) break;
        ans = xnew;
        double fnew = culvert_form1_residual(c, x, ans, T);
        if (SWB_SIGN(fm, fnew) != fm)        { xlo = xm;  flo = fm;  xhi = ans; fhi = fnew; }
        else if (SWB_SIGN(flo, fnew) != flo) { xhi = ans; fhi = fnew; }
        else if (SWB_SIGN(fhi, fnew) != fhi) { xlo = ans; flo = fnew; }
        else return;
        if (fabs(xhi - xlo) <= xacc) return;
    }
}
// unsubmerged inlet (culvert.c:256-282)
SWB_HD double culvert_unsubmerged(Culvert &c, const Xs &x, double h, const double *T)
{
    double q;
    c.kk = c.par[CP_K];
    c.mm = c.par[CP_M];
    double arg = h / c.yFull / c.kk;
    if (c.par[CP_FORM] == 1.0) {
        c.hPlus = h / c.yFull + c.scf;            // getForm1Flow (culvert.c:344-362)
        culvert_form1_solve(c, x, h, T);
        q = c.qc;
    }
    else q = c.ad * pow(arg, 1.0 / c.mm);
    c.dQdH = q / h / c.mm;
    return q;
}
// submerged inlet (culvert.c:286-309)
SWB_HD double culvert_submerged(Culvert &c, double h)
{
    double cc = c.par[CP_C], yy = c.par[CP_Y];
    double arg = (h / c.yFull - yy + c.scf) / cc;
    if (arg <= 0.0) { c.dQdH = 0.0; return SWB_BIG; }
    double q = sqrt(arg) * c.ad;
    c.dQdH = 0.5 * q / arg / c.yFull / cc;
    return q;
}
// culvert_getInflow (culvert.c:174-252): h is the head at the conduit's upstream end; returns the
// possibly reduced flow and, when the inlet controls, replaces dqdh and raises inletControl
SWB_NI double culvert_inflow(const Net &n, const Xs &x, int j, double q0, double h, double &dqdh,
                             unsigned char &inletControl, const double *T)
{
    int code = n.xs_culvert[j];
    if (code <= 0 || code > SWB_MAX_CULVERT_CODE) return q0;
    Culvert c;
    c.par = n.culvert_params + 5 * code;
    c.yFull = x.yFull;
    c.ad = x.aFull * sqrt(c.yFull);
    c.dQdH = 0.0; c.qc = 0.0;
    switch (code) {                               // mitered inlets
      case 5: case 37: case 46: c.scf = -7.0 * n.cond_slope[j]; break;
      default:                  c.scf = 0.5 * n.cond_slope[j];
    }
    double y = h - n.link_z1[j];
    double y2 = c.yFull * (16.0 * c.par[CP_C] + c.par[CP_Y] - c.scf);
    double q;
    if (y >= y2) q = culvert_submerged(c, y);
    else {
        double y1 = 0.95 * c.yFull;
        if (y <= y1) q = culvert_unsubmerged(c, x, y, T);
        else {                                    // transition zone (culvert.c:313-330)
            double q1 = culvert_unsubmerged(c, x, y1, T);
            double q2 = culvert_submerged(c, y2);
            q = q1 + (q2 - q1) * (y - y1) / (y2 - y1);
            c.dQdH = (q2 - q1) / (y2 - y1);
        }
    }
    if (q < q0) {
        inletControl = 1;
        dqdh = c.dQdH;
        return q;
    }
    return q0;
}

} // namespace swb
#endif
