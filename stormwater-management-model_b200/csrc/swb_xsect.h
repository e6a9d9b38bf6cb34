// swb_xsect.h -- cross-section geometry device library (K1b).
//
// Pure functions of a by-value parameter pack; the normalised shape tables live in ONE flat array
// (swb_xsect_tables.h) that every CTA stages into shared memory, so a table read is an LDS with a
// data-dependent index instead of a divergent constant/global load.  No function mutates its
// cross section (the reference's filled-circular helpers do, xsect.c:2462-2524; restated here on
// local copies).  Operation order follows the reference expression by expression because results
// must match its un-fused IEEE arithmetic bit for bit (compile with --fmad=false).
//
// Restates: xsect_getAofY/WofY/RofY/YofA/RofA/SofA/AofS/dSdA/Ycrit (xsect.c:714-1319),
// lookup/invLookup/locate (:1474-1608), generic_getAofS (:1359-1400), getYcritEnum/Ridder
// (:1634-1748), the per-shape closed forms (:1754-2357), circular specials (:2361-2618),
// findroot_Newton / findroot_Ridder (findroot.c:20-138).
#ifndef SWB_XSECT_H
#define SWB_XSECT_H

#include <math.h>
#include "swb_common.h"
#include "swb_xsect_tables.h"

namespace swb {

// enums.h:117-143
enum XsType {
    XS_DUMMY = 0, XS_CIRCULAR, XS_FILLED_CIRCULAR, XS_RECT_CLOSED, XS_RECT_OPEN, XS_TRAPEZOIDAL,
    XS_TRIANGULAR, XS_PARABOLIC, XS_POWERFUNC, XS_RECT_TRIANG, XS_RECT_ROUND, XS_MOD_BASKET,
    XS_HORIZ_ELLIPSE, XS_VERT_ELLIPSE, XS_ARCH, XS_EGGSHAPED, XS_HORSESHOE, XS_GOTHIC, XS_CATENARY,
    XS_SEMIELLIPTICAL, XS_BASKETHANDLE, XS_SEMICIRCULAR, XS_IRREGULAR, XS_CUSTOM, XS_FORCE_MAIN,
    XS_STREET
};

struct Xs {                 // objects.h:581-599, plus resolved per-object tables
    int    type;
    int    ntbl;            // entries in the per-object tables (IRREGULAR / CUSTOM / STREET)
    const double *atbl, *rtbl, *wtbl;
    double yFull, wMax, ywMax, aFull, rFull, sFull, sMax, yBot, aBot, sBot, rBot;
    double rYFull;          // exact_rcp(yFull): RN(1 / yFull) when y / yFull may go through div_rcp, else 0
};
SWB_FI double xs_ynorm(const Xs &x, double y) { return div_by(y, x.yFull, x.rYFull); }

// ratio of area at max. flow to full area (xsect.c:55-81); >= 1 means an open shape
SWB_FI double xs_amax_ratio(int type)
{
    switch (type) {
      case XS_CIRCULAR: case XS_FILLED_CIRCULAR: case XS_FORCE_MAIN: return 0.9756;
      case XS_RECT_CLOSED: return 0.97;
      case XS_RECT_TRIANG: case XS_RECT_ROUND: case XS_CATENARY: case XS_SEMIELLIPTICAL: return 0.98;
      case XS_MOD_BASKET: case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE: case XS_EGGSHAPED:
      case XS_HORSESHOE: case XS_GOTHIC: case XS_BASKETHANDLE: case XS_SEMICIRCULAR:
      case XS_CUSTOM: return 0.96;
      case XS_ARCH: return 0.92;
      default: return 1.0;
    }
}
SWB_FI bool xs_is_open(int type) { return xs_amax_ratio(type) >= 1.0; }

// ---- table primitives -----------------------------------------------------------------------
// xsect.c:1474-1507
SWB_FI double xs_lookup_body(double x, const double *tb, int n)
{
    double delta = 1.0 / ((double)n - 1);
    // 51- and 26-entry tables (every circular table): delta = RN(1/50), RN(1/25) pass exact_rcp()
    // with reciprocals 50 and 25, so both divisions by delta are three operations each
    // (tests/test_exact_division.py); other table sizes divide
    const bool viaRcp = (n == 51 || n == 26);
    const double rDelta = (double)n - 1;
    int i = (int)(viaRcp ? div_rcp(x, delta, rDelta) : x / delta);
    if (i >= n - 1) return tb[n - 1];
    double x0 = i * delta;
    double x1 = ((double)i + 1) * delta;
    double t0 = tb[i], t1 = tb[i + 1];
    double y = t0 + (viaRcp ? div_rcp((x - x0) * (t1 - t0), delta, rDelta) : (x - x0) * (t1 - t0) / delta);
    if (i < 2) {
        double y2 = y + (x - x0) * (x - x1) / (delta * delta) * (t0 / 2.0 - t1 + tb[i + 2] / 2.0);
        if (y2 > 0.0) y = y2;
    }
    if (y < 0.0) y = 0.0;
    return y;
}

#ifdef SWB_LOOKUP_NI
// one shared copy of the 51-entry lookup (all circular tables) instead of one per call site
template <int N> SWB_NI double xs_lookup_fixed(double x, const double *tb) { return xs_lookup_body(x, tb, N); }
#endif
SWB_FI double xs_lookup(double x, const double *tb, int n)
{
#ifdef SWB_LOOKUP_NI
    if (n == 51) return xs_lookup_fixed<51>(x, tb);
#endif
    return xs_lookup_body(x, tb, n);
}

// xsect.c:1571-1608
SWB_HD int xs_locate(double y, const double *tb, int jLast)
{
    int j1 = 0, j2 = jLast;
    if (y <= tb[0]) return 0;
    if (y >= tb[jLast]) return jLast;
    while (j2 - j1 > 1) {
        int j = (j1 + j2) >> 1;
        if (y >= tb[j]) j1 = j; else j2 = j;
    }
    return j1;
}

// xsect.c:1511-1567
SWB_HD double xs_inv_lookup(double y, const double *tb, int nItems)
{
    double dx = 1.0 / (double)((double)nItems - 1);
    int n = nItems, i;
    if (tb[n - 3] > tb[n - 1]) n = n - 2;
    if (n < nItems && y > tb[nItems - 1]) {
        if (y >= tb[nItems - 3]) return ((double)n - 1) * dx;
        if (y <= tb[nItems - 2]) i = nItems - 2;
        else i = nItems - 3;
    }
    else i = xs_locate(y, tb, n - 1);
    if (i >= n - 1) return ((double)n - 1) * dx;
    double x0 = i * dx, x;
    double dy = tb[i + 1] - tb[i];
    if (dy == 0.0) x = x0;
    else x = x0 + (y - tb[i]) * dx / dy;
    if (x < 0.0) x = 0.0;
    if (x > 1.0) x = 1.0;
    return x;
}

// ---- circular specials (xsect.c:2531-2618) ---------------------------------------------------
SWB_HD double xs_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

SWB_HD double circ_theta_of_alpha(double alpha)
{
    double theta, theta1, ap, d;
    if (alpha > 0.04) theta = 1.2 + 5.08 * (alpha - 0.04) / 0.96;
    else theta = 0.031715 - 12.79384 * alpha + 8.28479 * sqrt(alpha);
    theta1 = theta;
    ap = (2.0 * SWB_PI) * alpha;
    for (int k = 1; k <= 40; k++) {
        d = -(ap - theta + sin(theta)) / (1.0 - cos(theta));
        if (d > 1.0) d = xs_sign(1.0, d);
        theta = theta - d;
        if (fabs(d) <= 0.0001) return theta;
    }
    return theta1;
}

SWB_HD double circ_theta_of_psi(double psi)
{
    double theta, theta1, ap, tt, tt23, t3, d;
    if      (psi > 0.90)  theta = 4.17 + 1.12 * (psi - 0.90) / 0.176;
    else if (psi > 0.5)   theta = 3.14 + 1.03 * (psi - 0.5) / 0.4;
    else if (psi > 0.015) theta = 1.2 + 1.94 * (psi - 0.015) / 0.485;
    else                  theta = 0.12103 - 55.5075 * psi + 15.62254 * sqrt(psi);
    theta1 = theta;
    ap = (2.0 * SWB_PI) * psi;
    for (int k = 1; k <= 40; k++) {
        theta = fabs(theta);
        tt = theta - sin(theta);
        tt23 = pow(tt, 2. / 3.);
        t3 = pow(theta, 1. / 3.);
        d = ap * theta / t3 - tt * tt23;
        d = d / (ap * (2. / 3.) / t3 - (5. / 3.) * tt23 * (1.0 - cos(theta)));
        theta = theta - d;
        if (fabs(d) <= 0.0001) return theta;
    }
    return theta1;
}

SWB_HD double circ_y_norm(double alpha)            // getYcircular
{
    if (alpha >= 1.0) return 1.0;
    if (alpha <= 0.0) return 0.0;
    if (alpha <= 1.0e-5) {
        double theta = pow(37.6911 * alpha, 1. / 3.);
        return theta * theta / 16.0;
    }
    double theta = circ_theta_of_alpha(alpha);
    return (1.0 - cos(theta / 2.)) / 2.0;
}

SWB_HD double circ_s_norm(double alpha)            // getScircular
{
    if (alpha >= 1.0) return 1.0;
    if (alpha <= 0.0) return 0.0;
    if (alpha <= 1.0e-5) {
        double theta = pow(37.6911 * alpha, 1. / 3.);
        return pow(theta, 13. / 3.) / 124.4797;
    }
    double theta = circ_theta_of_alpha(alpha);
    return pow((theta - sin(theta)), 5. / 3.) / (2.0 * SWB_PI) / pow(theta, 2. / 3.);
}

SWB_HD double circ_a_norm(double psi)              // getAcircular
{
    if (psi >= 1.0) return 1.0;
    if (psi <= 0.0) return 0.0;
    if (psi <= 1.0e-6) {
        double theta = pow(124.4797 * psi, 3. / 13.);
        return theta * theta * theta / 37.6911;
    }
    double theta = circ_theta_of_psi(psi);
    return (theta - sin(theta)) / (2.0 * SWB_PI);
}

// xsect.c:2367-2376 (yFull, aFull passed so FILLED_CIRCULAR can use shifted values)
SWB_HD double circ_y_of_a(double yFull, double aFull, double a, const double *T)
{
    double alpha = a / aFull;
    if (alpha < 0.04) return yFull * circ_y_norm(alpha);
    return yFull * xs_lookup(alpha, T + XT_Y_CIRC, XN_Y_CIRC);
}

// ---- forward declarations ----------------------------------------------------------------------
// S >= 0 names the shape at compile time: `switch (S >= 0 ? S : x.type)` is then resolved by the
// front end itself (relying on the optimiser to fold a struct member proved unreliable: the member
// was promoted to a constant only after the last CFG simplification, leaving every shape's code
// inside the specialised conduit functions).  S < 0 (the default) dispatches at run time.
template <int S = -1> SWB_FI double xs_a_of_y(const Xs &x, double y, const double *T);
template <int S = -1> SWB_FI double xs_w_of_y(const Xs &x, double y, const double *T);
SWB_HD double xs_y_of_a(const Xs &x, double a, const double *T);
// The reference's getRofY / getRofA / getSofA call one another through their default branches
// (xsect.c:1094,1113,1137,768).  No shape ever goes round the cycle, so it is unrolled here into
// an acyclic chain r_of_y -> r_of_a -> s_of_a -> r_of_a_direct -> r_of_y_direct that the device
// compiler can inline completely (no stack frames, no recursion).

// ---- RECT_CLOSED (xsect.c:1754-1803) -----------------------------------------------------------
#define SWB_RECT_ALFMAX        0.97
#define SWB_RECT_TRIANG_ALFMAX 0.98
#define SWB_RECT_ROUND_ALFMAX  0.98

SWB_FI double rect_closed_r_of_a(const Xs &x, double a)
{
    if (a <= 0.0) return 0.0;
    double p = x.wMax + 2. * a / x.wMax;
    if (a / x.aFull > SWB_RECT_ALFMAX)
        p += (a / x.aFull - SWB_RECT_ALFMAX) / (1.0 - SWB_RECT_ALFMAX) * x.wMax;
    return a / p;
}

// ---- RECT_TRIANG (xsect.c:1836-1931) -----------------------------------------------------------
SWB_HD double rect_triang_y_of_a(const Xs &x, double a)
{
    if (a <= x.aBot) return sqrt(a / x.sBot);
    return x.yBot + (a - x.aBot) / x.wMax;
}
SWB_HD double rect_triang_r_of_a(const Xs &x, double a)
{
    if (a <= 0.0) return 0.0;
    double y = rect_triang_y_of_a(x, a);
    if (y <= x.yBot) return a / (2. * y * x.rBot);
    double p = 2. * x.yBot * x.rBot + 2. * (y - x.yBot);
    double alf = (a / x.aFull) - SWB_RECT_TRIANG_ALFMAX;
    if (alf > 0.0) p += alf / (1.0 - SWB_RECT_TRIANG_ALFMAX) * x.wMax;
    return a / p;
}
SWB_HD double rect_triang_r_of_y(const Xs &x, double y)
{
    if (y <= x.yBot) return y * x.sBot / (2. * x.rBot);
    double a = x.aBot + (y - x.yBot) * x.wMax;
    double p = 2. * x.yBot * x.rBot + 2. * (y - x.yBot);
    double alf = (a / x.aFull) - SWB_RECT_TRIANG_ALFMAX;
    if (alf > 0.0) p += alf / (1.0 - SWB_RECT_TRIANG_ALFMAX) * x.wMax;
    return a / p;
}

// ---- RECT_ROUND (xsect.c:1938-2072) ------------------------------------------------------------
SWB_HD double rect_round_y_of_a(const Xs &x, double a, const double *T)
{
    if (a > x.aBot) return x.yBot + (a - x.aBot) / x.wMax;
    double alpha = a / (SWB_PI * x.rBot * x.rBot);
    if (alpha < 0.04) return (2.0 * x.rBot) * circ_y_norm(alpha);
    return (2.0 * x.rBot) * xs_lookup(alpha, T + XT_Y_CIRC, XN_Y_CIRC);
}
SWB_HD double rect_round_r_of_a(const Xs &x, double a, const double *T)
{
    double y1, theta1, p, arg;
    if (a <= 0.0) return 0.0;
    if (a > x.aBot) {
        y1 = (a - x.aBot) / x.wMax;
        theta1 = 2.0 * asin(x.wMax / 2.0 / x.rBot);
        p = x.rBot * theta1 + 2.0 * y1;
        arg = (a / x.aFull) - SWB_RECT_ROUND_ALFMAX;
        if (arg > 0.0) p += arg / (1.0 - SWB_RECT_ROUND_ALFMAX) * x.wMax;
        return a / p;
    }
    y1 = rect_round_y_of_a(x, a, T);
    theta1 = 2.0 * acos(1.0 - y1 / x.rBot);
    p = x.rBot * theta1;
    return a / p;
}
SWB_HD double rect_round_a_of_y(const Xs &x, double y)
{
    if (y > x.yBot) return x.aBot + (y - x.yBot) * x.wMax;
    double theta1 = 2.0 * acos(1.0 - y / x.rBot);
    return 0.5 * x.rBot * x.rBot * (theta1 - sin(theta1));
}

// ---- MOD_BASKET (xsect.c:2082-2173) ------------------------------------------------------------
SWB_HD double mod_basket_y_of_a(const Xs &x, double a, const double *T)
{
    if (a <= x.aFull - x.aBot) return a / x.wMax;
    double alpha = (x.aFull - a) / (SWB_PI * x.rBot * x.rBot), y1;
    if (alpha < 0.04) y1 = circ_y_norm(alpha);
    else              y1 = xs_lookup(alpha, T + XT_Y_CIRC, XN_Y_CIRC);
    y1 = 2.0 * x.rBot * y1;
    return x.yFull - y1;
}
SWB_HD double mod_basket_r_of_a(const Xs &x, double a, const double *T)
{
    if (a <= x.aFull - x.aBot) return a / (x.wMax + 2.0 * a / x.wMax);
    double y1 = x.yFull - mod_basket_y_of_a(x, a, T);
    double theta1 = 2.0 * acos(1.0 - y1 / x.rBot);
    double p = (x.sBot - theta1) * x.rBot;
    y1 = x.yFull - x.yBot;
    p = p + 2.0 * y1 + x.wMax;
    return a / p;
}

// ---- TRAPEZOIDAL / TRIANGULAR / PARABOLIC / POWERFUNC (xsect.c:2184-2357) ----------------------
SWB_HD double trapez_y_of_a(const Xs &x, double a)
{
    if (x.sBot == 0.0) return a / x.yBot;
    return (sqrt(x.yBot * x.yBot + 4. * x.sBot * a) - x.yBot) / (2. * x.sBot);
}
SWB_HD double trapez_a_of_y(const Xs &x, double y) { return (x.yBot + x.sBot * y) * y; }
SWB_HD double parab_p_of_y(const Xs &x, double y)
{
    double xx = 2. * sqrt(y) / x.rBot;
    double t = sqrt(1.0 + xx * xx);
    return 0.5 * x.rBot * x.rBot * (xx * t + log(xx + t));
}
SWB_HD double parab_a_of_y(const Xs &x, double y) { return (4. / 3. * x.rBot * y * sqrt(y)); }
SWB_HD double parab_y_of_a(const Xs &x, double a) { return pow((3. / 4.) * a / x.rBot, 2. / 3.); }
SWB_HD double powerfunc_p_of_y(const Xs &x, double y)
{
    double dy1 = 0.02 * x.yFull;
    double h = (x.sBot + 1.0) * x.rBot / 2.0;
    double m = x.sBot, p = 0.0, y1 = 0.0, x1 = 0.0, x2, y2, dx, dy;
    do {
        y2 = y1 + dy1;
        if (y2 > y) y2 = y;
        x2 = h * pow(y2, m);
        dx = x2 - x1;
        dy = y2 - y1;
        p += sqrt(dx * dx + dy * dy);
        x1 = x2;
        y1 = y2;
    } while (y2 < y);
    return 2.0 * p;
}
SWB_HD double powerfunc_a_of_y(const Xs &x, double y) { return x.rBot * pow(y, x.sBot + 1.0); }
SWB_HD double powerfunc_y_of_a(const Xs &x, double a) { return pow(a / x.rBot, 1.0 / (x.sBot + 1.0)); }

// ---- FILLED_CIRCULAR (xsect.c:2462-2524) on shifted local values -------------------------------
SWB_HD double filled_circ_r_of_y(const Xs &x, double y, const double *T)
{
    double yF = x.yFull + x.yBot, aF = x.aFull + x.aBot;
    y += x.yBot;
    double a = aF * xs_lookup(y / yF, T + XT_A_CIRC, XN_A_CIRC);
    double r = 0.25 * yF * xs_lookup(y / yF, T + XT_R_CIRC, XN_R_CIRC);
    double p = (a / r);
    a = a - x.aBot;
    p = p - x.rBot + x.sBot;
    r = a / p;
    return r;
}

// ---- per-shape table descriptor for the "tabulated" families -----------------------------------
struct XsTabs { short aO, aN, rO, rN, yO, yN, sO, sN, wO, wN; };
#define SWB_T(NAME) (short)XT_##NAME, (short)XN_##NAME
SWB_FI XsTabs xs_tabs(int type)
{
    switch (type) {
      case XS_EGGSHAPED:      return XsTabs{SWB_T(A_EGG), SWB_T(R_EGG), SWB_T(Y_EGG), SWB_T(S_EGG), SWB_T(W_EGG)};
      case XS_HORSESHOE:      return XsTabs{SWB_T(A_HORSESHOE), SWB_T(R_HORSESHOE), SWB_T(Y_HORSESHOE), SWB_T(S_HORSESHOE), SWB_T(W_HORSESHOE)};
      case XS_BASKETHANDLE:   return XsTabs{SWB_T(A_BASKETHANDLE), SWB_T(R_BASKETHANDLE), SWB_T(Y_BASKETHANDLE), SWB_T(S_BASKETHANDLE), SWB_T(W_BASKETHANDLE)};
      case XS_GOTHIC:         return XsTabs{-1, 0, -1, 0, SWB_T(Y_GOTHIC), SWB_T(S_GOTHIC), SWB_T(W_GOTHIC)};
      case XS_CATENARY:       return XsTabs{-1, 0, -1, 0, SWB_T(Y_CATENARY), SWB_T(S_CATENARY), SWB_T(W_CATENARY)};
      case XS_SEMIELLIPTICAL: return XsTabs{-1, 0, -1, 0, SWB_T(Y_SEMIELLIP), SWB_T(S_SEMIELLIP), SWB_T(W_SEMIELLIP)};
      case XS_SEMICIRCULAR:   return XsTabs{-1, 0, -1, 0, SWB_T(Y_SEMICIRC), SWB_T(S_SEMICIRC), SWB_T(W_SEMICIRC)};
      case XS_HORIZ_ELLIPSE:  return XsTabs{SWB_T(A_HORIZELLIPSE), SWB_T(R_HORIZELLIPSE), -1, 0, -1, 0, SWB_T(W_HORIZELLIPSE)};
      case XS_VERT_ELLIPSE:   return XsTabs{SWB_T(A_VERTELLIPSE), SWB_T(R_VERTELLIPSE), -1, 0, -1, 0, SWB_T(W_VERTELLIPSE)};
      case XS_ARCH:           return XsTabs{SWB_T(A_ARCH), SWB_T(R_ARCH), -1, 0, -1, 0, SWB_T(W_ARCH)};
      default:                return XsTabs{-1, 0, -1, 0, -1, 0, -1, 0, -1, 0};
    }
}

// ---- A(y)  (xsect.c:857-939) -------------------------------------------------------------------
template <int S>
SWB_FI double xs_a_of_y(const Xs &x, double y, const double *T)
{
    double yNorm = xs_ynorm(x, y);
    if (y <= 0.0) return 0.0;
    switch (S >= 0 ? S : x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR:    return x.aFull * xs_lookup(yNorm, T + XT_A_CIRC, XN_A_CIRC);
      case XS_FILLED_CIRCULAR: {
        double yF = x.yFull + x.yBot, aF = x.aFull + x.aBot;
        double a = aF * xs_lookup((y + x.yBot) / yF, T + XT_A_CIRC, XN_A_CIRC);
        return a - x.aBot;
      }
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_BASKETHANDLE:
      case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE: case XS_ARCH: {
        XsTabs t = xs_tabs(x.type);
        return x.aFull * xs_lookup(yNorm, T + t.aO, t.aN);
      }
      case XS_GOTHIC: case XS_CATENARY: case XS_SEMIELLIPTICAL: case XS_SEMICIRCULAR: {
        XsTabs t = xs_tabs(x.type);
        return x.aFull * xs_inv_lookup(yNorm, T + t.yO, t.yN);
      }
      case XS_IRREGULAR: case XS_CUSTOM: case XS_STREET:
        return x.aFull * xs_lookup(yNorm, x.atbl, x.ntbl);
      case XS_RECT_CLOSED: return y * x.wMax;
      case XS_RECT_TRIANG:
        if (y <= x.yBot) return y * y * x.sBot;
        return x.aBot + (y - x.yBot) * x.wMax;
      case XS_RECT_ROUND:  return rect_round_a_of_y(x, y);
      case XS_RECT_OPEN:   return y * x.wMax;
      case XS_MOD_BASKET: {
        if (y <= x.yFull - x.yBot) return y * x.wMax;
        double y1 = x.yFull - y;
        double theta1 = 2.0 * acos(1.0 - y1 / x.rBot);
        double a1 = 0.5 * x.rBot * x.rBot * (theta1 - sin(theta1));
        return x.aFull - a1;
      }
      case XS_TRAPEZOIDAL: return trapez_a_of_y(x, y);
      case XS_TRIANGULAR:  return y * y * x.sBot;
      case XS_PARABOLIC:   return parab_a_of_y(x, y);
      case XS_POWERFUNC:   return powerfunc_a_of_y(x, y);
      default:             return 0.0;
    }
}

// ---- W(y)  (xsect.c:943-1027) ------------------------------------------------------------------
template <int S>
SWB_FI double xs_w_of_y(const Xs &x, double y, const double *T)
{
    double yNorm = xs_ynorm(x, y);
    switch (S >= 0 ? S : x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR:    return x.wMax * xs_lookup(yNorm, T + XT_W_CIRC, XN_W_CIRC);
      case XS_FILLED_CIRCULAR:
        yNorm = (y + x.yBot) / (x.yFull + x.yBot);
        return x.wMax * xs_lookup(yNorm, T + XT_W_CIRC, XN_W_CIRC);
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_BASKETHANDLE: case XS_GOTHIC: case XS_CATENARY:
      case XS_SEMIELLIPTICAL: case XS_SEMICIRCULAR: case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE:
      case XS_ARCH: {
        XsTabs t = xs_tabs(x.type);
        return x.wMax * xs_lookup(yNorm, T + t.wO, t.wN);
      }
      case XS_IRREGULAR: case XS_CUSTOM: case XS_STREET:
        return x.wMax * xs_lookup(yNorm, x.wtbl, x.ntbl);
      case XS_RECT_CLOSED:
        if (yNorm == 1.0) return 0.0;
        return x.wMax;
      case XS_RECT_TRIANG:
        if (y <= x.yBot) return 2.0 * x.sBot * y;
        return x.wMax;
      case XS_RECT_ROUND:
        if (y > x.yBot) return x.wMax;
        return 2.0 * sqrt(y * (2.0 * x.rBot - y));
      case XS_RECT_OPEN:   return x.wMax;
      case XS_MOD_BASKET: {
        if (y <= 0.0) return 0.0;
        if (y <= x.yFull - x.yBot) return x.wMax;
        double y1 = x.yFull - y;
        return 2.0 * sqrt(y1 * (2.0 * x.rBot - y1));
      }
      case XS_TRAPEZOIDAL: return x.yBot + 2.0 * y * x.sBot;
      case XS_TRIANGULAR:  return 2.0 * x.sBot * y;
      case XS_PARABOLIC:   return 2.0 * x.rBot * sqrt(y);
      case XS_POWERFUNC:   return (x.sBot + 1.0) * x.rBot * pow(y, x.sBot);
      default:             return 0.0;
    }
}

// ---- R(y)  (xsect.c:1031-1096) -----------------------------------------------------------------
template <int S = -1>
SWB_FI double xs_r_of_y_direct(const Xs &x, double y, const double *T)
{
    double yNorm = xs_ynorm(x, y);
    switch (S >= 0 ? S : x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR:    return x.rFull * xs_lookup(yNorm, T + XT_R_CIRC, XN_R_CIRC);
      case XS_FILLED_CIRCULAR:
        if (x.yBot == 0.0) return x.rFull * xs_lookup(yNorm, T + XT_R_CIRC, XN_R_CIRC);
        return filled_circ_r_of_y(x, y, T);
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_BASKETHANDLE:
      case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE: case XS_ARCH: {
        XsTabs t = xs_tabs(x.type);
        return x.rFull * xs_lookup(yNorm, T + t.rO, t.rN);
      }
      case XS_IRREGULAR: case XS_CUSTOM: case XS_STREET:
        return x.rFull * xs_lookup(yNorm, x.rtbl, x.ntbl);
      case XS_RECT_TRIANG: return rect_triang_r_of_y(x, y);
      case XS_RECT_ROUND: {
        if (y <= 0.0) return 0.0;
        if (y > x.yBot) return rect_round_r_of_a(x, rect_round_a_of_y(x, y), T);
        double theta1 = 2.0 * acos(1.0 - y / x.rBot);
        return 0.5 * x.rBot * (1.0 - sin(theta1)) / theta1;
      }
      case XS_TRAPEZOIDAL:
        if (y == 0.0) return 0.0;
        return trapez_a_of_y(x, y) / (x.yBot + y * x.rBot);
      case XS_TRIANGULAR:  return (y * x.sBot) / (2. * x.rBot);
      case XS_PARABOLIC:
        if (y <= 0.0) return 0.0;
        return parab_a_of_y(x, y) / parab_p_of_y(x, y);
      case XS_POWERFUNC:
        if (y <= 0.0) return 0.0;
        return powerfunc_a_of_y(x, y) / powerfunc_p_of_y(x, y);
      default:             return 0.0;     // shapes that go through R(A): see xs_r_of_y below
    }
}

// ---- Y(a)  (xsect.c:776-853) -------------------------------------------------------------------
SWB_HD double xs_y_of_a(const Xs &x, double a, const double *T)
{
    double alpha = a / x.aFull;
    switch (x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR:    return circ_y_of_a(x.yFull, x.aFull, a, T);
      case XS_FILLED_CIRCULAR: {
        double y = circ_y_of_a(x.yFull + x.yBot, x.aFull + x.aBot, a + x.aBot, T);
        return y - x.yBot;
      }
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_BASKETHANDLE: case XS_GOTHIC: case XS_CATENARY:
      case XS_SEMIELLIPTICAL: case XS_SEMICIRCULAR: {
        XsTabs t = xs_tabs(x.type);
        return x.yFull * xs_lookup(alpha, T + t.yO, t.yN);
      }
      case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE: case XS_ARCH: {
        XsTabs t = xs_tabs(x.type);
        return x.yFull * xs_inv_lookup(alpha, T + t.aO, t.aN);
      }
      case XS_IRREGULAR: case XS_CUSTOM: case XS_STREET:
        return x.yFull * xs_inv_lookup(alpha, x.atbl, x.ntbl);
      case XS_RECT_CLOSED: return a / x.wMax;
      case XS_RECT_TRIANG: return rect_triang_y_of_a(x, a);
      case XS_RECT_ROUND:  return rect_round_y_of_a(x, a, T);
      case XS_RECT_OPEN:   return a / x.wMax;
      case XS_MOD_BASKET:  return mod_basket_y_of_a(x, a, T);
      case XS_TRAPEZOIDAL: return trapez_y_of_a(x, a);
      case XS_TRIANGULAR:  return sqrt(a / x.sBot);
      case XS_PARABOLIC:   return parab_y_of_a(x, a);
      case XS_POWERFUNC:   return powerfunc_y_of_a(x, a);
      default:             return 0.0;
    }
}

// ---- R(a)  (xsect.c:1100-1142) -----------------------------------------------------------------
template <int S = -1>
SWB_FI double xs_r_of_a_direct(const Xs &x, double a, const double *T)
{
    if (a <= 0.0) return 0.0;
    switch (S >= 0 ? S : x.type) {
      case XS_HORIZ_ELLIPSE: case XS_VERT_ELLIPSE: case XS_ARCH: case XS_IRREGULAR:
      case XS_FILLED_CIRCULAR: case XS_CUSTOM: case XS_STREET:
        return xs_r_of_y_direct(x, xs_y_of_a(x, a, T), T);
      case XS_RECT_CLOSED: return rect_closed_r_of_a(x, a);
      case XS_RECT_OPEN:   return a / (x.wMax + (2. - x.sBot) * a / x.wMax);
      case XS_RECT_TRIANG: return rect_triang_r_of_a(x, a);
      case XS_RECT_ROUND:  return rect_round_r_of_a(x, a, T);
      case XS_MOD_BASKET:  return mod_basket_r_of_a(x, a, T);
      case XS_TRAPEZOIDAL: return a / (x.yBot + trapez_y_of_a(x, a) * x.rBot);
      case XS_TRIANGULAR:  return a / (2. * sqrt(a / x.sBot) * x.rBot);
      case XS_PARABOLIC:   return a / parab_p_of_y(x, parab_y_of_a(x, a));
      case XS_POWERFUNC:   return a / powerfunc_p_of_y(x, powerfunc_y_of_a(x, a));
      default: return 0.0;                  // shapes that go through S(A): see xs_r_of_a below
    }
}

// ---- S(a)  (xsect.c:714-772) -------------------------------------------------------------------
SWB_HD double xs_s_of_a(const Xs &x, double a, const double *T)
{
    double alpha = a / x.aFull;
    switch (x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR:
        if (alpha < 0.04) return x.sFull * circ_s_norm(alpha);
        return x.sFull * xs_lookup(alpha, T + XT_S_CIRC, XN_S_CIRC);
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_GOTHIC: case XS_CATENARY:
      case XS_SEMIELLIPTICAL: case XS_BASKETHANDLE: case XS_SEMICIRCULAR: {
        XsTabs t = xs_tabs(x.type);
        return x.sFull * xs_lookup(alpha, T + t.sO, t.sN);
      }
      case XS_RECT_CLOSED:
        if (a / x.aFull > SWB_RECT_ALFMAX)
            return x.sMax + (x.sFull - x.sMax) * (a / x.aFull - SWB_RECT_ALFMAX) / (1.0 - SWB_RECT_ALFMAX);
        return a * pow(xs_r_of_a_direct(x, a, T), 2. / 3.);
      case XS_RECT_OPEN: {
        double y = a / x.wMax;
        double r = a / ((2.0 - x.sBot) * y + x.wMax);
        return a * pow(r, 2. / 3.);
      }
      case XS_RECT_TRIANG:
        if (a / x.aFull > SWB_RECT_TRIANG_ALFMAX)
            return x.sMax + (x.sFull - x.sMax) * (a / x.aFull - SWB_RECT_TRIANG_ALFMAX) / (1.0 - SWB_RECT_TRIANG_ALFMAX);
        return a * pow(rect_triang_r_of_a(x, a), 2. / 3.);
      case XS_RECT_ROUND:
        if (a / x.aFull > SWB_RECT_ROUND_ALFMAX)
            return x.sMax + (x.sFull - x.sMax) * (a / x.aFull - SWB_RECT_ROUND_ALFMAX) / (1.0 - SWB_RECT_ROUND_ALFMAX);
        else if (a > x.aBot)
            return a * pow(xs_r_of_a_direct(x, a, T), 2. / 3.);
        else {
            double aF = SWB_PI * x.rBot * x.rBot;
            double al = a / aF;
            double sF = x.sBot;
            if (al < 0.04) return sF * circ_s_norm(al);
            return sF * xs_lookup(al, T + XT_S_CIRC, XN_S_CIRC);
        }
      default: {
        if (a == 0.0) return 0.0;
        double r = xs_r_of_a_direct(x, a, T);
        if (r < SWB_TINY) return 0.0;
        return a * pow(r, 2. / 3.);
      }
    }
}

// shapes whose R(A) is derived from the section factor (default branch of xsect.c:1137-1141)
SWB_FI bool xs_r_of_a_via_s(int type)
{
    switch (type) {
      case XS_DUMMY: case XS_CIRCULAR: case XS_FORCE_MAIN: case XS_EGGSHAPED: case XS_HORSESHOE:
      case XS_GOTHIC: case XS_CATENARY: case XS_SEMIELLIPTICAL: case XS_BASKETHANDLE:
      case XS_SEMICIRCULAR: return true;
      default: return false;
    }
}
template <int S = -1>
SWB_FI double xs_r_of_a(const Xs &x, double a, const double *T)
{
    if (a <= 0.0) return 0.0;
    const int type = (S >= 0 ? S : x.type);
    if (xs_r_of_a_via_s(type)) {
        if (type == XS_DUMMY) return 0.0;
        double cathy = xs_s_of_a(x, a, T);
        if (cathy < SWB_TINY || a < SWB_TINY) return 0.0;
        return pow(cathy / a, 3. / 2.);
    }
    return xs_r_of_a_direct<S>(x, a, T);
}
// shapes whose R(Y) is R(A(Y)) (default branch of xsect.c:1094)
template <int S = -1>
SWB_FI double xs_r_of_y(const Xs &x, double y, const double *T)
{
    switch (S >= 0 ? S : x.type) {
      case XS_DUMMY: case XS_RECT_CLOSED: case XS_RECT_OPEN: case XS_MOD_BASKET: case XS_GOTHIC:
      case XS_CATENARY: case XS_SEMIELLIPTICAL: case XS_SEMICIRCULAR:
        return xs_r_of_a<S>(x, xs_a_of_y<S>(x, y, T), T);
      default: return xs_r_of_y_direct<S>(x, y, T);
    }
}

// ---- dS/dA (xsect.c:1194-1253, 1424-1470) ------------------------------------------------------
SWB_HD double xs_generic_dsda(const Xs &x, double a, const double *T)
{
    double alpha = a / x.aFull;
    double alpha1 = alpha - 0.001;
    double alpha2 = alpha + 0.001;
    if (alpha1 < 0.0) alpha1 = 0.0;
    double a1 = alpha1 * x.aFull;
    double a2 = alpha2 * x.aFull;
    return (xs_s_of_a(x, a2, T) - xs_s_of_a(x, a1, T)) / (a2 - a1);
}
SWB_HD double xs_tabular_dsda(const Xs &x, double a, const double *tb, int n)
{
    double alpha = a / x.aFull;
    double delta = 1.0 / ((double)n - 1);
    int i = (int)(alpha / delta);
    if (i >= n - 1) i = n - 2;
    double dSdA = (tb[i + 1] - tb[i]) / delta;
    return dSdA * x.sFull / x.aFull;
}
SWB_HD double xs_dsda(const Xs &x, double a, const double *T)
{
    switch (x.type) {
      case XS_FORCE_MAIN:
      case XS_CIRCULAR: {
        double alpha = a / x.aFull;
        if (alpha <= 1.0e-30) return 1.0e-30;
        else if (alpha < 0.04) {
            double theta = circ_theta_of_alpha(alpha);
            double p = theta * x.yFull / 2.0;
            double r = a / p;
            double dPdA = 4.0 / x.yFull / (1. - cos(theta));
            return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
        }
        return xs_tabular_dsda(x, a, T + XT_S_CIRC, XN_S_CIRC);
      }
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_GOTHIC: case XS_CATENARY:
      case XS_SEMIELLIPTICAL: case XS_BASKETHANDLE: case XS_SEMICIRCULAR: {
        XsTabs t = xs_tabs(x.type);
        return xs_tabular_dsda(x, a, T + t.sO, t.sN);
      }
      case XS_RECT_CLOSED: {
        double alpha = a / x.aFull;
        if (alpha > SWB_RECT_ALFMAX) return (x.sFull - x.sMax) / ((1.0 - SWB_RECT_ALFMAX) * x.aFull);
        if (alpha <= 1.0e-30) return xs_generic_dsda(x, a, T);
        double r = xs_r_of_a(x, a, T);
        return (5. / 3. - (2. / 3.) * (2.0 / x.wMax) * r) * pow(r, 2. / 3.);
      }
      case XS_RECT_OPEN: {
        if (a / x.aFull <= 1.0e-30) return xs_generic_dsda(x, a, T);
        double r = xs_r_of_a(x, a, T);
        double dPdA = (2.0 - x.sBot) / x.wMax;
        return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
      }
      case XS_RECT_TRIANG: {
        double alpha = a / x.aFull, dPdA;
        if (alpha > SWB_RECT_TRIANG_ALFMAX)
            return (x.sFull - x.sMax) / ((1.0 - SWB_RECT_TRIANG_ALFMAX) * x.aFull);
        if (alpha <= 1.0e-30) return xs_generic_dsda(x, a, T);
        if (a > x.aBot) dPdA = 2.0 / x.wMax;
        else dPdA = x.rBot / sqrt(a * x.sBot);
        double r = rect_triang_r_of_a(x, a);
        return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
      }
      case XS_RECT_ROUND: {
        if (a / x.aFull > SWB_RECT_ROUND_ALFMAX)
            return (x.sFull - x.sMax) / ((1.0 - SWB_RECT_ROUND_ALFMAX) * x.aFull);
        else if (a > x.aBot) {
            double r = rect_round_r_of_a(x, a, T);
            double dPdA = 2.0 / x.wMax;
            return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
        }
        return xs_generic_dsda(x, a, T);
      }
      case XS_MOD_BASKET: {
        if (a <= x.aFull - x.aBot && a / x.aFull > 1.0e-30) {
            double r = a / (x.wMax + 2.0 * a / x.wMax);
            double dPdA = 2.0 / x.wMax;
            return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
        }
        return xs_generic_dsda(x, a, T);
      }
      case XS_TRAPEZOIDAL: {
        if (a / x.aFull <= 1.0e-30) return xs_generic_dsda(x, a, T);
        double r = a / (x.yBot + trapez_y_of_a(x, a) * x.rBot);
        double dPdA = x.rBot / sqrt(x.yBot * x.yBot + 4. * x.sBot * a);
        return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
      }
      case XS_TRIANGULAR: {
        if (a / x.aFull <= 1.0e-30) return xs_generic_dsda(x, a, T);
        double r = a / (2. * sqrt(a / x.sBot) * x.rBot);
        double dPdA = x.rBot / sqrt(a * x.sBot);
        return (5. / 3. - (2. / 3.) * dPdA * r) * pow(r, 2. / 3.);
      }
      default: return xs_generic_dsda(x, a, T);
    }
}

// ---- A(s): normal-depth side (xsect.c:1146-1190, 1359-1400; findroot.c:20-88) -----------------
SWB_HD double xs_amax(const Xs &x)
{
    if (x.type == XS_IRREGULAR || x.type == XS_CUSTOM) return x.aBot;
    return xs_amax_ratio(x.type) * x.aFull;
}

SWB_HD double xs_generic_a_of_s(const Xs &x, double s, const double *T)
{
    if (s <= 0.0) return 0.0;
    double a1, a2;
    if ((s <= x.sMax && s >= x.sFull) && x.sMax != x.sFull) { a1 = x.aFull; a2 = xs_amax(x); }
    else { a1 = 0.0; a2 = xs_amax(x); }
    double rts = 0.5 * (a1 + a2);
    double xacc = 0.0001 * x.aFull;
    // Newton-Raphson safeguarded by bisection on f(a) = S(a) - s
    double xlo = a1, xhi = a2, dxold = fabs(a2 - a1), dx = dxold, xx = rts, temp;
    double f = xs_s_of_a(x, xx, T) - s;
    double df = xs_dsda(x, xx, T);
    for (int j = 1; j <= 60; j++) {
        if ((((xx - xhi) * df - f) * ((xx - xlo) * df - f) >= 0.0 || (fabs(2.0 * f) > fabs(dxold * df)))) {
            dxold = dx;
            dx = 0.5 * (xhi - xlo);
            xx = xlo + dx;
            if (xlo == xx) break;
        } else {
            dxold = dx;
            dx = f / df;
            temp = xx;
            xx -= dx;
            if (temp == xx) break;
        }
        if (fabs(dx) < xacc) break;
        f = xs_s_of_a(x, xx, T) - s;
        df = xs_dsda(x, xx, T);
        if (f < 0.0) xlo = xx; else xhi = xx;
    }
    return xx;
}

SWB_HD double xs_a_of_s(const Xs &x, double s, const double *T)
{
    double psi = s / x.sFull;
    if (s <= 0.0) return 0.0;
    if (s > x.sMax) s = x.sMax;
    switch (x.type) {
      case XS_DUMMY: return 0.0;
      case XS_FORCE_MAIN:
      case XS_CIRCULAR: {
        double ps = s / x.sFull;                       // circ_getAofS recomputes psi from clamped s
        if (ps == 0.0) return 0.0;
        if (ps >= 1.0) return x.aFull;
        if (ps <= 0.015) return x.aFull * circ_a_norm(ps);
        return x.aFull * xs_inv_lookup(ps, T + XT_S_CIRC, XN_S_CIRC);
      }
      case XS_EGGSHAPED: case XS_HORSESHOE: case XS_GOTHIC: case XS_CATENARY:
      case XS_SEMIELLIPTICAL: case XS_BASKETHANDLE: case XS_SEMICIRCULAR: {
        XsTabs t = xs_tabs(x.type);
        return x.aFull * xs_inv_lookup(psi, T + t.sO, t.sN);    // unclamped psi, as the reference
      }
      default: return xs_generic_a_of_s(x, s, T);
    }
}

// ---- critical depth (xsect.c:1257-1319, 1612-1748; findroot.c:91-138) --------------------------
SWB_HD double xs_qcrit_residual(const Xs &x, double yc, double qTarget, const double *T)
{
    double a = xs_a_of_y(x, yc, T);
    double w = xs_w_of_y(x, yc, T);
    double qc = -qTarget;
    if (w > 0.0) qc = a * sqrt(SWB_GRAVITY * a / w) - qTarget;
    return qc;
}

SWB_HD double xs_ycrit_enum(const Xs &x, double q, double y0, const double *T)
{
    double dy = x.yFull / 25.;
    int i1 = (int)(y0 / dy);
    double q0 = xs_qcrit_residual(x, i1 * dy, 0.0, T), qc, yc;
    if (q0 < q) {
        yc = x.yFull;
        for (int i = i1 + 1; i <= 25; i++) {
            qc = xs_qcrit_residual(x, i * dy, 0.0, T);
            if (qc >= q) { yc = ((q - q0) / (qc - q0) + ((double)i - 1)) * dy; break; }
            q0 = qc;
        }
    } else {
        yc = 0.0;
        for (int i = i1 - 1; i >= 0; i--) {
            qc = xs_qcrit_residual(x, i * dy, 0.0, T);
            if (qc < q) { yc = ((q - qc) / (q0 - qc) + (double)i) * dy; break; }
            q0 = qc;
        }
    }
    return yc;
}

SWB_HD double xs_ycrit_ridder(const Xs &x, double q, double y0, const double *T)
{
    double y1 = 0.0, y2 = 0.99 * x.yFull;
    double q2 = xs_qcrit_residual(x, y2, 0.0, T);
    if (q2 < q) return x.yFull;
    double q0 = xs_qcrit_residual(x, y0, 0.0, T);
    double q1 = xs_qcrit_residual(x, 0.5 * x.yFull, 0.0, T);
    if (q0 > q) { y2 = y0; if (q1 < q) y1 = 0.5 * x.yFull; }
    else        { y1 = y0; if (q1 > q) y2 = 0.5 * x.yFull; }
    // Ridder's method, tolerance 0.001 ft
    const double xacc = 0.001;
    double flo = xs_qcrit_residual(x, y1, q, T);
    double fhi = xs_qcrit_residual(x, y2, q, T);
    if (flo == 0.0) return y1;
    if (fhi == 0.0) return y2;
    double ans = 0.5 * (y1 + y2);
    if ((flo > 0.0 && fhi < 0.0) || (flo < 0.0 && fhi > 0.0)) {
        double xlo = y1, xhi = y2;
        for (int j = 1; j <= 60; j++) {
            double xm = 0.5 * (xlo + xhi);
            double fm = xs_qcrit_residual(x, xm, q, T);
            double s = sqrt(fm * fm - flo * fhi);
            if (s == 0.0) return ans;
            double xnew = xm + (xm - xlo) * ((flo >= fhi ? 1.0 : -1.0) * fm / s);
            if (fabs(xnew - ans) <= xacc) break;
            ans = xnew;
            double fnew = xs_qcrit_residual(x, ans, q, T);
            if (xs_sign(fm, fnew) != fm) { xlo = xm; flo = fm; xhi = ans; fhi = fnew; }
            else if (xs_sign(flo, fnew) != flo) { xhi = ans; fhi = fnew; }
            else if (xs_sign(fhi, fnew) != fhi) { xlo = ans; flo = fnew; }
            else return ans;
            if (fabs(xhi - xlo) <= xacc) return ans;
        }
        return ans;
    }
    return -1.e20;
}

SWB_HD double xs_ycrit(const Xs &x, double q, const double *T)
{
    double q2g = (q * q) / SWB_GRAVITY;
    double y, r;
    if (q2g == 0.0) return 0.0;
    switch (x.type) {
      case XS_DUMMY: return 0.0;
      case XS_RECT_OPEN:
      case XS_RECT_CLOSED: y = pow(q2g / (x.wMax * x.wMax), 1. / 3.); break;
      case XS_TRIANGULAR:  y = pow(2.0 * q2g / (x.sBot * x.sBot), 1. / 5.); break;
      case XS_PARABOLIC:   y = pow(27. / 32. * q2g / (x.rBot * x.rBot), 1. / 4.); break;
      case XS_POWERFUNC:
        y = 1. / (2.0 * x.sBot + 3.0);
        y = pow(q2g * (x.sBot + 1.0) / (x.rBot * x.rBot), y);
        break;
      default:
        y = 1.01 * pow(q2g / x.yFull, 1. / 4.);
        if (y >= x.yFull) y = 0.97 * x.yFull;
        r = x.aFull / (SWB_PI / 4.0 * (x.yFull * x.yFull));
        if (r >= 0.5 && r <= 2.0) y = xs_ycrit_enum(x, q, y, T);
        else y = xs_ycrit_ridder(x, q, y, T);
    }
    return SWB_MIN(y, x.yFull);
}

// ---- run-time-shape entry points: real calls (one copy of each big switch in the binary) ---------
SWB_NI double xs_a_of_y_ni(const Xs &x, double y, const double *T) { return xs_a_of_y(x, y, T); }
SWB_NI double xs_w_of_y_ni(const Xs &x, double y, const double *T) { return xs_w_of_y(x, y, T); }
SWB_NI double xs_r_of_y_ni(const Xs &x, double y, const double *T) { return xs_r_of_y(x, y, T); }
SWB_NI double xs_y_of_a_ni(const Xs &x, double a, const double *T) { return xs_y_of_a(x, a, T); }
SWB_NI double xs_a_of_s_ni(const Xs &x, double s, const double *T) { return xs_a_of_s(x, s, T); }
SWB_NI double xs_ycrit_ni(const Xs &x, double q, const double *T) { return xs_ycrit(x, q, T); }

// S >= 0: shape known at compile time (inlined, switch folded); S < 0: run-time shape (call)
template <int S> SWB_FI double xa(const Xs &x, double y, const double *T)
{ if constexpr (S >= 0) return xs_a_of_y<S>(x, y, T); else return xs_a_of_y_ni(x, y, T); }
template <int S> SWB_FI double xw(const Xs &x, double y, const double *T)
{ if constexpr (S >= 0) return xs_w_of_y<S>(x, y, T); else return xs_w_of_y_ni(x, y, T); }
template <int S> SWB_FI double xr(const Xs &x, double y, const double *T)
{ if constexpr (S >= 0) return xs_r_of_y<S>(x, y, T); else return xs_r_of_y_ni(x, y, T); }

// known-answer dispatcher used by swb_xsect_eval
SWB_HD double xs_eval(int fn, const Xs &x, double arg, const double *T)
{
    switch (fn) {
      case 0: return xs_a_of_y(x, arg, T);
      case 1: return xs_w_of_y(x, arg, T);
      case 2: return xs_r_of_y(x, arg, T);
      case 3: return xs_y_of_a(x, arg, T);
      case 4: return xs_r_of_a(x, arg, T);
      case 5: return xs_s_of_a(x, arg, T);
      case 6: return xs_a_of_s(x, arg, T);
      case 7: return xs_dsda(x, arg, T);
      default: return xs_ycrit(x, arg, T);
    }
}

} // namespace swb
#endif
