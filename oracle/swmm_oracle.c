/*
 * swmm_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement of the reference's dynamic-wave + quality
 * routing hot path, used only as a checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  Nothing in the product path links, loads or calls this file.
 *
 * PARITY PIN: this oracle is pinned against trajectories of the UNMODIFIED reference engine
 * (the .npz fixtures under tests/golden, written by tests/golden/make_golden.py from oracle/_ref/libswmm5.so) and
 * against the live engine itself where oracle/_ref exists (tests/test_oracle.py): simulated time
 * and Picard iteration count after every step and all sampled depths / flows / concentrations
 * must be bit-identical.
 *
 * It is deliberately written the way the reference is -- one model, arrays of per-object structs,
 * sequential loops, scatter-style node sums in link order -- i.e. NOT like the device code
 * (member-major SoA, CSR gathers, fused phases, compaction), so that an agreement between the two
 * is meaningful.
 *
 * Coverage: junctions, outfalls (free / normal / stage), true conduits with CIRCULAR, RECT_CLOSED,
 * RECT_OPEN, TRAPEZOIDAL and TRIANGULAR sections, EXTRAN and SLOT surcharge, variable time step,
 * external inflow hydrographs, completely-mixed quality with first-order decay.  Regulators,
 * storage curves and the other 21 shapes are checked against the engine directly (golden
 * fixtures, seam drop-in tests), not through this file: oracle_create() returns NULL for them.
 *
 * Every function cites the reference lines it follows (SWMM 5.2.4, src/solver/).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "swmm_b200.h"
#include "../stormwater-management-model_b200/csrc/swb_xsect_tables.h"   /* table DATA only */

#define FUDGE 0.0001
#define TINY 1.E-6
#define ZERO 1.E-10
#define PI 3.141592654
#define GRAVITY 32.2
#define OMEGA 0.5
#define MAXVELOCITY 50.
#define MINTIMESTEP 0.001
#define ZeroVolume 0.0353147
#define ZeroDepth 0.003281
#define FLOW_TOL 0.00001
#define MIN(x, y) (((x) <= (y)) ? (x) : (y))
#define MAX(x, y) (((x) >= (y)) ? (x) : (y))
#define SGN(x) (((x) < 0) ? (-1) : (1))
#define MAXP 8

enum { CIRCULAR = 1, RECT_CLOSED = 3, RECT_OPEN = 4, TRAPEZOIDAL = 5, TRIANGULAR = 6 };
enum { DRY, UP_DRY, DN_DRY, SUBCRITICAL, SUPCRITICAL, UP_CRITICAL, DN_CRITICAL };

static const double Tab[] = { SWB_XS_TABLE_DATA };

typedef struct {
    int type;
    double yFull, wMax, ywMax, aFull, rFull, sFull, sMax, yBot, aBot, sBot, rBot;
} XS;

typedef struct {                       /* objects.h:490-530 + dynwave.c:72-79 (TXnode) */
    int type, degree, outfallType, outfallFlap;
    double invert, fullDepth, surDepth, pondedArea, fullVolume, crownElev;
    double inflow, outflow, losses, oldVolume, newVolume, overflow, oldDepth, newDepth;
    double newLatFlow, oldNetInflow, stage;
    double oldQual[MAXP], newQual[MAXP];
    int converged;
    double newSurfArea, oldSurfArea, sumdqdh, dYdT;
} ONode;

typedef struct {                       /* objects.h:664-733 */
    int node1, node2, direction, hasFlap, barrels, hasLosses, flowClass, bypassed;
    double offset1, offset2, qLimit, cLossInlet, cLossOutlet, cLossAvg, seepRate;
    XS xs;
    double length, modLength, roughFactor, slope, beta, qMax;
    double oldFlow, newFlow, oldDepth, newDepth, oldVolume, newVolume, surfArea1, surfArea2;
    double setting, froude, dqdh, a1, a2, q1;
    double oldQual[MAXP], newQual[MAXP], totalLoad[MAXP];
    int normalFlow, fullState, capacityLimited;
} OLink;

typedef struct oracle_model {
    int nN, nL, nP;
    swb_options opt;
    double crownCutoff, kDecay[MAXP];
    ONode *node;
    OLink *link;
    /* inflows */
    int nInf, *infNode, *infStart;
    double *infT, *infQ, *infSf, *infBase, *infC, scale, shift, startDay, startSecs;
    /* clock */
    double simTime, varStep, lastDt;
    int steps;                         /* Picard iterations of the current step */
    long long totIters, nonConv;
} oracle_model;

/* ---- xsect.c:1474-1507 ---------------------------------------------------------------------- */
static double lookup(double x, const double *table, int nItems)
{
    double delta, x0, x1, y, y2;
    int i;
    delta = 1.0 / ((double)nItems - 1);
    i = (int)(x / delta);
    if (i >= nItems - 1) return table[nItems - 1];
    x0 = i * delta;
    x1 = ((double)i + 1) * delta;
    y = table[i] + (x - x0) * (table[i + 1] - table[i]) / delta;
    if (i < 2) {
        y2 = y + (x - x0) * (x - x1) / (delta * delta) * (table[i] / 2.0 - table[i + 1] + table[i + 2] / 2.0);
        if (y2 > 0.0) y = y2;
    }
    if (y < 0.0) y = 0.0;
    return y;
}

static int isOpen(int type) { return type == RECT_OPEN || type == TRAPEZOIDAL || type == TRIANGULAR; }

/* ---- xsect.c:857-1096 for the covered shapes -------------------------------------------------- */
static double getAofY(const XS *x, double y)
{
    double yNorm = y / x->yFull;
    if (y <= 0.0) return 0.0;
    switch (x->type) {
      case CIRCULAR:    return x->aFull * lookup(yNorm, Tab + XT_A_CIRC, XN_A_CIRC);
      case RECT_CLOSED: return y * x->wMax;
      case RECT_OPEN:   return y * x->wMax;
      case TRAPEZOIDAL: return (x->yBot + x->sBot * y) * y;
      case TRIANGULAR:  return y * y * x->sBot;
    }
    return 0.0;
}
static double getWofY(const XS *x, double y)
{
    double yNorm = y / x->yFull;
    switch (x->type) {
      case CIRCULAR:    return x->wMax * lookup(yNorm, Tab + XT_W_CIRC, XN_W_CIRC);
      case RECT_CLOSED: if (yNorm == 1.0) return 0.0; return x->wMax;
      case RECT_OPEN:   return x->wMax;
      case TRAPEZOIDAL: return x->yBot + 2.0 * y * x->sBot;
      case TRIANGULAR:  return 2.0 * x->sBot * y;
    }
    return 0.0;
}
static double rectClosedRofA(const XS *x, double a)                      /* xsect.c:1793-1803 */
{
    double p;
    if (a <= 0.0) return 0.0;
    p = x->wMax + 2. * a / x->wMax;
    if (a / x->aFull > 0.97) p += (a / x->aFull - 0.97) / (1.0 - 0.97) * x->wMax;
    return a / p;
}
static double getRofY(const XS *x, double y)
{
    double yNorm = y / x->yFull, a;
    switch (x->type) {
      case CIRCULAR:    return x->rFull * lookup(yNorm, Tab + XT_R_CIRC, XN_R_CIRC);
      case RECT_CLOSED: return rectClosedRofA(x, getAofY(x, y));
      case RECT_OPEN:   a = getAofY(x, y);                                /* xsect.c:1094,1123 */
                        if (a <= 0.0) return 0.0;
                        return a / (x->wMax + (2. - x->sBot) * a / x->wMax);
      case TRAPEZOIDAL: if (y == 0.0) return 0.0;
                        return ((x->yBot + x->sBot * y) * y) / (x->yBot + y * x->rBot);
      case TRIANGULAR:  return (y * x->sBot) / (2. * x->rBot);
    }
    return 0.0;
}

/* ---- circular specials, xsect.c:2531-2618 ----------------------------------------------------- */
static double SIGNF(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }
static double getThetaOfAlpha(double alpha)
{
    int k;
    double theta, theta1, ap, d;
    if (alpha > 0.04) theta = 1.2 + 5.08 * (alpha - 0.04) / 0.96;
    else theta = 0.031715 - 12.79384 * alpha + 8.28479 * sqrt(alpha);
    theta1 = theta;
    ap = (2.0 * PI) * alpha;
    for (k = 1; k <= 40; k++) {
        d = -(ap - theta + sin(theta)) / (1.0 - cos(theta));
        if (d > 1.0) d = SIGNF(1.0, d);
        theta = theta - d;
        if (fabs(d) <= 0.0001) return theta;
    }
    return theta1;
}
static double getThetaOfPsi(double psi)
{
    int k;
    double theta, theta1, ap, tt, tt23, t3, d;
    if (psi > 0.90) theta = 4.17 + 1.12 * (psi - 0.90) / 0.176;
    else if (psi > 0.5) theta = 3.14 + 1.03 * (psi - 0.5) / 0.4;
    else if (psi > 0.015) theta = 1.2 + 1.94 * (psi - 0.015) / 0.485;
    else theta = 0.12103 - 55.5075 * psi + 15.62254 * sqrt(psi);
    theta1 = theta;
    ap = (2.0 * PI) * psi;
    for (k = 1; k <= 40; k++) {
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
static double getYcircular(double alpha)
{
    double theta;
    if (alpha >= 1.0) return 1.0;
    if (alpha <= 0.0) return 0.0;
    if (alpha <= 1.0e-5) { theta = pow(37.6911 * alpha, 1. / 3.); return theta * theta / 16.0; }
    theta = getThetaOfAlpha(alpha);
    return (1.0 - cos(theta / 2.)) / 2.0;
}
static double getAcircular(double psi)
{
    double theta;
    if (psi >= 1.0) return 1.0;
    if (psi <= 0.0) return 0.0;
    if (psi <= 1.0e-6) { theta = pow(124.4797 * psi, 3. / 13.); return theta * theta * theta / 37.6911; }
    theta = getThetaOfPsi(psi);
    return (theta - sin(theta)) / (2.0 * PI);
}

/* xsect.c:1511-1608 */
static int locate(double y, const double *table, int jLast)
{
    int j, j1 = 0, j2 = jLast;
    if (y <= table[0]) return 0;
    if (y >= table[jLast]) return jLast;
    while (j2 - j1 > 1) {
        j = (j1 + j2) >> 1;
        if (y >= table[j]) j1 = j; else j2 = j;
    }
    return j1;
}
static double invLookup(double y, const double *table, int nItems)
{
    double dx, x, x0, dy;
    int n, i;
    dx = 1.0 / (double)((double)nItems - 1);
    n = nItems;
    if (table[n - 3] > table[n - 1]) n = n - 2;
    if (n < nItems && y > table[nItems - 1]) {
        if (y >= table[nItems - 3]) return ((double)n - 1) * dx;
        if (y <= table[nItems - 2]) i = nItems - 2;
        else i = nItems - 3;
    }
    else i = locate(y, table, n - 1);
    if (i >= n - 1) return ((double)n - 1) * dx;
    x0 = i * dx;
    dy = table[i + 1] - table[i];
    if (dy == 0.0) x = x0;
    else x = x0 + (y - table[i]) * dx / dy;
    if (x < 0.0) x = 0.0;
    if (x > 1.0) x = 1.0;
    return x;
}

/* normal / critical depth are only needed at outfalls and offsets; circular only for normal depth:
 * link_getYnorm (link.c:783-804), circ_getAofS / getYofA (xsect.c:2367-2389), xsect_getYcrit
 * (xsect.c:1257-1319) with getYcritEnum (:1634-1696) */
static double getYnorm(const OLink *L, double q)
{
    const XS *x = &L->xs;
    double s, a, psi, alpha;
    q = fabs(q);
    if (q > L->qMax) q = L->qMax;
    if (q <= 0.0) return 0.0;
    s = q / L->beta;
    if (s > x->sMax) s = x->sMax;
    if (x->type == CIRCULAR) {
        psi = s / x->sFull;
        if (psi == 0.0) a = 0.0;
        else if (psi >= 1.0) a = x->aFull;
        else if (psi <= 0.015) a = x->aFull * getAcircular(psi);
        else a = x->aFull * invLookup(psi, Tab + XT_S_CIRC, XN_S_CIRC);
        alpha = a / x->aFull;
        if (alpha < 0.04) return x->yFull * getYcircular(alpha);
        return x->yFull * lookup(alpha, Tab + XT_Y_CIRC, XN_Y_CIRC);
    }
    return -1.0;     /* generic Newton solve (xsect.c:1359-1400) not restated: caller must avoid */
}
static double getQcritical(const XS *x, double yc)
{
    double a = getAofY(x, yc), w = getWofY(x, yc), qc = -0.0;
    if (w > 0.0) qc = a * sqrt(GRAVITY * a / w) - 0.0;
    return qc;
}
static double getYcrit(const XS *x, double q)
{
    double q2g = (q * q) / GRAVITY, y, r, dy, q0, qc, yc;
    int i1, i;
    if (q2g == 0.0) return 0.0;
    if (x->type == RECT_OPEN || x->type == RECT_CLOSED) y = pow(q2g / (x->wMax * x->wMax), 1. / 3.);
    else if (x->type == TRIANGULAR) y = pow(2.0 * q2g / (x->sBot * x->sBot), 1. / 5.);
    else {
        y = 1.01 * pow(q2g / x->yFull, 1. / 4.);
        if (y >= x->yFull) y = 0.97 * x->yFull;
        r = x->aFull / (PI / 4.0 * (x->yFull * x->yFull));
        if (!(r >= 0.5 && r <= 2.0)) return -1.0;      /* Ridder branch not restated */
        dy = x->yFull / 25.;
        i1 = (int)(y / dy);
        q0 = getQcritical(x, i1 * dy);
        if (q0 < q) {
            yc = x->yFull;
            for (i = i1 + 1; i <= 25; i++) {
                qc = getQcritical(x, i * dy);
                if (qc >= q) { yc = ((q - q0) / (qc - q0) + ((double)i - 1)) * dy; break; }
                q0 = qc;
            }
        } else {
            yc = 0.0;
            for (i = i1 - 1; i >= 0; i--) {
                qc = getQcritical(x, i * dy);
                if (qc < q) { yc = ((q - qc) / (q0 - qc) + (double)i) * dy; break; }
                q0 = qc;
            }
        }
        y = yc;
    }
    return MIN(y, x->yFull);
}

/* ---- dwflow.c:575-633 -------------------------------------------------------------------------- */
static double getSlotWidth(const oracle_model *M, const XS *x, double y)
{
    double yNorm = y / x->yFull;
    if (M->opt.surcharge_method != SWB_SLOT || isOpen(x->type) || yNorm < M->crownCutoff) return 0.0;
    if (yNorm > 1.78) return 0.01 * x->wMax;
    return x->wMax * 0.5423 * exp(-pow(yNorm, 2.4));
}
static double getWidth(const oracle_model *M, const XS *x, double y)
{
    double wSlot = getSlotWidth(M, x, y);
    if (wSlot > 0.0) return wSlot;
    if (y / x->yFull >= M->crownCutoff && !isOpen(x->type)) y = M->crownCutoff * x->yFull;
    return getWofY(x, y);
}
static double getArea(const XS *x, double y, double wSlot)
{
    if (y >= x->yFull) return x->aFull + (y - x->yFull) * wSlot;
    return getAofY(x, y);
}
static double getHydRad(const XS *x, double y)
{
    if (y >= x->yFull) return x->rFull;
    return getRofY(x, y);
}
static double getFroude(const XS *x, double v, double y)                 /* link.c:847-871 */
{
    if (y <= FUDGE) return 0.0;
    if (!isOpen(x->type) && x->yFull - y <= FUDGE) return 0.0;
    y = getAofY(x, y) / getWofY(x, y);
    return fabs(v) / sqrt(GRAVITY * y);
}
static int setFlapGate(const oracle_model *M, const OLink *L, double q)  /* link.c:643-670 */
{
    int n = -1;
    if (L->hasFlap) { if (q * (double)L->direction < 0.0) return 1; }
    if (q < 0.0) n = L->node2;
    if (q > 0.0) n = L->node1;
    if (n >= 0 && M->node[n].type == SWB_OUTFALL && M->node[n].outfallFlap) return 1;
    return 0;
}

/* ---- dwflow.c:297-413 -------------------------------------------------------------------------- */
static int getFlowClass(oracle_model *M, OLink *L, double q, double h1, double h2, double y1, double y2,
                        double *yC, double *yN, double *fasnh)
{
    ONode *N1 = &M->node[L->node1], *N2 = &M->node[L->node2];
    int flowClass = SUBCRITICAL;
    double ycMin, ycMax, z1 = L->offset1, z2 = L->offset2;
    if (N1->type == SWB_OUTFALL) z1 = MAX(0.0, (z1 - N1->newDepth));
    if (N2->type == SWB_OUTFALL) z2 = MAX(0.0, (z2 - N2->newDepth));
    *fasnh = 1.0;
    if (y1 > FUDGE && y2 > FUDGE) {
        if (q < 0.0) {
            if (z1 > 0.0) {
                *yN = getYnorm(L, fabs(q)); *yC = getYcrit(&L->xs, fabs(q));
                ycMin = MIN(*yN, *yC);
                if (y1 < ycMin) flowClass = UP_CRITICAL;
            }
        } else if (z2 > 0.0) {
            *yN = getYnorm(L, fabs(q)); *yC = getYcrit(&L->xs, fabs(q));
            ycMin = MIN(*yN, *yC); ycMax = MAX(*yN, *yC);
            if (y2 < ycMin) flowClass = DN_CRITICAL;
            else if (y2 < ycMax) {
                if (ycMax - ycMin < FUDGE) *fasnh = 0.0;
                else *fasnh = (ycMax - y2) / (ycMax - ycMin);
            }
        }
    }
    else if (y1 <= FUDGE && y2 <= FUDGE) flowClass = DRY;
    else if (y2 > FUDGE) {
        if (h2 < N1->invert + L->offset1) flowClass = UP_DRY;
        else if (z1 > 0.0) { *yN = getYnorm(L, fabs(q)); *yC = getYcrit(&L->xs, fabs(q)); flowClass = UP_CRITICAL; }
    } else {
        if (h1 < N2->invert + L->offset2) flowClass = DN_DRY;
        else if (z2 > 0.0) { *yN = getYnorm(L, fabs(q)); *yC = getYcrit(&L->xs, fabs(q)); flowClass = DN_CRITICAL; }
    }
    return flowClass;
}

/* ---- dwflow.c:417-550 -------------------------------------------------------------------------- */
static void findSurfArea(oracle_model *M, OLink *L, double q, double length, double *h1, double *h2,
                         double *y1, double *y2)
{
    ONode *N1 = &M->node[L->node1], *N2 = &M->node[L->node2];
    const XS *x = &L->xs;
    double fd1 = *y1, fd2 = *y2, fdMid, w1, w2, wMid, sa1 = 0.0, sa2 = 0.0;
    double normalDepth = (fd1 + fd2) / 2.0, criticalDepth = normalDepth, fasnh = 1.0;
    if (fd1 >= x->yFull && fd2 >= x->yFull) L->flowClass = SUBCRITICAL;
    else L->flowClass = getFlowClass(M, L, q, *h1, *h2, *y1, *y2, &criticalDepth, &normalDepth, &fasnh);
    switch (L->flowClass) {
      case SUBCRITICAL:
        fdMid = 0.5 * (fd1 + fd2);
        if (fdMid < FUDGE) fdMid = FUDGE;
        w1 = getWidth(M, x, fd1); w2 = getWidth(M, x, fd2); wMid = getWidth(M, x, fdMid);
        sa1 = (w1 + wMid) * length / 4.;
        sa2 = (wMid + w2) * length / 4. * fasnh;
        break;
      case UP_CRITICAL:
        fd1 = criticalDepth;
        if (normalDepth < criticalDepth) fd1 = normalDepth;
        fd1 = MAX(fd1, FUDGE);
        *h1 = N1->invert + L->offset1 + fd1;
        fdMid = 0.5 * (fd1 + fd2);
        if (fdMid < FUDGE) fdMid = FUDGE;
        w2 = getWidth(M, x, fd2); wMid = getWidth(M, x, fdMid);
        sa2 = (wMid + w2) * length * 0.5;
        break;
      case DN_CRITICAL:
        fd2 = criticalDepth;
        if (normalDepth < criticalDepth) fd2 = normalDepth;
        fd2 = MAX(fd2, FUDGE);
        *h2 = N2->invert + L->offset2 + fd2;
        w1 = getWidth(M, x, fd1);
        fdMid = 0.5 * (fd1 + fd2);
        if (fdMid < FUDGE) fdMid = FUDGE;
        wMid = getWidth(M, x, fdMid);
        sa1 = (w1 + wMid) * length * 0.5;
        break;
      case UP_DRY:
        fd1 = FUDGE;
        fdMid = 0.5 * (fd1 + fd2);
        if (fdMid < FUDGE) fdMid = FUDGE;
        w1 = getWidth(M, x, fd1); w2 = getWidth(M, x, fd2); wMid = getWidth(M, x, fdMid);
        sa2 = (wMid + w2) * length / 4.;
        if (L->offset1 <= 0.0) sa1 = (w1 + wMid) * length / 4.;
        break;
      case DN_DRY:
        fd2 = FUDGE;
        fdMid = 0.5 * (fd1 + fd2);
        if (fdMid < FUDGE) fdMid = FUDGE;
        w1 = getWidth(M, x, fd1); w2 = getWidth(M, x, fd2); wMid = getWidth(M, x, fdMid);
        sa1 = (wMid + w1) * length / 4.;
        if (L->offset2 <= 0.0) sa2 = (w2 + wMid) * length / 4.;
        break;
      case DRY:
        sa1 = FUDGE * length / 2.0;
        sa2 = sa1;
        break;
    }
    L->surfArea1 = sa1; L->surfArea2 = sa2;
    *y1 = fd1; *y2 = fd2;
}

/* ---- dwflow.c:57-293 (no evaporation / seepage / culverts / force mains in the covered set) --- */
static void findConduitFlow(oracle_model *M, OLink *L, int steps, double omega, double dt)
{
    ONode *N1 = &M->node[L->node1], *N2 = &M->node[L->node2];
    const XS *x = &L->xs;
    double z1, z2, h1, h2, y1, y2, a1, a2, r1, yMid, rMid, aMid, aWtd, rWtd, qLast, qOld, aOld, v, rho,
           sigma, length, wSlot, dq1, dq2, dq3, dq4, dq5, denom, q, barrels = L->barrels, losses, f1, qNorm;
    int isFull = 0, isClosed = (L->setting == 0), check, hasOutfall;

    qOld = L->oldFlow / barrels;
    qLast = L->q1;
    z1 = N1->invert + L->offset1;
    z2 = N2->invert + L->offset2;
    h1 = N1->newDepth + N1->invert;
    h2 = N2->newDepth + N2->invert;
    h1 = MAX(h1, z1);
    h2 = MAX(h2, z2);
    y1 = h1 - z1; y2 = h2 - z2;
    y1 = MAX(y1, FUDGE); y2 = MAX(y2, FUDGE);
    if (M->opt.surcharge_method != SWB_SLOT) { y1 = MIN(y1, x->yFull); y2 = MIN(y2, x->yFull); }
    aOld = L->a2;
    aOld = MAX(aOld, FUDGE);
    length = L->modLength;
    findSurfArea(M, L, qLast, length, &h1, &h2, &y1, &y2);
    wSlot = getSlotWidth(M, x, y1);
    a1 = getArea(x, y1, wSlot);
    r1 = getHydRad(x, y1);
    wSlot = getSlotWidth(M, x, y2);
    a2 = getArea(x, y2, wSlot);
    yMid = 0.5 * (y1 + y2);
    wSlot = getSlotWidth(M, x, yMid);
    aMid = getArea(x, yMid, wSlot);
    rMid = getHydRad(x, yMid);
    if (y1 >= x->yFull && y2 >= x->yFull) isFull = 1;
    if (L->flowClass == DRY || L->flowClass == UP_DRY || L->flowClass == DN_DRY || isClosed || aMid <= FUDGE) {
        L->a1 = 0.5 * (a1 + a2);
        L->q1 = 0.0;
        L->dqdh = GRAVITY * dt * aMid / length * barrels;
        L->froude = 0.0;
        L->newDepth = MIN(yMid, x->yFull);
        L->newVolume = L->a1 * L->length * barrels;
        L->newFlow = 0.0;
        return;
    }
    v = qLast / aMid;
    if (fabs(v) > MAXVELOCITY) v = MAXVELOCITY * SGN(qLast);
    L->froude = getFroude(x, v, yMid);
    if (L->flowClass == SUBCRITICAL && L->froude > 1.0) L->flowClass = SUPCRITICAL;
    if (L->froude <= 0.5) sigma = 1.0;
    else if (L->froude >= 1.0) sigma = 0.0;
    else sigma = 2.0 * (1.0 - L->froude);
    rho = 1.0;
    if (!isFull && qLast > 0.0 && h1 >= h2) rho = sigma;
    aWtd = a1 + (aMid - a1) * rho;
    rWtd = r1 + (rMid - r1) * rho;
    if (M->opt.inert_damping == SWB_NO_DAMPING) sigma = 1.0;
    else if (M->opt.inert_damping == SWB_FULL_DAMPING) sigma = 0.0;
    if (isFull && !isOpen(x->type)) sigma = 0.0;
    dq1 = dt * L->roughFactor / pow(rWtd, 1.33333) * fabs(v);
    dq2 = dt * GRAVITY * aWtd * (h2 - h1) / length;
    dq3 = 0.0; dq4 = 0.0;
    if (sigma > 0.0) {
        dq3 = 2.0 * v * (aMid - aOld) * sigma;
        dq4 = dt * v * v * (a2 - a1) / length * sigma;
    }
    dq5 = 0.0;
    if (L->hasLosses) {
        losses = 0.0;
        if (a1 > FUDGE) losses += L->cLossInlet * (fabs(qLast) / a1);
        if (a2 > FUDGE) losses += L->cLossOutlet * (fabs(qLast) / a2);
        if (aMid > FUDGE) losses += L->cLossAvg * (fabs(qLast) / aMid);
        dq5 = losses / 2.0 / length * dt;
    }
    denom = 1.0 + dq1 + dq5;
    q = (qOld - dq2 + dq3 + dq4) / denom;
    L->dqdh = 1.0 / denom * GRAVITY * dt * aWtd / length * barrels;
    L->normalFlow = 0;
    if (q > 0.0) {
        if (M->opt.normal_flow_ltd != SWB_NF_NEITHER && y1 < x->yFull &&
            (L->flowClass == SUBCRITICAL || L->flowClass == SUPCRITICAL)) {
            check = 0;                                                       /* dwflow.c:637-686 */
            hasOutfall = (N1->type == SWB_OUTFALL || N2->type == SWB_OUTFALL);
            if (M->opt.normal_flow_ltd == SWB_NF_SLOPE || M->opt.normal_flow_ltd == SWB_NF_BOTH || hasOutfall)
                if (y1 < y2) check = 1;
            if (!check && (M->opt.normal_flow_ltd == SWB_NF_FROUDE || M->opt.normal_flow_ltd == SWB_NF_BOTH) &&
                !hasOutfall) {
                if (y1 > FUDGE && y2 > FUDGE) {
                    f1 = getFroude(x, q / a1, y1);
                    if (f1 >= 1.0) check = 1;
                }
            }
            if (check) {
                qNorm = L->beta * a1 * pow(r1, 2. / 3.);
                if (qNorm < q) { L->normalFlow = 1; q = qNorm; }
            }
        }
    }
    if (steps > 0) {
        q = (1.0 - omega) * qLast + omega * q;
        if (q * qLast < 0.0) q = 0.001 * SGN(q);
    }
    if (L->qLimit > 0.0) { if (fabs(q) > L->qLimit) q = SGN(q) * L->qLimit; }
    if (setFlapGate(M, L, q)) q = 0.0;
    if (q > FUDGE && N1->newDepth <= FUDGE) q = FUDGE;
    if (q < -FUDGE && N2->newDepth <= FUDGE) q = -FUDGE;
    L->a1 = aMid;
    L->q1 = q;
    L->newDepth = MIN(yMid, x->yFull);
    aMid = (a1 + a2) / 2.0;
    L->fullState = (a1 >= x->aFull) ? ((a2 >= x->aFull) ? SWB_ALL_FULL : SWB_UP_FULL)
                                    : ((a2 >= x->aFull) ? SWB_DN_FULL : 0);
    L->newVolume = aMid * L->length * barrels;
    L->newFlow = q * barrels;
}

/* ---- dynwave.c:528-589 (scatter, in link order) ------------------------------------------------- */
static void updateNodeFlows(oracle_model *M, OLink *L)
{
    ONode *N1 = &M->node[L->node1], *N2 = &M->node[L->node2];
    double q = L->newFlow;
    if (q >= 0.0) { N1->outflow += q; N2->inflow += q; }
    else          { N1->inflow -= q;  N2->outflow -= q; }
    N1->newSurfArea += L->surfArea1 * L->barrels;
    N2->newSurfArea += L->surfArea2 * L->barrels;
    N1->sumdqdh += L->dqdh;
    N2->sumdqdh += L->dqdh;
}

/* ---- link.c:728-766 + node.c:1413-1492 ---------------------------------------------------------- */
static void setOutfallDepth(oracle_model *M, OLink *L)
{
    ONode *N;
    double z, q, yCrit, yNorm, yNew, stage;
    if (M->node[L->node2].type == SWB_OUTFALL) { N = &M->node[L->node2]; z = L->offset2; }
    else if (M->node[L->node1].type == SWB_OUTFALL) { N = &M->node[L->node1]; z = L->offset1; }
    else return;
    q = fabs(L->newFlow / L->barrels);
    yNorm = getYnorm(L, q);
    yCrit = getYcrit(&L->xs, q);
    switch (N->outfallType) {
      case SWB_FREE_OUTFALL:   if (z > 0.0) N->newDepth = 0.0; else N->newDepth = MIN(yNorm, yCrit); return;
      case SWB_NORMAL_OUTFALL: if (z > 0.0) N->newDepth = 0.0; else N->newDepth = yNorm; return;
    }
    stage = N->stage;
    yCrit = MIN(yCrit, yNorm);
    if (yCrit + z + N->invert < stage) yNew = stage - N->invert;
    else if (z > 0.0) {
        if (stage < N->invert + z) yNew = MAX(0.0, (stage - N->invert));
        else yNew = z + yCrit;
    }
    else yNew = yCrit;
    N->newDepth = yNew;
}

/* ---- dynwave.c:636-795 ---------------------------------------------------------------------------- */
static void setNodeDepth(oracle_model *M, ONode *N, double dt)
{
    int canPond, isPonded, isSurcharged = 0;
    double dQ, dV, dy, yMax, yOld, yLast, yNew, yCrown, surfArea, denom, corr, f;
    canPond = (M->opt.allow_ponding && N->pondedArea > 0.0);
    isPonded = (canPond && N->newDepth > N->fullDepth);
    yCrown = N->crownElev - N->invert;
    yOld = N->oldDepth;
    yLast = N->newDepth;
    N->overflow = 0.0;
    surfArea = N->newSurfArea;
    surfArea = MAX(surfArea, M->opt.min_surf_area);
    dQ = N->inflow - N->outflow;
    dV = 0.5 * (N->oldNetInflow + dQ) * dt;
    if (M->opt.surcharge_method == SWB_EXTRAN) {
        if (isPonded) isSurcharged = 0;
        else isSurcharged = (yCrown > 0.0 && yLast > yCrown);
    }
    if (!isSurcharged) {
        dy = dV / surfArea;
        yNew = yOld + dy;
        if (!isPonded) N->oldSurfArea = surfArea;
        if (M->steps > 0) yNew = (1.0 - OMEGA) * yLast + OMEGA * yNew;
        if (isPonded && yNew < N->fullDepth) yNew = N->fullDepth - FUDGE;
    } else {
        corr = 1.0;
        if (N->degree < 0) corr = 0.6;
        denom = N->sumdqdh;
        if (yLast < 1.25 * yCrown) {
            f = (yLast - yCrown) / yCrown;
            denom += (N->oldSurfArea / dt - N->sumdqdh) * exp(-15.0 * f);
        }
        if (denom == 0.0) dy = 0.0;
        else dy = corr * dQ / denom;
        yNew = yLast + dy;
        if (yNew < yCrown) yNew = yCrown - FUDGE;
        if (canPond && yNew > N->fullDepth) yNew = N->fullDepth + FUDGE;
    }
    if (yNew < 0) yNew = 0.0;
    yMax = N->fullDepth;
    if (!canPond) yMax += N->surDepth;
    if (yNew > yMax) {
        if (!canPond) { N->overflow = dV / dt; N->newVolume = N->fullVolume; yNew = yMax; }
        else {
            N->newVolume = MAX((N->oldVolume + dV), N->fullVolume);
            N->overflow = (N->newVolume - MAX(N->oldVolume, N->fullVolume)) / dt;
        }
        if (N->overflow < FUDGE) N->overflow = 0.0;
    }
    else N->newVolume = (N->fullDepth > 0.0) ? N->fullVolume * (yNew / N->fullDepth) : 0.0;
    N->dYdT = fabs(yNew - yOld) / dt;
    N->newDepth = yNew;
}

/* ---- dynwave.c:224-262 ---------------------------------------------------------------------------- */
static int dynwaveExecute(oracle_model *M, double tStep)
{
    int i, j, converged = 0;
    double yOld;
    M->steps = 0;
    for (i = 0; i < M->nN; i++) { M->node[i].converged = 0; M->node[i].dYdT = 0.0; }
    for (j = 0; j < M->nL; j++) { M->link[j].bypassed = 0; M->link[j].a2 = M->link[j].a1; }
    while (M->steps < M->opt.max_trials) {
        for (i = 0; i < M->nN; i++) {                                      /* initNodeStates :297-331 */
            ONode *N = &M->node[i];
            N->newSurfArea = 0.0;
            if (M->opt.allow_ponding && N->newDepth > N->fullDepth && N->pondedArea != 0.0)
                N->newSurfArea = N->pondedArea;
            N->inflow = 0.0;
            N->outflow = N->losses;
            if (N->newLatFlow >= 0.0) N->inflow += N->newLatFlow;
            else N->outflow -= N->newLatFlow;
            N->sumdqdh = 0.0;
        }
        for (j = 0; j < M->nL; j++)                                        /* findLinkFlows :382-412 */
            if (!M->link[j].bypassed) findConduitFlow(M, &M->link[j], M->steps, OMEGA, tStep);
        for (j = 0; j < M->nL; j++) updateNodeFlows(M, &M->link[j]);
        for (j = 0; j < M->nL; j++) setOutfallDepth(M, &M->link[j]);      /* findNodeDepths :593-632 */
        converged = 1;
        for (i = 0; i < M->nN; i++) {
            ONode *N = &M->node[i];
            if (N->type == SWB_OUTFALL) continue;
            yOld = N->newDepth;
            setNodeDepth(M, N, tStep);
            N->converged = 1;
            if (fabs(yOld - N->newDepth) > M->opt.head_tol) { N->converged = 0; converged = 0; }
        }
        M->steps++;
        if (M->steps > 1) {
            if (converged) break;
            for (j = 0; j < M->nL; j++)                                    /* findBypassedLinks :335 */
                M->link[j].bypassed = (M->node[M->link[j].node1].converged &&
                                       M->node[M->link[j].node2].converged);
        }
    }
    if (!converged) M->nonConv++;
    for (j = 0; j < M->nL; j++) {                                          /* findLimitedLinks :349 */
        OLink *L = &M->link[j];
        L->capacityLimited = 0;
        if (L->a1 >= L->xs.aFull) {
            double h1 = M->node[L->node1].newDepth + M->node[L->node1].invert;
            double h2 = M->node[L->node2].newDepth + M->node[L->node2].invert;
            if ((h1 - h2) > fabs(L->slope) * L->length) L->capacityLimited = 1;
        }
    }
    return M->steps;
}

/* ---- dynwave.c:195-220, 799-921 -------------------------------------------------------------------- */
static double getRoutingStep(oracle_model *M, double fixedStep)
{
    int i, j;
    double tMin, t, q, maxDepth;
    if (M->opt.courant_factor == 0.0) return fixedStep;
    if (fixedStep < MINTIMESTEP) return fixedStep;
    if (M->varStep == 0.0) M->varStep = M->opt.min_route_step;
    else {
        tMin = fixedStep;
        for (j = 0; j < M->nL; j++) {
            OLink *L = &M->link[j];
            q = fabs(L->newFlow) / L->barrels;
            if (q <= FUDGE || L->a1 <= FUDGE || L->froude <= 0.01) continue;
            t = L->newVolume / L->barrels / q;
            t = t * L->modLength / L->length;
            t = t * L->froude / (1.0 + L->froude) * M->opt.courant_factor;
            if (t < tMin) tMin = t;
        }
        for (i = 0; i < M->nN; i++) {
            ONode *N = &M->node[i];
            if (N->type == SWB_OUTFALL) continue;
            if (N->newDepth <= FUDGE) continue;
            if (N->newDepth + FUDGE >= N->crownElev - N->invert) continue;
            maxDepth = (N->crownElev - N->invert) * 0.25;
            if (maxDepth < FUDGE) continue;
            if (N->dYdT < FUDGE) continue;
            t = maxDepth / N->dYdT;
            if (t < tMin) tMin = t;
        }
        if (tMin < M->opt.min_route_step) tMin = M->opt.min_route_step;
        M->varStep = tMin;
    }
    M->varStep = floor(1000.0 * M->varStep) / 1000.0;
    return M->varStep;
}

/* ---- qualrout.c:146-174, 498-518 -------------------------------------------------------------------- */
static double getMixedQual(double c, double v1, double wIn, double qIn, double tStep)
{
    double vIn, cIn, cMax;
    if (qIn <= ZERO) return c;
    vIn = qIn * tStep;
    cIn = wIn * tStep / vIn;
    cMax = MAX(c, cIn);
    c = (c * v1 + wIn * tStep) / (v1 + vIn);
    c = MIN(c, cMax);
    c = MAX(c, 0.0);
    return c;
}
static double getReactedQual(double kDecay, double c, double tStep)
{
    double c2;
    if (kDecay == 0.0) return c;
    c2 = c * (1.0 - kDecay * tStep);
    return MAX(0.0, c2);
}

/* ---- qualrout.c:100-142 with :179-353, 398-474 ------------------------------------------------------ */
static void qualroutExecute(oracle_model *M, double tStep)
{
    int i, j, p, up;
    double qLink, qIn, v1, v2, c1, c2, wIn;
    for (j = 0; j < M->nL; j++) {                                         /* findLinkMassFlow */
        OLink *L = &M->link[j];
        ONode *N;
        qLink = L->newFlow;
        N = &M->node[qLink < 0.0 ? L->node1 : L->node2];
        qLink = fabs(qLink);
        for (p = 0; p < M->nP; p++) {
            double w = qLink * L->oldQual[p];
            N->newQual[p] += w;
            L->totalLoad[p] += w * tStep;
        }
    }
    for (i = 0; i < M->nN; i++) {
        ONode *N = &M->node[i];
        qIn = N->inflow;
        if (N->oldVolume > ZeroVolume) {                                    /* findStorageQual */
            v1 = N->oldVolume;
            for (p = 0; p < M->nP; p++) {
                c1 = getReactedQual(M->kDecay[p], N->oldQual[p], tStep);
                c2 = getMixedQual(c1, v1, N->newQual[p], qIn, tStep);
                if ((N->newVolume <= ZeroVolume || N->newDepth <= ZeroDepth) && qIn <= ZERO) c2 = 0.0;
                N->newQual[p] = c2;
            }
        } else if (qIn > ZERO) {                                            /* findNodeQual */
            for (p = 0; p < M->nP; p++) N->newQual[p] /= qIn;
        } else {
            for (p = 0; p < M->nP; p++) N->newQual[p] = (N->newDepth > ZeroDepth) ? N->oldQual[p] : 0.0;
        }
    }
    for (j = 0; j < M->nL; j++) {                                          /* findLinkQual */
        OLink *L = &M->link[j];
        up = (L->newFlow < 0.0) ? L->node2 : L->node1;
        qIn = fabs(L->q1) * L->barrels;
        v1 = L->oldVolume; v2 = L->newVolume;
        qIn = qIn + (v2 + 0.0 - v1) / tStep;
        qIn = MAX(qIn, 0.0);
        for (p = 0; p < M->nP; p++) {
            c1 = L->oldQual[p];
            c2 = getReactedQual(M->kDecay[p], c1, tStep);
            wIn = M->node[up].newQual[p] * qIn;
            c2 = getMixedQual(c2, v1, wIn, qIn, tStep);
            if (v2 < ZeroVolume || L->newDepth <= ZeroDepth) c2 = 0.0;
            L->newQual[p] = c2;
        }
    }
}

/* ---- external inflows: inflow.c:207-234, table.c:745-806 (extend = FALSE) ---------------------------- */
static double tseries(const oracle_model *M, int k, double x)
{
    int i0 = M->infStart[k], i1 = M->infStart[k + 1], i;
    double dx;
    if (i1 <= i0) return 0.0;
    if (x < M->infT[i0] || x > M->infT[i1 - 1]) return 0.0;
    for (i = i0 + 1; i < i1; i++)
        if (x <= M->infT[i]) {
            dx = M->infT[i] - M->infT[i - 1];
            if (fabs(dx) < 1.0e-20) return (M->infQ[i - 1] + M->infQ[i]) / 2.;
            return M->infQ[i - 1] + (x - M->infT[i - 1]) * (M->infQ[i] - M->infQ[i - 1]) / dx;
        }
    return 0.0;
}

/* ---- one routing step: execRouting (swmm5.c:514-575) + routing_execute (routing.c:203-266) ---------- */
int oracle_step(oracle_model *M, double tEnd)
{
    int i, j, k, p, iters;
    double dt, tms, nextms, endms, date, q;
    if (M->simTime >= tEnd) return 0;
    dt = getRoutingStep(M, M->opt.route_step);
    tms = 1000.0 * M->simTime; nextms = tms + 1000.0 * dt; endms = 1000.0 * tEnd;
    if (nextms > endms) { dt = (endms - tms) / 1000.0; dt = MAX(dt, 1. / 1000.0); }
    /* initSystemInflows (routing.c:312-336) */
    for (i = 0; i < M->nN; i++)
        for (p = 0; p < M->nP; p++) { M->node[i].oldQual[p] = M->node[i].newQual[p]; M->node[i].newQual[p] = 0.0; }
    for (j = 0; j < M->nL; j++)
        for (p = 0; p < M->nP; p++) { M->link[j].oldQual[p] = M->link[j].newQual[p]; M->link[j].newQual[p] = 0.0; }
    for (i = 0; i < M->nN; i++) { M->node[i].newLatFlow = 0.0; M->node[i].losses = 0.0; }
    /* addExternalInflows (routing.c:435-490) at getDateTime(NewRoutingTime) (swmm5.c:1543) */
    date = M->startDay + (M->startSecs + (1000.0 * M->simTime + 1.0) / 1000.0) / 86400.0;
    for (k = 0; k < M->nInf; k++) {
        ONode *N = &M->node[M->infNode[k]];
        q = tseries(M, k, date - M->shift) * (M->infSf[k] * M->scale) + M->infBase[k];
        if (fabs(q) < FLOW_TOL) q = 0.0;
        N->newLatFlow += q;
        if (q >= 0.0) for (p = 0; p < M->nP; p++) N->newQual[p] += M->infC[k * M->nP + p] * q;
    }
    /* routeFlow (routing.c:399-409) + flowrout_execute prologue (flowrout.c:153-162) */
    for (j = 0; j < M->nL; j++) {
        OLink *L = &M->link[j];
        L->oldDepth = L->newDepth; L->oldFlow = L->newFlow; L->oldVolume = L->newVolume;
    }
    for (i = 0; i < M->nN; i++) {
        ONode *N = &M->node[i];
        N->oldDepth = N->newDepth; N->oldVolume = N->newVolume;
        N->oldNetInflow = N->inflow - N->outflow;
        N->inflow = N->newLatFlow; N->outflow = N->losses;
        N->overflow = 0.0;
        if (N->newVolume > N->fullVolume) N->overflow = (N->newVolume - N->fullVolume) / dt;
    }
    iters = dynwaveExecute(M, dt);
    if (M->nP > 0 && !M->opt.ignore_quality) qualroutExecute(M, dt);
    M->simTime = (1000.0 * M->simTime + 1000.0 * dt) / 1000.0;
    M->lastDt = dt;
    M->totIters += iters;
    return iters;
}

double oracle_time(const oracle_model *M) { return M->simTime; }
long long oracle_total_iterations(const oracle_model *M) { return M->totIters; }
long long oracle_non_converged(const oracle_model *M) { return M->nonConv; }

/* ---- construction / field access --------------------------------------------------------------------- */
oracle_model *oracle_create(const swb_network_desc *d, const swb_options *o)
{
    int i, j, p;
    oracle_model *M;
    if (d->n_pollut > MAXP) return NULL;
    for (j = 0; j < d->n_links; j++) {
        int t = d->xs_type[j];
        if (d->link_type[j] != SWB_CONDUIT) return NULL;
        if (t != CIRCULAR && t != RECT_CLOSED && t != RECT_OPEN && t != TRAPEZOIDAL && t != TRIANGULAR) return NULL;
        if (d->xs_culvert[j] > 0 || d->link_seep_rate[j] > 0.0) return NULL;
        if (t != CIRCULAR && (d->link_offset1[j] > 0.0 || d->link_offset2[j] > 0.0 ||
            d->node_type[d->link_node1[j]] == SWB_OUTFALL || d->node_type[d->link_node2[j]] == SWB_OUTFALL))
            return NULL;               /* normal depth needs the generic Newton solve */
    }
    for (i = 0; i < d->n_nodes; i++) if (d->node_type[i] == SWB_STORAGE) return NULL;
    M = (oracle_model *)calloc(1, sizeof(*M));
    M->nN = d->n_nodes; M->nL = d->n_links; M->nP = d->n_pollut; M->opt = *o;
    M->crownCutoff = (o->surcharge_method == SWB_SLOT) ? 0.985257 : 0.96;
    M->node = (ONode *)calloc(M->nN, sizeof(ONode));
    M->link = (OLink *)calloc(M->nL, sizeof(OLink));
    M->scale = 1.0;
    for (p = 0; p < M->nP; p++) M->kDecay[p] = d->pollut_kdecay[p];
    for (i = 0; i < M->nN; i++) {
        ONode *N = &M->node[i];
        N->type = d->node_type[i]; N->degree = d->node_degree[i]; N->invert = d->node_invert[i];
        N->fullDepth = d->node_full_depth[i]; N->surDepth = d->node_sur_depth[i];
        N->pondedArea = d->node_ponded_area[i]; N->fullVolume = d->node_full_volume[i];
        N->crownElev = d->node_crown_elev[i]; N->outfallType = d->outfall_type[i];
        N->outfallFlap = d->outfall_flap[i];
    }
    for (j = 0; j < M->nL; j++) {
        OLink *L = &M->link[j];
        XS *x = &L->xs;
        L->node1 = d->link_node1[j]; L->node2 = d->link_node2[j]; L->direction = d->link_direction[j];
        L->hasFlap = d->link_has_flap[j]; L->barrels = d->cond_barrels[j]; L->hasLosses = d->cond_has_losses[j];
        L->offset1 = d->link_offset1[j]; L->offset2 = d->link_offset2[j]; L->qLimit = d->link_q_limit[j];
        L->cLossInlet = d->link_closs_in[j]; L->cLossOutlet = d->link_closs_out[j];
        L->cLossAvg = d->link_closs_avg[j];
        x->type = d->xs_type[j]; x->yFull = d->xs_yfull[j]; x->wMax = d->xs_wmax[j]; x->ywMax = d->xs_ywmax[j];
        x->aFull = d->xs_afull[j]; x->rFull = d->xs_rfull[j]; x->sFull = d->xs_sfull[j]; x->sMax = d->xs_smax[j];
        x->yBot = d->xs_ybot[j]; x->aBot = d->xs_abot[j]; x->sBot = d->xs_sbot[j]; x->rBot = d->xs_rbot[j];
        L->length = d->cond_length[j]; L->modLength = d->cond_mod_length[j];
        L->roughFactor = d->cond_rough_factor[j]; L->slope = d->cond_slope[j]; L->beta = d->cond_beta[j];
        L->qMax = d->cond_q_max[j]; L->setting = 1.0;
    }
    return M;
}

void oracle_destroy(oracle_model *M)
{
    if (!M) return;
    free(M->node); free(M->link); free(M->infNode); free(M->infStart); free(M->infT); free(M->infQ);
    free(M->infSf); free(M->infBase); free(M->infC);
    free(M);
}

void oracle_set_inflows(oracle_model *M, const swb_inflow_desc *f, double scale, double shift_days)
{
    int n = f->n_inflow_nodes, np = f->n_ts_pts, nc = n * (M->nP ? M->nP : 1);
    M->nInf = n; M->scale = scale; M->shift = shift_days;
    M->startDay = f->start_day; M->startSecs = f->start_secs;
    M->infNode = (int *)malloc(sizeof(int) * (n + 1));      memcpy(M->infNode, f->node, sizeof(int) * n);
    M->infStart = (int *)malloc(sizeof(int) * (n + 1));     memcpy(M->infStart, f->ts_start, sizeof(int) * (n + 1));
    M->infT = (double *)malloc(sizeof(double) * (np + 1));  memcpy(M->infT, f->ts_t, sizeof(double) * np);
    M->infQ = (double *)malloc(sizeof(double) * (np + 1));  memcpy(M->infQ, f->ts_q, sizeof(double) * np);
    M->infSf = (double *)malloc(sizeof(double) * (n + 1));  memcpy(M->infSf, f->sfactor, sizeof(double) * n);
    M->infBase = (double *)malloc(sizeof(double) * (n + 1)); memcpy(M->infBase, f->baseline, sizeof(double) * n);
    M->infC = (double *)calloc(nc + 1, sizeof(double));
    if (f->concen && M->nP) memcpy(M->infC, f->concen, sizeof(double) * nc);
}

#define NODE_RW(expr) for (i = 0; i < M->nN; i++) { if (set) expr = buf[i]; else buf[i] = (double)(expr); } return 0
#define LINK_RW(expr) for (i = 0; i < M->nL; i++) { if (set) expr = buf[i]; else buf[i] = (double)(expr); } return 0
static int field_rw(oracle_model *M, int field, double *buf, int set)
{
    int i, p;
    switch (field) {
      case SWB_NODE_NEW_DEPTH:      NODE_RW(M->node[i].newDepth);
      case SWB_NODE_OLD_DEPTH:      NODE_RW(M->node[i].oldDepth);
      case SWB_NODE_NEW_VOLUME:     NODE_RW(M->node[i].newVolume);
      case SWB_NODE_OLD_VOLUME:     NODE_RW(M->node[i].oldVolume);
      case SWB_NODE_NEW_LATFLOW:    NODE_RW(M->node[i].newLatFlow);
      case SWB_NODE_LOSSES:         NODE_RW(M->node[i].losses);
      case SWB_NODE_INFLOW:         NODE_RW(M->node[i].inflow);
      case SWB_NODE_OUTFLOW:        NODE_RW(M->node[i].outflow);
      case SWB_NODE_OVERFLOW:       NODE_RW(M->node[i].overflow);
      case SWB_NODE_OLD_NET_INFLOW: NODE_RW(M->node[i].oldNetInflow);
      case SWB_NODE_OUTFALL_STAGE:  NODE_RW(M->node[i].stage);
      case SWB_LINK_NEW_FLOW:       LINK_RW(M->link[i].newFlow);
      case SWB_LINK_OLD_FLOW:       LINK_RW(M->link[i].oldFlow);
      case SWB_LINK_NEW_DEPTH:      LINK_RW(M->link[i].newDepth);
      case SWB_LINK_OLD_DEPTH:      LINK_RW(M->link[i].oldDepth);
      case SWB_LINK_NEW_VOLUME:     LINK_RW(M->link[i].newVolume);
      case SWB_LINK_OLD_VOLUME:     LINK_RW(M->link[i].oldVolume);
      case SWB_LINK_SETTING:        LINK_RW(M->link[i].setting);
      case SWB_LINK_DQDH:           LINK_RW(M->link[i].dqdh);
      case SWB_LINK_FROUDE:         LINK_RW(M->link[i].froude);
      case SWB_LINK_FLOW_CLASS:     LINK_RW(M->link[i].flowClass);
      case SWB_COND_A1:             LINK_RW(M->link[i].a1);
      case SWB_COND_A2:             LINK_RW(M->link[i].a2);
      case SWB_COND_Q1:             LINK_RW(M->link[i].q1);
      case SWB_NODE_NEW_QUAL: case SWB_NODE_OLD_QUAL:
        for (i = 0; i < M->nN; i++) for (p = 0; p < M->nP; p++) {
            double *v = field == SWB_NODE_NEW_QUAL ? &M->node[i].newQual[p] : &M->node[i].oldQual[p];
            if (set) *v = buf[i * M->nP + p]; else buf[i * M->nP + p] = *v;
        }
        return 0;
      case SWB_LINK_NEW_QUAL: case SWB_LINK_OLD_QUAL: case SWB_LINK_TOTAL_LOAD:
        for (i = 0; i < M->nL; i++) for (p = 0; p < M->nP; p++) {
            double *v = field == SWB_LINK_NEW_QUAL ? &M->link[i].newQual[p] :
                        field == SWB_LINK_OLD_QUAL ? &M->link[i].oldQual[p] : &M->link[i].totalLoad[p];
            if (set) *v = buf[i * M->nP + p]; else buf[i * M->nP + p] = *v;
        }
        return 0;
    }
    return 1;     /* fields the oracle does not carry are simply ignored by the caller */
}
int oracle_set_field(oracle_model *M, int field, const double *buf) { return field_rw(M, field, (double *)buf, 1); }
int oracle_get_field(oracle_model *M, int field, double *buf) { return field_rw(M, field, buf, 0); }
