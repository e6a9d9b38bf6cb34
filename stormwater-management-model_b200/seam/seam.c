/*
 * seam.c -- libswmm5_b200_seam.so: the reference's routing seam, served by the B200 library.
 *
 * Exports exactly the seven functions the reference's upper layers call for dynamic-wave flow
 * routing and quality routing (funcs.h:229-237; callers flowrout.c:87,114,129,167, project.c:261,
 * routing.c:122,248) with the reference's own signatures, so that
 *       LD_PRELOAD=libswmm5_b200_seam.so runswmm model.inp model.rpt model.out
 * or linking the engine without dynwave.c / dwflow.c / qualrout.c runs the unmodified host engine
 * (input parsing, runoff, controls, mass balance, statistics, report, .out) over device routing.
 *
 * Compiled against the reference's headers (never copied): it reads and writes the engine's
 * exported globals Node[], Link[], Conduit[] ... (globals.h:146-169).
 *
 * Data exchange per routing step (correctness-first variant): every dynamic
 * field the device reads is uploaded before the call, every field the host engine reads afterwards
 * is downloaded after it.  Solver-private scratch (the reference's static Xnode[], dynwave.c:72-85)
 * lives only on the device.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "headers.h"
#include "flatten.h"
#include "swmm_b200.h"

static const double MINTIMESTEP      = 0.001;     /* dynwave.c:60-66 */
static const double DEFAULT_SURFAREA = 12.566;
static const double DEFAULT_HEADTOL  = 0.005;
static const int    DEFAULT_MAXTRIALS = 8;
static const double ZeroDepthQ = 0.003281;        /* qualrout.c:41 */

static swb_flat     g_flat;
static swb_network *g_net;
static swb_solver  *g_solver;
static double      *g_buf;        /* staging, max(nodes, links) x max(1, pollutants) */
static double      *g_mb_prev;    /* 3 x pollutants: cumulative mass-balance terms last seen */
static int          g_ready;
static int          g_hyd_current;   /* dynwave_execute ran in this routing step: device hydraulics are current */

static void seam_fail(const char *what)
{
    char msg[512];
    snprintf(msg, sizeof(msg), " B200 routing seam: %s (%s)", what, swb_last_error());
    report_writeErrorMsg(ERR_SYSTEM, msg);
}

/* fields the device reads that the host engine owns between calls */
static const int UP_HYD[] = {
    SWB_NODE_NEW_DEPTH, SWB_NODE_OLD_DEPTH, SWB_NODE_NEW_VOLUME, SWB_NODE_OLD_VOLUME,
    SWB_NODE_NEW_LATFLOW, SWB_NODE_LOSSES, SWB_NODE_INFLOW, SWB_NODE_OUTFLOW, SWB_NODE_OVERFLOW,
    SWB_NODE_OLD_NET_INFLOW, SWB_NODE_OUTFALL_STAGE,
    SWB_LINK_NEW_FLOW, SWB_LINK_OLD_FLOW, SWB_LINK_NEW_DEPTH, SWB_LINK_OLD_DEPTH, SWB_LINK_NEW_VOLUME,
    SWB_LINK_OLD_VOLUME, SWB_LINK_SETTING, SWB_LINK_TARGET_SETTING, SWB_COND_A1, SWB_COND_Q1,
    SWB_ORIF_CORIF, SWB_ORIF_CWEIR, SWB_ORIF_HCRIT, SWB_REG_SURF_AREA, SWB_WEIR_CSURCHARGE };
/* (dqdh, Froude number, flow class, normal-flow / inlet-control flags, full state and the conduit loss rates are
 * pure outputs of a routing step -- every link rewrites them in trial 0 before anything reads them -- so they
 * only travel down; dynwave_init zeroes them on both sides) */
/* fields the host engine reads after dynwave_execute (routing.c, stats.c, massbal.c, output.c) */
static const int DOWN_HYD[] = {
    SWB_NODE_NEW_DEPTH, SWB_NODE_NEW_VOLUME, SWB_NODE_INFLOW, SWB_NODE_OUTFLOW, SWB_NODE_OVERFLOW,
    SWB_LINK_NEW_FLOW, SWB_LINK_NEW_DEPTH, SWB_LINK_NEW_VOLUME, SWB_LINK_SETTING, SWB_LINK_DQDH,
    SWB_LINK_FROUDE, SWB_LINK_FLOW_CLASS, SWB_LINK_SURF_AREA1, SWB_LINK_SURF_AREA2, SWB_LINK_BYPASSED,
    SWB_LINK_NORMAL_FLOW, SWB_LINK_INLET_CONTROL, SWB_COND_A1, SWB_COND_A2, SWB_COND_Q1, SWB_COND_Q2,
    SWB_COND_FULL_STATE, SWB_COND_CAPACITY_LIMITED, SWB_COND_EVAP_LOSS, SWB_COND_SEEP_LOSS,
    SWB_REG_SURF_AREA };
static const int UP_QUAL[] = {
    SWB_NODE_NEW_QUAL, SWB_NODE_OLD_QUAL, SWB_LINK_NEW_QUAL, SWB_LINK_OLD_QUAL, SWB_LINK_TOTAL_LOAD,
    SWB_NODE_STORAGE_EVAP_LOSS, SWB_NODE_STORAGE_EXFIL_LOSS, SWB_NODE_HRT, SWB_NODE_OLD_VOLUME };
static const int DOWN_QUAL[] = {
    SWB_NODE_NEW_QUAL, SWB_LINK_NEW_QUAL, SWB_LINK_TOTAL_LOAD, SWB_NODE_HRT };
#define COUNT(a) ((int)(sizeof(a) / sizeof((a)[0])))

static int push(const int *ids, int n)
{
    int i;
    for (i = 0; i < n; i++) {
        if (swb_engine_get_field(ids[i], g_buf)) continue;
        if (swb_set_field(g_solver, ids[i], 0, 1, g_buf)) return 1;
    }
    return 0;
}
static int pull(const int *ids, int n)
{
    int i;
    for (i = 0; i < n; i++) {
        if (swb_get_field(g_solver, ids[i], 0, 1, g_buf)) return 1;
        swb_engine_set_field(ids[i], g_buf);
    }
    return 0;
}

static int ensure_device(void)
{
    int device = 0, nN = Nobjects[NODE], nL = Nobjects[LINK], nP = Nobjects[POLLUT], n, rc, i;
    const char *env = getenv("SWB_DEVICE");
    if (g_ready) return 1;
    if (env) device = atoi(env);
    for (i = 0; i < nN; i++)
        if (Node[i].treatment) {
            report_writeErrorMsg(ERR_SYSTEM, " B200 routing seam: treatment expressions are not supported");
            return 0;
        }
    rc = swb_flatten_network(&g_flat);
    if (rc) { seam_fail("network uses an element the device path does not cover"); return 0; }
    if (swb_network_create(&g_flat.desc, &g_flat.opt, device, &g_net)) { seam_fail("network upload"); return 0; }
    if (swb_solver_create(g_net, 1, &g_solver)) { seam_fail("solver creation"); return 0; }
    n = (nN > nL ? nN : nL) * (nP > 0 ? nP : 1);
    g_buf = (double *)calloc((size_t)n + 1, sizeof(double));
    g_mb_prev = (double *)calloc((size_t)3 * (nP > 0 ? nP : 1), sizeof(double));
    g_ready = 1;
    return 1;
}

/* ---- funcs.h:229-233 -------------------------------------------------------------------------- */
void dynwave_validate(void)                                   /* dynwave.c:177-191 */
{
    if (MinRouteStep > RouteStep) MinRouteStep = RouteStep;
    if (MinRouteStep < MINTIMESTEP) MinRouteStep = MINTIMESTEP;
    if (MinSurfArea == 0.0) MinSurfArea = DEFAULT_SURFAREA;
    else MinSurfArea /= UCF(LENGTH) * UCF(LENGTH);
    if (HeadTol == 0.0) HeadTol = DEFAULT_HEADTOL;
    else HeadTol /= UCF(LENGTH);
    if (MaxTrials == 0) MaxTrials = DEFAULT_MAXTRIALS;
}

void dynwave_init(void)                                       /* dynwave.c:117-161 */
{
    int i, j;
    double z;
    for (i = 0; i < Nobjects[NODE]; i++) Node[i].crownElev = Node[i].invertElev;
    for (i = 0; i < Nobjects[LINK]; i++) {
        j = Link[i].node1;
        z = Node[j].invertElev + Link[i].offset1 + Link[i].xsect.yFull;
        Node[j].crownElev = MAX(Node[j].crownElev, z);
        j = Link[i].node2;
        z = Node[j].invertElev + Link[i].offset2 + Link[i].xsect.yFull;
        Node[j].crownElev = MAX(Node[j].crownElev, z);
        Link[i].flowClass = DRY;
        Link[i].dqdh = 0.0;
    }
    if (SurchargeMethod == SLOT) CrownCutoff = 0.985257;
    else                         CrownCutoff = 0.96;
    /* the device image is built lazily at the first routing call: initial depths, hot start and
     * link settings are only final after flowrout_init / routing_open have finished */
    g_ready = 0;
}

void dynwave_close(void)                                      /* dynwave.c:165-173 */
{
    if (g_solver) swb_solver_destroy(g_solver);
    if (g_net) swb_network_destroy(g_net);
    g_solver = NULL; g_net = NULL;
    if (g_ready) swb_flat_free(&g_flat);
    free(g_buf); free(g_mb_prev);
    g_buf = NULL; g_mb_prev = NULL; g_ready = 0; g_hyd_current = 0;
}

double dynwave_getRoutingStep(double fixedStep)               /* dynwave.c:195-220 */
{
    double dt = fixedStep;
    if (CourantFactor == 0.0) return fixedStep;
    if (fixedStep < MINTIMESTEP) return fixedStep;
    if (ErrorCode || !ensure_device()) return fixedStep;
    if (swb_get_routing_step(g_solver, fixedStep, &dt)) { seam_fail("get_routing_step"); return fixedStep; }
    {
        swb_member_stats st;
        if (!swb_get_stats(g_solver, 0, 1, &st)) stats_updateCriticalTimeCount(st.crit_node, st.crit_link);
    }
    return dt;
}

int dynwave_execute(double tStep)                             /* dynwave.c:224-262 */
{
    int iters = 0, i, nN = Nobjects[NODE];
    swb_member_stats before, after;
    if (ErrorCode) return 0;
    if (!ensure_device()) return 0;
    if (swb_set_climate(g_solver, Evap.rate, Adjust.hydconFactor)) { seam_fail("climate"); return 0; }
    if (push(UP_HYD, COUNT(UP_HYD))) { seam_fail("state upload"); return 0; }
    swb_get_stats(g_solver, 0, 1, &before);
    if (swb_dynwave_execute(g_solver, &tStep, &iters)) { seam_fail("dynwave_execute"); return 0; }
    if (pull(DOWN_HYD, COUNT(DOWN_HYD))) { seam_fail("state download"); return 0; }
    g_hyd_current = 1;            /* the device's hydraulic image is this step's: qualrout_execute need not re-send it */
    swb_get_stats(g_solver, 0, 1, &after);
    if (after.non_converged > before.non_converged) {         /* updateConvergenceStats, :266-272 */
        NonConvergeCount++;
        if (!swb_get_field(g_solver, SWB_NODE_CONVERGED, 0, 1, g_buf))
            for (i = 0; i < nN; i++) stats_updateConvergenceStats(i, (int)g_buf[i]);
    }
    return iters;
}

/* ---- funcs.h:236-237 -------------------------------------------------------------------------- */
void qualrout_init(void)                                      /* qualrout.c:63-96 */
{
    int i, p, isWet;
    double c;
    for (i = 0; i < Nobjects[NODE]; i++) {
        isWet = (Node[i].newDepth > ZeroDepthQ);
        for (p = 0; p < Nobjects[POLLUT]; p++) {
            c = isWet ? Pollut[p].initConcen : 0.0;
            Node[i].oldQual[p] = c;
            Node[i].newQual[p] = c;
        }
    }
    for (i = 0; i < Nobjects[LINK]; i++) {
        isWet = (Link[i].newDepth > ZeroDepthQ);
        for (p = 0; p < Nobjects[POLLUT]; p++) {
            c = isWet ? Pollut[p].initConcen : 0.0;
            Link[i].oldQual[p] = c;
            Link[i].newQual[p] = c;
        }
    }
}

void qualrout_execute(double tStep)                           /* qualrout.c:100-142 */
{
    int p, nP = Nobjects[POLLUT];
    double mb[3 * 16];
    if (ErrorCode || nP == 0) return;
    if (!ensure_device()) return;
    /* steady-state periods skip dynwave_execute (routing.c:241-243): the hydraulic image the
     * quality step reads (flows, volumes, inflows) must then come from the host as well */
    if ((!g_hyd_current && push(UP_HYD, COUNT(UP_HYD))) || push(UP_QUAL, COUNT(UP_QUAL))) { seam_fail("quality upload"); return; }
    g_hyd_current = 0;
    if (swb_qualrout_execute(g_solver, &tStep)) { seam_fail("qualrout_execute"); return; }
    if (pull(DOWN_QUAL, COUNT(DOWN_QUAL))) { seam_fail("quality download"); return; }
    /* massbal_addReactedMass / addSeepageLoss take rates (mass/s), addToFinalStorage a mass */
    if (nP <= 16 && !swb_get_massbal(g_solver, 0, 1, mb, mb + nP, mb + 2 * nP)) {
        for (p = 0; p < nP; p++) {
            massbal_addReactedMass(p, (mb[p] - g_mb_prev[p]) / tStep);
            massbal_addSeepageLoss(p, (mb[nP + p] - g_mb_prev[nP + p]) / tStep);
            massbal_addToFinalStorage(p, mb[2 * nP + p] - g_mb_prev[2 * nP + p]);
        }
        memcpy(g_mb_prev, mb, sizeof(double) * 3 * nP);
    }
}
