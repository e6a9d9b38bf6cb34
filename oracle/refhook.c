/*
 * refhook.c -- TEST INFRASTRUCTURE (oracle/).  Small accessor library linked against the
 * unmodified reference engine (oracle/_ref/libswmm5.so) and compiled against the reference's own
 * headers.  It lets tests/ and bench.py's cpu_baseline leg
 *   - read the engine's global objects as the flat arrays of include/swmm_b200.h
 *     (via the product's seam/flatten.c, so the extraction code under test is the shipped one),
 *   - evaluate the reference's xsect_* functions on an arbitrary TXsect (known-answer vectors).
 * Nothing in the product path loads this file.
 */
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include "headers.h"
#include "flatten.h"

static swb_flat g_flat;
static int g_have;

const swb_network_desc *refhook_network(void)
{
    if (g_have) swb_flat_free(&g_flat);
    g_have = 1;
    swb_flatten_network(&g_flat);
    return &g_flat.desc;
}
const swb_options *refhook_options(void) { return &g_flat.opt; }
int refhook_get_field(int field, double *buf) { return swb_engine_get_field(field, buf); }
int refhook_set_field(int field, const double *buf) { return swb_engine_set_field(field, buf); }
int refhook_field_len(int field)
{ return swb_field_len(field, Nobjects[NODE], Nobjects[LINK], Nobjects[POLLUT]); }
double refhook_new_routing_time(void) { return NewRoutingTime; }
int refhook_non_converge_count(void) { return NonConvergeCount; }
int refhook_error_code(void) { return ErrorCode; }

/* p = {yFull,wMax,ywMax,aFull,rFull,sFull,sMax,yBot,aBot,sBot,rBot} */
static void mk(TXsect *x, int type, const double *p)
{
    memset(x, 0, sizeof(*x));
    x->type = type; x->transect = -1;
    x->yFull = p[0]; x->wMax = p[1]; x->ywMax = p[2]; x->aFull = p[3]; x->rFull = p[4];
    x->sFull = p[5]; x->sMax = p[6]; x->yBot = p[7]; x->aBot = p[8]; x->sBot = p[9]; x->rBot = p[10];
}
/* fn: 0 AofY 1 WofY 2 RofY 3 YofA 4 RofA 5 SofA 6 AofS 7 dSdA 8 Ycrit */
void refhook_xsect_eval(int fn, int type, const double *p, int n, const double *arg, double *out)
{
    TXsect x; int i;
    mk(&x, type, p);
    for (i = 0; i < n; i++) switch (fn) {
        case 0: out[i] = xsect_getAofY(&x, arg[i]); break;
        case 1: out[i] = xsect_getWofY(&x, arg[i]); break;
        case 2: out[i] = xsect_getRofY(&x, arg[i]); break;
        case 3: out[i] = xsect_getYofA(&x, arg[i]); break;
        case 4: out[i] = xsect_getRofA(&x, arg[i]); break;
        case 5: out[i] = xsect_getSofA(&x, arg[i]); break;
        case 6: out[i] = xsect_getAofS(&x, arg[i]); break;
        case 7: out[i] = xsect_getdSdA(&x, arg[i]); break;
        case 8: out[i] = xsect_getYcrit(&x, arg[i]); break;
    }
}
/* xsect_setParams (xsect.c:216) on user geometry -> the 11 derived parameters */
int refhook_xsect_set(int type, const double *geom4, double ucf, double *p)
{
    TXsect x; double g[4]; int ok;
    memset(&x, 0, sizeof(x));
    memcpy(g, geom4, sizeof(g));
    ok = xsect_setParams(&x, type, g, ucf);
    p[0] = x.yFull; p[1] = x.wMax; p[2] = x.ywMax; p[3] = x.aFull; p[4] = x.rFull; p[5] = x.sFull;
    p[6] = x.sMax; p[7] = x.yBot; p[8] = x.aBot; p[9] = x.sBot; p[10] = x.rBot;
    return ok;
}

/* ---- external inflow hydrographs of the live engine -> swb_inflow_desc (ensemble driver) ----
 * Nodes with a FLOW time-series inflow (no baseline pattern).  Concentration inflows are taken as
 * constants evaluated at the start date (the BASELINE configs use constant series). */
static swb_inflow_desc g_inf;
static int *g_inf_node, *g_inf_start;
static double *g_inf_t, *g_inf_q, *g_inf_sf, *g_inf_bl, *g_inf_c;

const swb_inflow_desc *refhook_inflows(void)
{
    int nN = Nobjects[NODE], nP = Nobjects[POLLUT], i, n = 0, npts = 0, k, p;
    TExtInflow *f;
    for (i = 0; i < nN; i++)
        for (f = Node[i].extInflow; f; f = f->next)
            if (f->type == FLOW_INFLOW && f->tSeries >= 0) {
                TTableEntry *e = Tseries[f->tSeries].firstEntry;
                n++;
                while (e) { npts++; e = e->next; }
            }
    free(g_inf_node); free(g_inf_start); free(g_inf_t); free(g_inf_q); free(g_inf_sf);
    free(g_inf_bl); free(g_inf_c);
    g_inf_node = calloc(n + 1, sizeof(int)); g_inf_start = calloc(n + 2, sizeof(int));
    g_inf_t = calloc(npts + 1, sizeof(double)); g_inf_q = calloc(npts + 1, sizeof(double));
    g_inf_sf = calloc(n + 1, sizeof(double)); g_inf_bl = calloc(n + 1, sizeof(double));
    g_inf_c = calloc((size_t)(n + 1) * (nP ? nP : 1), sizeof(double));
    k = 0; npts = 0;
    for (i = 0; i < nN; i++) {
        TExtInflow *flow = NULL;
        for (f = Node[i].extInflow; f; f = f->next)
            if (f->type == FLOW_INFLOW && f->tSeries >= 0) { flow = f; break; }
        if (!flow) continue;
        g_inf_node[k] = i; g_inf_start[k] = npts;
        g_inf_sf[k] = flow->sFactor * flow->cFactor;      /* cFactor = 1 / UCF(FLOW) */
        g_inf_bl[k] = flow->baseline * flow->cFactor;
        { TTableEntry *e = Tseries[flow->tSeries].firstEntry;
          while (e) { g_inf_t[npts] = e->x; g_inf_q[npts] = e->y; npts++; e = e->next; } }
        for (f = Node[i].extInflow; f; f = f->next)
            if (f->type == CONCEN_INFLOW) {
                p = f->param;
                g_inf_c[k * nP + p] = inflow_getExtInflow(f, StartDateTime);
            }
        k++;
    }
    g_inf_start[k] = npts;
    memset(&g_inf, 0, sizeof(g_inf));
    g_inf.n_inflow_nodes = n; g_inf.n_ts_pts = npts; g_inf.node = g_inf_node;
    g_inf.ts_start = g_inf_start; g_inf.ts_t = g_inf_t; g_inf.ts_q = g_inf_q;
    g_inf.sfactor = g_inf_sf; g_inf.baseline = g_inf_bl; g_inf.concen = g_inf_c;
    {
        int h, mi, s;
        datetime_decodeTime(StartDateTime, &h, &mi, &s);
        g_inf.start_day = floor(StartDateTime);
        g_inf.start_secs = 3600.0 * h + 60.0 * mi + s;
    }
    return &g_inf;
}
double refhook_total_duration(void) { return TotalDuration; }

/* ---- the shipped flatteners (seam/flatten.c) for everything swb_run_steps evaluates on the device:
 * external / dry-weather inflows with patterns and pollutant records, control rules, pump depths,
 * timed outfall stages */
static swb_flat g_flat_inf, g_flat_ctl;
static swb_inflow_desc g_inf_full;
static swb_controls_desc g_ctl;
const swb_inflow_desc *refhook_inflows_full(void)
{
    swb_flat_free(&g_flat_inf);
    if (swb_flatten_inflows(&g_flat_inf, &g_inf_full) != 0) return NULL;
    return &g_inf_full;
}
const swb_controls_desc *refhook_controls(void)
{
    swb_flat_free(&g_flat_ctl);
    if (swb_flatten_controls(&g_flat_ctl, &g_ctl) != 0) return NULL;
    return &g_ctl;
}

/* ---- statistics of the live engine in the plane order of include/swmm_b200.h (swb_node_stat /
 * swb_link_stat): NodeStats / StorageStats / OutfallStats / LinkStats / PumpStats (stats.c:63-68).
 * Dates become elapsed seconds since StartDateTime. */
extern TNodeStats *NodeStats;
extern TLinkStats *LinkStats;
extern TStorageStats *StorageStats;
extern TOutfallStats *OutfallStats;
extern TPumpStats *PumpStats;
extern double MaxOutfallFlow;
static double secs(DateTime d) { return (d - StartDateTime) * 86400.0; }

void refhook_node_stats(double *out)
{
    int nN = Nobjects[NODE], nP = Nobjects[POLLUT], i, p, k;
#define NS(plane) out[(size_t)(plane) * nN + i]
    for (i = 0; i < nN; i++) {
        TNodeStats *s = &NodeStats[i];
        NS(SWB_NS_SUM_DEPTH) = s->avgDepth; NS(SWB_NS_MAX_DEPTH) = s->maxDepth;
        NS(SWB_NS_MAX_DEPTH_TIME) = secs(s->maxDepthDate); NS(SWB_NS_TIME_FLOODED) = s->timeFlooded;
        NS(SWB_NS_VOL_FLOODED) = s->volFlooded; NS(SWB_NS_MAX_PONDED_VOL) = s->maxPondedVol;
        NS(SWB_NS_TIME_SURCHARGED) = s->timeSurcharged; NS(SWB_NS_TOT_LATFLOW) = s->totLatFlow;
        NS(SWB_NS_MAX_LATFLOW) = s->maxLatFlow; NS(SWB_NS_MAX_INFLOW) = s->maxInflow;
        NS(SWB_NS_MAX_INFLOW_TIME) = secs(s->maxInflowDate); NS(SWB_NS_MAX_OVERFLOW) = s->maxOverflow;
        NS(SWB_NS_MAX_OVERFLOW_TIME) = secs(s->maxOverflowDate);
        NS(SWB_NS_NONCONV_COUNT) = s->nonConvergedCount; NS(SWB_NS_TIME_COURANT) = s->timeCourantCritical;
        NS(SWB_NS_X_SUM) = NS(SWB_NS_X_MAX) = NS(SWB_NS_X_MAX_TIME) = NS(SWB_NS_X_MAX_FLOW) = 0.0;
        NS(SWB_NS_X_EVAP) = NS(SWB_NS_X_EXFIL) = 0.0;
        for (p = 0; p < nP; p++) NS(SWB_NS_LOAD0 + p) = 0.0;
        k = Node[i].subIndex;
        if (Node[i].type == STORAGE) {
            NS(SWB_NS_X_SUM) = StorageStats[k].avgVol; NS(SWB_NS_X_MAX) = StorageStats[k].maxVol;
            NS(SWB_NS_X_MAX_TIME) = secs(StorageStats[k].maxVolDate); NS(SWB_NS_X_MAX_FLOW) = StorageStats[k].maxFlow;
            NS(SWB_NS_X_EVAP) = StorageStats[k].evapLosses; NS(SWB_NS_X_EXFIL) = StorageStats[k].exfilLosses;
        } else if (Node[i].type == OUTFALL) {
            NS(SWB_NS_X_SUM) = OutfallStats[k].avgFlow; NS(SWB_NS_X_MAX) = OutfallStats[k].maxFlow;
            NS(SWB_NS_X_MAX_TIME) = OutfallStats[k].totalPeriods;
            for (p = 0; p < nP; p++) NS(SWB_NS_LOAD0 + p) = OutfallStats[k].totalLoad[p];
        }
    }
#undef NS
}

void refhook_link_stats(double *out)
{
    int nL = Nobjects[LINK], j, c, k;
#define LS(plane) out[(size_t)(plane) * nL + j]
    for (j = 0; j < nL; j++) {
        TLinkStats *s = &LinkStats[j];
        for (c = 0; c < SWB_LS_PLANES; c++) LS(c) = 0.0;
        LS(SWB_LS_MAX_FLOW) = s->maxFlow; LS(SWB_LS_MAX_FLOW_TIME) = secs(s->maxFlowDate);
        LS(SWB_LS_MAX_VELOC) = s->maxVeloc; LS(SWB_LS_MAX_DEPTH) = s->maxDepth;
        LS(SWB_LS_TIME_FULL_FLOW) = s->timeFullFlow; LS(SWB_LS_TIME_CAP_LIMITED) = s->timeCapacityLimited;
        LS(SWB_LS_TIME_SURCHARGED) = s->timeSurcharged; LS(SWB_LS_TIME_FULL_UP) = s->timeFullUpstream;
        LS(SWB_LS_TIME_FULL_DN) = s->timeFullDnstream; LS(SWB_LS_TURN_SIGN) = s->flowTurnSign;
        LS(SWB_LS_TURNS) = s->flowTurns; LS(SWB_LS_TIME_COURANT) = s->timeCourantCritical;
        if (Link[j].type == PUMP) {
            k = Link[j].subIndex;
            LS(SWB_LS_PUMP_MIN_FLOW) = PumpStats[k].minFlow; LS(SWB_LS_PUMP_SUM_FLOW) = PumpStats[k].avgFlow;
            LS(SWB_LS_PUMP_VOLUME) = PumpStats[k].volume; LS(SWB_LS_PUMP_UTILIZED) = PumpStats[k].utilized;
            LS(SWB_LS_PUMP_ENERGY) = PumpStats[k].energy; LS(SWB_LS_PUMP_OFF_LOW) = PumpStats[k].offCurveLow;
            LS(SWB_LS_PUMP_OFF_HIGH) = PumpStats[k].offCurveHigh; LS(SWB_LS_PUMP_STARTUPS) = PumpStats[k].startUps;
            LS(SWB_LS_PUMP_PERIODS) = PumpStats[k].totalPeriods;
        } else {
            LS(SWB_LS_TIME_NORMAL) = s->timeNormalFlow; LS(SWB_LS_TIME_INLET) = s->timeInletControl;
            for (c = 0; c < MAX_FLOW_CLASSES; c++) LS(SWB_LS_TIME_CLASS0 + c) = s->timeInFlowClass[c];
        }
    }
#undef LS
}
double refhook_max_outfall_flow(void) { return MaxOutfallFlow; }
