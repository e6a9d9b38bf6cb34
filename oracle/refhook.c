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
