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
