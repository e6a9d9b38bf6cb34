/*
 * refcount.c -- TEST / BENCH INFRASTRUCTURE (oracle/).  Interposer that counts and times the
 * reference's seam functions WITHOUT modifying the reference: it defines
 * the exported seam symbols, forwards to the real ones with dlsym(RTLD_NEXT) and accumulates the
 * return values / clock deltas.  Use with LD_PRELOAD for runswmm, or dlopen it RTLD_GLOBAL before
 * libswmm5.so from Python.  Nominal conduit-updates = sum of dynwave_execute return values x true
 * conduits.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int    (*real_exec)(double);
static double (*real_step)(double);
static void   (*real_qual)(double);
static void   (*real_route)(int, double);

static void   (*real_crit)(int, int);
static int g_crit_node = -1, g_crit_link = -1;
static long long g_steps, g_iters;
static int g_last_iters;
static double g_t_exec, g_t_step, g_t_qual, g_t_route;

int dynwave_execute(double tStep)
{
    if (!real_exec) real_exec = (int (*)(double))dlsym(RTLD_NEXT, "dynwave_execute");
    double t0 = now();
    int n = real_exec(tStep);
    g_t_exec += now() - t0;
    g_steps++; g_iters += n; g_last_iters = n;
    return n;
}
double dynwave_getRoutingStep(double fixedStep)
{
    if (!real_step) real_step = (double (*)(double))dlsym(RTLD_NEXT, "dynwave_getRoutingStep");
    double t0 = now();
    double r = real_step(fixedStep);
    g_t_step += now() - t0;
    return r;
}
void qualrout_execute(double tStep)
{
    if (!real_qual) real_qual = (void (*)(double))dlsym(RTLD_NEXT, "qualrout_execute");
    double t0 = now();
    real_qual(tStep);
    g_t_qual += now() - t0;
}
void routing_execute(int model, double tStep)
{
    if (!real_route) real_route = (void (*)(int, double))dlsym(RTLD_NEXT, "routing_execute");
    double t0 = now();
    real_route(model, tStep);
    g_t_route += now() - t0;
}

/* stats.c:522-532: records which element limited the variable time step */
void stats_updateCriticalTimeCount(int node, int link)
{
    if (!real_crit) real_crit = (void (*)(int, int))dlsym(RTLD_NEXT, "stats_updateCriticalTimeCount");
    g_crit_node = node; g_crit_link = link;
    if (real_crit) real_crit(node, link);
}
void refcount_bind_crit(void *f) { real_crit = (void (*)(int, int))f; }
void refcount_last_critical(int *node, int *link) { *node = g_crit_node; *link = g_crit_link; }

/* dlopen use: hand over the engine's own entry points explicitly (RTLD_NEXT only works for
 * LD_PRELOAD / link-order interposition) */
void refcount_bind(void *exec, void *step, void *qual, void *route)
{
    real_exec = (int (*)(double))exec; real_step = (double (*)(double))step;
    real_qual = (void (*)(double))qual; real_route = (void (*)(int, double))route;
}
int refcount_last_iterations(void) { return g_last_iters; }
void refcount_get(long long *steps, long long *iters, double *t4)
{
    *steps = g_steps; *iters = g_iters;
    t4[0] = g_t_exec; t4[1] = g_t_step; t4[2] = g_t_qual; t4[3] = g_t_route;
}
void refcount_reset(void)
{
    g_steps = g_iters = 0; g_last_iters = 0;
    g_t_exec = g_t_step = g_t_qual = g_t_route = 0.0;
}

/* with LD_PRELOAD: print the totals when the process ends (REFCOUNT_FILE=path to redirect) */
__attribute__((destructor)) static void report(void)
{
    const char *path = getenv("REFCOUNT_FILE");
    FILE *f;
    if (!g_steps || !path) return;
    f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "{\"steps\": %lld, \"iterations\": %lld, \"t_dynwave_execute\": %.6f, "
               "\"t_get_routing_step\": %.6f, \"t_qualrout_execute\": %.6f, \"t_routing_execute\": %.6f}\n",
            g_steps, g_iters, g_t_exec, g_t_step, g_t_qual, g_t_route);
    fclose(f);
}
