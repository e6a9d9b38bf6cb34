/*
 * swmm_b200.h -- C-ABI of libswmm_b200.so, the B200 (sm_100a) dynamic-wave flow routing and
 * water-quality transport library.
 *
 * This is the "core" half of the drop-in boundary.  The other half, include/swmm_b200_seam.h,
 * is the reference's own seam (dynwave_* / qualrout_*, funcs.h:229-237); the seam shim flattens
 * the reference engine's global Node[]/Link[] objects into the swb_network_desc below and drives
 * this API once per routing step.  Ensembles (many scenarios of one network, BASELINE.json
 * configs[3]) have no counterpart in the reference's one-project-per-process API
 * (swmm5.h:129-151) and are driven through this header directly.
 *
 * Conventions
 *   - plain pointers and sizes only; every array is caller-owned host memory unless stated;
 *   - all hydraulic quantities are in the reference's internal units (ft, cfs, s; consts.h:31-50),
 *     curve tables stay in user units exactly as the reference keeps them (link.c:1579-1620);
 *   - every function returns 0 on success or one of SWB_ERR_*; swb_last_error() gives the text;
 *   - host-side layout of a per-member field is [member][item] (one member is one contiguous
 *     array, like the reference's Node[]/Link[]); the device layout is [item][member].
 *   - integer codes (node/link/xsect/flow-class types) are the reference's enums.h values.
 */
#ifndef SWMM_B200_H
#define SWMM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_VERSION 100

/* error codes */
#define SWB_OK            0
#define SWB_ERR_ARG       1   /* bad argument / inconsistent descriptor                    */
#define SWB_ERR_CUDA      2   /* CUDA runtime failure (no device, launch failure, ...)     */
#define SWB_ERR_UNSUPP    3   /* network uses an element the device path does not cover    */
#define SWB_ERR_NAN       4   /* device state became non-finite (maps to ERR_SYSTEM, 500)  */

/* enums.h values used in the descriptor (restated so callers need no reference header) */
enum { SWB_JUNCTION = 0, SWB_OUTFALL = 1, SWB_STORAGE = 2, SWB_DIVIDER = 3 };          /* enums.h:70  */
enum { SWB_CONDUIT = 0, SWB_PUMP = 1, SWB_ORIFICE = 2, SWB_WEIR = 3, SWB_OUTLET = 4 }; /* enums.h:80  */
enum { SWB_TRANSVERSE_WEIR = 0, SWB_SIDEFLOW_WEIR, SWB_VNOTCH_WEIR, SWB_TRAPEZOIDAL_WEIR,
       SWB_ROADWAY_WEIR };                                                               /* enums.h:430 */
enum { SWB_EXTRAN = 0, SWB_SLOT = 1 };                                                  /* enums.h:369 */
enum { SWB_NO_DAMPING = 0, SWB_PARTIAL_DAMPING = 1, SWB_FULL_DAMPING = 2 };             /* enums.h:364 */
enum { SWB_NF_SLOPE = 0, SWB_NF_FROUDE = 1, SWB_NF_BOTH = 2, SWB_NF_NEITHER = 3 };      /* enums.h:358 */
enum { SWB_FREE_OUTFALL = 0, SWB_NORMAL_OUTFALL = 1, SWB_STAGE_OUTFALL = 2 };
       /* FIXED/TIDAL/TIMESERIES outfalls (enums.h:389) all reduce to "stage given per step"   */
enum { SWB_DRY = 0, SWB_UP_DRY, SWB_DN_DRY, SWB_SUBCRITICAL, SWB_SUPCRITICAL,
       SWB_UP_CRITICAL, SWB_DN_CRITICAL, SWB_MAX_FLOW_CLASSES, SWB_UP_FULL, SWB_DN_FULL,
       SWB_ALL_FULL };                                                                   /* enums.h:245 */

/* Options block: [OPTIONS] values the path reads (dynwave.c:177-191, project.c:840-900) */
typedef struct swb_options {
    int    surcharge_method;   /* SurchargeMethod                                           */
    int    inert_damping;      /* InertDamping                                              */
    int    normal_flow_ltd;    /* NormalFlowLtd                                             */
    int    allow_ponding;      /* AllowPonding                                              */
    int    max_trials;         /* MaxTrials (after dynwave_validate: 0 -> 8)                */
    int    force_main_eqn;     /* ForceMainEqn                                              */
    int    unit_system;        /* UnitSystem (0 US, 1 SI) -- regulators evaluate in user units */
    int    ignore_quality;     /* IgnoreQuality                                             */
    double head_tol;           /* HeadTol, ft (after validate)                              */
    double min_surf_area;      /* MinSurfArea, ft2 (after validate)                         */
    double courant_factor;     /* CourantFactor (0 = fixed step)                            */
    double min_route_step;     /* MinRouteStep, s                                           */
    double route_step;         /* RouteStep, s (the user's fixed / maximum step)            */
    double ucf_length;         /* UCF(LENGTH), UCF(VOLUME), UCF(FLOW) (swmm5.c:1378-1388)   */
    double ucf_volume;
    double ucf_flow;
    double evap_rate;          /* Evap.rate, ft/s (per-step scalar, see swb_set_climate)    */
    double hydcon_factor;      /* Adjust.hydconFactor                                       */
} swb_options;

/* Flattened network (static data).  Per-link arrays are indexed by LINK index for every link
 * type; entries that do not apply to a link's type are ignored. */
typedef struct swb_network_desc {
    int n_nodes, n_links, n_pollut, n_curves, n_curve_pts, n_shape_tbls, shape_tbl_len;
    int reserved0;

    /* nodes (objects.h:490-530) */
    const int    *node_type;        /* Node.type                                           */
    const int    *node_degree;      /* Node.degree (sign set by flowrout.c:330)            */
    const double *node_invert;      /* invertElev                                          */
    const double *node_full_depth;  /* fullDepth                                           */
    const double *node_sur_depth;   /* surDepth                                            */
    const double *node_ponded_area; /* pondedArea                                          */
    const double *node_full_volume; /* fullVolume                                          */
    const double *node_crown_elev;  /* crownElev as set by dynwave_init (dynwave.c:137-153)*/
    /* outfall / storage attributes, indexed by NODE (ignored for other node types) */
    const int    *outfall_type;     /* SWB_*_OUTFALL                                       */
    const int    *outfall_flap;     /* Outfall.hasFlapGate                                 */
    const int    *storage_shape;    /* Storage.shape (enums.h:396)                         */
    const int    *storage_curve;    /* Storage.aCurve (curve index or -1)                  */
    const double *storage_a0, *storage_a1, *storage_a2;

    /* links (objects.h:664-709) */
    const int    *link_type, *link_node1, *link_node2, *link_direction, *link_has_flap;
    const double *link_offset1, *link_offset2, *link_q_limit, *link_q_full;
    const double *link_closs_in, *link_closs_out, *link_closs_avg, *link_seep_rate;
    /* cross sections (objects.h:581-599) */
    const int    *xs_type, *xs_culvert, *xs_table;  /* xs_table: index into shape tables or -1 */
    const double *xs_yfull, *xs_wmax, *xs_ywmax, *xs_afull, *xs_rfull, *xs_sfull, *xs_smax;
    const double *xs_ybot, *xs_abot, *xs_sbot, *xs_rbot;
    /* conduits (objects.h:714-733) */
    const int    *cond_barrels, *cond_has_losses;
    const double *cond_length;      /* TRUE length, link_getLength (link.c:808,1195)       */
    const double *cond_user_length; /* Conduit.length as entered (dynwave.c:374)           */
    const double *cond_mod_length, *cond_rough_factor, *cond_slope, *cond_beta, *cond_q_max;
    /* regulators, indexed by LINK */
    const int    *pump_type, *pump_curve;           /* Pump.type (enums.h:418), pumpCurve  */
    const double *pump_xmin, *pump_xmax;
    const int    *orif_type;                        /* Orifice.type                        */
    const double *orif_cdisch, *orif_length;
    const int    *weir_type, *weir_can_surcharge, *weir_cd_curve;
    const double *weir_cdisch1, *weir_cdisch2, *weir_end_con, *weir_slope, *weir_length;
    const double *weir_road_width;                  /* Weir.roadWidth, ft (ROADWAY weirs)  */
    const int    *weir_road_surface;                /* Weir.roadSurface: 1 paved, 2 gravel */
    const int    *outlet_curve, *outlet_curve_type; /* Outlet.qCurve, curveType            */
    const double *outlet_qcoeff, *outlet_qexpon;

    /* curves flattened from the reference's linked lists (table.c), CSR by curve */
    const int    *curve_start;      /* n_curves+1                                          */
    const int    *curve_type;       /* Curve.curveType (enums.h:437)                       */
    const double *curve_x, *curve_y;
    /* per-object geometry tables for IRREGULAR / CUSTOM / STREET (objects.h:604-659):
     * n_shape_tbls tables of shape_tbl_len entries each, [table][entry]; table t has
     * shape_tbl_n[t] valid entries (51 except for STREET transects) */
    const int    *shape_tbl_n;
    const double *shape_area_tbl, *shape_hrad_tbl, *shape_width_tbl;

    /* pollutants */
    const double *pollut_kdecay;    /* Pollut.kDecay, 1/s                                  */
} swb_network_desc;

/* Fields that can be read or written with swb_get_field / swb_set_field.
 * NODE_* are n_nodes long per member, LINK_* n_links, *_QUAL and LINK_TOTAL_LOAD are
 * [item][pollutant] per member on the host side.  Integer fields travel as doubles. */
enum swb_field {
    SWB_NODE_NEW_DEPTH = 0, SWB_NODE_OLD_DEPTH, SWB_NODE_NEW_VOLUME, SWB_NODE_OLD_VOLUME,
    SWB_NODE_NEW_LATFLOW, SWB_NODE_LOSSES, SWB_NODE_INFLOW, SWB_NODE_OUTFLOW, SWB_NODE_OVERFLOW,
    SWB_NODE_OLD_NET_INFLOW, SWB_NODE_NEW_SURF_AREA, SWB_NODE_OLD_SURF_AREA, SWB_NODE_SUMDQDH,
    SWB_NODE_DYDT, SWB_NODE_CONVERGED, SWB_NODE_OUTFALL_STAGE, SWB_NODE_STORAGE_EVAP_LOSS,
    SWB_NODE_STORAGE_EXFIL_LOSS, SWB_NODE_HRT, SWB_NODE_NEW_QUAL, SWB_NODE_OLD_QUAL,
    SWB_NODE_OLD_LATFLOW, SWB_NODE_OLD_INFLOW,   /* Node.oldLatFlow / oldFlowInflow (report interpolation) */
    SWB_LINK_NEW_FLOW = 32, SWB_LINK_OLD_FLOW, SWB_LINK_NEW_DEPTH, SWB_LINK_OLD_DEPTH,
    SWB_LINK_NEW_VOLUME, SWB_LINK_OLD_VOLUME, SWB_LINK_SETTING, SWB_LINK_TARGET_SETTING,
    SWB_LINK_DQDH, SWB_LINK_FROUDE, SWB_LINK_FLOW_CLASS, SWB_LINK_SURF_AREA1, SWB_LINK_SURF_AREA2,
    SWB_LINK_BYPASSED, SWB_LINK_NORMAL_FLOW, SWB_LINK_INLET_CONTROL, SWB_COND_A1, SWB_COND_A2,
    SWB_COND_Q1, SWB_COND_Q2, SWB_COND_FULL_STATE, SWB_COND_CAPACITY_LIMITED, SWB_COND_EVAP_LOSS,
    SWB_COND_SEEP_LOSS, SWB_ORIF_CORIF, SWB_ORIF_CWEIR, SWB_ORIF_HCRIT, SWB_REG_SURF_AREA,
    SWB_WEIR_CSURCHARGE, SWB_LINK_NEW_QUAL, SWB_LINK_OLD_QUAL, SWB_LINK_TOTAL_LOAD,
    SWB_FIELD_COUNT = 96
};

/*
 * Per-object routing statistics of every member (stats.c:449-754), kept on the device once
 * swb_enable_statistics has been called.  Plane ids of swb_get_node_stats / swb_get_link_stats /
 * swb_get_system_stats; times are seconds, *_TIME planes hold the elapsed simulated time (new routing
 * time of the step) at which the maximum occurred, flows cfs, volumes ft3, depths ft.
 */
enum swb_node_stat {
    SWB_NS_SUM_DEPTH = 0,      /* NodeStats.avgDepth before the division by the step count   */
    SWB_NS_MAX_DEPTH, SWB_NS_MAX_DEPTH_TIME, SWB_NS_TIME_FLOODED, SWB_NS_VOL_FLOODED,
    SWB_NS_MAX_PONDED_VOL, SWB_NS_TIME_SURCHARGED, SWB_NS_TOT_LATFLOW, SWB_NS_MAX_LATFLOW,
    SWB_NS_MAX_INFLOW, SWB_NS_MAX_INFLOW_TIME, SWB_NS_MAX_OVERFLOW, SWB_NS_MAX_OVERFLOW_TIME,
    SWB_NS_NONCONV_COUNT,      /* stats_updateConvergenceStats                               */
    SWB_NS_TIME_COURANT,       /* times the node was time-step critical                      */
    /* storage units: sum of volume, max volume, its time, max outflow, evaporation, exfiltration;
       outfalls: sum of flow (periods with flow), max flow, number of such periods               */
    SWB_NS_X_SUM, SWB_NS_X_MAX, SWB_NS_X_MAX_TIME, SWB_NS_X_MAX_FLOW, SWB_NS_X_EVAP, SWB_NS_X_EXFIL,
    SWB_NS_LOAD0,              /* outfalls: total load of pollutant p in plane SWB_NS_LOAD0 + p */
    SWB_NS_PLANES = SWB_NS_LOAD0
};
enum swb_link_stat {
    SWB_LS_MAX_FLOW = 0, SWB_LS_MAX_FLOW_TIME, SWB_LS_MAX_VELOC, SWB_LS_MAX_DEPTH,
    SWB_LS_TIME_FULL_FLOW, SWB_LS_TIME_CAP_LIMITED, SWB_LS_TIME_SURCHARGED, SWB_LS_TIME_FULL_UP,
    SWB_LS_TIME_FULL_DN, SWB_LS_TURN_SIGN, SWB_LS_TURNS, SWB_LS_TIME_COURANT,
    /* conduits */
    SWB_LS_TIME_NORMAL, SWB_LS_TIME_INLET, SWB_LS_TIME_CLASS0,   /* + flow class 0..6 */
    /* pumps (same planes as the conduit-only block above) */
    SWB_LS_PUMP_MIN_FLOW = SWB_LS_TIME_NORMAL, SWB_LS_PUMP_SUM_FLOW, SWB_LS_PUMP_VOLUME,
    SWB_LS_PUMP_UTILIZED, SWB_LS_PUMP_ENERGY, SWB_LS_PUMP_OFF_LOW, SWB_LS_PUMP_OFF_HIGH,
    SWB_LS_PUMP_STARTUPS, SWB_LS_PUMP_PERIODS,
    SWB_LS_PLANES = SWB_LS_TIME_CLASS0 + 7
};
enum swb_system_stat {
    SWB_SS_REPORT_STEPS = 0,   /* ReportStepCount                                            */
    SWB_SS_ROUTING_SPAN,       /* RoutingTimeSpan, s                                         */
    SWB_SS_MAX_OUTFALL_FLOW,   /* MaxOutfallFlow                                             */
    SWB_SS_MIN_DT, SWB_SS_MAX_DT, SWB_SS_ROUTING_TIME, SWB_SS_STEP_COUNT, SWB_SS_TRIALS,  /* TimeStepStats */
    SWB_SS_PLANES
};

/* per-member accumulators kept on the device (massbal.c:517-555, stats.c:522-540) */
typedef struct swb_member_stats {
    double sim_time;          /* elapsed simulated time, s                                  */
    double last_dt;           /* routing step most recently used, s                         */
    double next_dt;           /* variable step for the next call (dynwave_getRoutingStep)   */
    long long steps;          /* routing steps taken                                        */
    long long iterations;     /* sum of Picard iterations (dynwave_execute return values)   */
    long long non_converged;  /* NonConvergeCount (dynwave.c:269)                           */
    int    crit_node, crit_link;  /* arg-min of the last Courant search (dynwave.c:827)     */
} swb_member_stats;

typedef struct swb_network swb_network;   /* device-resident static data       */
typedef struct swb_solver  swb_solver;    /* device-resident state of M members */

const char *swb_last_error(void);
int  swb_version(void);
int  swb_device_count(void);

/* static network -> HBM (SoA, shared by all members).  device = CUDA ordinal. */
int  swb_network_create(const swb_network_desc *desc, const swb_options *opt, int device,
                        swb_network **out);
void swb_network_destroy(swb_network *net);

/* M lockstep members of one network. */
int  swb_solver_create(swb_network *net, int n_members, swb_solver **out);
void swb_solver_destroy(swb_solver *s);
int  swb_solver_members(const swb_solver *s);

/* state exchange; buf is [n_members][items(*n_pollut)] doubles for members
 * [member0, member0+n_members) */
int  swb_set_field(swb_solver *s, int field, int member0, int n_members, const double *buf);
int  swb_get_field(swb_solver *s, int field, int member0, int n_members, double *buf);
/* same value for every member (a single-member image is broadcast) */
int  swb_broadcast_field(swb_solver *s, int field, const double *buf);
int  swb_set_climate(swb_solver *s, double evap_rate, double hydcon_factor);

/* qualrout_init (qualrout.c:63-96): c = init_concen[p] where depth > ZeroDepth else 0 */
int  swb_qual_init(swb_solver *s, const double *init_concen);

/*
 * The reference's per-step calls, one for one (all members in lockstep):
 *   swb_old_state_swap  = routing.c:399-409 + flowrout.c:153-162 (+ routing.c:312-336 quality swap
 *                         when with_quality): new->old, node_initFlows, overflow reset.  The seam
 *                         shim does NOT call this (the host engine has already done it and the
 *                         shim uploads the result); the ensemble driver does.
 *   swb_dynwave_execute = dynwave_execute (dynwave.c:224-262); dt[member]; returns iteration
 *                         counts in iters[member] (may be NULL).
 *   swb_qualrout_execute= qualrout_execute (qualrout.c:100-142).
 *   swb_get_routing_step= dynwave_getRoutingStep (dynwave.c:195-220) for every member.
 */
int  swb_old_state_swap(swb_solver *s, const double *dt, int with_quality);
int  swb_dynwave_execute(swb_solver *s, const double *dt, int *iters);
int  swb_qualrout_execute(swb_solver *s, const double *dt);
int  swb_get_routing_step(swb_solver *s, double fixed_step, double *dt_out);

/*
 * One routing step of all members with HOST buffers on both sides: the ensemble form of what the
 * seam does per step (routing.c:399-409 swap, lateral inflows and quality loads from the host
 * engine, dynwave_execute, qualrout_execute, dynwave_getRoutingStep).  Host arrays use the host
 * layout [member][item(,pollutant)]; copies are asynchronous on the solver's stream and overlap
 * the device-side transposes; buffers from swb_host_alloc are pinned (any host pointer works,
 * pageable memory just copies slower).
 */
typedef struct swb_step_io {
    const double *dt;           /* [M] routing step to take; NULL = the device's variable step   */
    const double *latflow;      /* [M][n_nodes] Node.newLatFlow (required)                        */
    const double *node_losses;  /* [M][n_nodes] Node.losses, or NULL (zero)                       */
    const double *qual_load;    /* [M][n_nodes][n_pollut] external mass-rate preload, or NULL     */
    double *node_depth;         /* out [M][n_nodes] Node.newDepth, or NULL                        */
    double *link_flow;          /* out [M][n_links] Link.newFlow, or NULL                         */
    double *next_dt;            /* out [M] next variable step, or NULL                            */
    int    *iters;              /* out [M] Picard iterations used, or NULL                        */
} swb_step_io;
void *swb_host_alloc(unsigned long long bytes);
void  swb_host_free(void *p);
int   swb_step_host(swb_solver *s, const swb_step_io *io);
/* One routing step of n solvers that hold disjoint member blocks of one ensemble, as a pipelined
 * batch: each solver's host->device copies, routing launch and device->host copies are queued on
 * its own stream, so the copies of one block overlap the routing kernel of another.  Results are
 * identical to calling swb_step_host(solvers[i], &io[i]) for i = 0..n-1; returns when all are done. */
int   swb_step_host_batch(swb_solver *const *solvers, const swb_step_io *io, int n);

/*
 * Ensemble driver ("runoff once, route many"): lateral inflows are evaluated on the device from
 * per-node hydrographs shared by all members and a per-member (scale, time shift) pair:
 *   latflow[m][node] = baseline[node] + scale[m] * sfactor[node] * TS_node(date_m - shift[m])
 * Times use the reference's DateTime encoding (decimal days) so that the interpolation is the
 * reference's own arithmetic: date_m = start_day + (start_secs + (elapsed_ms + 1) / 1000) / 86400
 * (getDateTime, swmm5.c:1543-1552; datetime_addSeconds, datetime.c:381-393), TS linear between
 * breakpoints and 0 outside them (table_tseriesLookup with extend = FALSE, table.c:745-806).
 * Pollutant p enters as concentration concen[node][p] * latflow (routing.c:476-489).
 */
typedef struct swb_inflow_desc {
    int n_inflow_nodes;
    int n_ts_pts;                   /* total breakpoints                                     */
    const int    *node;             /* node index per inflow                                 */
    const int    *ts_start;         /* n_inflow_nodes+1, CSR into ts_t/ts_q                  */
    const double *ts_t;             /* DateTime (days), as the reference stores Tseries x    */
    const double *ts_q;             /* cfs                                                   */
    const double *sfactor;          /* per inflow node                                       */
    const double *baseline;         /* per inflow node, cfs                                  */
    const double *concen;           /* [n_inflow_nodes][n_pollut], mass/ft3 (may be NULL)    */
    const double *member_scale;     /* [n_members] (NULL = 1)                                */
    const double *member_shift;     /* [n_members], days (NULL = 0)                          */
    double start_day;               /* floor(StartDateTime)                                  */
    double start_secs;              /* seconds of day of StartDateTime                       */
    /* ---- everything below is optional (NULL / 0 = absent): baseline patterns, time-varying CONCEN /
     * MASS pollutant inflows and dry-weather flow, evaluated on the device every step exactly like
     * inflow_getExtInflow / inflow_getDwfInflow / getPatternFactor (inflow.c:207-234, 361-392, 456-486)
     * and accumulated in the order of addExternalInflows / addDryWeatherInflows (routing.c:435-575).
     * Member scale and shift act on the FLOW time series only. */
    const double *cfactor;          /* per inflow node: TExtInflow.cFactor (NULL = 1)        */
    const int    *base_pattern;     /* per inflow node: baseline pattern index or -1         */
    int n_patterns;
    const int    *pattern_type;     /* MONTHLY 0, DAILY 1, HOURLY 2, WEEKEND 3 (enums.h)     */
    const double *pattern_factor;   /* [n_patterns][24]                                      */
    int n_qual_inflows;             /* pollutant inflow records; when > 0 `concen` is ignored */
    const int    *q_node, *q_pollut;
    const int    *q_type;           /* 1 CONCEN_INFLOW, 2 MASS_INFLOW (enums.h)              */
    const int    *q_series;         /* index into the series table below or -1               */
    const int    *q_pattern;        /* baseline pattern or -1                                */
    const double *q_cfactor, *q_sfactor, *q_baseline;
    int n_series;                   /* time series used by the pollutant inflows             */
    const int    *series_start;     /* n_series+1, CSR into series_t / series_v              */
    const double *series_t, *series_v;
    int n_dwf;                      /* dry-weather inflow records, a node's records in the   */
    const int    *dwf_node;         /*   order of its TDwfInflow list                        */
    const int    *dwf_param;        /* -1 flow, else pollutant index                         */
    const double *dwf_avg;          /* avgValue, internal units                              */
    const int    *dwf_patterns;     /* [n_dwf][4] pattern index per pattern type or -1       */
    const double *pollut_dwf_concen;/* [n_pollut] Pollut.dwfConcen (NULL = 0)                */
    /* routing interface file ("runoff once, route many"): the records of an inflows file, linear in time between
     * two records exactly like iface_getNumIfaceNodes / getIfaceFlow / getIfaceQual (iface.c:187-275), added after
     * the external and dry-weather inflows like addIfaceInflows (routing.c:736-775).  Flows in internal units,
     * concentrations per PROJECT pollutant (0 where the file carries none).  The same for every member. */
    int n_iface_nodes, n_iface_records;
    const int    *iface_node;       /* [n_iface_nodes] project node of each file node (-1: not in the project) */
    const double *iface_date;       /* [n_iface_records] DateTime of each record, ascending   */
    const double *iface_value;      /* [record][file node][1 + n_pollut]                      */
} swb_inflow_desc;

int  swb_set_inflows(swb_solver *s, const swb_inflow_desc *inflows);

/* ---- control rules and target settings on the device ------------------------------
 * The part of evaluateControlRules (routing.c:269-308) an ensemble member needs every step:
 * link_setTargetSetting (pump start-up / shut-off depths, link.c:604-624), controls_evaluate for rules
 * whose premises test node / link / simulation-time variables against numbers or other variables and
 * whose actions set a status / setting to a number, a curve of the control value, a time series or a PID
 * output (controls.c:495-552, 1086-1164, 1243-1450), then link_setSetting with orifice opening rates and
 * the weir's surcharge coefficient (link.c:626-639, 1729-1762, 2166-2190).  One thread per member walks
 * the rules in order; priorities resolve conflicts like updateActionList (controls.c:1168-1202).
 * RuleStep > 0 also shortens the routing step to land on rule times (routing.c:190-199).  The same pass sets
 * the stage of TIDAL / TIMESERIES outfalls for the step.
 * Not supported (rejected by the flattener): named variables / math expressions and rain-gage premises.
 * Object / attribute / relation codes are the reference's own (controls.c:45-76). */
enum { SWB_RULE_OBJ_SIM = -1, SWB_RULE_OBJ_GAGE = 0, SWB_RULE_OBJ_NODE = 1, SWB_RULE_OBJ_LINK = 2 };   /* TVariable.object as getPremiseVariable
     * stores it (controls.c:680-790): r_GAGE, r_NODE, r_LINK for every link type, -1 for SIMULATION variables */
typedef struct swb_controls_desc {
    int n_rules, n_premises, n_actions;
    double rule_step;               /* RuleStep, s (0 = rules are evaluated every routing step)          */
    const double *rule_priority;    /* per rule                                                          */
    const int    *rule_premise_start;   /* n_rules+1, CSR into the premise arrays                        */
    const int    *rule_then_start, *rule_else_start;  /* n_rules+1 each, CSR into act_then / act_else    */
    const int    *act_then, *act_else;  /* action indices, list order                                    */
    const int    *prem_type;        /* r_AND 2 (also for IF), r_OR 3 (RuleState, controls.c:65)          */
    const int    *prem_lhs_obj, *prem_lhs_index, *prem_lhs_attr;
    const int    *prem_rhs_is_var, *prem_rhs_obj, *prem_rhs_index, *prem_rhs_attr;
    const int    *prem_relation;    /* EQ NE LT LE GT GE = 0..5                                          */
    const double *prem_value;
    const int    *act_rule, *act_link, *act_attr, *act_curve, *act_tseries;
    const double *act_value, *act_kp, *act_ki, *act_kd;
    /* time series used by actions: CSR like the inflow series */
    int n_series;
    const int    *series_start;
    const double *series_t, *series_v;
    /* per link (n_links entries each) */
    const double *pump_y_on, *pump_y_off;   /* Pump.yOn / yOff, ft (0 = none)                            */
    const double *orif_orate;               /* Orifice.orate, s (0 = instantaneous)                      */
    const double *link_time_last_set;       /* Link.timeLastSet at the start (DateTime, days)            */
    /* clock */
    double start_datetime;                  /* StartDateTime (DateTime, days)                            */
    double start_day, start_secs;           /* floor(StartDateTime), seconds of day (as swb_inflow_desc) */
    /* outfalls whose stage varies in time (node.c:1437-1458): evaluated per member at the start of a step */
    int n_stage_nodes;
    const int    *stage_node;               /* node index                                                */
    const int    *stage_kind;               /* 1 TIDAL_OUTFALL (curve by hour of day), 2 TIMESERIES_OUTFALL */
    const int    *stage_table;              /* kind 1: curve index of the network; kind 2: series index above */
} swb_controls_desc;
/* Installs the rules (a second call replaces them); swb_run_steps then evaluates them at the start of every
 * routing step of every member.  Host-fed steps (swb_step_host) keep taking settings from the host. */
int  swb_set_controls(swb_solver *s, const swb_controls_desc *controls);
/* Re-enumerates the members of an ensemble on the device: afterwards member i is what member perm[i] was.  State,
 * clocks, iteration counters, statistics, routing totals, inflow scale / shift and rule state move together, so
 * every result of a member is the same bits as without the call -- only its index changes (the caller keeps the
 * map).  Why: a Picard trial works on the members that have not converged yet (dynwave.c:242-256); when those
 * sit next to each other a warp's 32 loads fill whole 32-byte sectors, when they are scattered each survivor
 * costs a sector of its own.  Trial counts persist from step to step, so sorting the members by the trials they
 * used recently (swb_get_stats: iterations) every few hundred steps keeps the late trials dense.  perm must be
 * a permutation of 0 .. n_members-1. */
int  swb_permute_members(swb_solver *s, const int *perm);
/* Advance every member n_steps routing steps entirely on the device (one cooperative launch):
 * dt from the Courant search, inflows from swb_set_inflows, dynamic wave + quality routing.
 * t_end: members stop stepping once their sim_time reaches it (last step shortened like
 * swmm5.c:541-546). */
int  swb_run_steps(swb_solver *s, int n_steps, double t_end);
int  swb_get_stats(swb_solver *s, int member0, int n_members, swb_member_stats *out);
/* cumulative quality mass-balance terms per member and pollutant, [member][pollutant]:
 * reacted and seepage in mass (sum of rate x dt, massbal.c:517-540), final_storage in mass (:545) */
int  swb_get_massbal(swb_solver *s, int member0, int n_members, double *reacted, double *seepage,
                     double *final_storage);
/* Routing totals per member since the solver was created (removeSystemOutflows
 * routing.c:776-925 + massbal_updateRoutingTotals massbal.c:587-617, kept on the device by
 * swb_run_steps / swb_step_host).  flow: [member][5] volumes (ft3) = {external inflow, flooding,
 * outflow, evaporation loss, seepage loss}; qual: [member][pollutant][6] masses = {external inflow,
 * flooding, outflow, reacted, seepage, moved to final storage}, each weighted exactly as the
 * reference's two half-step updates do.  With the initial / final stored volume and mass (sums of
 * NEW_VOLUME and NEW_VOLUME x NEW_QUAL over nodes and links) these give the continuity errors of
 * massbal_getFlowError / massbal_getQualError (massbal.c:858-960) for every member.  Either may be NULL. */
int  swb_get_routing_totals(swb_solver *s, int member0, int n_members, double *flow, double *qual);
/* conduit-updates performed so far: sum over members of Picard iterations x true conduits -- the metric of
 * BASELINE.json / SURVEY 8(d), i.e. trial SLOTS: a conduit the reference's bypass rule skips in a late trial
 * (dynwave.c:335-345) still counts, for the device and for the reference arm of bench.py alike */
long long swb_conduit_updates(swb_solver *s);

/* ---- one network partitioned over several GPUs (BASELINE.json configs[4]) --------------
 * The reference has no counterpart (one address space, dynwave.c:224-262 loops over all links and
 * nodes).  Here every rank holds the nodes it owns, the far-end ("ghost") nodes of the conduits that
 * cross its border, and every link with at least one owned end; a cut conduit is therefore computed
 * on both sides (dwflow.c:57-293 is a pure function of the two end depths and the conduit's own
 * state).  After every node update of a Picard trial the kernel itself stores the new depth and
 * converged flag of its border nodes into the peers' receive windows (peer memory over NVLink, no
 * host round trip, no NCCL on the data path), followed by a flag; the same exchange carries the OR of
 * the convergence test (dynwave.c:248-251) and, once per step, the MIN of the Courant search
 * (dynwave.c:799-921) and the border concentrations for qualrout.c:253-353.  Only true conduits may
 * be cut (the ordered regulator pass, SURVEY A.4, stays on one rank).  n_members must be 1.
 * Local numbering: nodes [0, n_owned_nodes) are owned, the rest are ghosts; links keep the relative
 * order of the unpartitioned network so every node sum is formed in the reference's order (A.3). */
typedef struct swb_partition_desc {
    int rank, n_ranks;         /* n_ranks <= 8                                                  */
    int n_owned_nodes;
    int n_send;                /* border values this rank publishes after every node update     */
    const int *send_node;      /* local (owned) node                                            */
    const int *send_rank;      /* rank that holds it as a ghost                                 */
    const int *send_slot;      /* slot in that rank's receive window                            */
    int n_recv;
    const int *recv_node;      /* local ghost node filled from receive slot r                   */
    const int *link_owned;     /* per local link: 0 = copy of a cut conduit another rank reports */
    double timeout_s;          /* a peer that does not answer within this time fails the launch
                                  with SWB_ERR_CUDA instead of hanging the device (0 -> 30 s)     */
} swb_partition_desc;
#define SWB_MAX_RANKS 8
#define SWB_WINDOW_HANDLE_BYTES 64
/* allocates this rank's receive window; then exchange handles (any transport: torch.distributed,
 * MPI, a file) and connect every peer before the first step */
int  swb_partition_attach(swb_solver *s, const swb_partition_desc *p);
int  swb_partition_export(swb_solver *s, void *handle /* SWB_WINDOW_HANDLE_BYTES */);
int  swb_partition_connect(swb_solver *s, int peer_rank, const void *handle);
/* halo exchanges completed so far (each is one barrier over all ranks) */
long long swb_partition_exchanges(swb_solver *s);

/* Report-time results: the float32 records the reference writes to the .out file,
 * computed on the device -- node_getResults (node.c:497-528) and link_getResults (link.c:674-724)
 * with the weighting factor of output.c:662-663, f[m] = (reportTime - OldRoutingTime) /
 * (NewRoutingTime - OldRoutingTime), one per member (n_members_total entries).  Only these 4-byte
 * records cross PCIe.  node_out: [n_members][n_nodes][6 + P] = {depth, head, volume, lateral inflow,
 * total inflow, overflow, quality...}; link_out: [n_members][n_links][5 + P] = {flow, depth,
 * velocity, volume, capacity, quality...} (enums.h:200-219), P = n_pollut unless ignore_quality.
 * Either output may be NULL. */
int  swb_get_results(swb_solver *s, const double *f, int member0, int n_members, float *node_out,
                     float *link_out);

/* Known-answer hook for the geometry library (K1b): evaluates one xsect function on the device for
 * n arguments.  fn: 0 AofY 1 WofY 2 RofY 3 YofA 4 RofA 5 SofA 6 AofS 7 dSdA 8 Ycrit (xsect.c:714-1319);
 * params = {yFull,wMax,ywMax,aFull,rFull,sFull,sMax,yBot,aBot,sBot,rBot} (objects.h:581-599). */
int  swb_xsect_eval(int device, int fn, int xs_type, const double *params, int n, const double *args,
                    double *out);

/* device-side time per phase of the persistent kernel since the last reset, ms (thread 0, globaltimer):
 * 0 prologue, 1 link phase, 2 regulator pass, 3 node phase, 4 loop control / compaction,
 * 5 epilogue, 6 quality nodes, 7 quality links, 8 next-step search, 9 halo exchanges (partitioned),
 * 10 the part of 9 spent waiting for the slowest peer */
/*
 * Per-object statistics.  swb_enable_statistics allocates and zeroes the planes;
 * from then on every routing step of swb_run_steps / swb_step_host updates them on the device
 * (stats_updateFlowStats, stats_updateTimeStepStats, stats_updateConvergenceStats,
 * stats_updateCriticalTimeCount) for steps whose new routing time is >= report_start_s.
 *   node_out  [n_members][SWB_NS_PLANES + n_pollut][n_nodes]
 *   link_out  [n_members][SWB_LS_PLANES][n_links]
 *   sys_out   [n_members][SWB_SS_PLANES]
 */
int  swb_enable_statistics(swb_solver *s, double report_start_s);
int  swb_get_node_stats(swb_solver *s, int member0, int n_members, double *node_out);
int  swb_get_link_stats(swb_solver *s, int member0, int n_members, double *link_out);
int  swb_get_system_stats(swb_solver *s, int member0, int n_members, double *sys_out);

int  swb_get_phase_times(swb_solver *s, double *ms, int n, int reset);
/* Profiling aid: n_steps of the phase mask `phases` (csrc/swb_engine.h PH_*) with the debug switches
   `debug` (DBG_*: skip the link or the node phase of the Picard loop); profile != 0 brackets the launch
   with cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly this launch. */
int  swb_debug_run(swb_solver *s, int phases, int n_steps, int debug, int profile);

/* launch bookkeeping for bench.py ("gpu_launches") and device timing of the last call */
long long swb_launch_count(const swb_solver *s);
/* Execution form.  Ensembles of at least `min_members` members (0 = never; default 256, or the environment
   variable SWB_STAGED_MIN_M) on an unpartitioned network run every routing step as a chain of kernels,
   one per phase with its own register budget (csrc/swb_staged.cuh); everything else runs the persistent
   cooperative kernel.  Both forms execute the same phase functions and give bit-identical states (the
   mass-balance sums agree to rounding: their atomic adds are unordered in either form).  Process-wide;
   returns the previous value. */
int  swb_set_staged_min_members(int min_members);
double    swb_last_kernel_ms(const swb_solver *s);
int       swb_sync(swb_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* SWMM_B200_H */
