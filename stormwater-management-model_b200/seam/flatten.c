/*
 * flatten.c -- see flatten.h.  Reads the live reference engine; owns no hydraulics.
 *
 * Field provenance is cited per block (objects.h line numbers of SWMM 5.2.4).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "headers.h"          /* reference: objects.h, globals.h, funcs.h, enums.h */
#include "flatten.h"

static void *keep(swb_flat *f, size_t bytes)
{
    void *p = calloc(bytes ? bytes : 8, 1);
    if (f->n_owned == f->cap_owned) {
        f->cap_owned = f->cap_owned ? 2 * f->cap_owned : 128;
        f->owned = (void **)realloc(f->owned, sizeof(void *) * f->cap_owned);
    }
    f->owned[f->n_owned++] = p;
    return p;
}
#define DARR(n) ((double *)keep(f, sizeof(double) * (size_t)(n)))
#define IARR(n) ((int *)keep(f, sizeof(int) * (size_t)(n)))

void swb_flat_free(swb_flat *f)
{
    int i;
    for (i = 0; i < f->n_owned; i++) free(f->owned[i]);
    free(f->owned);
    memset(f, 0, sizeof(*f));
}

static int curve_len(const TTable *t)
{
    int n = 0;
    const TTableEntry *e = t->firstEntry;
    while (e) { n++; e = e->next; }
    return n;
}

int swb_flatten_network(swb_flat *f)
{
    int nN = Nobjects[NODE], nL = Nobjects[LINK], nP = Nobjects[POLLUT], nC = Nobjects[CURVE];
    int i, j, k, n, rc = SWB_OK;
    swb_network_desc *d = &f->desc;
    swb_options *o = &f->opt;
    memset(f, 0, sizeof(*f));

    /* ---- options (globals.h:68-118), already adjusted by dynwave_validate (dynwave.c:177) */
    o->surcharge_method = SurchargeMethod;
    o->inert_damping    = InertDamping;
    o->normal_flow_ltd  = NormalFlowLtd;
    o->allow_ponding    = AllowPonding;
    o->max_trials       = MaxTrials;
    o->force_main_eqn   = ForceMainEqn;
    o->unit_system      = UnitSystem;
    o->ignore_quality   = IgnoreQuality;
    o->head_tol         = HeadTol;
    o->min_surf_area    = MinSurfArea;
    o->courant_factor   = CourantFactor;
    o->min_route_step   = MinRouteStep;
    o->route_step       = RouteStep;
    o->ucf_length       = UCF(LENGTH);
    o->ucf_volume       = UCF(VOLUME);
    o->ucf_flow         = UCF(FLOW);
    o->evap_rate        = Evap.rate;
    o->hydcon_factor    = Adjust.hydconFactor;

    d->n_nodes = nN; d->n_links = nL; d->n_pollut = nP; d->n_curves = nC;

    /* ---- nodes (objects.h:490-530), outfalls (:535-546), storage (:551-565) */
    {
        int *type = IARR(nN), *deg = IARR(nN), *otype = IARR(nN), *oflap = IARR(nN);
        int *sshape = IARR(nN), *scurve = IARR(nN);
        double *inv = DARR(nN), *fd = DARR(nN), *sd = DARR(nN), *pa = DARR(nN), *fv = DARR(nN),
               *ce = DARR(nN), *a0 = DARR(nN), *a1 = DARR(nN), *a2 = DARR(nN);
        for (i = 0; i < nN; i++) {
            type[i] = Node[i].type;  deg[i] = Node[i].degree;
            inv[i] = Node[i].invertElev;  fd[i] = Node[i].fullDepth;  sd[i] = Node[i].surDepth;
            pa[i] = Node[i].pondedArea;   fv[i] = Node[i].fullVolume; ce[i] = Node[i].crownElev;
            otype[i] = 0; oflap[i] = 0; sshape[i] = 0; scurve[i] = -1;
            if (Node[i].type == OUTFALL) {
                k = Node[i].subIndex;
                oflap[i] = Outfall[k].hasFlapGate;
                if (Outfall[k].type == FREE_OUTFALL)        otype[i] = SWB_FREE_OUTFALL;
                else if (Outfall[k].type == NORMAL_OUTFALL) otype[i] = SWB_NORMAL_OUTFALL;
                else                                        otype[i] = SWB_STAGE_OUTFALL;
            } else if (Node[i].type == STORAGE) {
                k = Node[i].subIndex;
                sshape[i] = Storage[k].shape;  scurve[i] = Storage[k].aCurve;
                a0[i] = Storage[k].a0;  a1[i] = Storage[k].a1;  a2[i] = Storage[k].a2;
            } else if (Node[i].type == DIVIDER) {
                /* dividers behave as junctions under DW (node.c:400-414 is not reached for
                 * true conduits); nothing extra to carry */
            }
        }
        d->node_type = type; d->node_degree = deg; d->node_invert = inv; d->node_full_depth = fd;
        d->node_sur_depth = sd; d->node_ponded_area = pa; d->node_full_volume = fv;
        d->node_crown_elev = ce; d->outfall_type = otype; d->outfall_flap = oflap;
        d->storage_shape = sshape; d->storage_curve = scurve;
        d->storage_a0 = a0; d->storage_a1 = a1; d->storage_a2 = a2;
    }

    /* ---- shape tables for IRREGULAR / CUSTOM / STREET cross sections (objects.h:604-659) */
    {
        int nT = 0;
        for (j = 0; j < nL; j++) {
            int t = Link[j].xsect.type;
            if (t == IRREGULAR || t == CUSTOM || t == STREET_XSECT) nT++;
        }
        d->n_shape_tbls = nT; d->shape_tbl_len = N_TRANSECT_TBL;
        {
            double *at = DARR(nT * N_TRANSECT_TBL), *ht = DARR(nT * N_TRANSECT_TBL),
                   *wt = DARR(nT * N_TRANSECT_TBL);
            int *tn = IARR(nT);
            d->shape_area_tbl = at; d->shape_hrad_tbl = ht; d->shape_width_tbl = wt;
            d->shape_tbl_n = tn;
        }
    }

    /* ---- links (objects.h:664-709), xsects (:581-599), conduits (:714-733), regulators */
    {
        int *type = IARR(nL), *n1 = IARR(nL), *n2 = IARR(nL), *dir = IARR(nL), *flap = IARR(nL);
        double *o1 = DARR(nL), *o2 = DARR(nL), *ql = DARR(nL), *qf = DARR(nL), *ci = DARR(nL),
               *co = DARR(nL), *ca = DARR(nL), *sr = DARR(nL);
        int *xt = IARR(nL), *xc = IARR(nL), *xtab = IARR(nL);
        double *yf = DARR(nL), *wm = DARR(nL), *yw = DARR(nL), *af = DARR(nL), *rf = DARR(nL),
               *sf = DARR(nL), *sm = DARR(nL), *yb = DARR(nL), *ab = DARR(nL), *sb = DARR(nL),
               *rb = DARR(nL);
        int *bar = IARR(nL), *hl = IARR(nL);
        double *len = DARR(nL), *ulen = DARR(nL), *ml = DARR(nL), *rgh = DARR(nL), *slp = DARR(nL), *beta = DARR(nL),
               *qmax = DARR(nL);
        int *pt = IARR(nL), *pc = IARR(nL), *ot = IARR(nL), *wt = IARR(nL), *wcs = IARR(nL),
            *wcc = IARR(nL), *olc = IARR(nL), *olt = IARR(nL), *wrs = IARR(nL);
        double *pmin = DARR(nL), *pmax = DARR(nL), *ocd = DARR(nL), *olen = DARR(nL),
               *wc1 = DARR(nL), *wc2 = DARR(nL), *wec = DARR(nL), *wsl = DARR(nL), *wlen = DARR(nL),
               *olq = DARR(nL), *ole = DARR(nL), *wrw = DARR(nL);
        int nT = 0;
        double *at = (double *)d->shape_area_tbl, *ht = (double *)d->shape_hrad_tbl,
               *wtb = (double *)d->shape_width_tbl;
        int *tn = (int *)d->shape_tbl_n;

        for (j = 0; j < nL; j++) {
            TXsect *x = &Link[j].xsect;
            type[j] = Link[j].type; n1[j] = Link[j].node1; n2[j] = Link[j].node2;
            dir[j] = Link[j].direction; flap[j] = Link[j].hasFlapGate;
            o1[j] = Link[j].offset1; o2[j] = Link[j].offset2; ql[j] = Link[j].qLimit;
            qf[j] = Link[j].qFull; ci[j] = Link[j].cLossInlet; co[j] = Link[j].cLossOutlet;
            ca[j] = Link[j].cLossAvg; sr[j] = Link[j].seepRate;
            xt[j] = x->type; xc[j] = x->culvertCode; xtab[j] = -1;
            yf[j] = x->yFull; wm[j] = x->wMax; yw[j] = x->ywMax; af[j] = x->aFull;
            rf[j] = x->rFull; sf[j] = x->sFull; sm[j] = x->sMax; yb[j] = x->yBot;
            ab[j] = x->aBot; sb[j] = x->sBot; rb[j] = x->rBot;
            bar[j] = 1; pc[j] = -1; wcc[j] = -1; olc[j] = -1;

            if (x->type == IRREGULAR || x->type == CUSTOM || x->type == STREET_XSECT) {
                const double *sa, *sh, *sw; int nt;
                if (x->type == IRREGULAR) {
                    TTransect *t = &Transect[x->transect];
                    sa = t->areaTbl; sh = t->hradTbl; sw = t->widthTbl; nt = N_TRANSECT_TBL;
                } else if (x->type == CUSTOM) {
                    TShape *s = &Shape[Curve[x->transect].refersTo];
                    sa = s->areaTbl; sh = s->hradTbl; sw = s->widthTbl; nt = N_SHAPE_TBL;
                } else {
                    TTransect *t = &Street[x->transect].transect;
                    sa = t->areaTbl; sh = t->hradTbl; sw = t->widthTbl; nt = t->nTbl;
                }
                memcpy(at + nT * N_TRANSECT_TBL, sa, sizeof(double) * N_TRANSECT_TBL);
                memcpy(ht + nT * N_TRANSECT_TBL, sh, sizeof(double) * N_TRANSECT_TBL);
                memcpy(wtb + nT * N_TRANSECT_TBL, sw, sizeof(double) * N_TRANSECT_TBL);
                tn[nT] = nt; xtab[j] = nT++;
            }
            k = Link[j].subIndex;
            switch (Link[j].type) {
              case CONDUIT:
                bar[j] = Conduit[k].barrels; hl[j] = Conduit[k].hasLosses;
                len[j] = link_getLength(j);              /* true length (link.c:808,1195) */
                ulen[j] = Conduit[k].length;
                ml[j] = Conduit[k].modLength; rgh[j] = Conduit[k].roughFactor;
                slp[j] = Conduit[k].slope; beta[j] = Conduit[k].beta; qmax[j] = Conduit[k].qMax;
                break;
              case PUMP:
                pt[j] = Pump[k].type; pc[j] = Pump[k].pumpCurve;
                pmin[j] = Pump[k].xMin; pmax[j] = Pump[k].xMax;
                break;
              case ORIFICE:
                ot[j] = Orifice[k].type; ocd[j] = Orifice[k].cDisch; olen[j] = Orifice[k].length;
                break;
              case WEIR:
                wt[j] = Weir[k].type; wcs[j] = Weir[k].canSurcharge; wcc[j] = Weir[k].cdCurve;
                wc1[j] = Weir[k].cDisch1; wc2[j] = Weir[k].cDisch2; wec[j] = Weir[k].endCon;
                wsl[j] = Weir[k].slope; wlen[j] = Weir[k].length;
                wrw[j] = Weir[k].roadWidth; wrs[j] = Weir[k].roadSurface;
                break;
              case OUTLET:
                olc[j] = Outlet[k].qCurve; olt[j] = Outlet[k].curveType;
                olq[j] = Outlet[k].qCoeff; ole[j] = Outlet[k].qExpon;
                break;
            }
        }
        d->link_type = type; d->link_node1 = n1; d->link_node2 = n2; d->link_direction = dir;
        d->link_has_flap = flap; d->link_offset1 = o1; d->link_offset2 = o2; d->link_q_limit = ql;
        d->link_q_full = qf; d->link_closs_in = ci; d->link_closs_out = co; d->link_closs_avg = ca;
        d->link_seep_rate = sr; d->xs_type = xt; d->xs_culvert = xc; d->xs_table = xtab;
        d->xs_yfull = yf; d->xs_wmax = wm; d->xs_ywmax = yw; d->xs_afull = af; d->xs_rfull = rf;
        d->xs_sfull = sf; d->xs_smax = sm; d->xs_ybot = yb; d->xs_abot = ab; d->xs_sbot = sb;
        d->xs_rbot = rb; d->cond_barrels = bar; d->cond_has_losses = hl; d->cond_length = len; d->cond_user_length = ulen;
        d->cond_mod_length = ml; d->cond_rough_factor = rgh; d->cond_slope = slp;
        d->cond_beta = beta; d->cond_q_max = qmax; d->pump_type = pt; d->pump_curve = pc;
        d->pump_xmin = pmin; d->pump_xmax = pmax; d->orif_type = ot; d->orif_cdisch = ocd;
        d->orif_length = olen; d->weir_type = wt; d->weir_can_surcharge = wcs;
        d->weir_cd_curve = wcc; d->weir_cdisch1 = wc1; d->weir_cdisch2 = wc2; d->weir_end_con = wec;
        d->weir_slope = wsl; d->weir_length = wlen; d->weir_road_width = wrw;
        d->weir_road_surface = wrs; d->outlet_curve = olc;
        d->outlet_curve_type = olt; d->outlet_qcoeff = olq; d->outlet_qexpon = ole;
    }

    /* ---- curves: linked lists (objects.h:87-111) -> CSR */
    {
        int *cs = IARR(nC + 1), *ct = IARR(nC);
        double *cx, *cy;
        n = 0;
        for (i = 0; i < nC; i++) { cs[i] = n; n += curve_len(&Curve[i]); ct[i] = Curve[i].curveType; }
        cs[nC] = n;
        cx = DARR(n); cy = DARR(n);
        for (i = 0; i < nC; i++) {
            const TTableEntry *e = Curve[i].firstEntry;
            k = cs[i];
            while (e) { cx[k] = e->x; cy[k] = e->y; k++; e = e->next; }
        }
        d->n_curve_pts = n; d->curve_start = cs; d->curve_type = ct; d->curve_x = cx; d->curve_y = cy;
    }

    /* ---- pollutants (objects.h:795-809) */
    {
        double *kd = DARR(nP);
        for (i = 0; i < nP; i++) kd[i] = Pollut[i].kDecay;
        d->pollut_kdecay = kd;
    }
    return rc;
}

int swb_field_len(int field, int nN, int nL, int nP)
{
    if (field == SWB_NODE_NEW_QUAL || field == SWB_NODE_OLD_QUAL) return nN * nP;
    if (field == SWB_LINK_NEW_QUAL || field == SWB_LINK_OLD_QUAL || field == SWB_LINK_TOTAL_LOAD)
        return nL * nP;
    return field < SWB_LINK_NEW_FLOW ? nN : nL;
}

/* outfall stage the device should impose this step (node.c:1437-1458) */
static double outfall_stage(int j)
{
    int i = Node[j].subIndex, k;
    double x, y, currentDate;
    switch (Outfall[i].type) {
      case FIXED_OUTFALL: return Outfall[i].fixedStage;
      case TIDAL_OUTFALL:
        k = Outfall[i].tideCurve;
        table_getFirstEntry(&Curve[k], &x, &y);
        currentDate = NewRoutingTime / MSECperDAY;
        x += (currentDate - floor(currentDate)) * 24.0;
        return table_lookup(&Curve[k], x) / UCF(LENGTH);
      case TIMESERIES_OUTFALL:
        k = Outfall[i].stageSeries;
        currentDate = StartDateTime + NewRoutingTime / MSECperDAY;
        return table_tseriesLookup(&Tseries[k], currentDate, TRUE) / UCF(LENGTH);
      default: return Node[j].invertElev;
    }
}

#define NODE_LOOP(expr)  for (i = 0; i < nN; i++) { if (set) { expr = buf[i]; } else buf[i] = (double)(expr); } return 0
#define LINK_LOOP(expr)  for (i = 0; i < nL; i++) { if (set) { expr = buf[i]; } else buf[i] = (double)(expr); } return 0
#define COND_LOOP(expr)  for (i = 0; i < nL; i++) { if (Link[i].type != CONDUIT) { if (!set) buf[i] = 0.0; continue; } \
                             k = Link[i].subIndex; if (set) { expr = buf[i]; } else buf[i] = (double)(expr); } return 0

static int field_rw(int field, double *buf, int set)
{
    int nN = Nobjects[NODE], nL = Nobjects[LINK], nP = Nobjects[POLLUT], i, k, p;
    switch (field) {
      case SWB_NODE_NEW_DEPTH:      NODE_LOOP(Node[i].newDepth);
      case SWB_NODE_OLD_DEPTH:      NODE_LOOP(Node[i].oldDepth);
      case SWB_NODE_NEW_VOLUME:     NODE_LOOP(Node[i].newVolume);
      case SWB_NODE_OLD_VOLUME:     NODE_LOOP(Node[i].oldVolume);
      case SWB_NODE_NEW_LATFLOW:    NODE_LOOP(Node[i].newLatFlow);
      case SWB_NODE_LOSSES:         NODE_LOOP(Node[i].losses);
      case SWB_NODE_INFLOW:         NODE_LOOP(Node[i].inflow);
      case SWB_NODE_OUTFLOW:        NODE_LOOP(Node[i].outflow);
      case SWB_NODE_OVERFLOW:       NODE_LOOP(Node[i].overflow);
      case SWB_NODE_OLD_NET_INFLOW: NODE_LOOP(Node[i].oldNetInflow);
      case SWB_NODE_OLD_LATFLOW:    NODE_LOOP(Node[i].oldLatFlow);
      case SWB_NODE_OLD_INFLOW:     NODE_LOOP(Node[i].oldFlowInflow);
      case SWB_NODE_OUTFALL_STAGE:
        for (i = 0; i < nN; i++)
            if (!set) buf[i] = (Node[i].type == OUTFALL) ? outfall_stage(i) : 0.0;
        return 0;
      case SWB_NODE_STORAGE_EVAP_LOSS:
      case SWB_NODE_STORAGE_EXFIL_LOSS:
      case SWB_NODE_HRT:
        for (i = 0; i < nN; i++) {
            double *v;
            if (Node[i].type != STORAGE) { if (!set) buf[i] = 0.0; continue; }
            k = Node[i].subIndex;
            v = field == SWB_NODE_HRT ? &Storage[k].hrt :
                field == SWB_NODE_STORAGE_EVAP_LOSS ? &Storage[k].evapLoss : &Storage[k].exfilLoss;
            if (set) *v = buf[i]; else buf[i] = *v;
        }
        return 0;
      case SWB_NODE_NEW_QUAL:
      case SWB_NODE_OLD_QUAL:
        for (i = 0; i < nN; i++) for (p = 0; p < nP; p++) {
            double *v = field == SWB_NODE_NEW_QUAL ? &Node[i].newQual[p] : &Node[i].oldQual[p];
            if (set) *v = buf[i * nP + p]; else buf[i * nP + p] = *v;
        }
        return 0;
      case SWB_LINK_NEW_FLOW:       LINK_LOOP(Link[i].newFlow);
      case SWB_LINK_OLD_FLOW:       LINK_LOOP(Link[i].oldFlow);
      case SWB_LINK_NEW_DEPTH:      LINK_LOOP(Link[i].newDepth);
      case SWB_LINK_OLD_DEPTH:      LINK_LOOP(Link[i].oldDepth);
      case SWB_LINK_NEW_VOLUME:     LINK_LOOP(Link[i].newVolume);
      case SWB_LINK_OLD_VOLUME:     LINK_LOOP(Link[i].oldVolume);
      case SWB_LINK_SETTING:        LINK_LOOP(Link[i].setting);
      case SWB_LINK_TARGET_SETTING: LINK_LOOP(Link[i].targetSetting);
      case SWB_LINK_DQDH:           LINK_LOOP(Link[i].dqdh);
      case SWB_LINK_FROUDE:         LINK_LOOP(Link[i].froude);
      case SWB_LINK_FLOW_CLASS:     LINK_LOOP(Link[i].flowClass);
      case SWB_LINK_SURF_AREA1:     LINK_LOOP(Link[i].surfArea1);
      case SWB_LINK_SURF_AREA2:     LINK_LOOP(Link[i].surfArea2);
      case SWB_LINK_BYPASSED:       LINK_LOOP(Link[i].bypassed);
      case SWB_LINK_NORMAL_FLOW:    LINK_LOOP(Link[i].normalFlow);
      case SWB_LINK_INLET_CONTROL:  LINK_LOOP(Link[i].inletControl);
      case SWB_COND_A1:             COND_LOOP(Conduit[k].a1);
      case SWB_COND_A2:             COND_LOOP(Conduit[k].a2);
      case SWB_COND_Q1:             COND_LOOP(Conduit[k].q1);
      case SWB_COND_Q2:             COND_LOOP(Conduit[k].q2);
      case SWB_COND_FULL_STATE:     COND_LOOP(Conduit[k].fullState);
      case SWB_COND_CAPACITY_LIMITED: COND_LOOP(Conduit[k].capacityLimited);
      case SWB_COND_EVAP_LOSS:      COND_LOOP(Conduit[k].evapLossRate);
      case SWB_COND_SEEP_LOSS:      COND_LOOP(Conduit[k].seepLossRate);
      case SWB_ORIF_CORIF: case SWB_ORIF_CWEIR: case SWB_ORIF_HCRIT:
        for (i = 0; i < nL; i++) {
            double *v;
            if (Link[i].type != ORIFICE) { if (!set) buf[i] = 0.0; continue; }
            k = Link[i].subIndex;
            v = field == SWB_ORIF_CORIF ? &Orifice[k].cOrif :
                field == SWB_ORIF_CWEIR ? &Orifice[k].cWeir : &Orifice[k].hCrit;
            if (set) *v = buf[i]; else buf[i] = *v;
        }
        return 0;
      case SWB_REG_SURF_AREA:
        for (i = 0; i < nL; i++) {
            double *v = NULL;
            k = Link[i].subIndex;
            if (Link[i].type == ORIFICE) v = &Orifice[k].surfArea;
            else if (Link[i].type == WEIR) v = &Weir[k].surfArea;
            if (!v) { if (!set) buf[i] = 0.0; continue; }
            if (set) *v = buf[i]; else buf[i] = *v;
        }
        return 0;
      case SWB_WEIR_CSURCHARGE:
        for (i = 0; i < nL; i++) {
            if (Link[i].type != WEIR) { if (!set) buf[i] = 0.0; continue; }
            k = Link[i].subIndex;
            if (set) Weir[k].cSurcharge = buf[i]; else buf[i] = Weir[k].cSurcharge;
        }
        return 0;
      case SWB_LINK_NEW_QUAL: case SWB_LINK_OLD_QUAL: case SWB_LINK_TOTAL_LOAD:
        for (i = 0; i < nL; i++) for (p = 0; p < nP; p++) {
            double *v = field == SWB_LINK_NEW_QUAL ? &Link[i].newQual[p] :
                        field == SWB_LINK_OLD_QUAL ? &Link[i].oldQual[p] : &Link[i].totalLoad[p];
            if (set) *v = buf[i * nP + p]; else buf[i * nP + p] = *v;
        }
        return 0;
    }
    return SWB_ERR_ARG;   /* solver-private fields (Xnode) have no engine-side home */
}

int swb_engine_get_field(int field, double *buf)       { return field_rw(field, buf, 0); }
int swb_engine_set_field(int field, const double *buf) { return field_rw(field, (double *)buf, 1); }

/* ---- inflows (objects.h:430-453, inflow.c) -------------------------------------------------------- */
static int series_len(int k)
{
    int n = 0;
    const TTableEntry *e = Tseries[k].firstEntry;
    while (e) { n++; e = e->next; }
    return n;
}

extern double Qcf[];                        /* swmm5.c: flow units conversion factors (iface.c:23) */

static int flatten_iface(swb_flat *f, swb_inflow_desc *d, const char *path)
{
    FILE *fp = fopen(path, "rt");
    char line[MAXLINE + 1], s1[MAXLINE + 1], s2[MAXLINE + 1];
    int nFilePol = -1, units, nNodes = 0, i, j, k, nP = Nobjects[POLLUT], cap = 0, nRec = 0, w = 1 + nP;
    int *polCol, *nodes;
    double *dates = NULL, *vals = NULL;
    if (!fp) return SWB_ERR_ARG;
    if (!fgets(line, MAXLINE, fp) || !sscanf(line, "%s", s1) || !strcomp(s1, "SWMM5")) { fclose(fp); return SWB_ERR_ARG; }
    fgets(line, MAXLINE, fp);                                   /* title */
    fgets(line, MAXLINE, fp);                                   /* reporting step */
    fgets(line, MAXLINE, fp);                                   /* number of constituents, FLOW included */
    if (sscanf(line, "%d", &nFilePol)) nFilePol--;
    if (nFilePol < 0) { fclose(fp); return SWB_ERR_ARG; }
    fgets(line, MAXLINE, fp);
    if (sscanf(line, "%s %s", s1, s2) < 2 || !strcomp(s1, "FLOW") || (units = findmatch(s2, FlowUnitWords)) < 0) { fclose(fp); return SWB_ERR_ARG; }
    polCol = IARR(nP + 1);
    for (i = 0; i < nP; i++) polCol[i] = -1;
    for (i = 0; i < nFilePol; i++) {
        fgets(line, MAXLINE, fp);
        if (sscanf(line, "%s %s", s1, s2) < 2) { fclose(fp); return SWB_ERR_ARG; }
        j = project_findObject(POLLUT, s1);
        if (j >= 0) polCol[j] = i;
    }
    fgets(line, MAXLINE, fp);
    if (!sscanf(line, "%d", &nNodes) || nNodes <= 0) { fclose(fp); return SWB_ERR_ARG; }
    nodes = IARR(nNodes);
    for (i = 0; i < nNodes; i++) {
        fgets(line, MAXLINE, fp);
        if (!sscanf(line, "%s", s1)) { fclose(fp); return SWB_ERR_ARG; }
        nodes[i] = project_findObject(NODE, s1);
    }
    fgets(line, MAXLINE, fp);                                   /* column headings */
    for (;;) {                                                  /* one record = one line per file node */
        int yr = 0, mon = 0, day = 0, hr = 0, mi = 0, sec = 0, ok = 1;
        if (nRec == cap) {
            cap = cap ? 2 * cap : 256;
            dates = (double *)realloc(dates, sizeof(double) * cap);
            vals = (double *)realloc(vals, sizeof(double) * (size_t)cap * nNodes * w);
        }
        for (i = 0; i < nNodes && ok; i++) {
            char *t;
            double *row = vals + ((size_t)nRec * nNodes + i) * w;
            double filev[64];
            if (feof(fp) || !fgets(line, MAXLINE, fp) || strtok(line, SEPSTR) == NULL) { ok = 0; break; }
#define NEXT_TOK() ((t = strtok(NULL, SEPSTR)) != NULL)
            if (!NEXT_TOK()) { ok = 0; break; } yr = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; } mon = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; } day = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; } hr = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; } mi = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; } sec = atoi(t);
            if (!NEXT_TOK()) { ok = 0; break; }
            row[0] = atof(t) / Qcf[units];
            for (j = 0; j < nFilePol && j < 64; j++) { if (!NEXT_TOK()) { ok = 0; break; } filev[j] = atof(t); }
#undef NEXT_TOK
            if (!ok) break;
            for (k = 0; k < nP; k++) row[1 + k] = (polCol[k] >= 0 && polCol[k] < 64) ? filev[polCol[k]] : 0.0;
        }
        if (!ok) break;
        dates[nRec] = datetime_encodeDate(yr, mon, day) + datetime_encodeTime(hr, mi, sec);
        nRec++;
    }
    fclose(fp);
    {
        double *dd = DARR(nRec + 1), *vv = DARR((size_t)(nRec + 1) * nNodes * w);
        if (nRec) { memcpy(dd, dates, sizeof(double) * nRec); memcpy(vv, vals, sizeof(double) * (size_t)nRec * nNodes * w); }
        free(dates); free(vals);
        d->n_iface_nodes = nNodes; d->n_iface_records = nRec; d->iface_node = nodes; d->iface_date = dd; d->iface_value = vv;
    }
    return SWB_OK;
}

int swb_flatten_inflows(swb_flat *f, swb_inflow_desc *d)
{
    int nN = Nobjects[NODE], nP = Nobjects[POLLUT], nPat = Nobjects[TIMEPATTERN], nTs = Nobjects[TSERIES];
    int i, k, p, n = 0, npts = 0, nq = 0, nd = 0, anyCf = 0, anyPat = 0;
    TExtInflow *x;
    TDwfInflow *w;
    int *node, *start, *bpat, *tsmap;
    double *tt, *tq, *sf, *bl, *cf, *cc;
    memset(d, 0, sizeof(*d));
    for (i = 0; i < nN; i++) {
        for (x = Node[i].extInflow; x; x = x->next) {
            if (x->type == FLOW_INFLOW) { n++; if (x->tSeries >= 0) npts += series_len(x->tSeries); break; }
        }
        for (x = Node[i].extInflow; x; x = x->next) if (x->type != FLOW_INFLOW) nq++;
        for (w = Node[i].dwfInflow; w; w = w->next) nd++;
    }
    node = IARR(n + 1); start = IARR(n + 2); bpat = IARR(n + 1);
    tt = DARR(npts + 1); tq = DARR(npts + 1); sf = DARR(n + 1); bl = DARR(n + 1); cf = DARR(n + 1);
    cc = DARR((size_t)(n + 1) * (nP ? nP : 1));
    k = 0; npts = 0;
    for (i = 0; i < nN; i++) {
        TExtInflow *flow = NULL;
        for (x = Node[i].extInflow; x; x = x->next) if (x->type == FLOW_INFLOW) { flow = x; break; }
        if (!flow) continue;
        node[k] = i; start[k] = npts;
        sf[k] = flow->sFactor; bl[k] = flow->baseline; cf[k] = flow->cFactor; bpat[k] = flow->basePat;
        if (flow->cFactor != 1.0) anyCf = 1;
        if (flow->basePat >= 0) anyPat = 1;
        if (flow->tSeries >= 0) {
            const TTableEntry *e = Tseries[flow->tSeries].firstEntry;
            while (e) { tt[npts] = e->x; tq[npts] = e->y; npts++; e = e->next; }
        }
        /* constant concentrations at the start date: the plain form kept for callers without records */
        for (x = Node[i].extInflow; x; x = x->next)
            if (x->type == CONCEN_INFLOW) cc[(size_t)k * nP + x->param] = inflow_getExtInflow(x, StartDateTime);
        k++;
    }
    start[k] = npts;
    d->n_inflow_nodes = n; d->n_ts_pts = npts; d->node = node; d->ts_start = start; d->ts_t = tt; d->ts_q = tq;
    d->sfactor = sf; d->baseline = bl; d->concen = cc;
    if (anyCf) d->cfactor = cf;
    else for (k = 0; k < n; k++) { /* cFactor == 1: nothing to fold */ }
    if (anyPat) d->base_pattern = bpat;
    {
        int h, mi, s;
        datetime_decodeTime(StartDateTime, &h, &mi, &s);
        d->start_day = floor(StartDateTime);
        d->start_secs = 3600.0 * h + 60.0 * mi + s;
    }
    /* patterns */
    if (nPat > 0) {
        int *pt = IARR(nPat);
        double *pf = DARR((size_t)nPat * 24);
        for (p = 0; p < nPat; p++) {
            pt[p] = Pattern[p].type;
            for (i = 0; i < 24; i++) pf[(size_t)p * 24 + i] = Pattern[p].factor[i];
        }
        d->n_patterns = nPat; d->pattern_type = pt; d->pattern_factor = pf;
    }
    /* pollutant inflow records, a node's records in list order; their time series as a shared table */
    if (nq > 0) {
        int *qn = IARR(nq), *qp = IARR(nq), *qt = IARR(nq), *qs = IARR(nq), *qb = IARR(nq);
        double *qc = DARR(nq), *qf = DARR(nq), *ql = DARR(nq);
        int ns = 0, tot = 0, *ss;
        double *st, *sv;
        tsmap = IARR(nTs + 1);
        for (i = 0; i < nTs; i++) tsmap[i] = -1;
        k = 0;
        for (i = 0; i < nN; i++)
            for (x = Node[i].extInflow; x; x = x->next) {
                if (x->type == FLOW_INFLOW) continue;
                qn[k] = i; qp[k] = x->param; qt[k] = (x->type == CONCEN_INFLOW) ? 1 : 2;
                qb[k] = x->basePat; qc[k] = x->cFactor; qf[k] = x->sFactor; ql[k] = x->baseline;
                qs[k] = -1;
                if (x->tSeries >= 0) {
                    if (tsmap[x->tSeries] < 0) { tsmap[x->tSeries] = ns++; tot += series_len(x->tSeries); }
                    qs[k] = tsmap[x->tSeries];
                }
                k++;
            }
        ss = IARR(ns + 1); st = DARR(tot + 1); sv = DARR(tot + 1);
        tot = 0;
        for (p = 0; p < ns; p++)
            for (i = 0; i < nTs; i++)
                if (tsmap[i] == p) {
                    const TTableEntry *e = Tseries[i].firstEntry;
                    ss[p] = tot;
                    while (e) { st[tot] = e->x; sv[tot] = e->y; tot++; e = e->next; }
                }
        ss[ns] = tot;
        d->n_qual_inflows = nq; d->q_node = qn; d->q_pollut = qp; d->q_type = qt; d->q_series = qs; d->q_pattern = qb;
        d->q_cfactor = qc; d->q_sfactor = qf; d->q_baseline = ql;
        d->n_series = ns; d->series_start = ss; d->series_t = st; d->series_v = sv;
    }
    /* dry-weather inflows */
    if (nd > 0) {
        int *dn = IARR(nd), *dp = IARR(nd), *dpat = IARR((size_t)nd * 4);
        double *da = DARR(nd), *dc = DARR(nP + 1);
        k = 0;
        for (i = 0; i < nN; i++)
            for (w = Node[i].dwfInflow; w; w = w->next) {
                dn[k] = i; dp[k] = w->param; da[k] = w->avgValue;
                for (p = 0; p < 4; p++) dpat[(size_t)k * 4 + p] = w->patterns[p];
                k++;
            }
        for (p = 0; p < nP; p++) dc[p] = Pollut[p].dwfConcen;
        d->n_dwf = nd; d->dwf_node = dn; d->dwf_param = dp; d->dwf_avg = da; d->dwf_patterns = dpat;
        d->pollut_dwf_concen = dc;
    }
    /* routing interface file with inflows ([FILES] USE INFLOWS): the engine keeps only a two-record window of
     * it (iface.c), so the file is read again here, header and records exactly as iface.c:380-600 does */
    if (Finflows.mode == USE_FILE) {
        int rc = flatten_iface(f, d, Finflows.name);
        if (rc) return rc;
    }
    return SWB_OK;
}

/* ---- control rules -----------------------------------------------------------------------------------
 * controls.c keeps its rule base in structs private to that file (controls.c:95-160); the engine exports
 * only the two globals below.  The layouts are restated here field for field so the list can be walked. */
struct swb_TVariable { int object, index, attribute; };
struct swb_TPremise {
    int type, exprIndex;
    struct swb_TVariable lhsVar, rhsVar;
    int relation;
    double value;
    struct swb_TPremise *next;
};
struct swb_TAction {
    int rule, link, attribute, curve, tseries;
    double value, kp, ki, kd, e1, e2;
    struct swb_TAction *next;
};
struct swb_TRule {
    char *ID;
    double priority;
    struct swb_TPremise *firstPremise, *lastPremise;
    struct swb_TAction *thenActions, *elseActions;
};
extern struct swb_TRule *Rules;        /* controls.c:163 */
extern int RuleCount;                  /* controls.c:167 */

int swb_flatten_controls(swb_flat *f, swb_controls_desc *d)
{
    int nL = Nobjects[LINK], nTs = Nobjects[TSERIES];
    int r, j, k, nPrem = 0, nAct = 0, nThen = 0, nElse = 0, ns = 0, tot = 0;
    struct swb_TPremise *p;
    struct swb_TAction *a;
    int *tsmap;
    memset(d, 0, sizeof(*d));
    for (r = 0; r < RuleCount; r++) {
        for (p = Rules[r].firstPremise; p; p = p->next) {
            nPrem++;
            if (p->exprIndex >= 0) return SWB_ERR_UNSUPP;                     /* math expression */
            if (p->lhsVar.object == 0 /* r_GAGE */) return SWB_ERR_UNSUPP;
            if (p->value == MISSING && p->rhsVar.object == 0) return SWB_ERR_UNSUPP;
        }
        for (a = Rules[r].thenActions; a; a = a->next) { nAct++; nThen++; }
        for (a = Rules[r].elseActions; a; a = a->next) { nAct++; nElse++; }
    }
    {
        double *pri = DARR(RuleCount + 1), *pv = DARR(nPrem + 1);
        int *ps = IARR(RuleCount + 2), *ts = IARR(RuleCount + 2), *es = IARR(RuleCount + 2);
        int *at = IARR(nThen + 1), *ae = IARR(nElse + 1);
        int *pt = IARR(nPrem + 1), *lo = IARR(nPrem + 1), *li = IARR(nPrem + 1), *la = IARR(nPrem + 1);
        int *rv = IARR(nPrem + 1), *ro = IARR(nPrem + 1), *ri = IARR(nPrem + 1), *ra = IARR(nPrem + 1);
        int *rel = IARR(nPrem + 1);
        int *ar = IARR(nAct + 1), *al = IARR(nAct + 1), *aa = IARR(nAct + 1), *ac = IARR(nAct + 1), *ats = IARR(nAct + 1);
        double *av = DARR(nAct + 1), *kp = DARR(nAct + 1), *ki = DARR(nAct + 1), *kd = DARR(nAct + 1);
        int ip = 0, ia = 0, it = 0, ie = 0;
        tsmap = IARR(nTs + 1);
        for (j = 0; j < nTs; j++) tsmap[j] = -1;
        for (r = 0; r < RuleCount; r++) {
            int pass;
            pri[r] = Rules[r].priority;
            ps[r] = ip; ts[r] = it; es[r] = ie;
            for (p = Rules[r].firstPremise; p; p = p->next) {
                pt[ip] = p->type; lo[ip] = p->lhsVar.object; li[ip] = p->lhsVar.index; la[ip] = p->lhsVar.attribute;
                rv[ip] = (p->value == MISSING); ro[ip] = p->rhsVar.object; ri[ip] = p->rhsVar.index;
                ra[ip] = p->rhsVar.attribute; rel[ip] = p->relation; pv[ip] = p->value;
                ip++;
            }
            for (pass = 0; pass < 2; pass++)
                for (a = pass ? Rules[r].elseActions : Rules[r].thenActions; a; a = a->next) {
                    ar[ia] = a->rule; al[ia] = a->link; aa[ia] = a->attribute; ac[ia] = a->curve; ats[ia] = -1;
                    av[ia] = a->value; kp[ia] = a->kp; ki[ia] = a->ki; kd[ia] = a->kd;
                    if (a->tseries >= 0) {
                        if (tsmap[a->tseries] < 0) { tsmap[a->tseries] = ns++; tot += series_len(a->tseries); }
                        ats[ia] = tsmap[a->tseries];
                    }
                    if (pass) ae[ie++] = ia; else at[it++] = ia;
                    ia++;
                }
        }
        ps[RuleCount] = ip; ts[RuleCount] = it; es[RuleCount] = ie;
        /* outfalls with a tidal curve or a stage time series (node.c:1437-1458) */
        {
            int nN = Nobjects[NODE], i, nst = 0;
            int *sn, *sk, *stb;
            for (i = 0; i < nN; i++)
                if (Node[i].type == OUTFALL && (Outfall[Node[i].subIndex].type == TIDAL_OUTFALL ||
                                                Outfall[Node[i].subIndex].type == TIMESERIES_OUTFALL)) nst++;
            sn = IARR(nst + 1); sk = IARR(nst + 1); stb = IARR(nst + 1);
            nst = 0;
            for (i = 0; i < nN; i++) {
                if (Node[i].type != OUTFALL) continue;
                k = Node[i].subIndex;
                if (Outfall[k].type == TIDAL_OUTFALL) { sn[nst] = i; sk[nst] = 1; stb[nst] = Outfall[k].tideCurve; nst++; }
                else if (Outfall[k].type == TIMESERIES_OUTFALL) {
                    j = Outfall[k].stageSeries;
                    if (tsmap[j] < 0) { tsmap[j] = ns++; tot += series_len(j); }
                    sn[nst] = i; sk[nst] = 2; stb[nst] = tsmap[j]; nst++;
                }
            }
            d->n_stage_nodes = nst; d->stage_node = sn; d->stage_kind = sk; d->stage_table = stb;
        }
        {
            int h, mi, sec;
            datetime_decodeTime(StartDateTime, &h, &mi, &sec);
            d->start_datetime = StartDateTime;
            d->start_day = floor(StartDateTime);
            d->start_secs = 3600.0 * h + 60.0 * mi + sec;
        }
        d->n_rules = RuleCount; d->n_premises = nPrem; d->n_actions = nAct;
        d->rule_step = RuleStep;
        d->rule_priority = pri; d->rule_premise_start = ps; d->rule_then_start = ts; d->rule_else_start = es;
        d->act_then = at; d->act_else = ae;
        d->prem_type = pt; d->prem_lhs_obj = lo; d->prem_lhs_index = li; d->prem_lhs_attr = la;
        d->prem_rhs_is_var = rv; d->prem_rhs_obj = ro; d->prem_rhs_index = ri; d->prem_rhs_attr = ra;
        d->prem_relation = rel; d->prem_value = pv;
        d->act_rule = ar; d->act_link = al; d->act_attr = aa; d->act_curve = ac; d->act_tseries = ats;
        d->act_value = av; d->act_kp = kp; d->act_ki = ki; d->act_kd = kd;
    }
    {
        int *ss = IARR(ns + 1);
        double *st = DARR(tot + 1), *sv = DARR(tot + 1);
        tot = 0;
        for (k = 0; k < ns; k++)
            for (j = 0; j < nTs; j++)
                if (tsmap[j] == k) {
                    const TTableEntry *e = Tseries[j].firstEntry;
                    ss[k] = tot;
                    while (e) { st[tot] = e->x; sv[tot] = e->y; tot++; e = e->next; }
                }
        ss[ns] = tot;
        d->n_series = ns; d->series_start = ss; d->series_t = st; d->series_v = sv;
    }
    {
        double *yon = DARR(nL + 1), *yoff = DARR(nL + 1), *orate = DARR(nL + 1), *tls = DARR(nL + 1);
        for (j = 0; j < nL; j++) {
            k = Link[j].subIndex;
            if (Link[j].type == PUMP) { yon[j] = Pump[k].yOn; yoff[j] = Pump[k].yOff; }
            if (Link[j].type == ORIFICE) orate[j] = Orifice[k].orate;
            tls[j] = Link[j].timeLastSet;
        }
        d->pump_y_on = yon; d->pump_y_off = yoff; d->orif_orate = orate; d->link_time_last_set = tls;
    }
    return SWB_OK;
}
