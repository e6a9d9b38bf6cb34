// swb_report.h -- report-time result extraction: the float32 records the
// reference writes to the .out file, computed on the device from the old / new state so that only
// 4-byte results cross PCIe at report times instead of the full fp64 state.
//
//   node_results <- node_getResults (node.c:497-528)
//   link_results <- link_getResults (link.c:674-724) with link_getVelocity (link.c:821-843)
//
// f is the reference's weighting factor (output.c:662-663):
//   f = (reportTime - OldRoutingTime) / (NewRoutingTime - OldRoutingTime), one value per member.
// Record layout = enums.h:200-219: node {DEPTH, HEAD, VOLUME, LATFLOW, INFLOW, OVERFLOW, QUAL...},
// link {FLOW, DEPTH, VELOCITY, VOLUME, CAPACITY, QUAL...}.
#ifndef SWB_REPORT_H
#define SWB_REPORT_H

#include "swb_dynwave.h"

namespace swb {

SWB_HD int node_record_len(const Net &n) { return 6 + (n.opt.ignore_quality ? 0 : n.nP); }
SWB_HD int link_record_len(const Net &n) { return 5 + (n.opt.ignore_quality ? 0 : n.nP); }

SWB_HD void node_results(const Net &n, const State &s, int i, int m, double f, float *x)
{
    const int M = s.M;
    const size_t ix = SWB_IX(i, m, M);
    const double f1 = 1.0 - f;
    double z = (f1 * s.n_old_depth[ix] + f * s.n_depth[ix]) * n.opt.ucf_length;
    x[0] = (float)z;
    z = n.node_invert[i] * n.opt.ucf_length;
    x[1] = x[0] + (float)z;
    z = (f1 * s.n_old_volume[ix] + f * s.n_volume[ix]) * n.opt.ucf_volume;
    x[2] = (float)z;
    z = (f1 * s.n_old_latflow[ix] + f * s.n_latflow[ix]) * n.opt.ucf_flow;
    x[3] = (float)z;
    z = (f1 * s.n_old_inflow[ix] + f * s.n_inflow[ix]) * n.opt.ucf_flow;
    x[4] = (float)z;
    z = s.n_overflow[ix] * n.opt.ucf_flow;
    x[5] = (float)z;
    if (!n.opt.ignore_quality)
        for (int p = 0; p < n.nP; p++) {
            size_t iq = SWB_IXP(p, i, n.nN, m, M);
            z = f1 * s.n_old_qual[iq] + f * s.n_qual[iq];
            x[6 + p] = (float)z;
        }
}

SWB_HD void link_results(const Net &n, const State &s, int j, int m, double f, float *x, const double *T)
{
    const int M = s.M;
    const size_t ix = SWB_IX(j, m, M);
    const double f1 = 1.0 - f;
    const double oldFlow = s.l_old_flow[ix], newFlow = s.l_flow[ix];
    double y = f1 * s.l_old_depth[ix] + f * s.l_depth[ix];
    double q = f1 * oldFlow + f * newFlow;
    double v = f1 * s.l_old_volume[ix] + f * s.l_volume[ix];
    const int type = n.link_type[j];
    double u = 0.0, c = 0.0;
    if (type == SWB_CONDUIT) {
        if (n.xs_type[j] != XS_DUMMY) {
            Xs x_ = load_xs(n, j);
            double area = xs_a_of_y_ni(x_, y, T);
            // link_getVelocity (link.c:834-841)
            if (!(y <= 0.01)) {
                double flow = q / (double)n.cond_barrels[j];
                if (area > SWB_FUDGE) u = flow / area;
            }
            c = area / x_.aFull;
        }
    }
    else c = s.l_setting[ix];
    // pump flow is not blended between its on and off states (link.c:701-705)
    if (type == SWB_PUMP && oldFlow * newFlow == 0.0) q = (f >= f1) ? newFlow : oldFlow;
    const double dir = (double)n.link_direction[j];
    y *= n.opt.ucf_length;
    v *= n.opt.ucf_volume;
    q *= n.opt.ucf_flow * dir;
    u *= n.opt.ucf_length * dir;
    x[0] = (float)q;
    x[1] = (float)y;
    x[2] = (float)u;
    x[3] = (float)v;
    x[4] = (float)c;
    if (!n.opt.ignore_quality)
        for (int p = 0; p < n.nP; p++) {
            size_t iq = SWB_IXP(p, j, n.nL, m, M);
            c = f1 * s.l_old_qual[iq] + f * s.l_qual[iq];
            x[5 + p] = (float)c;
        }
}

} // namespace swb
#endif
