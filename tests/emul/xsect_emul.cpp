// Host build of the device geometry library (csrc/swb_xsect.h) for CPU-only tests.
// Same signature as oracle/refhook.c:refhook_xsect_eval so a test can diff the two directly.
#include "swb_xsect.h"
using namespace swb;
static const double g_tab[] = { SWB_XS_TABLE_DATA };
extern "C" void emul_xsect_eval(int fn, int type, const double *p, int n, const double *arg, double *out)
{
    Xs x{};
    x.type = type; x.ntbl = 0; x.atbl = x.rtbl = x.wtbl = nullptr;
    x.yFull = p[0]; x.wMax = p[1]; x.ywMax = p[2]; x.aFull = p[3]; x.rFull = p[4];
    x.sFull = p[5]; x.sMax = p[6]; x.yBot = p[7]; x.aBot = p[8]; x.sBot = p[9]; x.rBot = p[10];
    x.rYFull = swb::exact_rcp(x.yFull);
    const double *T = g_tab;
    for (int i = 0; i < n; i++) switch (fn) {
        case 0: out[i] = xs_a_of_y(x, arg[i], T); break;
        case 1: out[i] = xs_w_of_y(x, arg[i], T); break;
        case 2: out[i] = xs_r_of_y(x, arg[i], T); break;
        case 3: out[i] = xs_y_of_a(x, arg[i], T); break;
        case 4: out[i] = xs_r_of_a(x, arg[i], T); break;
        case 5: out[i] = xs_s_of_a(x, arg[i], T); break;
        case 6: out[i] = xs_a_of_s(x, arg[i], T); break;
        case 7: out[i] = xs_dsda(x, arg[i], T); break;
        case 8: out[i] = xs_ycrit(x, arg[i], T); break;
    }
}
