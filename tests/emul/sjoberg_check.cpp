// Host check of the device-side slot exponent polynomial (csrc/swb_dynwave.h: dw_sjoberg_exponent_poly):
// exp(-p(y)) against the reference expression exp(-pow(y, 2.4)) over the range the caller can reach.
// Prints the largest difference in ulps.
#include <cstdio>
#include <cmath>
#include "swb_dynwave.h"
int main()
{
    double worst = 0.0, at = 0.0;
    const int n = 400000;
    for (int i = 0; i <= n; i++) {
        double y = 0.985257 + (1.78 - 0.985257) * i / n;
        double a = exp(-swb::dw_sjoberg_exponent_poly(y)), b = exp(-pow(y, 2.4));
        double ulp = nextafter(b, 1.0) - b;
        double d = fabs(a - b) / ulp;
        if (d > worst) { worst = d; at = y; }
    }
    printf("max difference %.2f ulp at yNorm = %.6f\n", worst, at);
    return worst <= 12.0 ? 0 : 1;
}
