// Host check of swb::div_rcp / swb::exact_rcp (csrc/swb_common.h): for divisors that pass the test,
// the three-operation quotient must equal the IEEE division for every dividend tried.
//   divcheck <samples per divisor>  -> prints "divisors D qualified Q mismatches M"
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include "swb_common.h"
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double rnd_double(int emin, int emax)
{
    uint64_t u = rnd();
    double m = 1.0 + (double)(u & 0xFFFFFFFFFFFFFull) / 4503599627370496.0;
    return ldexp(m, emin + (int)((u >> 52) % (unsigned)(emax - emin + 1)));
}
int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 200000;
    long bad = 0, qualified = 0, divisors = 0;
    auto check = [&](double d) {
        divisors++;
        double r = swb::exact_rcp(d);
        if (r == 0.0) return;
        qualified++;
        for (long i = 0; i < n; i++) {
            double x = rnd_double(-40, 40);
            if (i & 1) x = -x;
            if (swb::div_rcp(x, d, r) != x / d) { bad++; if (bad < 8) printf("d=%a x=%a got %a want %a\n", d, x, swb::div_rcp(x, d, r), x / d); }
        }
        for (int i = 0; i < 4096; i++) {               // multiples of d and their neighbours
            double x = i * d;
            if (swb::div_rcp(x, d, r) != x / d) { bad++; if (bad < 8) printf("S d=%a x=%a i=%d got %a want %a\n", d, x, i, swb::div_rcp(x, d, r), x / d); }
            double up = nextafter(x, 1e300), dn = nextafter(x, -1e300);
            if (i && swb::div_rcp(up, d, r) != up / d) bad++;       // (i == 0: denormals, out of scope)
            if (i && swb::div_rcp(dn, d, r) != dn / d) bad++;
        }
    };
    // the table spacings the geometry lookups divide by
    if (swb::exact_rcp(1.0 / 50.0) != 50.0 || swb::exact_rcp(1.0 / 25.0) != 25.0) { printf("table spacing rejected\n"); return 1; }
    const double fixed[] = { 1.0 / 50.0, 1.0 / 25.0, 1.5, 3.0, 0.75, 1.0, 2.0, 400.0, 100.0, 1.25, 2.5, 6.0, 12.0 };
    for (double d : fixed) check(d);
    for (int k = 0; k < 3000; k++) check(rnd_double(-8, 12));           // arbitrary significands
    for (int k = 1; k <= 2000; k++) check(0.25 * k);                      // quarter-foot sizes / lengths
    printf("divisors %ld qualified %ld mismatches %ld\n", divisors, qualified, bad);
    return bad != 0;
}
