// swb_common.h -- constants and macros shared by every device header.
//
// The device headers are plain C++ inline functions marked SWB_HD so that tests/emul can compile
// the very same kernel arithmetic for the host and check it on a machine without a GPU.  That
// host build is test scaffolding only: libswmm_b200.so contains the CUDA path alone and every
// entry point fails with SWB_ERR_CUDA when no device is present.
#ifndef SWB_COMMON_H
#define SWB_COMMON_H

// Inlining policy: every SWB_HD function is force-inlined on the device so that a call with a
// compile-time shape constant folds its switch away and keeps the cross section in registers;
// the few SWB_NI wrappers are real calls, used where the shape is only known at run time
// (generic conduit path, regulators, outfall boundary) to keep the code footprint bounded.
//   SWB_FI  small hot functions of the specialised conduit path (force-inlined)
//   SWB_HD  everything else (the compiler decides)
//   SWB_NI  real calls
#ifdef __CUDACC__
#define SWB_FI __host__ __device__ __forceinline__
#define SWB_HD __host__ __device__ inline
#define SWB_NI __host__ __device__ __noinline__ inline
#else
#define SWB_FI inline
#define SWB_HD inline
#define SWB_NI __attribute__((noinline)) inline
#endif

// consts.h:31-50 and the module constants of dynwave.c:60-66, dwflow.c:37, qualrout.c:40-41.
// PI is the reference's 10-digit value on purpose.
#define SWB_FUDGE       0.0001
#define SWB_TINY        1.E-6
#define SWB_ZERO        1.E-10
#define SWB_PI          3.141592654
#define SWB_GRAVITY     32.2
#define SWB_OMEGA       0.5
#define SWB_MAXVELOCITY 50.
#define SWB_MINTIMESTEP 0.001
#define SWB_ZERO_VOLUME 0.0353147
#define SWB_ZERO_DEPTH  0.003281
#define SWB_M3_PER_FT3  0.028317

// macros.h:25-33 semantics: MIN returns x when equal, SGN(0) = +1
#define SWB_MIN(x, y) (((x) <= (y)) ? (x) : (y))
#define SWB_MAX(x, y) (((x) >= (y)) ? (x) : (y))
#define SWB_SGN(x)    (((x) < 0) ? (-1) : (1))

#define SWB_MAX_POLLUT 8       // pollutants carried per member (register accumulators)

#include <string.h>
#include <math.h>
namespace swb {
// ---- exact division by a divisor known in advance ------------------------------------------------
// x / d == fma(fma(-q, d, x), r, q) with q = x * r and r = RN(1 / d), correctly rounded for EVERY
// finite x away from over/underflow, whenever exact_rcp_ok(d) holds (Markstein's correction step).
// With d * r = 1 + e, t = |e| / 2^-53 and m the significand of d in [1, 2):
//   * q is within 0.5 + t * m_Q / 2 < 1 ulp of Q = x / d when t <= 0.5, so the residual
//     x - q * d is exactly representable and the inner fma is exact;
//   * the outer fma then rounds Q + (Q - q) * e, which is within (0.5 + t) * t * 2^-53 ulp of Q,
//     while a quotient of two doubles is at least 2^-53 / m ulp away from any rounding midpoint:
//     same rounding whenever (0.5 + t) * t < 1 / m.
// Three dependent operations instead of the ~18 of the generic division sequence (reciprocal
// seed, four Newton steps, range check).  The reference's own `/` is what this reproduces bit for
// bit; divisors that fail the test keep the plain division.
SWB_FI double div_rcp(double x, double d, double r)
{
    double q = x * r;
    double e = fma(-q, d, x);
    return fma(e, r, q);
}
// host-side test for a divisor (network set-up); returns RN(1 / d) or 0 when d does not qualify
inline double exact_rcp(double d)
{
    if (!(d > 1.0e-100 && d < 1.0e100)) return 0.0;
    const double r = 1.0 / d;
    const double t = fabs(fma(d, r, -1.0)) / 1.1102230246251565e-16;      // exact residual / 2^-53
    int ex;
    const double m = 2.0 * frexp(d, &ex);
    return (t <= 0.5 && (0.5 + t) * t < 1.0 / m) ? r : 0.0;
}
// r == 0 marks "no exact reciprocal": warp-uniform choice (r is a per-link constant)
SWB_FI double div_by(double x, double d, double r) { return r != 0.0 ? div_rcp(x, d, r) : x / d; }

// order-preserving integer image of a non-negative double (for integer atomicMin)
SWB_HD unsigned long long dbits(double v)
{
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long u; memcpy(&u, &v, sizeof(u)); return u;
#endif
}
SWB_HD double dfrombits(unsigned long long u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double v; memcpy(&v, &u, sizeof(v)); return v;
#endif
}
}

#endif
