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
namespace swb {
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
