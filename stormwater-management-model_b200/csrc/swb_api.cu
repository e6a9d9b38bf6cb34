// swb_api.cu -- libswmm_b200.so: CUDA (sm_100a) backend of the C-ABI in include/swmm_b200.h.
//
// One cooperative, persistent kernel (swb_route_kernel) executes whole routing steps: every CTA
// stages the 12 KB of normalised cross-section tables into shared memory once, then runs the
// phase sequence of swb_engine.h with cooperative-groups grid barriers between phases.  The grid
// is sized to the machine (resident CTAs x SM count, trimmed so that threads % members == 0 and
// never larger than the work), not to the problem.
//
// There is no host fallback in this file: every entry point fails with SWB_ERR_CUDA when no
// device is usable.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <string>
#include "swb_state.h"
#include "swb_engine.h"

namespace cg = cooperative_groups;
using namespace swb;

#ifndef SWB_BLOCK
#define SWB_BLOCK 512
#endif
#ifndef SWB_MIN_BLOCKS
#define SWB_MIN_BLOCKS 2     // 64 registers/thread, 32 resident warps per SM: measured best (profiles/README.md)
#endif

#define SWB_MAX_MEMBERS 8192

struct CudaCtx {
    int tid, G, lane, block_size;
    const double *T;
    int warp_size, warp_lane;
    __device__ __forceinline__ unsigned long long next_ticket(unsigned long long *p)
    {
        unsigned long long t = 0;
        if (warp_lane == 0) t = atomicAdd(p, 1ull);
        return __shfl_sync(0xffffffffu, t, 0);
    }
    int *alive_list;            // shared memory, SWB_MAX_MEMBERS entries
    int *scan;                  // shared memory, 1 + warps entries
    // Ordered stream compaction of the members for which alive(m) holds; every CTA computes the
    // same list (ballot + popc within warps, warp totals through shared memory).
    template <class Pred>
    __device__ __forceinline__ int compact_members(int M, Pred alive)
    {
        const int lane_id = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (threadIdx.x == 0) scan[0] = 0;
        __syncthreads();
        for (int m0 = 0; m0 < M; m0 += blockDim.x) {
            int mm = m0 + threadIdx.x;
            bool flag = (mm < M) && alive(mm);
            unsigned b = __ballot_sync(0xffffffffu, flag);
            if (lane_id == 0) scan[1 + wid] = __popc(b);
            __syncthreads();
            int off = scan[0];
            for (int w = 0; w < wid; w++) off += scan[1 + w];
            if (flag) alive_list[off + __popc(b & ((1u << lane_id) - 1u))] = mm;
            __syncthreads();
            if (threadIdx.x == 0) { int t = scan[0]; for (int w = 0; w < nw; w++) t += scan[1 + w]; scan[0] = t; }
            __syncthreads();
        }
        return scan[0];
    }
    __device__ __forceinline__ void grid_sync() { cg::this_grid().sync(); }
    __device__ __forceinline__ unsigned long long now_ns()
    { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ bool block_or(bool b) { return __syncthreads_or(b ? 1 : 0) != 0; }
    __device__ __forceinline__ void atomic_min_u64(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
    __device__ __forceinline__ void atomic_add_f64(double *p, double v) { atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_min_i32(int *p, int v) { atomicMin(p, v); }
    __device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
    {
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = w < v ? w : v;
        }
        return v;
    }
    __device__ __forceinline__ double warp_sum_f64(double v)
    {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

__global__ void __launch_bounds__(SWB_BLOCK, SWB_MIN_BLOCKS)
swb_route_kernel(const __grid_constant__ Net net, const __grid_constant__ State st,
                 const __grid_constant__ RunArgs args)
{
    __shared__ double tab[XT_TOTAL];
    extern __shared__ int s_alive[];          // st.M entries (dynamic: keeps the L1 carve-out large)
    __shared__ int s_scan[1 + SWB_BLOCK / 32];
    for (int i = threadIdx.x; i < XT_TOTAL; i += blockDim.x) tab[i] = net.xs_tables[i];
    __syncthreads();
    CudaCtx ctx;
    ctx.alive_list = s_alive;
    ctx.warp_size = 32;
    ctx.warp_lane = threadIdx.x & 31;
    ctx.scan = s_scan;
    ctx.tid = blockIdx.x * blockDim.x + threadIdx.x;
    ctx.G = gridDim.x * blockDim.x;
    ctx.lane = threadIdx.x;
    ctx.block_size = blockDim.x;
    ctx.T = tab;
    engine_run(net, st, args, ctx);
}

__global__ void swb_xsect_kernel(int fn, Xs x, int n, const double *tables, const double *args, double *out)
{
    __shared__ double tab[XT_TOTAL];
    for (int i = threadIdx.x; i < XT_TOTAL; i += blockDim.x) tab[i] = tables[i];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = xs_eval(fn, x, args[i], tab);
}

// host layout [m][item][p]  ->  device layout [(p * items + item)][m]   (32 x 32 shared-memory tiles)
__global__ void swb_transpose_in_kernel(double *dst, const double *src, int M, int items, int planes)
{
    __shared__ double tile[32][33];
    const int cols = items * planes;                  // src is M rows x cols
    int c0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int m = m0 + r, c = c0 + threadIdx.x;
        if (m < M && c < cols) tile[r][threadIdx.x] = src[(size_t)m * cols + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, m = m0 + threadIdx.x;
        if (m < M && c < cols) {
            int item = c / planes, p = c - item * planes;
            dst[((size_t)p * items + item) * M + m] = tile[threadIdx.x][r];
        }
    }
}
// device layout [item][m] -> host layout [m][item]
__global__ void swb_transpose_out_kernel(double *dst, const double *src, int M, int items)
{
    __shared__ double tile[32][33];
    int i0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, m = m0 + threadIdx.x;
        if (i < items && m < M) tile[r][threadIdx.x] = src[(size_t)i * M + m];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int m = m0 + r, i = i0 + threadIdx.x;
        if (i < items && m < M) dst[(size_t)m * items + i] = tile[threadIdx.x][r];
    }
}

namespace swb { namespace backend {

static std::string cuda_err(const char *what, cudaError_t e)
{
    return std::string(what) + ": " + cudaGetErrorString(e);
}

static int g_device = -1, g_sms = 0, g_blocks_per_sm = 0;
static cudaEvent_t g_ev0, g_ev1;

static int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static bool init(int device, std::string &err)
{
    int n = device_count();
    if (n <= 0) { err = "no CUDA device available (libswmm_b200 has no CPU path)"; return false; }
    if (device < 0 || device >= n) { err = "CUDA device ordinal out of range"; return false; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { err = cuda_err("cudaSetDevice", e); return false; }
    if (g_device != device) {
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, device);
        if (e != cudaSuccess) { err = cuda_err("cudaGetDeviceProperties", e); return false; }
        if (!p.cooperativeLaunch) { err = "device does not support cooperative launch"; return false; }
        g_sms = p.multiProcessorCount;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_blocks_per_sm, swb_route_kernel, SWB_BLOCK,
                                                          sizeof(int) * SWB_MAX_MEMBERS);
        if (e != cudaSuccess || g_blocks_per_sm < 1) { err = cuda_err("occupancy query", e); return false; }
        if (g_device < 0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
        g_device = device;
    }
    return true;
}

static void *alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    return p;
}
static void free_(void *p) { cudaFree(p); }
static void upload(void *d, const void *s, size_t b) { if (b) cudaMemcpy(d, s, b, cudaMemcpyHostToDevice); }
static void download(void *d, const void *s, size_t b) { if (b) cudaMemcpy(d, s, b, cudaMemcpyDeviceToHost); }
static void zero(void *d, size_t b) { if (b) cudaMemset(d, 0, b); }

static void *host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    return p;
}
static void host_free(void *p) { if (p) cudaFreeHost(p); }
static void h2d_async(void *d, const void *s, size_t b) { if (b) cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, 0); }
static void d2h_async(void *d, const void *s, size_t b) { if (b) cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, 0); }
static void transpose_in(double *dev, const double *stage, int M, int items, int planes)
{
    dim3 grid((items * planes + 31) / 32, (M + 31) / 32), block(32, 8);
    swb_transpose_in_kernel<<<grid, block>>>(dev, stage, M, items, planes);
}
static void transpose_out(double *stage, const double *dev, int M, int items)
{
    dim3 grid((items + 31) / 32, (M + 31) / 32), block(32, 8);
    swb_transpose_out_kernel<<<grid, block>>>(stage, dev, M, items);
}

static bool sync(std::string &err)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { err = cuda_err("cudaDeviceSynchronize", e); return false; }
    return true;
}

static bool xsect_eval(int device, int fn, const Xs &x, int n, const double *args, double *out, std::string &err)
{
    if (!init(device, err)) return false;
    static const double tab[] = { SWB_XS_TABLE_DATA };
    double *dT = (double *)alloc(sizeof(tab)), *dA = (double *)alloc(sizeof(double) * n),
           *dO = (double *)alloc(sizeof(double) * n);
    upload(dT, tab, sizeof(tab));
    upload(dA, args, sizeof(double) * n);
    if (n > 0) swb_xsect_kernel<<<(n + 127) / 128, 128>>>(fn, x, n, dT, dA, dO);
    cudaError_t e = cudaDeviceSynchronize();
    download(out, dO, sizeof(double) * n);
    free_(dT); free_(dA); free_(dO);
    if (e != cudaSuccess) { err = cuda_err("swb_xsect_kernel", e); return false; }
    return true;
}

// CTAs to launch: all that can be co-resident, but (i) threads % M == 0 so a thread keeps one
// member, (ii) no more threads than (objects x members), (iii) at least one CTA per 256 members.
static int pick_blocks(int M, int maxItems)
{
    long long maxBlocks = (long long)g_sms * g_blocks_per_sm;
    long long work = ((long long)maxItems * M + SWB_BLOCK - 1) / SWB_BLOCK;
    long long blocks = std::max(1LL, std::min(maxBlocks, work));
    int g = M, b = SWB_BLOCK;                 // gcd(M, SWB_BLOCK)
    while (b) { int t = g % b; g = b; b = t; }
    long long unit = M / g;                   // blocks must be a multiple of this
    blocks = (blocks / unit) * unit;
    if (blocks < unit) blocks = unit;
    return (int)blocks;
}

static bool launch(const Net &net, const State &st, const RunArgs &args, int device, float *ms, std::string &err)
{
    if (!init(device, err)) return false;
    int blocks = pick_blocks(st.M, std::max(net.nN, net.nL));
    if ((long long)blocks > (long long)g_sms * g_blocks_per_sm) {
        err = "ensemble too wide for one cooperative launch on this device";
        return false;
    }
    void *kargs[] = { (void *)&net, (void *)&st, (void *)&args };
    cudaEventRecord(g_ev0, 0);
    size_t dyn = sizeof(int) * (size_t)((st.M + 31) / 32 * 32);
    cudaError_t e = cudaLaunchCooperativeKernel((void *)swb_route_kernel, dim3(blocks), dim3(SWB_BLOCK), kargs, dyn, 0);
    if (e != cudaSuccess) { err = cuda_err("cudaLaunchCooperativeKernel", e); return false; }
    cudaEventRecord(g_ev1, 0);
    e = cudaEventSynchronize(g_ev1);
    if (e != cudaSuccess) { err = cuda_err("swb_route_kernel", e); return false; }
    cudaEventElapsedTime(ms, g_ev0, g_ev1);
    return true;
}

} }

#include "swb_api_impl.h"
