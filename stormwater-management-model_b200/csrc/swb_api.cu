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
#include <cuda_profiler_api.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include "swb_state.h"
#include "swb_engine.h"
#include "swb_report.h"

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
    // one atomic for the whole CTA: returns the first of blockDim / 32 consecutive tiles
    unsigned long long *blk_ticket;        // shared memory, two slots (alternating rounds)
    int blk_round;
    __device__ __forceinline__ unsigned long long next_ticket_block(unsigned long long *p)
    {
        unsigned long long *slot = blk_ticket + (blk_round & 1);
        blk_round++;
        if (threadIdx.x == 0) *slot = atomicAdd(p, (unsigned long long)(blockDim.x >> 5));
        __syncthreads();
        return *slot;
    }
    // the same ticket in two halves: the atomic is issued now, its value is broadcast later
    __device__ __forceinline__ unsigned long long ticket_issue(unsigned long long *p)
    {
        unsigned long long t = 0;
        if (warp_lane == 0) t = atomicAdd(p, 1ull);
        return t;
    }
    __device__ __forceinline__ unsigned long long ticket_take(unsigned long long pending)
    { return __shfl_sync(0xffffffffu, pending, 0); }
    __device__ __forceinline__ void prefetch(const void *p)
    { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
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
    // src is R rows x C columns (row-major); writes dst[row(c) * R + r] = src[r * C + c] with
    // row(c) = c, or for planes > 0 (column c = item * planes + p) row(c) = p * (C / planes) + item.
    // Grid-wide: CTAs stride over 32 x 32 tiles staged in shared memory, both sides coalesced.
    double (*tile)[33];
    double *Tw;
    // geometry tables -> shared memory (after the input transposes, which borrow the same bytes)
    __device__ __forceinline__ void load_tables(const double *src)
    {
        __syncthreads();
        for (int i = threadIdx.x; i < XT_TOTAL; i += blockDim.x) Tw[i] = src[i];
        __syncthreads();
    }
    __device__ __forceinline__ void transpose(double *dst, const double *src, int R, int C, int planes)
    {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const int tilesC = (C + 31) / 32, total = ((R + 31) / 32) * tilesC;
        const int items = planes > 0 ? C / planes : 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int r0 = (t / tilesC) * 32, c0 = (t % tilesC) * 32;
            for (int r = ty; r < 32; r += nw)
                if (r0 + r < R && c0 + tx < C) tile[r][tx] = src[(size_t)(r0 + r) * C + c0 + tx];
            __syncthreads();
            for (int c = ty; c < 32; c += nw) {
                int cc = c0 + c;
                if (cc < C && r0 + tx < R) {
                    if (planes > 0) { int item = cc / planes; cc = (cc - item * planes) * items + item; }
                    dst[(size_t)cc * R + r0 + tx] = tile[tx][c];
                }
            }
            __syncthreads();
        }
    }
    // ---- partitioned network: system-scope accesses to windows that peers write over NVLink
    __device__ __forceinline__ void fence_system() { __threadfence_system(); }
    __device__ __forceinline__ void block_sync() { __syncthreads(); }
    __device__ __forceinline__ void store_release_sys(unsigned long long *p, unsigned long long v)
    { asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
    __device__ __forceinline__ unsigned long long load_acquire_sys(const unsigned long long *p)
    { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
    __device__ __forceinline__ double load_sys_f64(const double *p)
    { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
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
    __shared__ unsigned long long s_blk_ticket[2];
    static_assert(XT_TOTAL >= 32 * 33, "the transpose tile aliases the geometry tables");
    CudaCtx ctx;
    ctx.alive_list = s_alive;
    ctx.warp_size = 32;
    ctx.warp_lane = threadIdx.x & 31;
    ctx.scan = s_scan;
    ctx.blk_ticket = s_blk_ticket;
    ctx.blk_round = 0;
    ctx.tile = (double (*)[33])tab;      // used only before load_tables() and after the last step
    ctx.Tw = tab;
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

// Report-time extraction: one thread per (object, member), members fastest, so every fp64 load is
// coalesced; each thread writes its own float32 record (32 B for a node with two pollutants).
__global__ void swb_report_kernel(const __grid_constant__ Net net, const __grid_constant__ State st,
                                  const double *f, int m0, int nm, int links, float *out)
{
    __shared__ double tab[XT_TOTAL];
    if (links) {
        for (int i = threadIdx.x; i < XT_TOTAL; i += blockDim.x) tab[i] = net.xs_tables[i];
        __syncthreads();
    }
    const int nItems = links ? net.nL : net.nN;
    const int rec = links ? link_record_len(net) : node_record_len(net);
    const long long total = (long long)nItems * nm;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int item = (int)(t / nm), mm = (int)(t - (long long)item * nm), m = m0 + mm;
        float *x = out + ((size_t)mm * nItems + item) * rec;
        if (links) link_results(net, st, item, m, f[m], x, tab);
        else node_results(net, st, item, m, f[m], x);
    }
}

namespace swb { namespace backend {

static std::string cuda_err(const char *what, cudaError_t e)
{
    return std::string(what) + ": " + cudaGetErrorString(e);
}

// per-device state (a process may hold solvers on several devices: every entry point that allocates
// or launches calls init(device) first, which makes that device current and fills its record)
struct DevInfo { bool ready; int sms, blocks_per_sm; size_t persist_max, persist_set; cudaEvent_t ev0, ev1; };
#define SWB_MAX_DEVICES 64
static DevInfo g_dev[SWB_MAX_DEVICES];
static thread_local cudaStream_t g_stream = 0;   // stream the asynchronous operations below are queued on
static thread_local std::string g_cuda_msg;
static const char *last_error() { return g_cuda_msg.c_str(); }
static bool ok(const char *what, cudaError_t e)
{
    if (e == cudaSuccess) return true;
    g_cuda_msg = cuda_err(what, e);
    cudaGetLastError();                          // clear the non-sticky error state
    return false;
}

static void *stream_create()
{
    cudaStream_t st = 0;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return (void *)st;
}
static void stream_destroy(void *st) { if (st) cudaStreamDestroy((cudaStream_t)st); }
static void use_stream(void *st) { g_stream = (cudaStream_t)st; }

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
    if (device < 0 || device >= n || device >= SWB_MAX_DEVICES) { err = "CUDA device ordinal out of range"; return false; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { err = cuda_err("cudaSetDevice", e); return false; }
    DevInfo &D = g_dev[device];
    if (!D.ready) {
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, device);
        if (e != cudaSuccess) { err = cuda_err("cudaGetDeviceProperties", e); return false; }
        if (!p.cooperativeLaunch) { err = "device does not support cooperative launch"; return false; }
        D.sms = p.multiProcessorCount;
        D.persist_max = (size_t)p.persistingL2CacheMaxSize;
        // static (tables, transpose tile) + dynamic (member list, up to 32 KB) exceeds the 48 KB default
        e = cudaFuncSetAttribute(swb_route_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(int) * SWB_MAX_MEMBERS));
        if (e != cudaSuccess) { err = cuda_err("cudaFuncSetAttribute", e); return false; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D.blocks_per_sm, swb_route_kernel, SWB_BLOCK,
                                                          sizeof(int) * SWB_MAX_MEMBERS);
        if (e != cudaSuccess || D.blocks_per_sm < 1) { err = cuda_err("occupancy query", e); return false; }
        if ((e = cudaEventCreate(&D.ev0)) != cudaSuccess || (e = cudaEventCreate(&D.ev1)) != cudaSuccess) {
            err = cuda_err("cudaEventCreate", e); return false;
        }
        D.ready = true;
    }
    return true;
}

static void *alloc(size_t bytes)
{
    void *p = nullptr;
    if (!ok("cudaMalloc", cudaMalloc(&p, bytes ? bytes : 8))) return nullptr;
    return p;
}
static void free_(void *p) { cudaFree(p); }
static bool upload(void *d, const void *s, size_t b) { return !b || ok("cudaMemcpy(H2D)", cudaMemcpy(d, s, b, cudaMemcpyHostToDevice)); }
static bool download(void *d, const void *s, size_t b) { return !b || ok("cudaMemcpy(D2H)", cudaMemcpy(d, s, b, cudaMemcpyDeviceToHost)); }
static bool zero(void *d, size_t b) { return !b || ok("cudaMemset", cudaMemset(d, 0, b)); }
static bool copy2d(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, bool to_device)
{
    if (!width || !height) return true;
    // contiguous on both sides (all members of a field, or a single model): one linear copy -- a 2-D copy
    // of 20 000 eight-byte rows costs ~0.3 ms, 60 of them per routing step made the drop-in seam 3x slower
    if (dpitch == width && spitch == width)
        return ok("cudaMemcpy", cudaMemcpy(d, s, width * height, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
    return ok("cudaMemcpy2D", cudaMemcpy2D(d, dpitch, s, spitch, width, height,
                                           to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
}

// Pin [base, base + bytes) in L2 for the kernels launched on `stream`: accesses inside the window are
// "persisting" (not displaced by normal or streaming lines).  The window is the static network
// (Net::arena); the ensemble state streams past it.  Larger-than-L2 networks get the fraction that fits.
// Measured (profiles/README.md, round 2): +1 % on the 512-member ensemble, but the carve-out is taken from
// the L2 everything else uses -- with the 250 MB arena of the 1M-link model it doubled the kernel time.
// Off by default; networks up to SWB_PERSIST_MAX_BYTES may opt in at build time.
#ifndef SWB_PERSIST_NET
#define SWB_PERSIST_NET 0
#endif
#define SWB_PERSIST_MAX_BYTES ((size_t)16 << 20)
static void persist_window(int device, cudaStream_t stream, const void *base, size_t bytes)
{
#if SWB_PERSIST_NET
    DevInfo &D = g_dev[device];
    if (!base || !bytes || D.persist_max == 0 || bytes > SWB_PERSIST_MAX_BYTES) return;
    // the persisting carve-out is taken away from everything else in L2 (register spills, the state
    // streams): reserve what the network needs, never the device maximum
    const size_t want = std::min(D.persist_max, (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1));
    if (want > D.persist_set) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        D.persist_set = want;
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = bytes <= D.persist_set ? 1.0f : (float)((double)D.persist_set / (double)bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
#endif
}

// receive window of a partitioned solver: plain device memory, exported to the peer processes as a
// CUDA IPC handle and mapped by them with peer access over NVLink
static_assert(sizeof(cudaIpcMemHandle_t) <= SWB_WINDOW_HANDLE_BYTES, "handle size");
static void *window_alloc(size_t bytes, void *handle, std::string &err)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { err = cuda_err("cudaMalloc(window)", e); return nullptr; }
    cudaMemset(p, 0, bytes);
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { err = cuda_err("cudaIpcGetMemHandle", e); cudaFree(p); return nullptr; }
    memset(handle, 0, SWB_WINDOW_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return p;
}
static void *window_open(const void *handle, size_t, std::string &err)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { err = cuda_err("cudaIpcOpenMemHandle", e); return nullptr; }
    return p;
}
static void window_close(void *p) { if (p) cudaIpcCloseMemHandle(p); }
static void window_free(void *p, size_t, const void *) { if (p) cudaFree(p); }

// dst[r][i] = src[r][perm[i]] for a batch of rows
template <class T>
__global__ void swb_permute_rows(T *dst, const T *src, const int *perm, int M, size_t rows)
{
    const size_t n = rows * (size_t)M;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const size_t r = t / M;
        const int i = (int)(t - r * M);
        dst[t] = src[r * M + perm[i]];
    }
}
static bool permute_members(void *p, size_t rows, int esz, int M, const int *d_perm, void *tmp, size_t tmp_bytes)
{
    const size_t row_bytes = (size_t)M * esz;
    const size_t batch = std::max<size_t>(1, tmp_bytes / row_bytes);
    for (size_t r0 = 0; r0 < rows; r0 += batch) {
        const size_t nr = std::min(batch, rows - r0);
        char *base = (char *)p + r0 * row_bytes;
        const int blocks = (int)std::min<size_t>(148 * 8, (nr * M + 255) / 256);
        if (esz == 8) swb_permute_rows<unsigned long long><<<blocks, 256>>>((unsigned long long *)tmp, (const unsigned long long *)base, d_perm, M, nr);
        else if (esz == 4) swb_permute_rows<unsigned int><<<blocks, 256>>>((unsigned int *)tmp, (const unsigned int *)base, d_perm, M, nr);
        else swb_permute_rows<unsigned char><<<blocks, 256>>>((unsigned char *)tmp, (const unsigned char *)base, d_perm, M, nr);
        if (!ok("swb_permute_rows", cudaGetLastError())) return false;
        if (!ok("cudaMemcpyAsync(D2D)", cudaMemcpyAsync(base, tmp, nr * row_bytes, cudaMemcpyDeviceToDevice, 0))) return false;
    }
    return true;
}

static void *host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    return p;
}
static void host_free(void *p) { if (p) cudaFreeHost(p); }
static void h2d_async(void *d, const void *s, size_t b) { if (b) cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, g_stream); }
static void d2h_async(void *d, const void *s, size_t b) { if (b) cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, g_stream); }
static bool sync(std::string &err)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { err = cuda_err("cudaDeviceSynchronize", e); return false; }
    return true;
}

static void profiler(bool on) { if (on) cudaProfilerStart(); else { cudaDeviceSynchronize(); cudaProfilerStop(); } }

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

static bool report(int device, const Net &net, const State &st, const double *f, int m0, int nm, float *node_out,
                   float *link_out, std::string &err)
{
    for (int links = 0; links < 2; links++) {
        float *out = links ? link_out : node_out;
        if (!out) continue;
        long long total = (long long)(links ? net.nL : net.nN) * nm;
        int blocks = (int)std::min<long long>((total + 255) / 256, (long long)g_dev[device].sms * 8);
        swb_report_kernel<<<std::max(blocks, 1), 256, 0, g_stream>>>(net, st, f, m0, nm, links, out);
    }
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) { err = cuda_err("swb_report_kernel", e); return false; }
    return true;
}

// CTAs to launch: all that can be co-resident, but (i) threads % M == 0 so a thread keeps one
// member, (ii) no more threads than (objects x members), (iii) at least one CTA per 256 members.
static int pick_blocks(const DevInfo &D, int M, int maxItems)
{
    long long maxBlocks = (long long)D.sms * D.blocks_per_sm;
    long long work = ((long long)maxItems * M + SWB_BLOCK - 1) / SWB_BLOCK;
    long long blocks = std::max(1LL, std::min(maxBlocks, work));
    int g = M, b = SWB_BLOCK;                 // gcd(M, SWB_BLOCK)
    while (b) { int t = g % b; g = b; b = t; }
    long long unit = M / g;                   // blocks must be a multiple of this
    blocks = (blocks / unit) * unit;
    if (blocks < unit) blocks = unit;
    return (int)blocks;
}

static bool launch_persistent(const Net &net, const State &st, const RunArgs &args, int device, float *ms, std::string &err,
                              bool wait = true)
{
    if (!init(device, err)) return false;
    const DevInfo &D = g_dev[device];
    int blocks = pick_blocks(D, st.M, std::max(net.nN, net.nL));
    if ((long long)blocks > (long long)D.sms * D.blocks_per_sm) {
        err = "ensemble too wide for one cooperative launch on this device";
        return false;
    }
    void *kargs[] = { (void *)&net, (void *)&st, (void *)&args };
    size_t dyn = sizeof(int) * (size_t)((st.M + 31) / 32 * 32);
    persist_window(device, g_stream, net.arena, net.arena_bytes);
    if (!wait) {                          // left in flight on the current stream (swb_step_host_batch)
        cudaError_t e = cudaLaunchCooperativeKernel((void *)swb_route_kernel, dim3(blocks), dim3(SWB_BLOCK), kargs,
                                                    dyn, g_stream);
        if (e != cudaSuccess) { err = cuda_err("cudaLaunchCooperativeKernel", e); return false; }
        *ms = 0.f;
        return true;
    }
    cudaError_t e = cudaEventRecord(D.ev0, g_stream);
    if (e != cudaSuccess) { err = cuda_err("cudaEventRecord", e); return false; }
    e = cudaLaunchCooperativeKernel((void *)swb_route_kernel, dim3(blocks), dim3(SWB_BLOCK), kargs, dyn,
                                                g_stream);
    if (e != cudaSuccess) { err = cuda_err("cudaLaunchCooperativeKernel", e); return false; }
    e = cudaEventRecord(D.ev1, g_stream);
    if (e != cudaSuccess) { err = cuda_err("cudaEventRecord", e); return false; }
    e = cudaEventSynchronize(D.ev1);
    if (e != cudaSuccess) { err = cuda_err("swb_route_kernel", e); return false; }
    cudaEventElapsedTime(ms, D.ev0, D.ev1);
    return true;
}

} }

#include "swb_staged.cuh"
#include "swb_api_impl.h"
