// CPU emulation of the device engine -- TEST SCAFFOLDING, never part of libswmm_b200.so.
// Instantiates csrc/swb_api_impl.h with a host-memory backend whose "launch" runs the identical
// engine_run<> template on G host threads with a real barrier for every grid-wide sync, so the
// CPU-only suite exercises the kernels' arithmetic AND their control flow (barrier placement,
// per-member convergence bookkeeping) against the reference.
#include <barrier>
#include <cstdio>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "swb_state.h"
#include "swb_engine.h"
#include "swb_report.h"

namespace swb { namespace backend {
static bool init(int, std::string &) { return true; }
static void *alloc(size_t b) { return std::malloc(b ? b : 8); }
static void free_(void *p) { std::free(p); }
static bool upload(void *d, const void *s, size_t b) { std::memcpy(d, s, b); return true; }
static bool download(void *d, const void *s, size_t b) { std::memcpy(d, s, b); return true; }
static bool zero(void *d, size_t b) { std::memset(d, 0, b); return true; }
static bool copy2d(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, bool)
{
    for (size_t r = 0; r < height; r++) std::memcpy((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return true;
}
static const char *last_error() { return "host allocation failed"; }
static bool sync(std::string &) { return true; }
static void profiler(bool) {}
static void *stream_create() { return nullptr; }
static void stream_destroy(void *) {}
static void use_stream(void *) {}
static int device_count() { return 1; }
static bool permute_members(void *p, size_t rows, int esz, int M, const int *perm, void *, size_t)
{
    std::vector<char> row((size_t)M * esz);
    for (size_t r = 0; r < rows; r++) {
        char *base = (char *)p + r * (size_t)M * esz;
        for (int i = 0; i < M; i++) memcpy(row.data() + (size_t)i * esz, base + (size_t)perm[i] * esz, esz);
        memcpy(base, row.data(), row.size());
    }
    return true;
}
static void *host_alloc(size_t b) { return std::malloc(b ? b : 8); }
static void host_free(void *p) { std::free(p); }
static void h2d_async(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
static void d2h_async(void *d, const void *s, size_t b) { std::memcpy(d, s, b); }
// receive window of a partitioned solver: POSIX shared memory, so that the peers of the CPU tests can
// be threads of one process or separate processes (torch.distributed / gloo, world_size 2)
struct ShmHandle { char name[48]; unsigned long long bytes; };
static void *window_alloc(size_t bytes, void *handle, std::string &err)
{
    static int counter = 0;
    ShmHandle h;
    std::memset(&h, 0, sizeof(h));
    std::snprintf(h.name, sizeof(h.name), "/swb_%d_%d", (int)getpid(), __atomic_fetch_add(&counter, 1, __ATOMIC_RELAXED));
    h.bytes = bytes;
    int fd = shm_open(h.name, O_CREAT | O_RDWR | O_EXCL, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { err = "shm_open failed"; return nullptr; }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { err = "mmap failed"; return nullptr; }
    std::memset(p, 0, bytes);
    static_assert(sizeof(ShmHandle) <= SWB_WINDOW_HANDLE_BYTES, "handle size");
    std::memset(handle, 0, SWB_WINDOW_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    return p;
}
static void *window_open(const void *handle, size_t, std::string &err)
{
    ShmHandle h;
    std::memcpy(&h, handle, sizeof(h));
    int fd = shm_open(h.name, O_RDWR, 0600);
    if (fd < 0) { err = "shm_open(peer) failed"; return nullptr; }
    void *p = mmap(nullptr, h.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { err = "mmap(peer) failed"; return nullptr; }
    return p;
}
static void window_close(void *) {}          // mappings live until the test process exits
static void window_free(void *p, size_t bytes, const void *handle)
{
    ShmHandle h;
    std::memcpy(&h, handle, sizeof(h));
    shm_unlink(h.name);
    if (p) munmap(p, bytes);
}

static bool xsect_eval(int, int fn, const Xs &x, int n, const double *args, double *out, std::string &)
{
    static const double tab[] = { SWB_XS_TABLE_DATA };
    for (int i = 0; i < n; i++) out[i] = xs_eval(fn, x, args[i], tab);
    return true;
}

static bool report(int, const Net &net, const State &st, const double *f, int m0, int nm, float *node_out,
                   float *link_out, std::string &)
{
    const int nr = node_record_len(net), lr = link_record_len(net);
    for (int mm = 0; mm < nm; mm++) {
        if (node_out)
            for (int i = 0; i < net.nN; i++)
                node_results(net, st, i, m0 + mm, f[m0 + mm], node_out + ((size_t)mm * net.nN + i) * nr);
        if (link_out)
            for (int j = 0; j < net.nL; j++)
                link_results(net, st, j, m0 + mm, f[m0 + mm], link_out + ((size_t)mm * net.nL + j) * lr, net.xs_tables);
    }
    return true;
}

struct EmulCtx {
    int tid, G, lane, block_size;
    const double *T;
    std::barrier<> *bar;
    int *alive_list;
    std::vector<int> own_list;
    int warp_size = 1, warp_lane = 0;
    unsigned long long next_ticket(unsigned long long *p) { return __atomic_fetch_add(p, 1ull, __ATOMIC_RELAXED); }
    unsigned long long ticket_issue(unsigned long long *p) { return next_ticket(p); }
    unsigned long long ticket_take(unsigned long long pending) { return pending; }
    void prefetch(const void *) {}
    template <class Pred> int compact_members(int M, Pred alive)
    {
        own_list.resize(M);
        alive_list = own_list.data();
        int n = 0;
        for (int mm = 0; mm < M; mm++) if (alive(mm)) alive_list[n++] = mm;
        return n;
    }
    void grid_sync() { bar->arrive_and_wait(); }
    void load_tables(const double *) {}
    // same contract as CudaCtx::transpose, elements dealt round-robin to the G host threads
    void transpose(double *dst, const double *src, int R, int C, int planes)
    {
        const int items = planes > 0 ? C / planes : 0;
        for (size_t e = tid; e < (size_t)R * C; e += G) {
            int r = (int)(e / C), c = (int)(e % C), cc = c;
            if (planes > 0) { int item = c / planes; cc = (c - item * planes) * items + item; }
            dst[(size_t)cc * R + r] = src[e];
        }
    }
    void block_sync() {}                  // a host "block" is one thread
    void fence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
    void store_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
    unsigned long long load_acquire_sys(const unsigned long long *p)
    { std::this_thread::yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }      // only used in spin loops
    double load_sys_f64(const double *p) { return *(const volatile double *)p; }
    unsigned long long now_ns()
    { return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
          std::chrono::steady_clock::now().time_since_epoch()).count(); }
    bool block_or(bool b) { return b; }
    void atomic_min_u64(unsigned long long *p, unsigned long long v)
    {
        unsigned long long cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
    unsigned long long warp_min_u64(unsigned long long v) { return v; }
    double warp_sum_f64(double v) { return v; }
    void atomic_min_i32(int *p, int v)
    {
        int cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
    void atomic_add_f64(double *p, double v)
    {
        unsigned long long *u = (unsigned long long *)p, cur = __atomic_load_n(u, __ATOMIC_RELAXED), nxt;
        do { double d; std::memcpy(&d, &cur, 8); d += v; std::memcpy(&nxt, &d, 8); }
        while (!__atomic_compare_exchange_n(u, &cur, nxt, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    }
};

static int staged_max_threads(int) { return 0; }
static int set_staged_min_members(int) { return 0; }    // the staged kernels are CUDA only
static bool launch(const Net &net, const State &st, const RunArgs &args, int, float *ms, std::string &, bool = true,
                   int *n_kernels = nullptr)
{
    if (n_kernels) *n_kernels = 1;
    const int M = st.M;
    int k = 1;
    if (M < 4) k = 4 / M;
    const int G = M * k;
    auto t0 = std::chrono::steady_clock::now();
    std::barrier<> bar(G);
    std::vector<std::thread> th;
    for (int t = 0; t < G; t++)
        th.emplace_back([&, t]() {
            EmulCtx c{t, G, 0, 1, net.xs_tables, &bar, nullptr, {}, 1, 0};
            engine_run(net, st, args, c);
        });
    for (auto &x : th) x.join();
    *ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
}
} }

#include "swb_api_impl.h"
