// swb_api_impl.h -- the C-ABI of include/swmm_b200.h, written once against a tiny memory/launch
// backend.  csrc/swb_api.cu instantiates it with the CUDA backend (the product); tests/emul
// instantiates it with a host-thread backend so the CPU-only test suite can drive the very same
// entry points.  The including file must define, before including this header:
//
//   namespace swb { namespace backend {
//     bool  init(int device, std::string &err);
//     void *alloc(size_t bytes);            void free_(void *p);
//     void  upload(void *dst, const void *src, size_t bytes);
//     void  download(void *dst, const void *src, size_t bytes);
//     void  zero(void *dst, size_t bytes);
//     bool  launch(const Net &, const State &, const RunArgs &, int device, float *ms, std::string &err);
//     bool  sync(std::string &err);         int device_count();
//     void *host_alloc(size_t);  void host_free(void *);
//     bool permute_members(void *p, size_t rows, int esz, int M, const int *d_perm, void *tmp, size_t tmp_bytes);
//     void  h2d_async(void *dst, const void *src, size_t bytes);  void d2h_async(...);
//     bool  xsect_eval(int device, int fn, const Xs &x, int n, const double *args, double *out, std::string &err);
//   } }
#ifndef SWB_API_IMPL_H
#define SWB_API_IMPL_H

#include <cstddef>
#include <cstring>
#include <string>
#include <vector>
#include "swb_host.h"
#include "swb_engine.h"
#include "swb_report.h"

using namespace swb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

struct swb_network {
    Net net;
    Derived der;
    int device;
    std::vector<void *> allocs;
};

struct swb_solver {
    swb_network *net;
    State st;
    int M;
    std::vector<void *> allocs;
    std::vector<void *> inflow_allocs;      // device arrays of the current swb_set_inflows call
    Inflows inflows;
    bool have_inflows;
    void *xfer_buf = nullptr; size_t xfer_cap = 0;   // pinned staging of swb_set_field / swb_get_field
    void *perm_tmp = nullptr; int *perm_idx = nullptr;   // scratch of swb_permute_members
    std::vector<void *> control_allocs;     // device arrays of the current swb_set_controls call
    Controls controls;
    long long launches;
    float last_ms;
    std::vector<double> h_dt;
    // swb_step_host staging (device): host-layout landing zones and device-layout images
    double *stg_lat, *stg_loss, *stg_qual, *img_lat, *img_loss, *img_qual, *stg_depth, *stg_flow;
    void *stream;               // own stream for swb_step_host_batch (created on first use)
    // partitioned network: own receive window, the peers' mapped windows
    void *window; size_t window_bytes;
    unsigned char window_handle[SWB_WINDOW_HANDLE_BYTES];
    void *peer_window[SWB_MAX_RANKS];
};

static size_t window_size(int nRecv, int W)
{ return sizeof(unsigned long long) * (HALO_CTRL_WORDS + 2 * SWB_MAX_RANKS * HALO_RED)
         + sizeof(double) * 2 * (size_t)(nRecv > 0 ? nRecv : 1) * W; }
static void window_views(void *base, unsigned long long *&ctrl, unsigned long long *&red, double *&stage)
{
    ctrl = (unsigned long long *)base;
    red = ctrl + HALO_CTRL_WORDS;
    stage = (double *)(red + 2 * SWB_MAX_RANKS * HALO_RED);
}

// allocation failures surface as SWB_ERR_CUDA from the creating entry point (never as a null device
// pointer inside Net / State): the helpers throw, the entry points catch, free and report
struct DeviceError { std::string what; };
static void *checked_alloc(size_t bytes)
{
    void *p = backend::alloc(bytes ? bytes : 8);
    if (!p) throw DeviceError{"device allocation of " + std::to_string(bytes) + " bytes failed: " + backend::last_error()};
    return p;
}
static void checked(bool ok, const char *what) { if (!ok) throw DeviceError{std::string(what) + ": " + backend::last_error()}; }
template <class T>
static T *dev_copy(std::vector<void *> &allocs, const T *src, size_t n)
{
    T *p = (T *)checked_alloc(sizeof(T) * (n ? n : 1));
    allocs.push_back(p);
    if (n) checked(backend::upload(p, src, sizeof(T) * n), "host -> device copy");
    return p;
}
template <class T>
static T *dev_zero(std::vector<void *> &allocs, size_t n)
{
    T *p = (T *)checked_alloc(sizeof(T) * (n ? n : 1));
    allocs.push_back(p);
    checked(backend::zero(p, sizeof(T) * (n ? n : 1)), "device memset");
    return p;
}
// static network arrays: packed (256-byte aligned) into one host image, uploaded into one allocation
struct Arena {
    std::vector<char> image;
    std::vector<std::pair<const void **, size_t>> slots;
    template <class T> void add(const T *&slot, const T *src, size_t n)
    {
        size_t off = (image.size() + 255) & ~(size_t)255;
        image.resize(off + sizeof(T) * (n ? n : 1), 0);
        if (n) memcpy(image.data() + off, src, sizeof(T) * n);
        slots.push_back({(const void **)&slot, off});
    }
    void *commit(std::vector<void *> &allocs)
    {
        char *base = (char *)checked_alloc(image.size());
        allocs.push_back(base);
        checked(backend::upload(base, image.data(), image.size()), "host -> device copy");
        for (auto &sl : slots) *sl.first = base + sl.second;
        return base;
    }
};

// every solver entry point that allocates, copies or launches makes the solver's device current first
// (a process may hold networks on several devices)
static int enter(const swb_network *nw)
{
    std::string err;
    if (!backend::init(nw->device, err)) return fail(SWB_ERR_CUDA, err);
    return SWB_OK;
}
static int enter(const swb_solver *s) { return enter(s->net); }
#define SWB_ENTER(obj) do { int _rc = enter(obj); if (_rc) return _rc; } while (0)

extern "C" {

const char *swb_last_error(void) { return g_err.c_str(); }
int swb_version(void) { return SWB_VERSION; }
int swb_device_count(void) { return backend::device_count(); }

void swb_network_destroy(swb_network *nw);
int swb_network_create(const swb_network_desc *d, const swb_options *o, int device, swb_network **out)
{
    if (!d || !o || !out) return fail(SWB_ERR_ARG, "null argument");
    std::string why = validate_desc(*d, *o);
    if (!why.empty())
        return fail(why.find("not supported") != std::string::npos || why.find("dummy") != std::string::npos
                        ? SWB_ERR_UNSUPP : SWB_ERR_ARG, why);
    std::string err;
    if (!backend::init(device, err)) return fail(SWB_ERR_CUDA, err);
    swb_network *nw = new swb_network();
    nw->device = device;
    derive(*d, *o, nw->der);
    fill_net_scalars(nw->net, *d, *o, nw->der);
    try {
        Arena ar;
#define X(T, name, kind) ar.add<T>(nw->net.name, d->name, desc_count(*d, #kind[0], #kind[1]));
        SWB_DESC_ARRAYS(X)
#undef X
        const Derived &r = nw->der;
#define D(name) ar.add(nw->net.name, r.name.data(), r.name.size());
        D(outfall_nodes) D(adj_packed) D(link_flags) D(link_z1) D(link_z2) D(xs_rcp_yfull) D(cond_rcp_mod_length) D(adj_start) D(adj)
        D(adjq_start) D(adjq) D(nc_links) D(node_order) D(link_order) D(outfall_link) D(xs_tables)
        D(link_kernel) D(culvert_params) D(road_tables) D(link_rows) D(link_cols_d) D(link_cols_i) D(outfall_slot) D(link_pre_node) D(pre_links)
#undef D
        nw->net.arena_bytes = ar.image.size();
        nw->net.arena = ar.commit(nw->allocs);
    } catch (const DeviceError &e) {
        swb_network_destroy(nw);
        return fail(SWB_ERR_CUDA, e.what);
    }
    *out = nw;
    return SWB_OK;
}

void swb_network_destroy(swb_network *nw)
{
    if (!nw) return;
    for (void *p : nw->allocs) backend::free_(p);
    delete nw;
}

void swb_solver_destroy(swb_solver *s);
int swb_solver_create(swb_network *nw, int M, swb_solver **out)
{
    if (!nw || !out || M < 1) return fail(SWB_ERR_ARG, "bad solver arguments");
    if (M > 1 && (M % 32) != 0) return fail(SWB_ERR_ARG, "n_members must be 1 or a multiple of 32");
    if (M > 8192) return fail(SWB_ERR_ARG, "n_members > 8192: split the ensemble over several solvers");
    SWB_ENTER(nw);
    swb_solver *s = new swb_solver();
    s->net = nw; s->M = M; s->launches = 0; s->last_ms = 0.f; s->have_inflows = false; s->stream = nullptr;
    s->stg_lat = s->stg_loss = s->stg_qual = s->img_lat = s->img_loss = s->img_qual = nullptr;
    s->stg_depth = s->stg_flow = nullptr;
    s->window = nullptr; s->window_bytes = 0;
    memset(s->peer_window, 0, sizeof(s->peer_window));
    memset(&s->inflows, 0, sizeof(s->inflows));
    memset(&s->controls, 0, sizeof(s->controls));
    State &st = s->st;
    memset(&st, 0, sizeof(st));
    st.M = M;
    const int nN = nw->net.nN, nL = nw->net.nL, nP = nw->net.nP;
    try {
    for (const FieldInfo &f : field_table()) {
        size_t n = field_items(f, nN, nL, nP) * (size_t)M;
        void *p = f.is_u8 ? (void *)dev_zero<unsigned char>(s->allocs, n) : (void *)dev_zero<double>(s->allocs, n);
        *(void **)((char *)&st + f.offset) = p;
    }
    st.dt = dev_zero<double>(s->allocs, M);
    st.var_step = dev_zero<double>(s->allocs, M);
    st.o_ynorm = dev_zero<double>(s->allocs, (size_t)(nw->net.nOutfallNodes ? nw->net.nOutfallNodes : 1) * M);
    st.o_ycrit = dev_zero<double>(s->allocs, (size_t)(nw->net.nOutfallNodes ? nw->net.nOutfallNodes : 1) * M);
    st.sim_time = dev_zero<double>(s->allocs, M);
    st.time_ms = dev_zero<double>(s->allocs, M);
    std::vector<double> ev(M, nw->net.opt.evap_rate), hc(M, nw->net.opt.hydcon_factor);
    st.evap_rate = dev_copy<double>(s->allocs, ev.data(), M);
    st.hydcon = dev_copy<double>(s->allocs, hc.data(), M);
    st.iters = dev_zero<int>(s->allocs, M);
    st.tot_iters = dev_zero<long long>(s->allocs, M);
    st.tot_steps = dev_zero<long long>(s->allocs, M);
    st.non_conv = dev_zero<long long>(s->allocs, M);
    st.crit_node = dev_zero<int>(s->allocs, M);
    st.crit_link = dev_zero<int>(s->allocs, M);
    st.tmin_bits = dev_zero<unsigned long long>(s->allocs, 2 * (size_t)M);
    st.alive = dev_zero<int>(s->allocs, (size_t)M);
    st.ctl = dev_zero<int>(s->allocs, SWB_CTL_WORDS);
    st.dt_cand = nullptr; st.dt_cand_stride = 0;
    if (M >= 32 && backend::staged_max_threads(nw->device) > 0) {   // staged kernels: per-thread Courant candidates
        st.dt_cand_stride = backend::staged_max_threads(nw->device);
        st.dt_cand = dev_zero<double>(s->allocs, 4 * (size_t)st.dt_cand_stride);
    }
    st.not_conv = dev_zero<int>(s->allocs, (size_t)SWB_MAX_TRIALS_CAP * M);
    st.done = dev_zero<int>(s->allocs, M);
    st.mb_reacted = dev_zero<double>(s->allocs, (size_t)(nP ? nP : 1) * M);
    st.mb_seepage = dev_zero<double>(s->allocs, (size_t)(nP ? nP : 1) * M);
    st.mb_final_storage = dev_zero<double>(s->allocs, (size_t)(nP ? nP : 1) * M);
    st.mb_rate = dev_zero<double>(s->allocs, (size_t)(MB_FLOW_TERMS + MB_QUAL_TERMS * nP) * M);
    st.mb_total = dev_zero<double>(s->allocs, (size_t)(MB_FLOW_TERMS + MB_QUAL_TERMS * nP) * M);
    st.mb_dt_prev = dev_zero<double>(s->allocs, M);
    st.phase_ns = dev_zero<unsigned long long>(s->allocs, SWB_N_PHASES);
    st.tickets = dev_zero<unsigned long long>(s->allocs, SWB_TICKETS_PER_TRIAL * SWB_MAX_TRIALS_CAP);
    // conduit / link settings default to fully open (Link.setting = 1.0, link.c:142)
    std::vector<double> ones((size_t)nL * M, 1.0);
    checked(backend::upload(st.l_setting, ones.data(), sizeof(double) * ones.size()), "host -> device copy");
    checked(backend::upload(st.l_target_setting, ones.data(), sizeof(double) * ones.size()), "host -> device copy");
    } catch (const DeviceError &e) {
        swb_solver_destroy(s);
        return fail(SWB_ERR_CUDA, e.what);
    }
    *out = s;
    return SWB_OK;
}

void swb_solver_destroy(swb_solver *s)
{
    if (!s) return;
    for (void *p : s->allocs) backend::free_(p);
    for (void *p : s->inflow_allocs) backend::free_(p);
    for (void *p : s->control_allocs) backend::free_(p);
    if (s->xfer_buf) backend::host_free(s->xfer_buf);
    if (s->stream) backend::stream_destroy(s->stream);
    for (int p = 0; p < SWB_MAX_RANKS; p++) if (s->peer_window[p]) backend::window_close(s->peer_window[p]);
    if (s->window) backend::window_free(s->window, s->window_bytes, s->window_handle);
    delete s;
}
int swb_solver_members(const swb_solver *s) { return s ? s->M : 0; }

// host [member][item(,p)]  <->  device [(p,) item][member]
static int field_xfer(swb_solver *s, int field, int m0, int nm, double *buf, const double *cbuf, bool set,
                      bool broadcast)
{
    if (!s || (!buf && !cbuf)) return fail(SWB_ERR_ARG, "null argument");
    SWB_ENTER(s);
    const FieldInfo *f = find_field(field);
    if (!f) return fail(SWB_ERR_ARG, "unknown field id");
    const int M = s->M, nN = s->net->net.nN, nL = s->net->net.nL, nP = s->net->net.nP;
    if (broadcast) { m0 = 0; nm = M; }
    if (m0 < 0 || nm < 1 || m0 + nm > M) return fail(SWB_ERR_ARG, "member range out of bounds");
    const size_t items = field_items(*f, nN, nL, nP);
    const size_t base = (f->kind == 'n') ? nN : (f->kind == 'l') ? nL : items;   // objects per plane
    const int planes = (int)(items / (base ? base : 1));
    void *dev = *(void **)((char *)&s->st + f->offset);
    // only the member columns [m0, m0 + nm) of every row travel (strided 2-D copy): reading one
    // member of a wide ensemble does not download the whole field
    const size_t esz = f->is_u8 ? 1 : sizeof(double);
    const size_t cols = (size_t)nm;
    // one pinned staging buffer per solver (doubles, then the byte image of a flag field): the copies run at
    // full PCIe rate without the driver's bounce buffer and nothing is allocated per call -- the drop-in seam
    // moves ~100 fields per routing step through here
    const size_t need = items * cols * (sizeof(double) + 1);
    if (need > s->xfer_cap) {
        if (s->xfer_buf) backend::host_free(s->xfer_buf);
        s->xfer_cap = 0;
        s->xfer_buf = backend::host_alloc(need);
        if (!s->xfer_buf) return fail(SWB_ERR_CUDA, "pinned staging allocation failed");
        s->xfer_cap = need;
    }
    struct Span { double *p; size_t n; double *data() const { return p; } size_t size() const { return n; }
                  double &operator[](size_t i) const { return p[i]; } };
    struct Span8 { unsigned char *p; size_t n; unsigned char *data() const { return p; } size_t size() const { return n; }
                   unsigned char &operator[](size_t i) const { return p[i]; } };
    const Span h = { (double *)s->xfer_buf, items * cols };
    const Span8 h8 = { (unsigned char *)s->xfer_buf + items * cols * sizeof(double), f->is_u8 ? items * cols : 0 };
    void *hostp = h.data();
    if (f->is_u8) hostp = h8.data();
    char *devp = (char *)dev + (size_t)m0 * esz;
    if (!set) {
        if (!backend::copy2d(hostp, cols * esz, devp, (size_t)M * esz, cols * esz, items, false))
            return fail(SWB_ERR_CUDA, backend::last_error());
        if (f->is_u8) for (size_t i = 0; i < h8.size(); i++) h[i] = h8[i];
        for (size_t it = 0; it < base; it++)
            for (int p = 0; p < planes; p++) {
                const double *row = h.data() + ((size_t)p * base + it) * cols;
                double *dst = buf + it * planes + p;
                for (int mm = 0; mm < nm; mm++) dst[(size_t)mm * items] = row[mm];
            }
        return SWB_OK;
    }
    for (size_t it = 0; it < base; it++)
        for (int p = 0; p < planes; p++) {
            double *row = h.data() + ((size_t)p * base + it) * cols;
            const double *src = cbuf + it * planes + p;
            if (broadcast) { const double v = *src; for (int mm = 0; mm < nm; mm++) row[mm] = v; }
            else for (int mm = 0; mm < nm; mm++) row[mm] = src[(size_t)mm * items];
        }
    if (f->is_u8) for (size_t i = 0; i < h8.size(); i++) h8[i] = (unsigned char)h[i];
    if (!backend::copy2d(devp, (size_t)M * esz, hostp, cols * esz, cols * esz, items, true))
        return fail(SWB_ERR_CUDA, backend::last_error());
    return SWB_OK;
}

int swb_set_field(swb_solver *s, int field, int m0, int nm, const double *buf)
{ return field_xfer(s, field, m0, nm, nullptr, buf, true, false); }
int swb_get_field(swb_solver *s, int field, int m0, int nm, double *buf)
{ return field_xfer(s, field, m0, nm, buf, nullptr, false, false); }
int swb_broadcast_field(swb_solver *s, int field, const double *buf)
{ return field_xfer(s, field, 0, 0, nullptr, buf, true, true); }

int swb_set_climate(swb_solver *s, double evap_rate, double hydcon_factor)
{
    if (!s) return fail(SWB_ERR_ARG, "null solver");
    SWB_ENTER(s);
    std::vector<double> ev(s->M, evap_rate), hc(s->M, hydcon_factor);
    backend::upload(s->st.evap_rate, ev.data(), sizeof(double) * s->M);
    backend::upload(s->st.hydcon, hc.data(), sizeof(double) * s->M);
    return SWB_OK;
}

int swb_qual_init(swb_solver *s, const double *init_concen)
{
    if (!s) return fail(SWB_ERR_ARG, "null solver");
    SWB_ENTER(s);
    const int M = s->M, nN = s->net->net.nN, nL = s->net->net.nL, nP = s->net->net.nP;
    if (nP == 0) return SWB_OK;
    std::vector<double> nd((size_t)nN * M), ld((size_t)nL * M);
    backend::download(nd.data(), s->st.n_depth, sizeof(double) * nd.size());
    backend::download(ld.data(), s->st.l_depth, sizeof(double) * ld.size());
    std::vector<double> nq((size_t)nP * nN * M), lq((size_t)nP * nL * M);
    for (int p = 0; p < nP; p++) {
        double c0 = init_concen ? init_concen[p] : 0.0;
        for (size_t i = 0; i < nd.size(); i++) nq[(size_t)p * nN * M + i] = nd[i] > SWB_ZERO_DEPTH ? c0 : 0.0;
        for (size_t i = 0; i < ld.size(); i++) lq[(size_t)p * nL * M + i] = ld[i] > SWB_ZERO_DEPTH ? c0 : 0.0;
    }
    backend::upload(s->st.n_qual, nq.data(), sizeof(double) * nq.size());
    backend::upload(s->st.n_old_qual, nq.data(), sizeof(double) * nq.size());
    backend::upload(s->st.l_qual, lq.data(), sizeof(double) * lq.size());
    backend::upload(s->st.l_old_qual, lq.data(), sizeof(double) * lq.size());
    return SWB_OK;
}

static int run(swb_solver *s, int phases, int n_steps, double t_end, double fixed_step,
               const double *host_lat = nullptr, const double *host_losses = nullptr,
               const double *host_qual = nullptr, bool wait = true, const RunArgs *stg = nullptr)
{
    RunArgs a;
    memset(&a, 0, sizeof(a));
    if (stg) a = *stg;                       // host-layout staging pointers (swb_step_host)
    a.host_lat = host_lat; a.host_losses = host_losses; a.host_qual = host_qual;
    a.phases = phases; a.n_steps = n_steps; a.t_end = t_end; a.fixed_step = fixed_step;
    a.inflows = s->inflows;
    if (phases & PH_ADVANCE) a.controls = s->controls;      // (zeroed otherwise: inactive)
    std::string err;
    float ms = 0.f;
    if (s->st.halo.nRanks > 1) {
        for (int p = 0; p < s->st.halo.nRanks; p++)
            if (p != s->st.halo.rank && !s->peer_window[p])
                return fail(SWB_ERR_ARG, "partitioned solver: swb_partition_connect has not been called for every peer");
        if (!wait || (phases & PH_HOSTIN)) return fail(SWB_ERR_UNSUPP, "partitioned solver: use swb_run_steps");
    }
    int kernels = 1;
    if (!backend::launch(s->net->net, s->st, a, s->net->device, &ms, err, wait, &kernels)) return fail(SWB_ERR_CUDA, err);
    s->launches += kernels;
    s->last_ms = ms;
    if (s->st.halo.nRanks > 1) {
        unsigned long long flag = 0;
        backend::download(&flag, s->st.halo.ctrl + HALO_ERR, sizeof(flag));
        if (flag) return fail(SWB_ERR_CUDA, "halo exchange timed out: a peer rank did not answer");
    }
    return SWB_OK;
}

static int put_dt(swb_solver *s, const double *dt)
{
    if (!s || !dt) return fail(SWB_ERR_ARG, "null argument");
    SWB_ENTER(s);
    backend::upload(s->st.dt, dt, sizeof(double) * s->M);
    return SWB_OK;
}

int swb_old_state_swap(swb_solver *s, const double *dt, int with_quality)
{
    int rc = put_dt(s, dt);
    if (rc) return rc;
    return run(s, PH_SWAP | (with_quality ? PH_QSWAP : 0), 1, 0.0, s->net->net.opt.route_step);
}

int swb_dynwave_execute(swb_solver *s, const double *dt, int *iters)
{
    int rc = put_dt(s, dt);
    if (rc) return rc;
    rc = run(s, PH_DYNWAVE, 1, 0.0, s->net->net.opt.route_step);
    if (rc) return rc;
    if (iters) backend::download(iters, s->st.iters, sizeof(int) * s->M);
    return SWB_OK;
}

int swb_qualrout_execute(swb_solver *s, const double *dt)
{
    int rc = put_dt(s, dt);
    if (rc) return rc;
    return run(s, PH_QUALITY, 1, 0.0, s->net->net.opt.route_step);
}

int swb_get_routing_step(swb_solver *s, double fixed_step, double *dt_out)
{
    if (!s || !dt_out) return fail(SWB_ERR_ARG, "null argument");
    SWB_ENTER(s);
    int rc = run(s, PH_NEXTDT, 1, 0.0, fixed_step);
    if (rc) return rc;
    backend::download(dt_out, s->st.var_step, sizeof(double) * s->M);
    return SWB_OK;
}

void *swb_host_alloc(unsigned long long bytes) { return backend::host_alloc((size_t)bytes); }
void swb_host_free(void *p) { backend::host_free(p); }

static void ensure_staging(swb_solver *s)
{
    const Net &n = s->net->net;
    const int M = s->M, nN = n.nN, nL = n.nL, nP = n.nP;
    if (!s->stg_lat) {
        s->stg_lat = (double *)dev_zero<double>(s->allocs, (size_t)nN * M);
        s->img_lat = (double *)dev_zero<double>(s->allocs, (size_t)nN * M);
        s->stg_loss = (double *)dev_zero<double>(s->allocs, (size_t)nN * M);
        s->img_loss = (double *)dev_zero<double>(s->allocs, (size_t)nN * M);
        s->stg_qual = (double *)dev_zero<double>(s->allocs, (size_t)nN * M * (nP ? nP : 1));
        s->img_qual = (double *)dev_zero<double>(s->allocs, (size_t)nN * M * (nP ? nP : 1));
        s->stg_depth = (double *)dev_zero<double>(s->allocs, (size_t)nN * M);
        s->stg_flow = (double *)dev_zero<double>(s->allocs, (size_t)nL * M);
    }
}

// queues one solver's step on the backend's current stream; wait = false leaves it in flight
static int step_host_enqueue(swb_solver *s, const swb_step_io *io, bool wait)
{
    if (!s || !io || !io->latflow) return fail(SWB_ERR_ARG, "null argument");
    const Net &n = s->net->net;
    const int M = s->M, nN = n.nN, nL = n.nL, nP = n.nP;
    const size_t nb = sizeof(double) * (size_t)nN * M, lb = sizeof(double) * (size_t)nL * M;
    const bool withQual = nP > 0 && !n.opt.ignore_quality;
    SWB_ENTER(s);
    try { ensure_staging(s); } catch (const DeviceError &e) { return fail(SWB_ERR_CUDA, e.what); }
    RunArgs stg;
    memset(&stg, 0, sizeof(stg));
    backend::h2d_async(s->stg_lat, io->latflow, nb);
    stg.stg_lat = s->stg_lat;
    if (io->node_losses) {
        backend::h2d_async(s->stg_loss, io->node_losses, nb);
        stg.stg_losses = s->stg_loss;
    }
    const bool haveQ = withQual && io->qual_load;
    if (haveQ) {
        backend::h2d_async(s->stg_qual, io->qual_load, nb * nP);
        stg.stg_qual = s->stg_qual;
    }
    if (io->node_depth) stg.stg_depth = s->stg_depth;
    if (io->link_flow) stg.stg_flow = s->stg_flow;
    int phases = PH_SWAP | PH_HOSTIN | PH_DYNWAVE | PH_NEXTDT | PH_MASSBAL | PH_STATS;
    if (withQual) phases |= PH_QSWAP | PH_QUALITY;
    double t_end = 0.0;
    if (io->dt) {
        backend::h2d_async(s->st.dt, io->dt, sizeof(double) * M);
        // a host-chosen step means dynwave_getRoutingStep has been called before (dynwave.c:209):
        // the Courant search that follows this step must not take the first-call shortcut
        backend::h2d_async(s->st.var_step, io->dt, sizeof(double) * M);
    }
    else { phases |= PH_ADVANCE; t_end = 1.0e300; }
    int rc = run(s, phases, 1, t_end, n.opt.route_step, s->img_lat, io->node_losses ? s->img_loss : nullptr,
                 haveQ ? s->img_qual : nullptr, wait, &stg);
    if (rc) return rc;
    if (io->node_depth) backend::d2h_async(io->node_depth, s->stg_depth, nb);
    if (io->link_flow)  backend::d2h_async(io->link_flow, s->stg_flow, lb);
    if (io->next_dt) backend::d2h_async(io->next_dt, s->st.var_step, sizeof(double) * M);
    if (io->iters)   backend::d2h_async(io->iters, s->st.iters, sizeof(int) * M);
    if (!wait) return SWB_OK;
    std::string err;
    if (!backend::sync(err)) return fail(SWB_ERR_CUDA, err);
    return SWB_OK;
}

int swb_step_host(swb_solver *s, const swb_step_io *io) { return step_host_enqueue(s, io, true); }

int swb_step_host_batch(swb_solver *const *solvers, const swb_step_io *io, int n)
{
    if (!solvers || !io || n < 1) return fail(SWB_ERR_ARG, "null argument");
    std::string err;
    for (int i = 0; i < n; i++) {
        if (!solvers[i]) return fail(SWB_ERR_ARG, "null solver in batch");
        for (int k = 0; k < i; k++)
            if (solvers[k] == solvers[i]) return fail(SWB_ERR_ARG, "the same solver twice in one batch");
        SWB_ENTER(solvers[i]);
        if (!solvers[i]->stream) solvers[i]->stream = backend::stream_create();
        try { ensure_staging(solvers[i]); } catch (const DeviceError &e) { return fail(SWB_ERR_CUDA, e.what); }
    }
    // staging buffers are allocated (and zeroed on the default stream) at a solver's first step
    if (!backend::sync(err)) return fail(SWB_ERR_CUDA, err);
    int rc = SWB_OK;
    for (int i = 0; i < n && rc == SWB_OK; i++) {
        backend::use_stream(solvers[i]->stream);
        rc = step_host_enqueue(solvers[i], &io[i], false);
    }
    backend::use_stream(nullptr);
    if (!backend::sync(err)) return fail(SWB_ERR_CUDA, err);
    return rc;
}

int swb_set_inflows(swb_solver *s, const swb_inflow_desc *d)
{
    if (!s || !d) return fail(SWB_ERR_ARG, "null argument");
    const int nN = s->net->net.nN, nP = s->net->net.nP, n = d->n_inflow_nodes;
    std::vector<int> slot(nN, -1);
    for (int k = 0; k < n; k++) {
        if (d->node[k] < 0 || d->node[k] >= nN) return fail(SWB_ERR_ARG, "inflow node out of range");
        slot[d->node[k]] = k;
    }
    auto pat_ok = [&](int p) { return p < d->n_patterns; };
    for (int k = 0; k < n && d->base_pattern; k++)
        if (!pat_ok(d->base_pattern[k])) return fail(SWB_ERR_ARG, "baseline pattern out of range");
    for (int r = 0; r < d->n_qual_inflows; r++) {
        if (d->q_node[r] < 0 || d->q_node[r] >= nN || d->q_pollut[r] < 0 || d->q_pollut[r] >= nP)
            return fail(SWB_ERR_ARG, "pollutant inflow record out of range");
        if (d->q_series[r] >= d->n_series || !pat_ok(d->q_pattern[r])) return fail(SWB_ERR_ARG, "pollutant inflow series / pattern out of range");
    }
    for (int r = 0; r < d->n_dwf; r++) {
        if (d->dwf_node[r] < 0 || d->dwf_node[r] >= nN || d->dwf_param[r] >= nP) return fail(SWB_ERR_ARG, "dry-weather record out of range");
        for (int q = 0; q < 4; q++) if (!pat_ok(d->dwf_patterns[(size_t)r * 4 + q])) return fail(SWB_ERR_ARG, "dry-weather pattern out of range");
    }
    for (int p = 0; p < d->n_patterns; p++)
        if (d->pattern_type[p] < 0 || d->pattern_type[p] > 3) return fail(SWB_ERR_ARG, "pattern type out of range");
    if (d->n_iface_nodes < 0 || d->n_iface_records < 0) return fail(SWB_ERR_ARG, "negative interface-file count");
    for (int r = 1; r < d->n_iface_records && d->n_iface_nodes > 0; r++)
        if (!(d->iface_date[r] > d->iface_date[r - 1])) return fail(SWB_ERR_ARG, "interface-file records are not in ascending date order");
    SWB_ENTER(s);
    // a second call replaces the first: its device arrays are released, not leaked
    for (void *q : s->inflow_allocs) backend::free_(q);
    s->inflow_allocs.clear();
    s->have_inflows = false;
    Inflows &f = s->inflows;
    f.n = n; f.start_day = d->start_day; f.start_secs = d->start_secs;
    try {
        std::vector<void *> &al = s->inflow_allocs;
        f.node = dev_copy<int>(al, d->node, n);
        f.ts_start = dev_copy<int>(al, d->ts_start, n + 1);
        f.ts_t = dev_copy<double>(al, d->ts_t, d->n_ts_pts);
        f.ts_q = dev_copy<double>(al, d->ts_q, d->n_ts_pts);
        f.sfactor = dev_copy<double>(al, d->sfactor, n);
        f.baseline = dev_copy<double>(al, d->baseline, n);
        std::vector<double> zc((size_t)n * (nP ? nP : 1), 0.0);
        f.concen = dev_copy<double>(al, d->concen ? d->concen : zc.data(), (size_t)n * nP);
        std::vector<double> one(s->M, 1.0), zero(s->M, 0.0);
        f.member_scale = dev_copy<double>(al, d->member_scale ? d->member_scale : one.data(), s->M);
        f.member_shift = dev_copy<double>(al, d->member_shift ? d->member_shift : zero.data(), s->M);
        f.node_slot = dev_copy<int>(al, slot.data(), nN);
        // ---- optional parts: patterns, pollutant inflow records, dry-weather flow
        f.general = 0;
        f.cfactor = nullptr; f.base_pattern = nullptr; f.nPatterns = 0; f.pat_type = nullptr; f.pat_factor = nullptr;
        f.node_q_start = nullptr; f.node_dwf_start = nullptr; f.pollut_dwf_concen = nullptr;
        if (d->cfactor) { f.cfactor = dev_copy<double>(al, d->cfactor, n); f.general = 1; }
        if (d->n_patterns > 0) {
            f.nPatterns = d->n_patterns;
            f.pat_type = dev_copy<int>(al, d->pattern_type, d->n_patterns);
            f.pat_factor = dev_copy<double>(al, d->pattern_factor, (size_t)d->n_patterns * 24);
        }
        if (d->base_pattern) { f.base_pattern = dev_copy<int>(al, d->base_pattern, n); f.general = 1; }
        if (d->n_qual_inflows > 0) {
            // records grouped by node, a node's records in the order given (= the reference's list order)
            const int nq = d->n_qual_inflows;
            std::vector<int> start(nN + 1, 0), order(nq);
            for (int r = 0; r < nq; r++) start[d->q_node[r] + 1]++;
            for (int i = 0; i < nN; i++) start[i + 1] += start[i];
            { std::vector<int> fill(start.begin(), start.end() - 1);
              for (int r = 0; r < nq; r++) order[fill[d->q_node[r]]++] = r; }
            auto gi = [&](const int *src) { std::vector<int> v(nq); for (int r = 0; r < nq; r++) v[r] = src[order[r]]; return v; };
            auto gd = [&](const double *src) { std::vector<double> v(nq); for (int r = 0; r < nq; r++) v[r] = src[order[r]]; return v; };
            f.node_q_start = dev_copy<int>(al, start.data(), nN + 1);
            f.q_pollut = dev_copy<int>(al, gi(d->q_pollut).data(), nq);
            f.q_type = dev_copy<int>(al, gi(d->q_type).data(), nq);
            f.q_series = dev_copy<int>(al, gi(d->q_series).data(), nq);
            f.q_pattern = dev_copy<int>(al, gi(d->q_pattern).data(), nq);
            f.q_cfactor = dev_copy<double>(al, gd(d->q_cfactor).data(), nq);
            f.q_sfactor = dev_copy<double>(al, gd(d->q_sfactor).data(), nq);
            f.q_baseline = dev_copy<double>(al, gd(d->q_baseline).data(), nq);
            const int ns = d->n_series;
            std::vector<int> s0(1, 0);
            f.series_start = dev_copy<int>(al, ns > 0 ? d->series_start : s0.data(), (ns > 0 ? ns : 0) + 1);
            const int npts = ns > 0 ? d->series_start[ns] : 0;
            f.series_t = dev_copy<double>(al, d->series_t, npts);
            f.series_v = dev_copy<double>(al, d->series_v, npts);
            f.general = 1;
        }
        if (d->n_dwf > 0) {
            const int nd = d->n_dwf;
            std::vector<int> start(nN + 1, 0), order(nd);
            for (int r = 0; r < nd; r++) start[d->dwf_node[r] + 1]++;
            for (int i = 0; i < nN; i++) start[i + 1] += start[i];
            { std::vector<int> fill(start.begin(), start.end() - 1);
              for (int r = 0; r < nd; r++) order[fill[d->dwf_node[r]]++] = r; }
            std::vector<int> param(nd), pats((size_t)nd * 4);
            std::vector<double> avg(nd);
            for (int r = 0; r < nd; r++) {
                param[r] = d->dwf_param[order[r]]; avg[r] = d->dwf_avg[order[r]];
                for (int q = 0; q < 4; q++) pats[(size_t)r * 4 + q] = d->dwf_patterns[(size_t)order[r] * 4 + q];
            }
            f.node_dwf_start = dev_copy<int>(al, start.data(), nN + 1);
            f.dwf_param = dev_copy<int>(al, param.data(), nd);
            f.dwf_avg = dev_copy<double>(al, avg.data(), nd);
            f.dwf_patterns = dev_copy<int>(al, pats.data(), (size_t)nd * 4);
            if (d->pollut_dwf_concen && nP > 0) f.pollut_dwf_concen = dev_copy<double>(al, d->pollut_dwf_concen, nP);
            f.general = 1;
        }
        f.nIfaceNodes = f.nIfaceRec = 0; f.if_slot = nullptr; f.if_date = f.if_val = nullptr;
        if (d->n_iface_nodes > 0 && d->n_iface_records > 0) {
            std::vector<int> islot(nN, -1);
            for (int k = 0; k < d->n_iface_nodes; k++)
                if (d->iface_node[k] >= 0 && d->iface_node[k] < nN) islot[d->iface_node[k]] = k;
            f.nIfaceNodes = d->n_iface_nodes; f.nIfaceRec = d->n_iface_records;
            f.if_slot = dev_copy<int>(al, islot.data(), nN);
            f.if_date = dev_copy<double>(al, d->iface_date, d->n_iface_records);
            f.if_val = dev_copy<double>(al, d->iface_value, (size_t)d->n_iface_records * d->n_iface_nodes * (1 + nP));
            f.general = 1;
        }
    } catch (const DeviceError &e) {
        for (void *q : s->inflow_allocs) backend::free_(q);
        s->inflow_allocs.clear();
        memset(&s->inflows, 0, sizeof(s->inflows));
        return fail(SWB_ERR_CUDA, e.what);
    }
    s->have_inflows = true;
    return SWB_OK;
}

// Control rules, pump start-up / shut-off depths and timed outfall stages for swb_run_steps.
int swb_set_controls(swb_solver *s, const swb_controls_desc *d)
{
    if (!s || !d) return fail(SWB_ERR_ARG, "null argument");
    const Net &n = s->net->net;
    const int nL = n.nL, nN = n.nN, M = s->M;
    const int nR = d->n_rules, nPr = d->n_premises, nA = d->n_actions;
    if (nR < 0 || nPr < 0 || nA < 0 || d->n_series < 0 || d->n_stage_nodes < 0) return fail(SWB_ERR_ARG, "negative count");
    if (!d->pump_y_on || !d->pump_y_off || !d->orif_orate) return fail(SWB_ERR_ARG, "per-link arrays missing");
    if (s->st.halo.nRanks > 1) return fail(SWB_ERR_UNSUPP, "control rules on a partitioned network");
    std::vector<int> ltype(nL);
    backend::download(ltype.data(), n.link_type, sizeof(int) * nL);
    for (int r = 0; r < nR; r++) {
        if (d->rule_premise_start[r] > d->rule_premise_start[r + 1] || d->rule_then_start[r] > d->rule_then_start[r + 1] ||
            d->rule_else_start[r] > d->rule_else_start[r + 1]) return fail(SWB_ERR_ARG, "rule CSR arrays not ascending");
    }
    if (nR > 0 && (d->rule_premise_start[nR] != nPr || d->rule_then_start[nR] + d->rule_else_start[nR] != nA))
        return fail(SWB_ERR_ARG, "rule CSR arrays do not cover the premises / actions");
    auto var_ok = [&](int obj, int idx, int attr) -> int {
        if (obj == RO_GAGE) return SWB_ERR_UNSUPP;
        if (obj != RO_NODE && obj != RO_LINK && obj != RO_SIM && obj != -1) return SWB_ERR_ARG;
        if (obj == RO_NODE && (idx < 0 || idx >= nN)) return SWB_ERR_ARG;
        if (obj == RO_LINK && (idx < 0 || idx >= nL)) return SWB_ERR_ARG;
        if (attr < 0 || attr > RA_MONTH) return SWB_ERR_ARG;
        return SWB_OK;
    };
    for (int p = 0; p < nPr; p++) {
        int rc = var_ok(d->prem_lhs_obj[p], d->prem_lhs_index[p], d->prem_lhs_attr[p]);
        if (!rc && d->prem_rhs_is_var[p]) rc = var_ok(d->prem_rhs_obj[p], d->prem_rhs_index[p], d->prem_rhs_attr[p]);
        if (rc == SWB_ERR_UNSUPP) return fail(rc, "rain-gage premises are not supported on the device");
        if (rc) return fail(rc, "premise variable out of range");
        if (d->prem_relation[p] < 0 || d->prem_relation[p] > 5) return fail(SWB_ERR_ARG, "premise relation out of range");
    }
    for (int a = 0; a < nA; a++) {
        if (d->act_link[a] < 0 || d->act_link[a] >= nL) return fail(SWB_ERR_ARG, "action link out of range");
        if (d->act_rule[a] < 0 || d->act_rule[a] >= nR) return fail(SWB_ERR_ARG, "action rule out of range");
        if (d->act_curve[a] >= n.nCurves) return fail(SWB_ERR_ARG, "action curve out of range");
        if (d->act_tseries[a] >= d->n_series) return fail(SWB_ERR_ARG, "action time series out of range");
    }
    for (int k = 0; k < d->n_stage_nodes; k++) {
        if (d->stage_node[k] < 0 || d->stage_node[k] >= nN) return fail(SWB_ERR_ARG, "stage node out of range");
        const int kind = d->stage_kind[k], t = d->stage_table[k];
        if (!((kind == 1 && t >= 0 && t < n.nCurves) || (kind == 2 && t >= 0 && t < d->n_series)))
            return fail(SWB_ERR_ARG, "stage curve / series out of range");
    }
    // the watch list: links whose setting can change during the run
    std::vector<int> link_watch(nL, -1), watch;
    {
        std::vector<char> mark(nL, 0);
        for (int j = 0; j < nL; j++) if (ltype[j] == SWB_PUMP) mark[j] = 1;
        for (int a = 0; a < nA; a++) mark[d->act_link[a]] = 1;
        for (int j = 0; j < nL; j++) if (mark[j]) { link_watch[j] = (int)watch.size(); watch.push_back(j); }
    }
    const int nW = (int)watch.size();
    SWB_ENTER(s);
    for (void *q : s->control_allocs) backend::free_(q);
    s->control_allocs.clear();
    Controls &c = s->controls;
    memset(&c, 0, sizeof(c));
    try {
        std::vector<void *> &al = s->control_allocs;
        c.nRules = nR; c.nAct = nA; c.nWatch = nW; c.nStage = d->n_stage_nodes;
        c.rule_step = d->rule_step; c.start_datetime = d->start_datetime;
        c.start_day = d->start_day; c.start_secs = d->start_secs;
        c.rule_priority = dev_copy<double>(al, d->rule_priority, nR);
        c.rule_prem_start = dev_copy<int>(al, d->rule_premise_start, nR + 1);
        c.rule_then_start = dev_copy<int>(al, d->rule_then_start, nR + 1);
        c.rule_else_start = dev_copy<int>(al, d->rule_else_start, nR + 1);
        c.act_then = dev_copy<int>(al, d->act_then, nR > 0 ? d->rule_then_start[nR] : 0);
        c.act_else = dev_copy<int>(al, d->act_else, nR > 0 ? d->rule_else_start[nR] : 0);
        c.prem_type = dev_copy<int>(al, d->prem_type, nPr);
        c.prem_lhs_obj = dev_copy<int>(al, d->prem_lhs_obj, nPr);
        c.prem_lhs_index = dev_copy<int>(al, d->prem_lhs_index, nPr);
        c.prem_lhs_attr = dev_copy<int>(al, d->prem_lhs_attr, nPr);
        c.prem_rhs_is_var = dev_copy<int>(al, d->prem_rhs_is_var, nPr);
        c.prem_rhs_obj = dev_copy<int>(al, d->prem_rhs_obj, nPr);
        c.prem_rhs_index = dev_copy<int>(al, d->prem_rhs_index, nPr);
        c.prem_rhs_attr = dev_copy<int>(al, d->prem_rhs_attr, nPr);
        c.prem_relation = dev_copy<int>(al, d->prem_relation, nPr);
        c.prem_value = dev_copy<double>(al, d->prem_value, nPr);
        c.act_rule = dev_copy<int>(al, d->act_rule, nA);
        c.act_link = dev_copy<int>(al, d->act_link, nA);
        c.act_attr = dev_copy<int>(al, d->act_attr, nA);
        c.act_curve = dev_copy<int>(al, d->act_curve, nA);
        c.act_tseries = dev_copy<int>(al, d->act_tseries, nA);
        std::vector<int> slot(nA);
        for (int a = 0; a < nA; a++) slot[a] = link_watch[d->act_link[a]];
        c.act_slot = dev_copy<int>(al, slot.data(), nA);
        c.act_kp = dev_copy<double>(al, d->act_kp, nA);
        c.act_ki = dev_copy<double>(al, d->act_ki, nA);
        c.act_kd = dev_copy<double>(al, d->act_kd, nA);
        const int ns = d->n_series;
        std::vector<int> s0(1, 0);
        c.series_start = dev_copy<int>(al, ns > 0 ? d->series_start : s0.data(), ns + 1);
        const int npts = ns > 0 ? d->series_start[ns] : 0;
        c.series_t = dev_copy<double>(al, d->series_t, npts);
        c.series_v = dev_copy<double>(al, d->series_v, npts);
        c.pump_y_on = dev_copy<double>(al, d->pump_y_on, nL);
        c.pump_y_off = dev_copy<double>(al, d->pump_y_off, nL);
        c.orif_orate = dev_copy<double>(al, d->orif_orate, nL);
        c.watch = dev_copy<int>(al, watch.data(), nW);
        c.link_watch = dev_copy<int>(al, link_watch.data(), nL);
        c.stage_node = dev_copy<int>(al, d->stage_node, d->n_stage_nodes);
        c.stage_kind = dev_copy<int>(al, d->stage_kind, d->n_stage_nodes);
        c.stage_table = dev_copy<int>(al, d->stage_table, d->n_stage_nodes);
        // per-member rule state
        c.control_value = dev_zero<double>(al, M);
        c.set_point = dev_zero<double>(al, M);
        std::vector<double> av((size_t)nA * M), tls((size_t)nW * M);
        for (int a = 0; a < nA; a++) for (int m = 0; m < M; m++) av[(size_t)a * M + m] = d->act_value[a];
        for (int w = 0; w < nW; w++)
            for (int m = 0; m < M; m++) tls[(size_t)w * M + m] = d->link_time_last_set ? d->link_time_last_set[watch[w]] : d->start_datetime;
        c.act_val = dev_copy<double>(al, av.data(), av.size());
        c.act_e1 = dev_zero<double>(al, (size_t)nA * M);
        c.act_e2 = dev_zero<double>(al, (size_t)nA * M);
        c.time_last_set = dev_copy<double>(al, tls.data(), tls.size());
        c.winner = dev_zero<int>(al, (size_t)nW * M);
        c.new_rule_time = dev_zero<double>(al, M);
        c.active = 1;
    } catch (const DeviceError &e) {
        for (void *q : s->control_allocs) backend::free_(q);
        s->control_allocs.clear();
        memset(&s->controls, 0, sizeof(s->controls));
        return fail(SWB_ERR_CUDA, e.what);
    }
    return SWB_OK;
}

// ---- re-enumeration of the member axis ------------------------------------------------------------------
// Every array with a member axis, as (pointer, rows, element size); built on demand so that arrays created
// later (statistics, inflows, rule state) are covered.
struct MemberArray { void *p; size_t rows; int esz; };
static std::vector<MemberArray> member_arrays(swb_solver *s)
{
    const Net &n = s->net->net;
    const State &st = s->st;
    const int nN = n.nN, nL = n.nL, nP = n.nP;
    std::vector<MemberArray> v;
    auto add = [&](const void *p, size_t rows, int esz) { if (p && rows) v.push_back({const_cast<void *>(p), rows, esz}); };
    for (const FieldInfo &f : field_table())
        add(*(void **)((char *)&st + f.offset), field_items(f, nN, nL, nP), f.is_u8 ? 1 : 8);
    add(st.dt, 1, 8); add(st.var_step, 1, 8); add(st.sim_time, 1, 8); add(st.time_ms, 1, 8);
    add(st.evap_rate, 1, 8); add(st.hydcon, 1, 8); add(st.iters, 1, 4);
    add(st.tot_iters, 1, 8); add(st.tot_steps, 1, 8); add(st.non_conv, 1, 8);
    add(st.crit_node, 1, 4); add(st.crit_link, 1, 4); add(st.tmin_bits, 2, 8); add(st.done, 1, 4);
    add(st.o_ynorm, n.nOutfallNodes, 8); add(st.o_ycrit, n.nOutfallNodes, 8);
    add(st.not_conv, SWB_MAX_TRIALS_CAP, 4);
    add(st.mb_reacted, nP ? nP : 1, 8); add(st.mb_seepage, nP ? nP : 1, 8); add(st.mb_final_storage, nP ? nP : 1, 8);
    add(st.mb_rate, MB_FLOW_TERMS + MB_QUAL_TERMS * nP, 8); add(st.mb_total, MB_FLOW_TERMS + MB_QUAL_TERMS * nP, 8);
    add(st.mb_dt_prev, 1, 8);
    add(st.stat_node, (size_t)(SWB_NS_PLANES + nP) * nN, 8); add(st.stat_link, (size_t)SWB_LS_PLANES * nL, 8);
    add(st.stat_sys, SWB_SS_PLANES, 8);
    if (s->have_inflows) { add(s->inflows.member_scale, 1, 8); add(s->inflows.member_shift, 1, 8); }
    const Controls &c = s->controls;
    if (c.active) {
        add(c.control_value, 1, 8); add(c.set_point, 1, 8); add(c.new_rule_time, 1, 8);
        add(c.act_val, c.nAct, 8); add(c.act_e1, c.nAct, 8); add(c.act_e2, c.nAct, 8);
        add(c.time_last_set, c.nWatch, 8); add(c.winner, c.nWatch, 4);
    }
    return v;
}

// After the call member i is what member perm[i] was: state, clocks, counters, statistics, routing totals,
// inflow scale / shift and rule state move together (nothing else refers to a member by number).
int swb_permute_members(swb_solver *s, const int *perm)
{
    if (!s || !perm) return fail(SWB_ERR_ARG, "null argument");
    const int M = s->M;
    if (s->st.halo.nRanks > 1) return fail(SWB_ERR_UNSUPP, "a partitioned solver holds one member");
    std::vector<char> seen(M, 0);
    for (int i = 0; i < M; i++) {
        if (perm[i] < 0 || perm[i] >= M || seen[perm[i]]) return fail(SWB_ERR_ARG, "perm is not a permutation of the members");
        seen[perm[i]] = 1;
    }
    SWB_ENTER(s);
    { std::string err; if (!backend::sync(err)) return fail(SWB_ERR_CUDA, err); }     // nothing of this solver in flight
    int rc = SWB_OK;
    try {
        // rows are permuted in batches through one scratch buffer (64 MB, or one row of the widest element) that
        // stays with the solver: the call is meant to be repeated every few hundred steps
        const size_t tmp_bytes = std::max<size_t>((size_t)64 << 20, (size_t)M * 8);
        if (!s->perm_tmp) {
            s->perm_idx = dev_zero<int>(s->allocs, M);
            s->perm_tmp = checked_alloc(tmp_bytes);
            s->allocs.push_back(s->perm_tmp);
        }
        checked(backend::upload(s->perm_idx, perm, sizeof(int) * M), "host -> device copy");
        for (const MemberArray &a : member_arrays(s))
            if (!backend::permute_members(a.p, a.rows, a.esz, M, s->perm_idx, s->perm_tmp, tmp_bytes)) { rc = fail(SWB_ERR_CUDA, backend::last_error()); break; }
    } catch (const DeviceError &e) { rc = fail(SWB_ERR_CUDA, e.what); }
    { std::string err; if (!backend::sync(err) && rc == SWB_OK) rc = fail(SWB_ERR_CUDA, err); }
    return rc;
}

int swb_run_steps(swb_solver *s, int n_steps, double t_end)
{
    if (!s || n_steps < 1) return fail(SWB_ERR_ARG, "bad arguments");
    if (!s->have_inflows) return fail(SWB_ERR_ARG, "swb_set_inflows has not been called");
    SWB_ENTER(s);
    const Net &n = s->net->net;
    int phases = PH_ADVANCE | PH_SWAP | PH_INFLOWS | PH_DYNWAVE | PH_NEXTDT | PH_MASSBAL | PH_STATS;
    if (n.nP > 0 && !n.opt.ignore_quality) phases |= PH_QSWAP | PH_QUALITY;
    return run(s, phases, n_steps, t_end, n.opt.route_step);
}

int swb_partition_attach(swb_solver *s, const swb_partition_desc *p)
{
    if (!s || !p) return fail(SWB_ERR_ARG, "null argument");
    if (s->M != 1) return fail(SWB_ERR_ARG, "a partitioned solver holds one member");
    if (s->window) return fail(SWB_ERR_ARG, "partition already attached");
    SWB_ENTER(s);
    const Net &n = s->net->net;
    if (p->n_ranks < 1 || p->n_ranks > SWB_MAX_RANKS || p->rank < 0 || p->rank >= p->n_ranks)
        return fail(SWB_ERR_ARG, "bad rank / n_ranks");
    if (p->n_owned_nodes < 0 || p->n_owned_nodes > n.nN || p->n_send < 0 || p->n_recv < 0 ||
        p->n_owned_nodes + p->n_recv != n.nN || !p->link_owned)
        return fail(SWB_ERR_ARG, "inconsistent partition descriptor");
    for (int k = 0; k < p->n_send; k++)
        if (p->send_node[k] < 0 || p->send_node[k] >= p->n_owned_nodes || p->send_rank[k] < 0 ||
            p->send_rank[k] >= p->n_ranks || p->send_rank[k] == p->rank || p->send_slot[k] < 0)
            return fail(SWB_ERR_ARG, "bad send entry");
    for (int k = 0; k < p->n_recv; k++)
        if (p->recv_node[k] < p->n_owned_nodes || p->recv_node[k] >= n.nN) return fail(SWB_ERR_ARG, "bad receive entry");
    // only true conduits may touch a ghost node (ordered regulator pass)
    const swb_network *nw = s->net;
    std::vector<int> n1(n.nL), n2(n.nL);
    backend::download(n1.data(), n.link_node1, sizeof(int) * n.nL);
    backend::download(n2.data(), n.link_node2, sizeof(int) * n.nL);
    for (int j = 0; j < n.nL; j++)
        if ((n1[j] >= p->n_owned_nodes || n2[j] >= p->n_owned_nodes) && !(nw->der.link_flags[j] & LF_TRUE_CONDUIT))
            return fail(SWB_ERR_UNSUPP, "a pump / regulator / dummy link crosses the partition border");
    Halo &H = s->st.halo;
    memset(&H, 0, sizeof(H));
    try {
    H.rank = p->rank; H.nRanks = p->n_ranks; H.nOwnedN = p->n_owned_nodes;
    H.nSend = p->n_send; H.nRecv = p->n_recv; H.W = n.nP > 2 ? n.nP : 2;
    {   // send entries sorted by node + CSR over the owned nodes
        std::vector<int> idx(p->n_send), sn(p->n_send), sr(p->n_send), ss(p->n_send), start(p->n_owned_nodes + 1, 0);
        for (int k = 0; k < p->n_send; k++) idx[k] = k;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return p->send_node[a] < p->send_node[b]; });
        for (int k = 0; k < p->n_send; k++) {
            sn[k] = p->send_node[idx[k]]; sr[k] = p->send_rank[idx[k]]; ss[k] = p->send_slot[idx[k]];
            start[sn[k] + 1]++;
        }
        for (int i = 0; i < p->n_owned_nodes; i++) start[i + 1] += start[i];
        H.send_node = dev_copy<int>(s->allocs, sn.data(), sn.size());
        H.send_rank = dev_copy<int>(s->allocs, sr.data(), sr.size());
        H.send_slot = dev_copy<int>(s->allocs, ss.data(), ss.size());
        H.send_start = dev_copy<int>(s->allocs, start.data(), start.size());
    }
    H.recv_node = dev_copy<int>(s->allocs, p->recv_node, p->n_recv);
    H.link_owned = dev_copy<int>(s->allocs, p->link_owned, n.nL);
    {
        std::vector<int> type(n.nN), order;
        backend::download(type.data(), n.node_type, sizeof(int) * n.nN);
        node_order_expensive_first(type.data(), p->n_owned_nodes, n.nN, order);
        H.node_order = dev_copy<int>(s->allocs, order.data(), order.size());
    }
    } catch (const DeviceError &e) { memset(&H, 0, sizeof(H)); return fail(SWB_ERR_CUDA, e.what); }
    H.wait_ns = s->st.phase_ns + TP_HALO_WAIT;
    H.timeout_ns = (unsigned long long)((p->timeout_s > 0.0 ? p->timeout_s : 30.0) * 1.0e9);
    std::string err;
    s->window_bytes = window_size(H.nRecv, H.W);
    s->window = backend::window_alloc(s->window_bytes, s->window_handle, err);
    if (!s->window) { H.nRanks = 0; return fail(SWB_ERR_CUDA, err); }
    window_views(s->window, H.ctrl, H.red, H.stage);
    window_views(s->window, H.peer_ctrl[H.rank], H.peer_red[H.rank], H.peer_stage[H.rank]);
    return SWB_OK;
}

int swb_partition_export(swb_solver *s, void *handle)
{
    if (!s || !handle || !s->window) return fail(SWB_ERR_ARG, "no partition attached");
    memcpy(handle, s->window_handle, SWB_WINDOW_HANDLE_BYTES);
    return SWB_OK;
}

int swb_partition_connect(swb_solver *s, int peer, const void *handle)
{
    if (!s || !handle || !s->window) return fail(SWB_ERR_ARG, "no partition attached");
    Halo &H = s->st.halo;
    if (peer < 0 || peer >= H.nRanks || peer == H.rank) return fail(SWB_ERR_ARG, "bad peer rank");
    if (s->peer_window[peer]) return fail(SWB_ERR_ARG, "peer already connected");
    std::string err;
    void *w = backend::window_open(handle, 0, err);
    if (!w) return fail(SWB_ERR_CUDA, err);
    s->peer_window[peer] = w;
    window_views(w, H.peer_ctrl[peer], H.peer_red[peer], H.peer_stage[peer]);
    return SWB_OK;
}

long long swb_partition_exchanges(swb_solver *s)
{
    if (!s || !s->window) return 0;
    unsigned long long e = 0;
    backend::download(&e, s->st.halo.ctrl + HALO_EPOCH, sizeof(e));
    return (long long)e;
}

int swb_get_stats(swb_solver *s, int m0, int nm, swb_member_stats *out)
{
    if (!s || !out || m0 < 0 || nm < 1 || m0 + nm > s->M) return fail(SWB_ERR_ARG, "bad arguments");
    SWB_ENTER(s);
    const int M = s->M;
    std::vector<double> t(M), dt(M), vs(M);
    std::vector<long long> st(M), it(M), nc(M);
    std::vector<int> cn(M), cl(M);
    backend::download(t.data(), s->st.sim_time, sizeof(double) * M);
    backend::download(dt.data(), s->st.dt, sizeof(double) * M);
    backend::download(vs.data(), s->st.var_step, sizeof(double) * M);
    backend::download(st.data(), s->st.tot_steps, sizeof(long long) * M);
    backend::download(it.data(), s->st.tot_iters, sizeof(long long) * M);
    backend::download(nc.data(), s->st.non_conv, sizeof(long long) * M);
    backend::download(cn.data(), s->st.crit_node, sizeof(int) * M);
    backend::download(cl.data(), s->st.crit_link, sizeof(int) * M);
    for (int k = 0; k < nm; k++) {
        int m = m0 + k;
        out[k].sim_time = t[m]; out[k].last_dt = dt[m]; out[k].next_dt = vs[m];
        out[k].steps = st[m]; out[k].iterations = it[m]; out[k].non_converged = nc[m];
        out[k].crit_node = cn[m]; out[k].crit_link = cl[m];
    }
    return SWB_OK;
}

int swb_get_massbal(swb_solver *s, int m0, int nm, double *reacted, double *seepage, double *final_storage)
{
    if (!s || m0 < 0 || nm < 1 || m0 + nm > s->M) return fail(SWB_ERR_ARG, "bad arguments");
    SWB_ENTER(s);
    const int M = s->M, nP = s->net->net.nP;
    std::vector<double> h((size_t)(nP ? nP : 1) * M);
    double *dst[3] = { reacted, seepage, final_storage };
    const double *src[3] = { s->st.mb_reacted, s->st.mb_seepage, s->st.mb_final_storage };
    for (int k = 0; k < 3; k++) {
        if (!dst[k]) continue;
        backend::download(h.data(), src[k], sizeof(double) * h.size());
        for (int mm = 0; mm < nm; mm++)
            for (int p = 0; p < nP; p++) dst[k][mm * nP + p] = h[(size_t)p * M + m0 + mm];
    }
    return SWB_OK;
}

int swb_get_routing_totals(swb_solver *s, int m0, int nm, double *flow, double *qual)
{
    if (!s || m0 < 0 || nm < 1 || m0 + nm > s->M) return fail(SWB_ERR_ARG, "bad arguments");
    SWB_ENTER(s);
    const int M = s->M, nP = s->net->net.nP, nMb = MB_FLOW_TERMS + MB_QUAL_TERMS * nP;
    std::vector<double> tot((size_t)nMb * M), rate((size_t)nMb * M), dtp(M);
    backend::download(tot.data(), s->st.mb_total, sizeof(double) * tot.size());
    backend::download(rate.data(), s->st.mb_rate, sizeof(double) * rate.size());
    backend::download(dtp.data(), s->st.mb_dt_prev, sizeof(double) * M);
    for (int mm = 0; mm < nm; mm++) {
        const int m = m0 + mm;
        for (int q = 0; q < nMb; q++) {
            // the last step's rates over its own second half (routing.c:264), still pending on the device
            double v = tot[(size_t)q * M + m];
            const bool mass = q >= MB_FLOW_TERMS && (q - MB_FLOW_TERMS) % MB_QUAL_TERMS == MBQ_FINAL;
            v += mass ? rate[(size_t)q * M + m] : rate[(size_t)q * M + m] * (dtp[m] / 2.);
            if (q < MB_FLOW_TERMS) { if (flow) flow[mm * MB_FLOW_TERMS + q] = v; }
            else if (qual) qual[(size_t)mm * MB_QUAL_TERMS * nP + (q - MB_FLOW_TERMS)] = v;
        }
    }
    return SWB_OK;
}

long long swb_conduit_updates(swb_solver *s)
{
    if (!s || enter(s)) return 0;
    std::vector<long long> it(s->M);
    backend::download(it.data(), s->st.tot_iters, sizeof(long long) * s->M);
    long long sum = 0;
    for (long long v : it) sum += v;
    return sum * (long long)s->net->net.nTrue;
}

int swb_xsect_eval(int device, int fn, int xs_type, const double *p, int n, const double *args, double *out)
{
    if (!p || !args || !out || n < 0 || fn < 0 || fn > 8) return fail(SWB_ERR_ARG, "bad arguments");
    Xs x;
    x.type = xs_type; x.ntbl = 0; x.atbl = x.rtbl = x.wtbl = nullptr;
    x.yFull = p[0]; x.wMax = p[1]; x.ywMax = p[2]; x.aFull = p[3]; x.rFull = p[4]; x.sFull = p[5];
    x.sMax = p[6]; x.yBot = p[7]; x.aBot = p[8]; x.sBot = p[9]; x.rBot = p[10];
    x.rYFull = exact_rcp(x.yFull);
    std::string err;
    if (!backend::xsect_eval(device, fn, x, n, args, out, err)) return fail(SWB_ERR_CUDA, err);
    return SWB_OK;
}

int swb_get_results(swb_solver *s, const double *f, int m0, int nm, float *node_out, float *link_out)
{
    if (!s || !f || m0 < 0 || nm < 1 || m0 + nm > s->M) return fail(SWB_ERR_ARG, "bad arguments");
    const Net &n = s->net->net;
    const size_t nrec = (size_t)node_record_len(n) * n.nN * nm, lrec = (size_t)link_record_len(n) * n.nL * nm;
    std::string err;
    if (!backend::init(s->net->device, err)) return fail(SWB_ERR_CUDA, err);
    double *df = (double *)backend::alloc(sizeof(double) * s->M);
    float *dn = node_out ? (float *)backend::alloc(sizeof(float) * nrec) : nullptr;
    float *dl = link_out ? (float *)backend::alloc(sizeof(float) * lrec) : nullptr;
    backend::upload(df, f, sizeof(double) * s->M);
    bool ok = backend::report(s->net->device, n, s->st, df, m0, nm, dn, dl, err);
    if (ok && dn) backend::download(node_out, dn, sizeof(float) * nrec);
    if (ok && dl) backend::download(link_out, dl, sizeof(float) * lrec);
    backend::free_(df);
    if (dn) backend::free_(dn);
    if (dl) backend::free_(dl);
    if (!ok) return fail(SWB_ERR_CUDA, err);
    s->launches += (node_out ? 1 : 0) + (link_out ? 1 : 0);
    return SWB_OK;
}

int swb_get_phase_times(swb_solver *s, double *ms, int n, int reset)
{
    if (!s || !ms || n < 1) return fail(SWB_ERR_ARG, "bad arguments");
    SWB_ENTER(s);
    unsigned long long h[SWB_N_PHASES];
    backend::download(h, s->st.phase_ns, sizeof(h));
    for (int i = 0; i < n; i++) ms[i] = i < SWB_N_PHASES ? (double)h[i] * 1.0e-6 : 0.0;
    if (reset) backend::zero(s->st.phase_ns, sizeof(h));
    return SWB_OK;
}

int swb_enable_statistics(swb_solver *s, double report_start_s)
{
    if (!s) return fail(SWB_ERR_ARG, "null solver");
    SWB_ENTER(s);
    const Net &n = s->net->net;
    const size_t M = s->M;
    try {
        if (!s->st.stat_node) {
            s->st.stat_node = dev_zero<double>(s->allocs, (size_t)(SWB_NS_PLANES + n.nP) * n.nN * M);
            s->st.stat_link = dev_zero<double>(s->allocs, (size_t)SWB_LS_PLANES * n.nL * M);
            s->st.stat_sys = dev_zero<double>(s->allocs, (size_t)SWB_SS_PLANES * M);
        } else {
            checked(backend::zero(s->st.stat_node, sizeof(double) * (size_t)(SWB_NS_PLANES + n.nP) * n.nN * M), "memset");
            checked(backend::zero(s->st.stat_link, sizeof(double) * (size_t)SWB_LS_PLANES * n.nL * M), "memset");
            checked(backend::zero(s->st.stat_sys, sizeof(double) * (size_t)SWB_SS_PLANES * M), "memset");
        }
        // TimeStepStats.minTimeStep starts at RouteStep (stats.c:293)
        std::vector<double> mn(M, n.opt.route_step);
        checked(backend::upload(s->st.stat_sys + (size_t)SWB_SS_MIN_DT * M, mn.data(), sizeof(double) * M), "copy");
    } catch (const DeviceError &e) { s->st.stat_node = s->st.stat_link = s->st.stat_sys = nullptr; return fail(SWB_ERR_CUDA, e.what); }
    s->st.stat_report_start = report_start_s;
    return SWB_OK;
}

// planes [K][items][M] on the device -> [member][plane][item] on the host
static int stats_xfer(swb_solver *s, const double *dev, int planes, size_t items, int m0, int nm, double *out)
{
    if (!s || !out || m0 < 0 || nm < 1 || m0 + nm > s->M) return fail(SWB_ERR_ARG, "bad arguments");
    if (!dev) return fail(SWB_ERR_ARG, "swb_enable_statistics has not been called");
    SWB_ENTER(s);
    const size_t rows = (size_t)planes * items, cols = (size_t)nm;
    std::vector<double> h(rows * cols);
    if (!backend::copy2d(h.data(), cols * sizeof(double), dev + m0, (size_t)s->M * sizeof(double),
                         cols * sizeof(double), rows, false))
        return fail(SWB_ERR_CUDA, backend::last_error());
    for (size_t r = 0; r < rows; r++)
        for (int mm = 0; mm < nm; mm++) out[(size_t)mm * rows + r] = h[r * cols + mm];
    return SWB_OK;
}
int swb_get_node_stats(swb_solver *s, int m0, int nm, double *out)
{ return s ? stats_xfer(s, s->st.stat_node, SWB_NS_PLANES + s->net->net.nP, s->net->net.nN, m0, nm, out) : fail(SWB_ERR_ARG, "null solver"); }
int swb_get_link_stats(swb_solver *s, int m0, int nm, double *out)
{ return s ? stats_xfer(s, s->st.stat_link, SWB_LS_PLANES, s->net->net.nL, m0, nm, out) : fail(SWB_ERR_ARG, "null solver"); }
int swb_get_system_stats(swb_solver *s, int m0, int nm, double *out)
{
    if (!s) return fail(SWB_ERR_ARG, "null solver");
    return stats_xfer(s, s->st.stat_sys, SWB_SS_PLANES, 1, m0, nm, out);
}

// profiling aid (tools/profile_phase.py): n_steps of the given phase mask with debug switches
// (swb_engine.h: DBG_*), optionally bracketed by the profiler start / stop markers
int swb_debug_run(swb_solver *s, int phases, int n_steps, int debug, int profile)
{
    if (!s || n_steps < 1) return fail(SWB_ERR_ARG, "bad arguments");
    SWB_ENTER(s);
    RunArgs a;
    memset(&a, 0, sizeof(a));
    a.debug = debug;
    if (profile) backend::profiler(true);
    int rc = run(s, phases, n_steps, 1.0e300, s->net->net.opt.route_step, nullptr, nullptr, nullptr, true, &a);
    if (profile) backend::profiler(false);
    return rc;
}

long long swb_launch_count(const swb_solver *s) { return s ? s->launches : 0; }
int swb_set_staged_min_members(int min_members) { return backend::set_staged_min_members(min_members); }
double swb_last_kernel_ms(const swb_solver *s) { return s ? (double)s->last_ms : 0.0; }
int swb_sync(swb_solver *s)
{
    if (s) SWB_ENTER(s);
    std::string err;
    if (!backend::sync(err)) return fail(SWB_ERR_CUDA, err);
    return SWB_OK;
}

} // extern "C"
#endif
