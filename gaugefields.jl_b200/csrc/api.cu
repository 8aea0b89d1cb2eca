// api.cu -- the extern "C" boundary of libgfb200.so (declared in include/gfb200.h).
//
// Host-side orchestration only: handle management, host<->device layout conversion, t-slab halo
// exchange over NCCL, deterministic scalar reductions and the integrator loops that string the
// kernels of kernels.cu / stout.cu together.  No CPU fallback exists: without a usable GPU every
// entry point fails with GFB_ERR_NODEVICE / GFB_ERR_CUDA.
#include <cmath>
#include <cstring>

#include "gfb_internal.h"

namespace gfb {

static std::string g_init_error;

int fail(gfb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_init_error = msg;
    return code;
}

Geom make_geom(const gfb_ctx* ctx, int nx, int ny, int nz, int nt, int slab_global_index) {
    Geom g;
    const int G = ctx->nslabs_total;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.tloc = nt / G;
    g.v3 = nx * ny * nz;
    g.nt = nt;
    g.t_stride = 1;
    g.t0 = slab_global_index * g.tloc;
    if (G > 1) {
        g.t_up_wrap = g.tloc;
        g.t_dn_wrap = g.tloc + 1;
        g.nslots = g.tloc + 2;
    } else {
        g.t_up_wrap = 0;
        g.t_dn_wrap = g.tloc - 1;
        g.nslots = g.tloc;
    }
    // z-chunk of the traversal: largest divisor of nz whose xy*zc slab of links stays <= 24 MiB, so the
    // t+-1 neighbours of a chunk are still L2-resident when the sweep comes back to them
    const double budget = 24.0 * 1024 * 1024;
    int zc = nz;
    while (zc > 1 && (double)nx * ny * zc * 576.0 > budget) {
        int d = zc - 1;
        while (d > 1 && nz % d != 0) d--;
        zc = d;
    }
    g.zc = zc;
    return g;
}

static int check_dims(gfb_ctx* ctx, int nx, int ny, int nz, int nt) {
    if (!ctx) return fail(nullptr, GFB_ERR_ARG, "null context");
    if (nx <= 0 || ny <= 0 || nz <= 0 || nt <= 0) return fail(ctx, GFB_ERR_ARG, "all lattice extents must be positive");
    if (nt % ctx->nslabs_total != 0) return fail(ctx, GFB_ERR_ARG, "NT must be divisible by the number of GPUs (t-slab decomposition)");
    if (ctx->nslabs_total > 1 && nt / ctx->nslabs_total < 2) return fail(ctx, GFB_ERR_ARG, "each t-slab needs at least 2 time-slices");
    if ((double)nx * ny * nz * 36.0 * (nt / ctx->nslabs_total + 2) >= 2147483647.0)
        return fail(ctx, GFB_ERR_ARG, "local slab too large for 32-bit element indexing (nslots*36*NX*NY*NZ must stay below 2^31)");
    return GFB_OK;
}

static int ensure_partial(gfb_ctx* ctx, Slab& s, size_t n) {
    if (s.partial_cap >= n) return GFB_OK;
    GFB_CUDA(ctx, cudaSetDevice(s.device));
    if (s.d_partial) GFB_CUDA(ctx, cudaFree(s.d_partial));
    s.d_partial = nullptr;
    GFB_CUDA(ctx, cudaMalloc(&s.d_partial, n * sizeof(double)));
    s.partial_cap = n;
    return GFB_OK;
}
static int ensure_staging(gfb_ctx* ctx, Slab& s, size_t bytes) {
    if (s.staging_cap >= bytes) return GFB_OK;
    GFB_CUDA(ctx, cudaSetDevice(s.device));
    if (s.d_staging) GFB_CUDA(ctx, cudaFree(s.d_staging));
    s.d_staging = nullptr;
    GFB_CUDA(ctx, cudaMalloc(&s.d_staging, bytes));
    s.staging_cap = bytes;
    return GFB_OK;
}

static int init_slab(gfb_ctx* ctx, Slab& s) {
    GFB_CUDA(ctx, cudaSetDevice(s.device));
    GFB_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    {
        // the halo stream outranks the compute stream so that the NCCL send/recv kernels get SM slots while a large
        // interior grid is resident
        int lo = 0, hi = 0;
        GFB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GFB_CUDA(ctx, cudaStreamCreateWithPriority(&s.comm_stream, cudaStreamNonBlocking, hi));
    }
    GFB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_a, cudaEventDisableTiming));
    GFB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_b, cudaEventDisableTiming));
    GFB_CUDA(ctx, cudaEventCreateWithFlags(&s.ev_c, cudaEventDisableTiming));
    GFB_CUDA(ctx, cudaEventCreate(&s.ev_tic));
    GFB_CUDA(ctx, cudaEventCreate(&s.ev_toc));
    GFB_CUDA(ctx, cudaMalloc(&s.d_result, 64 * sizeof(double)));
    GFB_CUDA(ctx, cudaMallocHost(&s.h_result, 64 * sizeof(double)));
    GFB_CUDA(ctx, cudaMalloc(&s.d_flags, 64));
    GFB_CUDA(ctx, cudaMemset(s.d_flags, 0, 64));
    return GFB_OK;
}

// ---- peer access for the halo exchange by peer stores -------------------------------------------------------------------
// One process, several GPUs: enable peer access between ring neighbours; the neighbours' buffers are ordinary pointers.
static bool setup_peers_local(gfb_ctx* ctx) {
    const int n = (int)ctx->slabs.size();
    if (n < 2) return false;
    for (int i = 0; i < n; i++) {
        const int nb[2] = {(i + n - 1) % n, (i + 1) % n};
        for (int k = 0; k < 2; k++) {
            const int a = ctx->slabs[i].device, b = ctx->slabs[nb[k]].device;
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); return false; }
            cudaSetDevice(a);
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
            cudaGetLastError();
        }
        ctx->slabs[i].peer_flag_prev = ctx->slabs[nb[0]].d_flags + 1;
        ctx->slabs[i].peer_flag_next = ctx->slabs[nb[1]].d_flags + 0;
    }
    return true;
}
// One process per GPU: all-gather the CUDA IPC handle of `mine` over NCCL and map the two ring neighbours' allocations.
// Collective: every rank calls it at the same point (buffer allocation is collective in the SPMD host programs).
static int exchange_ipc(gfb_ctx* ctx, void* mine, void** prev, void** next) {
    Slab& s = ctx->slabs[0];
    const int G = ctx->nslabs_total, r = s.index;
    GFB_CUDA(ctx, cudaSetDevice(s.device));
    cudaIpcMemHandle_t h;
    GFB_CUDA(ctx, cudaIpcGetMemHandle(&h, mine));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    GFB_CHECK(ensure_staging(ctx, s, (size_t)64 * (G + 1)));
    char* stg = reinterpret_cast<char*>(s.d_staging);
    GFB_CUDA(ctx, cudaMemcpyAsync(stg, &h, 64, cudaMemcpyHostToDevice, s.stream));
    GFB_NCCL(ctx, ncclAllGather(stg, stg + 64, 64, ncclChar, s.nccl, s.stream));
    std::vector<char> all((size_t)64 * G);
    GFB_CUDA(ctx, cudaMemcpyAsync(all.data(), stg + 64, (size_t)64 * G, cudaMemcpyDeviceToHost, s.stream));
    GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
    auto open = [&](int rank, void** out) -> int {
        const std::string key(all.data() + (size_t)64 * rank, 64);
        auto it = ctx->ipc_opened.find(key);
        if (it != ctx->ipc_opened.end()) { *out = it->second; return GFB_OK; }
        cudaIpcMemHandle_t hh;
        std::memcpy(&hh, key.data(), 64);
        void* ptr = nullptr;
        GFB_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, hh, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened[key] = ptr;
        *out = ptr;
        return GFB_OK;
    };
    GFB_CHECK(open((r + G - 1) % G, prev));
    GFB_CHECK(open((r + 1) % G, next));
    return GFB_OK;
}
static bool setup_peers_distributed(gfb_ctx* ctx) {
    Slab& s = ctx->slabs[0];
    void *pp = nullptr, *pn = nullptr;
    int st = exchange_ipc(ctx, s.d_flags, &pp, &pn);
    // every rank must take the same path: agree on the outcome (sum of failures) before trusting it
    int bad = (st != GFB_OK) ? 1 : 0;
    if (cudaMemcpy(s.d_result, &bad, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (ncclAllReduce(s.d_result, s.d_result, 1, ncclInt, ncclSum, s.nccl, s.stream) != ncclSuccess) return false;
    if (cudaMemcpyAsync(&bad, s.d_result, sizeof(int), cudaMemcpyDeviceToHost, s.stream) != cudaSuccess) return false;
    if (cudaStreamSynchronize(s.stream) != cudaSuccess) return false;
    if (bad) { cudaGetLastError(); return false; }
    s.peer_flag_prev = reinterpret_cast<unsigned*>(pp) + 1;
    s.peer_flag_next = reinterpret_cast<unsigned*>(pn) + 0;
    return true;
}
static bool peer_halo_wanted() {
    const char* e = getenv("GFB200_HALO");  // "nccl": keep the send/recv exchange; default "peer"
    return !(e && std::string(e) == "nccl");
}

// Sum one scalar per local slab (already in d_result[slot]) over all slabs of all ranks, in slab order.
static int gather_scalars(gfb_ctx* ctx, int nslots, double* out, bool take_max = false) {
    const int G = ctx->nslabs_total;
    for (int k = 0; k < nslots; k++) out[k] = 0.0;
    auto fold = [&](double& acc, double v) { acc = take_max ? (v > acc || v != v ? v : acc) : acc + v; };
    if (!ctx->distributed) {
        for (auto& s : ctx->slabs) {
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            GFB_CUDA(ctx, cudaMemcpyAsync(s.h_result, s.d_result, nslots * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        }
        for (auto& s : ctx->slabs) {
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
            for (int k = 0; k < nslots; k++) fold(out[k], s.h_result[k]);
        }
        return GFB_OK;
    }
    // one process per GPU: all-gather the per-rank values and add them in rank order on every rank
    Slab& s = ctx->slabs[0];
    GFB_CUDA(ctx, cudaSetDevice(s.device));
    if (nslots * G > 56) return fail(ctx, GFB_ERR_ARG, "too many ranks for the scalar gather area");
    GFB_NCCL(ctx, ncclAllGather(s.d_result, s.d_result + 8, nslots, ncclDouble, s.nccl, s.stream));
    GFB_CUDA(ctx, cudaMemcpyAsync(s.h_result, s.d_result + 8, (size_t)nslots * G * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
    for (int r = 0; r < G; r++)
        for (int k = 0; k < nslots; k++) fold(out[k], s.h_result[r * nslots + k]);
    return GFB_OK;
}

// t-halo exchange of `buf` (layout of gfb_gauge::d) on the compute streams: slot tloc <- next slab's
// slice 0, slot tloc+1 <- previous slab's last slice (set_wing_U!/set_halo! in the
// reference, gaugefields_4D_MPILattice.jl:497-507).
// overlapped = false: on the compute streams (in stream order with everything else).
// overlapped = true : on the high-priority halo streams, after the work already queued on the compute streams (the
//   boundary time-slices); ev_b marks completion and the caller makes the compute stream wait on it only AFTER it has
//   queued the interior slices, so the exchange over NVLink runs concurrently with the interior compute.
// su3 = true (the buffer holds unitary links: outputs of the fused update passes): only rows 0 and 1 of every link travel
//   (planes k = 0..5 of each direction are contiguous in the slice) and the receiver rebuilds row 2 = conj(row0 x row1) on the
//   halo stream: 42 instead of 63 planes per step.  Opt-in (GFB200_HALO_SU3=1): measured +3 % at 64^4 on 2 GPUs with 8 slices
//   each but -7 % on 8 GPUs (14 instead of 4 NCCL operations per step and one more kernel; the 8-GPU step is not
//   bandwidth-bound) -- profiles/r1_tmarch.md.
// ncclGroupStart/End pair that is closed on every return path
struct NcclGroup {
    bool open = false;
    ncclResult_t start() { ncclResult_t r = ncclGroupStart(); open = (r == ncclSuccess); return r; }
    ncclResult_t end() { open = false; return ncclGroupEnd(); }
    ~NcclGroup() { if (open) ncclGroupEnd(); }
};
static int exchange_halo_buffers(gfb_ctx* ctx, const gfb_gauge* g, const std::vector<double2*>& buf, bool overlapped, bool su3 = false) {
    const int G = ctx->nslabs_total;
    if (G == 1) return GFB_OK;
    static const bool allow_su3 = [] { const char* e = getenv("GFB200_HALO_SU3"); return e ? atoi(e) != 0 : false; }();
    su3 = su3 && allow_su3;
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t slice = g->slice_elems() * 2;  // doubles
    const size_t up_count = (size_t)27 * v3 * 2;  // the t+1 halo only needs the three spatial links
    if (overlapped) {
        for (auto& s : ctx->slabs) {
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            GFB_CUDA(ctx, cudaEventRecord(s.ev_a, s.stream));
            GFB_CUDA(ctx, cudaStreamWaitEvent(s.comm_stream, s.ev_a, 0));
        }
    }
    NcclGroup grp;
    GFB_NCCL(ctx, grp.start());
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        cudaStream_t st = overlapped ? s.comm_stream : s.stream;
        const int prev = (s.index + G - 1) % G, next = (s.index + 1) % G;
        double* base = reinterpret_cast<double*>(buf[i]);
        double* first = base;
        double* last = base + (size_t)(g->tloc - 1) * slice;
        double* up = base + (size_t)g->tloc * slice;
        double* dn = base + (size_t)(g->tloc + 1) * slice;
        if (!su3) {
            GFB_NCCL(ctx, ncclSend(first, up_count, ncclDouble, prev, s.nccl, st));
            GFB_NCCL(ctx, ncclSend(last, slice, ncclDouble, next, s.nccl, st));
            GFB_NCCL(ctx, ncclRecv(up, up_count, ncclDouble, next, s.nccl, st));
            GFB_NCCL(ctx, ncclRecv(dn, slice, ncclDouble, prev, s.nccl, st));
        } else {
            const size_t dir = 9 * v3 * 2, rows01 = 6 * v3 * 2;  // doubles per direction / per two rows of a direction
            for (int mu = 0; mu < 3; mu++) GFB_NCCL(ctx, ncclSend(first + mu * dir, rows01, ncclDouble, prev, s.nccl, st));
            for (int mu = 0; mu < 4; mu++) GFB_NCCL(ctx, ncclSend(last + mu * dir, rows01, ncclDouble, next, s.nccl, st));
            for (int mu = 0; mu < 3; mu++) GFB_NCCL(ctx, ncclRecv(up + mu * dir, rows01, ncclDouble, next, s.nccl, st));
            for (int mu = 0; mu < 4; mu++) GFB_NCCL(ctx, ncclRecv(dn + mu * dir, rows01, ncclDouble, prev, s.nccl, st));
        }
    }
    GFB_NCCL(ctx, grp.end());
    if (su3) {
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            Slab& s = ctx->slabs[i];
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            launch_complete_su3_rows(overlapped ? s.comm_stream : s.stream, make_geom(ctx, g->nx, g->ny, g->nz, g->nt, s.index), buf[i]);
            ctx->launches += 1;
            GFB_CUDA(ctx, cudaGetLastError());
        }
    }
    if (overlapped) {
        for (auto& s : ctx->slabs) {
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            GFB_CUDA(ctx, cudaEventRecord(s.ev_b, s.comm_stream));
        }
    }
    return GFB_OK;
}
// compute streams wait for the overlapped exchange started last
static int join_halo_exchange(gfb_ctx* ctx) {
    if (ctx->nslabs_total == 1) return GFB_OK;
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaStreamWaitEvent(s.stream, s.ev_b, 0));
    }
    return GFB_OK;
}
static int ensure_halo(gfb_gauge* g) {
    if (!g->has_halo || g->halo_valid) return GFB_OK;
    GFB_CHECK(exchange_halo_buffers(g->ctx, g, g->d, false));
    g->halo_valid = true;
    return GFB_OK;
}

// link buffers (one per local slab).  One process per GPU with peer halos: from the context's cache, and the ring neighbours'
// buffers of the same collective call are mapped (PoolBuf::peer_prev/next).
static int alloc_like(gfb_ctx* ctx, const gfb_gauge* g, std::vector<double2*>& out) {
    out.assign(ctx->slabs.size(), nullptr);
    const size_t bytes = g->elems_per_slab() * sizeof(double2);
    if (ctx->distributed && ctx->peer_ok) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[0].device));
        PoolBuf* pb = nullptr;
        for (auto& b : ctx->pool)
            if (!b.used && b.bytes == bytes) { pb = &b; break; }
        if (!pb) {
            PoolBuf nb;
            GFB_CUDA(ctx, cudaMalloc(&nb.p, bytes));
            nb.bytes = bytes;
            ctx->pool.push_back(nb);
            pb = &ctx->pool.back();
        }
        pb->used = true;
        void *pp = nullptr, *pn = nullptr;
        GFB_CHECK(exchange_ipc(ctx, pb->p, &pp, &pn));
        pb->peer_prev = reinterpret_cast<double2*>(pp);
        pb->peer_next = reinterpret_cast<double2*>(pn);
        out[0] = reinterpret_cast<double2*>(pb->p);
        return GFB_OK;
    }
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMalloc(&out[i], bytes));
    }
    return GFB_OK;
}
static void release_like(gfb_ctx* ctx, std::vector<double2*>& v) {
    for (size_t i = 0; i < v.size(); i++) {
        if (!v[i]) continue;
        bool pooled = false;
        for (auto& b : ctx->pool)
            if (b.p == v[i]) { b.used = false; pooled = true; break; }
        if (!pooled) { cudaSetDevice(ctx->slabs[i].device); cudaFree(v[i]); }
    }
    v.clear();
}
// the ring neighbours' copies of local buffer i of `bufs` (null when the halo goes over NCCL)
static void peers_of(gfb_ctx* ctx, const std::vector<double2*>& bufs, size_t i, double2** prev, double2** next) {
    *prev = *next = nullptr;
    if (!ctx->peer_ok) return;
    if (!ctx->distributed) {
        const size_t n = bufs.size();
        *prev = bufs[(i + n - 1) % n];
        *next = bufs[(i + 1) % n];
        return;
    }
    for (auto& b : ctx->pool)
        if (b.p == bufs[i]) { *prev = b.peer_prev; *next = b.peer_next; return; }
}

}  // namespace gfb

using namespace gfb;

// workspaces owned by a gauge handle (double buffer for the fused updates, flow field Z): members of gfb_gauge
typedef gfb_gauge gfb_gauge_ws;
static int get_ws(gfb_gauge* g, bool need_alt, bool need_z, gfb_gauge_ws** out) {
    gfb_ctx* ctx = g->ctx;
    if (need_alt && g->alt.empty()) GFB_CHECK(alloc_like(ctx, g, g->alt));
    if (need_z && g->z.empty()) {
        g->z.assign(ctx->slabs.size(), nullptr);
        size_t n = (size_t)g->tloc * 32 * g->nx * g->ny * g->nz;
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
            GFB_CUDA(ctx, cudaMalloc(&g->z[i], n * sizeof(double)));
        }
    }
    *out = g;
    return GFB_OK;
}
// device memory of a handle (the host struct stays)
static void release_gauge_memory(gfb_gauge* g) {
    gfb_ctx* ctx = g->ctx;
    if (!ctx) return;
    release_like(ctx, g->alt);
    for (size_t i = 0; i < g->wide.size(); i++) { cudaSetDevice(ctx->slabs[i].device); cudaFree(g->wide[i]); }
    g->wide.clear();
    for (size_t i = 0; i < g->z.size(); i++) { cudaSetDevice(ctx->slabs[i].device); cudaFree(g->z[i]); }
    g->z.clear();
    release_like(ctx, g->d);
}

static bool same_shape(const gfb_gauge* a, const gfb_gauge* b) { return a->ctx == b->ctx && a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->nt == b->nt; }
static bool same_shape(const gfb_gauge* a, const gfb_mom* b) { return a->ctx == b->ctx && a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->nt == b->nt; }
static bool same_shape(const gfb_mom* a, const gfb_mom* b) { return a->ctx == b->ctx && a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->nt == b->nt; }
static Geom geom_of(const gfb_gauge* g, size_t i) { return make_geom(g->ctx, g->nx, g->ny, g->nz, g->nt, g->ctx->slabs[i].index); }
static Geom geom_of(const gfb_mom* p, size_t i) { return make_geom(p->ctx, p->nx, p->ny, p->nz, p->nt, p->ctx->slabs[i].index); }
static int post_launch(gfb_ctx* ctx, int nlaunch = 1) {
    ctx->launches += nlaunch;
    GFB_CUDA(ctx, cudaGetLastError());
    return GFB_OK;
}

extern "C" {

static int build_wide(gfb_gauge* g, const std::vector<double2*>& src, Geom* wide_geom);
static int links_are_unitary(gfb_gauge* g, bool* yes);

int gfb_version(void) { return 100; }

const char* gfb_last_error(const gfb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int gfb_init(int ngpu, const int* devices, gfb_ctx** out) {
    if (!out) return fail(nullptr, GFB_ERR_ARG, "out is null");
    *out = nullptr;
    if (ngpu <= 0) return fail(nullptr, GFB_ERR_ARG, "ngpu must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(nullptr, GFB_ERR_NODEVICE, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (ngpu > ndev) return fail(nullptr, GFB_ERR_ARG, "more GPUs requested than visible");
    gfb_ctx* ctx = new gfb_ctx();
    ctx->nslabs_total = ngpu;
    ctx->distributed = false;
    ctx->slabs.resize(ngpu);
    std::vector<int> devs(ngpu);
    for (int i = 0; i < ngpu; i++) {
        devs[i] = devices ? devices[i] : i;
        ctx->slabs[i].device = devs[i];
        ctx->slabs[i].index = i;
    }
    for (auto& s : ctx->slabs) {
        int st = init_slab(ctx, s);
        if (st != GFB_OK) { g_init_error = ctx->err; delete ctx; return st; }
    }
    if (ngpu > 1) {
        std::vector<ncclComm_t> comms(ngpu);
        ncclResult_t r = ncclCommInitAll(comms.data(), ngpu, devs.data());
        if (r != ncclSuccess) { g_init_error = std::string("ncclCommInitAll: ") + ncclGetErrorString(r); delete ctx; return GFB_ERR_NCCL; }
        for (int i = 0; i < ngpu; i++) ctx->slabs[i].nccl = comms[i];
        ctx->peer_ok = peer_halo_wanted() && setup_peers_local(ctx);
    }
    *out = ctx;
    return GFB_OK;
}

int gfb_nccl_unique_id(char* out128) {
    if (!out128) return fail(nullptr, GFB_ERR_ARG, "out128 is null");
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, GFB_ERR_NCCL, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(out128, &id, 128);
    return GFB_OK;
}

int gfb_init_rank(int rank, int nranks, const char* id128, int device, gfb_ctx** out) {
    if (!out) return fail(nullptr, GFB_ERR_ARG, "out is null");
    *out = nullptr;
    if (nranks <= 0 || rank < 0 || rank >= nranks) return fail(nullptr, GFB_ERR_ARG, "invalid rank/nranks");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(nullptr, GFB_ERR_NODEVICE, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, GFB_ERR_ARG, "invalid device index");
    gfb_ctx* ctx = new gfb_ctx();
    ctx->nslabs_total = nranks;
    ctx->distributed = nranks > 1;
    ctx->slabs.resize(1);
    ctx->slabs[0].device = device;
    ctx->slabs[0].index = rank;
    int st = init_slab(ctx, ctx->slabs[0]);
    if (st != GFB_OK) { g_init_error = ctx->err; delete ctx; return st; }
    if (nranks > 1) {
        if (!id128) { delete ctx; return fail(nullptr, GFB_ERR_ARG, "id128 is null"); }
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ncclResult_t r = ncclCommInitRank(&ctx->slabs[0].nccl, nranks, id, rank);
        if (r != ncclSuccess) { g_init_error = std::string("ncclCommInitRank: ") + ncclGetErrorString(r); delete ctx; return GFB_ERR_NCCL; }
        ctx->peer_ok = peer_halo_wanted() && setup_peers_distributed(ctx);
        ctx->err.clear();
    }
    *out = ctx;
    return GFB_OK;
}

static void release_field_memory(gfb_field* f);
static void release_mom_memory(gfb_mom* p);

int gfb_finalize(gfb_ctx* ctx) {
    if (!ctx) return GFB_OK;
    for (auto& s : ctx->slabs) {
        cudaSetDevice(s.device);
        cudaDeviceSynchronize();
    }
    // handles that outlive the context (host finalizers run in any order): release their device memory now and orphan them
    for (gfb_field* f : ctx->fields) { release_field_memory(f); f->ctx = nullptr; f->parent = nullptr; }
    for (gfb_gauge* g : ctx->gauges) { release_gauge_memory(g); g->ctx = nullptr; }
    for (gfb_mom* p : ctx->moms) { release_mom_memory(p); p->ctx = nullptr; }
    if (ctx->distributed && ctx->slabs[0].nccl && !ctx->ipc_opened.empty()) {
        // nobody may free an exported buffer while a neighbour still maps it: close our mappings, meet, then free
        Slab& s = ctx->slabs[0];
        cudaSetDevice(s.device);
        for (auto& kv : ctx->ipc_opened) cudaIpcCloseMemHandle(kv.second);
        ctx->ipc_opened.clear();
        if (ncclAllReduce(s.d_result, s.d_result, 1, ncclInt, ncclSum, s.nccl, s.stream) == ncclSuccess) cudaStreamSynchronize(s.stream);
        cudaGetLastError();
    }
    for (auto& b : ctx->pool) { cudaSetDevice(ctx->slabs[0].device); cudaFree(b.p); }
    ctx->pool.clear();
    for (auto& s : ctx->slabs) {
        cudaSetDevice(s.device);
        if (s.nccl) ncclCommDestroy(s.nccl);
        if (s.d_flags) cudaFree(s.d_flags);
        if (s.d_partial) cudaFree(s.d_partial);
        if (s.d_result) cudaFree(s.d_result);
        if (s.h_result) cudaFreeHost(s.h_result);
        if (s.d_staging) cudaFree(s.d_staging);
        cudaEventDestroy(s.ev_a); cudaEventDestroy(s.ev_b); cudaEventDestroy(s.ev_c);
        cudaEventDestroy(s.ev_tic); cudaEventDestroy(s.ev_toc);
        cudaStreamDestroy(s.stream); cudaStreamDestroy(s.comm_stream);
    }
    delete ctx;
    return GFB_OK;
}

int gfb_sync(gfb_ctx* ctx) {
    if (!ctx) return fail(nullptr, GFB_ERR_ARG, "null context");
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
        GFB_CUDA(ctx, cudaStreamSynchronize(s.comm_stream));
        if (ctx->peer_ok) {
            unsigned timed_out = 0;
            GFB_CUDA(ctx, cudaMemcpy(&timed_out, s.d_flags + 2, sizeof(unsigned), cudaMemcpyDeviceToHost));
            if (timed_out) return fail(ctx, GFB_ERR_NCCL, "halo exchange: a ring neighbour did not complete a pass within 20 s");
        }
    }
    return GFB_OK;
}

int gfb_num_slabs(const gfb_ctx* ctx, int* local, int* total) {
    if (!ctx) return GFB_ERR_ARG;
    if (local) *local = (int)ctx->slabs.size();
    if (total) *total = ctx->nslabs_total;
    return GFB_OK;
}

int gfb_timer_tic(gfb_ctx* ctx) {
    if (!ctx) return GFB_ERR_ARG;
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaEventRecord(s.ev_tic, s.stream));
    }
    return GFB_OK;
}
int gfb_timer_toc(gfb_ctx* ctx, double* ms) {
    if (!ctx || !ms) return GFB_ERR_ARG;
    double worst = 0.0;
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaEventRecord(s.ev_toc, s.stream));
    }
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaEventSynchronize(s.ev_toc));
        float f = 0.f;
        GFB_CUDA(ctx, cudaEventElapsedTime(&f, s.ev_tic, s.ev_toc));
        if (f > worst) worst = f;
    }
    *ms = worst;
    return GFB_OK;
}
int gfb_kernel_launches(const gfb_ctx* ctx, long long* count) {
    if (!ctx || !count) return GFB_ERR_ARG;
    *count = ctx->launches;
    return GFB_OK;
}
int gfb_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return GFB_ERR_ARG;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) return fail(nullptr, GFB_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    return GFB_OK;
}
int gfb_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
    return GFB_OK;
}

// ---- fields -----------------------------------------------------------------------------------
int gfb_gauge_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_gauge** out) {
    if (!out) return fail(ctx, GFB_ERR_ARG, "out is null");
    *out = nullptr;
    GFB_CHECK(check_dims(ctx, nx, ny, nz, nt));
    gfb_gauge* g = new gfb_gauge();
    g->ctx = ctx; g->nx = nx; g->ny = ny; g->nz = nz; g->nt = nt;
    g->tloc = nt / ctx->nslabs_total;
    g->has_halo = ctx->nslabs_total > 1;
    g->halo_valid = false;
    int st = alloc_like(ctx, g, g->d);
    if (st != GFB_OK) { release_like(ctx, g->d); delete g; return st; }
    ctx->gauges.insert(g);
    *out = g;
    return GFB_OK;
}
int gfb_gauge_free(gfb_gauge* g) {
    if (!g) return GFB_OK;
    if (g->ctx) {  // null: the context was finalized first and has already released the device memory
        // views of this configuration must not dangle
        for (gfb_field* f : g->ctx->fields)
            if (f->parent == g) { f->parent = nullptr; f->d.clear(); }
        release_gauge_memory(g);
        g->ctx->gauges.erase(g);
    }
    delete g;
    return GFB_OK;
}
int gfb_mom_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_mom** out) {
    if (!out) return fail(ctx, GFB_ERR_ARG, "out is null");
    *out = nullptr;
    GFB_CHECK(check_dims(ctx, nx, ny, nz, nt));
    gfb_mom* p = new gfb_mom();
    p->ctx = ctx; p->nx = nx; p->ny = ny; p->nz = nz; p->nt = nt;
    p->tloc = nt / ctx->nslabs_total;
    p->d.assign(ctx->slabs.size(), nullptr);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMalloc(&p->d[i], p->elems_per_slab() * sizeof(double)));
        GFB_CUDA(ctx, cudaMemsetAsync(p->d[i], 0, p->elems_per_slab() * sizeof(double), ctx->slabs[i].stream));
    }
    ctx->moms.insert(p);
    *out = p;
    return GFB_OK;
}
static void release_mom_memory(gfb_mom* p) {
    if (!p->ctx) return;
    for (size_t i = 0; i < p->d.size(); i++) { cudaSetDevice(p->ctx->slabs[i].device); cudaFree(p->d[i]); }
    p->d.clear();
}
int gfb_mom_free(gfb_mom* p) {
    if (!p) return GFB_OK;
    if (p->ctx) { release_mom_memory(p); p->ctx->moms.erase(p); }
    delete p;
    return GFB_OK;
}

int gfb_gauge_upload(gfb_gauge* g, int mu, const double* host) {
    if (!g || !host) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t bytes = v3 * g->tloc * 9 * sizeof(double2);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(g, i);
        const double* src = host + (size_t)geo.t0 * v3 * 18;
        GFB_CUDA(ctx, cudaMemcpyAsync(s.d_staging, src, bytes, cudaMemcpyHostToDevice, s.stream));
        launch_links_from_host_layout(s.stream, geo, mu, reinterpret_cast<const double2*>(s.d_staging), g->d[i]);
        GFB_CHECK(post_launch(ctx));
    }
    g->halo_valid = false;
    g->unitary = -1;  // checked on first use by a fused pass (links_are_unitary)
    return GFB_OK;
}
int gfb_gauge_download(const gfb_gauge* g, int mu, double* host) {
    if (!g || !host) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t bytes = v3 * g->tloc * 9 * sizeof(double2);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(g, i);
        launch_links_to_host_layout(s.stream, geo, mu, g->d[i], reinterpret_cast<double2*>(s.d_staging));
        GFB_CHECK(post_launch(ctx));
        double* dst = host + (size_t)geo.t0 * v3 * 18;
        GFB_CUDA(ctx, cudaMemcpyAsync(dst, s.d_staging, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ctx->slabs) { GFB_CUDA(ctx, cudaSetDevice(s.device)); GFB_CUDA(ctx, cudaStreamSynchronize(s.stream)); }
    return GFB_OK;
}
int gfb_gauge_upload_ildg(gfb_gauge* g, const void* payload, int precision) {
    if (!g || !payload) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (precision != 32 && precision != 64) return fail(ctx, GFB_ERR_ARG, "ILDG precision must be 32 or 64");
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t site_bytes = (size_t)4 * 9 * 2 * (precision / 8);
    const size_t bytes = v3 * g->tloc * site_bytes;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(g, i);
        const char* src = reinterpret_cast<const char*>(payload) + (size_t)geo.t0 * v3 * site_bytes;
        GFB_CUDA(ctx, cudaMemcpyAsync(s.d_staging, src, bytes, cudaMemcpyHostToDevice, s.stream));
        launch_links_from_ildg(s.stream, geo, precision, s.d_staging, g->d[i]);
        GFB_CHECK(post_launch(ctx));
    }
    g->halo_valid = false;
    g->unitary = -1;  // a 32-bit file is unitary to 1e-7 only: the passes then use full 3x3 products unless gfb_reunitarize is called
    return GFB_OK;
}
int gfb_gauge_download_ildg(const gfb_gauge* g, void* payload, int precision) {
    if (!g || !payload) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (precision != 32 && precision != 64) return fail(ctx, GFB_ERR_ARG, "ILDG precision must be 32 or 64");
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t site_bytes = (size_t)4 * 9 * 2 * (precision / 8);
    const size_t bytes = v3 * g->tloc * site_bytes;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(g, i);
        launch_links_to_ildg(s.stream, geo, precision, g->d[i], s.d_staging);
        GFB_CHECK(post_launch(ctx));
        char* dst = reinterpret_cast<char*>(payload) + (size_t)geo.t0 * v3 * site_bytes;
        GFB_CUDA(ctx, cudaMemcpyAsync(dst, s.d_staging, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ctx->slabs) { GFB_CUDA(ctx, cudaSetDevice(s.device)); GFB_CUDA(ctx, cudaStreamSynchronize(s.stream)); }
    return GFB_OK;
}
int gfb_mom_upload(gfb_mom* p, int mu, const double* host) {
    if (!p || !host) return fail(p ? p->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    const size_t v3 = (size_t)p->nx * p->ny * p->nz;
    const size_t bytes = v3 * p->tloc * 8 * sizeof(double);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(p, i);
        const double* src = host + (size_t)geo.t0 * v3 * 8;
        GFB_CUDA(ctx, cudaMemcpyAsync(s.d_staging, src, bytes, cudaMemcpyHostToDevice, s.stream));
        launch_mom_from_host_layout(s.stream, geo, mu, reinterpret_cast<const double*>(s.d_staging), p->d[i]);
        GFB_CHECK(post_launch(ctx));
    }
    return GFB_OK;
}
int gfb_mom_download(const gfb_mom* p, int mu, double* host) {
    if (!p || !host) return fail(p ? p->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    const size_t v3 = (size_t)p->nx * p->ny * p->nz;
    const size_t bytes = v3 * p->tloc * 8 * sizeof(double);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(p, i);
        launch_mom_to_host_layout(s.stream, geo, mu, p->d[i], reinterpret_cast<double*>(s.d_staging));
        GFB_CHECK(post_launch(ctx));
        double* dst = host + (size_t)geo.t0 * v3 * 8;
        GFB_CUDA(ctx, cudaMemcpyAsync(dst, s.d_staging, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ctx->slabs) { GFB_CUDA(ctx, cudaSetDevice(s.device)); GFB_CUDA(ctx, cudaStreamSynchronize(s.stream)); }
    return GFB_OK;
}

int gfb_gauge_copy(gfb_gauge* dst, const gfb_gauge* src) {
    if (!dst || !src) return fail(dst ? dst->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = dst->ctx;
    if (!same_shape(dst, src)) return fail(ctx, GFB_ERR_ARG, "destination and source lattice sizes differ");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMemcpyAsync(dst->d[i], src->d[i], src->elems_per_slab() * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->slabs[i].stream));
    }
    dst->halo_valid = src->halo_valid;
    dst->unitary = src->unitary;
    return GFB_OK;
}
int gfb_mom_copy(gfb_mom* dst, const gfb_mom* src) {
    if (!dst || !src) return fail(dst ? dst->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = dst->ctx;
    if (!same_shape(dst, src)) return fail(ctx, GFB_ERR_ARG, "destination and source lattice sizes differ");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMemcpyAsync(dst->d[i], src->d[i], src->elems_per_slab() * sizeof(double), cudaMemcpyDeviceToDevice, ctx->slabs[i].stream));
    }
    return GFB_OK;
}
int gfb_mom_zero(gfb_mom* p) {
    if (!p) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMemsetAsync(p->d[i], 0, p->elems_per_slab() * sizeof(double), ctx->slabs[i].stream));
    }
    return GFB_OK;
}
int gfb_mom_axpy(gfb_mom* p, double t, const gfb_mom* f) {
    if (!p || !f) return fail(p ? p->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    if (!same_shape(p, f)) return fail(ctx, GFB_ERR_ARG, "momentum fields differ in shape");
    if (!std::isfinite(t)) return fail(ctx, GFB_ERR_ARG, "the step size must be finite");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_axpy(ctx->slabs[i].stream, p->d[i], t, f->d[i], p->elems_per_slab());
        GFB_CHECK(post_launch(ctx));
    }
    return GFB_OK;
}

// add_U!(P[mu], t, F[mu]) on ONE direction (the loop body of update_momenta!, molecular_dynamics.jl:580-582): the momenta of a
// direction are the 8 coefficient planes [mu*8, mu*8+8) of every time-slice
int gfb_mom_axpy_dir(gfb_mom* p, int mu, double t, const gfb_mom* f) {
    if (!p || !f) return fail(p ? p->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    if (!same_shape(p, f)) return fail(ctx, GFB_ERR_ARG, "momentum fields differ in shape");
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    if (!std::isfinite(t)) return fail(ctx, GFB_ERR_ARG, "the step size must be finite");
    const size_t v3 = (size_t)p->nx * p->ny * p->nz;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        for (int tt = 0; tt < p->tloc; tt++) {
            const size_t off = ((size_t)tt * 32 + (size_t)mu * 8) * v3;
            launch_axpy(ctx->slabs[i].stream, p->d[i] + off, t, f->d[i] + off, 8 * v3);
        }
        GFB_CHECK(post_launch(ctx, p->tloc));
    }
    return GFB_OK;
}

// ---- initial fields -----------------------------------------------------------------------------
int gfb_set_cold(gfb_gauge* g) {
    if (!g) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_set_cold(ctx->slabs[i].stream, geom_of(g, i), g->d[i]);
        GFB_CHECK(post_launch(ctx));
    }
    g->halo_valid = false;
    g->unitary = 1;
    return GFB_OK;
}
int gfb_set_hot(gfb_gauge* g, uint64_t seed, int rng_alg) {
    if (!g) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (rng_alg != GFB_PHILOX4X32) return fail(ctx, GFB_ERR_ARG, "only the Philox4x32 site RNG is implemented");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_set_hot(ctx->slabs[i].stream, geom_of(g, i), g->d[i], seed);
        GFB_CHECK(post_launch(ctx));
    }
    g->halo_valid = false;
    g->unitary = 1;
    return GFB_OK;
}
int gfb_gaussian_momenta(gfb_mom* p, uint64_t seed, uint64_t sweep, double sigma, int rng_alg) {
    if (!p) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = p->ctx;
    if (rng_alg != GFB_PHILOX4X32) return fail(ctx, GFB_ERR_ARG, "only the Philox4x32 site RNG is implemented");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_gaussian(ctx->slabs[i].stream, geom_of(p, i), p->d[i], seed, sweep, sigma);
        GFB_CHECK(post_launch(ctx));
    }
    return GFB_OK;
}
int gfb_reunitarize(gfb_gauge* g) {
    if (!g) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_reunitarize(ctx->slabs[i].stream, geom_of(g, i), g->d[i]);
        GFB_CHECK(post_launch(ctx));
    }
    g->halo_valid = false;
    g->unitary = 1;
    return GFB_OK;
}

// ---- observables ----------------------------------------------------------------------------------
static int plaquette_partial(gfb_gauge* g, int slot) {
    gfb_ctx* ctx = g->ctx;
    GFB_CHECK(ensure_halo(g));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        Geom geo = geom_of(g, i);
        GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geo) * 2 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_plaquette(s.stream, geo, g->d[i], s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result + slot);
        GFB_CHECK(post_launch(ctx, 2));
    }
    return GFB_OK;
}
static int kinetic_partial(gfb_mom* p, int slot) {
    gfb_ctx* ctx = p->ctx;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_partial(ctx, s, (size_t)sumsq_blocks(p->elems_per_slab()) + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_sumsq(s.stream, p->d[i], p->elems_per_slab(), s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result + slot);
        GFB_CHECK(post_launch(ctx, 2));
    }
    return GFB_OK;
}

int gfb_plaquette_sum(gfb_gauge* g, double* out) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    GFB_CHECK(plaquette_partial(g, 0));
    return gather_scalars(g->ctx, 1, out);
}
int gfb_wilson_action(gfb_gauge* g, double beta, double* out) {
    double s = 0.0;
    GFB_CHECK(gfb_plaquette_sum(g, &s));
    *out = beta * s;
    return GFB_OK;
}
int gfb_kinetic(gfb_mom* p, double* out) {
    if (!p || !out) return fail(p ? p->ctx : nullptr, GFB_ERR_ARG, "null argument");
    GFB_CHECK(kinetic_partial(p, 0));
    return gather_scalars(p->ctx, 1, out);
}
int gfb_hamiltonian(gfb_gauge* g, gfb_mom* p, double beta, double* out) {
    if (!g || !p || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, p)) return fail(g->ctx, GFB_ERR_ARG, "U and p must have the same lattice");
    GFB_CHECK(plaquette_partial(g, 0));
    // the kinetic pass reuses d_partial: stream order keeps the two reductions apart
    GFB_CHECK(kinetic_partial(p, 1));
    double v[2];
    GFB_CHECK(gather_scalars(g->ctx, 2, v));
    *out = -(beta / 3.0) * v[0] + 0.5 * v[1];
    return GFB_OK;
}
// [sum_{x, mu<nu} Re tr plaquette, sum_x Re tr of the 12 "rectangular" loops] (make_loops_fromname, wilsonloops.jl:233-245)
int gfb_loop_sums(gfb_gauge* g, double* out2) {
    if (!g || !out2) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    Geom gw = geom_of(g, 0);
    const bool wide = ctx->nslabs_total > 1;
    if (wide) GFB_CHECK(build_wide(g, g->d, &gw));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geom_of(g, i)) * 2 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_loop_sums(s.stream, gw, wide ? 2 : 0, g->tloc, wide ? g->wide[i] : g->d[i], s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result);
        launch_final_reduce(s.stream, s.d_partial + nb, nb, s.d_result + 1);
        GFB_CHECK(post_launch(ctx, 3));
    }
    return gather_scalars(ctx, 2, out2);
}

// topological_charge_density / topological_charge (src/AbstractGaugefields.jl:1447-1490): method 0 plaquette, 1 clover,
// 2 improved (5/3 clover - 1/12 rectangle).  The density lands in a device buffer in host order (x fastest, t slowest).
static int topological_density(gfb_gauge* g, int method, std::vector<double*>& dens) {
    gfb_ctx* ctx = g->ctx;
    if (method < 0 || method > 2) return fail(ctx, GFB_ERR_ARG, "supported topological charge methods are plaquette (0), clover (1) and improved (2)");
    // the clover leaves reach one site back and forward in two directions, the rectangle loops two sites: on t-slabs all three
    // methods read the wide copy of the slab
    Geom gw = geom_of(g, 0);
    const bool wide = ctx->nslabs_total > 1;
    if (wide) GFB_CHECK(build_wide(g, g->d, &gw));
    const int tb = wide ? 2 : 0;
    dens.assign(ctx->slabs.size(), nullptr);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        const size_t n = (size_t)gw.v3 * g->tloc;
        GFB_CHECK(ensure_staging(ctx, s, n * sizeof(double)));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        dens[i] = reinterpret_cast<double*>(s.d_staging);
        const double2* u = wide ? g->wide[i] : g->d[i];
        if (method == 2) {
            launch_topological_density(s.stream, gw, tb, g->tloc, tb, u, dens[i], 1, 5.0 / 3.0, false);
            launch_topological_density(s.stream, gw, tb, g->tloc, tb, u, dens[i], 2, -1.0 / 12.0, true);
            GFB_CHECK(post_launch(ctx, 2));
        } else {
            launch_topological_density(s.stream, gw, tb, g->tloc, tb, u, dens[i], method, 1.0, false);
            GFB_CHECK(post_launch(ctx));
        }
    }
    return GFB_OK;
}
int gfb_topological_charge(gfb_gauge* g, int method, double* out) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    std::vector<double*> dens;
    GFB_CHECK(topological_density(g, method, dens));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        Geom geo = geom_of(g, i);
        GFB_CHECK(ensure_partial(ctx, s, 1024 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_sum_plain(s.stream, dens[i], (size_t)geo.v3 * geo.tloc, s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result);
        GFB_CHECK(post_launch(ctx, 2));
    }
    return gather_scalars(ctx, 1, out);
}
// host_density: NX*NY*NZ*NT doubles, x fastest (the reference's density[ix, iy, iz, it]); a rank of a one-process-per-GPU
// context passes the global array and fills only its own time-slices
int gfb_topological_charge_density(gfb_gauge* g, int method, double* host_density) {
    if (!g || !host_density) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    std::vector<double*> dens;
    GFB_CHECK(topological_density(g, method, dens));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        Geom geo = geom_of(g, i);
        const size_t n = (size_t)geo.v3 * geo.tloc;
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaMemcpyAsync(host_density + (size_t)geo.v3 * geo.t0, dens[i], n * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
    }
    return GFB_OK;
}

int gfb_energy_density(gfb_gauge* g, int kind, double* out) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    const double V = (double)g->nx * g->ny * g->nz * g->nt;
    if (kind == GFB_E_PLAQUETTE) {
        double s = 0.0;
        GFB_CHECK(gfb_plaquette_sum(g, &s));
        *out = 2.0 * (18.0 - s / V);
        return GFB_OK;
    }
    if (kind != GFB_E_CLOVER) return fail(ctx, GFB_ERR_ARG, "unknown energy-density kind");
    GFB_CHECK(ensure_halo(g));
    bool unitary = true;
    GFB_CHECK(links_are_unitary(g, &unitary));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        Geom geo = geom_of(g, i);
        GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geo) * 2 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_clover_energy(s.stream, geo, g->d[i], s.d_partial, &nb, !unitary);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result);
        GFB_CHECK(post_launch(ctx, 2));
    }
    double v = 0.0;
    GFB_CHECK(gather_scalars(ctx, 1, &v));
    *out = v / (V * 16.0);
    return GFB_OK;
}
int gfb_polyakov(gfb_gauge* g, double* out2) {
    if (!g || !out2) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    const int G = ctx->nslabs_total;
    const size_t v3 = (size_t)g->nx * g->ny * g->nz;
    const size_t field = 9 * v3;  // double2 elements of one running-product field
    // slab r continues the product of slabs 0..r-1 (handed over with NCCL send/recv) and the last slab takes the trace
    for (int r = 0; r < G; r++) {
        int li = -1, lprev = -1;
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            if (ctx->slabs[i].index == r) li = (int)i;
            if (ctx->slabs[i].index == r - 1) lprev = (int)i;
        }
        if (li >= 0) GFB_CHECK(ensure_staging(ctx, ctx->slabs[li], 2 * field * sizeof(double2)));
        if (r > 0 && (li >= 0 || lprev >= 0)) {
            NcclGroup grp;
            GFB_NCCL(ctx, grp.start());
            if (lprev >= 0) {
                Slab& sp = ctx->slabs[lprev];
                GFB_NCCL(ctx, ncclSend(reinterpret_cast<double2*>(sp.d_staging) + field, field * 2, ncclDouble, r, sp.nccl, sp.stream));
            }
            if (li >= 0) {
                Slab& sr = ctx->slabs[li];
                GFB_NCCL(ctx, ncclRecv(sr.d_staging, field * 2, ncclDouble, r - 1, sr.nccl, sr.stream));
            }
            GFB_NCCL(ctx, grp.end());
        }
        if (li < 0) continue;
        Slab& s = ctx->slabs[li];
        Geom geo = geom_of(g, li);
        GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geo) * 2 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        double2* const stg = reinterpret_cast<double2*>(s.d_staging);
        int nb = 0;
        launch_polyakov(s.stream, geo, g->d[li], r > 0 ? stg : nullptr, r < G - 1 ? stg + field : nullptr, s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result);  // zero on every slab but the last
        launch_final_reduce(s.stream, s.d_partial + nb, nb, s.d_result + 1);
        GFB_CHECK(post_launch(ctx, 3));
    }
    double v[2];
    GFB_CHECK(gather_scalars(ctx, 2, v));
    out2[0] = v[0] / v3;
    out2[1] = v[1] / v3;
    return GFB_OK;
}

// ---- updates -----------------------------------------------------------------------------------------
// Are all links SU(3) to 1e-12?  The fused kernels form staples from two rows of their unitary factors (su3.cuh); a
// configuration that came in through upload / ILDG / a written view is checked once (one read of U), and if it is not unitary
// every pass on it uses full 3x3 products, like the reference's staples, which hold for any matrices.
static int links_are_unitary(gfb_gauge* g, bool* yes) {
    gfb_ctx* ctx = g->ctx;
    if (g->unitary < 0) {
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            Slab& s = ctx->slabs[i];
            Geom geo = geom_of(g, i);
            GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geo) * 2 + 16));
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            int nb = 0;
            launch_unitarity_defect(s.stream, geo, g->d[i], s.d_partial, &nb);
            launch_final_max(s.stream, s.d_partial, nb, s.d_result);
            GFB_CHECK(post_launch(ctx, 2));
        }
        double worst = 0.0;
        GFB_CHECK(gather_scalars(ctx, 1, &worst, true));
        g->unitary = (worst <= 1e-12) ? 1 : 0;
    }
    *yes = g->unitary == 1;
    return GFB_OK;
}

// Wide copy of every local slab of `src` for kernels that reach two slices away (rectangle loops): tloc + 4 slices in natural t
// order -- [0, 2) the previous slab's last two slices, [2, tloc + 2) the slab itself, [tloc + 2, tloc + 4) the next slab's first
// two -- owned by the configuration handle.  One device copy plus one NCCL send/recv pair per neighbour (two slices each).
static int build_wide(gfb_gauge* g, const std::vector<double2*>& src, Geom* wide_geom) {
    gfb_ctx* ctx = g->ctx;
    const int G = ctx->nslabs_total;
    if (g->tloc < 2) return fail(ctx, GFB_ERR_ARG, "rectangle terms need at least two time-slices per GPU");
    const size_t slice = g->slice_elems();  // double2 per slice
    if (g->wide.empty()) {
        g->wide.assign(ctx->slabs.size(), nullptr);
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
            GFB_CUDA(ctx, cudaMalloc(&g->wide[i], (size_t)(g->tloc + 4) * slice * sizeof(double2)));
        }
    }
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaMemcpyAsync(g->wide[i] + 2 * slice, src[i], (size_t)g->tloc * slice * sizeof(double2), cudaMemcpyDeviceToDevice, s.stream));
    }
    NcclGroup grp;
    GFB_NCCL(ctx, grp.start());
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        const int prev = (s.index + G - 1) % G, next = (s.index + 1) % G;
        const size_t two = 2 * slice * 2;  // doubles in two slices
        double* own = reinterpret_cast<double*>(src[i]);
        double* w = reinterpret_cast<double*>(g->wide[i]);
        GFB_NCCL(ctx, ncclSend(own, two, ncclDouble, prev, s.nccl, s.stream));                                         // my first two -> prev's top halo
        GFB_NCCL(ctx, ncclSend(own + (size_t)(g->tloc - 2) * slice * 2, two, ncclDouble, next, s.nccl, s.stream));     // my last two -> next's bottom halo
        GFB_NCCL(ctx, ncclRecv(w + (size_t)(g->tloc + 2) * slice * 2, two, ncclDouble, next, s.nccl, s.stream));
        GFB_NCCL(ctx, ncclRecv(w, two, ncclDouble, prev, s.nccl, s.stream));
    }
    GFB_NCCL(ctx, grp.end());
    Geom gw = geom_of(g, 0);
    gw.tloc = g->tloc + 4;
    gw.nslots = g->tloc + 4;
    gw.t_up_wrap = 0;            // never taken: the computed slices [2, tloc + 2) reach at most two slices away
    gw.t_dn_wrap = g->tloc + 3;
    *wide_geom = gw;
    return GFB_OK;
}

// one fused pass over all local slabs: Z' = a*TA(U V^dag) + b*Z ; Uout = exp(c Z') Uin.
// With several slabs and an output link field, on return (in stream order) uout's halo slots are valid:
//   peer halos (default): the kernels store their boundary slices into the neighbours' halo slots themselves (NVLink peer
//     stores, tmarch.cu / kernels.cu); one launch per slab, then one single-thread kernel that signals both ring neighbours and
//     waits for their signal.  No pack, no send/recv kernel, no event chain, and the slab is not split.
//   NCCL (GFB200_HALO=nccl or no peer access): boundary slices -> send/recv on the halo stream || interior slices -> join.
static int fused_pass(gfb_gauge* g, const std::vector<double2*>& uin, const std::vector<double2*>* uout, const std::vector<double*>* zin,
                      const std::vector<double*>* zout, FusedArgs fa) {
    gfb_ctx* ctx = g->ctx;
    if (fa.c_rect != 0.0) {
        {
            bool su3 = true;
            GFB_CHECK(links_are_unitary(g, &su3));
            fa.full3 = !su3;
        }
        // rectangle staples reach two sites away.  One slab: plain periodic addressing.  t-slabs: the one-slice halo slots are
        // not enough, so the pass runs on a wide copy of the slab (build_wide) and writes into the slab's own slices; the
        // output's one-slice halo is NOT exchanged here (callers clear halo_valid).
        if (ctx->nslabs_total == 1) {
            Slab& s = ctx->slabs[0];
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            launch_force_general(s.stream, geom_of(g, 0), 0, g->tloc, 0, uin[0], uout ? (*uout)[0] : nullptr, zin ? (*zin)[0] : nullptr, zout ? (*zout)[0] : nullptr, fa);
            return post_launch(ctx);
        }
        Geom gw;
        GFB_CHECK(build_wide(g, uin, &gw));
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            Slab& s = ctx->slabs[i];
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            launch_force_general(s.stream, gw, 2, g->tloc, 2, g->wide[i], uout ? (*uout)[i] : nullptr, zin ? (*zin)[i] : nullptr, zout ? (*zout)[i] : nullptr, fa);
            GFB_CHECK(post_launch(ctx));
        }
        return GFB_OK;
    }
    fa.a *= fa.c_plaq;  // plaquette-only actions: the coefficient folds into the force scale and the Wilson kernels run
    fa.c_plaq = 1.0;
    bool unitary = true;
    GFB_CHECK(links_are_unitary(g, &unitary));
    fa.full3 = !unitary;
    const bool slabs = ctx->nslabs_total > 1 && uout != nullptr;
    auto launch = [&](size_t i, int t0, int tc, int stride, const FusedArgs& f) -> int {
        Slab& s = ctx->slabs[i];
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(g, i);
        geo.t_stride = stride;
        if (tc <= 0) return GFB_OK;
        launch_force_fused(s.stream, geo, t0, tc, uin[i], uout ? (*uout)[i] : nullptr, zin ? (*zin)[i] : nullptr, zout ? (*zout)[i] : nullptr, f);
        return post_launch(ctx);
    };
    if (!slabs) {
        for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 0, g->tloc, 1, fa));
        return GFB_OK;
    }
    if (ctx->peer_ok) {
        const unsigned serial = ++ctx->pass_serial;
        for (size_t i = 0; i < ctx->slabs.size(); i++) {
            FusedArgs f = fa;
            peers_of(ctx, *uout, i, &f.peer_prev, &f.peer_next);
            if (!f.peer_prev || !f.peer_next) return fail(ctx, GFB_ERR_ARG, "internal: output buffer without peer mappings");
            GFB_CHECK(launch(i, 0, g->tloc, 1, f));
        }
        static const bool nowait = getenv("GFB200_HALO_NOWAIT") != nullptr;  // timing experiments only: passes race
        for (auto& s : ctx->slabs) {
            if (nowait) break;
            GFB_CUDA(ctx, cudaSetDevice(s.device));
            launch_halo_signal_wait(s.stream, s.peer_flag_prev, s.peer_flag_next, s.d_flags, serial);
            GFB_CHECK(post_launch(ctx));
        }
        return GFB_OK;
    }
    fa.leave_sms = true;
    for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 0, 2, g->tloc - 1, fa));  // slices 0 and tloc-1 in one launch
    GFB_CHECK(exchange_halo_buffers(ctx, g, *uout, true, unitary));  // unitary links: two rows may travel (GFB200_HALO_SU3)
    for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 1, g->tloc - 2, 1, fa));
    return join_halo_exchange(ctx);
}

int gfb_force(gfb_mom* f, gfb_gauge* g, double beta) {
    if (!f || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, f)) return fail(g->ctx, GFB_ERR_ARG, "force and U must have the same lattice");
    GFB_CHECK(ensure_halo(g));
    FusedArgs fa;
    fa.a = -beta / 6.0;  // -(1/NC) * (beta/2)
    fa.write_z = true;
    return fused_pass(g, g->d, nullptr, nullptr, &f->d, fa);
}
int gfb_flow_force(gfb_mom* f, gfb_gauge* g) {
    if (!f || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, f)) return fail(g->ctx, GFB_ERR_ARG, "force and U must have the same lattice");
    GFB_CHECK(ensure_halo(g));
    FusedArgs fa;
    fa.a = 1.0;
    fa.write_z = true;
    return fused_pass(g, g->d, nullptr, nullptr, &f->d, fa);
}
int gfb_update_momenta(gfb_mom* p, gfb_gauge* g, double eps, double beta) {
    if (!p || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, p)) return fail(g->ctx, GFB_ERR_ARG, "P and U must have the same lattice");
    if (!std::isfinite(eps)) return fail(g->ctx, GFB_ERR_ARG, "the momentum step size must be finite");
    GFB_CHECK(ensure_halo(g));
    FusedArgs fa;
    fa.a = eps * (-beta / 6.0);
    fa.b = 1.0;
    fa.read_z = true;
    fa.write_z = true;
    return fused_pass(g, g->d, nullptr, &p->d, &p->d, fa);
}
int gfb_update_links(gfb_gauge* g, const gfb_mom* p, double eps) {
    if (!p || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (!same_shape(g, p)) return fail(ctx, GFB_ERR_ARG, "U and P must have the same lattice");
    if (!std::isfinite(eps)) return fail(ctx, GFB_ERR_ARG, "the gauge-field step size must be finite");
    const bool split = ctx->nslabs_total > 1;
    auto launch = [&](size_t i, int t0, int tc, int stride = 1) -> int {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        if (tc <= 0) return GFB_OK;
        Geom geo = geom_of(g, i);
        geo.t_stride = stride;
        launch_update_links(ctx->slabs[i].stream, geo, t0, tc, g->d[i], g->d[i], p->d[i], eps);
        return post_launch(ctx);
    };
    if (!split) {
        for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 0, g->tloc));
        g->halo_valid = false;
        return GFB_OK;
    }
    // site-local update: boundary slices first, their exchange overlaps the interior (set_wing_U!/set_halo! of the
    // reference happens inside substitute_U!, gaugefields_4D_MPILattice.jl:497-512)
    for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 0, 2, g->tloc - 1));
    GFB_CHECK(exchange_halo_buffers(ctx, g, g->d, true, true));
    for (size_t i = 0; i < ctx->slabs.size(); i++) GFB_CHECK(launch(i, 1, g->tloc - 2));
    GFB_CHECK(join_halo_exchange(ctx));
    g->halo_valid = true;
    return GFB_OK;
}
int gfb_exp_aF_U(gfb_gauge* w, double a, const gfb_mom* f, const gfb_gauge* u) {
    if (!w || !f || !u) return fail(w ? w->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = w->ctx;
    if (!same_shape(w, u) || !same_shape(w, f)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    if (a == 0.0) return fail(ctx, GFB_ERR_ARG, "the step must not be zero in exp_aF_U");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_update_links(ctx->slabs[i].stream, geom_of(w, i), 0, w->tloc, u->d[i], w->d[i], f->d[i], a);
        GFB_CHECK(post_launch(ctx));
    }
    w->unitary = u->unitary;
    w->halo_valid = false;
    return GFB_OK;
}

// S = -(2/NC) (c_plaq sum Re tr plaquettes + c_rect sum Re tr rectangles): the potential of a GaugeAction with the terms
// (c_plaq, plaquette + plaquette') and (c_rect, rectangular + rectangular') (md_potential, molecular_dynamics.jl:247-249)
static int hamiltonian_general(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, double* out) {
    if (c_rect == 0.0) return gfb_hamiltonian(g, p, 2.0 * c_plaq, out);
    double ls[2], kin = 0.0;
    GFB_CHECK(gfb_loop_sums(g, ls));
    GFB_CHECK(gfb_kinetic(p, &kin));
    *out = -(2.0 / 3.0) * (c_plaq * ls[0] + c_rect * ls[1]) + 0.5 * kin;
    return GFB_OK;
}
static int update_momenta_general(gfb_mom* p, gfb_gauge* g, double eps, double c_plaq, double c_rect) {
    FusedArgs fa;
    fa.a = eps * (-1.0 / 3.0); fa.b = 1.0; fa.read_z = true; fa.write_z = true;
    fa.c_plaq = c_plaq; fa.c_rect = c_rect;
    GFB_CHECK(ensure_halo(g));
    return fused_pass(g, g->d, nullptr, &p->d, &p->d, fa);
}
static int md_trajectory_impl(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, int steps, double tau, int integrator, int fused, double* H) {
    if (!p || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (!same_shape(g, p)) return fail(ctx, GFB_ERR_ARG, "U and P must have the same lattice");
    if (steps <= 0) return fail(ctx, GFB_ERR_ARG, "steps must be positive");
    if (!std::isfinite(tau)) return fail(ctx, GFB_ERR_ARG, "trajectory_length must be finite");
    if (tau == 0.0) return fail(ctx, GFB_ERR_ARG, "trajectory_length must not be zero");
    if (integrator != GFB_QPQ && integrator != GFB_PQP) return fail(ctx, GFB_ERR_ARG, "integrator must be QPQ or PQP");
    if (H) GFB_CHECK(hamiltonian_general(g, p, c_plaq, c_rect, &H[0]));
    const double eps = tau / steps;
    if (!fused) {
        for (int k = 0; k < steps; k++) {
            if (integrator == GFB_QPQ) {
                GFB_CHECK(gfb_update_links(g, p, eps / 2));
                GFB_CHECK(update_momenta_general(p, g, eps, c_plaq, c_rect));
                GFB_CHECK(gfb_update_links(g, p, eps / 2));
            } else {
                GFB_CHECK(update_momenta_general(p, g, eps / 2, c_plaq, c_rect));
                GFB_CHECK(gfb_update_links(g, p, eps));
                GFB_CHECK(update_momenta_general(p, g, eps / 2, c_plaq, c_rect));
            }
        }
    } else {
        // leapfrog with the kick and the following drift in one kernel (double-buffered links) and the
        // adjacent half drifts (QPQ) / half kicks (PQP) of consecutive steps merged
        gfb_gauge_ws* ws = nullptr;
        GFB_CHECK(get_ws(g, true, false, &ws));
        const double kf = -1.0 / 3.0;
        FusedArgs fa;
        fa.c_plaq = c_plaq; fa.c_rect = c_rect;
        fa.b = 1.0; fa.read_z = true; fa.write_z = true; fa.do_exp = true;
        if (integrator == GFB_QPQ) {
            GFB_CHECK(gfb_update_links(g, p, eps / 2));
            for (int k = 0; k < steps; k++) {
                GFB_CHECK(ensure_halo(g));
                fa.a = eps * kf;
                fa.c = (k == steps - 1) ? eps / 2 : eps;
                GFB_CHECK(fused_pass(g, g->d, &ws->alt, &p->d, &p->d, fa));
                std::swap(g->d, ws->alt);
                g->halo_valid = (c_rect == 0.0);  // exchanged inside the pass (the rectangle path works on wide copies instead)
            }
        } else {
            for (int k = 0; k < steps; k++) {
                GFB_CHECK(ensure_halo(g));
                fa.a = ((k == 0) ? eps / 2 : eps) * kf;
                fa.c = eps;
                GFB_CHECK(fused_pass(g, g->d, &ws->alt, &p->d, &p->d, fa));
                std::swap(g->d, ws->alt);
                g->halo_valid = (c_rect == 0.0);  // exchanged inside the pass (the rectangle path works on wide copies instead)
            }
            GFB_CHECK(update_momenta_general(p, g, eps / 2, c_plaq, c_rect));
        }
    }
    if (H) GFB_CHECK(hamiltonian_general(g, p, c_plaq, c_rect, &H[1]));
    return GFB_OK;
}

int gfb_md_trajectory(gfb_gauge* g, gfb_mom* p, double beta, int steps, double tau, int integrator, int fused, double* H) {
    return md_trajectory_impl(g, p, beta / 2.0, 0.0, steps, tau, integrator, fused, H);
}
int gfb_md_trajectory_general(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, int steps, double tau, int integrator, int fused, double* H) {
    return md_trajectory_impl(g, p, c_plaq, c_rect, steps, tau, integrator, fused, H);
}
int gfb_force_general(gfb_mom* f, gfb_gauge* g, double c_plaq, double c_rect) {
    if (!f || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, f)) return fail(g->ctx, GFB_ERR_ARG, "force and U must have the same lattice");
    GFB_CHECK(ensure_halo(g));
    FusedArgs fa;
    fa.a = -1.0 / 3.0;  // -(1/NC)
    fa.write_z = true;
    fa.c_plaq = c_plaq; fa.c_rect = c_rect;
    return fused_pass(g, g->d, nullptr, nullptr, &f->d, fa);
}
int gfb_update_momenta_general(gfb_mom* p, gfb_gauge* g, double eps, double c_plaq, double c_rect) {
    if (!p || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, p)) return fail(g->ctx, GFB_ERR_ARG, "P and U must have the same lattice");
    if (!std::isfinite(eps)) return fail(g->ctx, GFB_ERR_ARG, "the momentum step size must be finite");
    return update_momenta_general(p, g, eps, c_plaq, c_rect);
}
int gfb_hamiltonian_general(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, double* out) {
    if (!g || !p || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (!same_shape(g, p)) return fail(g->ctx, GFB_ERR_ARG, "U and p must have the same lattice");
    return hamiltonian_general(g, p, c_plaq, c_rect, out);
}

static int flow_impl(gfb_gauge* g, double eps, int nsteps, double c_plaq, double c_rect) {
    if (!g) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (nsteps <= 0) return fail(ctx, GFB_ERR_ARG, "steps must be positive");
    if (!(eps > 0.0) || !std::isfinite(eps)) return fail(ctx, GFB_ERR_ARG, "step_size must be positive");
    gfb_gauge_ws* ws = nullptr;
    GFB_CHECK(get_ws(g, true, true, &ws));
    // Luescher RK3 in 2N-storage form (identical to gradientflow.jl:192-226 up to rounding):
    //   Z0 = -eps F(U);             W1 = exp(Z0/4) U
    //   Z1 = -(8/9) eps F(W1) - (17/36) Z0;  W2 = exp(Z1) W1
    //   Z2 = -(3/4) eps F(W2) - Z1;          U' = exp(Z2) W2
    for (int k = 0; k < nsteps; k++) {
        FusedArgs fa;
        fa.write_z = true; fa.do_exp = true;
        fa.c_plaq = c_plaq; fa.c_rect = c_rect;
        GFB_CHECK(ensure_halo(g));
        fa.a = -eps; fa.b = 0.0; fa.c = 0.25; fa.read_z = false;
        GFB_CHECK(fused_pass(g, g->d, &ws->alt, nullptr, &ws->z, fa));
        std::swap(g->d, ws->alt); g->halo_valid = (c_rect == 0.0);
        GFB_CHECK(ensure_halo(g));
        fa.a = -(8.0 / 9.0) * eps; fa.b = -17.0 / 36.0; fa.c = 1.0; fa.read_z = true;
        GFB_CHECK(fused_pass(g, g->d, &ws->alt, &ws->z, &ws->z, fa));
        std::swap(g->d, ws->alt); g->halo_valid = (c_rect == 0.0);
        GFB_CHECK(ensure_halo(g));
        fa.a = -(3.0 / 4.0) * eps; fa.b = -1.0; fa.c = 1.0; fa.read_z = true;
        GFB_CHECK(fused_pass(g, g->d, &ws->alt, &ws->z, &ws->z, fa));
        std::swap(g->d, ws->alt); g->halo_valid = (c_rect == 0.0);
    }
    return GFB_OK;
}
int gfb_flow(gfb_gauge* g, double eps, int nsteps) { return flow_impl(g, eps, nsteps, 1.0, 0.0); }
// flow!(U, ::Gradientflow_general) (src/smearing/gradientflow.jl:240-316) for link values (c_plaq, c_rect) of the loop sets
// ("plaquette", "rectangular"): F = TAcoeffs(U dSdU), the same RK3 scheme; (1, 0) is the Wilson flow
int gfb_flow_general(gfb_gauge* g, double eps, int nsteps, double c_plaq, double c_rect) { return flow_impl(g, eps, nsteps, c_plaq, c_rect); }

// heatbath!(U, ::Heatbath) / overrelaxation!(U, ...) for the Wilson action (src/heatbath/heatbathmodule.jl:481-650, 1799-1860):
// for every direction and both checkerboard colours, refresh the halo, then update every link of that colour in place
static int heatbath_sweep(gfb_gauge* g, double beta, uint64_t seed, uint64_t sweep, int rng_alg, bool overrelax) {
    if (!g) return fail(nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (rng_alg != GFB_PHILOX4X32) return fail(ctx, GFB_ERR_ARG, "the B200 backend implements the Philox4x32 site RNG");
    if ((g->nx | g->ny | g->nz | g->nt) & 1) return fail(ctx, GFB_ERR_ARG, "the checkerboard update needs even lattice extents");
    if (!overrelax && !(beta > 0.0 && std::isfinite(beta))) return fail(ctx, GFB_ERR_ARG, "beta must be positive and finite");
    for (auto& s : ctx->slabs) {
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaMemsetAsync(s.d_flags + 8, 0, sizeof(int), s.stream));
    }
    for (int mu = 0; mu < 4; mu++)
        for (int colour = 0; colour < 2; colour++) {
            GFB_CHECK(ensure_halo(g));
            for (size_t i = 0; i < ctx->slabs.size(); i++) {
                Slab& s = ctx->slabs[i];
                GFB_CUDA(ctx, cudaSetDevice(s.device));
                launch_heatbath(s.stream, geom_of(g, i), g->d[i], mu, colour, beta, seed, sweep, overrelax, reinterpret_cast<int*>(s.d_flags + 8));
                GFB_CHECK(post_launch(ctx));
            }
            g->halo_valid = false;
        }
    g->unitary = 1;  // every updated link was reunitarised; the update preserves the group manifold
    int total = 0;
    for (auto& s : ctx->slabs) {
        int f = 0;
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        GFB_CUDA(ctx, cudaMemcpyAsync(&f, s.d_flags + 8, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        GFB_CUDA(ctx, cudaStreamSynchronize(s.stream));
        total += f;
    }
    if (total) return fail(ctx, GFB_ERR_NUMERIC, overrelax ? "overrelaxation normalization failed at " + std::to_string(total) + " site(s)"
                                                           : "SU(3) heatbath failed at " + std::to_string(total) + " site(s) after 100000 tries");
    return GFB_OK;
}
int gfb_heatbath(gfb_gauge* g, double beta, uint64_t seed, uint64_t sweep, int rng_alg) { return heatbath_sweep(g, beta, seed, sweep, rng_alg, false); }
int gfb_overrelaxation(gfb_gauge* g, double beta, uint64_t seed, uint64_t sweep, int rng_alg) { return heatbath_sweep(g, beta, seed, sweep, rng_alg, true); }

int gfb_stout_forward(gfb_gauge* out, gfb_gauge* in, double rho, gfb_mom* q) {
    if (!out || !in) return fail(in ? in->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = in->ctx;
    if (out == in) return fail(ctx, GFB_ERR_ARG, "stout forward needs distinct input and output configurations");
    if (!same_shape(out, in) || (q && !same_shape(in, q))) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    GFB_CHECK(ensure_halo(in));
    // Q_mu = TA(rho * V_mu U_mu^dag) = -rho * TA(U_mu V_mu^dag);  U' = exp(Q_mu) U_mu
    FusedArgs fa;
    fa.a = -rho; fa.c = 1.0; fa.do_exp = true; fa.write_z = (q != nullptr);
    GFB_CHECK(fused_pass(in, in->d, &out->d, nullptr, q ? &q->d : nullptr, fa));
    out->unitary = in->unitary;
    out->halo_valid = true;
    return GFB_OK;
}

int gfb_wilson_dSdU(gfb_gauge* d, gfb_gauge* g, double beta) {
    if (!d || !g) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = g->ctx;
    if (d == g || !same_shape(d, g)) return fail(ctx, GFB_ERR_ARG, "derivative field must be a distinct configuration of the same shape");
    GFB_CHECK(ensure_halo(g));
    bool unitary = true;
    GFB_CHECK(links_are_unitary(g, &unitary));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_staple_field(ctx->slabs[i].stream, geom_of(g, i), g->d[i], d->d[i], beta / 2.0, !unitary);
        GFB_CHECK(post_launch(ctx));
    }
    d->unitary = 0;  // a derivative field, not a configuration
    d->halo_valid = false;
    return GFB_OK;
}
int gfb_kick_from_dSdU(gfb_mom* p, gfb_gauge* u, gfb_gauge* dsdu, double factor) {
    if (!p || !u || !dsdu) return fail(u ? u->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = u->ctx;
    if (!same_shape(u, dsdu) || !same_shape(u, p)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_kick_from_dsdu(ctx->slabs[i].stream, geom_of(u, i), u->d[i], dsdu->d[i], p->d[i], factor);
        GFB_CHECK(post_launch(ctx));
    }
    return GFB_OK;
}

int gfb_stout_backward(gfb_gauge* d_in, gfb_gauge* d_out, gfb_gauge* in, double rho) {
    if (!d_in || !d_out || !in) return fail(in ? in->ctx : nullptr, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = in->ctx;
    if (d_in == d_out || d_in == in) return fail(ctx, GFB_ERR_ARG, "stout backward needs a distinct output field");
    if (!same_shape(d_in, in) || !same_shape(d_out, in)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    if (!std::isfinite(rho)) return fail(ctx, GFB_ERR_ARG, "rho must be finite");
    gfb_gauge_ws* ws = nullptr;
    GFB_CHECK(get_ws(d_in, true, false, &ws));  // Lambda = dS/dC lives in d_in's spare buffer (same layout, with halo slots)
    GFB_CHECK(ensure_halo(in));
    bool unitary = true;
    GFB_CHECK(links_are_unitary(in, &unitary));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_stout_lambda(ctx->slabs[i].stream, geom_of(in, i), in->d[i], d_out->d[i], ws->alt[i], d_in->d[i], rho, !unitary);
        GFB_CHECK(post_launch(ctx));
    }
    GFB_CHECK(exchange_halo_buffers(ctx, d_in, ws->alt, false));  // neighbours' Lambda on the slab faces
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_stout_backward(ctx->slabs[i].stream, geom_of(in, i), in->d[i], ws->alt[i], d_in->d[i], rho);
        GFB_CHECK(post_launch(ctx));
    }
    d_in->unitary = 0;
    d_in->halo_valid = false;
    return GFB_OK;
}

// ---- primitive table ------------------------------------------------------------------------------------
static FieldRef ref_of(const gfb_field* f, size_t i) {
    if (f->is_view) return FieldRef{f->parent->d[i] + (size_t)f->mu * 9 * f->nx * f->ny * f->nz, 36};
    return FieldRef{f->d[i], 9};
}
// a view whose configuration was freed (or any handle whose context is gone) cannot be used any more
static int check_field(const gfb_field* f) {
    if (!f || !f->ctx) return fail(nullptr, GFB_ERR_ARG, "null or orphaned field handle");
    if (f->is_view && !f->parent) return fail(f->ctx, GFB_ERR_ARG, "this view's gauge configuration has been freed");
    return GFB_OK;
}
static const double2* plane0(const gfb_field* f) { return ref_of(f, 0).p; }
static Geom geom_of(const gfb_field* f, size_t i) { return make_geom(f->ctx, f->nx, f->ny, f->nz, f->nt, f->ctx->slabs[i].index); }
static bool same_shape(const gfb_field* a, const gfb_field* b) { return a->ctx == b->ctx && a->nx == b->nx && a->ny == b->ny && a->nz == b->nz && a->nt == b->nt; }
static void mark_written(gfb_field* f) {
    f->halo_valid = false;
    if (f->parent) { f->parent->halo_valid = false; f->parent->unitary = -1; }
}
static bool is_zero(const Shift4& s) { return !s.v[0] && !s.v[1] && !s.v[2] && !s.v[3]; }
static int read_shift(gfb_ctx* ctx, const int* shift4, Shift4* out) {
    for (int k = 0; k < 4; k++) out->v[k] = shift4 ? shift4[k] : 0;
    if (ctx->nslabs_total > 1 && std::abs(out->v[3]) > 1)
        return fail(ctx, GFB_ERR_ARG, "t-shifts beyond the halo width 1 are not supported on a t-slab decomposition");
    return GFB_OK;
}
// t-halo of ONE field: its 9 planes of the first / last local slice go to the neighbours' halo slots (all 9 both ways)
static int field_halo(gfb_field* f, const Shift4& s) {
    gfb_ctx* ctx = f->ctx;
    const int G = ctx->nslabs_total;
    if (G == 1 || s.v[3] == 0) return GFB_OK;
    if (f->halo_valid && !f->is_view) return GFB_OK;
    const size_t v3 = (size_t)f->nx * f->ny * f->nz;
    const size_t slice = (size_t)f->slice_planes() * v3 * 2, count = 9 * v3 * 2;  // doubles
    NcclGroup grp;
    GFB_NCCL(ctx, grp.start());
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& sl = ctx->slabs[i];
        const int prev = (sl.index + G - 1) % G, next = (sl.index + 1) % G;
        double* base = reinterpret_cast<double*>(ref_of(f, i).p);
        GFB_NCCL(ctx, ncclSend(base, count, ncclDouble, prev, sl.nccl, sl.stream));
        GFB_NCCL(ctx, ncclSend(base + (size_t)(f->tloc - 1) * slice, count, ncclDouble, next, sl.nccl, sl.stream));
        GFB_NCCL(ctx, ncclRecv(base + (size_t)f->tloc * slice, count, ncclDouble, next, sl.nccl, sl.stream));
        GFB_NCCL(ctx, ncclRecv(base + (size_t)(f->tloc + 1) * slice, count, ncclDouble, prev, sl.nccl, sl.stream));
    }
    GFB_NCCL(ctx, grp.end());
    f->halo_valid = true;
    return GFB_OK;
}

int gfb_field_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_field** out) {
    if (!out) return fail(ctx, GFB_ERR_ARG, "out is null");
    *out = nullptr;
    GFB_CHECK(check_dims(ctx, nx, ny, nz, nt));
    gfb_field* f = new gfb_field();
    f->ctx = ctx; f->nx = nx; f->ny = ny; f->nz = nz; f->nt = nt;
    f->tloc = nt / ctx->nslabs_total;
    f->has_halo = ctx->nslabs_total > 1;
    f->d.assign(ctx->slabs.size(), nullptr);
    const size_t elems = (size_t)(f->tloc + (f->has_halo ? 2 : 0)) * 9 * (size_t)nx * ny * nz;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        GFB_CUDA(ctx, cudaMalloc(&f->d[i], elems * sizeof(double2)));
        GFB_CUDA(ctx, cudaMemsetAsync(f->d[i], 0, elems * sizeof(double2), ctx->slabs[i].stream));
    }
    ctx->fields.insert(f);
    *out = f;
    return GFB_OK;
}
int gfb_field_view(gfb_gauge* g, int mu, gfb_field** out) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, GFB_ERR_ARG, "null argument");
    if (mu < 0 || mu > 3) return fail(g->ctx, GFB_ERR_ARG, "mu must be in 0..3");
    gfb_field* f = new gfb_field();
    f->ctx = g->ctx; f->nx = g->nx; f->ny = g->ny; f->nz = g->nz; f->nt = g->nt; f->tloc = g->tloc;
    f->has_halo = g->has_halo; f->parent = g; f->mu = mu; f->is_view = true;
    // a view holds no pointers of its own: the fused MD / flow passes swap the configuration's buffers, so the planes are
    // resolved from the parent at every use (ref_of)
    g->ctx->fields.insert(f);
    *out = f;
    return GFB_OK;
}
static void release_field_memory(gfb_field* f) {
    if (!f->ctx || f->is_view) return;
    for (size_t i = 0; i < f->d.size(); i++) { cudaSetDevice(f->ctx->slabs[i].device); cudaFree(f->d[i]); }
    f->d.clear();
}
int gfb_field_free(gfb_field* f) {
    if (!f) return GFB_OK;
    if (f->ctx) { release_field_memory(f); f->ctx->fields.erase(f); }
    delete f;
    return GFB_OK;
}
static int field_transfer(gfb_field* f, double* host, bool to_host) {
    GFB_CHECK(check_field(f));
    gfb_ctx* ctx = f->ctx;
    const size_t v3 = (size_t)f->nx * f->ny * f->nz;
    const size_t bytes = v3 * f->tloc * 9 * sizeof(double2);
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        GFB_CHECK(ensure_staging(ctx, s, bytes));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        Geom geo = geom_of(f, i);
        double* h = host + (size_t)geo.t0 * v3 * 18;
        if (!to_host) GFB_CUDA(ctx, cudaMemcpyAsync(s.d_staging, h, bytes, cudaMemcpyHostToDevice, s.stream));
        launch_prim_host(s.stream, geo, ref_of(f, i), reinterpret_cast<double2*>(s.d_staging), to_host ? 1 : 0);
        GFB_CHECK(post_launch(ctx));
        if (to_host) GFB_CUDA(ctx, cudaMemcpyAsync(h, s.d_staging, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ctx->slabs) { GFB_CUDA(ctx, cudaSetDevice(s.device)); GFB_CUDA(ctx, cudaStreamSynchronize(s.stream)); }
    return GFB_OK;
}
int gfb_field_upload(gfb_field* f, const double* host) {
    if (!f || !host) return fail(f ? f->ctx : nullptr, GFB_ERR_ARG, "null argument");
    mark_written(f);
    return field_transfer(f, const_cast<double*>(host), false);
}
int gfb_field_download(gfb_field* f, double* host) {
    if (!f || !host) return fail(f ? f->ctx : nullptr, GFB_ERR_ARG, "null argument");
    return field_transfer(f, host, true);
}
static int field_fill(gfb_field* f, double diag) {
    GFB_CHECK(check_field(f));
    gfb_ctx* ctx = f->ctx;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_prim_fill(ctx->slabs[i].stream, geom_of(f, i), ref_of(f, i), diag);
        GFB_CHECK(post_launch(ctx));
    }
    mark_written(f);
    return GFB_OK;
}
int gfb_field_clear(gfb_field* f) { return field_fill(f, 0.0); }
int gfb_field_unit(gfb_field* f) { return field_fill(f, 1.0); }

static int axpy_impl(gfb_field* c, double2 alpha, gfb_field* a, const int* shift4, int dag, int assign) {
    GFB_CHECK(check_field(c));
    GFB_CHECK(check_field(a));
    gfb_ctx* ctx = c->ctx;
    if (!same_shape(c, a)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    Shift4 s;
    GFB_CHECK(read_shift(ctx, shift4, &s));
    if (!is_zero(s) && plane0(c) == plane0(a)) return fail(ctx, GFB_ERR_ARG, "a shifted source must not alias the destination");
    GFB_CHECK(field_halo(a, s));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_prim_axpy(ctx->slabs[i].stream, geom_of(c, i), ref_of(c, i), alpha, ref_of(a, i), s, dag ? 1 : 0, assign);
        GFB_CHECK(post_launch(ctx));
    }
    mark_written(c);
    return GFB_OK;
}
int gfb_field_copy(gfb_field* dst, gfb_field* src, const int* shift4, int dagger) { return axpy_impl(dst, make_double2(1.0, 0.0), src, shift4, dagger, 1); }
int gfb_axpy(gfb_field* c, double alpha_re, double alpha_im, gfb_field* a, int dag_a) { return axpy_impl(c, make_double2(alpha_re, alpha_im), a, nullptr, dag_a, 0); }

int gfb_mul(gfb_field* c, gfb_field* a, const int* shift_a4, int dag_a, gfb_field* b, const int* shift_b4, int dag_b, double alpha_re, double alpha_im,
            double beta_re, double beta_im) {
    GFB_CHECK(check_field(c));
    GFB_CHECK(check_field(a));
    GFB_CHECK(check_field(b));
    gfb_ctx* ctx = c->ctx;
    if (!same_shape(c, a) || !same_shape(c, b)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    Shift4 sa, sb;
    GFB_CHECK(read_shift(ctx, shift_a4, &sa));
    GFB_CHECK(read_shift(ctx, shift_b4, &sb));
    if ((plane0(c) == plane0(a) && !is_zero(sa)) || (plane0(c) == plane0(b) && !is_zero(sb)))
        return fail(ctx, GFB_ERR_ARG, "the destination of mul! must not alias a shifted operand");
    GFB_CHECK(field_halo(a, sa));
    GFB_CHECK(field_halo(b, sb));
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_prim_mul(ctx->slabs[i].stream, geom_of(c, i), ref_of(c, i), ref_of(a, i), sa, dag_a ? 1 : 0, ref_of(b, i), sb, dag_b ? 1 : 0,
                        make_double2(alpha_re, alpha_im), make_double2(beta_re, beta_im));
        GFB_CHECK(post_launch(ctx));
    }
    mark_written(c);
    return GFB_OK;
}
static int trace_impl(gfb_field* a, gfb_field* b, double* out2) {
    GFB_CHECK(check_field(a));
    if (b) GFB_CHECK(check_field(b));
    if (!out2) return fail(a->ctx, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = a->ctx;
    if (b && !same_shape(a, b)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        Slab& s = ctx->slabs[i];
        Geom geo = geom_of(a, i);
        GFB_CHECK(ensure_partial(ctx, s, (size_t)plaquette_blocks(geo) * 2 + 16));
        GFB_CUDA(ctx, cudaSetDevice(s.device));
        int nb = 0;
        launch_prim_trace(s.stream, geo, ref_of(a, i), b ? ref_of(b, i) : ref_of(a, i), b ? 1 : 0, s.d_partial, &nb);
        launch_final_reduce(s.stream, s.d_partial, nb, s.d_result);
        launch_final_reduce(s.stream, s.d_partial + nb, nb, s.d_result + 1);
        GFB_CHECK(post_launch(ctx, 3));
    }
    return gather_scalars(ctx, 2, out2);
}
int gfb_tr(gfb_field* a, double* out2) { return trace_impl(a, nullptr, out2); }
int gfb_tr2(gfb_field* a, gfb_field* b, double* out2) {
    if (!b) return fail(a ? a->ctx : nullptr, GFB_ERR_ARG, "null argument");
    return trace_impl(a, b, out2);
}
static int ta_exp_impl(gfb_field* out, gfb_field* in, int mode, double t) {
    GFB_CHECK(check_field(out));
    GFB_CHECK(check_field(in));
    gfb_ctx* ctx = out->ctx;
    if (!same_shape(out, in)) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    if (!std::isfinite(t)) return fail(ctx, GFB_ERR_ARG, "t must be finite");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_prim_ta_exp(ctx->slabs[i].stream, geom_of(out, i), ref_of(out, i), ref_of(in, i), mode, t);
        GFB_CHECK(post_launch(ctx));
    }
    mark_written(out);
    return GFB_OK;
}
int gfb_ta_project(gfb_field* q, gfb_field* m) { return ta_exp_impl(q, m, 0, 1.0); }
int gfb_exp(gfb_field* e, double t, gfb_field* q) { return ta_exp_impl(e, q, 1, t); }
static int mom_impl(gfb_field* f, gfb_mom* p, int mu, int mode, double s) {
    GFB_CHECK(check_field(f));
    if (!p) return fail(f->ctx, GFB_ERR_ARG, "null argument");
    gfb_ctx* ctx = f->ctx;
    if (mu < 0 || mu > 3) return fail(ctx, GFB_ERR_ARG, "mu must be in 0..3");
    if (f->ctx != p->ctx || f->nx != p->nx || f->ny != p->ny || f->nz != p->nz || f->nt != p->nt) return fail(ctx, GFB_ERR_ARG, "fields differ in shape");
    if (!std::isfinite(s)) return fail(ctx, GFB_ERR_ARG, "the scalar must be finite");
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        GFB_CUDA(ctx, cudaSetDevice(ctx->slabs[i].device));
        launch_prim_mom(ctx->slabs[i].stream, geom_of(f, i), ref_of(f, i), p->d[i], mu, mode, s);
        GFB_CHECK(post_launch(ctx));
    }
    if (mode == 1) mark_written(f);
    return GFB_OK;
}
int gfb_ta_coeffs_add(gfb_mom* p, int mu, double factor, gfb_field* m) { return mom_impl(m, p, mu, 0, factor); }
int gfb_exp_mom(gfb_field* e, double t, gfb_mom* p, int mu) { return mom_impl(e, p, mu, 1, t); }

}  // extern "C"
