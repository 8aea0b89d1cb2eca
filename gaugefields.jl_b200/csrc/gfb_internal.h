// gfb_internal.h -- host-side structures of libgfb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/gfb200.h"
#include "lattice.cuh"

namespace gfb {

// One t-slab of the lattice resident on one GPU.
struct Slab {
    int device = 0;
    int index = 0;  // global slab index (== NCCL rank)
    cudaStream_t stream = nullptr;       // compute stream
    cudaStream_t comm_stream = nullptr;  // halo stream (overlapped exchange)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_tic = nullptr, ev_toc = nullptr;
    ncclComm_t nccl = nullptr;
    double* d_partial = nullptr;  // per-block partial sums
    size_t partial_cap = 0;
    double* d_result = nullptr;   // 8 result slots + gather area
    double* h_result = nullptr;   // pinned mirror
    void* d_staging = nullptr;    // AoS <-> SoA staging for upload/download
    size_t staging_cap = 0;
    // t-slab halo by peer stores: flag words the neighbours write (their pass serial) and we spin on
    //   d_flags[0] <- previous slab, d_flags[1] <- next slab, d_flags[2] = timeout marker
    unsigned* d_flags = nullptr;
    unsigned* peer_flag_prev = nullptr;  // previous slab's d_flags[1] (we are its next)
    unsigned* peer_flag_next = nullptr;  // next slab's d_flags[0] (we are its previous)
};

// link buffers of a one-process-per-GPU context come from a per-context cache: they are exported to the two ring neighbours
// (CUDA IPC) and are therefore never handed back to the driver before gfb_finalize (cudaFree of an exported allocation while
// a neighbour still maps it is undefined); a released buffer is reused by the next request of the same size.
struct PoolBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool used = false;
    double2* peer_prev = nullptr;  // the previous / next slab's buffer allocated by the same collective call
    double2* peer_next = nullptr;
};

}  // namespace gfb

struct gfb_gauge;
struct gfb_mom;
struct gfb_field;
struct gfb_ctx {
    std::vector<gfb::Slab> slabs;  // local slabs
    int nslabs_total = 1;
    bool distributed = false;  // one process per GPU
    std::string err;
    long long launches = 0;
    // halo exchange by peer stores inside the fused kernels (api.cu, fused_pass); false -> NCCL send/recv
    bool peer_ok = false;
    unsigned pass_serial = 0;
    std::vector<gfb::PoolBuf> pool;               // distributed contexts only
    std::map<std::string, void*> ipc_opened;      // IPC handle bytes -> mapping in this process
    // live handles: gfb_finalize releases their device memory and orphans them (ctx = nullptr), so that a handle freed after
    // its context (finalizers run in any order in the Julia/Python hosts) only deletes the host struct
    std::set<gfb_gauge*> gauges;
    std::set<gfb_mom*> moms;
    std::set<gfb_field*> fields;
};

struct gfb_gauge {
    gfb_ctx* ctx = nullptr;
    int nx = 0, ny = 0, nz = 0, nt = 0, tloc = 0;
    bool has_halo = false;
    bool halo_valid = false;
    // 1: every link is SU(3) to 1e-12 (two-row products allowed), 0: not (full 3x3 products), -1: unknown (checked on first use)
    int unitary = -1;
    std::vector<double2*> d;  // per local slab
    // workspaces owned by the handle: double buffer of the fused updates, flow field Z
    std::vector<double2*> alt;
    std::vector<double*> z;
    std::vector<double2*> wide;  // t-slab decompositions, general-action path: tloc + 4 slices (two halo slices either side, natural order)
    size_t elems_per_slab() const { return (size_t)(tloc + (has_halo ? 2 : 0)) * 36 * (size_t)nx * ny * nz; }
    size_t slice_elems() const { return (size_t)36 * nx * ny * nz; }
};

struct gfb_mom {
    gfb_ctx* ctx = nullptr;
    int nx = 0, ny = 0, nz = 0, nt = 0, tloc = 0;
    std::vector<double*> d;
    size_t elems_per_slab() const { return (size_t)tloc * 32 * (size_t)nx * ny * nz; }
};

// one 3x3 matrix field (primitive table): either its own buffer (9 planes per slice) or a view of U[mu] (36 planes per slice)
struct gfb_field {
    gfb_ctx* ctx = nullptr;
    int nx = 0, ny = 0, nz = 0, nt = 0, tloc = 0;
    bool has_halo = false, halo_valid = false;
    gfb_gauge* parent = nullptr;  // views: the configuration (null once it has been freed)
    bool is_view = false;
    int mu = 0;
    std::vector<double2*> d;      // own buffers only (views resolve their planes from the parent at every use)
    int slice_planes() const { return is_view ? 36 : 9; }
};

namespace gfb {

struct FieldRef {
    double2* p;
    int slice_planes;
};
struct Shift4 {
    int v[4];
};
void launch_prim_mul(cudaStream_t st, const Geom& g, FieldRef c, FieldRef a, Shift4 sa, int da, FieldRef b, Shift4 sb, int db, double2 alpha, double2 beta);
void launch_prim_axpy(cudaStream_t st, const Geom& g, FieldRef c, double2 alpha, FieldRef a, Shift4 sa, int da, int assign);
void launch_prim_fill(cudaStream_t st, const Geom& g, FieldRef c, double diag);
void launch_prim_trace(cudaStream_t st, const Geom& g, FieldRef a, FieldRef b, int two, double* partial, int* nblocks);
void launch_prim_ta_exp(cudaStream_t st, const Geom& g, FieldRef out, FieldRef in, int mode, double t);
void launch_prim_mom(cudaStream_t st, const Geom& g, FieldRef f, double* p, int mu, int mode, double s);
void launch_prim_host(cudaStream_t st, const Geom& g, FieldRef f, double2* staging, int to_host);

Geom make_geom(const gfb_ctx* ctx, int nx, int ny, int nz, int nt, int slab_global_index);

int fail(gfb_ctx* ctx, int code, const std::string& msg);

#define GFB_CUDA(ctx, call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return gfb::fail(ctx, GFB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)
#define GFB_NCCL(ctx, call)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return gfb::fail(ctx, GFB_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r_));          \
    } while (0)
#define GFB_CHECK(expr)                    \
    do {                                   \
        int s_ = (expr);                   \
        if (s_ != GFB_OK) return s_;       \
    } while (0)

// ---- kernel launch wrappers (kernels.cu); all asynchronous on the given stream ----------------
struct FusedArgs {
    double a = 0, b = 0, c = 0;  // Z' = a*TA(U V^dag) + b*Z ; Uout = exp(c*Z') U
    bool read_z = false, write_z = false, do_exp = false;
    // t-slab halo by peer stores: the neighbours' copies of the output link buffer, or null
    double2* peer_prev = nullptr;
    double2* peer_next = nullptr;
    bool full3 = false;      // links not unitary to 1e-12: full 3x3 products (k_force_fused only, as the reference's staples)
    bool leave_sms = false;
    // general-action path (general.cu): V = c_plaq V_plaq + c_rect V_rect; c_rect != 0 selects k_force_general
    double c_plaq = 1.0, c_rect = 0.0;  // NCCL-overlapped slab interior: one CTA per item instead of a persistent grid, so the send/recv kernels get SMs
};
void launch_force_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                        const FusedArgs& fa);
// persistent t-marching shared-memory variant of launch_force_fused (tmarch.cu); false = launch not covered, nothing launched
bool launch_tmarch_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                         const FusedArgs& fa);
// general-action path (general.cu).  Links are read at slices [t_begin, t_begin + t_count) of `u` (whose neighbours two slices away
// must be addressable: a single-slab field, or the wide copy of a slab) and results are written at slice t - t_shift.
void launch_force_general(cudaStream_t st, const Geom& g, int t_begin, int t_count, int t_shift, const double2* uin, double2* uout, const double* zin, double* zout,
                          const FusedArgs& fa);
// partial[0..nb) = block sums of sum_{mu<nu} Re tr P, partial[nb..2nb) = block sums of Re tr over the 12 rectangle loops
void launch_loop_sums(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* u, double* partial, int* nblocks);
// kind 0 plaquette, 1 clover, 2 rectangle field strengths; density[x + nx*(y + ny*(z + nz*(t - t_shift)))] (+)= weight * q(x)
void launch_topological_density(cudaStream_t st, const Geom& g, int t_begin, int t_count, int t_shift, const double2* u, double* density, int kind, double weight,
                                bool accumulate);
void launch_sum_plain(cudaStream_t st, const double* v, size_t n, double* partial, int* nblocks);
// one checkerboard colour of direction mu: Cabibbo-Marinari heatbath (Kennedy-Pendleton) or overrelaxation, in place (heatbath.cu)
void launch_heatbath(cudaStream_t st, const Geom& g, double2* u, int mu, int colour, double beta, unsigned long long seed, unsigned long long sweep, bool overrelax,
                     int* failures);
// peer-store halo exchange: tell both ring neighbours that pass `serial` is complete here, then wait for theirs
void launch_halo_signal_wait(cudaStream_t st, unsigned* peer_flag_prev, unsigned* peer_flag_next, unsigned* my_flags, unsigned serial);
// max over the local links of |U U^dag - 1|_max and |det U - 1| -> partial[0..nblocks) (block maxima)
void launch_unitarity_defect(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks);
void launch_final_max(cudaStream_t st, const double* partial, int n, double* out);
void launch_update_links(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* z, double c);
void launch_plaquette(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks);
void launch_sumsq(cudaStream_t st, const double* p, size_t n, double* partial, int* nblocks);
void launch_clover_energy(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks, bool full3);
void launch_polyakov(cudaStream_t st, const Geom& g, const double2* u, const double2* pin, double2* pout, double* partial, int* nblocks);  // two interleaved partial arrays
void launch_final_reduce(cudaStream_t st, const double* partial, int n, double* out);
int plaquette_blocks(const Geom& g);
int sumsq_blocks(size_t n);
void launch_links_from_host_layout(cudaStream_t st, const Geom& g, int mu, const double2* staging, double2* u);
void launch_links_to_host_layout(cudaStream_t st, const Geom& g, int mu, const double2* u, double2* staging);
// ILDG payload (big-endian [t][z][y][x][mu][row][col], 32- or 64-bit) <-> device layout for this slab's time-slices
void launch_links_from_ildg(cudaStream_t st, const Geom& g, int precision, const void* payload, double2* u);
void launch_links_to_ildg(cudaStream_t st, const Geom& g, int precision, const double2* u, void* payload);
void launch_mom_from_host_layout(cudaStream_t st, const Geom& g, int mu, const double* staging, double* p);
void launch_mom_to_host_layout(cudaStream_t st, const Geom& g, int mu, const double* p, double* staging);
void launch_set_cold(cudaStream_t st, const Geom& g, double2* u);
void launch_set_hot(cudaStream_t st, const Geom& g, double2* u, unsigned long long seed);
void launch_gaussian(cudaStream_t st, const Geom& g, double* p, unsigned long long seed, unsigned long long sweep, double sigma);
void launch_reunitarize(cudaStream_t st, const Geom& g, double2* u);
// rebuilds row 2 = conj(row0 x row1) of the links in the two halo slots (3 spatial directions in the t+1 slot, 4 in the t-1 slot)
void launch_complete_su3_rows(cudaStream_t st, const Geom& g, double2* u);
void launch_axpy(cudaStream_t st, double* y, double a, const double* x, size_t n);
void launch_staple_field(cudaStream_t st, const Geom& g, const double2* u, double2* out, double scale, bool full3);
void launch_kick_from_dsdu(cudaStream_t st, const Geom& g, const double2* u, const double2* d, double* p, double factor);
void launch_stout_lambda(cudaStream_t st, const Geom& g, const double2* u, const double2* dout, double2* lambda, double2* din, double rho, bool full3);
void launch_stout_backward(cudaStream_t st, const Geom& g, const double2* u, const double2* lambda, double2* din, double rho);

}  // namespace gfb
