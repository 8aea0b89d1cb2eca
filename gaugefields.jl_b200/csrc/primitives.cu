// primitives.cu -- the primitive table: element-wise operations on single 3x3 matrix fields with lazy shift / adjoint views
// (the set LatticeMatrices implements for Gaugefields_4D_MPILattice, src/4D/mpi_jacc/gaugefields_4D_MPILattice.jl:474-842).
// They keep every generic algorithm of the reference expressible on this backend; the hot path uses the fused kernels.
// All are HBM-bound streaming kernels: one thread per site, nine coalesced 128-bit accesses per operand.
#include "gfb_internal.h"
#include "stencil.cuh"

namespace gfb {

// a field inside a slab: element (t, k, s3) at base[((t*S + k) * v3) + s3]; S = 9 for its own buffer, 36 for a view of U[mu]
__device__ __forceinline__ unsigned fld_offset(const Geom& g, const FieldRef& f, const Coord& c) {
    return (unsigned)(c.t * f.slice_planes) * (unsigned)g.v3 + (unsigned)s3_of(g, c);
}
__device__ __forceinline__ Coord site_of(const Geom& g, long n) {
    Coord c;
    c.x = (int)(n % g.nx); n /= g.nx;
    c.y = (int)(n % g.ny); n /= g.ny;
    c.z = (int)(n % g.nz);
    c.t = (int)(n / g.nz);
    return c;
}
__device__ __forceinline__ Coord shifted(const Geom& g, Coord c, const Shift4& s) {
    if (s.v[0]) { c.x = (c.x + s.v[0]) % g.nx; if (c.x < 0) c.x += g.nx; }
    if (s.v[1]) { c.y = (c.y + s.v[1]) % g.ny; if (c.y < 0) c.y += g.ny; }
    if (s.v[2]) { c.z = (c.z + s.v[2]) % g.nz; if (c.z < 0) c.z += g.nz; }
    if (s.v[3]) {
        if (g.nslots == g.tloc) {  // single slab: periodic wrap of any t shift
            c.t = (c.t + s.v[3]) % g.tloc;
            if (c.t < 0) c.t += g.tloc;
        } else {  // t-slabs: |shift| <= 1 (checked on the host), halo slots hold the neighbours' faces
            if (s.v[3] > 0) c.t = (c.t == g.tloc - 1) ? g.t_up_wrap : c.t + 1;
            else c.t = (c.t == 0) ? g.t_dn_wrap : c.t - 1;
        }
    }
    return c;
}
__device__ __forceinline__ M3 fld_load(const FieldRef& f, const Geom& g, const Coord& c, bool dagger) {
    M3 m = m3_load_rw(f.p + fld_offset(g, f, c), (unsigned)g.v3);
    return dagger ? m3_dagger(m) : m;
}

__global__ void __launch_bounds__(128) k_prim_mul(Geom g, FieldRef c, FieldRef a, Shift4 sa, int da, FieldRef b, Shift4 sb, int db, double2 alpha, double2 beta) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    const M3 ma = fld_load(a, g, shifted(g, x, sa), da), mb = fld_load(b, g, shifted(g, x, sb), db);
    M3 r = mul_nn(ma, mb);
    double2* dst = c.p + fld_offset(g, c, x);
    const bool use_c = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
    for (int k = 0; k < 9; k++) {
        double2 v = cmul(alpha, r.e[k]);
        if (use_c) cmac(v, beta, dst[(size_t)k * g.v3]);
        dst[(size_t)k * g.v3] = v;
    }
}
__global__ void __launch_bounds__(256) k_prim_axpy(Geom g, FieldRef c, double2 alpha, FieldRef a, Shift4 sa, int da, int assign) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    const M3 ma = fld_load(a, g, shifted(g, x, sa), da);
    double2* dst = c.p + fld_offset(g, c, x);
#pragma unroll
    for (int k = 0; k < 9; k++) {
        double2 v = assign ? make_double2(0.0, 0.0) : dst[(size_t)k * g.v3];
        cmac(v, alpha, ma.e[k]);
        dst[(size_t)k * g.v3] = v;
    }
}
__global__ void __launch_bounds__(256) k_prim_fill(Geom g, FieldRef c, double diag) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    double2* dst = c.p + fld_offset(g, c, x);
#pragma unroll
    for (int k = 0; k < 9; k++) dst[(size_t)k * g.v3] = make_double2((k == 0 || k == 4 || k == 8) ? diag : 0.0, 0.0);
}
// partial[b] = Re, partial[gridDim + b] = Im of sum_x tr(A) (b == nullptr) or tr(A B)
__global__ void __launch_bounds__(128) k_prim_trace(Geom g, FieldRef a, FieldRef b, int two, double* __restrict__ partial) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double re = 0.0, im = 0.0;
    if (n < (long)g.v3 * g.tloc) {
        const Coord x = site_of(g, n);
        const M3 ma = fld_load(a, g, x, false);
        if (two) {
            const double2 t = tr_nn(ma, fld_load(b, g, x, false));
            re = t.x; im = t.y;
        } else {
            re = ma.e[0].x + ma.e[4].x + ma.e[8].x;
            im = ma.e[0].y + ma.e[4].y + ma.e[8].y;
        }
    }
    const double r = block_sum(re);
    const double i = block_sum(im);
    if (threadIdx.x == 0) { partial[blockIdx.x] = r; partial[gridDim.x + blockIdx.x] = i; }
}
// mode 0: Q = TA(M);  mode 1: E = exp(t * TA(M))
__global__ void __launch_bounds__(128) k_prim_ta_exp(Geom g, FieldRef out, FieldRef in, int mode, double t) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    const M3 m = fld_load(in, g, x, false);
    M3 r;
    if (mode == 0) r = ta_matrix(m);
    else {
        double c[8];
        ta_coeffs(m, c);
        r = exp_ta(c, t);
    }
    m3_store(out.p + fld_offset(g, out, x), (unsigned)g.v3, r);
}
// mode 0: P_mu += factor * coeffs(TA(M));  mode 1: E = exp(t * P_mu)
__global__ void __launch_bounds__(128) k_prim_mom(Geom g, FieldRef f, double* __restrict__ p, int mu, int mode, double s) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    double* pp = p + mom_offset(g, x, mu);
    if (mode == 0) {
        double c[8];
        ta_coeffs(fld_load(f, g, x, false), c);
#pragma unroll
        for (int k = 0; k < 8; k++) pp[(size_t)k * g.v3] = fma(s, c[k], pp[(size_t)k * g.v3]);
    } else {
        double c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = pp[(size_t)k * g.v3];
        m3_store(f.p + fld_offset(g, f, x), (unsigned)g.v3, exp_ta(c, s));
    }
}
// host layout <-> field (ComplexF64[3,3,NX,NY,NZ,T], element (i,j) at i + 3j)
__global__ void __launch_bounds__(256) k_prim_host(Geom g, FieldRef f, double2* __restrict__ staging, int to_host) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = site_of(g, n);
    double2* d = f.p + fld_offset(g, f, x);
    double2* h = staging + (size_t)n * 9;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (to_host) h[i + 3 * j] = d[(size_t)(3 * i + j) * g.v3];
            else d[(size_t)(3 * i + j) * g.v3] = h[i + 3 * j];
        }
}

static inline unsigned grid_for(const Geom& g, int bs) { return (unsigned)(((long)g.v3 * g.tloc + bs - 1) / bs); }

void launch_prim_mul(cudaStream_t st, const Geom& g, FieldRef c, FieldRef a, Shift4 sa, int da, FieldRef b, Shift4 sb, int db, double2 alpha, double2 beta) {
    k_prim_mul<<<grid_for(g, 128), 128, 0, st>>>(g, c, a, sa, da, b, sb, db, alpha, beta);
}
void launch_prim_axpy(cudaStream_t st, const Geom& g, FieldRef c, double2 alpha, FieldRef a, Shift4 sa, int da, int assign) {
    k_prim_axpy<<<grid_for(g, 256), 256, 0, st>>>(g, c, alpha, a, sa, da, assign);
}
void launch_prim_fill(cudaStream_t st, const Geom& g, FieldRef c, double diag) { k_prim_fill<<<grid_for(g, 256), 256, 0, st>>>(g, c, diag); }
void launch_prim_trace(cudaStream_t st, const Geom& g, FieldRef a, FieldRef b, int two, double* partial, int* nblocks) {
    const unsigned nb = grid_for(g, 128);
    k_prim_trace<<<nb, 128, 0, st>>>(g, a, b, two, partial);
    *nblocks = (int)nb;
}
void launch_prim_ta_exp(cudaStream_t st, const Geom& g, FieldRef out, FieldRef in, int mode, double t) {
    k_prim_ta_exp<<<grid_for(g, 128), 128, 0, st>>>(g, out, in, mode, t);
}
void launch_prim_mom(cudaStream_t st, const Geom& g, FieldRef f, double* p, int mu, int mode, double s) {
    k_prim_mom<<<grid_for(g, 128), 128, 0, st>>>(g, f, p, mu, mode, s);
}
void launch_prim_host(cudaStream_t st, const Geom& g, FieldRef f, double2* staging, int to_host) {
    k_prim_host<<<grid_for(g, 256), 256, 0, st>>>(g, f, staging, to_host);
}

}  // namespace gfb
