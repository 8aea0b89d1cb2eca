// kernels.cu -- hand-written sm_100a kernels of the SU(3) Wilson update path.
//
// All arithmetic is fp64 complex on the FP64 pipe (3x3 complex products are not a dense
// contraction, so tensor cores do not apply); every global access is a coalesced 128-bit (links)
// or 64-bit (momenta) access to a structure-of-arrays plane.  Roofline per kernel: DESIGN.md.
#include <cstdio>

#include "gfb_internal.h"
#include "su3.cuh"
#include "stencil.cuh"

// tuning knobs of the fused force kernel (see profiles/ and DESIGN.md)
#ifndef GFB_FF_MINBLOCKS
#define GFB_FF_MINBLOCKS 4
#endif
#ifndef GFB_FF_L2PF
#define GFB_FF_L2PF 0  // >0: each block bulk-prefetches into L2 the first-touch data (t+1 links, momenta) of the block this many blocks ahead
#endif
#ifndef GFB_FF_ROWS
#define GFB_FF_ROWS 1  // x-rows (32-site groups) per block: neighbouring rows share staple operands through L1
#endif

namespace gfb {

// ------------------------------------------------------------------------------------------------
// fused staple -> TA force -> (momentum / flow-field update) -> (exp * U)
//   Z' = a * TAcoeffs(U_mu V_mu^dag) + b * Z ;  Uout_mu = exp(c * Z') * Uin_mu
// One thread per (site, mu): threadIdx.x = site within the block, threadIdx.y = mu, so a warp holds
// 32 x-consecutive sites of one direction (coalesced, no divergence in the nu loop).
// Covers md_force!/update_momenta! (molecular_dynamics.jl:251-267,539-551), the fused leapfrog
// step, the three RK3 flow stages (gradientflow.jl:192-226) and the stout forward layer
// (stout_fast.jl:250-274).
// ------------------------------------------------------------------------------------------------
template <bool READ_Z, bool WRITE_Z, bool DO_EXP, bool FULL3>
__global__ void __launch_bounds__(128 * GFB_FF_ROWS, (GFB_FF_MINBLOCKS / GFB_FF_ROWS) > 0 ? (GFB_FF_MINBLOCKS / GFB_FF_ROWS) : 1)
k_force_fused(Geom g, int t_begin, int t_count, const double2* __restrict__ uin, double2* __restrict__ uout, const double* __restrict__ zin,
              double* __restrict__ zout, double a, double b, double c, double2* __restrict__ peer_prev, double2* __restrict__ peer_next) {
    const int mu = threadIdx.y;
    const long n = ((long)blockIdx.x * GFB_FF_ROWS + threadIdx.z) * blockDim.x + threadIdx.x;
#if GFB_FF_L2PF > 0
    if (threadIdx.x == 0 && threadIdx.z == 0) {
        // DRAM -> L2 ahead of the sweep: the first touch of a time-slice is the "+t" neighbour access, and of the momenta the
        // kick itself.  One bulk prefetch per 512-byte plane row, issued by one lane per direction.
        const long np = ((long)blockIdx.x + GFB_FF_L2PF) * GFB_FF_ROWS * 32;
        if (np + 32 * GFB_FF_ROWS <= (long)g.v3 * t_count) {
            const Coord xp = decode_site(g, np, t_begin, t_count);
            const Coord xt = step(g, xp, 3, +1);
            const unsigned sb = (unsigned)g.v3 * 16u;
            if (mu < 3) {
                const char* b = reinterpret_cast<const char*>(uin + link_offset(g, xt, mu));
#pragma unroll
                for (int k = 0; k < 9; k++)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b + (size_t)k * sb), "r"(512 * GFB_FF_ROWS) : "memory");
            }
            if (READ_Z) {
                const char* b = reinterpret_cast<const char*>(zin + mom_offset(g, xp, mu));
#pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b + (size_t)k * (sb / 2)), "r"(256 * GFB_FF_ROWS) : "memory");
            }
        }
    }
#endif
    if (n >= (long)g.v3 * t_count) return;
    const Coord x = decode_site(g, n, t_begin, t_count);
    M3 s = staple_sum<FULL3>(uin, g, x, mu);
    const M3 umu = load_link(uin, g, x, mu);
    double z[8];
    ta_coeffs_nd(umu, s, z);
    const unsigned zo = mom_offset(g, x, mu);
    const unsigned zsb = (unsigned)g.v3 * 8u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double v = a * z[k];
        if (READ_Z) v = fma(b, __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(zin + zo) + (size_t)k * zsb)), v);
        z[k] = v;
        if (WRITE_Z) *reinterpret_cast<double*>(reinterpret_cast<char*>(zout + zo) + (size_t)k * zsb) = v;
    }
    if (DO_EXP) {
        const M3 r = FULL3 ? mul_nn(exp_ta(z, c), umu) : exp_ta_times_su3(z, c, umu);
        store_link(uout, g, x, mu, r);
        // t-slab halo by peer stores (same scheme as tmarch.cu): our boundary slices go straight into the neighbours' halo slots
        const unsigned s3 = (unsigned)s3_of(g, x);
        if (peer_prev != nullptr && x.t == 0 && mu < 3) m3_store(peer_prev + ((unsigned)(g.tloc * 36 + mu * 9) * (unsigned)g.v3 + s3), (unsigned)g.v3, r);
        if (peer_next != nullptr && x.t == g.tloc - 1) m3_store(peer_next + ((unsigned)((g.tloc + 1) * 36 + mu * 9) * (unsigned)g.v3 + s3), (unsigned)g.v3, r);
    }
}

void launch_force_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                        const FusedArgs& fa) {
    const int SITES = 32 * GFB_FF_ROWS;
    dim3 block(32, 4, GFB_FF_ROWS);
    long nsites = (long)g.v3 * t_count;
    if (nsites <= 0) return;
    if (!fa.full3 && launch_tmarch_fused(st, g, t_begin, t_count, uin, uout, zin, zout, fa)) return;
    dim3 grid((unsigned)((nsites + SITES - 1) / SITES));
    double2* const pp = fa.do_exp ? fa.peer_prev : nullptr;
    double2* const pn = fa.do_exp ? fa.peer_next : nullptr;
#define GFB_LAUNCH_FF(R, W, E)                                                                                                       \
    do {                                                                                                                             \
        if (fa.full3) k_force_fused<R, W, E, true><<<grid, block, 0, st>>>(g, t_begin, t_count, uin, uout, zin, zout, fa.a, fa.b, fa.c, pp, pn);  \
        else k_force_fused<R, W, E, false><<<grid, block, 0, st>>>(g, t_begin, t_count, uin, uout, zin, zout, fa.a, fa.b, fa.c, pp, pn);          \
    } while (0)
    if (fa.read_z) {
        if (fa.do_exp) GFB_LAUNCH_FF(true, true, true);
        else GFB_LAUNCH_FF(true, true, false);
    } else {
        if (fa.do_exp) {
            if (fa.write_z) GFB_LAUNCH_FF(false, true, true);
            else GFB_LAUNCH_FF(false, false, true);
        } else GFB_LAUNCH_FF(false, true, false);
    }
#undef GFB_LAUNCH_FF
}

// ------------------------------------------------------------------------------------------------
// link update  Uout_mu = exp(c * Z_mu) * Uin_mu  (update_gaugefields!, molecular_dynamics.jl:513-531;
// exp_aF_U!, AbstractGaugefields.jl:2810-2841).  Site-local, so uout may alias uin.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_update_links(Geom g, int t_begin, int t_count, const double2* uin, double2* uout, const double* __restrict__ z, double c) {
    const int mu = blockIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * t_count) return;
    Coord x;
    x.t = t_begin + (int)(n / g.v3) * g.t_stride;
    int s3 = (int)(n % g.v3);
    const size_t zo = (size_t)(x.t * 32 + mu * 8) * g.v3 + s3;
    double zz[8];
#pragma unroll
    for (int k = 0; k < 8; k++) zz[k] = __ldg(z + zo + (size_t)k * g.v3);
    const size_t uo = (size_t)(x.t * 36 + mu * 9) * g.v3 + s3;
    M3 u;
#pragma unroll
    for (int k = 0; k < 9; k++) u.e[k] = uin[uo + (size_t)k * g.v3];
    M3 e = exp_ta(zz, c);
    M3 r = mul_nn(e, u);
#pragma unroll
    for (int k = 0; k < 9; k++) uout[uo + (size_t)k * g.v3] = r.e[k];
}
void launch_update_links(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* z, double c) {
    long n = (long)g.v3 * t_count;
    if (n <= 0) return;
    dim3 grid((unsigned)((n + 255) / 256), 4);
    k_update_links<<<grid, 256, 0, st>>>(g, t_begin, t_count, uin, uout, z, c);
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
// sum_{mu<nu} Re tr P_munu(x), one thread per site (calculate_Plaquette, AbstractGaugefields.jl:2684-2699)
__global__ void __launch_bounds__(128) k_plaquette(Geom g, const double2* __restrict__ u, double* __restrict__ partial) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (n < (long)g.v3 * g.tloc) {
        const Coord x = decode_site(g, n, 0, g.tloc);
#pragma unroll 1
        for (int mu = 0; mu < 3; mu++) {
            const Coord xm = step(g, x, mu, +1);
            const M3 umu = load_link(u, g, x, mu);
#pragma unroll 1
            for (int nu = mu + 1; nu < 4; nu++) {
                const Coord xn = step(g, x, nu, +1);
                M3 b = load_link(u, g, xm, nu);
                M3 ab = mul_nn(umu, b);  // U_mu(x) U_nu(x+mu)
                b = load_link(u, g, x, nu);
                M3 c = load_link(u, g, xn, mu);
                M3 cd = mul_nn(b, c);  // U_nu(x) U_mu(x+nu)
                acc += retr_nd(ab, cd);
            }
        }
    }
    double r = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
int plaquette_blocks(const Geom& g) { return (int)(((long)g.v3 * g.tloc + 127) / 128); }
void launch_plaquette(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks) {
    int nb = plaquette_blocks(g);
    k_plaquette<<<nb, 128, 0, st>>>(g, u, partial);
    *nblocks = nb;
}

// sum of squares of a flat fp64 array (p*p, TA_Gaugefields.jl:127-137)
__global__ void __launch_bounds__(256) k_sumsq(const double* __restrict__ p, size_t n, double* __restrict__ partial) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < n + 1; i += stride) {
        if (i + 1 < n) {
            double2 v = __ldg(reinterpret_cast<const double2*>(p + i));
            acc = fma(v.x, v.x, acc);
            acc = fma(v.y, v.y, acc);
        } else if (i < n) {
            double v = p[i];
            acc = fma(v, v, acc);
        }
    }
    double r = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
int sumsq_blocks(size_t n) {
    size_t nb = (n / 2 + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    return (int)nb;
}
void launch_sumsq(cudaStream_t st, const double* p, size_t n, double* partial, int* nblocks) {
    int nb = sumsq_blocks(n);
    k_sumsq<<<nb, 256, 0, st>>>(p, n, partial);
    *nblocks = nb;
}

// fixed-order final sum of the per-block partials (one block): deterministic for a given launch shape
__global__ void __launch_bounds__(1024) k_final_reduce(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    double r = block_sum(acc);
    if (threadIdx.x == 0) *out = r;
}
void launch_final_reduce(cudaStream_t st, const double* partial, int n, double* out) { k_final_reduce<<<1, 1024, 0, st>>>(partial, n, out); }

// clover energy density numerator: sum_x sum_{mu<nu} sum_a c_a(G_munu)^2 / 2 with G = TA(4 leaves)
// (samples/measurements/energydensity.jl:4-78; -tr(G G)/2 summed over mu != nu equals sum_{mu<nu} sum_a c_a^2/2 * 2 / 2)
__global__ void __launch_bounds__(128) k_clover_energy(Geom g, const double2* __restrict__ u, double* __restrict__ partial) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (n < (long)g.v3 * g.tloc) {
        const Coord x = decode_site(g, n, 0, g.tloc);
#pragma unroll 1
        for (int mu = 0; mu < 3; mu++) {
#pragma unroll 1
            for (int nu = mu + 1; nu < 4; nu++) {
                const Coord xpm = step(g, x, mu, +1), xpn = step(g, x, nu, +1);
                const Coord xmm = step(g, x, mu, -1), xmn = step(g, x, nu, -1);
                M3 w;
                {  // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
                    M3 a = load_link(u, g, x, mu), b = load_link(u, g, xpm, nu);
                    M3 ab = mul_nn(a, b);
                    a = load_link(u, g, x, nu); b = load_link(u, g, xpn, mu);
                    M3 cd = mul_nn(a, b);
                    w = mul_nd(ab, cd);
                }
                {  // leaf 2: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu)
                    const Coord xmmpn = step(g, xmm, nu, +1);
                    M3 a = load_link(u, g, x, nu), b = load_link(u, g, xmmpn, mu);
                    M3 ab = mul_nd(a, b);
                    a = load_link(u, g, xmm, nu); b = load_link(u, g, xmm, mu);
                    M3 cd = mul_dn(a, b);
                    mac_nn(w, ab, cd);
                }
                {  // leaf 3: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) U_mu(x)^dag
                    const Coord xmnpm = step(g, xmn, mu, +1);
                    M3 a = load_link(u, g, xmn, nu), b = load_link(u, g, xmn, mu);
                    M3 ab = mul_dn(a, b);
                    a = load_link(u, g, xmnpm, nu); b = load_link(u, g, x, mu);
                    M3 cd = mul_nd(a, b);
                    mac_nn(w, ab, cd);
                }
                {  // leaf 4: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
                    const Coord xmmmn = step(g, xmm, nu, -1);
                    M3 a = load_link(u, g, xmmmn, nu), b = load_link(u, g, xmm, mu);
                    M3 ba = mul_nn(a, b);  // U_nu(x-mu-nu) U_mu(x-mu); leaf starts with its dagger
                    a = load_link(u, g, xmmmn, mu); b = load_link(u, g, xmn, nu);
                    M3 cd = mul_nn(a, b);
                    mac_dn(w, ba, cd);
                }
                double c[8];
                ta_coeffs(w, c);
#pragma unroll
                for (int k = 0; k < 8; k++) acc = fma(0.5 * c[k], c[k], acc);
            }
        }
    }
    double r = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
// The same sum for links that are unitary to 1e-12 (the configuration's flag, api.cu links_are_unitary): every leaf is a product of
// four SU(3) links, so two rows are carried through the three products (3 x 72 FMA) and the third row is rebuilt when the leaf is
// added (36): 252 instead of 324 FP64 instructions per leaf, and the leading link of each leaf is read as two rows.
#ifndef GFB_CLOVER_MINBLOCKS
#define GFB_CLOVER_MINBLOCKS 4  // measured at 32^4 (E(t) call incl. reduction + sync): 2 blocks 0.722 ms, 3 blocks 0.763, 4 blocks 0.702; full 3x3 kernel 0.813
#endif
__global__ void __launch_bounds__(128, GFB_CLOVER_MINBLOCKS) k_clover_energy_su3(Geom g, const double2* __restrict__ u, double* __restrict__ partial) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    const unsigned pl = (unsigned)g.v3;
    if (n < (long)g.v3 * g.tloc) {
        const Coord x = decode_site(g, n, 0, g.tloc);
#pragma unroll 1
        for (int mu = 0; mu < 3; mu++) {
            const Coord xpm = step(g, x, mu, +1), xmm = step(g, x, mu, -1);
#pragma unroll 1
            for (int nu = mu + 1; nu < 4; nu++) {
                const Coord xpn = step(g, x, nu, +1), xmn = step(g, x, nu, -1);
                M3 w = m3_zero();
                {  // leaf 1: U_mu(x) U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag
                    R2 r = r2_load_rows01(u + link_offset(g, x, mu), pl);
                    r = r2_mul_nn(r, load_link(u, g, xpm, nu));
                    r = r2_mul_nd(r, load_link(u, g, xpn, mu));
                    r = r2_mul_nd(r, load_link(u, g, x, nu));
                    acc_su3(w, r);
                }
                {  // leaf 2: U_nu(x) U_mu(x-mu+nu)^dag U_nu(x-mu)^dag U_mu(x-mu)
                    R2 r = r2_load_rows01(u + link_offset(g, x, nu), pl);
                    r = r2_mul_nd(r, load_link(u, g, step(g, xmm, nu, +1), mu));
                    r = r2_mul_nd(r, load_link(u, g, xmm, nu));
                    r = r2_mul_nn(r, load_link(u, g, xmm, mu));
                    acc_su3(w, r);
                }
                {  // leaf 3: U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) U_mu(x)^dag
                    R2 r = r2_load_dag_rows01(u + link_offset(g, xmn, nu), pl);
                    r = r2_mul_nn(r, load_link(u, g, xmn, mu));
                    r = r2_mul_nn(r, load_link(u, g, step(g, xmn, mu, +1), nu));
                    r = r2_mul_nd(r, load_link(u, g, x, mu));
                    acc_su3(w, r);
                }
                {  // leaf 4: U_mu(x-mu)^dag U_nu(x-mu-nu)^dag U_mu(x-mu-nu) U_nu(x-nu)
                    const Coord xmmmn = step(g, xmm, nu, -1);
                    R2 r = r2_load_dag_rows01(u + link_offset(g, xmm, mu), pl);
                    r = r2_mul_nd(r, load_link(u, g, xmmmn, nu));
                    r = r2_mul_nn(r, load_link(u, g, xmmmn, mu));
                    r = r2_mul_nn(r, load_link(u, g, xmn, nu));
                    acc_su3(w, r);
                }
                double c[8];
                ta_coeffs(w, c);
#pragma unroll
                for (int k = 0; k < 8; k++) acc = fma(0.5 * c[k], c[k], acc);
            }
        }
    }
    double r = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
void launch_clover_energy(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks, bool full3) {
    int nb = plaquette_blocks(g);
    if (full3) k_clover_energy<<<nb, 128, 0, st>>>(g, u, partial);
    else k_clover_energy_su3<<<nb, 128, 0, st>>>(g, u, partial);
    *nblocks = nb;
}

// Polyakov loop, single slab: one thread per spatial site multiplies U_t along t
// (calculate_Polyakov_loop, AbstractGaugefields.jl:2929-2956).  partial[b] = Re, partial[gridDim+b] = Im
// pin  != null: continue the product of the previous slabs (9 planes of v3 elements);  pout != null: hand the running product
// to the next slab instead of taking the trace (t-slabs multiply in slab order: the loop winds once around the global t extent)
__global__ void __launch_bounds__(128) k_polyakov(Geom g, const double2* __restrict__ u, const double2* __restrict__ pin, double2* __restrict__ pout,
                                                  double* __restrict__ partial) {
    const int s3 = blockIdx.x * blockDim.x + threadIdx.x;
    double re = 0.0, im = 0.0;
    if (s3 < g.v3) {
        M3 p = pin ? m3_load(pin + s3, (unsigned)g.v3) : m3_load(u + (size_t)(0 * 36 + 27) * g.v3 + s3, (unsigned)g.v3);
        for (int t = pin ? 0 : 1; t < g.tloc; t++) {
            M3 q = m3_load(u + (size_t)(t * 36 + 27) * g.v3 + s3, (unsigned)g.v3);
            p = mul_nn(p, q);
        }
        if (pout) m3_store(pout + s3, (unsigned)g.v3, p);
        else {
            re = p.e[0].x + p.e[4].x + p.e[8].x;
            im = p.e[0].y + p.e[4].y + p.e[8].y;
        }
    }
    double r = block_sum(re);
    double i = block_sum(im);
    if (threadIdx.x == 0) { partial[blockIdx.x] = r; partial[gridDim.x + blockIdx.x] = i; }
}
void launch_polyakov(cudaStream_t st, const Geom& g, const double2* u, const double2* pin, double2* pout, double* partial, int* nblocks) {
    int nb = (g.v3 + 127) / 128;
    k_polyakov<<<nb, 128, 0, st>>>(g, u, pin, pout, partial);
    *nblocks = nb;
}

// halo slots after a two-row exchange of unitary links: row 2 = conj(row0 x row1) (the identity reunitarize() ends with)
__global__ void __launch_bounds__(256) k_complete_su3_rows(Geom g, double2* __restrict__ u) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_dir = g.v3;
    if (n >= 7 * per_dir) return;
    const int d = (int)(n / per_dir);  // 0..2: t+1 slot, directions 0..2;  3..6: t-1 slot, directions 0..3
    const int s3 = (int)(n - (long)d * per_dir);
    const int slot = d < 3 ? g.tloc : g.tloc + 1;
    const int mu = d < 3 ? d : d - 3;
    double2* p = u + (size_t)(slot * 36 + mu * 9) * g.v3 + s3;
    double2 r[6];
#pragma unroll
    for (int k = 0; k < 6; k++) r[k] = p[(size_t)k * g.v3];
    double2 t;
    t = cmul(r[1], r[5]); { double2 v = cmul(r[2], r[4]); t.x -= v.x; t.y -= v.y; } p[(size_t)6 * g.v3] = make_double2(t.x, -t.y);
    t = cmul(r[2], r[3]); { double2 v = cmul(r[0], r[5]); t.x -= v.x; t.y -= v.y; } p[(size_t)7 * g.v3] = make_double2(t.x, -t.y);
    t = cmul(r[0], r[4]); { double2 v = cmul(r[1], r[3]); t.x -= v.x; t.y -= v.y; } p[(size_t)8 * g.v3] = make_double2(t.x, -t.y);
}
void launch_complete_su3_rows(cudaStream_t st, const Geom& g, double2* u) {
    if (g.nslots <= g.tloc) return;
    const long n = 7L * g.v3;
    k_complete_su3_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, u);
}

// ------------------------------------------------------------------------------------------------
// t-slab halo exchange by peer stores: after a fused pass every rank tells its two ring neighbours "pass `serial` is complete
// here" (its boundary slices already sit in their halo slots: the pass stored them there) and waits for theirs.  One thread;
// runs in stream order behind the pass, so the kernel boundary has made the pass's peer stores visible system-wide before the
// flag is written.  The spin is bounded (about 20 s): a dead neighbour must not hang the GPU; d_flags[2] records the timeout
// and the host turns it into an error at the next synchronising call.
// ------------------------------------------------------------------------------------------------
__global__ void k_halo_signal_wait(unsigned* peer_flag_prev, unsigned* peer_flag_next, unsigned* my_flags, unsigned serial) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag_prev), "r"(serial) : "memory");
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag_next), "r"(serial) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned a, b;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(a) : "l"(my_flags) : "memory");
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(b) : "l"(my_flags + 1) : "memory");
        if ((int)(a - serial) >= 0 && (int)(b - serial) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) { my_flags[2] = serial; break; }
        __nanosleep(200);
    }
}
void launch_halo_signal_wait(cudaStream_t st, unsigned* peer_flag_prev, unsigned* peer_flag_next, unsigned* my_flags, unsigned serial) {
    k_halo_signal_wait<<<1, 1, 0, st>>>(peer_flag_prev, peer_flag_next, my_flags, serial);
}

// how far the links are from SU(3): max over the 9 entries of |U U^dag - 1| and |det U - 1| (two-row products need 1e-12)
__global__ void __launch_bounds__(128) k_unitarity_defect(Geom g, const double2* __restrict__ u, double* __restrict__ partial) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double worst = 0.0;
    if (n < (long)g.v3 * g.tloc) {
        const Coord x = decode_site(g, n, 0, g.tloc);
#pragma unroll 1
        for (int mu = 0; mu < 4; mu++) {
            const M3 a = load_link(u, g, x, mu);
            M3 w = mul_nd(a, a);
            w.e[0].x -= 1.0; w.e[4].x -= 1.0; w.e[8].x -= 1.0;
#pragma unroll
            for (int k = 0; k < 9; k++) worst = fmax(worst, fmax(fabs(w.e[k].x), fabs(w.e[k].y)));
            // det = row0 . (row1 x row2)
            double2 c0 = cmul(a.e[4], a.e[8]); { const double2 v = cmul(a.e[5], a.e[7]); c0.x -= v.x; c0.y -= v.y; }
            double2 c1 = cmul(a.e[5], a.e[6]); { const double2 v = cmul(a.e[3], a.e[8]); c1.x -= v.x; c1.y -= v.y; }
            double2 c2 = cmul(a.e[3], a.e[7]); { const double2 v = cmul(a.e[4], a.e[6]); c2.x -= v.x; c2.y -= v.y; }
            double2 det = cmul(a.e[0], c0);
            cmac(det, a.e[1], c1);
            cmac(det, a.e[2], c2);
            worst = fmax(worst, fmax(fabs(det.x - 1.0), fabs(det.y)));
            if (!(worst == worst)) worst = 1e300;  // NaN links are not unitary
        }
    }
    __shared__ double wp[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_down_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = worst;
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = fmax(fmax(wp[0], wp[1]), fmax(wp[2], wp[3]));
}
void launch_unitarity_defect(cudaStream_t st, const Geom& g, const double2* u, double* partial, int* nblocks) {
    const int nb = plaquette_blocks(g);
    k_unitarity_defect<<<nb, 128, 0, st>>>(g, u, partial);
    *nblocks = nb;
}
__global__ void __launch_bounds__(1024) k_final_max(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, partial[i]);
    __shared__ double wp[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wp[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int i = 0; i < 32; i++) r = fmax(r, wp[i]);
        *out = r;
    }
}
void launch_final_max(cudaStream_t st, const double* partial, int n, double* out) { k_final_max<<<1, 1024, 0, st>>>(partial, n, out); }

// ------------------------------------------------------------------------------------------------
// host layout <-> device layout.  staging holds this slab's chunk of the reference's gathered array
// ComplexF64[3,3,NX,NY,NZ,T] (element (i,j) of a site at i + 3j; src/API.jl:516-529)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_links_from_host(Geom g, int mu, const double2* __restrict__ staging, double2* __restrict__ u) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    const double2* src = staging + (size_t)n * 9;
    double2* dst = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) dst[(size_t)(3 * i + j) * g.v3] = src[i + 3 * j];
}
__global__ void __launch_bounds__(256) k_links_to_host(Geom g, int mu, const double2* __restrict__ u, double2* __restrict__ staging) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    double2* dst = staging + (size_t)n * 9;
    const double2* src = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) dst[i + 3 * j] = src[(size_t)(3 * i + j) * g.v3];
}
// ILDG binary payload (big-endian, [t][z][y][x][mu][row][col] complex; src/output/ildg_format.jl:697-746 writes it site by
// site on the host) <-> device layout: byte swap, precision conversion and the AoS -> SoA transpose in one pass over the slab,
// so a production configuration goes file -> pinned buffer -> GPU without a host-side reshuffle.  One thread per local site.
__device__ __forceinline__ unsigned long long bswap64(unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
template <typename F>
__global__ void __launch_bounds__(256) k_links_from_ildg(Geom g, const void* __restrict__ payload, double2* __restrict__ u) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
#pragma unroll 1
    for (int mu = 0; mu < 4; mu++) {
        double2* dst = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const size_t e = ((size_t)n * 4 + mu) * 9 + k;  // complex element index in the payload
            double re, im;
            if (sizeof(F) == 8) {
                const unsigned long long* src = reinterpret_cast<const unsigned long long*>(payload) + 2 * e;
                re = __longlong_as_double((long long)bswap64(src[0]));
                im = __longlong_as_double((long long)bswap64(src[1]));
            } else {
                const unsigned* src = reinterpret_cast<const unsigned*>(payload) + 2 * e;
                re = (double)__uint_as_float(__byte_perm(src[0], 0, 0x0123));
                im = (double)__uint_as_float(__byte_perm(src[1], 0, 0x0123));
            }
            dst[(size_t)k * g.v3] = make_double2(re, im);
        }
    }
}
template <typename F>
__global__ void __launch_bounds__(256) k_links_to_ildg(Geom g, const double2* __restrict__ u, void* __restrict__ payload) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
#pragma unroll 1
    for (int mu = 0; mu < 4; mu++) {
        const double2* src = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const size_t e = ((size_t)n * 4 + mu) * 9 + k;
            const double2 v = src[(size_t)k * g.v3];
            if (sizeof(F) == 8) {
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(payload) + 2 * e;
                dst[0] = bswap64((unsigned long long)__double_as_longlong(v.x));
                dst[1] = bswap64((unsigned long long)__double_as_longlong(v.y));
            } else {
                unsigned* dst = reinterpret_cast<unsigned*>(payload) + 2 * e;
                dst[0] = __byte_perm(__float_as_uint((float)v.x), 0, 0x0123);
                dst[1] = __byte_perm(__float_as_uint((float)v.y), 0, 0x0123);
            }
        }
    }
}
void launch_links_from_ildg(cudaStream_t st, const Geom& g, int precision, const void* payload, double2* u) {
    const long n = (long)g.v3 * g.tloc;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (precision == 64) k_links_from_ildg<double><<<nb, 256, 0, st>>>(g, payload, u);
    else k_links_from_ildg<float><<<nb, 256, 0, st>>>(g, payload, u);
}
void launch_links_to_ildg(cudaStream_t st, const Geom& g, int precision, const double2* u, void* payload) {
    const long n = (long)g.v3 * g.tloc;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (precision == 64) k_links_to_ildg<double><<<nb, 256, 0, st>>>(g, u, payload);
    else k_links_to_ildg<float><<<nb, 256, 0, st>>>(g, u, payload);
}

__global__ void __launch_bounds__(256) k_mom_from_host(Geom g, int mu, const double* __restrict__ staging, double* __restrict__ p) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    const double2* src = reinterpret_cast<const double2*>(staging + (size_t)n * 8);
    double* dst = p + (size_t)(t * 32 + mu * 8) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double2 v = src[k];
        dst[(size_t)(2 * k) * g.v3] = v.x;
        dst[(size_t)(2 * k + 1) * g.v3] = v.y;
    }
}
__global__ void __launch_bounds__(256) k_mom_to_host(Geom g, int mu, const double* __restrict__ p, double* __restrict__ staging) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    double2* dst = reinterpret_cast<double2*>(staging + (size_t)n * 8);
    const double* src = p + (size_t)(t * 32 + mu * 8) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = make_double2(src[(size_t)(2 * k) * g.v3], src[(size_t)(2 * k + 1) * g.v3]);
}
static inline unsigned site_grid(const Geom& g, int bs) { return (unsigned)(((long)g.v3 * g.tloc + bs - 1) / bs); }
void launch_links_from_host_layout(cudaStream_t st, const Geom& g, int mu, const double2* staging, double2* u) {
    k_links_from_host<<<site_grid(g, 256), 256, 0, st>>>(g, mu, staging, u);
}
void launch_links_to_host_layout(cudaStream_t st, const Geom& g, int mu, const double2* u, double2* staging) {
    k_links_to_host<<<site_grid(g, 256), 256, 0, st>>>(g, mu, u, staging);
}
void launch_mom_from_host_layout(cudaStream_t st, const Geom& g, int mu, const double* staging, double* p) {
    k_mom_from_host<<<site_grid(g, 256), 256, 0, st>>>(g, mu, staging, p);
}
void launch_mom_to_host_layout(cudaStream_t st, const Geom& g, int mu, const double* p, double* staging) {
    k_mom_to_host<<<site_grid(g, 256), 256, 0, st>>>(g, mu, p, staging);
}

// ------------------------------------------------------------------------------------------------
// initial fields, RNG
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_set_cold(Geom g, double2* __restrict__ u) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int mu = blockIdx.y;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    double2* dst = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 9; k++) dst[(size_t)k * g.v3] = make_double2((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0, 0.0);
}
void launch_set_cold(cudaStream_t st, const Geom& g, double2* u) { k_set_cold<<<dim3(site_grid(g, 256), 4), 256, 0, st>>>(g, u); }

// stream key = first two words of Philox((seed, sweep), (tag, direction)); DESIGN.md "Random streams"
static void host_stream_key(unsigned long long seed, unsigned long long sweep, unsigned direction, unsigned tag, unsigned* k0, unsigned* k1) {
    unsigned o[4];
    philox4x32_10((unsigned)seed, (unsigned)(seed >> 32), (unsigned)sweep, (unsigned)(sweep >> 32), tag, direction, o);
    *k0 = o[0];
    *k1 = o[1];
}
struct Keys4 {
    unsigned k0[4], k1[4];
};

// hot start: 9 complex entries uniform in (-1/2,1/2)^2, filled column by column, then reunitarised
// (randomGaugefields: gaugefields_4D_nowing.jl:240-279; keyed per global site as gaugefields_4D_MPILattice.jl:430-472)
__global__ void __launch_bounds__(128) k_set_hot(Geom g, double2* __restrict__ u, Keys4 keys) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int mu = blockIdx.y;
    Coord x;
    x.t = (int)(n / g.v3);
    int s3 = (int)(n - (long)x.t * g.v3);
    const unsigned long long gs = (unsigned long long)s3 + (unsigned long long)g.v3 * (unsigned long long)(g.t0 + x.t);
    M3 m;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        double u0, u1;
        site_uniform_pair(keys.k0[mu], keys.k1[mu], gs, (unsigned)k, u0, u1);
        m.e[3 * (k % 3) + (k / 3)] = make_double2(u0 - 0.5, u1 - 0.5);
    }
    m = reunitarize(m);
    double2* dst = u + (size_t)(x.t * 36 + mu * 9) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 9; k++) dst[(size_t)k * g.v3] = m.e[k];
}
void launch_set_hot(cudaStream_t st, const Geom& g, double2* u, unsigned long long seed) {
    Keys4 keys;
    for (int mu = 0; mu < 4; mu++) host_stream_key(seed, 0ull, (unsigned)(mu + 1), 0x00484f54u, &keys.k0[mu], &keys.k1[mu]);
    k_set_hot<<<dim3(site_grid(g, 128), 4), 128, 0, st>>>(g, u, keys);
}

// Gaussian momenta: 4 Box-Muller (value, spare) pairs per (site, direction)
// (gauss_distribution!, TA_gaugefields_4D_MPILattice.jl:157-193; stream structure
//  test/MPIJACCtest/random_fields_site_rng.jl:22-42)
__global__ void __launch_bounds__(256) k_gaussian(Geom g, double* __restrict__ p, Keys4 keys, double sigma) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int mu = blockIdx.y;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    const unsigned long long gs = (unsigned long long)s3 + (unsigned long long)g.v3 * (unsigned long long)(g.t0 + t);
    double* dst = p + (size_t)(t * 32 + mu * 8) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double u0, u1;
        site_uniform_pair(keys.k0[mu], keys.k1[mu], gs, (unsigned)k, u0, u1);
        double r = sigma * sqrt(-2.0 * log(1.0 - u0));
        double sn, cs;
        sincos(6.283185307179586476925286766559 * u1, &sn, &cs);
        dst[(size_t)(2 * k) * g.v3] = r * cs;
        dst[(size_t)(2 * k + 1) * g.v3] = r * sn;
    }
}
void launch_gaussian(cudaStream_t st, const Geom& g, double* p, unsigned long long seed, unsigned long long sweep, double sigma) {
    Keys4 keys;
    for (int mu = 0; mu < 4; mu++) host_stream_key(seed, sweep, (unsigned)(mu + 1), 0x47415553u, &keys.k0[mu], &keys.k1[mu]);
    k_gaussian<<<dim3(site_grid(g, 256), 4), 256, 0, st>>>(g, p, keys, sigma);
}

__global__ void __launch_bounds__(256) k_reunitarize(Geom g, double2* __restrict__ u) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int mu = blockIdx.y;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    double2* p = u + (size_t)(t * 36 + mu * 9) * g.v3 + s3;
    M3 m;
#pragma unroll
    for (int k = 0; k < 9; k++) m.e[k] = p[(size_t)k * g.v3];
    m = reunitarize(m);
#pragma unroll
    for (int k = 0; k < 9; k++) p[(size_t)k * g.v3] = m.e[k];
}
void launch_reunitarize(cudaStream_t st, const Geom& g, double2* u) { k_reunitarize<<<dim3(site_grid(g, 256), 4), 256, 0, st>>>(g, u); }

__global__ void __launch_bounds__(256) k_axpy(double* __restrict__ y, double a, const double* __restrict__ x, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fma(a, x[i], y[i]);
}
void launch_axpy(cudaStream_t st, double* y, double a, const double* x, size_t n) {
    size_t nb = (n + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    k_axpy<<<(unsigned)nb, 256, 0, st>>>(y, a, x, n);
}

// ------------------------------------------------------------------------------------------------
// derivative-field helpers for the generic (non-fused) force path
// ------------------------------------------------------------------------------------------------
// out_mu(x) = scale * sum of the six staples A with tr(loop) = tr(U_mu A), i.e. scale * V_mu(x)^dagger
// (calc_dSdUmu!, GaugeActions.jl:95-123 for the plaquette+plaquette' action with scale = beta/2)
template <bool FULL3>
__global__ void __launch_bounds__(128, 3) k_staple_field(Geom g, const double2* __restrict__ u, double2* __restrict__ out, double scale) {
    const int mu = threadIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = decode_site(g, n, 0, g.tloc);
    M3 s = staple_sum<FULL3>(u, g, x, mu);
    M3 d = m3_dagger(s);
#pragma unroll
    for (int k = 0; k < 9; k++) { d.e[k].x *= scale; d.e[k].y *= scale; }
    store_link(out, g, x, mu, d);
}
void launch_staple_field(cudaStream_t st, const Geom& g, const double2* u, double2* out, double scale, bool full3) {
    dim3 block(32, 4);
    long nsites = (long)g.v3 * g.tloc;
    if (full3) k_staple_field<true><<<(unsigned)((nsites + 31) / 32), block, 0, st>>>(g, u, out, scale);
    else k_staple_field<false><<<(unsigned)((nsites + 31) / 32), block, 0, st>>>(g, u, out, scale);
}

// P_mu += factor * TAcoeffs(U_mu * D_mu)  (md_force! tail, molecular_dynamics.jl:255-265)
__global__ void __launch_bounds__(256) k_kick_from_dsdu(Geom g, const double2* __restrict__ u, const double2* __restrict__ d, double* __restrict__ p,
                                                        double factor) {
    const int mu = blockIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const int t = (int)(n / g.v3);
    const int s3 = (int)(n - (long)t * g.v3);
    const size_t uo = (size_t)(t * 36 + mu * 9) * g.v3 + s3;
    M3 a = m3_load(u + uo, (unsigned)g.v3), b = m3_load(d + uo, (unsigned)g.v3);
    M3 w = mul_nn(a, b);
    double c[8];
    ta_coeffs(w, c);
    double* dst = p + (size_t)(t * 32 + mu * 8) * g.v3 + s3;
#pragma unroll
    for (int k = 0; k < 8; k++) dst[(size_t)k * g.v3] = fma(factor, c[k], dst[(size_t)k * g.v3]);
}
void launch_kick_from_dsdu(cudaStream_t st, const Geom& g, const double2* u, const double2* d, double* p, double factor) {
    k_kick_from_dsdu<<<dim3(site_grid(g, 256), 4), 256, 0, st>>>(g, u, d, p, factor);
}

}  // namespace gfb
