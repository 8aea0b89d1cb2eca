// general.cu -- the general-action path: plaquette + rectangle (Symanzik / Iwasaki / DBW2 type) actions and the loop
// observables built from the same path products (topological charge by the plaquette, clover and improved definitions).
//
// Reference behaviour: a GaugeAction holds terms (coefficient, loops + adjoint loops); calc_dSdUmu! sums coefficient x the
// staples obtained by deleting U_mu from every loop (src/action/GaugeActions.jl:95-123), F_update! / md_force! project
// U_mu dSdU_mu on the algebra (src/smearing/gradientflow.jl:318-334, src/molecular_dynamics.jl:251-267), and
// Gradientflow_general integrates that force with the same RK3 scheme as the Wilson flow (gradientflow.jl:240-316).  The
// "rectangular" loop set is the 1x2 and 2x1 rectangles of every plane (src/autostaples/wilsonloops.jl:233-245, 304-330).
// Topological charge: src/AbstractGaugefields.jl:1184-1400.
//
// Here the sum over terms is ONE staple field V = c_plaq V_plaq + c_rect V_rect evaluated inside the fused
// force -> (momentum / flow field) -> exp kernel, like the Wilson path (kernels.cu): one launch per kick or RK3 stage instead of
// ~40 whole-field kernels per rectangle staple.  Links are read with plain coalesced 128-bit loads (the t-marching tile kernel
// covers the plaquette stencil only); on SU(3) configurations (the handle's unitarity flag) the path products are two-row, otherwise
// full 3x3 as in the reference.  Rectangles reach two sites away, so
// on a t-slab decomposition the API first assembles a "wide" copy of the slab with two halo slices on either side in natural
// t order (api.cu, build_wide): the kernels then read links at wide slice t and write their results at slab slice t - t_shift.
#include "gfb_internal.h"
#include "stencil.cuh"
#include "su3.cuh"

namespace gfb {

namespace {

// m <- m * U_{+-dir}(y) and y moves along the step:  forward = U_dir(y), y += dir;  backward: y -= dir, U_dir(y)^dagger
__device__ __forceinline__ void path_step(M3& m, const double2* __restrict__ u, const Geom& g, Coord& y, int dir, int sgn) {
    if (sgn > 0) {
        const M3 l = load_link(u, g, y, dir);
        m = mul_nn(m, l);
        y = step(g, y, dir, +1);
    } else {
        y = step(g, y, dir, -1);
        const M3 l = load_link(u, g, y, dir);
        m = mul_nd(m, l);
    }
}
// first factor of a path (avoids multiplying the identity)
__device__ __forceinline__ M3 path_first(const double2* __restrict__ u, const Geom& g, Coord& y, int dir, int sgn) {
    if (sgn > 0) {
        const M3 l = load_link(u, g, y, dir);
        y = step(g, y, dir, +1);
        return l;
    }
    y = step(g, y, dir, -1);
    return m3_dagger(load_link(u, g, y, dir));
}

// the same for SU(3) links: rows 0,1 of the path product (72 instead of 108 FP64 instructions per factor); the third row is
// rebuilt when the finished path is accumulated (acc_su3), as in the Wilson kernels (su3.cuh)
__device__ __forceinline__ void path_step(R2& r, const double2* __restrict__ u, const Geom& g, Coord& y, int dir, int sgn) {
    if (sgn > 0) {
        const M3 l = load_link(u, g, y, dir);
        r = r2_mul_nn(r, l);
        y = step(g, y, dir, +1);
    } else {
        y = step(g, y, dir, -1);
        const M3 l = load_link(u, g, y, dir);
        r = r2_mul_nd(r, l);
    }
}
__device__ __forceinline__ R2 path_first_r2(const double2* __restrict__ u, const Geom& g, Coord& y, int dir, int sgn) {
    if (sgn > 0) {
        const R2 r = r2_load_rows01(u + link_offset(g, y, dir), (unsigned)g.v3);
        y = step(g, y, dir, +1);
        return r;
    }
    y = step(g, y, dir, -1);
    return r2_load_dag_rows01(u + link_offset(g, y, dir), (unsigned)g.v3);
}
template <bool FULL3>
struct PathOf { typedef M3 type; };
template <>
struct PathOf<false> { typedef R2 type; };
template <bool FULL3>
__device__ __forceinline__ typename PathOf<FULL3>::type path_begin(const double2* __restrict__ u, const Geom& g, Coord& y, int dir, int sgn) {
    if constexpr (FULL3) return path_first(u, g, y, dir, sgn);
    else return path_first_r2(u, g, y, dir, sgn);
}
__device__ __forceinline__ void path_end(M3& v, const M3& m) { m3_add(v, m); }
__device__ __forceinline__ void path_end(M3& v, const R2& r) { acc_su3(v, r); }

// Sum of the 18 rectangle staples of link (x, mu): every path S from x to x+mu such that U_mu(x) S^dagger is a 1x2 or 2x1
// rectangle (both orientations of the plane): for each nu != mu and s = +-1
//   (a) s nu, mu, mu, -s nu, -mu      (2x1, the long side ahead of the link)
//   (b) -mu, s nu, mu, mu, -s nu      (2x1, the long side behind the link)
//   (c) s nu, s nu, mu, -s nu, -s nu  (1x2)
template <bool FULL3>
__device__ __forceinline__ M3 rect_staple_sum(const double2* __restrict__ u, const Geom& g, const Coord& x, int mu) {
    M3 v = m3_zero();
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
        int nu = mu + 1 + j;
        if (nu >= 4) nu -= 4;
#pragma unroll 1
        for (int s = -1; s <= 1; s += 2) {
            {
                Coord y = x;
                auto m = path_begin<FULL3>(u, g, y, nu, s);
                path_step(m, u, g, y, mu, +1);
                path_step(m, u, g, y, mu, +1);
                path_step(m, u, g, y, nu, -s);
                path_step(m, u, g, y, mu, -1);
                path_end(v, m);
            }
            {
                Coord y = x;
                auto m = path_begin<FULL3>(u, g, y, mu, -1);
                path_step(m, u, g, y, nu, s);
                path_step(m, u, g, y, mu, +1);
                path_step(m, u, g, y, mu, +1);
                path_step(m, u, g, y, nu, -s);
                path_end(v, m);
            }
            {
                Coord y = x;
                auto m = path_begin<FULL3>(u, g, y, nu, s);
                path_step(m, u, g, y, nu, s);
                path_step(m, u, g, y, mu, +1);
                path_step(m, u, g, y, nu, -s);
                path_step(m, u, g, y, nu, -s);
                path_end(v, m);
            }
        }
    }
    return v;
}

__device__ __forceinline__ void m3_axpy(M3& acc, double a, const M3& m) {
#pragma unroll
    for (int k = 0; k < 9; k++) {
        acc.e[k].x = fma(a, m.e[k].x, acc.e[k].x);
        acc.e[k].y = fma(a, m.e[k].y, acc.e[k].y);
    }
}

// Z' = a * TAcoeffs(U_mu (c_plaq V_plaq + c_rect V_rect)^dag) + b * Z ;  Uout_mu = exp(c Z') Uin_mu
#ifndef GFB_GEN_MINB
#define GFB_GEN_MINB 3
#endif
template <bool READ_Z, bool WRITE_Z, bool DO_EXP, bool FULL3>
__global__ void __launch_bounds__(128, GFB_GEN_MINB)
k_force_general(Geom g, int t_begin, int t_count, int t_shift, const double2* __restrict__ uin, double2* __restrict__ uout, const double* __restrict__ zin,
                double* __restrict__ zout, double a, double b, double c, double c_plaq, double c_rect) {
    const int mu = threadIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * t_count) return;
    const Coord x = decode_site(g, n, t_begin, t_count);
    Coord xo = x;  // where the results go: the slab's own slice numbering
    xo.t -= t_shift;
    M3 v = m3_zero();
    if (c_plaq != 0.0) m3_axpy(v, c_plaq, staple_sum<FULL3>(uin, g, x, mu));
    if (c_rect != 0.0) m3_axpy(v, c_rect, rect_staple_sum<FULL3>(uin, g, x, mu));
    const M3 umu = load_link(uin, g, x, mu);
    double z[8];
    ta_coeffs_nd(umu, v, z);
    const unsigned zo = mom_offset(g, xo, mu);
    const unsigned zs = (unsigned)g.v3;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double w = a * z[k];
        if (READ_Z) w = fma(b, zin[zo + k * zs], w);
        z[k] = w;
        if (WRITE_Z) zout[zo + k * zs] = w;
    }
    if (DO_EXP) store_link(uout, g, xo, mu, FULL3 ? mul_nn(exp_ta(z, c), umu) : exp_ta_times_su3(z, c, umu));
}

// per site: sum_{mu<nu} Re tr P_munu  and  sum over the 12 rectangle loops of Re tr  (evaluate_GaugeAction's two building blocks)
__global__ void __launch_bounds__(128) k_loop_sums(Geom g, int t_begin, int t_count, const double2* __restrict__ u, double* __restrict__ partial, int nblocks) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double plaq = 0.0, rect = 0.0;
    if (n < (long)g.v3 * t_count) {
        const Coord x = decode_site(g, n, t_begin, t_count);
#pragma unroll 1
        for (int mu = 0; mu < 3; mu++) {
#pragma unroll 1
            for (int nu = mu + 1; nu < 4; nu++) {
                // A = U_mu(x) U_nu(x+mu), B = U_nu(x) U_mu(x+nu): plaquette = A B^dag
                Coord y = x;
                M3 a = path_first(u, g, y, mu, +1);
                const Coord xpm = y;
                path_step(a, u, g, y, nu, +1);
                y = x;
                M3 bb = path_first(u, g, y, nu, +1);
                const Coord xpn = y;
                path_step(bb, u, g, y, mu, +1);
                plaq += retr_nd(a, bb);
                // (mu,1)(nu,2)(mu,-1)(nu,-2):  [U_mu(x) U_nu(x+mu) U_nu(x+mu+nu)] [U_nu(x) U_nu(x+nu) U_mu(x+2nu)]^dag
                {
                    M3 l = a;
                    Coord z1 = step(g, xpm, nu, +1);
                    path_step(l, u, g, z1, nu, +1);
                    Coord z2 = xpn;
                    M3 r = load_link(u, g, x, nu);
                    path_step(r, u, g, z2, nu, +1);
                    path_step(r, u, g, z2, mu, +1);
                    rect += retr_nd(l, r);
                }
                // (mu,2)(nu,1)(mu,-2)(nu,-1):  [U_mu(x) U_mu(x+mu) U_nu(x+2mu)] [U_nu(x) U_mu(x+nu) U_mu(x+nu+mu)]^dag
                {
                    Coord z1 = xpm;
                    M3 l = load_link(u, g, x, mu);
                    path_step(l, u, g, z1, mu, +1);
                    path_step(l, u, g, z1, nu, +1);
                    M3 r = bb;
                    Coord z2 = step(g, xpn, mu, +1);
                    path_step(r, u, g, z2, mu, +1);
                    rect += retr_nd(l, r);
                }
            }
        }
    }
    const double rp = block_sum(plaq);
    const double rr = block_sum(rect);
    if (threadIdx.x == 0) { partial[blockIdx.x] = rp; partial[nblocks + blockIdx.x] = rr; }
}

// closed loop from x along four signed segments (d0, n0) (d1, n1) (d2, n2) (d3, n3), |n| steps each
__device__ __forceinline__ M3 loop4(const double2* __restrict__ u, const Geom& g, const Coord& x, int d0, int n0, int d1, int n1, int d2, int n2, int d3, int n3) {
    Coord y = x;
    const int d[4] = {d0, d1, d2, d3}, n[4] = {n0, n1, n2, n3};
    M3 m = path_first(u, g, y, d[0], n[0] > 0 ? 1 : -1);
    for (int k = 1; k < (n[0] > 0 ? n[0] : -n[0]); k++) path_step(m, u, g, y, d[0], n[0] > 0 ? 1 : -1);
#pragma unroll 1
    for (int i = 1; i < 4; i++) {
        const int sg = n[i] > 0 ? 1 : -1, cnt = n[i] > 0 ? n[i] : -n[i];
        for (int k = 0; k < cnt; k++) path_step(m, u, g, y, d[i], sg);
    }
    return m;
}

// TA coefficients of the field strength F_munu(x) by one of three loop sets (AbstractGaugefields.jl:1184-1351):
//   0 plaquette: the loop (mu,1)(nu,1)(mu,-1)(nu,-1);  1 clover: make_cloverloops (src/autostaples/wilsonloops.jl:166-177);
//   2 rectangle: _rectangle_loops (AbstractGaugefields.jl:1316-1330)
__device__ __forceinline__ void field_strength(const double2* __restrict__ u, const Geom& g, const Coord& x, int mu, int nu, int kind, double* c) {
    M3 w;
    if (kind == 0) {
        w = loop4(u, g, x, mu, 1, nu, 1, mu, -1, nu, -1);
    } else if (kind == 1) {
        w = loop4(u, g, x, mu, 1, nu, 1, mu, -1, nu, -1);
        m3_add(w, loop4(u, g, x, nu, 1, mu, -1, nu, -1, mu, 1));
        m3_add(w, loop4(u, g, x, nu, -1, mu, 1, nu, 1, mu, -1));
        m3_add(w, loop4(u, g, x, mu, -1, nu, -1, mu, 1, nu, 1));
    } else {
        w = loop4(u, g, x, mu, 2, nu, 1, mu, -2, nu, -1);
        m3_add(w, loop4(u, g, x, nu, 1, mu, -2, nu, -1, mu, 2));
        m3_add(w, loop4(u, g, x, nu, -1, mu, 2, nu, 1, mu, -2));
        m3_add(w, loop4(u, g, x, mu, -2, nu, -1, mu, 2, nu, 1));
        m3_add(w, loop4(u, g, x, mu, 1, nu, 2, mu, -1, nu, -2));
        m3_add(w, loop4(u, g, x, nu, 2, mu, -1, nu, -2, mu, 1));
        m3_add(w, loop4(u, g, x, nu, -2, mu, 1, nu, 2, mu, -1));
        m3_add(w, loop4(u, g, x, mu, -1, nu, -2, mu, 1, nu, 2));
    }
    ta_coeffs(w, c);
}

// q(x) = -Re sum_{mu nu rho sigma} eps tr(F_munu F_rhosigma) / (32 pi^2 n^2), n = loops per field strength.  With
// F = sum_a c_a i lambda_a / 2:  tr(F F') = -(1/2) sum_a c_a c'_a, and the 24 permutations are 8 x (01|23) - (02|13) + (03|12).
// density[site] (host order x fastest, then y, z, local t) gets `weight` times the kind's density added (improved = 5/3 clover - 1/12 rectangle)
__global__ void __launch_bounds__(128) k_topological_density(Geom g, int t_begin, int t_count, int t_shift, const double2* __restrict__ u, double* __restrict__ density, int kind,
                                                             double weight, int accumulate) {
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * t_count) return;
    const Coord x = decode_site(g, n, t_begin, t_count);
    const int pairs[3][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2}};
    double q = 0.0;
#pragma unroll 1
    for (int p = 0; p < 3; p++) {
        double ca[8], cb[8];
        field_strength(u, g, x, pairs[p][0], pairs[p][1], kind, ca);
        field_strength(u, g, x, pairs[p][2], pairs[p][3], kind, cb);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) s = fma(ca[k], cb[k], s);
        q += (p == 1) ? -s : s;
    }
    const double nl = kind == 0 ? 1.0 : (kind == 1 ? 4.0 : 8.0);
    const double rect_factor = kind == 2 ? 2.0 : 1.0;
    // -Re(8 * (-1/2) * q) / (32 pi^2 n^2)
    const double val = weight * rect_factor * 4.0 * q / (32.0 * 9.869604401089358 * nl * nl);
    const size_t idx = (size_t)s3_of(g, x) + (size_t)g.v3 * (x.t - t_shift);
    density[idx] = accumulate ? density[idx] + val : val;
}

__global__ void __launch_bounds__(256) k_sum_plain(const double* __restrict__ v, size_t n, double* __restrict__ partial) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += v[i];
    const double r = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

}  // namespace

void launch_force_general(cudaStream_t st, const Geom& g, int t_begin, int t_count, int t_shift, const double2* uin, double2* uout, const double* zin, double* zout,
                          const FusedArgs& fa) {
    const long nsites = (long)g.v3 * t_count;
    if (nsites <= 0) return;
    dim3 block(32, 4), grid((unsigned)((nsites + 31) / 32));
#define GFB_LAUNCH_FG(R, W, E)                                                                                                                      \
    do {                                                                                                                                             \
        if (fa.full3) k_force_general<R, W, E, true><<<grid, block, 0, st>>>(g, t_begin, t_count, t_shift, uin, uout, zin, zout, fa.a, fa.b, fa.c, fa.c_plaq, fa.c_rect); \
        else k_force_general<R, W, E, false><<<grid, block, 0, st>>>(g, t_begin, t_count, t_shift, uin, uout, zin, zout, fa.a, fa.b, fa.c, fa.c_plaq, fa.c_rect);       \
    } while (0)
    if (fa.read_z) {
        if (fa.do_exp) GFB_LAUNCH_FG(true, true, true);
        else GFB_LAUNCH_FG(true, true, false);
    } else {
        if (fa.do_exp) {
            if (fa.write_z) GFB_LAUNCH_FG(false, true, true);
            else GFB_LAUNCH_FG(false, false, true);
        } else GFB_LAUNCH_FG(false, true, false);
    }
#undef GFB_LAUNCH_FG
}

void launch_loop_sums(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* u, double* partial, int* nblocks) {
    const int nb = (int)(((long)g.v3 * t_count + 127) / 128);
    k_loop_sums<<<nb, 128, 0, st>>>(g, t_begin, t_count, u, partial, nb);
    *nblocks = nb;
}

void launch_topological_density(cudaStream_t st, const Geom& g, int t_begin, int t_count, int t_shift, const double2* u, double* density, int kind, double weight,
                                bool accumulate) {
    const int nb = (int)(((long)g.v3 * t_count + 127) / 128);
    k_topological_density<<<nb, 128, 0, st>>>(g, t_begin, t_count, t_shift, u, density, kind, weight, accumulate ? 1 : 0);
}

void launch_sum_plain(cudaStream_t st, const double* v, size_t n, double* partial, int* nblocks) {
    int nb = (int)((n + 255) / 256);
    if (nb > 1024) nb = 1024;
    if (nb < 1) nb = 1;
    k_sum_plain<<<nb, 256, 0, st>>>(v, n, partial);
    *nblocks = nb;
}

}  // namespace gfb
