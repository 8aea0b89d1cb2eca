// stencil.cuh -- device helpers shared by kernels.cu and stout.cu: link access, the six-staple sum and the
// deterministic block reduction.
#pragma once
#include "lattice.cuh"
#include "su3.cuh"

#ifndef GFB_FF_UNROLL
#define GFB_FF_UNROLL 1
#endif

namespace gfb {

constexpr int kStapleUnroll = GFB_FF_UNROLL;

__device__ __forceinline__ M3 load_link(const double2* __restrict__ u, const Geom& g, const Coord& c, int mu) {
    return m3_load(u + link_offset(g, c, mu), (unsigned)g.v3);
}
__device__ __forceinline__ void store_link(double2* __restrict__ u, const Geom& g, const Coord& c, int mu, const M3& m) {
    m3_store(u + link_offset(g, c, mu), (unsigned)g.v3, m);
}

// deterministic block sum (fixed shuffle tree + fixed-order warp combine); result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    const int nthreads = blockDim.x * blockDim.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_part[tid >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (tid < 32) {
        r = (tid < (nthreads + 31) / 32) ? warp_part[tid] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

// U_mu(x) * V_mu(x)^dagger with V_mu the sum of the six plaquette staples
//   V_mu = sum_{nu != mu} [ U_nu(x) U_mu(x+nu) U_nu(x+mu)^dag + U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu) ]
// (src/autostaples/wilsonloops.jl:468-484, construct_double_staple! src/AbstractGaugefields.jl:2856-2871)
#ifndef GFB_FF_NOMATH
#define GFB_FF_NOMATH 0
#endif
#ifndef GFB_FF_XOR
#define GFB_FF_XOR 0
#endif
#ifndef GFB_FF_PIPE
#define GFB_FF_PIPE 0  // 0: compiler-scheduled loads; 1: all three operands of a staple requested up front; 2: next staple prefetched in registers
#endif

// 128-bit read-only load that the compiler may not sink towards its first use
__device__ __forceinline__ double2 ldg_pinned(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ M3 load_link_pinned(const double2* __restrict__ u, const Geom& g, const Coord& c, int mu) {
    M3 r;
    const char* b = reinterpret_cast<const char*>(u + link_offset(g, c, mu));
    const unsigned sb = (unsigned)g.v3 * 16u;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = ldg_pinned(reinterpret_cast<const double2*>(b + (size_t)k * sb));
    return r;
}

struct StapleOps {
    M3 a, b, c;
};
// operands of staple j (0..5) of link (x, mu): j>>1 picks nu = (mu+1+(j>>1)) mod 4, j&1 = 0 upper / 1 lower
__device__ __forceinline__ StapleOps fetch_staple(const double2* __restrict__ u, const Geom& g, const Coord& x, const Coord& xm, int mu, int j) {
#if GFB_FF_XOR
    const int nu = mu ^ ((j >> 1) + 1);  // the two links of a plane visit it in the same step: their shared operands hit L1 together
#else
    int nu = mu + 1 + (j >> 1);
    if (nu >= 4) nu -= 4;
#endif
    StapleOps o;
    if ((j & 1) == 0) {
        const Coord xn = step(g, x, nu, +1);
        o.a = load_link_pinned(u, g, x, nu);
        o.b = load_link_pinned(u, g, xn, mu);
        o.c = load_link_pinned(u, g, xm, nu);
    } else {
        const Coord xd = step(g, x, nu, -1);
        const Coord xdm = step(g, xd, mu, +1);
        o.a = load_link_pinned(u, g, xd, nu);
        o.b = load_link_pinned(u, g, xd, mu);
        o.c = load_link_pinned(u, g, xdm, nu);
    }
    return o;
}
__device__ __forceinline__ void accumulate_staple(M3& s, const StapleOps& o, int j) {
#if GFB_FF_NOMATH  // diagnostic build: memory traffic only (profiles/r1_ncu_force_fused.md)
    m3_add(s, o.a); m3_add(s, o.b); m3_add(s, o.c);
    return;
#endif
    if ((j & 1) == 0) {
        M3 t = mul_nn(o.a, o.b);
        mac_nd(s, t, o.c);
    } else {
        M3 t = mul_dn(o.a, o.b);
        mac_nn(s, t, o.c);
    }
}

// FULL3 = true: generic 3x3 products for links that are not (yet) unitary -- exactly the reference's staples
template <bool FULL3 = false>
__device__ __forceinline__ M3 staple_sum(const double2* __restrict__ u, const Geom& g, const Coord& x, int mu) {
    M3 s = m3_zero();
    const Coord xm = step(g, x, mu, +1);
    if (FULL3) {
#pragma unroll 1
        for (int j = 0; j < 6; j++) {
            const StapleOps o = fetch_staple(u, g, x, xm, mu, j);
            accumulate_staple(s, o, j);
        }
        return s;
    }
#if GFB_FF_PIPE == 1
#pragma unroll 1
    for (int j = 0; j < 6; j++) {
        const StapleOps o = fetch_staple(u, g, x, xm, mu, j);
        accumulate_staple(s, o, j);
    }
#elif GFB_FF_PIPE == 2
    StapleOps cur = fetch_staple(u, g, x, xm, mu, 0);
#pragma unroll
    for (int j = 0; j < 6; j++) {
        StapleOps nxt;
        if (j < 5) nxt = fetch_staple(u, g, x, xm, mu, j + 1);
        accumulate_staple(s, cur, j);
        if (j < 5) cur = nxt;
    }
#else
#pragma unroll kStapleUnroll
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        // two-row products (su3.cuh): the links are unitary, so rows 0,1 of A (6 of its 9 elements) carry the staple and its
        // third row is rebuilt inside the accumulation
        {
            const Coord xn = step(g, x, nu, +1);
            const R2 a = r2_load_rows01(u + link_offset(g, x, nu), (unsigned)g.v3);
            const M3 b = load_link(u, g, xn, mu);
            const R2 t = r2_mul_nn(a, b);
            const M3 c = load_link(u, g, xm, nu);
            acc_su3(s, r2_mul_nd(t, c));
        }
        {
            const Coord xd = step(g, x, nu, -1);
            const Coord xdm = step(g, xd, mu, +1);
            const R2 a = r2_load_dag_rows01(u + link_offset(g, xd, nu), (unsigned)g.v3);
            const M3 b = load_link(u, g, xd, mu);
            const R2 t = r2_mul_nn(a, b);
            const M3 c = load_link(u, g, xdm, nu);
            acc_su3(s, r2_mul_nn(t, c));
        }
    }
#endif
    return s;
}

}  // namespace gfb
