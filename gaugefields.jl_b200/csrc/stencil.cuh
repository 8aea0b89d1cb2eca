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
__device__ __forceinline__ M3 staple_sum(const double2* __restrict__ u, const Geom& g, const Coord& x, int mu) {
    M3 s = m3_zero();
    const Coord xm = step(g, x, mu, +1);
#pragma unroll kStapleUnroll
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        {
            const Coord xn = step(g, x, nu, +1);
            M3 a = load_link(u, g, x, nu);
            M3 b = load_link(u, g, xn, mu);
            M3 t = mul_nn(a, b);
            a = load_link(u, g, xm, nu);
            mac_nd(s, t, a);
        }
        {
            const Coord xd = step(g, x, nu, -1);
            const Coord xdm = step(g, xd, mu, +1);
            M3 a = load_link(u, g, xd, nu);
            M3 b = load_link(u, g, xd, mu);
            M3 t = mul_dn(a, b);
            a = load_link(u, g, xdm, nu);
            mac_nn(s, t, a);
        }
    }
    return s;
}

}  // namespace gfb
