// su3.cuh -- register-resident 3x3 complex fp64 algebra for sm_100a.
//
// A matrix is 9 double2 (re, im) values, row-major (element (i,j) at 3*i+j), always indexed
// with compile-time constants so that it lives in registers.  Device memory holds one
// double2 plane per element (structure of arrays), so every load/store below is a 128-bit
// access that is perfectly coalesced across the sites of a warp.
//
// Conventions follow the reference (file:line relative to /root/reference):
//   * TA projection to Gell-Mann coefficients  -- src/4D/TA_gaugefields_4D_serial.jl:181-269
//   * Hermitian matrix from coefficients       -- src/4D/TA_gaugefields_4D_serial.jl:779-847
//   * exp(t * sum_a c_a i lambda_a/2)          -- exptU!, src/4D/TA_gaugefields_4D_serial.jl:760-848
//     (evaluated here by a Cayley-Hamilton Horner recursion instead of the legacy
//      eigen-decomposition; see DESIGN.md "SU(3) exponential")
#pragma once
#include <cuda_runtime.h>

namespace gfb {

struct M3 {
    double2 e[9];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc += a*b
__device__ __forceinline__ void cmac(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += a*conj(b)
__device__ __forceinline__ void cmac_c(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.y, b.x, acc.y); acc.y = fma(-a.x, b.y, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cmac_ca(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

__device__ __forceinline__ M3 m3_zero() {
    M3 r;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = make_double2(0.0, 0.0);
    return r;
}
__device__ __forceinline__ M3 m3_identity() {
    M3 r = m3_zero();
    r.e[0].x = 1.0; r.e[4].x = 1.0; r.e[8].x = 1.0;
    return r;
}

// C = A*B
__device__ __forceinline__ M3 mul_nn(const M3& a, const M3& b) {
    M3 c;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = cmul(a.e[3 * i], b.e[j]);
            cmac(s, a.e[3 * i + 1], b.e[3 + j]);
            cmac(s, a.e[3 * i + 2], b.e[6 + j]);
            c.e[3 * i + j] = s;
        }
    return c;
}
// C += A*B
__device__ __forceinline__ void mac_nn(M3& c, const M3& a, const M3& b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = c.e[3 * i + j];
            cmac(s, a.e[3 * i], b.e[j]);
            cmac(s, a.e[3 * i + 1], b.e[3 + j]);
            cmac(s, a.e[3 * i + 2], b.e[6 + j]);
            c.e[3 * i + j] = s;
        }
}
// C += A*B^dagger
__device__ __forceinline__ void mac_nd(M3& c, const M3& a, const M3& b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = c.e[3 * i + j];
            cmac_c(s, a.e[3 * i], b.e[3 * j]);
            cmac_c(s, a.e[3 * i + 1], b.e[3 * j + 1]);
            cmac_c(s, a.e[3 * i + 2], b.e[3 * j + 2]);
            c.e[3 * i + j] = s;
        }
}
// C = A*B^dagger
__device__ __forceinline__ M3 mul_nd(const M3& a, const M3& b) {
    M3 c = m3_zero();
    mac_nd(c, a, b);
    return c;
}
// C = A^dagger*B
__device__ __forceinline__ M3 mul_dn(const M3& a, const M3& b) {
    M3 c;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = make_double2(0.0, 0.0);
            cmac_ca(s, a.e[i], b.e[j]);
            cmac_ca(s, a.e[3 + i], b.e[3 + j]);
            cmac_ca(s, a.e[6 + i], b.e[6 + j]);
            c.e[3 * i + j] = s;
        }
    return c;
}
// C += A^dagger * B
__device__ __forceinline__ void mac_dn(M3& c, const M3& a, const M3& b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = c.e[3 * i + j];
            cmac_ca(s, a.e[i], b.e[j]);
            cmac_ca(s, a.e[3 + i], b.e[3 + j]);
            cmac_ca(s, a.e[6 + i], b.e[6 + j]);
            c.e[3 * i + j] = s;
        }
}
// Re tr(A*B^dagger)
__device__ __forceinline__ double retr_nd(const M3& a, const M3& b) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 9; k++) { s = fma(a.e[k].x, b.e[k].x, s); s = fma(a.e[k].y, b.e[k].y, s); }
    return s;
}
// tr(A*B)
__device__ __forceinline__ double2 tr_nn(const M3& a, const M3& b) {
    double2 s = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) cmac(s, a.e[3 * i + j], b.e[3 * j + i]);
    return s;
}
__device__ __forceinline__ M3 m3_dagger(const M3& a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.e[3 * i + j] = make_double2(a.e[3 * j + i].x, -a.e[3 * j + i].y);
    return r;
}
__device__ __forceinline__ void m3_add(M3& c, const M3& a) {
#pragma unroll
    for (int k = 0; k < 9; k++) { c.e[k].x += a.e[k].x; c.e[k].y += a.e[k].y; }
}

// ---- two-row arithmetic for SU(3) operands ---------------------------------------------------------------------------
// A product of unitary links is unitary, so only its rows 0 and 1 are computed (2 x 72 FMA for a staple A B C from the first
// two rows of A) and row 2 = conj(row0 x row1) is folded into the accumulation of the staple sum: 180 instead of 216 FP64
// instructions per staple.  Used by both fused force kernels; every caller passes gauge links (unitary to rounding).
struct R2 {
    double2 e[6];  // rows 0 and 1
};
__device__ __forceinline__ R2 rows01(const M3& a) {
    R2 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.e[k] = a.e[k];
    return r;
}
__device__ __forceinline__ R2 rows01_dag(const M3& a) {
    R2 r;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.e[3 * i + j] = make_double2(a.e[3 * j + i].x, -a.e[3 * j + i].y);
    return r;
}
// (2x3) * (3x3)
__device__ __forceinline__ R2 r2_mul_nn(const R2& a, const M3& b) {
    R2 c;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = cmul(a.e[3 * i], b.e[j]);
            cmac(s, a.e[3 * i + 1], b.e[3 + j]);
            cmac(s, a.e[3 * i + 2], b.e[6 + j]);
            c.e[3 * i + j] = s;
        }
    return c;
}
// (2x3) * (3x3)^dagger
__device__ __forceinline__ R2 r2_mul_nd(const R2& a, const M3& b) {
    R2 c;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double2 s = make_double2(0.0, 0.0);
            cmac_c(s, a.e[3 * i], b.e[3 * j]);
            cmac_c(s, a.e[3 * i + 1], b.e[3 * j + 1]);
            cmac_c(s, a.e[3 * i + 2], b.e[3 * j + 2]);
            c.e[3 * i + j] = s;
        }
    return c;
}
// acc += conj(a*b - c*d)
__device__ __forceinline__ void cross_acc(double2& acc, double2 a, double2 b, double2 c, double2 d) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x); acc.x = fma(-c.x, d.x, acc.x); acc.x = fma(c.y, d.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y); acc.y = fma(c.x, d.y, acc.y); acc.y = fma(c.y, d.x, acc.y);
}
// v += the SU(3) matrix whose rows 0,1 are r (row 2 = conj(row0 x row1))
__device__ __forceinline__ void acc_su3(M3& v, const R2& r) {
#pragma unroll
    for (int k = 0; k < 6; k++) { v.e[k].x += r.e[k].x; v.e[k].y += r.e[k].y; }
    cross_acc(v.e[6], r.e[1], r.e[5], r.e[2], r.e[4]);
    cross_acc(v.e[7], r.e[2], r.e[3], r.e[0], r.e[5]);
    cross_acc(v.e[8], r.e[0], r.e[4], r.e[1], r.e[3]);
}
__device__ __forceinline__ M3 complete_su3(const R2& r) {
    M3 v;
#pragma unroll
    for (int k = 0; k < 6; k++) v.e[k] = r.e[k];
    v.e[6] = v.e[7] = v.e[8] = make_double2(0.0, 0.0);
    cross_acc(v.e[6], r.e[1], r.e[5], r.e[2], r.e[4]);
    cross_acc(v.e[7], r.e[2], r.e[3], r.e[0], r.e[5]);
    cross_acc(v.e[8], r.e[0], r.e[4], r.e[1], r.e[3]);
    return v;
}

// rows 0,1 of the link stored structure-of-arrays at p (element k at p + k*plane)
__device__ __forceinline__ R2 r2_load_rows01(const double2* __restrict__ p, unsigned plane) {
    R2 r;
    const char* b = reinterpret_cast<const char*>(p);
    const unsigned sb = plane * 16u;
#pragma unroll
    for (int k = 0; k < 6; k++) r.e[k] = __ldg(reinterpret_cast<const double2*>(b + (size_t)k * sb));
    return r;
}
// rows 0,1 of its adjoint: (A^dag)[i][j] = conj(A[j][i])
__device__ __forceinline__ R2 r2_load_dag_rows01(const double2* __restrict__ p, unsigned plane) {
    R2 r;
    const char* b = reinterpret_cast<const char*>(p);
    const unsigned sb = plane * 16u;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(b + (size_t)(3 * j + i) * sb));
            r.e[3 * i + j] = make_double2(v.x, -v.y);
        }
    return r;
}

// structure-of-arrays access: element k of the matrix lives `k*plane` elements after element 0.
// The address of element k is formed as base + k*(plane*16) with a 32-bit byte stride, which ptxas
// turns into ONE IMAD.WIDE.U32 per 128-bit access (64-bit index arithmetic costs 4 integer
// instructions per load and steals issue slots from the FP64 pipe).
__device__ __forceinline__ M3 m3_load(const double2* __restrict__ p, unsigned plane) {
    M3 r;
    const char* b = reinterpret_cast<const char*>(p);
    const unsigned sb = plane * 16u;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = __ldg(reinterpret_cast<const double2*>(b + (size_t)k * sb));
    return r;
}
__device__ __forceinline__ M3 m3_load_rw(const double2* p, unsigned plane) {
    M3 r;
    const char* b = reinterpret_cast<const char*>(p);
    const unsigned sb = plane * 16u;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = *reinterpret_cast<const double2*>(b + (size_t)k * sb);
    return r;
}
__device__ __forceinline__ void m3_store(double2* __restrict__ p, unsigned plane, const M3& m) {
    char* b = reinterpret_cast<char*>(p);
    const unsigned sb = plane * 16u;
#pragma unroll
    for (int k = 0; k < 9; k++) *reinterpret_cast<double2*>(b + (size_t)k * sb) = m.e[k];
}

#define GFB_SR3I 0.57735026918962576451  // 1/sqrt(3)

// 8 Gell-Mann coefficients of the traceless anti-Hermitian part of M
// (src/4D/TA_gaugefields_4D_serial.jl:181-269):  TA(M) = sum_a c_a * i*lambda_a/2
__device__ __forceinline__ void ta_coeffs(const M3& m, double* c) {
    // y = (m - m^dagger)/2 ; diagonal is purely imaginary
    double d0 = m.e[0].y, d1 = m.e[4].y, d2 = m.e[8].y;
    double tri = (d0 + d1 + d2) * (1.0 / 3.0);
    d0 -= tri; d1 -= tri; d2 -= tri;
    // y12 = (m12 - conj(m21))/2 ; y21 = -conj(y12)
    double y01r = 0.5 * (m.e[1].x - m.e[3].x), y01i = 0.5 * (m.e[1].y + m.e[3].y);
    double y02r = 0.5 * (m.e[2].x - m.e[6].x), y02i = 0.5 * (m.e[2].y + m.e[6].y);
    double y12r = 0.5 * (m.e[5].x - m.e[7].x), y12i = 0.5 * (m.e[5].y + m.e[7].y);
    c[0] = 2.0 * y01i;
    c[1] = 2.0 * y01r;
    c[2] = d0 - d1;
    c[3] = 2.0 * y02i;
    c[4] = 2.0 * y02r;
    c[5] = 2.0 * y12i;
    c[6] = 2.0 * y12r;
    c[7] = GFB_SR3I * (d0 + d1 - 2.0 * d2);
}

// ta_coeffs(U * V^dagger) without forming the product: only the anti-Hermitian part of W = U V^dag is needed --
// Im W_ii (6 FMA each) and the three pairs W_ij, W_ji (12 FMA each): 90 FP64 instructions instead of 108 + the projection
__device__ __forceinline__ void ta_coeffs_nd(const M3& u, const M3& v, double* c) {
    double d[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) { s = fma(u.e[3 * i + k].y, v.e[3 * i + k].x, s); s = fma(-u.e[3 * i + k].x, v.e[3 * i + k].y, s); }
        d[i] = s;
    }
    auto w = [&](int i, int j) {
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 3; k++) cmac_c(s, u.e[3 * i + k], v.e[3 * j + k]);
        return s;
    };
    const double2 w01 = w(0, 1), w10 = w(1, 0), w02 = w(0, 2), w20 = w(2, 0), w12 = w(1, 2), w21 = w(2, 1);
    const double tri = (d[0] + d[1] + d[2]) * (1.0 / 3.0);
    const double d0 = d[0] - tri, d1 = d[1] - tri, d2 = d[2] - tri;
    c[0] = w01.y + w10.y;
    c[1] = w01.x - w10.x;
    c[2] = d0 - d1;
    c[3] = w02.y + w20.y;
    c[4] = w02.x - w20.x;
    c[5] = w12.y + w21.y;
    c[6] = w12.x - w21.x;
    c[7] = GFB_SR3I * (d0 + d1 - 2.0 * d2);
}

// matrix-valued TA: Q = (M - M^dagger)/2 - tr/3 (src/4D/nowing/gaugefields_4D_nowing.jl:1253-1345)
__device__ __forceinline__ M3 ta_matrix(const M3& m) {
    M3 q;
    double tri = (m.e[0].y + m.e[4].y + m.e[8].y) * (1.0 / 3.0);
    q.e[0] = make_double2(0.0, m.e[0].y - tri);
    q.e[4] = make_double2(0.0, m.e[4].y - tri);
    q.e[8] = make_double2(0.0, m.e[8].y - tri);
    q.e[1] = make_double2(0.5 * (m.e[1].x - m.e[3].x), 0.5 * (m.e[1].y + m.e[3].y));
    q.e[2] = make_double2(0.5 * (m.e[2].x - m.e[6].x), 0.5 * (m.e[2].y + m.e[6].y));
    q.e[5] = make_double2(0.5 * (m.e[5].x - m.e[7].x), 0.5 * (m.e[5].y + m.e[7].y));
    q.e[3] = make_double2(-q.e[1].x, q.e[1].y);
    q.e[6] = make_double2(-q.e[2].x, q.e[2].y);
    q.e[7] = make_double2(-q.e[5].x, q.e[5].y);
    return q;
}

// coefficients -> anti-Hermitian matrix sum_a c_a i lambda_a / 2
__device__ __forceinline__ M3 ta_from_coeffs(const double* c) {
    M3 q;
    double h00 = 0.5 * (c[2] + GFB_SR3I * c[7]);
    double h11 = 0.5 * (-c[2] + GFB_SR3I * c[7]);
    double h22 = -GFB_SR3I * c[7];
    q.e[0] = make_double2(0.0, h00);
    q.e[4] = make_double2(0.0, h11);
    q.e[8] = make_double2(0.0, h22);
    // H01 = (c1 - i c2)/2 -> i*H01 = (c2 + i c1)/2
    q.e[1] = make_double2(0.5 * c[1], 0.5 * c[0]);
    q.e[2] = make_double2(0.5 * c[4], 0.5 * c[3]);
    q.e[5] = make_double2(0.5 * c[6], 0.5 * c[5]);
    q.e[3] = make_double2(-q.e[1].x, q.e[1].y);
    q.e[6] = make_double2(-q.e[2].x, q.e[2].y);
    q.e[7] = make_double2(-q.e[5].x, q.e[5].y);
    return q;
}


// Hermitian traceless Q in compact form: 3 real diagonal + 3 complex upper entries.
struct H3 {
    double d0, d1, d2;
    double2 o01, o02, o12;
};

__device__ __forceinline__ H3 h3_from_coeffs(const double* c, double t) {
    H3 q;
    double s = 0.5 * t;
    q.d0 = s * (c[2] + GFB_SR3I * c[7]);
    q.d1 = s * (-c[2] + GFB_SR3I * c[7]);
    q.d2 = -2.0 * s * GFB_SR3I * c[7];
    q.o01 = make_double2(s * c[0], -s * c[1]);
    q.o02 = make_double2(s * c[3], -s * c[4]);
    q.o12 = make_double2(s * c[5], -s * c[6]);
    return q;
}

// f0,f1,f2 with exp(iQ) = f0 + f1 Q + f2 Q^2 for Hermitian traceless Q with
// c0 = det Q, c1 = tr(Q^2)/2 (Q^3 = c1 Q + c0).  Horner evaluation of the Taylor series
// reduced with the Cayley-Hamilton relation; valid for c1 <= 0.75 (|eigenvalues| <= 1).
// Fully unrolled with immediate coefficients: the recursion is linear with REAL coefficients (c0, c1), so the real
// parts (driven by the even a_n = i^n/n!) and the imaginary parts (odd n) are two independent real recursions of
// 2 FMA per term each.  (The first version looped with a constant-memory table and per-term selects: ~45 instructions
// per term and the top stall of the t-marching kernel -- profiles/r1_tmarch.md.)
template <int n>
struct InvFactorial {
    static constexpr double v = InvFactorial<n - 1>::v / (double)n;
};
template <>
struct InvFactorial<0> {
    static constexpr double v = 1.0;
};
// one Horner term:  p <- a_n + Q p  reduced mod Q^3 = c1 Q + c0:  (p0, p1, p2) <- (a_n + c0 p2, p0 + c1 p2, p1)
template <int n>
__device__ __forceinline__ void ch_terms(double c0, double c1, double2& p0, double2& p1, double2& p2) {
    constexpr double a = ((n & 2) ? -1.0 : 1.0) * InvFactorial<n>::v;  // a_n = i^n/n!: real for even n, imaginary for odd n
    double2 n0, n1;
    if ((n & 1) == 0) { n0.x = fma(c0, p2.x, a); n0.y = c0 * p2.y; }
    else { n0.x = c0 * p2.x; n0.y = fma(c0, p2.y, a); }
    n1.x = fma(c1, p2.x, p0.x);
    n1.y = fma(c1, p2.y, p0.y);
    p2 = p1; p1 = n1; p0 = n0;
    if constexpr (n > 0) ch_terms<n - 1>(c0, c1, p0, p1, p2);
}
template <int N>
__device__ __forceinline__ void ch_coefficients_n(double c0, double c1, double2& f0, double2& f1, double2& f2) {
    double2 p0 = make_double2(0.0, 0.0), p1 = make_double2(0.0, 0.0), p2 = make_double2(0.0, 0.0);
    ch_terms<N>(c0, c1, p0, p1, p2);
    f0 = p0; f1 = p1; f2 = p2;
}
// Truncation: the largest |eigenvalue| x of Q obeys x^2 <= 4 c1 / 3, and the remainder after N terms is below x^(N+1)/(N+1)!:
//   c1 <= 0.01 (x <= 0.116): N = 10 -> 1e-18;  c1 <= 3/64 (x <= 0.25): N = 12 -> 2.4e-18;  c1 <= 3/4 (x <= 1): N = 19 -> 4e-19
__device__ __forceinline__ void ch_coefficients(double c0, double c1, double2& f0, double2& f1, double2& f2) {
    if (c1 <= 0.01) ch_coefficients_n<10>(c0, c1, f0, f1, f2);
    else if (c1 <= 0.046875) ch_coefficients_n<12>(c0, c1, f0, f1, f2);
    else ch_coefficients_n<19>(c0, c1, f0, f1, f2);
}

// E = exp(i Q).  Arguments with spectral radius > 1 are scaled by 2^-s and squared back.
__device__ __forceinline__ M3 exp_iH(const H3& qin) {
    H3 q = qin;
    double a01 = q.o01.x * q.o01.x + q.o01.y * q.o01.y;
    double a02 = q.o02.x * q.o02.x + q.o02.y * q.o02.y;
    double a12 = q.o12.x * q.o12.x + q.o12.y * q.o12.y;
    double c1 = 0.5 * (q.d0 * q.d0 + q.d1 * q.d1 + q.d2 * q.d2) + a01 + a02 + a12;
    int s = 0;
    if (c1 > 0.75) {
        double sc = 1.0;
        while (c1 * sc * sc > 0.75 && s < 60) { sc *= 0.5; s++; }
        q.d0 *= sc; q.d1 *= sc; q.d2 *= sc;
        q.o01.x *= sc; q.o01.y *= sc; q.o02.x *= sc; q.o02.y *= sc; q.o12.x *= sc; q.o12.y *= sc;
        a01 *= sc * sc; a02 *= sc * sc; a12 *= sc * sc;
        c1 *= sc * sc;
    }
    // t = q01*q12*conj(q02)
    double2 t = cmul(q.o01, q.o12);
    double re3 = t.x * q.o02.x + t.y * q.o02.y;
    double c0 = q.d0 * q.d1 * q.d2 + 2.0 * re3 - q.d0 * a12 - q.d1 * a02 - q.d2 * a01;
    double2 f0, f1, f2;
    ch_coefficients(c0, c1, f0, f1, f2);
    // Q^2 (Hermitian)
    double s00 = q.d0 * q.d0 + a01 + a02;
    double s11 = a01 + q.d1 * q.d1 + a12;
    double s22 = a02 + a12 + q.d2 * q.d2;
    double2 s01, s02, s12;
    // Q2_01 = (d0+d1) q01 + q02 conj(q12)
    s01 = make_double2((q.d0 + q.d1) * q.o01.x, (q.d0 + q.d1) * q.o01.y);
    cmac_c(s01, q.o02, q.o12);
    // Q2_02 = (d0+d2) q02 + q01 q12
    s02 = make_double2((q.d0 + q.d2) * q.o02.x, (q.d0 + q.d2) * q.o02.y);
    cmac(s02, q.o01, q.o12);
    // Q2_12 = (d1+d2) q12 + conj(q01) q02
    s12 = make_double2((q.d1 + q.d2) * q.o12.x, (q.d1 + q.d2) * q.o12.y);
    cmac_ca(s12, q.o01, q.o02);
    M3 e;
    // diagonal: f0 + f1*d + f2*s
    e.e[0] = make_double2(f0.x + f1.x * q.d0 + f2.x * s00, f0.y + f1.y * q.d0 + f2.y * s00);
    e.e[4] = make_double2(f0.x + f1.x * q.d1 + f2.x * s11, f0.y + f1.y * q.d1 + f2.y * s11);
    e.e[8] = make_double2(f0.x + f1.x * q.d2 + f2.x * s22, f0.y + f1.y * q.d2 + f2.y * s22);
    // upper
    e.e[1] = cmul(f1, q.o01); cmac(e.e[1], f2, s01);
    e.e[2] = cmul(f1, q.o02); cmac(e.e[2], f2, s02);
    e.e[5] = cmul(f1, q.o12); cmac(e.e[5], f2, s12);
    // lower: f1*conj(q) + f2*conj(s)
    e.e[3] = make_double2(0.0, 0.0); cmac_c(e.e[3], f1, q.o01); cmac_c(e.e[3], f2, s01);
    e.e[6] = make_double2(0.0, 0.0); cmac_c(e.e[6], f1, q.o02); cmac_c(e.e[6], f2, s02);
    e.e[7] = make_double2(0.0, 0.0); cmac_c(e.e[7], f1, q.o12); cmac_c(e.e[7], f2, s12);
    for (int k = 0; k < s; k++) e = mul_nn(e, e);
    return e;
}

// exp(t * sum_a c_a i lambda_a/2)
__device__ __forceinline__ M3 exp_ta(const double* c, double t) { return exp_iH(h3_from_coeffs(c, t)); }

// exp(t * sum_a c_a i lambda_a/2) * U for a unitary U: the product is unitary, so rows 0,1 are computed (the third row of the
// exponential is then dead code) and row 2 = conj(row0 x row1): 96 instead of 108 FP64 instructions, ~24 less in the exponential
__device__ __forceinline__ M3 exp_ta_times_su3(const double* c, double t, const M3& u) {
    const M3 e = exp_ta(c, t);
    return complete_su3(r2_mul_nn(rows01(e), u));
}

// SU(3) reunitarisation (src/4D/nowing/gaugefields_4D_nowing.jl:2387-2458, :195-238)
__device__ __forceinline__ M3 reunitarize(const M3& m) {
    M3 u = m;
    double2 w1 = make_double2(0.0, 0.0);
    double w2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cmac_c(w1, u.e[3 + c], u.e[c]);
        w2 += u.e[c].x * u.e[c].x + u.e[c].y * u.e[c].y;
    }
    w1.x = -w1.x / w2; w1.y = -w1.y / w2;
    double w3 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double2 x = u.e[3 + c];
        cmac(x, w1, u.e[c]);
        u.e[3 + c] = x;
        w3 += x.x * x.x + x.y * x.y;
    }
    double s2 = 1.0 / sqrt(w2), s3 = 1.0 / sqrt(w3);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        u.e[c].x *= s2; u.e[c].y *= s2;
        u.e[3 + c].x *= s3; u.e[3 + c].y *= s3;
    }
    // row 3 = conj(row1 x row2)
    double2 t;
    t = cmul(u.e[1], u.e[5]); { double2 v = cmul(u.e[2], u.e[4]); t.x -= v.x; t.y -= v.y; } u.e[6] = make_double2(t.x, -t.y);
    t = cmul(u.e[2], u.e[3]); { double2 v = cmul(u.e[0], u.e[5]); t.x -= v.x; t.y -= v.y; } u.e[7] = make_double2(t.x, -t.y);
    t = cmul(u.e[0], u.e[4]); { double2 v = cmul(u.e[1], u.e[3]); t.x -= v.x; t.y -= v.y; } u.e[8] = make_double2(t.x, -t.y);
    return u;
}

// Philox4x32-10 (Salmon et al., SC11), checked against the Random123 known-answer vectors
__device__ __host__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
        unsigned n1 = (unsigned)p1;
        unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        unsigned n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void site_uniform_pair(unsigned k0, unsigned k1, unsigned long long gsite, unsigned n, double& u0, double& u1) {
    unsigned o[4];
    philox4x32_10((unsigned)gsite, (unsigned)(gsite >> 32), n, 0u, k0, k1, o);
    unsigned long long a = ((unsigned long long)o[1] << 32) | o[0], b = ((unsigned long long)o[3] << 32) | o[2];
    u0 = (double)(a >> 11) * 0x1.0p-53;
    u1 = (double)(b >> 11) * 0x1.0p-53;
}

}  // namespace gfb
