// stout.cu -- backward pass (chain rule) of one plaquette-staple stout layer with scalar rho.
//
// Reference: layer_pullback! -> backward_dSdUαUβρ_add! (src/smearing/stout_fast.jl:222-245, 317-407) with the
// pieces calc_dSdu1! (:629), calc_dSdQ!/CdexpQdQ! (:636, 888-946, 1031-1081), calc_dSdΩ! (:683), calc_dSdC!
// (:688), calc_dSdUdag! (:693) and calc_dSdUν_fromdSCμ_add! (:712-785), whose symbolic dCμ/dUν and dCμ†/dUν
// tables (src/smearing/stout_dataset.jl:22-95) are expanded here by hand for the plaquette staple.
//
// Convention (src/molecular_dynamics.jl:255-265): dS = sum_x,mu tr(D_mu(x) dU_mu(x)) + c.c.; D is "dSdU".
// Layer: C_mu = rho V_mu, Omega = C U^dag, Q = TA(Omega), U' = exp(Q) U.  Given D' = dS/dU':
//   kernel 1 (k_stout_local, one thread per link, needs the staple stencil of U):
//       D_mu    = D' exp(Q) + (dS/dOmega C)^dag
//       Lambda  = dS/dC = U^dag dS/dOmega,  dS/dOmega = TA(L),  tr(L dQ) = tr(U D' d exp(Q))
//   kernel 2 (k_stout_gather, one thread per link, stencil over U and Lambda):
//       D_alpha(y) += rho * sum_{beta != alpha} [ six terms, see below ]
// Two kernels because Lambda of the neighbours must be complete before it is gathered.
// HBM bytes per site (fp64): kernel 1 reads U 576 + D' 576, writes D 576 + Lambda 576; kernel 2 reads U 576 +
// Lambda 576 + D 576, writes D 576 => 4608 B/site per layer.
#include "gfb_internal.h"
#include "stencil.cuh"

namespace gfb {

struct C3 {
    double2 v0, v1, v2;
};

// Pull-back of the exponential, L with tr(L dQ) = tr(C d exp(Q)) for Q = i*H, H Hermitian traceless given by its
// 8 coefficients (the tape of the forward pass).  With exp(iH) = f0 + f1 H + f2 H^2 (Cayley-Hamilton, c0 = det H,
// c1 = tr H^2 / 2) one has (Morningstar-Peardon; the reference's closed form, src/smearing/stout_fast.jl:1046-1081)
//   i L = tr(C B1) H + tr(C B2) H^2 + f1 C + f2 (H C + C H),   B1 = sum_j (df_j/dc1) H^j,  B2 = sum_j (df_j/dc0) H^j.
// Instead of the reference's trigonometric b_ij (singular for w -> 0 and 9u^2 = w^2, src/AbstractGaugefields.jl:3302-3343)
// f_j and both derivative sets come from forward-mode differentiation of the Horner recursion of the Taylor polynomial
// reduced with H^3 = c1 H + c0 -- a polynomial in (c0, c1), regular everywhere (pb_terms below).
// one Horner term of f(H) = sum_n a_n H^n (a_n = i^n / n!) and of its derivatives, reduced with H^3 = c1 H + c0:
//   p <- a_n + H p :  (p0, p1, p2) <- (a_n + c0 p2, p0 + c1 p2, p1)
//   d = dp/dc0     :  (d0, d1, d2) <- (p2 + c0 d2, d0 + c1 d2, d1)
//   e = dp/dc1     :  (e0, e1, e2) <- (c0 e2, e0 + p2 + c1 e2, e1)
// c0, c1 are real, so real and imaginary parts never mix: 14 FP64 instructions per term (the first version iterated
// P <- 1 + (iH/n) P with complex scalings from a constant table: 33 per term plus the loop; same truncated series).
template <int n>
__device__ __forceinline__ void pb_terms(double c0, double c1, C3& p, C3& d, C3& e) {
    constexpr double a = ((n & 2) ? -1.0 : 1.0) * InvFactorial<n>::v;
    C3 np, nd, ne;
    if ((n & 1) == 0) { np.v0.x = fma(c0, p.v2.x, a); np.v0.y = c0 * p.v2.y; }
    else { np.v0.x = c0 * p.v2.x; np.v0.y = fma(c0, p.v2.y, a); }
    np.v1 = make_double2(fma(c1, p.v2.x, p.v0.x), fma(c1, p.v2.y, p.v0.y));
    np.v2 = p.v1;
    nd.v0 = make_double2(fma(c0, d.v2.x, p.v2.x), fma(c0, d.v2.y, p.v2.y));
    nd.v1 = make_double2(fma(c1, d.v2.x, d.v0.x), fma(c1, d.v2.y, d.v0.y));
    nd.v2 = d.v1;
    ne.v0 = make_double2(c0 * e.v2.x, c0 * e.v2.y);
    ne.v1 = make_double2(fma(c1, e.v2.x, e.v0.x + p.v2.x), fma(c1, e.v2.y, e.v0.y + p.v2.y));
    ne.v2 = e.v1;
    p = np; d = nd; e = ne;
    if constexpr (n > 0) pb_terms<n - 1>(c0, c1, p, d, e);
}
template <int N>
__device__ __forceinline__ void pb_series(double c0, double c1, C3& p, C3& d, C3& e) {
    const double2 z = make_double2(0.0, 0.0);
    p = {z, z, z}; d = {z, z, z}; e = {z, z, z};
    pb_terms<N>(c0, c1, p, d, e);
}

__device__ __forceinline__ M3 exp_pullback(const M3& cm, const double* q) {
    const H3 h = h3_from_coeffs(q, 1.0);
    const double a01 = h.o01.x * h.o01.x + h.o01.y * h.o01.y;
    const double a02 = h.o02.x * h.o02.x + h.o02.y * h.o02.y;
    const double a12 = h.o12.x * h.o12.x + h.o12.y * h.o12.y;
    const double c1 = 0.5 * (h.d0 * h.d0 + h.d1 * h.d1 + h.d2 * h.d2) + a01 + a02 + a12;
    const double2 t3 = cmul(h.o01, h.o12);
    const double c0 = h.d0 * h.d1 * h.d2 + 2.0 * (t3.x * h.o02.x + t3.y * h.o02.y) - h.d0 * a12 - h.d1 * a02 - h.d2 * a01;
    // |eigenvalues| x <= sqrt(4 c1 / 3); the derivative series lose one power: remainder ~ x^N / N!.
    //   c1 <= 3/64 (x <= 1/4): N = 14 -> 4e-20;  c1 <= 3/4 (x <= 1): N = 21 -> 2e-20;  c1 <= 12 (x <= 4): N = 39 -> 1e-23 * 4^39 = 1.5e-23..
    C3 p, d0, d1;  // f_j, df_j/dc0, df_j/dc1
    if (c1 <= 0.046875) pb_series<14>(c0, c1, p, d0, d1);
    else if (c1 <= 0.75) pb_series<21>(c0, c1, p, d0, d1);
    else pb_series<39>(c0, c1, p, d0, d1);
    // H and H^2 as full matrices
    M3 hm, h2;
    hm.e[0] = make_double2(h.d0, 0.0); hm.e[4] = make_double2(h.d1, 0.0); hm.e[8] = make_double2(h.d2, 0.0);
    hm.e[1] = h.o01; hm.e[2] = h.o02; hm.e[5] = h.o12;
    hm.e[3] = make_double2(h.o01.x, -h.o01.y); hm.e[6] = make_double2(h.o02.x, -h.o02.y); hm.e[7] = make_double2(h.o12.x, -h.o12.y);
    h2 = mul_nn(hm, hm);
    const double2 trc = make_double2(cm.e[0].x + cm.e[4].x + cm.e[8].x, cm.e[0].y + cm.e[4].y + cm.e[8].y);
    const double2 trch = tr_nn(cm, hm), trch2 = tr_nn(cm, h2);
    double2 t1 = cmul(d1.v0, trc); cmac(t1, d1.v1, trch); cmac(t1, d1.v2, trch2);  // tr(C B1)
    double2 t2 = cmul(d0.v0, trc); cmac(t2, d0.v1, trch); cmac(t2, d0.v2, trch2);  // tr(C B2)
    M3 hc = mul_nn(hm, cm);
    mac_nn(hc, cm, hm);
    M3 r;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        double2 v = cmul(t1, hm.e[k]);
        cmac(v, t2, h2.e[k]);
        cmac(v, p.v1, cm.e[k]);
        cmac(v, p.v2, hc.e[k]);
        r.e[k] = make_double2(v.y, -v.x);  // divide by i
    }
    return r;
}

#ifndef GFB_STOUT_LOCAL_MINB
#define GFB_STOUT_LOCAL_MINB 3  // 168 registers, 308 bytes of spills in L1: back_prop 61.1 -> 60.0 ms at 48^3x96 (the gather is slower at 3)
#endif
#ifndef GFB_STOUT_GATHER_MINB
#define GFB_STOUT_GATHER_MINB 2
#endif
// kernel 1: site-local part of the pull-back and Lambda = dS/dC
template <bool FULL3>
__global__ void __launch_bounds__(128, GFB_STOUT_LOCAL_MINB)
k_stout_local(Geom g, const double2* __restrict__ u, const double2* __restrict__ dout, double2* __restrict__ lambda, double2* __restrict__ din, double rho) {
    const int mu = threadIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord x = decode_site(g, n, 0, g.tloc);
    M3 c = staple_sum<FULL3>(u, g, x, mu);  // V_mu
    const M3 umu = load_link(u, g, x, mu);
    double q[8];
    {
        M3 w = mul_nd(umu, c);  // U V^dag ; Q = TA(rho V U^dag) = -rho TA(U V^dag), exactly as the forward kernel forms it
        ta_coeffs(w, q);
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] *= -rho;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) { c.e[k].x *= rho; c.e[k].y *= rho; }  // C_mu
    const M3 dp = load_link(dout, g, x, mu);
    M3 acc;
    {
        const M3 eq = exp_ta(q, 1.0);
        acc = mul_nn(dp, eq);  // calc_dSdu1!
    }
    M3 dsdo;
    {
        const M3 cc = mul_nn(umu, dp);        // calc_dSdQ!: C = U dS/dU'
        dsdo = ta_matrix(exp_pullback(cc, q));  // calc_dSdΩ!
    }
    store_link(lambda, g, x, mu, mul_dn(umu, dsdo));  // calc_dSdC!: U^dag dS/dOmega
    {
        // calc_dSdUdag! then add_U!(dSdU, dSdUdag'): (dS/dOmega C)^dag = C^dag dS/dOmega^dag
        const M3 t = mul_nn(dsdo, c);
        m3_add(acc, m3_dagger(t));
    }
    store_link(din, g, x, mu, acc);
}

// kernel 2: D_alpha(y) += rho * sum_{beta != alpha} of the six terms in which U_alpha(y) enters C_beta / C_alpha
// (a) U_b(y+a) U_a(y+b)^dag L_b(y)            U_alpha(x) in the upper staple of C_beta(x), x = y
// (c) U_b(y+a) L_a(y+b) U_b(y)^dag            U_alpha(x-beta) in the lower staple of C_alpha(x), x = y+beta
// (f) L_b(y+a)^dag U_a(y+b)^dag U_b(y)^dag    U_alpha(x-alpha) in the adjoint lower staple of C_beta(x)^dag, x = y+alpha
// (b) U_b(y-b+a)^dag L_a(y-b) U_b(y-b)        U_alpha(x+beta) in the upper staple of C_alpha(x), x = y-beta
// (e) U_b(y-b+a)^dag U_a(y-b)^dag L_b(y-b)^dag  U_alpha(x+beta) in the adjoint upper staple of C_beta(x)^dag, x = y-beta
// (d) L_b(y+a-b) U_a(y-b)^dag U_b(y-b)        U_alpha(x-beta+alpha)... in the lower staple of C_beta(x), x = y+alpha-beta
// (a = alpha, b = beta, L = Lambda).  Pairs sharing a factor are combined: 10 products per beta.
__global__ void __launch_bounds__(128, GFB_STOUT_GATHER_MINB)
k_stout_gather(Geom g, const double2* __restrict__ u, const double2* __restrict__ lambda, double2* __restrict__ din, double rho) {
    const int al = threadIdx.y;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (long)g.v3 * g.tloc) return;
    const Coord y = decode_site(g, n, 0, g.tloc);
    const Coord ya = step(g, y, al, +1);
    M3 acc = m3_zero();
#pragma unroll 1
    for (int be = 0; be < 4; be++) {
        if (be == al) continue;
        const Coord yb = step(g, y, be, +1);
        const Coord ym = step(g, y, be, -1);
        const Coord yma = step(g, ym, al, +1);
        {
            // upper: (c) + (f) = [U_b(y+a) L_a(y+b) + L_b(y+a)^dag U_a(y+b)^dag] U_b(y)^dag ; (a) = U_b(y+a) U_a(y+b)^dag L_b(y)
            const M3 uba = load_link(u, g, ya, be);
            M3 t = mul_nn(uba, load_link(lambda, g, yb, al));
            const M3 uab = load_link(u, g, yb, al);
            {
                const M3 lba = load_link(lambda, g, ya, be);
                // t += L_b(y+a)^dag U_a(y+b)^dag = (U_a(y+b) L_b(y+a))^dag
                const M3 w = mul_nn(uab, lba);
                m3_add(t, m3_dagger(w));
            }
            mac_nd(acc, t, load_link(u, g, y, be));
            t = mul_nd(uba, uab);
            mac_nn(acc, t, load_link(lambda, g, y, be));
        }
        {
            // lower: (b) + (e) = U_b(y-b+a)^dag [L_a(y-b) U_b(y-b) + U_a(y-b)^dag L_b(y-b)^dag] ; (d) = L_b(y+a-b) U_a(y-b)^dag U_b(y-b)
            const M3 ubm = load_link(u, g, ym, be);
            M3 t = mul_nn(load_link(lambda, g, ym, al), ubm);
            const M3 uam = load_link(u, g, ym, al);
            {
                // t += U_a(y-b)^dag L_b(y-b)^dag = (L_b(y-b) U_a(y-b))^dag
                const M3 w = mul_nn(load_link(lambda, g, ym, be), uam);
                m3_add(t, m3_dagger(w));
            }
            mac_dn(acc, load_link(u, g, yma, be), t);
            t = mul_dn(uam, ubm);
            mac_nn(acc, load_link(lambda, g, yma, be), t);
        }
    }
    const unsigned off = link_offset(g, y, al);
    M3 d = m3_load_rw(din + off, (unsigned)g.v3);
#pragma unroll
    for (int k = 0; k < 9; k++) { d.e[k].x = fma(rho, acc.e[k].x, d.e[k].x); d.e[k].y = fma(rho, acc.e[k].y, d.e[k].y); }
    m3_store(din + off, (unsigned)g.v3, d);
}

void launch_stout_lambda(cudaStream_t st, const Geom& g, const double2* u, const double2* dout, double2* lambda, double2* din, double rho, bool full3) {
    dim3 block(32, 4);
    long nsites = (long)g.v3 * g.tloc;
    if (full3) k_stout_local<true><<<(unsigned)((nsites + 31) / 32), block, 0, st>>>(g, u, dout, lambda, din, rho);
    else k_stout_local<false><<<(unsigned)((nsites + 31) / 32), block, 0, st>>>(g, u, dout, lambda, din, rho);
}
void launch_stout_backward(cudaStream_t st, const Geom& g, const double2* u, const double2* lambda, double2* din, double rho) {
    dim3 block(32, 4);
    long nsites = (long)g.v3 * g.tloc;
    k_stout_gather<<<(unsigned)((nsites + 31) / 32), block, 0, st>>>(g, u, lambda, din, rho);
}

}  // namespace gfb
