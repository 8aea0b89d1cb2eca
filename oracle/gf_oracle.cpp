// gf_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17 restatement of the math on Gaugefields.jl's quenched SU(3)
// Wilson update path, used ONLY as the checker for the CUDA library
// (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
// legs).  Nothing under gaugefields.jl_b200/ may include, link or call this.
//
// Parity status: PINNED against the reference's own golden values
//   * legacy "Reproducible" hot start (StableRNG(123)) 4^4 SU(3) plaquette
//     0.008449494077606137              (test/init.jl:276-283)
//   * 4^4 SU(3) Wilson flow, 100 x eps=0.01, plaquette 0.8786515255315753
//                                        (test/gradientflow_test.jl:129-139)
//   * cold start invariants, MD reversibility < 2e-12, force additivity
//                                        (test/md_driver.jl:371-395, 417-482)
//   * SU(2)-embedded one-instanton plaquette 0.9796864531099871
//                                        (test/init.jl:351-371)
// UNPINNED (LatticeMatrices.jl is not vendored): the Philox key schedule of
// the LM site streams; the Philox-keyed hot start / Gaussian fill below follow
// the reference's *stream structure* (test/MPIJACCtest/random_fields_site_rng.jl:22-42)
// with a key schedule defined in DESIGN.md.
//
// Every function cites the reference file:line (relative to /root/reference)
// whose behaviour it restates.  Storage is the reference's host layout:
//   links   ComplexF64[3,3,NX,NY,NZ,NT] column-major per direction (src/API.jl:516-529)
//   momenta Float64[8,1,NX,NY,NZ,NT]                        (TA_gaugefields_4D_MPILattice.jl:34-35)
// with the 4 directions stored back to back.

#include <omp.h>

#include <complex>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

typedef std::complex<double> cd;

namespace {

struct M3 {
    cd a[3][3];
};

inline M3 zero3() { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = 0.0; return r; }
inline M3 ident3() { M3 r = zero3(); for (int i = 0; i < 3; i++) r.a[i][i] = 1.0; return r; }
inline M3 mul(const M3& x, const M3& y) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cd s = 0.0;
            for (int k = 0; k < 3; k++) s += x.a[i][k] * y.a[k][j];
            r.a[i][j] = s;
        }
    return r;
}
inline M3 dag(const M3& x) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = std::conj(x.a[j][i]); return r; }
inline M3 add(const M3& x, const M3& y) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = x.a[i][j] + y.a[i][j]; return r; }
inline M3 sub(const M3& x, const M3& y) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = x.a[i][j] - y.a[i][j]; return r; }
inline M3 scale(cd s, const M3& x) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.a[i][j] = s * x.a[i][j]; return r; }
inline cd trace(const M3& x) { return x.a[0][0] + x.a[1][1] + x.a[2][2]; }
inline double norm1(const M3& x) { double m = 0; for (int j = 0; j < 3; j++) { double s = 0; for (int i = 0; i < 3; i++) s += std::abs(x.a[i][j]); m = std::max(m, s); } return m; }

struct Lat {
    int n[4];
    long V;
    explicit Lat(const int* d) { for (int i = 0; i < 4; i++) n[i] = d[i]; V = (long)n[0] * n[1] * n[2] * n[3]; }
    inline long idx(const int* x) const { return x[0] + (long)n[0] * (x[1] + (long)n[1] * (x[2] + (long)n[2] * x[3])); }
    inline void coord(long s, int* x) const { for (int i = 0; i < 4; i++) { x[i] = (int)(s % n[i]); s /= n[i]; } }
};

// link accessor in the reference host layout (element (i,j) at i + 3*j, src/API.jl:516-529)
inline M3 getU(const double* U, const Lat& L, int mu, long s) {
    const double* p = U + ((long)mu * L.V + s) * 18;
    M3 r;
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) r.a[i][j] = cd(p[2 * (i + 3 * j)], p[2 * (i + 3 * j) + 1]);
    return r;
}
inline void setU(double* U, const Lat& L, int mu, long s, const M3& m) {
    double* p = U + ((long)mu * L.V + s) * 18;
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) { p[2 * (i + 3 * j)] = m.a[i][j].real(); p[2 * (i + 3 * j) + 1] = m.a[i][j].imag(); }
}

// Ordered product of links along a path starting at site x.
// A step (mu,+1) multiplies by U_mu(y) and moves y -> y+mu; a step (mu,-1) moves
// y -> y-mu first and multiplies by U_mu(y)^dagger.  Periodic wrap.
// Restates evaluate_gaugelinks! (src/AbstractGaugefields.jl:1797-1862) with the
// shift semantics of src/4D/nowing/gaugefields_4D_nowing.jl:380-412.
struct Step { int mu; int sgn; };
inline M3 path_product(const double* U, const Lat& L, const int* x0, const Step* st, int nst) {
    int y[4] = {x0[0], x0[1], x0[2], x0[3]};
    M3 r = ident3();
    for (int k = 0; k < nst; k++) {
        int mu = st[k].mu;
        if (st[k].sgn > 0) {
            r = mul(r, getU(U, L, mu, L.idx(y)));
            y[mu] = (y[mu] + 1) % L.n[mu];
        } else {
            y[mu] = (y[mu] + L.n[mu] - 1) % L.n[mu];
            r = mul(r, dag(getU(U, L, mu, L.idx(y))));
        }
    }
    return r;
}

const double SR3 = std::sqrt(3.0);

// Traceless anti-Hermitian projection to 8 Gell-Mann coefficients,
// TA(M) = sum_a c_a * i*lambda_a/2.  Restates
// src/4D/TA_gaugefields_4D_serial.jl:181-269.
inline void ta_coeffs(const M3& m, double* c) {
    M3 y;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) y.a[i][j] = 0.5 * (m.a[i][j] - std::conj(m.a[j][i]));
    cd tr = trace(y) / 3.0;
    for (int i = 0; i < 3; i++) y.a[i][i] -= tr;
    c[0] = y.a[0][1].imag() + y.a[1][0].imag();
    c[1] = y.a[0][1].real() - y.a[1][0].real();
    c[2] = y.a[0][0].imag() - y.a[1][1].imag();
    c[3] = y.a[0][2].imag() + y.a[2][0].imag();
    c[4] = y.a[0][2].real() - y.a[2][0].real();
    c[5] = y.a[1][2].imag() + y.a[2][1].imag();
    c[6] = y.a[1][2].real() - y.a[2][1].real();
    c[7] = (y.a[0][0].imag() + y.a[1][1].imag() - 2.0 * y.a[2][2].imag()) / SR3;
}

// matrix-valued projection Q = (M-M^dag)/2 - tr/3 (src/4D/nowing/gaugefields_4D_nowing.jl:1253-1345)
inline M3 ta_matrix(const M3& m) {
    M3 y;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) y.a[i][j] = 0.5 * (m.a[i][j] - std::conj(m.a[j][i]));
    cd tr = trace(y) / 3.0;
    for (int i = 0; i < 3; i++) y.a[i][i] -= tr;
    return y;
}

// Hermitian matrix H = sum_a c_a lambda_a / 2 (src/4D/TA_gaugefields_4D_serial.jl:779-847)
inline M3 hermitian_from_coeffs(const double* u, double t) {
    double c[8];
    for (int a = 0; a < 8; a++) c[a] = t * u[a] * 0.5;
    M3 h;
    h.a[0][0] = c[2] + c[7] / SR3;
    h.a[0][1] = cd(c[0], -c[1]);
    h.a[0][2] = cd(c[3], -c[4]);
    h.a[1][0] = cd(c[0], c[1]);
    h.a[1][1] = -c[2] + c[7] / SR3;
    h.a[1][2] = cd(c[5], -c[6]);
    h.a[2][0] = cd(c[3], c[4]);
    h.a[2][1] = cd(c[5], c[6]);
    h.a[2][2] = -2.0 * c[7] / SR3;
    return h;
}

// exp of a general 3x3 complex matrix by scaling-and-squaring Taylor series
// (accurate to ~1e-16; independent of both the legacy eigen route and the
// Cayley-Hamilton route used on the GPU).
inline M3 exp_taylor(const M3& x) {
    double nrm = norm1(x);
    int s = 0;
    while (nrm > 1.0) { nrm *= 0.5; s++; }  // few squarings: each one doubles the rounding error
    M3 y = scale(std::ldexp(1.0, -s), x);
    M3 r = ident3();
    M3 term = ident3();
    for (int n = 1; n <= 30; n++) {
        term = scale(1.0 / n, mul(term, y));
        r = add(r, term);
    }
    for (int k = 0; k < s; k++) r = mul(r, r);
    return r;
}

// exp(i H), H Hermitian traceless, via the closed-form eigenvalues (trigonometric
// cubic solution) and spectral projectors.  Follows the eigen-decomposition route
// of exptU! (src/4D/TA_gaugefields_4D_serial.jl:850-1074) without its tinyvalue /
// Taylor-4 artefacts: near-degenerate spectra fall back to exp_taylor.
inline M3 exp_iH_eigen(const M3& h, bool* used_fallback) {
    const double PI23 = 2.0 * M_PI / 3.0;
    M3 h2 = mul(h, h);
    double c1 = 0.5 * trace(h2).real();                 // = -p of the depressed cubic  e^3 - c1 e - c0 = 0
    cd det = h.a[0][0] * (h.a[1][1] * h.a[2][2] - h.a[1][2] * h.a[2][1]) - h.a[0][1] * (h.a[1][0] * h.a[2][2] - h.a[1][2] * h.a[2][0]) +
             h.a[0][2] * (h.a[1][0] * h.a[2][1] - h.a[1][1] * h.a[2][0]);
    double c0 = det.real();
    *used_fallback = false;
    if (c1 < 1e-6) { *used_fallback = true; return exp_taylor(scale(cd(0, 1), h)); }
    double r = 2.0 * std::sqrt(c1 / 3.0);
    double arg = 3.0 * c0 / (c1 * r);
    arg = std::min(1.0, std::max(-1.0, arg));
    double th = std::acos(arg) / 3.0;
    double e[3] = {r * std::cos(th), r * std::cos(th + PI23), 0.0};
    e[2] = -e[0] - e[1];
    double gap = std::min({std::abs(e[0] - e[1]), std::abs(e[1] - e[2]), std::abs(e[0] - e[2])});
    if (gap < 1e-3 * r) { *used_fallback = true; return exp_taylor(scale(cd(0, 1), h)); }
    // spectral projectors P_k = prod_{l != k} (H - e_l)/(e_k - e_l)
    M3 out = zero3();
    for (int k = 0; k < 3; k++) {
        int l1 = (k + 1) % 3, l2 = (k + 2) % 3;
        M3 a = h, b = h;
        for (int i = 0; i < 3; i++) { a.a[i][i] -= e[l1]; b.a[i][i] -= e[l2]; }
        M3 p = scale(1.0 / ((e[k] - e[l1]) * (e[k] - e[l2])), mul(a, b));
        out = add(out, scale(cd(std::cos(e[k]), std::sin(e[k])), p));
    }
    return out;
}

// exp(t * sum_a u_a i lambda_a/2): exptU! for TA fields (src/4D/TA_gaugefields_4D_serial.jl:760-848)
inline M3 exp_ta(const double* u, double t, int route) {
    M3 h = hermitian_from_coeffs(u, t);
    if (route == 1) { bool fb; return exp_iH_eigen(h, &fb); }
    return exp_taylor(scale(cd(0, 1), h));
}

// ---------------------------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------------------------

// StableRNGs.jl LehmerRNG (un-vendored dependency, StableRNGs 1.x, Project.toml:48): 128-bit
// multiplicative congruential generator, state = (seed<<1)|1, output = high 64 bits;
// rand(Float64) = reinterpret(0x3ff0...|low52(u64)) - 1.  Pinned by the golden plaquette
// 0.008449494077606137 (test/init.jl:281), which only this variant reproduces.
struct Lehmer {
    unsigned __int128 state;
    explicit Lehmer(uint64_t seed) { state = (((unsigned __int128)seed) << 1) | 1; }
    inline uint64_t u64() {
        const unsigned __int128 mult = (((unsigned __int128)0x45a31efc5a35d971ULL) << 64) | 0x261fd0407a968addULL;
        state *= mult;
        return (uint64_t)(state >> 64);
    }
    inline double f64() {
        uint64_t b = 0x3ff0000000000000ULL | (u64() & 0x000fffffffffffffULL);
        double d; std::memcpy(&d, &b, 8);
        return d - 1.0;
    }
};

// Philox4x32-10 (Salmon et al. 2011).  Used for the decomposition-independent per-global-site
// streams (structure: test/MPIJACCtest/random_fields_site_rng.jl:22-42; key schedule: DESIGN.md).
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// stream key for (seed, sweep, direction, tag): first two words of Philox((seed,sweep),(tag,direction))
inline void stream_key(uint64_t seed, uint64_t sweep, uint32_t direction, uint32_t tag, uint32_t key[2]) {
    uint32_t ctr[4] = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)sweep, (uint32_t)(sweep >> 32)};
    uint32_t k[2] = {tag, direction};
    uint32_t o[4];
    philox4x32_10(ctr, k, o);
    key[0] = o[0]; key[1] = o[1];
}
// two uniforms in [0,1) from draw `n` of the stream of global site `gsite`
inline void site_uniform_pair(const uint32_t key[2], uint64_t gsite, uint32_t n, double* u0, double* u1) {
    uint32_t ctr[4] = {(uint32_t)gsite, (uint32_t)(gsite >> 32), n, 0u};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    uint64_t a = ((uint64_t)o[1] << 32) | o[0], b = ((uint64_t)o[3] << 32) | o[2];
    *u0 = (double)(a >> 11) * 0x1.0p-53;
    *u1 = (double)(b >> 11) * 0x1.0p-53;
}

const uint32_t TAG_HOT = 0x00484f54u;    // src/AbstractGaugefields.jl:121-123
const uint32_t TAG_GAUSS = 0x47415553u;

// SU(3) reunitarisation: row-1 normalise, row-2 Gram-Schmidt, row 3 = conj(row1 x row2)
// (src/4D/nowing/gaugefields_4D_nowing.jl:2387-2458 and :195-238)
inline M3 reunitarize(const M3& m) {
    M3 u = m;
    cd w1 = 0.0, w2 = 0.0;
    for (int c = 0; c < 3; c++) { w1 += u.a[1][c] * std::conj(u.a[0][c]); w2 += u.a[0][c] * std::conj(u.a[0][c]); }
    w1 = -w1 / w2;
    cd x[3];
    for (int c = 0; c < 3; c++) x[c] = u.a[1][c] + w1 * u.a[0][c];
    cd w3 = 0.0;
    for (int c = 0; c < 3; c++) w3 += x[c] * std::conj(x[c]);
    cd s3 = 1.0 / std::sqrt(w3), s2 = 1.0 / std::sqrt(w2);
    for (int c = 0; c < 3; c++) { u.a[0][c] = u.a[0][c] * s2; u.a[1][c] = x[c] * s3; }
    u.a[2][0] = std::conj(u.a[0][1] * u.a[1][2] - u.a[0][2] * u.a[1][1]);
    u.a[2][1] = std::conj(u.a[0][2] * u.a[1][0] - u.a[0][0] * u.a[1][2]);
    u.a[2][2] = std::conj(u.a[0][0] * u.a[1][1] - u.a[0][1] * u.a[1][0]);
    return u;
}

// sum over the two plaquette-type closed loops through U_mu(x) in the (mu,nu) plane:
//   U_mu(x) * [upper staple]^dag-type product, i.e. the loops (mu,nu,-mu,-nu) and (mu,-nu,-mu,nu).
// Their sum over nu equals U_mu(x) * V_mu(x)^dagger with V_mu the 6-staple sum of
// src/autostaples/wilsonloops.jl:468-484 / construct_double_staple! (src/AbstractGaugefields.jl:2856-2871).
inline M3 U_times_Vdag(const double* U, const Lat& L, const int* x, int mu) {
    M3 acc = zero3();
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        Step up[4] = {{mu, 1}, {nu, 1}, {mu, -1}, {nu, -1}};
        Step dn[4] = {{mu, 1}, {nu, -1}, {mu, -1}, {nu, 1}};
        acc = add(acc, path_product(U, L, x, up, 4));
        acc = add(acc, path_product(U, L, x, dn, 4));
    }
    return acc;
}

// the staple sum V_mu(x) itself (needed by stout: C_mu = rho * V_mu)
inline M3 staple_sum(const double* U, const Lat& L, const int* x, int mu) {
    M3 acc = zero3();
    for (int nu = 0; nu < 4; nu++) {
        if (nu == mu) continue;
        Step up[3] = {{nu, 1}, {mu, 1}, {nu, -1}};
        Step dn[3] = {{nu, -1}, {mu, 1}, {nu, 1}};
        acc = add(acc, path_product(U, L, x, up, 3));
        acc = add(acc, path_product(U, L, x, dn, 3));
    }
    return acc;
}

}  // namespace

extern "C" {

// ----------------------------------------------------------------------------- initial fields

// threads the OpenMP loops below run on (bench.py reports it as cpu_baseline.cores); n > 0 sets it first
int orc_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}

void orc_set_cold(double* U, const int* dims) {
    Lat L(dims);
    M3 one = ident3();
    for (int mu = 0; mu < 4; mu++) for (long s = 0; s < L.V; s++) setU(U, L, mu, s, one);
}

// legacy "Reproducible" hot start: every direction re-seeds StableRNG(123)
// (src/4D/nowing/gaugefields_4D_nowing.jl:240-279; Appendix B of SURVEY.md)
void orc_hot_start_stable123(double* U, const int* dims) {
    Lat L(dims);
    for (int mu = 0; mu < 4; mu++) {
        Lehmer rng(123);
        for (long s = 0; s < L.V; s++) {  // it,iz,iy,ix loops with ix fastest == linear site order
            M3 m;
            for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) { double re = rng.f64() - 0.5; double im = rng.f64() - 0.5; m.a[i][j] = cd(re, im); }
            setU(U, L, mu, s, reunitarize(m));
        }
    }
}

// Philox-keyed hot start: stream (seed, 0, mu+1, tag HOT) per global site
// (structure of src/4D/mpi_jacc/gaugefields_4D_MPILattice.jl:430-472)
void orc_hot_start_philox(double* U, const int* dims, uint64_t seed) {
    Lat L(dims);
    for (int mu = 0; mu < 4; mu++) {
        uint32_t key[2];
        stream_key(seed, 0, (uint32_t)(mu + 1), TAG_HOT, key);
#pragma omp parallel for
        for (long s = 0; s < L.V; s++) {
            M3 m;
            for (int k = 0; k < 9; k++) {
                double u0, u1;
                site_uniform_pair(key, (uint64_t)s, (uint32_t)k, &u0, &u1);
                m.a[k % 3][k / 3] = cd(u0 - 0.5, u1 - 0.5);
            }
            setU(U, L, mu, s, reunitarize(m));
        }
    }
}

// Gaussian momenta: per global site one stream keyed (seed, sweep, mu+1, tag GAUSS); the 8
// coefficients consume 4 Box-Muller (value, spare) pairs (TA_gaugefields_4D_MPILattice.jl:157-193,
// test/MPIJACCtest/random_fields_site_rng.jl:22-42)
void orc_gaussian_momenta(double* P, const int* dims, uint64_t seed, uint64_t sweep, double sigma) {
    Lat L(dims);
    for (int mu = 0; mu < 4; mu++) {
        uint32_t key[2];
        stream_key(seed, sweep, (uint32_t)(mu + 1), TAG_GAUSS, key);
#pragma omp parallel for
        for (long s = 0; s < L.V; s++) {
            double* p = P + ((long)mu * L.V + s) * 8;
            for (int k = 0; k < 4; k++) {
                double u0, u1;
                site_uniform_pair(key, (uint64_t)s, (uint32_t)k, &u0, &u1);
                double r = std::sqrt(-2.0 * std::log(1.0 - u0));
                double th = 2.0 * M_PI * u1;
                p[2 * k] = sigma * r * std::cos(th);
                p[2 * k + 1] = sigma * r * std::sin(th);
            }
        }
    }
}

void orc_reunitarize(double* U, const int* dims) {
    Lat L(dims);
    for (int mu = 0; mu < 4; mu++) for (long s = 0; s < L.V; s++) setU(U, L, mu, s, reunitarize(getU(U, L, mu, s)));
}

// ----------------------------------------------------------------------------- observables

// calculate_Plaquette (src/AbstractGaugefields.jl:2684-2699): sum_{x,mu<nu} Re tr P_munu, un-normalised
double orc_plaquette_sum(const double* U, const int* dims) {
    Lat L(dims);
    double tot = 0.0;
#pragma omp parallel for reduction(+ : tot)
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        double acc = 0.0;
        for (int mu = 0; mu < 4; mu++)
            for (int nu = mu + 1; nu < 4; nu++) {
                Step pl[4] = {{mu, 1}, {nu, 1}, {mu, -1}, {nu, -1}};
                acc += trace(path_product(U, L, x, pl, 4)).real();
            }
        tot += acc;
    }
    return tot;
}

// p*p (src/4D/TA_gaugefields_4D_serial.jl:106-129 summed over directions, src/TA_Gaugefields.jl:127-137)
double orc_momentum_norm2(const double* P, const int* dims) {
    Lat L(dims);
    double tot = 0.0;
    long n = 4 * L.V * 8;
#pragma omp parallel for reduction(+ : tot)
    for (long i = 0; i < n; i++) tot += P[i] * P[i];
    return tot;
}

// md_hamiltonian (src/molecular_dynamics.jl:494-505) for the Wilson action pushed as beta/2 * (plaq + plaq')
double orc_hamiltonian(const double* U, const double* P, const int* dims, double beta) {
    return -(beta / 3.0) * orc_plaquette_sum(U, dims) + 0.5 * orc_momentum_norm2(P, dims);
}

// Polyakov loop in the t direction (src/AbstractGaugefields.jl:2929-2956): sum over spatial sites of tr prod_t U_4 / (NX NY NZ)
void orc_polyakov(const double* U, const int* dims, double* out2) {
    Lat L(dims);
    cd tot = 0.0;
    for (int z = 0; z < L.n[2]; z++) for (int y = 0; y < L.n[1]; y++) for (int x = 0; x < L.n[0]; x++) {
        M3 r = ident3();
        for (int t = 0; t < L.n[3]; t++) { int c[4] = {x, y, z, t}; r = mul(r, getU(U, L, 3, L.idx(c))); }
        tot += trace(r);
    }
    tot /= (double)((long)L.n[0] * L.n[1] * L.n[2]);
    out2[0] = tot.real(); out2[1] = tot.imag();
}

// clover energy density (samples/measurements/energydensity.jl:4-78)
double orc_energy_density_clover(const double* U, const int* dims) {
    Lat L(dims);
    double tot = 0.0;
#pragma omp parallel for reduction(+ : tot)
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        double acc = 0.0;
        for (int mu = 0; mu < 4; mu++)
            for (int nu = 0; nu < 4; nu++) {
                if (mu == nu) continue;
                Step l1[4] = {{mu, 1}, {nu, 1}, {mu, -1}, {nu, -1}};
                Step l2[4] = {{nu, 1}, {mu, -1}, {nu, -1}, {mu, 1}};
                Step l3[4] = {{nu, -1}, {mu, 1}, {nu, 1}, {mu, -1}};
                Step l4[4] = {{mu, -1}, {nu, -1}, {mu, 1}, {nu, 1}};
                M3 w = add(add(path_product(U, L, x, l1, 4), path_product(U, L, x, l2, 4)), add(path_product(U, L, x, l3, 4), path_product(U, L, x, l4, 4)));
                M3 g = ta_matrix(w);
                acc += (-trace(mul(g, g)) / 2.0).real();
            }
        tot += acc;
    }
    return tot / ((double)L.V * 16.0);
}

// ----------------------------------------------------------------------------- force, updates

// md_force! with the plaquette+plaquette' action at coefficient beta/2
// (src/molecular_dynamics.jl:251-267, src/action/GaugeActions.jl:95-123):
//   F_mu = -(1/3) * TAcoeffs( U_mu * (beta/2) * sum of 6 staples )
void orc_force(double* F, const double* U, const int* dims, double beta) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) {
            M3 w = scale(beta / 2.0, U_times_Vdag(U, L, x, mu));
            double c[8];
            ta_coeffs(w, c);
            double* f = F + ((long)mu * L.V + s) * 8;
            for (int a = 0; a < 8; a++) f[a] = (-1.0 / 3.0) * c[a];
        }
    }
}

// add_force!(F, U; plaqonly=true, factor=1) after clear (src/AbstractGaugefields.jl:2717-2762): F_mu = TAcoeffs(U_mu V_mu^dag)
void orc_flow_force(double* F, const double* U, const int* dims) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) ta_coeffs(U_times_Vdag(U, L, x, mu), F + ((long)mu * L.V + s) * 8);
    }
}

// update_gaugefields! / exp_aF_U! : U_mu <- exp(eps * P_mu) U_mu
// (src/molecular_dynamics.jl:513-531, src/AbstractGaugefields.jl:2810-2841); route 0 = Taylor, 1 = eigen
void orc_update_links_route(double* Uout, const double* Uin, const double* P, const int* dims, double eps, int route) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++)
        for (int mu = 0; mu < 4; mu++) {
            M3 e = exp_ta(P + ((long)mu * L.V + s) * 8, eps, route);
            setU(Uout, L, mu, s, mul(e, getU(Uin, L, mu, s)));
        }
}
void orc_update_links(double* U, const double* P, const int* dims, double eps) { orc_update_links_route(U, U, P, dims, eps, 0); }

// update_momenta!: P += eps * force(U) (src/molecular_dynamics.jl:539-551)
void orc_update_momenta(double* P, const double* U, const int* dims, double eps, double beta) {
    Lat L(dims);
    std::vector<double> F((size_t)4 * L.V * 8);
    orc_force(F.data(), U, dims, beta);
    long n = 4 * L.V * 8;
#pragma omp parallel for
    for (long i = 0; i < n; i++) P[i] += eps * F[i];
}

// md_step! QPQ / PQP (src/molecular_dynamics.jl:604-616); integrator 0 = QPQ, 1 = PQP
void orc_md_step(double* U, double* P, const int* dims, double beta, double eps, int integrator) {
    if (integrator == 0) {
        orc_update_links(U, P, dims, eps / 2);
        orc_update_momenta(P, U, dims, eps, beta);
        orc_update_links(U, P, dims, eps / 2);
    } else {
        orc_update_momenta(P, U, dims, eps / 2, beta);
        orc_update_links(U, P, dims, eps);
        orc_update_momenta(P, U, dims, eps / 2, beta);
    }
}

// md_trajectory! with diagnostics (src/molecular_dynamics.jl:712-730): H[0] initial, H[1] final
void orc_md_trajectory(double* U, double* P, const int* dims, double beta, int steps, double tau, int integrator, double* H) {
    H[0] = orc_hamiltonian(U, P, dims, beta);
    double eps = tau / steps;
    for (int k = 0; k < steps; k++) orc_md_step(U, P, dims, beta, eps, integrator);
    H[1] = orc_hamiltonian(U, P, dims, beta);
}

// one Luescher RK3 step of flow! (src/smearing/gradientflow.jl:171-238)
void orc_flow_step(double* U, const int* dims, double eps) {
    Lat L(dims);
    size_t nf = (size_t)4 * L.V * 8, nu = (size_t)4 * L.V * 18;
    std::vector<double> F0(nf), F1(nf), F2(nf), Ft(nf), W1(nu), W2(nu);
    orc_flow_force(F0.data(), U, dims);
    orc_update_links_route(W1.data(), U, F0.data(), dims, -eps / 4, 0);
    orc_flow_force(F1.data(), W1.data(), dims);
    for (size_t i = 0; i < nf; i++) Ft[i] = -(8 * eps / 9) * F1[i] + (17 * eps / 36) * F0[i];
    orc_update_links_route(W2.data(), W1.data(), Ft.data(), dims, 1.0, 0);
    orc_flow_force(F2.data(), W2.data(), dims);
    for (size_t i = 0; i < nf; i++) Ft[i] = -(3 * eps / 4) * F2[i] + (8 * eps / 9) * F1[i] - (17 * eps / 36) * F0[i];
    orc_update_links_route(U, W2.data(), Ft.data(), dims, 1.0, 0);
}


// ----------------------------------------------------------------------------- stout smearing

// STOUT_Layer forward! (src/smearing/stout_fast.jl:250-274, calc_C! :603-624) for the plaquette staple with scalar rho:
//   C_mu = rho * V_mu ;  Q_mu = TA(C_mu U_mu^dag) (matrix-valued, nowing:1253-1345) ;  U'_mu = exp(Q_mu) U_mu
// Qout (may be NULL) receives the 8 coefficients of Q_mu.
void orc_stout_forward(double* Uout, const double* Uin, const int* dims, double rho, double* Qout) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) {
            M3 u = getU(Uin, L, mu, s);
            M3 c = scale(rho, staple_sum(Uin, L, x, mu));
            M3 q = ta_matrix(mul(c, dag(u)));
            if (Qout) ta_coeffs(q, Qout + ((long)mu * L.V + s) * 8);
            setU(Uout, L, mu, s, mul(exp_taylor(q), u));
        }
    }
}

}  // extern "C"

// ----------------------------------------------------------------------------- stout backward

namespace {

// Pull-back of the matrix exponential, L = C * d exp(Q)/dQ, defined by tr(L dQ) = tr(C d(exp Q)).
// Route 0 (independent of the reference's closed form): term-by-term derivative of the Taylor
// series, d(Q^{n+1}) = sum_k Q^k dQ Q^{n-k}  =>  L = sum_n S_n/(n+1)!,  S_0 = C, S_{n+1} = Q S_n + C Q^{n+1}.
inline M3 exp_pullback_series(const M3& C, const M3& Q) {
    // scaling keeps the series short and well conditioned: exp(Q) = exp(Q/2^s)^(2^s) is NOT used here
    // (the derivative of a power is awkward); stout arguments are small, so 60 terms cover |Q| < 8.
    M3 L = C, S = C, CQn = C;
    double inv_fact = 1.0;  // 1/(n+1)!
    for (int n = 0; n < 60; n++) {
        CQn = mul(CQn, Q);               // C Q^{n+1}
        S = add(mul(Q, S), CQn);         // S_{n+1}
        inv_fact /= (double)(n + 2);     // 1/(n+2)!
        L = add(L, scale(inv_fact, S));
    }
    return L;
}

// Route 1: the reference's closed form (Morningstar-Peardon), CdexpQdQ! for NC=3
// (src/smearing/stout_fast.jl:888-946, 1031-1081) with calc_coefficients_Q
// (src/AbstractGaugefields.jl:3284-3343): with Qt = Q/i Hermitian,
//   L = [ tr(C B1) Qt + tr(C B2) Qt^2 + f1 C + f2 (Qt C + C Qt) ] / i,   B_i = b_i0 + b_i1 Qt + b_i2 Qt^2.
// Like the reference it returns false (output untouched) when |tr Q^2| <= 1e-18; unlike the reference it
// reflects c0 < 0 (the reference's formula is only valid for c0 >= 0) so it can be used on any input.
inline bool exp_pullback_closed(const M3& C, const M3& Q, M3* out) {
    cd trq2 = trace(mul(Q, Q));
    if (std::abs(trq2) <= 1e-18) return false;
    M3 qt = scale(cd(0, -1), Q);
    M3 qt2 = mul(qt, qt);
    cd det = qt.a[0][0] * (qt.a[1][1] * qt.a[2][2] - qt.a[1][2] * qt.a[2][1]) - qt.a[0][1] * (qt.a[1][0] * qt.a[2][2] - qt.a[1][2] * qt.a[2][0]) +
             qt.a[0][2] * (qt.a[1][0] * qt.a[2][1] - qt.a[1][1] * qt.a[2][0]);
    double c0 = det.real();
    double c1 = 0.5 * trace(qt2).real();
    bool reflect = c0 < 0;
    if (reflect) c0 = -c0;
    double c0max = 2.0 * std::pow(c1 / 3.0, 1.5);
    double th = std::acos(std::min(1.0, c0 / c0max));
    double u = std::sqrt(c1 / 3.0) * std::cos(th / 3.0);
    double w = std::sqrt(c1) * std::sin(th / 3.0);
    double w2 = w * w, u2 = u * u;
    double xi0, xi1;
    if (std::abs(w) < 0.05) {
        xi0 = 1.0 - w2 / 6.0 * (1.0 - w2 / 20.0 * (1.0 - w2 / 42.0 * (1.0 - w2 / 72.0)));
        xi1 = -(1.0 / 3.0 - w2 / 30.0 * (1.0 - w2 / 28.0 * (1.0 - w2 / 54.0)));
    } else {
        xi0 = std::sin(w) / w;
        xi1 = std::cos(w) / w2 - std::sin(w) / (w2 * w);
    }
    const cd I(0, 1);
    cd emiu = std::exp(-I * u), e2iu = std::exp(2.0 * I * u);
    double cw = std::cos(w);
    cd h0 = (u2 - w2) * e2iu + emiu * (8.0 * u2 * cw + 2.0 * I * u * (3.0 * u2 + w2) * xi0);
    cd h1 = 2.0 * u * e2iu - emiu * (2.0 * u * cw - I * (3.0 * u2 - w2) * xi0);
    cd h2 = e2iu - emiu * (cw + 3.0 * I * u * xi0);
    double denom = 9.0 * u2 - w2;
    cd f0 = h0 / denom, f1 = h1 / denom, f2 = h2 / denom;
    cd r10 = 2.0 * (u + I * (u2 - w2)) * e2iu + 2.0 * emiu * (4.0 * u * (2.0 - I * u) * cw + I * (9.0 * u2 + w2 - I * u * (3.0 * u2 + w2)) * xi0);
    cd r11 = 2.0 * (1.0 + 2.0 * I * u) * e2iu + emiu * (-2.0 * (1.0 - I * u) * cw + I * (6.0 * u + I * (w2 - 3.0 * u2)) * xi0);
    cd r12 = 2.0 * I * e2iu + I * emiu * (cw - 3.0 * (1.0 - I * u) * xi0);
    cd r20 = -2.0 * e2iu + 2.0 * I * u * emiu * (cw + (1.0 + 4.0 * I * u) * xi0 + 3.0 * u2 * xi1);
    cd r21 = -I * emiu * (cw + (1.0 + 2.0 * I * u) * xi0 - 3.0 * u2 * xi1);
    cd r22 = emiu * (xi0 - 3.0 * I * u * xi1);
    double d2 = 2.0 * denom * denom;
    cd b10 = (2.0 * u * r10 + (3.0 * u2 - w2) * r20 - 2.0 * (15.0 * u2 + w2) * f0) / d2;
    cd b11 = (2.0 * u * r11 + (3.0 * u2 - w2) * r21 - 2.0 * (15.0 * u2 + w2) * f1) / d2;
    cd b12 = (2.0 * u * r12 + (3.0 * u2 - w2) * r22 - 2.0 * (15.0 * u2 + w2) * f2) / d2;
    cd b20 = (r10 - 3.0 * u * r20 - 24.0 * u * f0) / d2;
    cd b21 = (r11 - 3.0 * u * r21 - 24.0 * u * f1) / d2;
    cd b22 = (r12 - 3.0 * u * r22 - 24.0 * u * f2) / d2;
    if (reflect) {  // f_j(-c0,c1) = (-1)^j conj f_j(c0,c1);  b_1j -> (-1)^j conj,  b_2j -> (-1)^(j+1) conj
        f0 = std::conj(f0); f1 = -std::conj(f1); f2 = std::conj(f2);
        b10 = std::conj(b10); b11 = -std::conj(b11); b12 = std::conj(b12);
        b20 = -std::conj(b20); b21 = std::conj(b21); b22 = -std::conj(b22);
    }
    M3 B1 = add(add(scale(b10, ident3()), scale(b11, qt)), scale(b12, qt2));
    M3 B2 = add(add(scale(b20, ident3()), scale(b21, qt)), scale(b22, qt2));
    cd t1 = trace(mul(C, B1)), t2 = trace(mul(C, B2));
    M3 r = add(add(scale(t1, qt), scale(t2, qt2)), add(scale(f1, C), scale(f2, add(mul(qt, C), mul(C, qt)))));
    *out = scale(cd(0, -1), r);
    return true;
}

}  // namespace

extern "C" {

// back-propagation through ONE plaquette-staple stout layer with scalar rho
// (layer_pullback! -> backward_dSdUαUβρ_add!, src/smearing/stout_fast.jl:222-245, 317-407; pieces
//  calc_dSdu1! :629, calc_dSdQ! :636, calc_dSdΩ! :683, calc_dSdC! :688, calc_dSdUdag! :693,
//  calc_dSdUν_fromdSCμ_add! :712-785 with the dCμ/dUν, dCμ†/dUν tables of src/smearing/stout_dataset.jl:22-95
//  expanded by hand for the plaquette staple).
// Convention (src/molecular_dynamics.jl:255-265): dS = sum tr(D_mu(x) dU_mu(x)) + c.c., D = "dSdU".
// dOut = dS/dU' (U' = smeared links), dIn = dS/dU.  route: 0 = series pull-back, 1 = closed form.
void orc_stout_backward(double* dIn, const double* dOut, const double* Uin, const int* dims, double rho, int route) {
    Lat L(dims);
    size_t nu = (size_t)4 * L.V * 18;
    std::vector<double> Lam(nu);  // Lambda_mu(x) = dS/dC_mu(x) = U_mu^dag dS/dOmega_mu
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) {
            M3 u = getU(Uin, L, mu, s), dp = getU(dOut, L, mu, s);
            M3 c = scale(rho, staple_sum(Uin, L, x, mu));
            M3 q = ta_matrix(mul(c, dag(u)));
            M3 eq = exp_taylor(q);
            M3 acc = mul(dp, eq);                        // calc_dSdu1!: dS/dU' * exp(Q)
            M3 cc = mul(u, dp);                          // calc_dSdQ!: C = U * dS/dU'
            M3 dsdq = zero3();                           // reference leaves the output untouched (zeroed temp) below eps_Q
            if (route == 1) exp_pullback_closed(cc, q, &dsdq);
            else dsdq = exp_pullback_series(cc, q);
            M3 dsdo = ta_matrix(dsdq);                   // calc_dSdΩ!
            setU(Lam.data(), L, mu, s, mul(dag(u), dsdo));  // calc_dSdC!
            acc = add(acc, dag(mul(dsdo, c)));           // calc_dSdUdag! then add_U!(dSdU, dSdUdag')
            setU(dIn, L, mu, s, acc);
        }
    }
    // dS/dC_mu star dC_mu/dU_nu and dS/dC_mu^dag star dC_mu^dag/dU_nu, gathered per target link (y, alpha)
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int y[4];
        L.coord(s, y);
        auto at = [&](int d1, int s1, int d2, int s2) {
            int z[4] = {y[0], y[1], y[2], y[3]};
            if (s1) z[d1] = (z[d1] + s1 + L.n[d1]) % L.n[d1];
            if (s2) z[d2] = (z[d2] + s2 + L.n[d2]) % L.n[d2];
            return L.idx(z);
        };
        for (int al = 0; al < 4; al++) {
            M3 acc = zero3();
            for (int be = 0; be < 4; be++) {
                if (be == al) continue;
                const int mu = be, nu = al;  // target link plays the role of U_nu in C_mu
                // (a) U_nu(x) in the upper staple of C_mu(x), x = y
                acc = add(acc, mul(mul(getU(Uin, L, mu, at(nu, 1, 0, 0)), dag(getU(Uin, L, nu, at(mu, 1, 0, 0)))), getU(Lam.data(), L, mu, s)));
                // (d) U_nu(x-nu+mu) in the lower staple of C_mu(x), x = y+nu-mu
                acc = add(acc, mul(getU(Lam.data(), L, mu, at(nu, 1, mu, -1)), mul(dag(getU(Uin, L, nu, at(mu, -1, 0, 0))), getU(Uin, L, mu, at(mu, -1, 0, 0)))));
                // (e) U_nu(x+mu) in the adjoint upper staple of C_mu(x)^dag, x = y-mu
                acc = add(acc, mul(mul(dag(getU(Uin, L, mu, at(mu, -1, nu, 1))), dag(getU(Uin, L, nu, at(mu, -1, 0, 0)))), dag(getU(Lam.data(), L, mu, at(mu, -1, 0, 0)))));
                // (f) U_nu(x-nu) in the adjoint lower staple of C_mu(x)^dag, x = y+nu
                acc = add(acc, mul(dag(getU(Lam.data(), L, mu, at(nu, 1, 0, 0))), mul(dag(getU(Uin, L, nu, at(mu, 1, 0, 0))), dag(getU(Uin, L, mu, s)))));
                // target link plays the role of U_mu in C_mu (mu = al), staple direction be
                const int m = al, n = be;
                // (b) U_mu(x+nu) in the upper staple of C_mu(x), x = y-n
                acc = add(acc, mul(mul(dag(getU(Uin, L, n, at(n, -1, m, 1))), getU(Lam.data(), L, m, at(n, -1, 0, 0))), getU(Uin, L, n, at(n, -1, 0, 0))));
                // (c) U_mu(x-nu) in the lower staple of C_mu(x), x = y+n
                acc = add(acc, mul(mul(getU(Uin, L, n, at(m, 1, 0, 0)), getU(Lam.data(), L, m, at(n, 1, 0, 0))), dag(getU(Uin, L, n, s))));
            }
            setU(dIn, L, al, s, add(getU(dIn, L, al, s), scale(rho, acc)));
        }
    }
}

// calc_dSdUmu! for the Wilson action pushed as beta/2 (plaq + plaq') (src/action/GaugeActions.jl:95-123):
// D_mu(x) = (beta/2) * sum of the six staples = (beta/2) * V_mu(x)^dagger
void orc_wilson_dSdU(double* D, const double* U, const int* dims, double beta) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) setU(D, L, mu, s, scale(beta / 2.0, dag(staple_sum(U, L, x, mu))));
    }
}

// P_mu += factor * TAcoeffs(U_mu * D_mu)  (md_force! tail, src/molecular_dynamics.jl:255-265)
void orc_kick_from_dSdU(double* P, const double* U, const double* D, const int* dims, double factor) {
    Lat L(dims);
#pragma omp parallel for
    for (long s = 0; s < L.V; s++)
        for (int mu = 0; mu < 4; mu++) {
            double c[8];
            ta_coeffs(mul(getU(U, L, mu, s), getU(D, L, mu, s)), c);
            double* p = P + ((long)mu * L.V + s) * 8;
            for (int a = 0; a < 8; a++) p[a] += factor * c[a];
        }
}

// single-matrix probe of the exp pull-back: C, Q, out in the host layout (column-major 3x3); returns 0 if skipped
int orc_exp_pullback(const double* c18, const double* q18, int route, double* out18) {
    M3 C, Q, R = zero3();
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) {
        C.a[i][j] = cd(c18[2 * (i + 3 * j)], c18[2 * (i + 3 * j) + 1]);
        Q.a[i][j] = cd(q18[2 * (i + 3 * j)], q18[2 * (i + 3 * j) + 1]);
    }
    int ok = 1;
    if (route == 1) ok = exp_pullback_closed(C, Q, &R) ? 1 : 0;
    else R = exp_pullback_series(C, Q);
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) { out18[2 * (i + 3 * j)] = R.a[i][j].real(); out18[2 * (i + 3 * j) + 1] = R.a[i][j].imag(); }
    return ok;
}

}  // extern "C"

extern "C" {

// ----------------------------------------------------------------------------- single-matrix probes (unit tests)

void orc_exp_ta(const double* c8, double t, int route, double* out18) {
    M3 e = exp_ta(c8, t, route);
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) { out18[2 * (i + 3 * j)] = e.a[i][j].real(); out18[2 * (i + 3 * j) + 1] = e.a[i][j].imag(); }
}
void orc_ta_coeffs(const double* m18, double* c8) {
    M3 m;
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) m.a[i][j] = cd(m18[2 * (i + 3 * j)], m18[2 * (i + 3 * j) + 1]);
    ta_coeffs(m, c8);
}
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void orc_staple_sum(double* V, const double* U, const int* dims) {
    Lat L(dims);
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) setU(V, L, mu, s, staple_sum(U, L, x, mu));
    }
}

}  // extern "C"

// ================================================================================================
// General-action path: plaquette + rectangle terms, loop sums, topological charge.
// Restates the GENERIC machinery of the reference (not the hand-derived staples the CUDA kernels use): an action term is a
// coefficient times a set of closed loops plus their adjoints (GaugeAction/push!, src/action/GaugeActions.jl:23-62;
// Gradientflow_general appends loop' itself, src/smearing/gradientflow.jl:88-101); dS/dU_mu is obtained by deleting U_mu
// from every loop that contains it (make_staple in Wilsonloop.jl; calc_dSdUmu!, GaugeActions.jl:95-123), so
// U_mu(x) dSdU_mu(x) = sum over loops W in (set + adjoint set) and over the +mu steps k of W of the loop re-started at that
// step and evaluated with the step at x.
// ================================================================================================
namespace {

typedef std::vector<Step> Loop;

// "plaquette": (mu,1)(nu,1)(mu,-1)(nu,-1) for mu < nu;  "rectangular": (mu,1)(nu,2)(mu,-1)(nu,-2) and (mu,2)(nu,1)(mu,-2)(nu,-1)
// (src/autostaples/wilsonloops.jl:219-245)
std::vector<Loop> loop_set(int kind) {
    std::vector<Loop> out;
    auto seg = [](Loop& l, int mu, int n) { for (int k = 0; k < std::abs(n); k++) l.push_back(Step{mu, n > 0 ? 1 : -1}); };
    for (int mu = 0; mu < 4; mu++)
        for (int nu = mu + 1; nu < 4; nu++) {
            if (kind == 0) {
                Loop l; seg(l, mu, 1); seg(l, nu, 1); seg(l, mu, -1); seg(l, nu, -1); out.push_back(l);
            } else {
                Loop a; seg(a, mu, 1); seg(a, nu, 2); seg(a, mu, -1); seg(a, nu, -2); out.push_back(a);
                Loop b; seg(b, mu, 2); seg(b, nu, 1); seg(b, mu, -2); seg(b, nu, -1); out.push_back(b);
            }
        }
    return out;
}
// adjoint loop: reversed path with reversed steps (Wilsonline adjoint)
Loop adjoint_loop(const Loop& l) {
    Loop r;
    for (int k = (int)l.size() - 1; k >= 0; k--) r.push_back(Step{l[k].mu, -l[k].sgn});
    return r;
}
// all loops of (set + adjoints) re-started at each of their +mu steps: closed paths that begin with U_mu(x)
std::vector<Loop> rotations_through(int kind, int mu) {
    std::vector<Loop> out;
    std::vector<Loop> all = loop_set(kind);
    const size_t n0 = all.size();
    for (size_t i = 0; i < n0; i++) all.push_back(adjoint_loop(all[i]));
    for (const Loop& l : all)
        for (size_t k = 0; k < l.size(); k++)
            if (l[k].mu == mu && l[k].sgn > 0) {
                Loop r;
                for (size_t j = 0; j < l.size(); j++) r.push_back(l[(k + j) % l.size()]);
                out.push_back(r);
            }
    return out;
}

M3 U_dSdU_general(const double* U, const Lat& L, const int* x, const std::vector<Loop>& plaq, const std::vector<Loop>& rect, double c_plaq, double c_rect) {
    M3 acc = zero3();
    if (c_plaq != 0.0) for (const Loop& l : plaq) acc = add(acc, scale(c_plaq, path_product(U, L, x, l.data(), (int)l.size())));
    if (c_rect != 0.0) for (const Loop& l : rect) acc = add(acc, scale(c_rect, path_product(U, L, x, l.data(), (int)l.size())));
    return acc;
}

int eps4(int a, int b, int c, int d) {
    int p[4] = {a, b, c, d};
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) if (p[i] == p[j]) return 0;
    int inv = 0;
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) inv += p[i] > p[j];
    return (inv % 2 == 0) ? 1 : -1;
}

}  // namespace

extern "C" {

// scale = -1/3: md_force! (molecular_dynamics.jl:251-267);  scale = 1: F_update! of the general flow (gradientflow.jl:318-334)
void orc_force_general(double* F, const double* U, const int* dims, double c_plaq, double c_rect, double scale_) {
    Lat L(dims);
    std::vector<Loop> plaq[4], rect[4];
    for (int mu = 0; mu < 4; mu++) { plaq[mu] = rotations_through(0, mu); rect[mu] = rotations_through(1, mu); }
#pragma omp parallel for
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (int mu = 0; mu < 4; mu++) {
            double c[8];
            ta_coeffs(U_dSdU_general(U, L, x, plaq[mu], rect[mu], c_plaq, c_rect), c);
            double* f = F + ((long)mu * L.V + s) * 8;
            for (int a = 0; a < 8; a++) f[a] = scale_ * c[a];
        }
    }
}
// number of loop rotations through a link: 6 for the plaquette set, 18 for the rectangular set (sanity probe for the tests)
int orc_rotations_through(int kind, int mu) { return (int)rotations_through(kind, mu).size(); }

// out2 = { sum_x sum_{plaquette loops} Re tr, sum_x sum_{rectangular loops} Re tr } (loops without their adjoints)
void orc_loop_sums(const double* U, const int* dims, double* out2) {
    Lat L(dims);
    const std::vector<Loop> plaq = loop_set(0), rect = loop_set(1);
    double sp = 0.0, sr = 0.0;
#pragma omp parallel for reduction(+ : sp, sr)
    for (long s = 0; s < L.V; s++) {
        int x[4];
        L.coord(s, x);
        for (const Loop& l : plaq) sp += trace(path_product(U, L, x, l.data(), (int)l.size())).real();
        for (const Loop& l : rect) sr += trace(path_product(U, L, x, l.data(), (int)l.size())).real();
    }
    out2[0] = sp; out2[1] = sr;
}

// topological_charge_density(U; method) (src/AbstractGaugefields.jl:1184-1400): method 0 plaquette, 1 clover, 2 improved.
// q(x) = -Re sum_{mu nu rho sigma} eps tr(F_munu F_rhosigma) / (32 pi^2 n^2) with F = TA(sum of the method's loops)
void orc_topological_charge_density(const double* U, const int* dims, int method, double* out) {
    Lat L(dims);
    auto field_loops = [](int kind, int mu, int nu) {
        std::vector<Loop> ls;
        auto mk = [&](std::initializer_list<std::pair<int, int>> segs) {
            Loop l;
            for (auto sg : segs) for (int k = 0; k < std::abs(sg.second); k++) l.push_back(Step{sg.first, sg.second > 0 ? 1 : -1});
            ls.push_back(l);
        };
        if (kind == 0) {
            mk({{mu, 1}, {nu, 1}, {mu, -1}, {nu, -1}});
        } else if (kind == 1) {  // make_cloverloops, src/autostaples/wilsonloops.jl:166-177
            mk({{mu, 1}, {nu, 1}, {mu, -1}, {nu, -1}});
            mk({{nu, 1}, {mu, -1}, {nu, -1}, {mu, 1}});
            mk({{nu, -1}, {mu, 1}, {nu, 1}, {mu, -1}});
            mk({{mu, -1}, {nu, -1}, {mu, 1}, {nu, 1}});
        } else {  // _rectangle_loops, src/AbstractGaugefields.jl:1316-1330
            mk({{mu, 2}, {nu, 1}, {mu, -2}, {nu, -1}});
            mk({{nu, 1}, {mu, -2}, {nu, -1}, {mu, 2}});
            mk({{nu, -1}, {mu, 2}, {nu, 1}, {mu, -2}});
            mk({{mu, -2}, {nu, -1}, {mu, 2}, {nu, 1}});
            mk({{mu, 1}, {nu, 2}, {mu, -1}, {nu, -2}});
            mk({{nu, 2}, {mu, -1}, {nu, -2}, {mu, 1}});
            mk({{nu, -2}, {mu, 1}, {nu, 2}, {mu, -1}});
            mk({{mu, -1}, {nu, -2}, {mu, 1}, {nu, 2}});
        }
        return ls;
    };
    auto density = [&](int kind, double weight, bool accumulate) {
        const double nl = kind == 0 ? 1.0 : (kind == 1 ? 4.0 : 8.0), rect_factor = kind == 2 ? 2.0 : 1.0;
        std::vector<Loop> loops[4][4];
        for (int mu = 0; mu < 4; mu++) for (int nu = 0; nu < 4; nu++) if (mu != nu) loops[mu][nu] = field_loops(kind, mu, nu);
#pragma omp parallel for
        for (long s = 0; s < L.V; s++) {
            int x[4];
            L.coord(s, x);
            M3 Fs[4][4];
            for (int mu = 0; mu < 4; mu++)
                for (int nu = 0; nu < 4; nu++) {
                    if (mu == nu) continue;
                    M3 w = zero3();
                    for (const Loop& l : loops[mu][nu]) w = add(w, path_product(U, L, x, l.data(), (int)l.size()));
                    Fs[mu][nu] = ta_matrix(w);
                }
            cd q = 0.0;
            for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) for (int c = 0; c < 4; c++) for (int d = 0; d < 4; d++) {
                const int e = eps4(a, b, c, d);
                if (e != 0) q += (double)e * trace(mul(Fs[a][b], Fs[c][d]));
            }
            const double v = weight * (-rect_factor * q.real() / (32.0 * M_PI * M_PI * nl * nl));
            out[s] = accumulate ? out[s] + v : v;
        }
    };
    if (method == 2) { density(1, 5.0 / 3.0, false); density(2, -1.0 / 12.0, true); }
    else density(method, 1.0, false);
}

}  // extern "C"

// ================================================================================================
// Heatbath and overrelaxation for the SU(3) Wilson action.  Restates the site algorithm of the reference
// (src/heatbath/portable/kernels.jl:14-270: project_onto_SU2!, _su2_update_kp_core!, _su3_update_subgroup!, the fixed subgroup
// sequence (1,2),(2,3),(1,3); overrelaxation: src/heatbath/heatbathmodule.jl:1243-1322) and the sweep order of
// heatbath!(U, ::Heatbath) (direction, then even / odd sites, :481-650).  The random STREAMS are this repository's
// (Philox keyed by seed, sweep, direction, colour, subgroup and global site; the reference's bits come from the un-vendored
// LatticeMatrices.jl -- parity unpinned), drawn in the reference's order: (R, R'), (R'', R''') per try, then
// (phi, cos theta).
// ================================================================================================
namespace {

const uint32_t TAG_HEATBATH = 0x48424154u, TAG_OVERRELAX = 0x4f56524cu;

inline void hb_stream_key(uint64_t seed, uint64_t sweep, uint32_t direction, uint32_t colour, uint32_t subgroup, uint32_t tag, uint32_t key[2]) {
    uint32_t ctr[4] = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)sweep, (uint32_t)(sweep >> 32)};
    uint32_t k[2] = {tag, direction | (colour << 8) | (subgroup << 16)};
    uint32_t o[4];
    philox4x32_10(ctr, k, o);
    key[0] = o[0]; key[1] = o[1];
}

struct S2 { cd a[2][2]; };

inline S2 sub_projected(const M3& uv, int n, int m) {
    S2 s;
    s.a[0][0] = uv.a[n][n]; s.a[0][1] = uv.a[n][m]; s.a[1][0] = uv.a[m][n]; s.a[1][1] = uv.a[m][m];
    cd alpha = 0.5 * s.a[0][0] + 0.5 * std::conj(s.a[1][1]);
    cd beta = 0.5 * s.a[1][0] - 0.5 * std::conj(s.a[0][1]);
    s.a[0][0] = alpha; s.a[1][0] = beta; s.a[0][1] = -std::conj(beta); s.a[1][1] = std::conj(alpha);
    return s;
}
inline M3 embed(const S2& k, int n, int m) {
    M3 a = ident3();
    a.a[n][n] = k.a[0][0]; a.a[n][m] = k.a[0][1]; a.a[m][n] = k.a[1][0]; a.a[m][m] = k.a[1][1];
    return a;
}
inline S2 mul2(const S2& x, const S2& y) {
    S2 r;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) r.a[i][j] = x.a[i][0] * y.a[0][j] + x.a[i][1] * y.a[1][j];
    return r;
}

// _su2_update_kp_core! (portable/kernels.jl:63-150)
inline bool su2_update_kp(const S2& V, double beta, const uint32_t key[2], uint64_t gsite, S2* out) {
    const double rho0 = (V.a[0][0] + V.a[1][1]).real() / 2, rho1 = -(V.a[0][1] + V.a[1][0]).imag() / 2;
    const double rho2 = (V.a[1][0] - V.a[0][1]).real() / 2, rho3 = (V.a[1][1] - V.a[0][0]).imag() / 2;
    const double rho = std::sqrt(rho0 * rho0 + rho1 * rho1 + rho2 * rho2 + rho3 * rho3);
    const cd det = V.a[0][0] * V.a[1][1] - V.a[0][1] * V.a[1][0];
    S2 V0;
    V0.a[0][0] = rho * V.a[1][1] / det; V0.a[0][1] = -rho * V.a[0][1] / det;
    V0.a[1][0] = -rho * V.a[1][0] / det; V0.a[1][1] = rho * V.a[0][0] / det;
    const double k = 2.0 * (beta / 3.0) * rho;
    uint32_t draw = 0;
    double delta = 0.0;
    bool accepted = false;
    for (int tries = 0; tries < 100000; tries++) {
        double R, Rp, Rpp, Rppp;
        site_uniform_pair(key, gsite, draw++, &R, &Rp);
        site_uniform_pair(key, gsite, draw++, &Rpp, &Rppp);
        const double X = -std::log(1.0 - R) / k, Xp = -std::log(1.0 - Rp) / k;
        const double c = std::cos(2.0 * M_PI * Rpp);
        delta = Xp + X * c * c;
        if (Rppp * Rppp <= 1.0 - 0.5 * delta) { accepted = true; break; }
    }
    if (!accepted) return false;
    const double a1 = 1.0 - delta, rr = std::sqrt(std::max(1.0 - a1 * a1, 0.0));
    double uphi, ucos;
    site_uniform_pair(key, gsite, draw++, &uphi, &ucos);
    const double phi = 2.0 * M_PI * uphi, costheta = 2.0 * (ucos - 0.5), sintheta = std::sqrt(std::max(1.0 - costheta * costheta, 0.0));
    const double a2 = rr * std::cos(phi) * sintheta, a3 = rr * std::sin(phi) * sintheta, a4 = rr * costheta;
    S2 temp;
    temp.a[0][0] = cd(a1, a4); temp.a[0][1] = cd(a3, a2); temp.a[1][0] = cd(-a3, a2); temp.a[1][1] = cd(a1, -a4);
    S2 U = mul2(temp, V0);
    const cd alpha = 0.5 * U.a[0][0] + 0.5 * std::conj(U.a[1][1]), b2 = 0.5 * U.a[1][0] - 0.5 * std::conj(U.a[0][1]);
    const double detU = std::norm(alpha) + std::norm(b2);
    out->a[0][0] = alpha / detU; out->a[1][0] = b2 / detU; out->a[0][1] = -std::conj(b2) / detU; out->a[1][1] = std::conj(alpha) / detU;
    return true;
}

}  // namespace

extern "C" {

// one sweep; returns the number of failed sites (0 = ok).  overrelax = 0: heatbath, 1: overrelaxation
int orc_heatbath_sweep(double* U, const int* dims, double beta, uint64_t seed, uint64_t sweep, int overrelax) {
    Lat L(dims);
    int failures = 0;
    for (int mu = 0; mu < 4; mu++)
        for (int colour = 0; colour < 2; colour++) {
            uint32_t keys[3][2];
            for (uint32_t s = 0; s < 3; s++) hb_stream_key(seed, sweep, (uint32_t)(mu + 1), (uint32_t)colour, s, overrelax ? TAG_OVERRELAX : TAG_HEATBATH, keys[s]);
            // in place: the staples of a link of one colour hold no mu-link of that colour
#pragma omp parallel for reduction(+ : failures)
            for (long s = 0; s < L.V; s++) {
                int x[4];
                L.coord(s, x);
                if (((x[0] + x[1] + x[2] + x[3]) & 1) != colour) continue;
                const M3 V = dag(staple_sum(U, L, x, mu));  // tr(u V) = the six plaquettes through the link
                M3 u = getU(U, L, mu, s);
                bool ok = true;
                if (!overrelax) {
                    const int sub[3][2] = {{0, 1}, {1, 2}, {0, 2}};
                    for (int k = 0; k < 3 && ok; k++) {
                        const S2 w = sub_projected(mul(u, V), sub[k][0], sub[k][1]);
                        S2 K;
                        ok = su2_update_kp(w, beta, keys[k], (uint64_t)s, &K);
                        if (ok) u = mul(embed(K, sub[k][0], sub[k][1]), u);
                    }
                } else {
                    for (uint32_t k = 0; k < 3 && ok; k++) {
                        double u0, u1;
                        site_uniform_pair(keys[0], (uint64_t)s, k, &u0, &u1);
                        const int n = (int)(u0 * 2.0), m = n + 1 + (int)(u1 * (double)(2 - n));
                        const S2 w = sub_projected(mul(u, V), n, m);
                        S2 wd;  // w^dagger
                        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) wd.a[i][j] = std::conj(w.a[j][i]);
                        S2 h = mul2(wd, wd);
                        const double nrm = std::sqrt(std::norm(h.a[0][0]) + std::norm(h.a[1][0]));
                        if (!(nrm > 0.0)) { ok = false; break; }
                        const cd al = h.a[0][0] / nrm, be = h.a[1][0] / nrm;
                        h.a[0][0] = al; h.a[1][0] = be; h.a[0][1] = -std::conj(be); h.a[1][1] = std::conj(al);
                        u = mul(embed(h, n, m), u);
                    }
                }
                if (!ok) { failures++; continue; }
                setU(U, L, mu, s, reunitarize(u));
            }
        }
    return failures;
}

}  // extern "C"
