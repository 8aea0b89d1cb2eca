// heatbath.cu -- the other quenched updater: Cabibbo-Marinari pseudo-heatbath (Kennedy-Pendleton SU(2) sampling) and
// microcanonical overrelaxation for the SU(3) Wilson action, one checkerboard colour of one direction per launch.
//
// Reference: heatbath!(U, ::Heatbath) / overrelaxation! -> heatbath_su3_sites! / _overrelaxation_sites!
// (src/heatbath/heatbathmodule.jl:1799-1860, 1383-1436) with the site kernels of src/heatbath/portable/kernels.jl:
//   * SU(3) update = the fixed subgroup sequence (1,2), (2,3), (1,3), each: UV = u V, S = 2x2 block of UV projected on
//     r * SU(2) (project_onto_SU2!, :14-25), K ~ Kennedy-Pendleton(S, beta) (_su2_update_kp_core!, :63-150), u <- embed(K) u
//     (:205-235), finally reunitarise (heatbath_normalize3!);
//   * overrelaxation = three subgroup hits with RANDOM subgroups n < m (SUN_overrelaxation_rng!, heatbathmodule.jl:1295-1322):
//     h = normalise((w^dag)^2), u <- embed(h) u, which leaves Re tr(u V) unchanged.
// V is the reference's staple sum (tr(u V) = sum of the six plaquettes through the link) = (sum of the paths x -> x+mu)^dagger.
// Random numbers: one counter-based stream per (seed, sweep, direction, colour, subgroup) and GLOBAL site, so a sweep does not
// depend on the slab decomposition; like every random field of this backend the bits are this backend's own (the reference's
// streams live in the un-vendored LatticeMatrices.jl: SURVEY.md 8c, "parity unpinned").
//
// One thread per site of the selected colour; the update is in place: the staples of a link (x, mu) with x of one colour
// contain only links of other directions and mu-links of the other colour.
#include "gfb_internal.h"
#include "stencil.cuh"
#include "su3.cuh"

namespace gfb {

namespace {

constexpr unsigned kTagHeatbath = 0x48424154u;        // "HBAT"
constexpr unsigned kTagOverrelaxation = 0x4f56524cu;  // "OVRL"
constexpr int kIterationMax = 100000;                 // ITERATION_MAX of the reference (heatbathmodule.jl:62)

struct Su2 {  // a0 + i (a1 sigma1 + a2 sigma2 + a3 sigma3) scaled: [[alpha, -conj(beta)], [beta, conj(alpha)]]
    double2 alpha, beta;
};

// the (n, m) 2x2 block of M projected on r * SU(2): alpha = (M_nn + conj(M_mm)) / 2, beta = (M_mn - conj(M_nm)) / 2
__device__ __forceinline__ Su2 project_su2(const M3& mm, int n, int m) {
    const double2 a = mm.e[3 * n + n], b = mm.e[3 * n + m], c = mm.e[3 * m + n], d = mm.e[3 * m + m];
    Su2 s;
    s.alpha = make_double2(0.5 * (a.x + d.x), 0.5 * (a.y - d.y));
    s.beta = make_double2(0.5 * (c.x - b.x), 0.5 * (c.y + b.y));
    return s;
}
// u <- embed(K) u: only rows n and m change
__device__ __forceinline__ void apply_su2(M3& u, const Su2& k, int n, int m) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double2 un = u.e[3 * n + j], um = u.e[3 * m + j];
        // row n: alpha u_n - conj(beta) u_m ; row m: beta u_n + conj(alpha) u_m
        double2 rn = cmul(k.alpha, un);
        rn.x -= k.beta.x * um.x + k.beta.y * um.y;
        rn.y -= k.beta.x * um.y - k.beta.y * um.x;
        double2 rm = cmul(k.beta, un);
        rm.x += k.alpha.x * um.x + k.alpha.y * um.y;
        rm.y += k.alpha.x * um.y - k.alpha.y * um.x;
        u.e[3 * n + j] = rn;
        u.e[3 * m + j] = rm;
    }
}

// Kennedy-Pendleton: K distributed as exp((beta/NC) Re tr(K S)) for S = rho * (SU(2) element); draws come in pairs from the
// site's stream (`draw` counts pairs).  Returns false when no candidate is accepted in kIterationMax tries.
__device__ __forceinline__ bool su2_update_kp(const Su2& s, double beta, unsigned k0, unsigned k1, unsigned long long gsite, unsigned& draw, Su2& out) {
    const double rho = sqrt(s.alpha.x * s.alpha.x + s.alpha.y * s.alpha.y + s.beta.x * s.beta.x + s.beta.y * s.beta.y);
    // V0 = rho * S^{-1} = S^dagger / rho for S = rho * g:  [[conj(alpha), conj(beta)], [-beta, alpha]] / rho
    const double ir = 1.0 / rho;
    const double k = 2.0 * (beta / 3.0) * rho;
    double delta = 0.0;
    bool accepted = false;
    for (int tries = 0; tries < kIterationMax; tries++) {
        double r, rp, rpp, rppp;
        site_uniform_pair(k0, k1, gsite, draw++, r, rp);
        site_uniform_pair(k0, k1, gsite, draw++, rpp, rppp);
        const double x = -log(1.0 - r) / k, xp = -log(1.0 - rp) / k;  // 1 - u lies in (0, 1]
        const double c = cos(6.283185307179586476925286766559 * rpp);
        delta = xp + x * c * c;
        if (rppp * rppp <= 1.0 - 0.5 * delta) { accepted = true; break; }
    }
    if (!accepted) return false;
    const double a1 = 1.0 - delta;
    const double rr = sqrt(fmax(1.0 - a1 * a1, 0.0));
    double uphi, ucos;
    site_uniform_pair(k0, k1, gsite, draw++, uphi, ucos);
    const double costheta = 2.0 * (ucos - 0.5);
    const double sintheta = sqrt(fmax(1.0 - costheta * costheta, 0.0));
    double sphi, cphi;
    sincos(6.283185307179586476925286766559 * uphi, &sphi, &cphi);
    const double a2 = rr * cphi * sintheta, a3 = rr * sphi * sintheta, a4 = rr * costheta;
    // temp = [[a1 + i a4, a3 + i a2], [-a3 + i a2, a1 - i a4]]  -> alpha_t = a1 + i a4, beta_t = -a3 + i a2
    const double2 at = make_double2(a1, a4), bt = make_double2(-a3, a2);
    // Unew = temp * V0, V0 = [[conj(alpha), conj(beta)], [-beta, alpha]] / rho.  First column of Unew:
    //   U11 = at conj(alpha) + (-conj(bt)) (-beta) ;  U21 = bt conj(alpha) + conj(at) (-beta)
    double2 u11 = make_double2(at.x * s.alpha.x + at.y * s.alpha.y, at.y * s.alpha.x - at.x * s.alpha.y);
    u11.x += bt.x * s.beta.x + bt.y * s.beta.y;
    u11.y += bt.x * s.beta.y - bt.y * s.beta.x;
    double2 u21 = make_double2(bt.x * s.alpha.x + bt.y * s.alpha.y, bt.y * s.alpha.x - bt.x * s.alpha.y);
    u21.x -= at.x * s.beta.x + at.y * s.beta.y;
    u21.y -= at.x * s.beta.y - at.y * s.beta.x;
    u11.x *= ir; u11.y *= ir; u21.x *= ir; u21.y *= ir;
    // the product of two SU(2) elements is SU(2): the reference's final re-projection only removes rounding; normalise the same way
    const double det = u11.x * u11.x + u11.y * u11.y + u21.x * u21.x + u21.y * u21.y;
    const double id = 1.0 / det;
    out.alpha = make_double2(u11.x * id, u11.y * id);
    out.beta = make_double2(u21.x * id, u21.y * id);
    return true;
}

struct HbKeys {
    unsigned k0[3], k1[3];  // heatbath: one stream per subgroup; overrelaxation: entry 0 only
};

template <bool OVERRELAX>
__global__ void __launch_bounds__(128) k_heatbath(Geom g, double2* __restrict__ u, int mu, int colour, double beta, HbKeys keys, int* __restrict__ failures) {
    // sites of one colour: enumerate pairs along x (nx is even) and pick the member with the right parity
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long half = (long)g.v3 * g.tloc / 2;
    if (n >= half) return;
    const int hx = g.nx / 2;
    Coord x;
    long r = n;
    const int xh = (int)(r % hx); r /= hx;
    x.y = (int)(r % g.ny); r /= g.ny;
    x.z = (int)(r % g.nz);
    x.t = (int)(r / g.nz);
    const int rest = x.y + x.z + x.t + g.t0;
    x.x = 2 * xh + ((rest + colour) & 1);
    const unsigned long long gsite = global_site_id(g, x);
    M3 vs = staple_sum<true>(u, g, x, mu);  // sum of the paths x -> x+mu; the reference's V is its adjoint
    M3 um = load_link(u, g, x, mu);
    bool ok = true;
    if (!OVERRELAX) {
        const int sub[3][2] = {{0, 1}, {1, 2}, {0, 2}};
#pragma unroll
        for (int s = 0; s < 3; s++) {
            const M3 uv = mul_nd(um, vs);
            const Su2 w = project_su2(uv, sub[s][0], sub[s][1]);
            Su2 k;
            unsigned draw = 0;
            if (!su2_update_kp(w, beta, keys.k0[s], keys.k1[s], gsite, draw, k)) { ok = false; break; }
            apply_su2(um, k, sub[s][0], sub[s][1]);
        }
    } else {
#pragma unroll 1
        for (int s = 0; s < 3; s++) {
            double u0, u1;
            site_uniform_pair(keys.k0[0], keys.k1[0], gsite, (unsigned)s, u0, u1);
            // n uniform in {1, 2}, m uniform in {n+1, .., 3} (rand_bounded semantics of SUN_overrelaxation_rng!)
            const int nn = (int)(u0 * 2.0);                       // 0 or 1
            const int mm = nn + 1 + (int)(u1 * (double)(2 - nn));  // nn = 0: 1 or 2; nn = 1: 2
            const M3 uv = mul_nd(um, vs);
            const Su2 w = project_su2(uv, nn, mm);
            // h = (w^dag)^2 normalised.  For w = [[a, -conj(b)], [b, conj(a)]]: w^dag = [[conj(a), conj(b)], [-b, a]],
            // (w^dag)^2 first column = (conj(a)^2 - |b|^2, -b (conj(a) + a))
            const double2 a = w.alpha, b = w.beta;
            double2 h11 = make_double2(a.x * a.x - a.y * a.y - (b.x * b.x + b.y * b.y), -2.0 * a.x * a.y);
            double2 h21 = make_double2(-2.0 * a.x * b.x, -2.0 * a.x * b.y);
            const double nrm = sqrt(h11.x * h11.x + h11.y * h11.y + h21.x * h21.x + h21.y * h21.y);
            if (!(nrm > 0.0)) { ok = false; break; }
            const double in = 1.0 / nrm;
            Su2 h;
            h.alpha = make_double2(h11.x * in, h11.y * in);
            h.beta = make_double2(h21.x * in, h21.y * in);
            apply_su2(um, h, nn, mm);
        }
    }
    if (!ok) { atomicAdd(failures, 1); return; }
    store_link(u, g, x, mu, reunitarize(um));
}

}  // namespace

static void hb_stream_key(unsigned long long seed, unsigned long long sweep, unsigned direction, unsigned colour, unsigned subgroup, unsigned tag, unsigned* k0, unsigned* k1) {
    unsigned o[4];
    // same construction as the other random fields (kernels.cu, host_stream_key) with (direction, colour, subgroup) packed into one word
    philox4x32_10((unsigned)seed, (unsigned)(seed >> 32), (unsigned)sweep, (unsigned)(sweep >> 32), tag, direction | (colour << 8) | (subgroup << 16), o);
    *k0 = o[0];
    *k1 = o[1];
}

void launch_heatbath(cudaStream_t st, const Geom& g, double2* u, int mu, int colour, double beta, unsigned long long seed, unsigned long long sweep, bool overrelax,
                     int* failures) {
    HbKeys keys;
    for (int s = 0; s < 3; s++)
        hb_stream_key(seed, sweep, (unsigned)(mu + 1), (unsigned)colour, (unsigned)s, overrelax ? kTagOverrelaxation : kTagHeatbath, &keys.k0[s], &keys.k1[s]);
    const long half = (long)g.v3 * g.tloc / 2;
    const unsigned nb = (unsigned)((half + 127) / 128);
    if (overrelax) k_heatbath<true><<<nb, 128, 0, st>>>(g, u, mu, colour, beta, keys, failures);
    else k_heatbath<false><<<nb, 128, 0, st>>>(g, u, mu, colour, beta, keys, failures);
}

}  // namespace gfb
