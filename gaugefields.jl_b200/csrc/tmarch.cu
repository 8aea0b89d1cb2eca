// tmarch.cu -- the fused staple -> TA force -> kick -> exp(eps P) U pass as a persistent t-marching kernel.
//
// Same result as k_force_fused (kernels.cu) to rounding; different data movement and fewer FP64 instructions.
//
// Why (profiles/r1_ncu_force_fused.md): k_force_fused requests 19 link matrices per link through L1 (10.9 KB/site), 6.0 KB/site of
// which come from L2 at the ~7 TB/s the L2->SM path delivers at 16 warps/SM; neither DRAM nor the FP64 pipe is the limit.
// Here a CTA owns a spatial tile of 8x4x2 sites and marches along t.  Every link matrix the six-staple stencil of a slice needs is
// copied into shared memory ONCE per tile and slice with 16-byte cp.async (LDGSTS, L1-bypassing), one whole slice ahead of its
// use, so L2->SM traffic drops to 2.64 matrix loads per link (1.5 KB/site) and every operand is a fixed-latency LDS.128.
// The backward-t staple is carried in registers from the previous slice by the thread that owns the link, so slice t-1 is never
// resident (tmarch_geom.h has the exact residency sets and the ring layout: 230400 bytes of shared memory, one CTA per SM).
//
// FP64 work: links are SU(3), so every staple A B C is formed from the first two rows of A only (2 x 72 FMA) and its third row
// is reconstructed as conj(row0 x row1) folded into the accumulation (24 FMA + 12 adds): 180 instead of 216 FP64 instructions per
// staple.  The staple sum itself is not unitary, so U V^dag, the TA projection, the exponential and exp*U stay full 3x3.
//
// One thread per (site, mu); a warp holds 32 sites of one direction (mu is a template parameter of the per-warp body, so every
// operand's ring and part are compile-time and its byte offset is one of 19 per-thread registers computed once per CTA).
#include <cstdint>
#include <cstdlib>

#include "gfb_internal.h"
#include "stencil.cuh"
#include "tmarch_geom.h"

#ifndef GFB_TM_DEBUG
#define GFB_TM_DEBUG 0  // 1: no staple arithmetic (copies + operand reads only); 2: no global->shared copies (arithmetic on stale smem)
#endif

namespace gfb {

namespace {

struct TmPlan {
    int t_begin, t_count;  // local slices covered by the launch (contiguous)
    int seg_len, nseg;     // t-segments: item = (segment, tile)
    int ntx, nty, ntz, ntiles;
};

struct R2 {
    double2 e[6];  // rows 0 and 1
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
#if GFB_TM_DEBUG != 2
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ M3 lds_m3(const unsigned char* p) {
    M3 r;
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = q[k];
    return r;
}
__device__ __forceinline__ R2 lds_rows01(const unsigned char* p) {
    R2 r;
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int k = 0; k < 6; k++) r.e[k] = q[k];
    return r;
}
// rows 0,1 of A^dagger: (A^dag)[i][j] = conj(A[j][i])
__device__ __forceinline__ R2 lds_dag_rows01(const unsigned char* p) {
    R2 r;
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double2 v = q[3 * j + i];
            r.e[3 * i + j] = make_double2(v.x, -v.y);
        }
    return r;
}
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

__device__ __forceinline__ int wrap(int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); }

// The whole persistent loop of one link-thread with direction MU.
template <int MU, bool READ_Z, bool WRITE_Z, bool DO_EXP>
__device__ __forceinline__ void tm_run(const Geom& g, const TmPlan& pl, const double2* __restrict__ uin, double2* __restrict__ uout,
                                       const double* __restrict__ zin, double* __restrict__ zout, double a, double b, double c,
                                       unsigned char* smem, const tm::Box* boxes) {
    const int tid = threadIdx.x;
    const int sidx = tid & (tm::SITES - 1);
    const int sx = sidx & (tm::BX - 1), sy = (sidx / tm::BX) & (tm::BY - 1), sz = sidx / (tm::BX * tm::BY);
    unsigned char* const sS = smem;
    unsigned char* const sR = smem + tm::S_RING * tm::S_BYTES;

    // ---- consumer side: operand offsets (tile independent)
    tm::Operands op;
    tm::make_operands(boxes, sx, sy, sz, MU, &op);

    // ---- producer side: this thread copies S slot tid and R slots tid, tid+256 of every slice
    int plam[3], px[3], py[3], pz[3];
    bool pvalid[3];
    pvalid[0] = tm::slot_to_pos(boxes, 0, tid, &plam[0], &px[0], &py[0], &pz[0]);
    pvalid[1] = tm::slot_to_pos(boxes, 1, tid, &plam[1], &px[1], &py[1], &pz[1]);
    pvalid[2] = tm::slot_to_pos(boxes, 1, tid + tm::NTHREADS, &plam[2], &px[2], &py[2], &pz[2]);
    const unsigned sb = (unsigned)g.v3 * 16u;          // bytes between element planes
    const size_t slice_bytes = (size_t)36 * sb;        // bytes of one time-slice of links

    const long nitems = (long)pl.ntiles * pl.nseg;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int seg = (int)(item / pl.ntiles);
        int tile = (int)(item % pl.ntiles);
        const int x0 = (tile % pl.ntx) * tm::BX; tile /= pl.ntx;
        const int y0 = (tile % pl.nty) * tm::BY;
        const int z0 = (tile / pl.nty) * tm::BZ;
        const int tb = pl.t_begin + seg * pl.seg_len;
        const int len = min(pl.seg_len, pl.t_begin + pl.t_count - tb);

        unsigned psrc[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int s3 = wrap(x0 + px[q], g.nx) + g.nx * (wrap(y0 + py[q], g.ny) + g.ny * wrap(z0 + pz[q], g.nz));
            psrc[q] = pvalid[q] ? ((unsigned)(plam[q] * 9) * (unsigned)g.v3 + (unsigned)s3) * 16u : 0u;
        }
        auto copy_mat = [&](int q, int tslot, unsigned char* dst_part, int slot) {
            if (!pvalid[q]) return;
            const char* src = reinterpret_cast<const char*>(uin) + (size_t)tslot * slice_bytes + psrc[q];
            const unsigned dst = smem_u32(dst_part + slot * tm::MAT_BYTES);
#pragma unroll
            for (int k = 0; k < 9; k++) cp_async16(dst + 16 * k, src + (size_t)k * sb);
        };
        auto copy_S = [&](int tslot, int ring) { copy_mat(0, tslot, sS + ring * tm::S_BYTES, tid); };
        auto copy_R = [&](int tslot, int ring) {
            copy_mat(1, tslot, sR + ring * tm::R_BYTES, tid);
            copy_mat(2, tslot, sR + ring * tm::R_BYTES, tid + tm::NTHREADS);
        };
        auto t_up = [&](int t) { return (t == g.tloc - 1) ? g.t_up_wrap : t + 1; };

        // ---- prologue: slice tb (full) and the S part of slice tb+1; the backward-t staple of slice tb from global memory
        copy_S(tb, 0);
        copy_R(tb, 0);
        copy_S(t_up(tb), 1);
        cp_async_commit();

        Coord x;
        x.x = x0 + sx; x.y = y0 + sy; x.z = z0 + sz; x.t = tb;
        const unsigned s3 = (unsigned)s3_of(g, x);
        R2 G;
        if (MU < 3) {
            const Coord y = step(g, x, 3, -1);
            const Coord ym = step(g, y, MU, +1);
            const M3 A = load_link(uin, g, y, 3);
            const M3 U = load_link(uin, g, y, MU);
            const M3 C = load_link(uin, g, ym, 3);
            G = r2_mul_nn(r2_mul_nn(rows01_dag(A), U), C);
        }
        cp_async_wait_all();
        __syncthreads();

        int rs = 0;  // j % 3
        for (int j = 0; j < len; j++) {
            const int t = tb + j;
            const int rs1 = (rs == 2) ? 0 : rs + 1;   // (j+1) % 3
            const int rs2 = (rs1 == 2) ? 0 : rs1 + 1; // (j+2) % 3
            if (j + 1 < len) {
                copy_R(t + 1, (j + 1) & 1);
                copy_S(t_up(t + 1), rs2);
            }
            cp_async_commit();

            const unsigned char* const Sc = sS + rs * tm::S_BYTES;
            const unsigned char* const Sn = sS + rs1 * tm::S_BYTES;
            const unsigned char* const Rc = sR + (j & 1) * tm::R_BYTES;
            auto cen = [&](int off) -> const unsigned char* { return ((off & 1) ? Rc : Sc) + (off & ~1); };
            auto nxt = [&](int off) -> const unsigned char* { return Sn + (off & ~1); };

            const unsigned zo = (unsigned)(t * 32 + MU * 8) * (unsigned)g.v3 + s3;
            const unsigned zsb = (unsigned)g.v3 * 8u;
            double z[8];
            if (READ_Z) {
#pragma unroll
                for (int k = 0; k < 8; k++) z[k] = __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(zin + zo) + (size_t)k * zsb));
            }

            const M3 U = lds_m3(cen(op.own));
            M3 V;
            R2 Gn;
            if (MU < 3) V = complete_su3(G);
            else V = m3_zero();
#pragma unroll
            for (int jj = 0; jj < 3; jj++) {
                const int nu = (MU + 1 + jj) & 3;
#if GFB_TM_DEBUG == 1
                m3_add(V, lds_m3(cen(op.up[jj][0])));
                m3_add(V, lds_m3((MU < 3 && nu == 3) ? nxt(op.up[jj][1]) : cen(op.up[jj][1])));
                m3_add(V, lds_m3((MU == 3) ? nxt(op.up[jj][2]) : cen(op.up[jj][2])));
                if (nu < 3) {
                    m3_add(V, lds_m3(cen(op.dn[jj][0])));
                    m3_add(V, lds_m3(cen(op.dn[jj][1])));
                    m3_add(V, lds_m3((MU == 3) ? nxt(op.dn[jj][2]) : cen(op.dn[jj][2])));
                }
                if (MU < 3) Gn = G;
                continue;
#endif
                if (MU < 3 && nu < 3) {
                    {
                        const R2 A = lds_rows01(cen(op.up[jj][0]));
                        const M3 B = lds_m3(cen(op.up[jj][1]));
                        const R2 T = r2_mul_nn(A, B);
                        const M3 C = lds_m3(cen(op.up[jj][2]));
                        acc_su3(V, r2_mul_nd(T, C));
                    }
                    {
                        const R2 A = lds_dag_rows01(cen(op.dn[jj][0]));
                        const M3 B = lds_m3(cen(op.dn[jj][1]));
                        const R2 T = r2_mul_nn(A, B);
                        const M3 C = lds_m3(cen(op.dn[jj][2]));
                        acc_su3(V, r2_mul_nn(T, C));
                    }
                } else if (MU < 3) {  // nu = t: upper staple from slices t, t+1; the lower one was carried in G; next G from slice t
                    const M3 A = lds_m3(cen(op.up[jj][0]));
                    const M3 C = lds_m3(cen(op.up[jj][2]));
                    {
                        const M3 B = lds_m3(nxt(op.up[jj][1]));
                        const R2 T = r2_mul_nn(rows01(A), B);
                        acc_su3(V, r2_mul_nd(T, C));
                    }
                    {
                        const R2 T = r2_mul_nn(rows01_dag(A), U);
                        Gn = r2_mul_nn(T, C);
                    }
                } else {  // MU = t, nu spatial
                    {
                        const R2 A = lds_rows01(cen(op.up[jj][0]));
                        const M3 B = lds_m3(cen(op.up[jj][1]));
                        const R2 T = r2_mul_nn(A, B);
                        const M3 C = lds_m3(nxt(op.up[jj][2]));
                        acc_su3(V, r2_mul_nd(T, C));
                    }
                    {
                        const R2 A = lds_dag_rows01(cen(op.dn[jj][0]));
                        const M3 B = lds_m3(cen(op.dn[jj][1]));
                        const R2 T = r2_mul_nn(A, B);
                        const M3 C = lds_m3(nxt(op.dn[jj][2]));
                        acc_su3(V, r2_mul_nn(T, C));
                    }
                }
            }
            if (MU < 3) G = Gn;

            double f[8];
            {
                const M3 w = mul_nd(U, V);
                ta_coeffs(w, f);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double v = a * f[k];
                if (READ_Z) v = fma(b, z[k], v);
                f[k] = v;
                if (WRITE_Z) *reinterpret_cast<double*>(reinterpret_cast<char*>(zout + zo) + (size_t)k * zsb) = v;
            }
            if (DO_EXP) {
                const M3 e = exp_ta(f, c);
                const M3 r = mul_nn(e, U);
                const unsigned uo = (unsigned)(t * 36 + MU * 9) * (unsigned)g.v3 + s3;
                m3_store(uout + uo, (unsigned)g.v3, r);
            }

            cp_async_wait_all();
            __syncthreads();
            rs = rs1;
        }
    }
}

template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__global__ void __launch_bounds__(tm::NTHREADS, 1)
k_tmarch_fused(Geom g, TmPlan pl, const double2* __restrict__ uin, double2* __restrict__ uout, const double* __restrict__ zin,
               double* __restrict__ zout, double a, double b, double c) {
    extern __shared__ __align__(128) unsigned char smem[];
    tm::Box* const boxes = reinterpret_cast<tm::Box*>(smem + tm::SMEM_DATA);
    if (threadIdx.x == 0) tm::make_boxes(boxes);
    __syncthreads();
    const int mu = threadIdx.x / tm::SITES;  // warp-uniform: two warps per direction
    switch (mu) {
        case 0: tm_run<0, READ_Z, WRITE_Z, DO_EXP>(g, pl, uin, uout, zin, zout, a, b, c, smem, boxes); break;
        case 1: tm_run<1, READ_Z, WRITE_Z, DO_EXP>(g, pl, uin, uout, zin, zout, a, b, c, smem, boxes); break;
        case 2: tm_run<2, READ_Z, WRITE_Z, DO_EXP>(g, pl, uin, uout, zin, zout, a, b, c, smem, boxes); break;
        default: tm_run<3, READ_Z, WRITE_Z, DO_EXP>(g, pl, uin, uout, zin, zout, a, b, c, smem, boxes); break;
    }
}

constexpr size_t kTmSmem = tm::SMEM_DATA + tm::NBOX * sizeof(tm::Box);

// t-segments: enough (segment, tile) items to fill the SMs evenly, as few segment prologues as possible
TmPlan make_plan(const Geom& g, int t_begin, int t_count, int nsm) {
    TmPlan pl;
    pl.t_begin = t_begin; pl.t_count = t_count;
    pl.ntx = g.nx / tm::BX; pl.nty = g.ny / tm::BY; pl.ntz = g.nz / tm::BZ;
    pl.ntiles = pl.ntx * pl.nty * pl.ntz;
    double best = 1e300;
    int best_nseg = 1;
    const double prologue = 1.5;  // cost of a segment start in units of one slice step (exposed first copies + G from global memory)
    for (int nseg = 1; nseg <= t_count; nseg++) {
        const int len = (t_count + nseg - 1) / nseg;
        const int real_nseg = (t_count + len - 1) / len;
        const long items = (long)pl.ntiles * real_nseg;
        const long rounds = (items + nsm - 1) / nsm;
        const double cost = (double)rounds * (len + prologue);
        if (cost < best - 1e-9) { best = cost; best_nseg = real_nseg; }
    }
    pl.seg_len = (t_count + best_nseg - 1) / best_nseg;
    pl.nseg = (t_count + pl.seg_len - 1) / pl.seg_len;
    return pl;
}

}  // namespace

// Returns false when the launch is not covered (tile does not divide the lattice, strided slice set, in-place links):
// the caller then uses k_force_fused.  GFB200_TMARCH=0 disables the kernel.
bool launch_tmarch_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                         const FusedArgs& fa) {
    const char* em = getenv("GFB200_TMARCH");  // read per launch so that tests can compare both kernels in one process
    const int mode = em ? atoi(em) : 1;
    if (!mode) return false;
    if (g.t_stride != 1 || t_count < 2) return false;
    if (g.nx % tm::BX || g.ny % tm::BY || g.nz % tm::BZ) return false;
    if (uout == uin) return false;
    static int nsm = 0;
    if (nsm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    TmPlan pl = make_plan(g, t_begin, t_count, nsm);
    if (const char* e = getenv("GFB200_TMARCH_SEGLEN")) {  // test hook: force the t-segment length
        const int len = atoi(e);
        if (len >= 1) { pl.seg_len = len < t_count ? len : t_count; pl.nseg = (t_count + pl.seg_len - 1) / pl.seg_len; }
    }
    const long nitems = (long)pl.ntiles * pl.nseg;
    const unsigned grid = (unsigned)(nitems < nsm ? nitems : nsm);
#define GFB_LAUNCH_TM(R, W, E)                                                                                                  \
    do {                                                                                                                        \
        auto kern = k_tmarch_fused<R, W, E>;                                                                                    \
        static bool attr_set = false;                                                                                           \
        if (!attr_set) {                                                                                                        \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmSmem) != cudaSuccess) {         \
                cudaGetLastError();                                                                                             \
                return false;                                                                                                   \
            }                                                                                                                   \
            attr_set = true;                                                                                                    \
        }                                                                                                                       \
        kern<<<grid, tm::NTHREADS, kTmSmem, st>>>(g, pl, uin, uout, zin, zout, fa.a, fa.b, fa.c);                                \
    } while (0)
    if (fa.read_z) {
        if (fa.do_exp) GFB_LAUNCH_TM(true, true, true);
        else GFB_LAUNCH_TM(true, true, false);
    } else {
        if (fa.do_exp) {
            if (fa.write_z) GFB_LAUNCH_TM(false, true, true);
            else GFB_LAUNCH_TM(false, false, true);
        } else GFB_LAUNCH_TM(false, true, false);
    }
#undef GFB_LAUNCH_TM
    return true;
}

}  // namespace gfb
