// tmarch.cu -- the fused staple -> TA force -> kick -> exp(eps P) U pass as a persistent t-marching kernel.
//
// Same result as k_force_fused (kernels.cu) to rounding; different data movement and fewer FP64 instructions.
//
// Why (profiles/r1_ncu_force_fused.md): k_force_fused requests 19 link matrices per link through L1 (10.9 KB/site), 6.0 KB/site of
// which come from L2 at the ~7 TB/s the L2->SM path delivers at 16 warps/SM; neither DRAM nor the FP64 pipe is the limit.
// Here a CTA owns a spatial tile of 8x4x2 sites and marches along t.  Every link matrix the six-staple stencil of a slice needs is
// copied into shared memory ONCE per tile and slice by TMA tensor copies (21 boxes per slice, one cp.async.bulk.tensor.4d each,
// issued by the lanes of warp 0, completion on mbarriers), one whole slice ahead of its use, so L2->SM traffic drops to 2.64
// matrix loads per link (1.5 KB/site) and every operand is a fixed-latency LDS.128.  (The first version copied with per-thread
// 16-byte cp.async: issuing 27 LDGSTS per thread and slice cost 24 % of the warp time -- profiles/r1_tmarch.md.)
// The backward-t staple is carried in registers from the previous slice by the thread that owns the link, so slice t-1 is never
// resident (tmarch_geom.h has the exact residency sets and the ring layout: 231 KB of shared memory, one CTA per SM).
//
// FP64 work: links are SU(3), so every staple A B C is formed from the first two rows of A only (2 x 72 FMA) and its third row
// is reconstructed as conj(row0 x row1) folded into the accumulation (24 FMA + 12 adds): 180 instead of 216 FP64 instructions per
// staple.  The staple sum itself is not unitary, so U V^dag, the TA projection, the exponential and exp*U stay full 3x3.
//
// One thread per (site, mu); a warp holds 32 sites of one direction (mu is a template parameter of the per-warp body, so every
// operand's ring and part are compile-time and its byte offset is one of 19 per-thread registers computed once per CTA).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gfb_internal.h"
#include "stencil.cuh"
#include "tmarch_geom.h"

#ifndef GFB_TM_PIPE
#define GFB_TM_PIPE 0  // 1: operands requested one product ahead in program order (tm_step)
#endif
#ifndef GFB_TM_PREGS
#define GFB_TM_PREGS 40  // registers per producer-group thread; the consumers get (64512 - 128*PREGS)/256 rounded down to 8
#endif
#ifndef GFB_TM_RSYNC_LATE
#define GFB_TM_RSYNC_LATE 1  // round barrier after (1) or before (0) the prologue copies of the next item
#endif
#ifndef GFB_TM_DEBUG
#define GFB_TM_DEBUG 0  // 1: no staple arithmetic (copies + operand reads only); 2: no global->shared copies (barriers only; arithmetic on stale smem); 3: no shared-memory reads (arithmetic on fabricated operands)
#endif

namespace gfb {

namespace {

struct TmPlan {
    int t_begin, t_count;  // local slices covered by the launch (contiguous)
    int seg_len, nseg;     // t-segments: item = (segment, tile)
    int ntx, nty, ntz, ntiles;
    // L2 blocking of the tile order: tiles are enumerated x fastest inside blocks of ntx x by x bz tiles (about one block = the
    // tiles the SMs march through concurrently), blocks y fastest.  The halo links of a tile then belong to tiles that march
    // at the same time (L2 hits); with the plain x,y,z order a wave of 148 tiles at 64^3 is one z-layer of tiles whose z-halos
    // were loaded a whole wave (hundreds of MB) earlier: 2783 B/site of DRAM traffic instead of 1664 (profiles/r2_tmarch.md)
    int by, bz;
};


__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// box table of tmarch_geom.h (tm::Tables::box) in constant memory: uniform loads feed the uniform-datapath TMA instructions
__constant__ int c_tm_box[tm::NBOX][8];

__device__ __forceinline__ bool elect_one() {
    unsigned p;
    asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}\n" : "=r"(p));
    return p != 0;
}

struct TmMaps {
    CUtensorMap m[tm::NMAP];  // per box shape and number of merged directions: tensor = [plane][z][y][2*x doubles], box = 9*nlam planes x ez x ey x 2*ex
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
#if GFB_TM_DEBUG == 2
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
#endif
}
#ifndef GFB_TM_WAIT_HINT
#define GFB_TM_WAIT_HINT 1  // 1: try_wait with a suspend-time hint (the hardware parks the warp instead of re-issuing the probe)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
#if GFB_TM_WAIT_HINT
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
#endif
}
// global -> shared 4-D tensor copy (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_4d(unsigned dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
#if GFB_TM_DEBUG != 2
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst_smem),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
#endif
}

// operand in shared memory: element k at p + k*stride (box-dense [k][pos] layout, tmarch_geom.h); p is a 32-bit shared address
struct SmOp {
    unsigned p;
    unsigned stride;
};
__device__ __forceinline__ double2 lds_el(const SmOp& o, int k) {
    double2 v;
#if GFB_TM_DEBUG == 3  // diagnostic: no shared-memory reads at all; operands fabricated in registers (one conversion + one add each)
    const unsigned a = o.p + k * o.stride;  // doubles in [0.25, 0.5) built with integer instructions only (no FP64-pipe work added)
    v.x = __hiloint2double(0x3FD00000 | (a & 0xFFFF), a);
    v.y = __hiloint2double(0x3FD00000 | ((a >> 4) & 0xFFFF), ~a);
#else
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(o.p + k * o.stride));
#endif
    return v;
}
__device__ __forceinline__ M3 lds_m3(const SmOp& o) {
    M3 r;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = lds_el(o, k);
    return r;
}
__device__ __forceinline__ R2 lds_rows01(const SmOp& o) {
    R2 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.e[k] = lds_el(o, k);
    return r;
}
// rows 0,1 of A^dagger: (A^dag)[i][j] = conj(A[j][i])
__device__ __forceinline__ R2 lds_dag_rows01(const SmOp& o) {
    R2 r;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double2 v = lds_el(o, 3 * j + i);
            r.e[3 * i + j] = make_double2(v.x, -v.y);
        }
    return r;
}
__device__ __forceinline__ int wrap(int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); }

// What a fused pass needs besides the lattice: field pointers, the coefficients of Z' = a TA(U V^dag) + b Z, U' = exp(c Z') U,
// the swizzle mask of the tile boxes (0x70 with CU_TENSOR_MAP_SWIZZLE_128B maps, 0 with linear ones) and -- on a t-slab
// decomposition -- the neighbours' copies of the OUTPUT link buffer: the boundary slices of U' are stored to the neighbours'
// halo slots by the threads that compute them (peer stores over NVLink inside the compute kernel; no pack, no send/recv kernel).
struct TmArgs {
    const double2* uin;
    double2* uout;
    const double* zin;
    double* zout;
    double a, b, c;
    unsigned swz;
    double2* peer_prev;  // previous slab's output buffer: our slice 0 (spatial links) goes to its slot tloc
    double2* peer_next;  // next slab's output buffer: our slice tloc-1 (all links) goes to its slot tloc+1
    // round barrier of the persistent grid (nullptr: off): every CTA's producer arrives once per round (= its k-th item) and
    // starts the copies of round k+1 only when all CTAs have arrived k+1 times; zeroed by the host before the launch
    unsigned long long* round_ctr;
};

// item -> (tile origin, t-segment)
struct TmItem {
    int x0, y0, z0, tb, len;
};
__device__ __forceinline__ TmItem decode_item(const TmPlan& pl, long item) {
    TmItem it;
    const int seg = (int)(item / pl.ntiles);
    const int r = (int)(item % pl.ntiles);
    int tx, ty, tz;
    tm::tile_of(pl.ntx, pl.nty, pl.ntz, pl.by, pl.bz, r, &tx, &ty, &tz);
    it.x0 = tx * tm::BX; it.y0 = ty * tm::BY; it.z0 = tz * tm::BZ;
    it.tb = pl.t_begin + seg * pl.seg_len;
    it.len = min(pl.seg_len, pl.t_begin + pl.t_count - it.tb);
    return it;
}

// descriptor (tmarch_geom.h, make_descriptors) -> shared-memory operand of the current step: sc = S slot of slice t,
// d_r / d_n = distance from there to the R slot of slice t / the S slot of slice t+1
__device__ __forceinline__ SmOp sm_operand(int d, unsigned sc, unsigned d_r, unsigned d_n, unsigned swz) {
    SmOp o;
    o.p = sc + ((unsigned)d & 0xFFFFu) + (((unsigned)d >> 28) & 1u) * d_r + (((unsigned)d >> 29) & 1u) * d_n;
    o.p ^= (o.p >> 3) & ((((unsigned)d >> 30) & 1u) * swz);  // 128-byte swizzle of the tile boxes (tmarch_geom.h, lookup())
    o.stride = (((unsigned)d >> 16) & 0xFFu) * 16u;
    return o;
}

// One slice step of one link-thread: the six staples from shared memory (the backward-t staple G is carried), the TA force,
// Z' and U' = exp(c Z') U.  `release` is called once every shared-memory operand of the step has been consumed.
template <bool READ_Z, bool WRITE_Z, bool DO_EXP, class Release>
__device__ __forceinline__ void tm_step(const int MU, const Geom& g, const TmArgs& ar, const int* od, const int t, const unsigned s3, const unsigned sc,
                                        const unsigned d_r, const unsigned d_n, R2& G, Release release) {
    auto at = [&](int d) -> SmOp { return sm_operand(d, sc, d_r, d_n, ar.swz); };
    const unsigned zo = (unsigned)(t * 32 + MU * 8) * (unsigned)g.v3 + s3;
    const unsigned zsb = (unsigned)g.v3 * 8u;
    double z[8];
    if (READ_Z) {
#pragma unroll
        for (int k = 0; k < 8; k++) z[k] = __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(ar.zin + zo) + (size_t)k * zsb));
    }

    M3 V, U;
    if (MU < 3) V = complete_su3(G);
    else V = m3_zero();
#if GFB_TM_DEBUG == 1
#pragma unroll
    for (int jj = 0; jj < 3; jj++) {
        m3_add(V, lds_m3(at(od[1 + 6 * jj]))); m3_add(V, lds_m3(at(od[2 + 6 * jj]))); m3_add(V, lds_m3(at(od[3 + 6 * jj])));
        if (jj < 2 || MU == 3) { m3_add(V, lds_m3(at(od[4 + 6 * jj]))); m3_add(V, lds_m3(at(od[5 + 6 * jj]))); m3_add(V, lds_m3(at(od[6 + 6 * jj]))); }
    }
    U = lds_m3(at(od[0]));
#elif GFB_TM_PIPE
    // Software pipeline in PROGRAM order (the loads are volatile asm, which ptxas keeps in order): every operand is requested one
    // whole 72-FMA product before its first use -- C of staple s ahead of T = A B, A and B of staple s+1 ahead of R = T C -- so
    // that with two warps per scheduler a warp does not sit on the short scoreboard a few FMAs after each LDS.  Live registers:
    // V 36 + T 24 + R 24 + C 36 + A' 24 + B' 36 + addressing ~45 = ~225 (the backward staple G is dead here, z comes later).
    {
        // staple s = 0..3: (upper, lower) of the two spatial-or-first directions; descriptors 1+6j.. (upper A,B,C), 4+6j.. (lower)
        R2 A = lds_rows01(at(od[1]));
        M3 B = lds_m3(at(od[2]));
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const int jj = s4 >> 1, lo = s4 & 1;
            const M3 C = lds_m3(at(od[(lo ? 6 : 3) + 6 * jj]));
            const R2 T = r2_mul_nn(A, B);
            M3 Bn;
            R2 An;
            M3 Afull;  // s4 == 3: operand 13 is needed in full by the nu = t tail of a spatial link
            if (s4 < 3) {
                const int s5 = s4 + 1, j2 = s5 >> 1, l2 = s5 & 1;
                An = l2 ? lds_dag_rows01(at(od[4 + 6 * j2])) : lds_rows01(at(od[1 + 6 * j2]));
                Bn = lds_m3(at(od[(l2 ? 5 : 2) + 6 * j2]));
            } else {
                Afull = lds_m3(at(od[13]));
                Bn = lds_m3(at(od[14]));
                An = rows01(Afull);
            }
            acc_su3(V, lo ? r2_mul_nn(T, C) : r2_mul_nd(T, C));
            if (s4 == 3) {
                if (MU == 3) {
                    // upper(2), lower(2) of a t-link: same pipeline, two more staples, then the own link
                    const M3 C2 = lds_m3(at(od[15]));
                    const R2 T2 = r2_mul_nn(An, Bn);
                    const R2 A3 = lds_dag_rows01(at(od[16]));
                    const M3 B3 = lds_m3(at(od[17]));
                    acc_su3(V, r2_mul_nd(T2, C2));
                    const M3 C3 = lds_m3(at(od[18]));
                    const R2 T3 = r2_mul_nn(A3, B3);
                    U = lds_m3(at(od[0]));
                    acc_su3(V, r2_mul_nn(T3, C3));
                } else {
                    // nu = t: the upper staple U_t(x) U_mu(x+t) U_t(x+mu)^dag and the NEXT slice's backward staple
                    // U_t(x)^dag U_mu(x) U_t(x+mu) share A = U_t(x) and C = U_t(x+mu); B of the latter is the own link
                    const M3 Ct = lds_m3(at(od[15]));
                    const R2 T2 = r2_mul_nn(An, Bn);
                    U = lds_m3(at(od[0]));
                    acc_su3(V, r2_mul_nd(T2, Ct));
                    const R2 T3 = r2_mul_nn(rows01_dag(Afull), U);
                    G = r2_mul_nn(T3, Ct);
                }
            }
            A = An;
            B = Bn;
        }
    }
#else
    auto upper = [&](int jj) {  // A B C^dag
        const R2 A = lds_rows01(at(od[1 + 6 * jj]));
        const M3 B = lds_m3(at(od[2 + 6 * jj]));
        const R2 T = r2_mul_nn(A, B);
        const M3 C = lds_m3(at(od[3 + 6 * jj]));
        acc_su3(V, r2_mul_nd(T, C));
    };
    auto lower = [&](int jj) {  // A^dag B C
        const R2 A = lds_dag_rows01(at(od[4 + 6 * jj]));
        const M3 B = lds_m3(at(od[5 + 6 * jj]));
        const R2 T = r2_mul_nn(A, B);
        const M3 C = lds_m3(at(od[6 + 6 * jj]));
        acc_su3(V, r2_mul_nn(T, C));
    };
    upper(0); lower(0);
    upper(1); lower(1);
    if (MU == 3) {
        upper(2); lower(2);
        U = lds_m3(at(od[0]));
    } else {
        // nu = t: the upper staple U_t(x) U_mu(x+t) U_t(x+mu)^dag and the NEXT slice's backward staple
        // U_t(x)^dag U_mu(x) U_t(x+mu) share A = U_t(x) and C = U_t(x+mu); B of the latter is the own link
        const M3 A = lds_m3(at(od[13]));
        const M3 C = lds_m3(at(od[15]));
        {
            const M3 B = lds_m3(at(od[14]));
            const R2 T = r2_mul_nn(rows01(A), B);
            acc_su3(V, r2_mul_nd(T, C));
        }
        U = lds_m3(at(od[0]));
        const R2 T = r2_mul_nn(rows01_dag(A), U);
        G = r2_mul_nn(T, C);
    }
#endif

    double f[8];
    ta_coeffs_nd(U, V, f);
    release();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double v = ar.a * f[k];
        if (READ_Z) v = fma(ar.b, z[k], v);
        f[k] = v;
        if (WRITE_Z) *reinterpret_cast<double*>(reinterpret_cast<char*>(ar.zout + zo) + (size_t)k * zsb) = v;
    }
    if (DO_EXP) {
        const M3 r = exp_ta_times_su3(f, ar.c, U);
        const unsigned uo = (unsigned)(t * 36 + MU * 9) * (unsigned)g.v3 + s3;
        m3_store(ar.uout + uo, (unsigned)g.v3, r);
        // t-slab halo: the neighbours' halo slots are filled from here (slot tloc = their t+1 halo needs our slice 0's three
        // spatial links, slot tloc+1 = their t-1 halo needs our last slice's four links; SURVEY 8e, set_wing_U! in the reference)
        if (ar.peer_prev != nullptr && t == 0 && MU < 3)
            m3_store(ar.peer_prev + ((unsigned)(g.tloc * 36 + MU * 9) * (unsigned)g.v3 + s3), (unsigned)g.v3, r);
        if (ar.peer_next != nullptr && t == g.tloc - 1)
            m3_store(ar.peer_next + ((unsigned)((g.tloc + 1) * 36 + MU * 9) * (unsigned)g.v3 + s3), (unsigned)g.v3, r);
    }
}

// backward-t staple of the first slice of a segment, from global memory (afterwards it is carried in registers)
__device__ __forceinline__ R2 first_backward_staple(const int MU, const Geom& g, const double2* __restrict__ uin, const Coord& x) {
    const Coord y = step(g, x, 3, -1);
    const Coord ym = step(g, y, MU, +1);
    const M3 A = load_link(uin, g, y, 3);
    const M3 U = load_link(uin, g, y, MU);
    const M3 C = load_link(uin, g, ym, 3);
    return r2_mul_nn(r2_mul_nn(rows01_dag(A), U), C);
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant A (GFB200_TMARCH_WS=0): 8 warps, warp 0 also issues the copies, one CTA-wide barrier per slice step.
// mu is warp-uniform but NOT a template parameter: all eight warps run the same instructions (a per-direction instantiation
// made the straight-line staple code four times larger than the instruction cache could hold, profiles/r1_tmarch.md).
// ---------------------------------------------------------------------------------------------------------------------
template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__device__ __forceinline__ void tm_run(const int MU, const TmMaps& maps, const Geom& g, const TmPlan& pl, const TmArgs& ar, unsigned char* smem,
                                       uint64_t* bars, const tm::Tables* __restrict__ tab) {
    const int tid = threadIdx.x;
    const int sidx = tid & (tm::SITES - 1);
    const int sx = sidx & (tm::BX - 1), sy = (sidx / tm::BX) & (tm::BY - 1), sz = sidx / (tm::BX * tm::BY);
    unsigned char* const sS = smem;
    unsigned char* const sR = smem + tm::S_RING * tm::S_SLOT;

    // ---- consumer side: operand descriptors (tile independent; tm::make_descriptors, tabulated once on the host)
    int od[tm::NDESC];
#pragma unroll
    for (int i = 0; i < tm::NDESC; i++) od[i] = tab->desc[i][tid];

    // ---- producer side: lane b of warp 0 owns box b (21 boxes per slice)
    const bool is_producer = tid < tm::NBOX;
    int bx_o[3] = {0, 0, 0}, b_lam = 0, b_isr = 0, b_base = 0, b_shape = 0;
    if (is_producer) {
        bx_o[0] = tab->box[tid][0]; bx_o[1] = tab->box[tid][1]; bx_o[2] = tab->box[tid][2];
        b_lam = tab->box[tid][3]; b_isr = tab->box[tid][4]; b_base = tab->box[tid][5]; b_shape = tab->box[tid][6];
    }
    uint64_t* const barS = bars;                 // [S_RING]
    uint64_t* const barR = bars + tm::S_RING;    // [R_RING]
    unsigned phases = 0;                         // bit s: parity of the next completion of barrier s (uniform over the CTA)
    auto wait_bar = [&](int s) {
        mbar_wait(bars + s, (phases >> s) & 1u);
        phases ^= 1u << s;
    };

    const long nitems = (long)pl.ntiles * pl.nseg;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const TmItem it = decode_item(pl, item);
        const int x0 = it.x0, y0 = it.y0, z0 = it.z0, tb = it.tb, len = it.len;
        const int cx = 2 * wrap(x0 + bx_o[0], g.nx), cy = wrap(y0 + bx_o[1], g.ny), cz = wrap(z0 + bx_o[2], g.nz);
        // one part (all its boxes) of the slice in storage slot `tslot` into ring buffer `ring`; warp 0 only
        auto copy_part = [&](int is_r, int tslot, int ring) {
            if (tid >= 32) return;
            uint64_t* const bar = is_r ? barR + ring : barS + ring;
            if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)((is_r ? tm::R_MATS : tm::S_MATS) * tm::MAT_BYTES));
            __syncwarp();
            if (is_producer && b_isr == is_r) {
                unsigned char* const dst = (is_r ? sR + ring * tm::R_SLOT : sS + ring * tm::S_SLOT) + b_base;
                tma_load_4d(smem_u32(dst), &maps.m[b_shape], cx, cy, cz, tslot * 36 + b_lam * 9, bar);
            }
        };
        auto t_up = [&](int t) { return (t == g.tloc - 1) ? g.t_up_wrap : t + 1; };

        // ---- prologue: slice tb (full) and the S part of slice tb+1; the backward-t staple of slice tb from global memory
        copy_part(0, tb, 0);
        copy_part(1, tb, 0);
        copy_part(0, t_up(tb), 1);

        Coord x;
        x.x = x0 + sx; x.y = y0 + sy; x.z = z0 + sz; x.t = tb;
        const unsigned s3 = (unsigned)s3_of(g, x);
        R2 G;
        if (MU < 3) G = first_backward_staple(MU, g, ar.uin, x);
        wait_bar(0);
        wait_bar(tm::S_RING + 0);
        wait_bar(1);

        int rs = 0;  // j % 3
        for (int j = 0; j < len; j++) {
            const int t = tb + j;
            const int rs1 = (rs == 2) ? 0 : rs + 1;   // (j+1) % 3
            const int rs2 = (rs1 == 2) ? 0 : rs1 + 1; // (j+2) % 3
            const bool more = j + 1 < len;
            if (more) {
                copy_part(1, t + 1, (j + 1) & 1);
                copy_part(0, t_up(t + 1), rs2);
            }
            // branch-free operand addressing (a ternary on the three ring bases compiled to divergent-branch regions that
            // ptxas could not schedule loads across): 32-bit shared address = Sc + offset + isR*(Rc-Sc) + isNext*(Sn-Sc)
            const unsigned sc = smem_u32(sS + rs * tm::S_SLOT);
            const unsigned d_r = smem_u32(sR + (j & 1) * tm::R_SLOT) - sc;
            const unsigned d_n = smem_u32(sS + rs1 * tm::S_SLOT) - sc;
            tm_step<READ_Z, WRITE_Z, DO_EXP>(MU, g, ar, od, t, s3, sc, d_r, d_n, G, [] {});
            // the next slice's parts (requested at the top of this step) must have landed; then every thread is done
            // reading this step's buffers and warp 0 may overwrite them
            if (more) {
                wait_bar(tm::S_RING + ((j + 1) & 1));
                wait_bar(rs2);
            }
            __syncthreads();
            rs = rs1;
        }
    }
}

template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__global__ void __launch_bounds__(tm::NTHREADS, 1)
k_tmarch_fused(const __grid_constant__ TmMaps maps, const tm::Tables* __restrict__ tab, Geom g, TmPlan pl, TmArgs ar) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + tm::BAR_OFF);
    if (threadIdx.x == 0) {
        for (int i = 0; i < tm::S_RING + tm::R_RING; i++) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int mu = threadIdx.x / tm::SITES;  // warp-uniform: two warps per direction
    tm_run<READ_Z, WRITE_Z, DO_EXP>(mu, maps, g, pl, ar, smem, bars, tab);
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant B (default): warp-specialised.  Warps 0-7 are the 256 link-threads (consumers), warps 8-11 a producer warpgroup whose
// GFB_TM_NPW warps issue the tensor copies (boxes dealt round-robin).  Registers move from the producer group to the consumers
// (setmaxnreg: 40 / 232 per thread, of the 64512 registers of a 384 x 168 launch).  No CTA-wide barrier in the march: every ring slot
// has a FULL mbarrier (the copies' byte count) and an EMPTY mbarrier (one arrival per consumer warp, given as soon as the
// warp has consumed the slot's operands, i.e. BEFORE its exponential), so the eight consumer warps drift apart by up to a
// slice instead of meeting once per step behind the warp that also had to issue 21 copies (variant A: 8 % barrier stall,
// 2100 of a step's 11250 cycles spent by warp 0 on UTMALDG issue -- profiles/r1_tmarch.md).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWsThreads = tm::NTHREADS + 128;
constexpr int kNBars = tm::S_RING + tm::R_RING;  // FULL barriers [0, kNBars), EMPTY barriers [kNBars, 2 kNBars)

template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    // relaxed, not acquire: the counter orders nothing (a locality hint), and an acquire load at gpu scope invalidates the SM's L1
    // on every poll (CCTL.IVALL), which is where the link warps' spilled descriptors live
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef GFB_TM_NPW
#define GFB_TM_NPW 4  // producer warps that issue tensor copies (of the 4 in the producer warpgroup); boxes are dealt round-robin
#endif
__device__ __noinline__ void tm_producer(const TmMaps& maps, const Geom& g, const TmPlan& pl, unsigned char* smem, uint64_t* bars,
                                         unsigned long long* round_ctr_arg, const int pw) {
    // producer warp 0 owns the barrier bookkeeping (the byte count of a fill, the round counter); the others only add copies
    unsigned long long* const round_ctr = pw == 0 ? round_ctr_arg : nullptr;
    unsigned char* const sS = smem;
    unsigned char* const sR = smem + tm::S_RING * tm::S_SLOT;
    unsigned ephase = ~0u;  // EMPTY barriers start "released": the first wait on each passes (parity of the preceding phase)
    const bool leader = elect_one();
    const long nitems = (long)pl.ntiles * pl.nseg;
    // Round barrier (persistent grid, one CTA per SM, all co-resident).  Without it nothing keeps the 148 marches aligned: a
    // CTA that is 1 % faster is 17 slices ahead after the 27 rounds x 64 slices of a 64^4 pass, its neighbours' halo links
    // have left L2 by the time it asks for them, and they are read from DRAM again (2644 instead of 1664 B/site,
    // profiles/r2_tmarch.md).  With it all CTAs start a round together and drift only within one march.
    long round = 0;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x, round++) {
        const TmItem it = decode_item(pl, item);
        auto round_wait = [&]() {
            if (round_ctr != nullptr && round > 0) {
                if (leader) {
                    // The barrier is a locality hint, not a correctness condition, so it gives up after 2 ms: if the CTAs of this
                    // launch are NOT all co-resident (another context's persistent kernel holds SMs) it must not deadlock.
                    const unsigned long long want = (unsigned long long)round * gridDim.x;
                    const unsigned long long t0 = global_timer_ns();
                    while (ld_acquire_gpu(round_ctr) < want) {
                        __nanosleep(200);
                        if (global_timer_ns() - t0 > 2000000ull) break;
                    }
                }
                __syncwarp();
            }
        };
#if !GFB_TM_RSYNC_LATE
        round_wait();
#endif
        // every box of one part of the slice in storage slot `tslot` into ring slot `ring`, once the consumers released it
        auto fill = [&](int is_r, int tslot, int ring) {
            const int s = is_r ? tm::S_RING + ring : ring;
            mbar_wait(bars + kNBars + s, (ephase >> s) & 1u);
            ephase ^= 1u << s;
            if (leader) {
                uint64_t* const bar = bars + s;
                unsigned char* const part = is_r ? sR + ring * tm::R_SLOT : sS + ring * tm::S_SLOT;
                // one arrival + the byte count of ALL boxes of the part; copies of the other producer warps may complete before
                // this is performed (the transaction count goes negative, the phase cannot complete without the arrival)
                if (pw == 0) mbar_arrive_expect_tx(bar, (unsigned)((is_r ? tm::R_MATS : tm::S_MATS) * tm::MAT_BYTES));
                const int b0 = is_r ? tm::NBOX_S : 0, b1 = is_r ? tm::NBOX : tm::NBOX_S;
                for (int b = b0 + pw; b < b1; b += GFB_TM_NPW) {
                    const int bcx = 2 * wrap(it.x0 + c_tm_box[b][0], g.nx), bcy = wrap(it.y0 + c_tm_box[b][1], g.ny), bcz = wrap(it.z0 + c_tm_box[b][2], g.nz);
                    tma_load_4d(smem_u32(part + c_tm_box[b][5]), &maps.m[c_tm_box[b][6]], bcx, bcy, bcz, tslot * 36 + c_tm_box[b][3] * 9, bar);
                }
            }
            __syncwarp();
        };
        auto t_up = [&](int t) { return (t == g.tloc - 1) ? g.t_up_wrap : t + 1; };
        fill(0, it.tb, 0);
        fill(1, it.tb, 0);
        fill(0, t_up(it.tb), 1);
#if GFB_TM_RSYNC_LATE
        // the barrier gates the march, not its prologue: the first slices of the next tile are requested while the slower CTAs
        // finish their round, so a CTA resumes one slice after the barrier opens instead of one copy latency + one slice
        round_wait();
#endif
        int rs = 0;
        for (int j = 0; j + 1 < it.len; j++) {
            const int t = it.tb + j;
            const int rs1 = (rs == 2) ? 0 : rs + 1, rs2 = (rs1 == 2) ? 0 : rs1 + 1;
            fill(1, t + 1, (j + 1) & 1);
            fill(0, t_up(t + 1), rs2);
            rs = rs1;
        }
        // all copies of this item are issued (the link warps are at most one slice behind)
        if (round_ctr != nullptr && leader) atomicAdd(round_ctr, 1ull);
    }
    // CTAs that have no item in the last round still arrive for it
    if (round_ctr != nullptr && leader && round * (long)gridDim.x < nitems) atomicAdd(round_ctr, 1ull);
}

template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__device__ __forceinline__ void tm_consumer(const int MU, const Geom& g, const TmPlan& pl, const TmArgs& ar, unsigned char* smem, uint64_t* bars,
                                            const tm::Tables* __restrict__ tab) {
    const int tid = threadIdx.x;
    const int sidx = tid & (tm::SITES - 1);
    const int sx = sidx & (tm::BX - 1), sy = (sidx / tm::BX) & (tm::BY - 1), sz = sidx / (tm::BX * tm::BY);
    unsigned char* const sS = smem;
    unsigned char* const sR = smem + tm::S_RING * tm::S_SLOT;
    int od[tm::NDESC];
#pragma unroll
    for (int i = 0; i < tm::NDESC; i++) od[i] = tab->desc[i][tid];
    unsigned fphase = 0;  // bit s: parity of the next completion of FULL barrier s
    auto wait_full = [&](int s) {
        mbar_wait(bars + s, (fphase >> s) & 1u);
        fphase ^= 1u << s;
    };
    const bool lane0 = (tid & 31) == 0;
    const long nitems = (long)pl.ntiles * pl.nseg;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const TmItem it = decode_item(pl, item);
        Coord x;
        x.x = it.x0 + sx; x.y = it.y0 + sy; x.z = it.z0 + sz; x.t = it.tb;
        const unsigned s3 = (unsigned)s3_of(g, x);
        R2 G;
        if (MU < 3) G = first_backward_staple(MU, g, ar.uin, x);
        wait_full(0);               // S part of slice tb
        wait_full(tm::S_RING + 0);  // R part of slice tb
        wait_full(1);               // S part of slice tb+1
        int rs = 0;
        for (int j = 0; j < it.len; j++) {
            const int t = it.tb + j;
            const int rs1 = (rs == 2) ? 0 : rs + 1;
            if (j > 0) {  // filled while the previous step ran
                wait_full(tm::S_RING + (j & 1));
                wait_full(rs1);
            }
            const unsigned sc = smem_u32(sS + rs * tm::S_SLOT);
            const unsigned d_r = smem_u32(sR + (j & 1) * tm::R_SLOT) - sc;
            const unsigned d_n = smem_u32(sS + rs1 * tm::S_SLOT) - sc;
            const bool last = j + 1 == it.len;
            tm_step<READ_Z, WRITE_Z, DO_EXP>(MU, g, ar, od, t, s3, sc, d_r, d_n, G, [&] {
                // this warp is done with slice t's S and R parts (the S part of slice t+1 stays for the next step, unless this
                // was the last step of the segment)
                __syncwarp();
                if (lane0) {
                    mbar_arrive(bars + kNBars + rs);
                    mbar_arrive(bars + kNBars + tm::S_RING + (j & 1));
                    if (last) mbar_arrive(bars + kNBars + rs1);
                }
            });
            rs = rs1;
        }
    }
}

template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__global__ void __launch_bounds__(kWsThreads, 1)
k_tmarch_ws(const __grid_constant__ TmMaps maps, const tm::Tables* __restrict__ tab, Geom g, TmPlan pl, TmArgs ar) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + tm::BAR_OFF);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNBars; i++) mbar_init(bars + i, 1);
        for (int i = 0; i < kNBars; i++) mbar_init(bars + kNBars + i, tm::NTHREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= tm::NTHREADS) {
        reg_dec<GFB_TM_PREGS>();
        if (threadIdx.x < tm::NTHREADS + 32 * GFB_TM_NPW) tm_producer(maps, g, pl, smem, bars, ar.round_ctr, (threadIdx.x - tm::NTHREADS) >> 5);
        return;
    }
    reg_inc<((64512 - 128 * GFB_TM_PREGS) / 256) & ~7>();
    const int mu = threadIdx.x / tm::SITES;  // warp-uniform: two warps per direction
    tm_consumer<READ_Z, WRITE_Z, DO_EXP>(mu, g, pl, ar, smem, bars, tab);
}

constexpr size_t kTmSmem = tm::SMEM_DATA;  // 232448 = the opt-in maximum; the mbarriers live in the tail of R slot 0

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
    // block of concurrently marched tiles: ntx x by x bz ~ nsm tiles, about square in sites (BY*by ~ BZ*bz)
    const double per_x = (double)nsm / pl.ntx;
    int by = (int)(std::sqrt(per_x * tm::BZ / tm::BY) + 0.5);
    by = by < 1 ? 1 : (by > pl.nty ? pl.nty : by);
    int bz = (int)(per_x / by + 0.5);
    bz = bz < 1 ? 1 : (bz > pl.ntz ? pl.ntz : bz);
    pl.by = by; pl.bz = bz;
    return pl;
}

// geometry tables in device memory, one copy per device
const tm::Tables* device_tables(int dev) {
    static std::mutex mtx;
    static const tm::Tables* per_dev[64] = {};
    std::lock_guard<std::mutex> lock(mtx);
    if (per_dev[dev & 63]) return per_dev[dev & 63];
    static tm::Tables host;
    static bool built = false;
    if (!built) { tm::make_tables(&host); built = true; }
    tm::Tables* d = nullptr;
    if (cudaMalloc(&d, sizeof(tm::Tables)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(d, &host, sizeof(tm::Tables), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
    if (cudaMemcpyToSymbol(c_tm_box, host.box, sizeof(host.box)) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
    per_dev[dev & 63] = d;
    return d;
}

// round-barrier counter of a persistent launch, one per (device, stream): launches on one stream are serialised
unsigned long long* round_counter_for(int dev, cudaStream_t st) {
    static std::mutex mtx;
    static std::unordered_map<unsigned long long, unsigned long long*> per;
    std::lock_guard<std::mutex> lock(mtx);
    const unsigned long long key = ((unsigned long long)(uintptr_t)st << 6) ^ (unsigned long long)(dev & 63);
    auto it = per.find(key);
    if (it != per.end()) return it->second;
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    per.emplace(key, d);
    return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor maps of a link buffer viewed as [nslots*36 planes][nz][ny][2*nx doubles], one per box shape; cached per buffer
// Returned BY VALUE (copied under the lock): the cache may be cleared by another thread's call.
bool tensor_maps_for(const double2* u, const Geom& g, bool swizzle, TmMaps* out) {
    struct Key {
        const void* p;
        int nx, ny, nz, nslots, swz;
        bool operator==(const Key& o) const { return p == o.p && nx == o.nx && ny == o.ny && nz == o.nz && nslots == o.nslots && swz == o.swz; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            return std::hash<const void*>()(k.p) ^ ((size_t)k.nx * 1315423911u) ^ ((size_t)k.ny << 12) ^ ((size_t)k.nz << 24) ^ ((size_t)k.nslots << 36) ^ (size_t)k.swz;
        }
    };
    static std::unordered_map<Key, TmMaps, Hash> cache;
    static std::mutex mtx;
    static EncodeTiledFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mtx);
    const Key key{u, g.nx, g.ny, g.nz, g.nslots, swizzle ? 1 : 0};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    TmMaps maps;
    const cuuint64_t gdim[4] = {(cuuint64_t)g.nx * 2, (cuuint64_t)g.ny, (cuuint64_t)g.nz, (cuuint64_t)g.nslots * 36};
    const cuuint64_t gstride[3] = {(cuuint64_t)g.nx * 16, (cuuint64_t)g.nx * g.ny * 16, (cuuint64_t)g.v3 * 16};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int s = 0; s < tm::NMAP; s++) {
        int e[3];
        tm::shape_extent(s / 3, e);
        const cuuint32_t box[4] = {(cuuint32_t)e[0] * 2, (cuuint32_t)e[1], (cuuint32_t)e[2], (cuuint32_t)(9 * (s % 3 + 1))};
        // full-tile boxes (shape 0: rows of 8 x-sites = 128 bytes) are stored with the 128-byte swizzle (tmarch_geom.h, lookup())
        const CUtensorMapSwizzle sw = (swizzle && s / 3 == 0) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        if (encode(&maps.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double2*>(u), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    if (cache.size() > 1024) cache.clear();  // buffers come and go with the fields; the maps are cheap to rebuild
    cache.emplace(key, maps);
    *out = maps;
    return true;
}

}  // namespace

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Returns false when the launch is not covered (tile does not divide the lattice, strided slice set, in-place links):
// the caller then uses k_force_fused.  Environment (read per launch so that tests can compare variants in one process):
//   GFB200_TMARCH=0 disables the kernel;  GFB200_TMARCH_WS=0 selects variant A (no producer warpgroup);
//   GFB200_TMARCH_SWIZZLE=0 copies the tile boxes linearly;  GFB200_TMARCH_SEGLEN=n forces the t-segment length.
bool launch_tmarch_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                         const FusedArgs& fa) {
    if (!env_int("GFB200_TMARCH", 1)) return false;
    if (g.t_stride != 1 || t_count < 1) return false;
    if (g.nx % tm::BX || g.ny % tm::BY || g.nz % tm::BZ) return false;
    if (uout == uin) return false;
    int dev = 0;
    cudaGetDevice(&dev);
    static int nsm = 0;
    if (nsm == 0) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const bool ws = env_int("GFB200_TMARCH_WS", 1) != 0;
    const bool swizzle = env_int("GFB200_TMARCH_SWIZZLE", 1) != 0;
    TmPlan pl = make_plan(g, t_begin, t_count, nsm);
    {
        const int len = env_int("GFB200_TMARCH_SEGLEN", 0);  // test hook
        if (len >= 1) { pl.seg_len = len < t_count ? len : t_count; pl.nseg = (t_count + pl.seg_len - 1) / pl.seg_len; }
    }
    {
        const int by = env_int("GFB200_TMARCH_BY", 0), bz = env_int("GFB200_TMARCH_BZ", 0);  // tuning hooks: tile-block shape
        if (by >= 1) pl.by = by < pl.nty ? by : pl.nty;
        if (bz >= 1) pl.bz = bz < pl.ntz ? bz : pl.ntz;
    }
    const long nitems = (long)pl.ntiles * pl.nseg;
    // persistent: one CTA per SM (232448 bytes of shared memory each).  NCCL-overlapped slab interiors (fa.leave_sms) launch one
    // CTA per item instead, so that SMs free up for the send/recv kernels of the halo stream (profiles/r1_tmarch.md)
    const unsigned grid = (unsigned)((nitems < nsm || fa.leave_sms) ? nitems : nsm);
    TmMaps maps;
    const tm::Tables* tab = device_tables(dev);
    if (!tab || !tensor_maps_for(uin, g, swizzle, &maps)) return false;
    TmArgs ar;
    ar.uin = uin; ar.uout = uout; ar.zin = zin; ar.zout = zout;
    ar.a = fa.a; ar.b = fa.b; ar.c = fa.c;
    ar.swz = swizzle ? 0x70u : 0u;
    ar.peer_prev = fa.do_exp ? fa.peer_prev : nullptr;
    ar.peer_next = fa.do_exp ? fa.peer_next : nullptr;
    // round barrier: persistent launches of the warp-specialised kernel with at least two full rounds of long marches.  Measured
    // (profiles/r2_tmarch.md): 64^4 (28 rounds x 64 slices) DRAM read 1880 -> 1097 B/site, 10.29 -> 9.83 ms; 32^4 (7 rounds x 16
    // slices, no re-read problem) 0.619 -> 0.625 ms, hence the length threshold.  GFB200_TMARCH_ROUNDSYNC=0 / 2: off / always.
    ar.round_ctr = nullptr;
    const int rsync = env_int("GFB200_TMARCH_ROUNDSYNC", 1);
    if (ws && grid == (unsigned)nsm && nitems >= 2L * nsm && rsync && (pl.seg_len >= 32 || rsync == 2)) {
        ar.round_ctr = round_counter_for(dev, st);
        if (ar.round_ctr && cudaMemsetAsync(ar.round_ctr, 0, sizeof(unsigned long long), st) != cudaSuccess) { cudaGetLastError(); ar.round_ctr = nullptr; }
    }
    if (env_int("GFB200_PEER_NOSTORE", 0)) ar.peer_prev = ar.peer_next = nullptr;  // timing experiments only: halos stay stale
#define GFB_LAUNCH_TM(R, W, E)                                                                                                 \
    do {                                                                                                                        \
        static bool attr_set[2][64] = {};  /* per device: one process may drive several GPUs */                                 \
        if (ws) {                                                                                                               \
            auto kern = k_tmarch_ws<R, W, E>;                                                                                   \
            if (!attr_set[1][dev & 63]) {                                                                                       \
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmSmem) != cudaSuccess) {     \
                    cudaGetLastError();                                                                                         \
                    return false;                                                                                               \
                }                                                                                                               \
                attr_set[1][dev & 63] = true;                                                                                   \
            }                                                                                                                   \
            kern<<<grid, kWsThreads, kTmSmem, st>>>(maps, tab, g, pl, ar);                                                      \
        } else {                                                                                                                \
            auto kern = k_tmarch_fused<R, W, E>;                                                                                \
            if (!attr_set[0][dev & 63]) {                                                                                       \
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmSmem) != cudaSuccess) {     \
                    cudaGetLastError();                                                                                         \
                    return false;                                                                                               \
                }                                                                                                               \
                attr_set[0][dev & 63] = true;                                                                                   \
            }                                                                                                                   \
            kern<<<grid, tm::NTHREADS, kTmSmem, st>>>(maps, tab, g, pl, ar);                                                    \
        }                                                                                                                       \
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
