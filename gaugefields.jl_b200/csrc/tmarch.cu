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
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gfb_internal.h"
#include "stencil.cuh"
#include "tmarch_geom.h"

#ifndef GFB_TM_UNIFORM_ISSUE
#define GFB_TM_UNIFORM_ISSUE 0  // 1: one elected thread issues all tensor copies back to back from uniform operands (box table in constant memory); measured 32^4 0.654 ms vs 0.620, 64^4 10.33 vs 10.38: not adopted
#endif
#ifndef GFB_TM_DEBUG
#define GFB_TM_DEBUG 0  // 1: no staple arithmetic (copies + operand reads only); 2: no global->shared copies (barriers only; arithmetic on stale smem)
#endif

namespace gfb {

namespace {

struct TmPlan {
    int t_begin, t_count;  // local slices covered by the launch (contiguous)
    int seg_len, nseg;     // t-segments: item = (segment, tile)
    int ntx, nty, ntz, ntiles;
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
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
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(o.p + k * o.stride));
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

// The whole persistent loop of one link-thread.  mu is warp-uniform but NOT a template parameter: all eight warps run the same
// instructions (a per-direction instantiation made the straight-line staple code four times larger than the instruction
// cache could hold: "no instruction" was the top stall of the first version, profiles/r1_tmarch.md).
template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__device__ __forceinline__ void tm_run(const int MU, const TmMaps& maps, const Geom& g, const TmPlan& pl, const double2* __restrict__ uin,
                                       double2* __restrict__ uout, const double* __restrict__ zin, double* __restrict__ zout, double a, double b,
                                       double c, unsigned char* smem, uint64_t* bars, const tm::Tables* __restrict__ tab) {
    const int tid = threadIdx.x;
    const int sidx = tid & (tm::SITES - 1);
    const int sx = sidx & (tm::BX - 1), sy = (sidx / tm::BX) & (tm::BY - 1), sz = sidx / (tm::BX * tm::BY);
    unsigned char* const sS = smem;
    unsigned char* const sR = smem + tm::S_RING * tm::S_BYTES;

    // ---- consumer side: operand descriptors (tile independent; tm::make_descriptors, tabulated once on the host)
    int od[tm::NDESC];
#pragma unroll
    for (int i = 0; i < tm::NDESC; i++) od[i] = tab->desc[i][tid];

#if !GFB_TM_UNIFORM_ISSUE
    // ---- producer side: lane b of warp 0 owns box b (21 boxes per slice)
    const bool is_producer = tid < tm::NBOX;
    int bx_o[3] = {0, 0, 0}, b_lam = 0, b_isr = 0, b_base = 0, b_shape = 0;
    if (is_producer) {
        bx_o[0] = tab->box[tid][0]; bx_o[1] = tab->box[tid][1]; bx_o[2] = tab->box[tid][2];
        b_lam = tab->box[tid][3]; b_isr = tab->box[tid][4]; b_base = tab->box[tid][5]; b_shape = tab->box[tid][6];
    }
#endif
    uint64_t* const barS = bars;                 // [S_RING]
    uint64_t* const barR = bars + tm::S_RING;    // [R_RING]
    unsigned phases = 0;                         // bit s: parity of the next completion of barrier s (uniform over the CTA)
    auto wait_bar = [&](int s) {
        mbar_wait(bars + s, (phases >> s) & 1u);
        phases ^= 1u << s;
    };

    const long nitems = (long)pl.ntiles * pl.nseg;
    for (long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int seg = (int)(item / pl.ntiles);
        int tile = (int)(item % pl.ntiles);
        const int x0 = (tile % pl.ntx) * tm::BX; tile /= pl.ntx;
        const int y0 = (tile % pl.nty) * tm::BY;
        const int z0 = (tile / pl.nty) * tm::BZ;
        const int tb = pl.t_begin + seg * pl.seg_len;
        const int len = min(pl.seg_len, pl.t_begin + pl.t_count - tb);

#if GFB_TM_UNIFORM_ISSUE
        // one part (all its boxes) of the slice in storage slot `tslot` into ring buffer `ring`: one elected thread of warp 0
        // issues the copies back to back; every operand is warp-uniform (tile origin, box table from constant memory), so the
        // compiler feeds the uniform-datapath UTMALDG without the per-lane ELECT / 8 x R2UR loop of a lane-per-box issue
        auto copy_part = [&](int is_r, int tslot, int ring) {
            if ((tid >> 5) != 0) return;
            uint64_t* const bar = is_r ? barR + ring : barS + ring;
            unsigned char* const part = is_r ? sR + ring * tm::R_BYTES : sS + ring * tm::S_BYTES;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar, (unsigned)((is_r ? tm::R_MATS : tm::S_MATS) * tm::MAT_BYTES));
                const int b0 = is_r ? tm::NBOX_S : 0, b1 = is_r ? tm::NBOX : tm::NBOX_S;
#pragma unroll
                for (int b = b0; b < b1; b++) {
                    const int bcx = 2 * wrap(x0 + c_tm_box[b][0], g.nx), bcy = wrap(y0 + c_tm_box[b][1], g.ny), bcz = wrap(z0 + c_tm_box[b][2], g.nz);
                    tma_load_4d(smem_u32(part + c_tm_box[b][5]), &maps.m[c_tm_box[b][6]], bcx, bcy, bcz, tslot * 36 + c_tm_box[b][3] * 9, bar);
                }
            }
            __syncwarp();
        };
#else
        const int cx = 2 * wrap(x0 + bx_o[0], g.nx), cy = wrap(y0 + bx_o[1], g.ny), cz = wrap(z0 + bx_o[2], g.nz);
        // one part (all its boxes) of the slice in storage slot `tslot` into ring buffer `ring`; warp 0 only
        auto copy_part = [&](int is_r, int tslot, int ring) {
            if (tid >= 32) return;
            uint64_t* const bar = is_r ? barR + ring : barS + ring;
            if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)((is_r ? tm::R_MATS : tm::S_MATS) * tm::MAT_BYTES));
            __syncwarp();
            if (is_producer && b_isr == is_r) {
                unsigned char* const dst = (is_r ? sR + ring * tm::R_BYTES : sS + ring * tm::S_BYTES) + b_base;
                tma_load_4d(smem_u32(dst), &maps.m[b_shape], cx, cy, cz, tslot * 36 + b_lam * 9, bar);
            }
        };
#endif
        auto t_up = [&](int t) { return (t == g.tloc - 1) ? g.t_up_wrap : t + 1; };

        // ---- prologue: slice tb (full) and the S part of slice tb+1; the backward-t staple of slice tb from global memory
        copy_part(0, tb, 0);
        copy_part(1, tb, 0);
        copy_part(0, t_up(tb), 1);

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
            const unsigned sc = smem_u32(sS + rs * tm::S_BYTES);
            const unsigned d_r = smem_u32(sR + (j & 1) * tm::R_BYTES) - sc;
            const unsigned d_n = smem_u32(sS + rs1 * tm::S_BYTES) - sc;
            auto at = [&](int d) -> SmOp {
                SmOp o;
                o.p = sc + ((unsigned)d & 0xFFFFu) + (((unsigned)d >> 28) & 1u) * d_r + (((unsigned)d >> 29) & 1u) * d_n;
                o.stride = (((unsigned)d >> 16) & 0xFFu) * 16u;
                return o;
            };

            const unsigned zo = (unsigned)(t * 32 + MU * 8) * (unsigned)g.v3 + s3;
            const unsigned zsb = (unsigned)g.v3 * 8u;
            double z[8];
            if (READ_Z) {
#pragma unroll
                for (int k = 0; k < 8; k++) z[k] = __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(zin + zo) + (size_t)k * zsb));
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
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double v = a * f[k];
                if (READ_Z) v = fma(b, z[k], v);
                f[k] = v;
                if (WRITE_Z) *reinterpret_cast<double*>(reinterpret_cast<char*>(zout + zo) + (size_t)k * zsb) = v;
            }
            if (DO_EXP) {
                const M3 r = exp_ta_times_su3(f, c, U);
                const unsigned uo = (unsigned)(t * 36 + MU * 9) * (unsigned)g.v3 + s3;
                m3_store(uout + uo, (unsigned)g.v3, r);
            }

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

constexpr size_t kTmBarOff = tm::SMEM_DATA;

template <bool READ_Z, bool WRITE_Z, bool DO_EXP>
__global__ void __launch_bounds__(tm::NTHREADS, 1)
k_tmarch_fused(const __grid_constant__ TmMaps maps, const tm::Tables* __restrict__ tab, Geom g, TmPlan pl, const double2* __restrict__ uin, double2* __restrict__ uout,
               const double* __restrict__ zin, double* __restrict__ zout, double a, double b, double c) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kTmBarOff);
    if (threadIdx.x == 0) {
        for (int i = 0; i < tm::S_RING + tm::R_RING; i++) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int mu = threadIdx.x / tm::SITES;  // warp-uniform: two warps per direction
    tm_run<READ_Z, WRITE_Z, DO_EXP>(mu, maps, g, pl, uin, uout, zin, zout, a, b, c, smem, bars, tab);
}

constexpr size_t kTmSmem = kTmBarOff + 8 * (tm::S_RING + tm::R_RING);

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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor maps of a link buffer viewed as [nslots*36 planes][nz][ny][2*nx doubles], one per box shape; cached per buffer
const TmMaps* tensor_maps_for(const double2* u, const Geom& g) {
    struct Key {
        const void* p;
        int nx, ny, nz, nslots;
        bool operator==(const Key& o) const { return p == o.p && nx == o.nx && ny == o.ny && nz == o.nz && nslots == o.nslots; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            return std::hash<const void*>()(k.p) ^ ((size_t)k.nx * 1315423911u) ^ ((size_t)k.ny << 12) ^ ((size_t)k.nz << 24) ^ ((size_t)k.nslots << 36);
        }
    };
    static std::unordered_map<Key, TmMaps, Hash> cache;
    static std::mutex mtx;
    static EncodeTiledFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mtx);
    const Key key{u, g.nx, g.ny, g.nz, g.nslots};
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return nullptr;
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
        if (encode(&maps.m[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double2*>(u), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return nullptr;
    }
    if (cache.size() > 1024) cache.clear();  // buffers come and go with the fields; the maps are cheap to rebuild
    return &cache.emplace(key, maps).first->second;
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
    // Short slab interiors stay in k_force_fused: with 8 slices per GPU (6 interior) the step is set by the exchange chain, a
    // segment start is 10 % of a 6-slice march, and all-k_force_fused measured 3-6 % faster on 8 GPUs (profiles/r1_tmarch.md);
    // from 16 slices per GPU on the t-marching interior wins (2.98 vs ~3.6 ms at 64^4 on 4 GPUs).  GFB200_TMARCH=2 forces it.
    if (g.nslots > g.tloc && t_count < 8 && mode != 2) return false;
    if (g.nx % tm::BX || g.ny % tm::BY || g.nz % tm::BZ) return false;
    if (uout == uin) return false;
    int dev = 0;
    cudaGetDevice(&dev);
    static int nsm = 0;
    if (nsm == 0) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // Multi-GPU slabs: a persistent grid fills every SM for the whole pass (231 KB of shared memory, 255 registers per thread), so
    // NCCL's send/recv kernels of the overlapped halo exchange could not start before it ends (measured: 7.5 ms instead of 5.2 ms
    // per 64^4/2 step).  There the grid is one CTA per (tile, segment) item instead: SMs free up every item and the high-priority
    // halo stream gets them first.  GFB200_TMARCH_RESERVE_SMS=n additionally keeps n SMs out of the plan.
    const bool slab = g.nslots > g.tloc;
    int reserve = 0;
    if (const char* er = getenv("GFB200_TMARCH_RESERVE_SMS")) reserve = atoi(er);
    if (reserve < 0 || reserve >= nsm) reserve = 0;
    int persistent = slab ? 0 : 1;
    if (const char* ep = getenv("GFB200_TMARCH_PERSISTENT")) persistent = atoi(ep);
    const int nsm_use = nsm - reserve;
    TmPlan pl = make_plan(g, t_begin, t_count, nsm_use);
    if (const char* e = getenv("GFB200_TMARCH_SEGLEN")) {  // test hook: force the t-segment length
        const int len = atoi(e);
        if (len >= 1) { pl.seg_len = len < t_count ? len : t_count; pl.nseg = (t_count + pl.seg_len - 1) / pl.seg_len; }
    }
    const TmMaps* maps = tensor_maps_for(uin, g);
    const tm::Tables* tab = device_tables(dev);
    if (!maps || !tab) return false;
    const long nitems = (long)pl.ntiles * pl.nseg;
    const unsigned grid = (unsigned)((!persistent || nitems < nsm_use) ? nitems : nsm_use);
#define GFB_LAUNCH_TM(R, W, E)                                                                                                  \
    do {                                                                                                                        \
        auto kern = k_tmarch_fused<R, W, E>;                                                                                    \
        static bool attr_set[64] = {};  /* per device: one process may drive several GPUs */                                    \
        if (!attr_set[dev & 63]) {                                                                                              \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmSmem) != cudaSuccess) {         \
                cudaGetLastError();                                                                                             \
                return false;                                                                                                   \
            }                                                                                                                   \
            attr_set[dev & 63] = true;                                                                                          \
        }                                                                                                                       \
        kern<<<grid, tm::NTHREADS, kTmSmem, st>>>(*maps, tab, g, pl, uin, uout, zin, zout, fa.a, fa.b, fa.c);                                \
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
