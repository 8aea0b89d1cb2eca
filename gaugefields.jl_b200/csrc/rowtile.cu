// rowtile.cu -- the fused staple -> TA force -> kick -> exp(eps P) U pass as a persistent, TMA-fed row-tile kernel.
//
// Same arithmetic as k_force_fused (kernels.cu); different data movement.  Why (profiles/r1_ncu_force_fused.md): with
// per-thread LDG the pass is bound by L2 -> SM delivery (7 TB/s of 6.3 GB per 32^4 launch; a load-only build of
// k_force_fused is no faster than the full kernel and an L2-resident lattice is no faster per site), not by DRAM or the
// FP64 pipe.  Here a producer thread streams every link row a tile needs into shared memory with TMA tensor copies
// (cp.async.bulk.tensor.2d -> UTMALDG, completion on mbarriers), each distinct row once per tile (34 rows instead of
// 76 link loads per site), and the link-threads read their operands with fixed-latency LDS.128.
//
// STATUS (round 1): parity-green (tests/test_gpu_rowtile.py) but not yet faster than k_force_fused -- 14.8 ms vs 14.7 ms
// per 64^4 pass, 1.16 ms vs 0.84 ms at 32^4 -- so it is OFF by default (GFB200_ROWTILE=1 enables it).  Diagnostic builds
// (GFB_RT_DEBUG): copies only 10.8 ms, arithmetic only 12.3 ms at 64^4: the ring is too shallow for the ~2 us copy round
// trip and 8 consumer warps per SM do not cover the LDS/DFMA latencies.  Next steps in DESIGN.md section 9.
//
//   tile      one x-row (y, z, t) of NX = 32 or 64 sites; 4*NX link-threads (warp = 32 sites of one direction mu)
//   CTA       persistent, one per SM: 4*NX/32 consumer warps + 1 producer warp; tiles r = blockIdx.x, +gridDim.x, ...
//             in the same L2-friendly sweep order as decode_site()
//   rows      a "row" is one link direction of one lattice row: 9 planes x NX double2 (4.6 / 9.2 KB), fetched by ONE 2-D
//             tensor copy (box = 9 planes x 2*NX doubles of the [plane][site] view of the link field; issuing nine 1-D
//             bulk copies per row instead made the single producer thread the bottleneck at ~85 cycles per copy).
//             Per tile 34 rows instead of 76 per-thread link loads:
//               O_lam = U_lam @ row            lam = 0..3   resident for the whole tile   (double buffered)
//               M_lam = U_lam @ row - lam^     lam = 1..3   resident for the lower halves
//               ring  : 3 rows (upper) / 6 rows (lower) per half-stage h = 2(p-1)+{0,1}, p = 1..3, nu = mu XOR p
//             which operand of which link-thread reads which row at which x-shift is tabulated in rowtile_tables.h
//             (generated and checked against the oracle's staple sum by scripts/gen_rowtile_tables.py)
//   pipeline  full/empty mbarrier pairs per buffer; the producer runs up to one O set, one M set and NS ring slots ahead,
//             across tile boundaries.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gfb_internal.h"
#include "rowtile_tables.h"
#include "stencil.cuh"

#ifndef GFB_RT_DEBUG
#define GFB_RT_DEBUG 0  // 1: consumers skip the staple arithmetic; 2: producer skips the copies (pipeline diagnostics only)
#endif

namespace gfb {

namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
#if GFB_RT_DEBUG == 2
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
// global -> shared 2-D tensor copy (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst_smem)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

struct RowCoord {
    int y, z, t;
};
// r-th row of a launch covering slices t_begin + j*t_stride, j < t_count: y, z-in-chunk fastest, then t, then z-chunk
__device__ __forceinline__ RowCoord decode_row(const Geom& g, long r, int t_begin, int t_count) {
    RowCoord c;
    c.y = (int)(r % g.ny); r /= g.ny;
    const int zi = (int)(r % g.zc); r /= g.zc;
    c.t = t_begin + (int)(r % t_count) * g.t_stride;
    c.z = (int)(r / t_count) * g.zc + zi;
    return c;
}

template <int NX>
__device__ __forceinline__ M3 row_load(const unsigned char* row, int xs) {
    M3 r;
    const double2* p = reinterpret_cast<const double2*>(row) + xs;
#pragma unroll
    for (int k = 0; k < 9; k++) r.e[k] = p[k * NX];
    return r;
}

}  // namespace

template <int NXW, bool READ_Z, bool WRITE_Z, bool DO_EXP>
__global__ void __launch_bounds__((4 * NXW + 1) * 32, 1)
k_rowtile_fused(const __grid_constant__ CUtensorMap tmap, Geom g, int t_begin, int t_count, double2* __restrict__ uout, const double* __restrict__ zin,
                double* __restrict__ zout, double a, double b, double c) {
    constexpr int NX = 32 * NXW;
    constexpr int NCW = 4 * NXW;           // consumer warps
    constexpr int ROWB = 9 * NX * 16;      // bytes of one row
    constexpr int NO = 2;                  // O sets
    constexpr int NM = (NXW == 1) ? 2 : 1; // M sets
    constexpr int NS = (NXW == 1) ? 4 : 2; // ring slots (each kRingRowsMax rows)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* const o_base = smem;
    unsigned char* const m_base = o_base + NO * 4 * ROWB;
    unsigned char* const r_base = m_base + NM * 3 * ROWB;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(r_base + NS * kRingRowsMax * ROWB);
    uint64_t* const o_full = bars;
    uint64_t* const o_empty = o_full + NO;
    uint64_t* const m_full = o_empty + NO;
    uint64_t* const m_empty = m_full + NM;
    uint64_t* const r_full = m_empty + NM;
    uint64_t* const r_empty = r_full + NS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long nrows = (long)g.ny * g.nz * t_count;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NO; i++) { mbar_init(o_full + i, 1); mbar_init(o_empty + i, NCW); }
        for (int i = 0; i < NM; i++) { mbar_init(m_full + i, 1); mbar_init(m_empty + i, NCW); }
        for (int i = 0; i < NS; i++) { mbar_init(r_full + i, 1); mbar_init(r_empty + i, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------------------------------------------------------- producer: one thread issues all bulk copies
        if (lane != 0) return;
        auto issue_row = [&](unsigned char* dst, uint64_t* bar, const RowCoord& rc, int lam, int dy, int dz, int dt) {
            int y = rc.y + dy, z = rc.z + dz, t = rc.t;
            if (y < 0) y += g.ny; else if (y >= g.ny) y -= g.ny;
            if (z < 0) z += g.nz; else if (z >= g.nz) z -= g.nz;
            if (dt > 0) t = (t == g.tloc - 1) ? g.t_up_wrap : t + 1;
            else if (dt < 0) t = (t == 0) ? g.t_dn_wrap : t - 1;
            // tensor view: dim0 = doubles within a plane (2 per site), dim1 = plane index t*36 + lam*9 + k
#if GFB_RT_DEBUG == 2  // diagnostic: no data movement, barriers only
            if (lam >= 0) return;
#endif
            tma_load_2d(dst, &tmap, 2 * NX * (y + g.ny * z), t * 36 + lam * 9, bar);
        };
        unsigned it = 0;
        for (long r = blockIdx.x; r < nrows; r += gridDim.x, it++) {
            const RowCoord rc = decode_row(g, r, t_begin, t_count);
            {
                const unsigned ob = it % NO;
                mbar_wait(o_empty + ob, ((it / NO) & 1) ^ 1);
                mbar_arrive_expect_tx(o_full + ob, 4 * ROWB);
                for (int lam = 0; lam < 4; lam++) issue_row(o_base + (ob * 4 + lam) * ROWB, o_full + ob, rc, lam, 0, 0, 0);
            }
#pragma unroll 1
            for (int h = 0; h < 6; h++) {
                if (h == 1) {
                    const unsigned mb = it % NM;
                    mbar_wait(m_empty + mb, ((it / NM) & 1) ^ 1);
                    mbar_arrive_expect_tx(m_full + mb, 3 * ROWB);
                    for (int lam = 1; lam < 4; lam++)
                        issue_row(m_base + (mb * 3 + lam - 1) * ROWB, m_full + mb, rc, lam, lam == 1 ? -1 : 0, lam == 2 ? -1 : 0, lam == 3 ? -1 : 0);
                }
                const unsigned q = it * 6 + h, rs = q % NS;
                const int cnt = (h & 1) ? 6 : 3;
                mbar_wait(r_empty + rs, ((q / NS) & 1) ^ 1);
                mbar_arrive_expect_tx(r_full + rs, cnt * ROWB);
                for (int i = 0; i < cnt; i++)
                    issue_row(r_base + (rs * kRingRowsMax + i) * ROWB, r_full + rs, rc, c_ring_rows[h][i][0], c_ring_rows[h][i][1], c_ring_rows[h][i][2],
                              c_ring_rows[h][i][3]);
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers: one thread per (site, mu)
    const int mu = warp / NXW;
    const int xsite = (warp % NXW) * 32 + lane;
    unsigned it = 0;
    for (long r = blockIdx.x; r < nrows; r += gridDim.x, it++) {
        const RowCoord rc = decode_row(g, r, t_begin, t_count);
        const unsigned ob = it % NO, mb = it % NM;
        const unsigned char* const orow = o_base + ob * 4 * ROWB;
        const unsigned char* const mrow = m_base + mb * 3 * ROWB;
        Coord x;
        x.x = xsite; x.y = rc.y; x.z = rc.z; x.t = rc.t;
        const unsigned zo = mom_offset(g, x, mu);
        const unsigned zsb = (unsigned)g.v3 * 8u;
        double z[8];
        mbar_wait(o_full + ob, (it / NO) & 1);
        M3 s = m3_zero();
#pragma unroll
        for (int p = 0; p < 3; p++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const unsigned q = it * 6 + 2 * p + half, rs = q % NS;
                const unsigned char* const rrow = r_base + rs * kRingRowsMax * ROWB;
                if (p == 0 && half == 1) mbar_wait(m_full + mb, (it / NM) & 1);
                mbar_wait(r_full + rs, (q / NS) & 1);
                if (READ_Z && p == 2 && half == 0) {
                    // momenta of this link: requested one stage early so their latency hides behind the last staples
#pragma unroll
                    for (int k = 0; k < 8; k++) z[k] = __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(zin + zo) + (size_t)k * zsb));
                }
                auto operand = [&](int o) -> M3 {
                    const int buf = c_ops[p][mu][3 * half + o][0], slot = c_ops[p][mu][3 * half + o][1], sh = c_ops[p][mu][3 * half + o][2];
                    const unsigned char* base = (buf == 0) ? orow : (buf == 1) ? mrow : rrow;
                    return row_load<NX>(base + slot * ROWB, (xsite + sh) & (NX - 1));
                };
#if GFB_RT_DEBUG == 1
                if (p == 2 && half == 1) m3_add(s, operand(2));
#else
                if (half == 0) {
                    M3 t = mul_nn(operand(0), operand(1));
                    mac_nd(s, t, operand(2));
                } else {
                    M3 t = mul_dn(operand(0), operand(1));
                    mac_nn(s, t, operand(2));
                }
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(r_empty + rs);
            }
        }
        if (lane == 0) mbar_arrive(m_empty + mb);
        const M3 umu = row_load<NX>(orow + mu * ROWB, xsite);
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty + ob);
        double f[8];
        {
            M3 w = mul_nd(umu, s);
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
            M3 e = exp_ta(f, c);
            M3 res = mul_nn(e, umu);
            store_link(uout, g, x, mu, res);
        }
    }
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of a link buffer viewed as [nslots*36 planes][2*V3 doubles], box = 9 planes x 2*NX doubles; cached per buffer
const CUtensorMap* tensor_map_for(const double2* u, const Geom& g) {
    struct Key {
        const void* p;
        int v3, nslots, nx;
        bool operator==(const Key& o) const { return p == o.p && v3 == o.v3 && nslots == o.nslots && nx == o.nx; }
    };
    struct Hash {
        size_t operator()(const Key& k) const { return std::hash<const void*>()(k.p) ^ ((size_t)k.v3 * 1315423911u) ^ ((size_t)k.nslots << 20) ^ (size_t)k.nx; }
    };
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    static std::mutex mtx;
    static EncodeTiledFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mtx);
    Key key{u, g.v3, g.nslots, g.nx};
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return nullptr;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)g.v3 * 2, (cuuint64_t)g.nslots * 36};
    const cuuint64_t gstride[1] = {(cuuint64_t)g.v3 * 16};
    const cuuint32_t box[2] = {(cuuint32_t)g.nx * 2, 9};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double2*>(u), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return nullptr;
    if (cache.size() > 4096) cache.clear();  // buffers come and go with the fields; the map is cheap to rebuild
    return &cache.emplace(key, m).first->second;
}

}  // namespace

template <int NXW>
static bool launch_rowtile_nx(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                              const FusedArgs& fa) {
    constexpr int NX = 32 * NXW;
    constexpr int ROWB = 9 * NX * 16;
    constexpr int NO = 2, NM = (NXW == 1) ? 2 : 1, NS = (NXW == 1) ? 4 : 2;
    const size_t smem = (size_t)(NO * 4 + NM * 3 + NS * kRingRowsMax) * ROWB + 64 * sizeof(uint64_t);
    const CUtensorMap* tm = tensor_map_for(uin, g);
    if (!tm) return false;
    static int nsm = 0;
    if (nsm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    const long nrows = (long)g.ny * g.nz * t_count;
    const unsigned grid = (unsigned)(nrows < nsm ? nrows : nsm);
    const unsigned block = (4 * NXW + 1) * 32;
#define GFB_LAUNCH_RT(R, W, E)                                                                                            \
    do {                                                                                                                  \
        auto kern = k_rowtile_fused<NXW, R, W, E>;                                                                        \
        static bool attr_set[64] = {};  /* per device */                                                                   \
        int dev_ = 0;                                                                                                     \
        cudaGetDevice(&dev_);                                                                                             \
        if (!attr_set[dev_ & 63]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set[dev_ & 63] = true; } \
        kern<<<grid, block, smem, st>>>(*tm, g, t_begin, t_count, uout, zin, zout, fa.a, fa.b, fa.c);                      \
    } while (0)
    if (fa.read_z) {
        if (fa.do_exp) GFB_LAUNCH_RT(true, true, true);
        else GFB_LAUNCH_RT(true, true, false);
    } else {
        if (fa.do_exp) {
            if (fa.write_z) GFB_LAUNCH_RT(false, true, true);
            else GFB_LAUNCH_RT(false, false, true);
        } else GFB_LAUNCH_RT(false, true, false);
    }
#undef GFB_LAUNCH_RT
    return true;
}

// Returns false when the geometry is not covered (NX must be 32 or 64): the caller then uses k_force_fused.
bool launch_rowtile_fused(cudaStream_t st, const Geom& g, int t_begin, int t_count, const double2* uin, double2* uout, const double* zin, double* zout,
                          const FusedArgs& fa) {
    static const int mode = [] {
        const char* e = getenv("GFB200_ROWTILE");  // 1 enables the row-tile kernel (off by default: see the header note on status)
        return e ? atoi(e) : 0;
    }();
    if (!mode) return false;
    if (uout == uin) return false;  // tiles read neighbouring rows: the output must be a different buffer
    if (g.nx == 32) return launch_rowtile_nx<1>(st, g, t_begin, t_count, uin, uout, zin, zout, fa);
    if (g.nx == 64) return launch_rowtile_nx<2>(st, g, t_begin, t_count, uin, uout, zin, zout, fa);
    return false;
}

}  // namespace gfb
