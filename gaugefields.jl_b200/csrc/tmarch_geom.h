// tmarch_geom.h -- shared-memory geometry of the t-marching fused pass (tmarch.cu).
//
// Plain C++ (no CUDA types) so that the same functions run in the kernel and in the host checker
// tests/host/tmarch_check.cpp, which replays producer copies and consumer reads on symbolic link ids.
//
// A CTA owns a spatial tile of BX x BY x BZ sites and marches along t.  For the six-staple stencil of slice t it needs
//   * the "full" set of slice t:  for a spatial link direction lam the offsets
//         {0, -e_lam} + {0, +e_i (i != lam)}   and   -e_j (j != lam)          (tile, faces, the (+i,-lam) edges)
//     and for lam = t the tile and its six spatial faces;
//   * of slice t+1 only the spatial links on the tile and on the -e_lam face.
// The backward-t staple of slice t is carried in registers from slice t-1 (the same thread computed it there),
// so slice t-1 is never resident.  The full set is split in two parts with separate rings:
//   S part  (248 matrices)  spatial lam on the box  tile extended by one site towards -e_lam    [needed as t+1 AND as t]
//   R part  (428 matrices)  everything else                                                    [needed only as t]
// Ring depth: S 3 (t, t+1 and the t+2 being fetched), R 2 (t and the t+1 being fetched): 1600 matrices = 230400 bytes.
// Inside a part a matrix is 144 contiguous bytes (array of structures): the 8 x-consecutive lanes of an LDS.128 phase
// hit stride-9 16-byte groups, which is conflict-free.
#pragma once

#if defined(__CUDACC__)
#define TM_HD __host__ __device__ __forceinline__
#define TM_UNROLL _Pragma("unroll")
#else
#define TM_HD inline
#define TM_UNROLL
#endif

namespace gfb {
namespace tm {

constexpr int BX = 8, BY = 4, BZ = 2;
constexpr int SITES = BX * BY * BZ;      // 64 sites, 256 link-threads
constexpr int NTHREADS = 4 * SITES;
constexpr int MAT_BYTES = 144;
constexpr int NBOX = 22;
constexpr int S_MATS = 248, R_MATS = 428;
constexpr int S_BYTES = S_MATS * MAT_BYTES, R_BYTES = R_MATS * MAT_BYTES;
constexpr int S_RING = 3, R_RING = 2;
constexpr int SMEM_DATA = S_RING * S_BYTES + R_RING * R_BYTES;  // 230400

struct Box {
    signed char lam, is_r;
    signed char o[3];  // origin relative to the tile origin (-1 .. B)
    signed char e[3];  // extents
    short base;        // first slot inside its part
};

TM_HD int tile_extent(int d) { return d == 0 ? BX : d == 1 ? BY : BZ; }

// Fills the 22 boxes: 3 S boxes (lam = 0,1,2), then R boxes: per spatial lam [+i1, +i2, -i1, -i2] (i1 < i2 the other two
// spatial directions), then lam = 3: tile, +x, +y, +z, -x, -y, -z.  Returns the number of boxes.
TM_HD int make_boxes(Box* b) {
    int n = 0;
    int sbase = 0, rbase = 0;
    for (int lam = 0; lam < 3; lam++) {
        Box x;
        x.lam = (signed char)lam; x.is_r = 0;
        for (int d = 0; d < 3; d++) { x.o[d] = (signed char)(d == lam ? -1 : 0); x.e[d] = (signed char)(tile_extent(d) + (d == lam ? 1 : 0)); }
        x.base = (short)sbase;
        sbase += x.e[0] * x.e[1] * x.e[2];
        b[n++] = x;
    }
    for (int lam = 0; lam < 3; lam++) {
        for (int sgn = +1; sgn >= -1; sgn -= 2) {
            for (int i = 0; i < 3; i++) {
                if (i == lam) continue;
                Box x;
                x.lam = (signed char)lam; x.is_r = 1;
                for (int d = 0; d < 3; d++) {
                    if (d == i) { x.o[d] = (signed char)(sgn > 0 ? tile_extent(d) : -1); x.e[d] = 1; }
                    else if (d == lam && sgn > 0) { x.o[d] = -1; x.e[d] = (signed char)(tile_extent(d) + 1); }
                    else { x.o[d] = 0; x.e[d] = (signed char)tile_extent(d); }
                }
                x.base = (short)rbase;
                rbase += x.e[0] * x.e[1] * x.e[2];
                b[n++] = x;
            }
        }
    }
    {
        Box x;
        x.lam = 3; x.is_r = 1;
        for (int d = 0; d < 3; d++) { x.o[d] = 0; x.e[d] = (signed char)tile_extent(d); }
        x.base = (short)rbase;
        rbase += SITES;
        b[n++] = x;
        for (int sgn = +1; sgn >= -1; sgn -= 2)
            for (int i = 0; i < 3; i++) {
                Box y;
                y.lam = 3; y.is_r = 1;
                for (int d = 0; d < 3; d++) {
                    if (d == i) { y.o[d] = (signed char)(sgn > 0 ? tile_extent(d) : -1); y.e[d] = 1; }
                    else { y.o[d] = 0; y.e[d] = (signed char)tile_extent(d); }
                }
                y.base = (short)rbase;
                rbase += y.e[0] * y.e[1] * y.e[2];
                b[n++] = y;
            }
    }
    return n;
}

// byte offset (within its part) of link lam at tile-relative position (x, y, z), bit 0 set when it lives in the R part;
// -1 when the position is not resident (never happens for the stencil's operands: checked by the host checker)
TM_HD int lookup(const Box* b, int lam, int x, int y, int z) {
    for (int i = 0; i < NBOX; i++) {
        if (b[i].lam != lam) continue;
        const int dx = x - b[i].o[0], dy = y - b[i].o[1], dz = z - b[i].o[2];
        if (dx < 0 || dy < 0 || dz < 0 || dx >= b[i].e[0] || dy >= b[i].e[1] || dz >= b[i].e[2]) continue;
        const int slot = b[i].base + dx + b[i].e[0] * (dy + b[i].e[1] * dz);
        return slot * MAT_BYTES + (b[i].is_r ? 1 : 0);
    }
    return -1;
}

// inverse map used by the producer: slot m of part `is_r` -> (lam, x, y, z); returns false past the end of the part
TM_HD bool slot_to_pos(const Box* b, int is_r, int m, int* lam, int* x, int* y, int* z) {
    for (int i = 0; i < NBOX; i++) {
        if (b[i].is_r != is_r) continue;
        const int n = b[i].e[0] * b[i].e[1] * b[i].e[2];
        const int r = m - b[i].base;
        if (r < 0 || r >= n) continue;
        *lam = b[i].lam;
        *x = b[i].o[0] + r % b[i].e[0];
        *y = b[i].o[1] + (r / b[i].e[0]) % b[i].e[1];
        *z = b[i].o[2] + r / (b[i].e[0] * b[i].e[1]);
        return true;
    }
    return false;
}

// Operand table of link-thread (site, mu).  For j = 0..2, nu = (mu+1+j) & 3:
//   upper staple  A B C^dag   with A = U_nu(x), B = U_mu(x+nu), C = U_nu(x+mu)
//   lower staple  A^dag B C   with A = U_nu(x-nu), B = U_mu(x-nu), C = U_nu(x-nu+mu)     (nu spatial only; nu = t is carried)
// Which slice an operand lives in follows from (mu, nu) alone:  a +t shift -> S part of slice t+1, otherwise slice t.
struct Operands {
    int up[3][3];  // [j][A,B,C] byte offset | is_r
    int dn[3][3];
    int own;
};
TM_HD int shifted(const Box* b, int lam, const int* p, int d_plus, int d_minus) {
    int q[3] = {p[0], p[1], p[2]};
    if (d_plus >= 0 && d_plus < 3) q[d_plus] += 1;
    if (d_minus >= 0 && d_minus < 3) q[d_minus] -= 1;
    return lookup(b, lam, q[0], q[1], q[2]);
}
TM_HD void make_operands(const Box* b, int sx, int sy, int sz, int mu, Operands* o) {
    const int p[3] = {sx, sy, sz};
    o->own = lookup(b, mu, sx, sy, sz);
    TM_UNROLL
    for (int j = 0; j < 3; j++) {
        const int nu = (mu + 1 + j) & 3;
        // shifts along t select the slice (compile-time in the kernel); here only the spatial part of the shift matters
        o->up[j][0] = shifted(b, nu, p, -1, -1);
        o->up[j][1] = shifted(b, mu, p, nu, -1);
        o->up[j][2] = shifted(b, nu, p, mu, -1);
        if (nu < 3) {
            o->dn[j][0] = shifted(b, nu, p, -1, nu);
            o->dn[j][1] = shifted(b, mu, p, -1, nu);
            o->dn[j][2] = shifted(b, nu, p, mu, nu);
        } else {
            o->dn[j][0] = o->dn[j][1] = o->dn[j][2] = -1;
        }
    }
}

}  // namespace tm
}  // namespace gfb
