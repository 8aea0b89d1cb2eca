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
//   S part  (248 matrices)  spatial lam on the tile and on its -e_lam face                     [needed as t+1 AND as t]
//   R part  (428 matrices)  everything else                                                    [needed only as t]
// Ring depth: S 3 (t, t+1 and the t+2 being fetched), R 2 (t and the t+1 being fetched).
// Every part is a list of BOXES (21 per slice), each fetched by one TMA tensor copy of the 9 element planes of one to three
// consecutive link directions (cp.async.bulk.tensor.4d, box = 2*ex doubles x ey x ez x 9*nlam planes); a box never crosses the periodic boundary because it is
// either inside the tile's extent or a single halo layer in each direction, so its origin is wrapped per coordinate.
// Inside a box the layout is the copy's: [k][z][y][x] 16-byte elements, i.e. element k of the matrix at box position i sits at
// box_base + (k*n + i)*16 with n the box volume: the 8 x-consecutive lanes of an LDS.128 phase read 128 contiguous bytes.
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
constexpr int NBOX = 21;
constexpr int NBOX_S = 4;               // the S part's boxes come first in the table
constexpr int NSHAPE = 7;
constexpr int NMAP = 3 * NSHAPE;        // tensor maps: box shape x number of merged directions (1..3)
constexpr int S_MATS = 248, R_MATS = 428;
constexpr int S_BYTES = 35712;           // 248 matrices, every S box is a multiple of 8 matrices (no padding)
constexpr int R_BYTES = 61952;           // 428 matrices + 320 bytes of padding (box bases are 128-byte aligned for TMA)
constexpr int S_RING = 3, R_RING = 2;
// Ring strides are multiples of 1024 bytes: the tile boxes (8 x-sites = one 128-byte row per (k, z, y)) are copied with the
// 128-byte TMA swizzle, whose XOR pattern is a function of the ABSOLUTE shared-memory address bits 7-9; with every tile box at
// a 1024-byte multiple the pattern is the same in every ring slot (see lookup(): swizzled operands).  The tail of R slot 0
// (R_SLOT - R_BYTES = 512 bytes) holds the mbarriers, so the whole 227 KB opt-in maximum is used: 3*35840 + 2*62464 = 232448.
constexpr int S_SLOT = 35840, R_SLOT = 62464;
constexpr int SMEM_DATA = S_RING * S_SLOT + R_RING * R_SLOT;  // 232448
constexpr int BAR_OFF = S_RING * S_SLOT + R_BYTES;             // mbarriers: up to 64 of them in the tail of R slot 0

// One TMA copy: the links of nlam consecutive directions lam .. lam+nlam-1 on a box of positions (their 9*nlam element planes
// are contiguous in a slice, so directions that need the same positions share a copy: 21 copies per slice instead of 31).
// Destination layout [plane][z][y][x]: direction lam+i occupies the sub-box at base + i*9*n*16, each sub-box is [k][pos].
struct Box {
    signed char lam, nlam, is_r;
    signed char o[3];  // origin relative to the tile origin (-1 .. B)
    signed char e[3];  // extents
    int base;          // byte offset inside its part (128-byte aligned)
};

// L2-blocked tile order: r in [0, ntx*nty*ntz) -> tile (tx, ty, tz).  Tiles are enumerated x fastest inside blocks of
// ntx x by x bz tiles, blocks y fastest; the last block of a row / the last row are ragged when by, bz do not divide nty, ntz.
// A bijection for any 1 <= by <= nty, 1 <= bz <= ntz (tests/host/tmarch_check.cpp).
TM_HD void tile_of(int ntx, int nty, int ntz, int by, int bz, int r, int* tx, int* ty, int* tz) {
    const int row = ntx * nty * bz;
    const int zb = r / row;
    r -= zb * row;
    const int hz = (ntz - zb * bz) < bz ? (ntz - zb * bz) : bz;
    const int blk = ntx * by * hz;
    const int yb = r / blk;
    r -= yb * blk;
    const int hy = (nty - yb * by) < by ? (nty - yb * by) : by;
    *tx = r % ntx;
    r /= ntx;
    *ty = yb * by + r % hy;
    *tz = zb * bz + r / hy;
}
TM_HD int tile_extent(int d) { return d == 0 ? BX : d == 1 ? BY : BZ; }
TM_HD int box_volume(const Box& b) { return b.e[0] * b.e[1] * b.e[2]; }
TM_HD int pad128(int n) { return (n + 127) & ~127; }
// index of the box shape among the NSHAPE distinct (ex, ey, ez): one tensor map per shape
TM_HD int shape_index(int ex, int ey, int ez) {
    if (ex == BX && ey == BY && ez == BZ) return 0;
    if (ex == 1 && ey == BY && ez == BZ) return 1;
    if (ex == BX && ey == 1 && ez == BZ) return 2;
    if (ex == BX && ey == BY && ez == 1) return 3;
    if (ex == 1 && ey == 1 && ez == BZ) return 4;
    if (ex == 1 && ey == BY && ez == 1) return 5;
    if (ex == BX && ey == 1 && ez == 1) return 6;
    return -1;
}
TM_HD void shape_extent(int s, int* e) {
    e[0] = (s == 1 || s == 4 || s == 5) ? 1 : BX;
    e[1] = (s == 2 || s == 4 || s == 6) ? 1 : BY;
    e[2] = (s == 3 || s == 5 || s == 6) ? 1 : BZ;
}

// Fills the NBOX boxes.  A box is described by the set of directions in which it is a halo layer:
//   lo[d] = -1 / +1  -> the single layer at -1 / at B_d ;  0 -> the tile's extent in d
TM_HD void put_box(Box* b, int* n, int* sbase, int* rbase, int lam, int nlam, int is_r, int l0, int l1, int l2) {
    const int lo[3] = {l0, l1, l2};
    Box x;
    x.lam = (signed char)lam; x.nlam = (signed char)nlam; x.is_r = (signed char)is_r;
    for (int d = 0; d < 3; d++) {
        x.o[d] = (signed char)(lo[d] < 0 ? -1 : (lo[d] > 0 ? tile_extent(d) : 0));
        x.e[d] = (signed char)(lo[d] != 0 ? 1 : tile_extent(d));
    }
    int* base = is_r ? rbase : sbase;
    x.base = *base;
    *base += pad128(nlam * box_volume(x) * MAT_BYTES);
    b[(*n)++] = x;
}
// a face layer (sign sgn, direction i) in the R part for every link direction except `skip` (-1: none), merged into runs
TM_HD void put_face_runs(Box* b, int* n, int* sbase, int* rbase, int i, int sgn, int skip) {
    int l[3] = {0, 0, 0};
    l[i] = sgn;
    int lam = 0;
    while (lam < 4) {
        if (lam == skip) { lam++; continue; }
        int run = 1;
        while (lam + run < 4 && lam + run != skip) run++;
        put_box(b, n, sbase, rbase, lam, run, 1, l[0], l[1], l[2]);
        lam += run;
    }
}
TM_HD int make_boxes(Box* b) {
    int n = 0, sbase = 0, rbase = 0;
    // S part: the three spatial directions on the tile (one copy), each one's -e_lam layer
    put_box(b, &n, &sbase, &rbase, 0, 3, 0, 0, 0, 0);
    for (int lam = 0; lam < 3; lam++) put_box(b, &n, &sbase, &rbase, lam, 1, 0, lam == 0 ? -1 : 0, lam == 1 ? -1 : 0, lam == 2 ? -1 : 0);
    // R part: U_t on the tile
    put_box(b, &n, &sbase, &rbase, 3, 1, 1, 0, 0, 0);
    for (int i = 0; i < 3; i++) {
        put_face_runs(b, &n, &sbase, &rbase, i, +1, i);  // +e_i face: every direction but i
        put_face_runs(b, &n, &sbase, &rbase, i, -1, i);  // -e_i face: every direction but i (direction i's is the S layer)
    }
    for (int lam = 0; lam < 3; lam++)                    // (+e_i, -e_lam) edges of a spatial direction
        for (int i = 0; i < 3; i++) {
            if (i == lam) continue;
            int l[3] = {0, 0, 0};
            l[i] = +1; l[lam] = -1;
            put_box(b, &n, &sbase, &rbase, lam, 1, 1, l[0], l[1], l[2]);
        }
    return n;
}

// operand descriptor of link lam at tile-relative position (x, y, z):
//   bits 0-15 byte offset of element 0 inside its part, bits 16-23 box volume n (element k is k*n*16 bytes further),
//   bit 24 set when the box belongs to the R part, bit 25 when the box is a full tile (copied with the 128-byte swizzle: the
//   16-byte chunk index, address bits 4-6, is XORed with address bits 7-9; n*16 = 1024 there, so the XOR term is the same for
//   all nine elements);  -1 when the position is not resident.
// Why the swizzle: a warp reads 8 x-consecutive sites per LDS.128 phase.  Unshifted that is one 128-byte row.  Shifted by +x
// it is positions 1..7 of the row plus one element of the +x face box, which in the linear layout lands on the banks of one
// of the seven (2 wavefronts instead of 1; 19 % of all wavefronts were such conflicts, profiles/r1_tmarch.md).  With the XOR
// swizzle, row r = y + 4z (+ 8k) holds position p at chunk p ^ (r & 7), so positions 1..7 leave exactly chunk r & 7 free --
// and the face box [k][z][y] (dense, 128-byte aligned) holds element (y, z) at chunk (y + 4z) & 7 = r & 7: conflict-free.
TM_HD int lookup(const Box* b, int lam, int x, int y, int z) {
    for (int i = 0; i < NBOX; i++) {
        if (lam < b[i].lam || lam >= b[i].lam + b[i].nlam) continue;
        const int dx = x - b[i].o[0], dy = y - b[i].o[1], dz = z - b[i].o[2];
        if (dx < 0 || dy < 0 || dz < 0 || dx >= b[i].e[0] || dy >= b[i].e[1] || dz >= b[i].e[2]) continue;
        const int idx = dx + b[i].e[0] * (dy + b[i].e[1] * dz);
        const int n = box_volume(b[i]);
        const int swz = (b[i].e[0] == BX && b[i].e[1] == BY && b[i].e[2] == BZ) ? 1 : 0;
        return (b[i].base + (lam - b[i].lam) * 9 * n * 16 + idx * 16) | (n << 16) | ((b[i].is_r ? 1 : 0) << 24) | (swz << 25);
    }
    return -1;
}

// j-th staple direction of a link in direction mu: the other three directions in ascending order, so that nu = t is the LAST
// iteration of a spatial link (the kernel fuses it with the carried backward-t staple)
TM_HD int staple_dir(int mu, int j) { return j < mu ? j : j + 1; }

// Operand table of link-thread (site, mu).  For j = 0..2, nu = staple_dir(mu, j):
//   upper staple  A B C^dag   with A = U_nu(x), B = U_mu(x+nu), C = U_nu(x+mu)
//   lower staple  A^dag B C   with A = U_nu(x-nu), B = U_mu(x-nu), C = U_nu(x-nu+mu)     (nu spatial only; nu = t is carried)
// Which slice an operand lives in follows from (mu, nu) alone:  a +t shift -> S part of slice t+1, otherwise slice t.
struct Operands {
    int up[3][3];  // [j][A,B,C] descriptors (lookup())
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
        const int nu = staple_dir(mu, j);
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

// Final per-thread operand descriptors of the kernel, in the order the hot loop uses them:
//   d[0] own link;  d[1+6j .. 6+6j] = up A,B,C, dn A,B,C of iteration j (dn of the nu = t iteration of a spatial link: unused, 0)
//   bits 0-15 byte offset, 16-23 box volume, bit 28: R part of slice t, bit 29: S part of slice t+1 (neither: S part of slice t),
//   bit 30: the box is swizzled (full tile)
constexpr int NDESC = 19;
TM_HD void make_descriptors(const Box* b, int sx, int sy, int sz, int mu, int* d) {
    Operands op;
    make_operands(b, sx, sy, sz, mu, &op);
    // lookup() bits 24 (R part) and 25 (swizzled) move to bits 28 and 30
    auto fin = [](int v, bool next) { return (v & 0xFFFFFF) | (next ? (2 << 28) : (((v >> 24) & 1) << 28)) | (((v >> 25) & 1) << 30); };
    d[0] = fin(op.own, false);
    for (int j = 0; j < 3; j++) {
        const int nu = staple_dir(mu, j);
        for (int o = 0; o < 3; o++) {
            // a +t shift moves the operand to slice t+1: B of the upper staple when nu = t, C of both staples when mu = t
            const bool next_up = (o == 1 && mu < 3 && nu == 3) || (o == 2 && mu == 3);
            const bool next_dn = (o == 2 && mu == 3);
            const int u = op.up[j][o], l = op.dn[j][o];
            d[1 + 6 * j + o] = fin(u, next_up);
            d[4 + 6 * j + o] = (nu == 3) ? 0 : fin(l, next_dn);
        }
    }
}

// what the kernel reads at start-up instead of recomputing the geometry per CTA (built once on the host)
struct Tables {
    int desc[NDESC][NTHREADS];  // [i][thread]: coalesced
    int box[NBOX][8];           // o0, o1, o2, first direction, is_r, base, tensor-map index, bytes
};
inline void make_tables(Tables* t) {
    Box b[NBOX];
    make_boxes(b);
    for (int tid = 0; tid < NTHREADS; tid++) {
        const int mu = tid / SITES, sidx = tid % SITES;
        int d[NDESC];
        make_descriptors(b, sidx % BX, (sidx / BX) % BY, sidx / (BX * BY), mu, d);
        for (int i = 0; i < NDESC; i++) t->desc[i][tid] = d[i];
    }
    for (int i = 0; i < NBOX; i++) {
        t->box[i][0] = b[i].o[0]; t->box[i][1] = b[i].o[1]; t->box[i][2] = b[i].o[2];
        t->box[i][3] = b[i].lam; t->box[i][4] = b[i].is_r; t->box[i][5] = b[i].base;
        t->box[i][6] = 3 * shape_index(b[i].e[0], b[i].e[1], b[i].e[2]) + (b[i].nlam - 1);
        t->box[i][7] = b[i].nlam * box_volume(b[i]) * MAT_BYTES;
    }
}

}  // namespace tm
}  // namespace gfb
