// Host replay of the t-marching kernel's data movement (gaugefields.jl_b200/csrc/tmarch.cu) on symbolic link ids:
// the producer copies and the consumer operand reads use the same tmarch_geom.h functions and the same ring rotation as
// the kernel; every operand a link-thread reads must be exactly the link the six-staple stencil names
// (src/autostaples/wilsonloops.jl:468-484 in the reference).  Test infrastructure only (run by tests/test_tmarch_host.py).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gaugefields.jl_b200/csrc/tmarch_geom.h"

using namespace gfb::tm;

static int NXg, NYg, NZg, NTg;
static int wrapc(int c, int n) { return ((c % n) + n) % n; }
static long link_id(int lam, int x, int y, int z, int t) {
    return (((long)wrapc(t, NTg) * 4 + lam) * NZg + wrapc(z, NZg)) * NYg * NXg + (long)wrapc(y, NYg) * NXg + wrapc(x, NXg);
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: nx ny nz nt seg_len\n"); return 2; }
    NXg = atoi(argv[1]); NYg = atoi(argv[2]); NZg = atoi(argv[3]); NTg = atoi(argv[4]);
    const int seg_len = atoi(argv[5]);
    Box boxes[NBOX];
    if (make_boxes(boxes) != NBOX) { printf("box count\n"); return 1; }
    // part sizes, alignment, shapes
    {
        int s = 0, r = 0, sm = 0, rm = 0;
        for (int i = 0; i < NBOX; i++) {
            const int n = box_volume(boxes[i]);
            int& base = boxes[i].is_r ? r : s;
            if (boxes[i].base != base || (base & 127)) { printf("box base %d\n", i); return 1; }
            if ((i < NBOX_S) != (boxes[i].is_r == 0)) { printf("S boxes must come first\n"); return 1; }
            if (boxes[i].nlam < 1 || boxes[i].nlam > 3 || boxes[i].lam + boxes[i].nlam > 4) { printf("box directions %d\n", i); return 1; }
            base += pad128(boxes[i].nlam * n * MAT_BYTES);
            (boxes[i].is_r ? rm : sm) += boxes[i].nlam * n;
            const int sh = shape_index(boxes[i].e[0], boxes[i].e[1], boxes[i].e[2]);
            int e[3];
            if (sh < 0) { printf("box shape %d\n", i); return 1; }
            shape_extent(sh, e);
            if (e[0] != boxes[i].e[0] || e[1] != boxes[i].e[1] || e[2] != boxes[i].e[2]) { printf("shape_extent %d\n", i); return 1; }
            // a box never crosses the periodic boundary: inside the tile's extent or a single layer
            for (int d = 0; d < 3; d++)
                if (!((boxes[i].o[d] == 0 && boxes[i].e[d] == tile_extent(d)) || (boxes[i].e[d] == 1 && (boxes[i].o[d] == -1 || boxes[i].o[d] == tile_extent(d))))) { printf("box %d spans\n", i); return 1; }
        }
        if (s != S_BYTES || r != R_BYTES || sm != S_MATS || rm != R_MATS) { printf("part sizes %d %d %d %d\n", s, r, sm, rm); return 1; }
        if (R_BYTES >= 65536) { printf("descriptor offset overflow\n"); return 1; }
    }
    // ring strides: multiples of 1024 bytes (the swizzle pattern must be the same in every slot), mbarriers in the tail of R slot 0
    if (S_SLOT % 1024 || R_SLOT % 1024 || S_SLOT < S_BYTES || R_SLOT < R_BYTES + 8 * 16 || SMEM_DATA != S_RING * S_SLOT + R_RING * R_SLOT ||
        SMEM_DATA > 232448 || BAR_OFF < S_RING * S_SLOT + R_BYTES || BAR_OFF + 8 * 16 > S_RING * S_SLOT + R_SLOT) { printf("ring layout\n"); return 1; }
    // L2-blocked tile order (tile_of): a bijection onto the tile grid for every block shape, ragged blocks included
    {
        const int ntx = NXg / BX, nty = NYg / BY, ntz = NZg / BZ, nt3 = ntx * nty * ntz;
        for (int by = 1; by <= nty; by++)
            for (int bz = 1; bz <= ntz; bz++) {
                std::vector<char> seen(nt3, 0);
                for (int r = 0; r < nt3; r++) {
                    int tx, ty, tz;
                    tile_of(ntx, nty, ntz, by, bz, r, &tx, &ty, &tz);
                    if (tx < 0 || tx >= ntx || ty < 0 || ty >= nty || tz < 0 || tz >= ntz || seen[(tz * nty + ty) * ntx + tx]++) { printf("tile_of by %d bz %d r %d\n", by, bz, r); return 1; }
                }
            }
    }
    // shared memory (absolute byte address / 16) as 16-byte elements holding (link id * 9 + k); tile boxes are written the way
    // the 128-byte TMA swizzle writes them: chunk index (address bits 4-6) XOR address bits 7-9
    const bool swizzle = argc > 6 ? atoi(argv[6]) != 0 : true;
    std::vector<long> M(SMEM_DATA / 16, -1);
    auto part_base = [&](int is_r, int ring) { return is_r ? S_RING * S_SLOT + ring * R_SLOT : ring * S_SLOT; };
    auto swz = [&](unsigned a) { return swizzle ? a ^ (((a >> 7) & 7u) << 4) : a; };
    long checked = 0, wavefronts = 0, ideal = 0, shifted_px_conflicts = 0;
    const int nseg = (NTg + seg_len - 1) / seg_len;
    for (int z0 = 0; z0 < NZg; z0 += BZ) for (int y0 = 0; y0 < NYg; y0 += BY) for (int x0 = 0; x0 < NXg; x0 += BX)
    for (int seg = 0; seg < nseg; seg++) {
        const int tb = seg * seg_len;
        const int len = (seg_len < NTg - tb) ? seg_len : NTg - tb;
        // one TMA tensor copy per box: origin wrapped per coordinate, box-dense destination [k][z][y][x]
        auto copy_part = [&](int is_r, int t, int ring) {
            for (int i = 0; i < NBOX; i++) {
                const Box& b = boxes[i];
                if (b.is_r != is_r) continue;
                const int ox = wrapc(x0 + b.o[0], NXg), oy = wrapc(y0 + b.o[1], NYg), oz = wrapc(z0 + b.o[2], NZg);
                if (ox + b.e[0] > NXg || oy + b.e[1] > NYg || oz + b.e[2] > NZg) { printf("box crosses the boundary\n"); exit(1); }
                const bool tile = b.e[0] == BX && b.e[1] == BY && b.e[2] == BZ;
                unsigned dst = (unsigned)(part_base(is_r, ring) + b.base);
                if (tile && (dst & 1023)) { printf("swizzled box not 1024-aligned\n"); exit(1); }
                // destination = the copy's own order: [plane = (direction - lam)*9 + k][z][y][x]
                for (int pl = 0; pl < 9 * b.nlam; pl++) for (int z = 0; z < b.e[2]; z++) for (int y = 0; y < b.e[1]; y++) for (int x = 0; x < b.e[0]; x++) {
                    M[(tile ? swz(dst) : dst) / 16] = link_id(b.lam + pl / 9, ox + x, oy + y, oz + z, t) * 9 + pl % 9;
                    dst += 16;
                }
            }
        };
        copy_part(0, tb, 0); copy_part(1, tb, 0); copy_part(0, tb + 1, 1);
        int rs = 0;
        for (int j = 0; j < len; j++) {
            const int t = tb + j;
            const int rs1 = (rs + 1) % 3, rs2 = (rs1 + 1) % 3;
            // the kernel's address arithmetic (tm_step, `at`): base of the S slot of slice t + offset + R / next-S displacement,
            // XOR for swizzled boxes; element k is k*n*16 bytes further
            const unsigned sc = (unsigned)part_base(0, rs), d_r = (unsigned)part_base(1, j & 1) - sc, d_n = (unsigned)part_base(0, rs1) - sc;
            auto addr = [&](int d) {
                unsigned p = sc + ((unsigned)d & 0xFFFFu) + (((unsigned)d >> 28) & 1u) * d_r + (((unsigned)d >> 29) & 1u) * d_n;
                p ^= (p >> 3) & ((((unsigned)d >> 30) & 1u) * (swizzle ? 0x70u : 0u));
                return p;
            };
            // emulate the asynchronous prefetch landing at the END of the step: read first, copy afterwards
            // an operand is read as 9 elements at stride n*16 bytes; all nine must be the same link, k = 0..8 in order
            auto rd = [&](int d) {
                const unsigned p = addr(d), n = ((unsigned)d >> 16) & 0xFFu;
                long id = -1;
                for (int k = 0; k < 9; k++) {
                    const long v = M[(p + k * n * 16) / 16];
                    if (v < 0 || v % 9 != k || (k > 0 && v / 9 != id)) { printf("bad element k=%d\n", k); exit(1); }
                    id = v / 9;
                }
                return id;
            };
            // bank conflicts of one LDS.128 phase: the 8 x-consecutive lanes of a (mu, sz, sy) row read element 0 of descriptor i
            auto phase_wavefronts = [&](int mu, int sz, int sy, int i) {
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
                for (int sx = 0; sx < BX; sx++) {
                    int d[NDESC];
                    make_descriptors(boxes, sx, sy, sz, mu, d);
                    if (d[i] == 0 && i != 0) return 0;  // unused slot (nu = t lower staple of a spatial link)
                    const int c = (addr(d[i]) >> 4) & 7;
                    if (++cnt[c] > worst) worst = cnt[c];
                }
                return worst;
            };
            if (z0 == 0 && y0 == 0 && x0 == 0 && seg == 0)
                for (int mu = 0; mu < 4; mu++) for (int sz = 0; sz < BZ; sz++) for (int sy = 0; sy < BY; sy++) for (int i = 0; i < NDESC; i++) {
                    const int w = phase_wavefronts(mu, sz, sy, i);
                    if (!w) continue;
                    wavefronts += w; ideal += 1;
                    // the reads the swizzle is for: +x shifted operands of interior rows (tile box + the +x face box)
                    const bool px = (i == 2 && staple_dir(mu, 0) == 0) || (mu == 0 && (i == 3 || i == 9 || i == 15));
                    if (swizzle && px && w != 1) shifted_px_conflicts++;
                }
            for (int mu = 0; mu < 4; mu++) for (int sz = 0; sz < BZ; sz++) for (int sy = 0; sy < BY; sy++) for (int sx = 0; sx < BX; sx++) {
                int d[NDESC];
                make_descriptors(boxes, sx, sy, sz, mu, d);
                const int X = x0 + sx, Y = y0 + sy, Z = z0 + sz;
                int e[4][4] = {{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1}};
                auto id = [&](int lam, int dplus, int dminus) {
                    int p[4] = {X, Y, Z, t};
                    if (dplus >= 0) for (int q = 0; q < 4; q++) p[q] += e[dplus][q];
                    if (dminus >= 0) for (int q = 0; q < 4; q++) p[q] -= e[dminus][q];
                    return link_id(lam, p[0], p[1], p[2], p[3]);
                };
#define EXPECT(got, want, what) do { if ((got) != (want)) { printf("mismatch %s mu=%d site=%d,%d,%d t=%d tile=%d,%d,%d got %ld want %ld\n", what, mu, sx, sy, sz, t, x0, y0, z0, (long)(got), (long)(want)); return 1; } checked++; } while (0)
                EXPECT(rd(d[0]), id(mu, -1, -1), "own");
                for (int jj = 0; jj < 3; jj++) {
                    const int nu = staple_dir(mu, jj);
                    // upper staple  U_nu(x) U_mu(x+nu) U_nu(x+mu)^dag
                    EXPECT(rd(d[1 + 6 * jj]), id(nu, -1, -1), "upA");
                    EXPECT(rd(d[2 + 6 * jj]), id(mu, nu, -1), "upB");
                    EXPECT(rd(d[3 + 6 * jj]), id(nu, mu, -1), "upC");
                    if (nu < 3) {  // lower staple  U_nu(x-nu)^dag U_mu(x-nu) U_nu(x-nu+mu)
                        EXPECT(rd(d[4 + 6 * jj]), id(nu, -1, nu), "dnA");
                        EXPECT(rd(d[5 + 6 * jj]), id(mu, -1, nu), "dnB");
                        EXPECT(rd(d[6 + 6 * jj]), id(nu, mu, nu), "dnC");
                    } else if (jj != 2) { printf("nu = t must be the last iteration\n"); return 1; }
                }
            }
            if (j + 1 < len) { copy_part(1, t + 1, (j + 1) & 1); copy_part(0, t + 2, rs2); }
            rs = rs1;
        }
    }
    if (shifted_px_conflicts) { printf("+x shifted reads of tile boxes still conflict: %ld\n", shifted_px_conflicts); return 1; }
    printf("ok %ld operand reads, LDS wavefronts per phase %.3f (%s)\n", checked, ideal ? (double)wavefronts / ideal : 0.0, swizzle ? "swizzled" : "linear");
    return 0;
}
