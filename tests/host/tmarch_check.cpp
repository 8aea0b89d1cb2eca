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
    // part sizes
    {
        int s = 0, r = 0;
        for (int i = 0; i < NBOX; i++) {
            int n = boxes[i].e[0] * boxes[i].e[1] * boxes[i].e[2];
            if (boxes[i].is_r) { if (boxes[i].base != r) { printf("R base\n"); return 1; } r += n; }
            else { if (boxes[i].base != s) { printf("S base\n"); return 1; } s += n; }
        }
        if (s != S_MATS || r != R_MATS) { printf("part sizes %d %d\n", s, r); return 1; }
    }
    std::vector<long> S(S_RING * S_MATS, -1), R(R_RING * R_MATS, -1);
    long checked = 0;
    const int nseg = (NTg + seg_len - 1) / seg_len;
    for (int z0 = 0; z0 < NZg; z0 += BZ) for (int y0 = 0; y0 < NYg; y0 += BY) for (int x0 = 0; x0 < NXg; x0 += BX)
    for (int seg = 0; seg < nseg; seg++) {
        const int tb = seg * seg_len;
        const int len = (seg_len < NTg - tb) ? seg_len : NTg - tb;
        auto copy_part = [&](int is_r, int t, int ring) {
            const int n = is_r ? R_MATS : S_MATS;
            for (int m = 0; m < n + 40; m++) {  // the kernel probes slots tid and tid+256 past the end as well
                int lam, x, y, z;
                if (!slot_to_pos(boxes, is_r, m, &lam, &x, &y, &z)) { if (m < n) { printf("slot_to_pos hole\n"); exit(1); } continue; }
                if (m >= n) { printf("slot_to_pos past end\n"); exit(1); }
                (is_r ? R : S)[ring * n + m] = link_id(lam, x0 + x, y0 + y, z0 + z, t);
            }
        };
        copy_part(0, tb, 0); copy_part(1, tb, 0); copy_part(0, tb + 1, 1);
        int rs = 0;
        for (int j = 0; j < len; j++) {
            const int t = tb + j;
            const int rs1 = (rs + 1) % 3, rs2 = (rs1 + 1) % 3;
            // emulate the asynchronous prefetch landing at the END of the step: read first, copy afterwards
            auto cen = [&](int off) { if (off < 0) { printf("missing operand\n"); exit(1); } return (off & 1) ? R[(j & 1) * R_MATS + (off >> 1) / (MAT_BYTES / 2)] : S[rs * S_MATS + off / MAT_BYTES]; };
            auto nxt = [&](int off) { if (off < 0 || (off & 1)) { printf("next-slice operand not in S\n"); exit(1); } return S[rs1 * S_MATS + off / MAT_BYTES]; };
            for (int mu = 0; mu < 4; mu++) for (int sz = 0; sz < BZ; sz++) for (int sy = 0; sy < BY; sy++) for (int sx = 0; sx < BX; sx++) {
                Operands op;
                make_operands(boxes, sx, sy, sz, mu, &op);
                const int X = x0 + sx, Y = y0 + sy, Z = z0 + sz;
                int e[4][4] = {{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,1}};
                auto id = [&](int lam, int dplus, int dminus) {
                    int p[4] = {X, Y, Z, t};
                    if (dplus >= 0) for (int d = 0; d < 4; d++) p[d] += e[dplus][d];
                    if (dminus >= 0) for (int d = 0; d < 4; d++) p[d] -= e[dminus][d];
                    return link_id(lam, p[0], p[1], p[2], p[3]);
                };
#define EXPECT(got, want, what) do { if ((got) != (want)) { printf("mismatch %s mu=%d site=%d,%d,%d t=%d tile=%d,%d,%d got %ld want %ld\n", what, mu, sx, sy, sz, t, x0, y0, z0, (long)(got), (long)(want)); return 1; } checked++; } while (0)
                EXPECT(cen(op.own), id(mu, -1, -1), "own");
                for (int jj = 0; jj < 3; jj++) {
                    const int nu = (mu + 1 + jj) & 3;
                    if (mu < 3 && nu < 3) {
                        EXPECT(cen(op.up[jj][0]), id(nu, -1, -1), "upA");
                        EXPECT(cen(op.up[jj][1]), id(mu, nu, -1), "upB");
                        EXPECT(cen(op.up[jj][2]), id(nu, mu, -1), "upC");
                        EXPECT(cen(op.dn[jj][0]), id(nu, -1, nu), "dnA");
                        EXPECT(cen(op.dn[jj][1]), id(mu, -1, nu), "dnB");
                        EXPECT(cen(op.dn[jj][2]), id(nu, mu, nu), "dnC");
                    } else if (mu < 3) {
                        EXPECT(cen(op.up[jj][0]), id(3, -1, -1), "tA");
                        EXPECT(nxt(op.up[jj][1]), id(mu, 3, -1), "tB");
                        EXPECT(cen(op.up[jj][2]), id(3, mu, -1), "tC");
                    } else {
                        EXPECT(cen(op.up[jj][0]), id(nu, -1, -1), "3upA");
                        EXPECT(cen(op.up[jj][1]), id(3, nu, -1), "3upB");
                        EXPECT(nxt(op.up[jj][2]), id(nu, 3, -1), "3upC");
                        EXPECT(cen(op.dn[jj][0]), id(nu, -1, nu), "3dnA");
                        EXPECT(cen(op.dn[jj][1]), id(3, -1, nu), "3dnB");
                        EXPECT(nxt(op.dn[jj][2]), id(nu, 3, nu), "3dnC");
                    }
                }
            }
            if (j + 1 < len) { copy_part(1, t + 1, (j + 1) & 1); copy_part(0, t + 2, rs2); }
            rs = rs1;
        }
    }
    printf("ok %ld operand reads\n", checked);
    return 0;
}
