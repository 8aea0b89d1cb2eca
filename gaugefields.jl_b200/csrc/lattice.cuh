// lattice.cuh -- slab geometry and device-side indexing.
//
// Device layout (per GPU slab, see DESIGN.md "Data layout in HBM"):
//   links    double2 U[t][mu*9 + k][s3]      t in [0, tloc + 2*has_halo), k = 3*row + col
//   momenta  double  P[t][mu*8 + a][s3]      t in [0, tloc)
// with s3 = x + nx*(y + ny*z) (x fastest, the reference's site order, src/API.jl:516-529).
// One time-slice of all four links is a single contiguous chunk, so the t-halo exchange
//  sends/receives straight from/to field memory without packing.
// Halo slots: t = tloc holds the neighbour's first slice (the "t+1" halo), t = tloc+1 the
// neighbour's last slice (the "t-1" halo).  On a single slab the wrap is done by indexing.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace gfb {

struct Geom {
    int nx, ny, nz, tloc;  // local extents (x, y, z undivided; t = slab)
    int v3;                // nx*ny*nz
    int zc;                // z-chunk of the L2-blocked traversal (divides nz)
    int t_up_wrap;         // storage slot of t+1 seen from t = tloc-1
    int t_dn_wrap;         // storage slot of t-1 seen from t = 0
    int t0;                // global t of local slice 0
    int nt;                // global NT
    int nslots;            // tloc (+2 with halo)
    int t_stride;          // launch covers slices t_begin + j*t_stride, j < t_count (1 = contiguous; tloc-1 = the two boundary slices)
};

struct Coord {
    int x, y, z, t;
};

// n-th site of a launch covering local slices [t_begin, t_begin + t_count): x, y, z-in-chunk
// fastest, then t, then z-chunk.  Keeps the t+-1 reuse distance at nx*ny*zc sites so that the
// neighbouring time-slices are served from L2 even at 64^3 spatial volume.
__device__ __forceinline__ Coord decode_site(const Geom& g, long n, int t_begin, int t_count) {
    Coord c;
    c.x = (int)(n % g.nx); n /= g.nx;
    c.y = (int)(n % g.ny); n /= g.ny;
    int zi = (int)(n % g.zc); n /= g.zc;
    c.t = t_begin + (int)(n % t_count) * g.t_stride;
    int zo = (int)(n / t_count);
    c.z = zo * g.zc + zi;
    return c;
}

__device__ __forceinline__ Coord step(const Geom& g, Coord c, int dir, int sgn) {
    if (dir == 0) {
        c.x += sgn;
        if (c.x == g.nx) c.x = 0; else if (c.x < 0) c.x = g.nx - 1;
    } else if (dir == 1) {
        c.y += sgn;
        if (c.y == g.ny) c.y = 0; else if (c.y < 0) c.y = g.ny - 1;
    } else if (dir == 2) {
        c.z += sgn;
        if (c.z == g.nz) c.z = 0; else if (c.z < 0) c.z = g.nz - 1;
    } else {
        if (sgn > 0) c.t = (c.t == g.tloc - 1) ? g.t_up_wrap : c.t + 1;
        else c.t = (c.t == 0) ? g.t_dn_wrap : c.t - 1;
    }
    return c;
}

__device__ __forceinline__ int s3_of(const Geom& g, const Coord& c) { return c.x + g.nx * (c.y + g.ny * c.z); }

// offset (in double2) of element 0 of link mu at site c; element k is k*v3 further
// (32-bit: check_dims guarantees nslots*36*v3 < 2^31)
__device__ __forceinline__ unsigned link_offset(const Geom& g, const Coord& c, int mu) {
    return (unsigned)(c.t * 36 + mu * 9) * (unsigned)g.v3 + (unsigned)s3_of(g, c);
}
// offset (in double) of coefficient 0 of momentum mu at site c; coefficient a is a*v3 further
__device__ __forceinline__ unsigned mom_offset(const Geom& g, const Coord& c, int mu) {
    return (unsigned)(c.t * 32 + mu * 8) * (unsigned)g.v3 + (unsigned)s3_of(g, c);
}
// global site id (x fastest, global t) used to key the per-site RNG streams
__device__ __forceinline__ unsigned long long global_site_id(const Geom& g, const Coord& c) {
    return (unsigned long long)s3_of(g, c) + (unsigned long long)g.v3 * (unsigned long long)(g.t0 + c.t);
}

}  // namespace gfb
