// Microbenchmark: L2->SM bandwidth for the fused kernel's access pattern -- every warp gathers 512-byte chunks from many
// planes that are a fixed stride apart (structure-of-arrays link planes) -- against the same volume read contiguously.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) k_gather(const double2* __restrict__ p, size_t plane_elems, int nplanes, int rows, int reps, double2* out) {
    double2 acc = make_double2(0, 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int r = 0; r < reps; r++)
        for (int row = blockIdx.x * 4 + w; row < rows; row += gridDim.x * 4) {
            const double2* q = p + (size_t)row * 32 + lane;
#pragma unroll 9
            for (int k = 0; k < nplanes; k++) {
                double2 v = __ldg(q + (size_t)k * plane_elems);
                acc.x += v.x; acc.y += v.y;
            }
        }
    if (acc.x == 1.2345) out[0] = acc;
}
int main() {
    double2 *p, *out; size_t maxb = (size_t)2 << 30;
    cudaMalloc(&p, maxb); cudaMalloc(&out, 1024); cudaMemset(p, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    // total bytes = nplanes * rows * 512
    struct Cfg { const char* name; size_t plane_elems; int nplanes; int rows; };
    Cfg cfgs[] = {
        {"36 planes x 512 KB stride (32^3 slice), 18.9 MB", 32768, 36, 1024},
        {"108 planes x 512 KB stride (3 slices), 56.6 MB", 32768, 108, 1024},
        {"36 planes x 4 MB stride (64^3 slice, 64 MB of 151 MB)", 262144, 36, 3456},
        {"36 planes, contiguous rows (plane = rows*32), 18.9 MB", 1024 * 32, 36, 1024},
        {"1 plane contiguous 18.9 MB", 0, 1, 36864},
        {"36 planes x (512 KB + 512 B) stride", 32768 + 32, 36, 1024},
        {"36 planes x (512 KB + 8 KB) stride", 32768 + 512, 36, 1024},
    };
    for (auto& c : cfgs) {
        for (int bps : {4, 8, 16}) {
            int reps = 20;
            k_gather<<<148 * bps, 128>>>(p, c.plane_elems, c.nplanes, c.rows, 2, out);
            cudaEventRecord(e0);
            k_gather<<<148 * bps, 128>>>(p, c.plane_elems, c.nplanes, c.rows, reps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("%-58s blocks/SM %2d : %6.2f TB/s\n", c.name, bps, (double)c.nplanes * c.rows * 512.0 * reps / ms / 1e9);
        }
    }
    return 0;
}
