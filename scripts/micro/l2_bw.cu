// Microbenchmark: L2->SM read bandwidth on B200 (is k_force_fused bound by the L2 crossbar?).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ p, size_t n, int reps, double2* out) {
    double2 acc = make_double2(0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            double2 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
            acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
        }
    }
    if (acc.x == 1.2345) out[0] = acc;
}
int main() {
    double2 *p, *out; size_t maxb = (size_t)4 << 30;
    cudaMalloc(&p, maxb); cudaMalloc(&out, 1024); cudaMemset(p, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096}) {
        size_t n = mb * 1024 * 1024 / 16;
        int reps = (int)(16384 / mb) + 1;
        for (int bps : {4, 8}) {
            k_read<<<148 * bps, 256>>>(p, n, 1, out);
            cudaEventRecord(e0);
            k_read<<<148 * bps, 256>>>(p, n, reps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("buffer %5zu MB  blocks/SM %d : %.2f TB/s\n", mb, bps, (double)n * 16 * reps / ms / 1e9);
        }
    }
    return 0;
}
