// Microbenchmark: does a non-FP64 instruction cost FP64-pipe time on B200?  Streams of DFMA (16 independent chains per
// thread) interleaved with K independent integer IMADs and L LDS.128 per 16 DFMA, at 1 and 2 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int L>
__global__ void __launch_bounds__(256, 1) k_mix(double* out, int iters, double a, double b, int m) {
    __shared__ double2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i * 1e-9, 1.0);
    __syncthreads();
    double acc[16];
    int ia[8];
#pragma unroll
    for (int c = 0; c < 16; c++) acc[c] = threadIdx.x + c;
#pragma unroll
    for (int c = 0; c < 8; c++) ia[c] = threadIdx.x * (c + 1);
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 255) * 16;
    double2 ld[L > 0 ? L : 1];
#pragma unroll
    for (int l = 0; l < (L > 0 ? L : 1); l++) ld[l] = make_double2(0, 0);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int l = 0; l < L; l++) {
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr + (((l + rep * L) & 7) * 4096) % 28672));
                ld[l] = v;
            }
#pragma unroll
            for (int c = 0; c < 16; c++) {
                acc[c] = fma(acc[c], a, (L > 0 && c < L) ? ld[c].x : b);
                if (c * K / 16 != (c + 1) * K / 16) {
                    const int j = (c * K / 16) & 7;
                    ia[j] = ia[j] * m + (int)i;
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) s += acc[c];
    int t = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) t ^= ia[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int K, int L>
void run(double* out, int warps_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int threads = 32 * warps_per_sm, blocks = 148, iters = 20000;
    k_mix<K, L><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9, 3);
    cudaEventRecord(e0);
    k_mix<K, L><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 64.0 * iters * blocks * threads / 32;  // warp instructions
    const double cyc_per_dfma_per_sched = ms * 1e-3 * 1.9e9 / (dfma / (148 * 4));
    printf("K=%2d IMAD, L=%d LDS.128 per 16 DFMA, %d warps/SM: %.3f ms, %.2f TFLOP/s, %.2f cycles per DFMA per scheduler (at 1.9 GHz)\n", K, L, warps_per_sm, ms,
           2.0 * 32 * dfma / ms / 1e9, cyc_per_dfma_per_sched);
}

int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    for (int w : {4, 8}) {
        run<0, 0>(out, w); run<4, 0>(out, w); run<8, 0>(out, w); run<16, 0>(out, w);
        run<0, 1>(out, w); run<0, 2>(out, w); run<0, 4>(out, w); run<8, 2>(out, w); run<8, 4>(out, w);
    }
    return 0;
}
