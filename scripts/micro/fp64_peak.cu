// Microbenchmark: achievable FP64 rate on B200 (calibrates the FP64 roof of k_force_fused).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gaugefields.jl_b200/csrc/su3.cuh"
using namespace gfb;

template <int CH>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// register-resident SU(3) products: m <- (m * c) * d^dagger, the staple inner loop without any loads
__global__ void __launch_bounds__(128, 4) k_su3(double2* out, int iters, double seed) {
    M3 m = m3_identity(), c = m3_identity(), d = m3_identity();
    c.e[1] = make_double2(seed * threadIdx.x, 1e-3); c.e[5] = make_double2(-1e-3, seed);
    d.e[2] = make_double2(seed, seed * 2); d.e[6] = make_double2(1e-4 * threadIdx.x, 0.0);
    M3 s = m3_zero();
    for (int i = 0; i < iters; i++) {
        M3 t = mul_nn(m, c);
        mac_nd(s, t, d);
        m = t;
    }
    double2 r = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < 9; k++) { r.x += s.e[k].x + m.e[k].x; r.y += s.e[k].y + m.e[k].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    double* out; cudaMalloc(&out, 148 * 32 * 1024 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        int threads = 128, blocks = 148 * warps_per_sm / 4;
        int iters = 20000;
        k_dfma<8><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * (double)blocks * threads;
        printf("dfma  ch=8  warps/SM=%2d : %.2f TFLOP/s\n", warps_per_sm, fl / ms / 1e9);
    }
    for (int ch : {1, 2, 4}) {
        int threads = 128, blocks = 148 * 4, iters = 20000;
        cudaEventRecord(e0);
        if (ch == 1) k_dfma<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        if (ch == 2) k_dfma<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        if (ch == 4) k_dfma<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * ch * iters * (double)blocks * threads;
        printf("dfma  ch=%d  warps/SM=16 : %.2f TFLOP/s  (=> DFMA latency %.1f cycles at 4 warps/SMSP)\n", ch, fl / ms / 1e9, 0.0);
    }
    for (int bps : {1, 2, 4}) {
        int blocks = 148 * bps, iters = 4000;
        k_su3<<<blocks, 128>>>((double2*)out, 10, 1e-3);
        cudaEventRecord(e0);
        k_su3<<<blocks, 128>>>((double2*)out, iters, 1e-3);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 216 * iters * (double)blocks * 128;  // 216 DFMA-class ops per iteration
        printf("su3 chain blocks/SM=%d (%2d warps/SM): %.2f TFLOP/s, %.3f ms\n", bps, bps * 4, fl / ms / 1e9, ms);
    }
    return 0;
}
