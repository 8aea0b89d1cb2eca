// lds_bw.cu -- shared-memory read throughput of LDS.128 at low occupancy (one CTA per SM, 8 or 16 warps), the operand path of
// the t-marching kernel: each warp reads 9-element matrices ([k][pos] layout, 8 x-consecutive lanes = 128 contiguous bytes,
// the four 8-lane groups in different rows) and folds them into registers with a few DADDs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bw lds_bw.cu && ./lds_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int NMAT, int ADDS>
__global__ void __launch_bounds__(512, 1) k_lds(double* out, int iters, long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = i * 1e-9;
    __syncthreads();
    // position of this lane inside a 64-position box; rows of 8
    const unsigned pos = (lane & 7) + 8 * ((lane >> 3) + 4 * (warp & 1));
    const unsigned stride = 64 * 16;
    double acc[ADDS > 0 ? ADDS : 1] = {0};
    long long t0 = clock64();
    unsigned base = (unsigned)__cvta_generic_to_shared(smem) + pos * 16 + (warp >> 1) * 9 * stride;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int m = 0; m < NMAT; m++) {
            double2 v[9];
#pragma unroll
            for (int k = 0; k < 9; k++)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "r"(base + (m * 9 + k) * stride));
#pragma unroll
            for (int k = 0; k < 9; k++) {
                if (ADDS > 0) { acc[k % ADDS] += v[k].x; acc[(k + 1) % ADDS] += v[k].y; }
                else acc[0] = __hiloint2double(__double2hiint(acc[0]) ^ __double2hiint(v[k].x), __double2loint(acc[0]) ^ __double2loint(v[k].y));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < (ADDS > 0 ? ADDS : 1); i++) s += acc[i];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NMAT, int ADDS>
void run(int threads, const char* name) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    auto kern = k_lds<NMAT, ADDS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    kern<<<148, threads, 200 * 1024>>>(out, iters, cyc);
    kern<<<148, threads, 200 * 1024>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double lds_per_sm = (double)iters * NMAT * 9 * (threads / 32);
    printf("%-28s threads %3d: %.2f cycles per LDS.128 per SM  (%.1f B/clk/SM)  err=%s\n", name, threads, avg / lds_per_sm, 512.0 * lds_per_sm / avg,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<2, 0>(256, "2 matrices, xor fold");
    run<2, 0>(512, "2 matrices, xor fold");
    run<2, 0>(128, "2 matrices, xor fold");
    run<2, 6>(256, "2 matrices, 18 DADD each");
    run<2, 6>(512, "2 matrices, 18 DADD each");
    run<1, 6>(256, "1 matrix, 18 DADD");
    run<4, 0>(256, "4 matrices, xor fold");
    return 0;
}
