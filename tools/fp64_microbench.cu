// Is the FP64 pipe of B200 a usable third integer-ish pipe?  Measures DFMA throughput alone and next to IMAD / IADD3 streams.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, uint32_t* iout, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 0.5;
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, y = blockIdx.x | 1;
    for (int i = 0; i < iters; i++) {
#pragma unroll 8
        for (int u = 0; u < 8; u++) {
            if (MODE == 0 || MODE == 2 || MODE == 3) {  // 8 independent DFMA
                a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
                a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
            }
            if (MODE == 1 || MODE == 2) {  // 8 IMAD (32-bit)
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x0) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x1) : "r"(y));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x2) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x3) : "r"(y));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x0) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x1) : "r"(y));
                asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x2) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x3) : "r"(y));
            }
            if (MODE == 3) {  // 8 VIADDMNMX-class (min of add)
                x0 = min(x0 + y, x1); x1 = min(x1 + y, x2); x2 = min(x2 + y, x3); x3 = min(x3 + y, x0);
                x0 = min(x0 + y, x2); x1 = min(x1 + y, x3); x2 = min(x2 + y, x0); x3 = min(x3 + y, x1);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
}
template <int MODE>
void run(const char* name, double* d, uint32_t* di, int per_iter) {
    const int iters = 2000, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, di, iters, 1.5); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, di, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * 8 * iters * 8 * per_iter;     // warps x iterations x unroll x instructions
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s %.3f ms  %.2f clk per warp-instruction per SMSP\n", name, ms, clk * 148 * 4 / warp_instr);
}
int main() {
    double* d; uint32_t* di;
    cudaMalloc(&d, 148 * 8 * 256 * 8); cudaMalloc(&di, 148 * 8 * 256 * 4);
    run<0>("DFMA alone", d, di, 8);
    run<1>("IMAD alone", d, di, 8);
    run<2>("DFMA + IMAD (1:1)", d, di, 16);
    run<3>("DFMA + add/min (1:1)", d, di, 16);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
