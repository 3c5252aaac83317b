// Which IADD3 operand forms issue at the full ALU rate on sm_100a?  (round 3 of the pipe microbenchmarks)
//   mode 0: IADD3 R, R, R, RZ       (ptxas is free to pick IMAD.IADD instead -- check the SASS)
//   mode 1: IADD3 R, R, UR, R       (third addend an opaque zero held in a UNIFORM register: what bb::aadd compiles to)
//   mode 2: IADD3 R, R, R, R        (third addend an opaque zero held in a VECTOR register)
//   mode 3..5: the same three forms interleaved 1:1 with VIADDMNMX (a canonical add = IADD3 + VIADDMNMX)
//   mode 6..8: the same three forms inside a Montgomery-product-shaped group: IMAD.WIDE, IMAD, IMAD.HI, IADD3(form), VIADDMNMX
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __constant__ uint32_t K_ZERO = 0u;
constexpr uint32_t P = 0x78000001u, MU = 0x88000001u;
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* zeros, int reps) {
    uint32_t x[8];
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 8 + j + blockIdx.x;
    const uint32_t zu = K_ZERO;
    const uint32_t zv = zeros[threadIdx.x];
    uint32_t y = zeros[threadIdx.x + 256] + 12345u;
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                constexpr int F = MODE % 3;
                if (MODE >= 6) {
                    int64_t t = (int64_t)(int32_t)x[j] * (int64_t)(int32_t)y;
                    int32_t q = (int32_t)((uint32_t)t * MU);
                    int32_t qh = __mulhi(q, (int32_t)P);
                    uint32_t d = F == 0 ? (uint32_t)(t >> 32) - (uint32_t)qh : (F == 1 ? (uint32_t)(t >> 32) - (uint32_t)qh + zu : (uint32_t)(t >> 32) - (uint32_t)qh + zv);
                    x[j] = min(d, d + P);
                } else {
                    const uint32_t o = x[(j + 3) & 7];
                    uint32_t s = F == 0 ? x[j] + o : (F == 1 ? x[j] + o + zu : x[j] + o + zv);
                    if (MODE >= 3) s = min(s, s - P);
                    x[j] = s;
                }
            }
        }
    }
    uint32_t a = 0;
    for (int j = 0; j < 8; j++) a ^= x[j];
    out[blockIdx.x * 256 + threadIdx.x] = a;
}
template <int MODE>
void run(const char* name, uint32_t* out, uint32_t* zeros, int per_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 512, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, zeros, reps); cudaDeviceSynchronize();
    float best = 1e9;
    for (int t = 0; t < 3; t++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, zeros, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double groups = (double)blocks * 8 /*warps*/ * reps * 64;  // warp-groups executed
    const double clk = best * 1e-3 * 1.965e9 * 148 * 4 / groups;
    printf("mode %d %-58s %.3f ms  %.2f clk per warp-group per SMSP (%d instr -> %.2f clk/instr)\n", MODE, name, best, clk, per_iter, clk / per_iter);
}
int main() {
    uint32_t *out, *zeros;
    cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&zeros, 4096); cudaMemset(zeros, 0, 4096);
    run<0>("add R,R", out, zeros, 1);
    run<1>("IADD3 R,R,UR,R (uniform opaque zero)", out, zeros, 1);
    run<2>("IADD3 R,R,R,R (vector opaque zero)", out, zeros, 1);
    run<3>("add R,R + VIADDMNMX", out, zeros, 2);
    run<4>("IADD3 R,R,UR,R + VIADDMNMX", out, zeros, 2);
    run<5>("IADD3 R,R,R,R + VIADDMNMX", out, zeros, 2);
    run<6>("IMAD.WIDE, IMAD, IMAD.HI, sub R,R, VIADDMNMX", out, zeros, 5);
    run<7>("IMAD.WIDE, IMAD, IMAD.HI, IADD3 -R,R,UR, VIADDMNMX", out, zeros, 5);
    run<8>("IMAD.WIDE, IMAD, IMAD.HI, IADD3 -R,R,R, VIADDMNMX", out, zeros, 5);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
