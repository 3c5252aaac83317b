// Butterfly pipe microbenchmark (round 2): the exact arithmetic of ntt::radix_round's Shoup butterfly on registers only (no
// shared memory, no barriers), to separate "instruction mix cannot dual-issue" from "shared memory / barriers / latency".
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/bfly_microbench tools/bfly_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkvm_prover_b200/csrc/bb31.cuh"
#ifndef NV
#define NV 32  // values per thread (8 rows x 4 columns)
#endif
#ifndef CTAS
#define CTAS 2
#endif
template <int MODE>
__global__ void __launch_bounds__(256, CTAS) k(uint32_t* io, const uint2* tw, int iters) {
    uint32_t x[NV];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NV; i++) x[i] = io[(size_t)i * gridDim.x * blockDim.x + tid];
    uint2 w[7];
#pragma unroll
    for (int i = 0; i < 7; i++) w[i] = tw[(threadIdx.x & 31) * 7 + i];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        // radix-8 on 8 "rows" x (NV/8) columns: 3 stages, 12 butterflies per column
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const int half = 4 >> v;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (q & half) continue;
                const uint2 ww = w[(v == 0 ? 0 : (v == 1 ? 4 : 6)) + (v == 2 ? 0 : (q & (half - 1)))];
#pragma unroll
                for (int c = 0; c < NV / 8; c++) {
                    const uint32_t a = x[q * (NV / 8) + c], b = x[(q + half) * (NV / 8) + c];
                    if (MODE == 0) {  // full butterfly
                        x[q * (NV / 8) + c] = bb::add(a, b);
                        const uint32_t d = a - b + bb::P;
                        const uint32_t qq = __umulhi(d, ww.y);
                        x[(q + half) * (NV / 8) + c] = bb::red2p(d * ww.x - qq * bb::P);
                    } else if (MODE == 1) {  // FMA-pipe part only (3 instructions)
                        const uint32_t qq = __umulhi(b, ww.y);
                        x[(q + half) * (NV / 8) + c] = b * ww.x - qq * bb::P + a;
                    } else if (MODE == 2) {  // ALU part only (4 instructions)
                        x[q * (NV / 8) + c] = bb::add(a, b);
                        const uint32_t d = a - b + bb::P;
                        x[(q + half) * (NV / 8) + c] = bb::red2p(d);
                    } else if (MODE == 3) {  // Montgomery butterfly: IMAD.WIDE + IMAD + IMAD.HI
                        x[q * (NV / 8) + c] = bb::add(a, b);
                        x[(q + half) * (NV / 8) + c] = bb::mul(bb::sub(a, b), ww.x);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; i++) io[(size_t)i * gridDim.x * blockDim.x + tid] = x[i];
}
template <int MODE>
void run(const char* name, uint32_t* io, uint2* tw, int sms) {
    const int grid = sms * CTAS, iters = 2000;
    k<MODE><<<grid, 256>>>(io, tw, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(io, tw, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // warp-butterflies per SMSP: warps/SMSP * iters * 12 * NV/8
    const double wb = (CTAS * 8.0 / 4) * iters * 12.0 * (NV / 8);
    printf("%-44s %8.3f ms  %6.2f clk per warp-butterfly per SMSP (%d warps/SMSP)\n", name, best, best * 1e-3 * 1.965e9 / wb, CTAS * 2);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* io; uint2* tw;
    cudaMalloc(&io, (size_t)NV * sms * CTAS * 256 * 4); cudaMalloc(&tw, 32 * 7 * 8);
    cudaMemset(io, 0x11, (size_t)NV * sms * CTAS * 256 * 4); cudaMemset(tw, 0x23, 32 * 7 * 8);
    run<0>("Shoup butterfly (3 FMA + 4 ALU instr)", io, tw, sms);
    run<1>("FMA part only (IMAD.HI + 2 IMAD)", io, tw, sms);
    run<2>("ALU part only (2 IADD3 + 2 VIADDMNMX)", io, tw, sms);
    run<3>("Montgomery butterfly", io, tw, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
