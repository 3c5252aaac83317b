// Poseidon2 phase costs (experiment): clocks per external round and per internal round when each runs alone.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkvm_prover_b200/csrc/poseidon2.cuh"
#ifndef MINB
#define MINB 2
#endif
template <int MODE>
__global__ void __launch_bounds__(256, MINB) k(uint32_t* st, uint64_t n, int reps) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    for (int j = 0; j < 16; j++) s[j] = st[16 * i + j];
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) {
#pragma unroll 1
            for (int q = 0; q < 8; q++) p2::external_round(s, P2_TAB.ext[q]);
        } else if (MODE == 1) {
#pragma unroll 1
            for (int q = 0; q < 13; q++) p2::internal_round(s, P2_TAB.in[q]);
        } else if (MODE == 2) {  // S-boxes only (16 per iteration)
#pragma unroll 1
            for (int q = 0; q < 8; q++) {
#pragma unroll
                for (int j = 0; j < 16; j++) s[j] = p2::sbox7((int32_t)bb::aadd(s[j], (uint32_t)P2_TAB.ext[q][j]));
            }
        } else {  // linear layer only
#pragma unroll 1
            for (int q = 0; q < 8; q++) p2::mds_light(s);
        }
    }
    for (int j = 0; j < 16; j++) st[16 * i + j] = s[j];
}
template <int MODE>
void run(const char* name, uint32_t* a, uint64_t n, int per_rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 16;
    k<MODE><<<(n + 255) / 256, 256>>>(a, n, reps); cudaDeviceSynchronize();
    float best = 1e9;
    for (int t = 0; t < 3; t++) {
        cudaEventRecord(e0); k<MODE><<<(n + 255) / 256, 256>>>(a, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // clocks per round per warp per SMSP: time * f * (148 * 4 SMSPs) / (warps * rounds)
    double warps = (double)n / 32, rounds = (double)reps * per_rep;
    printf("%-28s %.3f ms  -> %.1f clk per round per warp-slot (1.965 GHz)\n", name, best, best * 1e-3 * 1.965e9 * 148 * 4 / (warps * rounds));
}
int main() {
    const uint64_t n = 148ull * 2048 * 4;
    uint32_t* a; cudaMalloc(&a, n * 64);
    uint32_t* h = (uint32_t*)malloc(n * 64);
    for (uint64_t i = 0; i < n * 16; i++) h[i] = (uint32_t)((i * 2654435761ull) % bb::P);
    cudaMemcpy(a, h, n * 64, cudaMemcpyHostToDevice);
    run<0>("external rounds only", a, n, 8);
    run<1>("internal rounds only", a, n, 13);
    run<2>("16 S-boxes only", a, n, 8);
    run<3>("mds_light only", a, n, 8);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
