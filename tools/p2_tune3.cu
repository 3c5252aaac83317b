// Poseidon2 scheduling experiments (not part of the library): -DVARIANT=
//   0  one state per thread, __launch_bounds__(256)            (the shipped arrangement)
//   1  one state per thread, __launch_bounds__(256, 2)         (up to 128 registers)
//   2  one state per thread, __launch_bounds__(256, 3)
//   3  two states per thread half a permutation apart: the 8 external rounds of one (FMA-pipe heavy) are issued
//      together with the 13 internal rounds of the other (ALU-pipe heavy)
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkvm_prover_b200/csrc/poseidon2.cuh"
#ifndef VARIANT
#define VARIANT 0
#endif
#ifndef MINB
#define MINB 2
#endif
#if VARIANT == 0
#define LB __launch_bounds__(256)
#elif VARIANT == 1
#define LB __launch_bounds__(256, 2)
#elif VARIANT == 2
#define LB __launch_bounds__(256, 3)
#else
#define LB __launch_bounds__(256, MINB)
#endif
#ifndef SKEW
#define SKEW 13
#endif

#if VARIANT < 3 || VARIANT >= 5
__global__ void LB k(uint32_t* st, uint64_t n, int reps) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    for (int j = 0; j < 16; j++) s[j] = st[16 * i + j];
#if VARIANT >= 5
    // phase offset experiment: odd warps burn SKEW internal rounds on a scratch state first, so that afterwards they sit
    // in the ALU-heavy internal rounds while the even warps sit in the FMA-heavy external rounds
    if ((threadIdx.x >> 5) & 1) {
        uint32_t d[16];
        for (int j = 0; j < 16; j++) d[j] = s[j] ^ 1;
#pragma unroll 1
        for (int r = 0; r < SKEW; r++) p2::internal_round(d, P2_TAB.in[r % 13]);
        if (d[0] == 0xffffffffu) st[0] = d[1];  // never true (values are < p); keeps the loop alive
    }
#endif
    for (int r = 0; r < reps; r++) p2::permute(s);
    for (int j = 0; j < 16; j++) st[16 * i + j] = s[j];
}
#else
__device__ __forceinline__ void fused(uint32_t (&X)[16], uint32_t (&Y)[16]) {
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        if (j == 2) p2::mds_light(X);
        const int e = (4 + 2 * j) & 7;
        p2::external_round(X, P2_TAB.ext[e]);
        p2::internal_round(Y, P2_TAB.in[3 * j]);
        p2::internal_round(Y, P2_TAB.in[3 * j + 1]);
        p2::external_round(X, P2_TAB.ext[e + 1]);
        p2::internal_round(Y, P2_TAB.in[3 * j + 2]);
    }
    p2::internal_round(Y, P2_TAB.in[12]);
}
__global__ void LB k(uint32_t* st, uint64_t n, int reps) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= n) return;
    uint32_t A[16], B[16];
    for (int j = 0; j < 16; j++) { A[j] = st[16 * (2 * i) + j]; B[j] = st[16 * (2 * i + 1) + j]; }
    p2::mds_light(A);
#pragma unroll 1
    for (int r = 0; r < 4; r++) p2::external_round(A, P2_TAB.ext[r]);
#pragma unroll 1
    for (int r = 0; r < 13; r++) p2::internal_round(A, P2_TAB.in[r]);
    p2::mds_light(B);
#pragma unroll 1
    for (int r = 0; r < 4; r++) p2::external_round(B, P2_TAB.ext[r]);
#pragma unroll 1
    for (int it = 0; it < 2 * (reps - 1); it++) {
        fused(A, B);
#pragma unroll
        for (int j = 0; j < 16; j++) { uint32_t t = A[j]; A[j] = B[j]; B[j] = t; }
    }
    // reps even or odd, the swaps come in pairs: A is before its last four external rounds, B before its internal rounds
#pragma unroll 1
    for (int r = 4; r < 8; r++) p2::external_round(A, P2_TAB.ext[r]);
#pragma unroll 1
    for (int r = 0; r < 13; r++) p2::internal_round(B, P2_TAB.in[r]);
#pragma unroll 1
    for (int r = 4; r < 8; r++) p2::external_round(B, P2_TAB.ext[r]);
    for (int j = 0; j < 16; j++) { st[16 * (2 * i) + j] = A[j]; st[16 * (2 * i + 1) + j] = B[j]; }
}
#endif
__global__ void __launch_bounds__(256) kplain(uint32_t* st, uint64_t n, int reps) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    for (int j = 0; j < 16; j++) s[j] = st[16 * i + j];
    for (int r = 0; r < reps; r++) p2::permute_plain(s);
    for (int j = 0; j < 16; j++) st[16 * i + j] = s[j];
}
int main() {
    const uint64_t n = 148ull * 2048 * 4;
    const int PER = (VARIANT < 3 || VARIANT >= 5) ? 1 : 2;
    uint32_t *a, *b;
    cudaMalloc(&a, n * 64); cudaMalloc(&b, n * 64);
    uint32_t* h = (uint32_t*)malloc(n * 64);
    for (uint64_t i = 0; i < n * 16; i++) h[i] = (uint32_t)((i * 2654435761ull) % bb::P);
    cudaMemcpy(a, h, n * 64, cudaMemcpyHostToDevice); cudaMemcpy(b, h, n * 64, cudaMemcpyHostToDevice);
    const unsigned grid = (unsigned)((n / PER + 255) / 256);
    k<<<grid, 256>>>(a, n, 3); kplain<<<(n + 255) / 256, 256>>>(b, n, 3);
    uint32_t* h2 = (uint32_t*)malloc(n * 64);
    cudaMemcpy(h, a, n * 64, cudaMemcpyDeviceToHost); cudaMemcpy(h2, b, n * 64, cudaMemcpyDeviceToHost);
    int ok = 1; for (uint64_t i = 0; i < n * 16; i++) if (h[i] != h2[i]) { ok = 0; break; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 16;
    k<<<grid, 256>>>(a, n, reps); cudaDeviceSynchronize();
    float best = 1e9;
    for (int t = 0; t < 3; t++) {
        cudaEventRecord(e0); k<<<grid, 256>>>(a, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("VARIANT=%d MINB=%d match_plain=%d  %.3f ms  %.3f Gperm/s  (%s)\n", VARIANT, MINB, ok, best, n * reps / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return !ok;
}
