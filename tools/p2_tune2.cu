// Poseidon2 harness, variant: TWO independent states per thread advanced in lockstep inside the same round loops.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkvm_prover_b200/csrc/poseidon2.cuh"
__device__ __forceinline__ void permute2(uint32_t (&a)[16], uint32_t (&b)[16]) {
    p2::mds_light(a); p2::mds_light(b);
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) { p2::external_round(a, P2_TAB.ext[4 * half + r]); p2::external_round(b, P2_TAB.ext[4 * half + r]); }
        if (half == 0) {
#pragma unroll 1
            for (int r = 0; r < 13; r++) { p2::internal_round(a, P2_TAB.in[r]); p2::internal_round(b, P2_TAB.in[r]); }
        }
    }
}
template <int TWO>
__global__ void __launch_bounds__(256) k(uint32_t* st, uint64_t n, int reps) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (TWO) {
        if (2 * i + 1 >= n) return;
        uint32_t a[16], b[16];
        for (int j = 0; j < 16; j++) { a[j] = st[16 * (2 * i) + j]; b[j] = st[16 * (2 * i + 1) + j]; }
        for (int r = 0; r < reps; r++) permute2(a, b);
        for (int j = 0; j < 16; j++) { st[16 * (2 * i) + j] = a[j]; st[16 * (2 * i + 1) + j] = b[j]; }
    } else {
        if (i >= n) return;
        uint32_t s[16];
        for (int j = 0; j < 16; j++) s[j] = st[16 * i + j];
        for (int r = 0; r < reps; r++) p2::permute(s);
        for (int j = 0; j < 16; j++) st[16 * i + j] = s[j];
    }
}
int main() {
    const uint64_t n = 148ull * 2048 * 4;
    uint32_t *a, *b;
    cudaMalloc(&a, n * 64); cudaMalloc(&b, n * 64);
    uint32_t* h = (uint32_t*)malloc(n * 64); uint32_t* h2 = (uint32_t*)malloc(n * 64);
    for (uint64_t i = 0; i < n * 16; i++) h[i] = (uint32_t)((i * 2654435761ull) % bb::P);
    cudaMemcpy(a, h, n * 64, cudaMemcpyHostToDevice); cudaMemcpy(b, h, n * 64, cudaMemcpyHostToDevice);
    k<0><<<(n + 255) / 256, 256>>>(a, n, 1); k<1><<<(n / 2 + 255) / 256, 256>>>(b, n, 1);
    cudaMemcpy(h, a, n * 64, cudaMemcpyDeviceToHost); cudaMemcpy(h2, b, n * 64, cudaMemcpyDeviceToHost);
    int ok = 1; for (uint64_t i = 0; i < n * 16; i++) if (h[i] != h2[i]) { ok = 0; break; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 8; float ms;
    k<0><<<(n + 255) / 256, 256>>>(a, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<0><<<(n + 255) / 256, 256>>>(a, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("one state / thread : %.3f ms  %.3f Gperm/s\n", ms, n * reps / (ms * 1e-3) / 1e9);
    k<1><<<(n / 2 + 255) / 256, 256>>>(b, n, reps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<1><<<(n / 2 + 255) / 256, 256>>>(b, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("two states / thread: %.3f ms  %.3f Gperm/s   match=%d\n", ms, n * reps / (ms * 1e-3) / 1e9, ok);
    return !ok;
}
