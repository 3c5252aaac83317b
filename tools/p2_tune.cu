// Poseidon2 pipe-assignment tuning harness: times one-permutation-per-thread throughput for the knob settings
// given with -D (P2_RC_FMA, P2_MDS_FMA_MASK, P2_INT_MODE) and checks the result against the plain formulation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zkvm_prover_b200/csrc/poseidon2.cuh"
__global__ void __launch_bounds__(256) k(uint32_t* st, uint64_t n, int reps, int plain) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
    for (int j = 0; j < 16; j++) s[j] = st[16 * i + j];
    for (int r = 0; r < reps; r++) { if (plain) p2::permute_plain(s); else p2::permute(s); }
    for (int j = 0; j < 16; j++) st[16 * i + j] = s[j];
}
int main() {
    const uint64_t n = 148ull * 2048 * 4;
    uint32_t *a, *b;
    cudaMalloc(&a, n * 64); cudaMalloc(&b, n * 64);
    uint32_t* h = (uint32_t*)malloc(n * 64);
    for (uint64_t i = 0; i < n * 16; i++) h[i] = (uint32_t)((i * 2654435761ull) % bb::P);
    cudaMemcpy(a, h, n * 64, cudaMemcpyHostToDevice); cudaMemcpy(b, h, n * 64, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(a, n, 1, 0); k<<<(n + 255) / 256, 256>>>(b, n, 1, 1);
    uint32_t* h2 = (uint32_t*)malloc(n * 64);
    cudaMemcpy(h, a, n * 64, cudaMemcpyDeviceToHost); cudaMemcpy(h2, b, n * 64, cudaMemcpyDeviceToHost);
    int ok = 1; for (uint64_t i = 0; i < n * 16; i++) if (h[i] != h2[i]) { ok = 0; break; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 8;
    k<<<(n + 255) / 256, 256>>>(a, n, reps, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<(n + 255) / 256, 256>>>(a, n, reps, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("RC_FMA=%d MDS_MODE=%d INT_MODE=%d  match_plain=%d  %.3f ms  %.3f Gperm/s\n", P2_RC_FMA, P2_MDS_FMA_MASK, P2_INT_MODE, ok, ms, n * reps / (ms * 1e-3) / 1e9);
    return !ok;
}
