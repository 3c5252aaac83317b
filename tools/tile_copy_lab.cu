// How fast can HBM serve the strided row-tile pattern of an NTT pass, as a function of the tile shape?
// A pass over a 2^n x W row-major u32 matrix moves tiles of 2^K rows (row stride 2^L rows, L = n - K: the first pass of a
// transform; or L = 0: the last pass) x C columns; the row segment is C * 4 bytes.  Wide-and-short tiles (2^8 x 32 columns,
// 128-byte segments) need three passes for n = 23; tall-and-narrow tiles (2^12 x 8 columns, 32-byte segments = one DRAM
// sector) would need two.  This lab copies in -> out with exactly that access pattern and nothing else (one CTA per tile,
// tiles of one row set adjacent in the grid), so the number is the memory-system ceiling of the shape.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
template <int VEC>  // 16-byte units per thread access: 1 = uint4, 2 = 2 x uint4 (32 B per thread)
__global__ void __launch_bounds__(512) copy_tiles(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int K, int L, int wq /* row pitch in uint4 */, int cq /* tile columns in uint4 */,
                                                   int store_L) {
    const uint32_t col_tiles = wq / cq;
    const uint32_t ct = blockIdx.x % col_tiles;
    const uint64_t rt = blockIdx.x / col_tiles;
    const uint64_t low = rt & ((1ull << L) - 1), high = rt >> L;
    const uint64_t row_base = (high << (L + K)) + low;
    // the store side may use another stride (e.g. a pass that reads strided and writes strided elsewhere): same tile id
    const uint64_t slow = rt & ((1ull << store_L) - 1), shigh = rt >> store_L;
    const uint64_t srow_base = (shigh << (store_L + K)) + slow;
    const int units = (cq << K);
    for (int i0 = threadIdx.x; i0 < units; i0 += blockDim.x * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x;
            if (i < units) {
                const int t = i / cq, c = i % cq;
                v[u] = in[(row_base + ((uint64_t)t << L)) * wq + ct * cq + c];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x;
            if (i < units) {
                const int t = i / cq, c = i % cq;
                out[(srow_base + ((uint64_t)t << store_L)) * wq + ct * cq + c] = v[u];
            }
        }
    }
}
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 23, W = 256;
    const size_t bytes = ((size_t)W << n) * 4;
    uint4 *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int K, cols, threads; };
    const Cfg cfgs[] = {{8, 32, 256}, {8, 32, 512}, {9, 32, 512}, {10, 16, 512}, {11, 8, 512}, {12, 8, 512}, {12, 8, 1024}, {12, 4, 512}, {11, 16, 512}, {12, 16, 1024}, {10, 32, 512}, {6, 128, 256}};
    printf("matrix 2^%d x %d u32 (%.2f GB), copy in -> out tile by tile; GB/s counts read + write\n", n, W, bytes / 1e9);
    for (const Cfg& c : cfgs) {
        for (int pat = 0; pat < 3; pat++) {  // 0: strided load + strided store (first pass), 1: contiguous both (last pass), 2: contiguous load, strided store (scatter)
            const int L = n - c.K;
            const int lL = pat == 0 ? L : 0, sL = pat == 1 ? 0 : L;
            const int wq = W / 4, cq = c.cols / 4;
            const uint32_t grid = (uint32_t)(((size_t)1 << (n - c.K)) * (wq / cq));
            float best = 1e9;
            for (int t = 0; t < 4; t++) {
                cudaEventRecord(e0);
                copy_tiles<1><<<grid, c.threads>>>(a, b, n, c.K, lL, wq, cq, sL);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (t && ms < best) best = ms;
            }
            printf("tile 2^%-2d rows x %3d cols (%3d B segments, %3d KB) %4d thr  %-28s %7.3f ms  %6.0f GB/s\n", c.K, c.cols, c.cols * 4, (c.cols * 4 << c.K) >> 10, c.threads,
                   pat == 0 ? "strided -> strided" : (pat == 1 ? "contiguous -> contiguous" : "contiguous -> strided"), best, 2.0 * bytes / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
