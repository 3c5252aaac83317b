// K2: batched BabyBear NTT over the rows of a row-major matrix (every column is one transform).
// Replaces p3_dft::Radix2DitParallel<BabyBear> behind TwoAdicSubgroupDft (p3-dft 0.4.3, Cargo.lock:5590
// of the reference; not vendored).  Results are field-exact, so any correct decomposition is bit-identical.
//
// Decomposition (validated against the oracle's NaiveDft): an n-stage radix-2 decimation-in-frequency
// transform (natural order in, bit-reversed order out) is cut into passes of K <= 10 stages.  A pass owns
// tiles of 2^K rows x 32 columns: rows  high*2^(n-s0) + t*2^L + low  (t = 0..2^K-1, L = n-s0-K).  Inside a
// tile the pass is a plain 2^K-point DIF with the local roots w_{2^K}^i, followed by one "twist" multiply
// of slot t' by w_{2^(n-s0)}^(low * bitrev_K(t')) -- the four-step regrouping of the merged twiddles.
// A tile lives in shared memory (64 KB for K=9: two CTAs per SM, so one CTA's global loads overlap the
// other's butterflies); a thread keeps 8 rows x 4 adjacent columns in registers and does three stages
// per shared-memory round trip; the twiddle is shared by the 4 columns (128-bit accesses everywhere).
// Rows are >= 128 B contiguous segments, so the strided row gathers of the early passes and the
// bit-reversal scatter stay fully coalesced.
#pragma once
#include "bb31.cuh"

namespace ntt {

constexpr int TILE_COLS = 32;
constexpr int THREADS = 512;
constexpr int LO_BITS = 12;  // two-level power tables: x^e = hi[e >> 12] * lo[e & 4095]

struct PassParams {
    const uint32_t* in;
    uint32_t* out;
    uint32_t width;       // columns (row pitch, elements)
    int n;                // log2 of the transform size
    int s0;               // first DIF stage done by this pass
    int K;                // stages in this pass (tile = 2^K rows)
    int inverse;          // use inverse roots
    const uint32_t* tw_local;  // 2^(K-1): w_{2^K}^(+-i)
    const uint32_t* tw_lo;     // w_N^i, i < 2^min(n,12)      (forward roots; inverse uses N - e)
    const uint32_t* tw_hi;     // w_N^(i << 12), i < 2^max(n-12,0)
    const uint32_t* pre_lo;    // optional: multiply input row j by pre_hi[j >> 12] * pre_lo[j & 4095]
    const uint32_t* pre_hi;
    const uint32_t* post_lo;   // optional (last pass only): multiply logical output index j likewise
    const uint32_t* post_hi;
    int out_natural;      // last pass only: write logical index j = bitrev_n(position) to row j
};

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    uint32_t v[4];
    __device__ __forceinline__ void load(const uint32_t* p) { uint4 t = *reinterpret_cast<const uint4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ __forceinline__ void store(uint32_t* p) const { *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<1> {
    uint32_t v[1];
    __device__ __forceinline__ void load(const uint32_t* p) { v[0] = *p; }
    __device__ __forceinline__ void store(uint32_t* p) const { *p = v[0]; }
};

__device__ __forceinline__ uint32_t pow2level(const uint32_t* lo, const uint32_t* hi, uint64_t e) {
    uint32_t l = __ldg(lo + (e & ((1u << LO_BITS) - 1)));
    return bb::mul(l, __ldg(hi + (e >> LO_BITS)));  // hi[] may carry a scale factor, so always multiply
}

// k stages (radix 2^k) at local stage u of the 2^K-point DIF, on registers; sm = tile [2^K][TILE_COLS]
template <int k, int VEC>
__device__ __forceinline__ void radix_round(uint32_t* sm, const uint32_t* sm_tw, int K, int u, int tid) {
    constexpr int LANES = TILE_COLS / VEC;  // threads per row
    constexpr int R = 1 << k;
    const int lowbits = K - u - k;
    const int groups = (1 << (K - k)) * LANES;
    for (int gi = tid; gi < groups; gi += THREADS) {
        const int lane = gi % LANES;
        const int g = gi / LANES;
        const int highpart = g >> lowbits, lowpart = g & ((1 << lowbits) - 1);
        const int base = (highpart << (K - u)) + lowpart;
        Vec<VEC> x[R];
#pragma unroll
        for (int q = 0; q < R; q++) x[q].load(sm + (base + (q << lowbits)) * TILE_COLS + lane * VEC);
#pragma unroll
        for (int v = 0; v < k; v++) {
            const int half = 1 << (k - 1 - v);
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (q & half) continue;
                const int e = (((q & (half - 1)) << lowbits) + lowpart) << (u + v);
                const uint32_t w = sm_tw[e];
#pragma unroll
                for (int c = 0; c < VEC; c++) {
                    uint32_t a = x[q].v[c], b = x[q + half].v[c];
                    x[q].v[c] = bb::add(a, b);
                    x[q + half].v[c] = bb::canon(bb::smul((int32_t)(a - b), (int32_t)w));  // a-b in (-p,p) as signed
                }
            }
        }
#pragma unroll
        for (int q = 0; q < R; q++) x[q].store(sm + (base + (q << lowbits)) * TILE_COLS + lane * VEC);
    }
}

template <int VEC>
__global__ void __launch_bounds__(THREADS, 2) pass_kernel(const PassParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int LANES = TILE_COLS / VEC;
    const int K = p.K, n = p.n, s0 = p.s0;
    const int L = n - s0 - K;
    const int R = 1 << K;
    uint32_t* sm = smem;                           // [R][TILE_COLS]
    uint32_t* sm_tw = smem + R * TILE_COLS;        // [R/2] local roots
    uint32_t* sm_row = sm_tw + (R > 1 ? R / 2 : 1);  // [R] per-row factor (prescale, then twist/postscale)
    const int tid = threadIdx.x;

    const uint32_t col_tiles = (p.width + TILE_COLS - 1) / TILE_COLS;
    const uint32_t ct = blockIdx.x % col_tiles;
    const uint64_t rt = blockIdx.x / col_tiles;    // row tile: (high, low)
    const uint64_t low = rt & ((1ull << L) - 1);
    const uint64_t high = rt >> L;
    const uint64_t row_base = (high << (n - s0)) + low;  // row of slot t = row_base + (t << L)
    const uint32_t col0 = ct * TILE_COLS;

    for (int i = tid; i < R / 2; i += THREADS) sm_tw[i] = __ldg(p.tw_local + i);
    if (p.pre_lo)
        for (int t = tid; t < R; t += THREADS) sm_row[t] = pow2level(p.pre_lo, p.pre_hi, row_base + ((uint64_t)t << L));
    if (p.pre_lo) __syncthreads();

    // ---- load tile (each row segment is 128 B contiguous)
    for (int i = tid; i < R * LANES; i += THREADS) {
        const int t = i / LANES, lane = i % LANES;
        const uint32_t col = col0 + lane * VEC;
        Vec<VEC> x;
        if (col < p.width) {
            x.load(p.in + (row_base + ((uint64_t)t << L)) * p.width + col);
            if (p.pre_lo) {
                const uint32_t f = sm_row[t];
#pragma unroll
                for (int c = 0; c < VEC; c++) x.v[c] = bb::mul(x.v[c], f);
            }
        } else {
#pragma unroll
            for (int c = 0; c < VEC; c++) x.v[c] = 0;
        }
        x.store(sm + t * TILE_COLS + lane * VEC);
    }
    __syncthreads();

    // ---- per-row output factor: twist w_B^(+-low*bitrev(t')) (L > 0) and/or post scale (last pass)
    const bool need_factor = (L > 0 && low != 0) || p.post_lo;
    if (need_factor) {
        for (int t = tid; t < R; t += THREADS) {
            uint32_t f = bb::ONE;
            if (L > 0 && low != 0) {
                uint64_t e = (low * (uint64_t)bb::bitrev((uint32_t)t, K)) << s0;  // exponent of w_N, < N
                if (p.inverse) e = ((1ull << n) - e) & ((1ull << n) - 1);
                f = pow2level(p.tw_lo, p.tw_hi, e);
            }
            if (p.post_lo) {
                const uint64_t pos = row_base + ((uint64_t)t << L);
                const uint64_t j = (uint64_t)bb::bitrev((uint32_t)pos, n);
                uint32_t g = pow2level(p.post_lo, p.post_hi, j);
                f = (L > 0 && low != 0) ? bb::mul(f, g) : g;
            }
            sm_row[t] = f;
        }
    }

    // ---- K stages: rounds of 3 (radix 8), remainder first
    int u = 0;
    const int rem = K % 3;
    if (rem == 1) { radix_round<1, VEC>(sm, sm_tw, K, u, tid); u += 1; __syncthreads(); }
    if (rem == 2) { radix_round<2, VEC>(sm, sm_tw, K, u, tid); u += 2; __syncthreads(); }
    for (; u < K; u += 3) { radix_round<3, VEC>(sm, sm_tw, K, u, tid); __syncthreads(); }

    // ---- store
    for (int i = tid; i < R * LANES; i += THREADS) {
        const int t = i / LANES, lane = i % LANES;
        const uint32_t col = col0 + lane * VEC;
        if (col >= p.width) continue;
        Vec<VEC> x;
        x.load(sm + t * TILE_COLS + lane * VEC);
        if (need_factor) {
            const uint32_t f = sm_row[t];
#pragma unroll
            for (int c = 0; c < VEC; c++) x.v[c] = bb::mul(x.v[c], f);
        }
        uint64_t row = row_base + ((uint64_t)t << L);
        if (p.out_natural) row = bb::bitrev((uint32_t)row, n);
        x.store(p.out + row * p.width + col);
    }
}

// power tables: lo[i] = base^i (i < 2^12), hi[i] = scale * base^(i << 12) (i < n_hi)
__global__ void pow_table_kernel(uint32_t* lo, uint32_t* hi, uint32_t base, uint32_t scale, uint32_t n_lo, uint32_t n_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lo) lo[i] = bb::pow(base, i);
    if (i < n_hi) hi[i] = bb::mul(scale, bb::pow(base, (uint64_t)i << LO_BITS));
}

// out[bitrev(i)] = in[i] (row permutation, out of place)
template <int VEC>
__global__ void bitrev_rows_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n, uint32_t width) {
    const uint32_t lanes = (width + VEC - 1) / VEC;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t row = i / lanes;
    uint32_t col = (uint32_t)(i % lanes) * VEC;
    if (row >> n) return;
    Vec<VEC> x;
    x.load(in + row * width + col);
    x.store(out + (uint64_t)bb::bitrev((uint32_t)row, n) * width + col);
}

}  // namespace ntt
