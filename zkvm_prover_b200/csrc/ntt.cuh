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

constexpr int TILE_ELEMS_LOG = 14;  // a tile is 2^14 elements (64 KB): 2^K rows x 2^(14-K) columns, so all 512 threads have a radix-8 group
constexpr int MAX_TILE_COLS_LOG = 7;  // at most 128 columns (512 B row segments)
#ifndef NTT_THREADS
#define NTT_THREADS 512
#endif
#ifndef NTT_MIN_CTAS
#define NTT_MIN_CTAS 2
#endif
constexpr int THREADS = NTT_THREADS;
constexpr int LO_BITS = 12;  // two-level power tables: x^e = hi[e >> 12] * lo[e & 4095]

struct PassParams {
    const uint32_t* in;
    uint32_t* out;
    uint32_t width;       // columns processed
    uint32_t in_pitch;    // row pitch of `in` in elements (>= width; a column strip of a wider matrix has pitch > width)
    uint32_t out_pitch;   // row pitch of `out`
    int n;                // log2 of the transform size
    int s0;               // first DIF stage done by this pass
    int K;                // stages in this pass (tile = 2^K rows)
    int lc;               // log2 of the tile's column count (tile = 2^K rows x 2^lc columns)
    int inverse;          // use inverse roots
    const uint2* tw_local;     // 2^(K-1): (w, floor(w * 2^32 / p)) with w = w_{2^K}^(+-i) as a PLAIN integer (Shoup pairs)
    const uint32_t* tw_lo;     // w_N^i, i < 2^min(n,12)      (forward roots; inverse uses N - e)
    const uint32_t* tw_hi;     // w_N^(i << 12), i < 2^max(n-12,0)
    const uint32_t* pre_lo;    // optional: multiply input row j by pre_hi[j >> 12] * pre_lo[j & 4095]
    const uint32_t* pre_hi;
    const uint32_t* post_lo;   // optional (last pass only): multiply logical output index j likewise
    const uint32_t* post_hi;
    int out_natural;      // last pass only: write logical index j = bitrev_n(position) to row j
    uint32_t rt_base;     // TMA kernel: first row tile of this launch (a pass may be issued in several launches, one per destination)
    uint32_t prefetch_dist;  // CTAs resident at once: each CTA prefetches (to L2) the tile of block blockIdx.x + prefetch_dist
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

// Shoup product with a precomputed pair (w, w' = floor(w 2^32 / p)), w a PLAIN integer < p: d*w mod p for any 32-bit d, so
// Montgomery-form data stays in Montgomery form.  IMAD.HI + 2 IMAD (8 FMA-pipe clocks, a Montgomery product needs 10) and
// one ALU instruction (Montgomery: two).
__device__ __forceinline__ uint32_t shoup_mul(uint32_t d, uint2 w) {
    const uint32_t q = __umulhi(d, w.y);
    return bb::red2p(d * w.x - q * bb::P);
}
// the pair for a factor given in Montgomery form f = w 2^32 mod p:  w 2^32 = w' p + f  =>  w' = -f p^-1 (mod 2^32)
__device__ __forceinline__ uint2 shoup_pair(uint32_t f_monty) { return make_uint2(bb::from_monty(f_monty), (0u - f_monty) * bb::MU); }

// k stages (radix 2^k) at local stage u of the 2^K-point DIF, on registers; sm = tile [2^K][TILE_COLS]
// LAST: this round ends the tile's DIF (lowbits == 0), so the twiddle index of a butterfly depends only on q and the
// butterflies with (q & (half-1)) == 0 multiply by w^0 = 1: they are done as plain subtractions (7 of the 12 butterflies
// of a radix-8 round).
// BREV_OUT (final round of a scatter pass, TMA kernel only): row t' is written to slot bitrev_K(t') so the tile leaves in
// natural coefficient order; every thread has exactly one group then, and a named barrier separates loads from stores.
template <int k, int VEC, bool LAST, int NT = THREADS, bool BREV_OUT = false>
__device__ __forceinline__ void radix_round(uint32_t* sm, const uint2* sm_tw, int K, int lc, int u, int tid, const uint2* fac_pre = nullptr,
                                            const uint2* fac_post = nullptr) {
    constexpr int LV = VEC == 4 ? 2 : 0;
    const int ll = lc - LV;                 // log2(threads per row)
    const int TILE_COLS = 1 << lc;
    constexpr int R = 1 << k;
    const int lowbits = K - u - k;
    const int groups = 1 << (K - k + ll);
    for (int gi = tid; gi < groups; gi += NT) {
        const int lane = gi & ((1 << ll) - 1);
        const int g = gi >> ll;
        const int highpart = g >> lowbits, lowpart = g & ((1 << lowbits) - 1);
        const int base = (highpart << (K - u)) + lowpart;
        Vec<VEC> x[R];
#pragma unroll
        for (int q = 0; q < R; q++) x[q].load(sm + (base + (q << lowbits)) * TILE_COLS + lane * VEC);
        if (fac_pre) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = fac_pre[base + (q << lowbits)];
#pragma unroll
                for (int c = 0; c < VEC; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
#pragma unroll
        for (int v = 0; v < k; v++) {
            const int half = 1 << (k - 1 - v);
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (q & half) continue;
                if (LAST && (q & (half - 1)) == 0) {
#pragma unroll
                    for (int c = 0; c < VEC; c++) {
                        uint32_t a = x[q].v[c], b = x[q + half].v[c];
                        x[q].v[c] = bb::add(a, b);
                        x[q + half].v[c] = bb::sub(a, b);
                    }
                    continue;
                }
                const int e = (((q & (half - 1)) << lowbits) + lowpart) << (u + v);
                // Shoup product with the precomputed quotient w' = floor(w 2^32 / p): d*w mod p = d*w - hi(d*w') * p, in [0, 2p).
                // w is the plain integer value of the root, so Montgomery-form data stays in Montgomery form.
                // IMAD.HI + 2 IMAD = 8 FMA-pipe clocks (a Montgomery product needs 10) and one ALU instruction less.
                const uint2 w = sm_tw[e];
#pragma unroll
                for (int c = 0; c < VEC; c++) {
                    const uint32_t a = x[q].v[c], b = x[q + half].v[c];
                    x[q].v[c] = bb::add(a, b);
                    const uint32_t d = a - b + bb::P;                 // in (0, 2p)
                    const uint32_t qq = __umulhi(d, w.y);
                    x[q + half].v[c] = bb::red2p(d * w.x - qq * bb::P);
                }
            }
        }
        if (fac_post) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = fac_post[base + (q << lowbits)];
#pragma unroll
                for (int c = 0; c < VEC; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
        if (BREV_OUT) {
            // all loads of the round are done before any permuted store; the non-.aligned form because a narrow tile leaves
            // part of a warp without a group (those lanes arrive from the statement after the loop)
            asm volatile("barrier.sync 1, %0;" ::"n"(NT) : "memory");
#pragma unroll
            for (int q = 0; q < R; q++) x[q].store(sm + bb::bitrev((uint32_t)(base + (q << lowbits)), K) * TILE_COLS + lane * VEC);
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) x[q].store(sm + (base + (q << lowbits)) * TILE_COLS + lane * VEC);
        }
    }
    if (BREV_OUT && tid >= groups) asm volatile("barrier.sync 1, %0;" ::"n"(NT) : "memory");  // idle threads still join the barrier
}

template <int VEC>
__global__ void __launch_bounds__(THREADS, NTT_MIN_CTAS) pass_kernel(const PassParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int LV = VEC == 4 ? 2 : 0;
    const int lc = p.lc, ll = lc - LV;
    const int TILE_COLS = 1 << lc, LANES = 1 << ll;
    const int K = p.K, n = p.n, s0 = p.s0;
    const int L = n - s0 - K;
    const int R = 1 << K;
    uint32_t* sm = smem;                           // [R][TILE_COLS]
    uint2* sm_tw = reinterpret_cast<uint2*>(smem + R * TILE_COLS);        // [R/2] local roots (Shoup pairs)
    uint32_t* sm_row = reinterpret_cast<uint32_t*>(sm_tw + (R > 1 ? R / 2 : 1));  // [R] per-row factor (prescale, then twist/postscale)
    const int tid = threadIdx.x;

    const uint32_t col_tiles = (p.width + TILE_COLS - 1) / TILE_COLS;
    const uint32_t ct = blockIdx.x % col_tiles;
    const uint64_t rt = blockIdx.x / col_tiles;    // row tile: (high, low)
    const uint64_t low = rt & ((1ull << L) - 1);
    const uint64_t high = rt >> L;
    const uint64_t row_base = (high << (n - s0)) + low;  // row of slot t = row_base + (t << L)
    const uint32_t col0 = ct * TILE_COLS;

    // L2 prefetch of the tile that the CTA scheduled `prefetch_dist` blocks later will load: by the time it starts,
    // its rows are in L2 and the load phase sees L2 latency instead of HBM latency (ncu: long_scoreboard was the top stall)
    if (p.prefetch_dist && blockIdx.x + p.prefetch_dist < gridDim.x) {
        const uint32_t fb = blockIdx.x + p.prefetch_dist;
        const uint32_t fct = fb % col_tiles;
        const uint64_t frt = fb / col_tiles;
        const uint64_t frow_base = ((frt >> L) << (n - s0)) + (frt & ((1ull << L) - 1));
        const int lines_per_row = (TILE_COLS * 4 + 127) / 128;  // 128 B lines per row segment
        for (int i = tid; i < R * lines_per_row; i += THREADS) {
            const int t = i / lines_per_row, ln = i % lines_per_row;
            const uint32_t col = fct * TILE_COLS + ln * 32;
            if (col < p.width) {
                const uint32_t* ptr = p.in + (frow_base + ((uint64_t)t << L)) * p.in_pitch + col;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        }
    }
    for (int i = tid; i < R / 2; i += THREADS) sm_tw[i] = __ldg(p.tw_local + i);
    if (p.pre_lo)
        for (int t = tid; t < R; t += THREADS) sm_row[t] = pow2level(p.pre_lo, p.pre_hi, row_base + ((uint64_t)t << L));
    if (p.pre_lo) __syncthreads();

    // ---- load tile (each row segment is 2^lc * 4 B contiguous)
    for (int i = tid; i < R * LANES; i += THREADS) {
        const int t = i >> ll, lane = i & (LANES - 1);
        const uint32_t col = col0 + lane * VEC;
        Vec<VEC> x;
        if (col < p.width) {
            x.load(p.in + (row_base + ((uint64_t)t << L)) * p.in_pitch + col);
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
#ifndef NTT_NO_TRIVIAL
    constexpr bool SKIP = true;
#else
    constexpr bool SKIP = false;
#endif
    if (rem == 1) { if (K == 1) radix_round<1, VEC, SKIP>(sm, sm_tw, K, lc, u, tid); else radix_round<1, VEC, false>(sm, sm_tw, K, lc, u, tid); u += 1; __syncthreads(); }
    if (rem == 2) { if (K == 2) radix_round<2, VEC, SKIP>(sm, sm_tw, K, lc, u, tid); else radix_round<2, VEC, false>(sm, sm_tw, K, lc, u, tid); u += 2; __syncthreads(); }
    for (; u + 3 < K; u += 3) { radix_round<3, VEC, false>(sm, sm_tw, K, lc, u, tid); __syncthreads(); }
    if (u < K) { radix_round<3, VEC, SKIP>(sm, sm_tw, K, lc, u, tid); __syncthreads(); }

    // ---- store
    for (int i = tid; i < R * LANES; i += THREADS) {
        const int t = i >> ll, lane = i & (LANES - 1);
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
        x.store(p.out + row * p.out_pitch + col);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Direct pass (round 2): one CTA per tile, no staging copy.  The FIRST register round loads its rows straight from global
// memory (LDG.128, eight lanes per 128-byte row segment) and the LAST one stores straight to global memory, so a tile crosses
// shared memory twice per interior round boundary only: 4 tile-sized shared-memory transfers for K = 8 instead of the 8 of
// the TMA ring (fill + 3 x (load, store) + drain).  Measured on B200 (profiles/ntt_lab_r02.txt): in the TMA kernel the
// shared-memory time of a tile (2048 clk) ADDS to its arithmetic (2700 clk) instead of hiding under it, because all warps
// of a CTA sit in the same phase; here 3-4 independent CTAs per SM are resident in different phases (global-load wait,
// arithmetic, exchange), and the per-row factor tables are computed while the tile's loads are in flight.
// K, LC are compile-time (index arithmetic folds away): tiles are 2^K rows x 2^LC columns = 2^13 elements, 256 threads,
// 32 elements per thread as 8 rows x 4 adjacent columns.
#ifndef NTT_DIRECT_CTAS
#define NTT_DIRECT_CTAS 3
#endif
constexpr int DIRECT_THREADS = 256;

template <int K, int LC, int k, int U, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void direct_round(uint32_t* sm, const uint2* sm_tw, const PassParams& p, uint64_t row_base, uint32_t col0, int L, const uint2* fac_pre,
                                             const uint2* fac_post, int tid) {
    constexpr int LL = LC - 2, TILE_COLS = 1 << LC, R = 1 << k;
    constexpr int lowbits = K - U - k;
    constexpr int groups = 1 << (K - k + LL);
    constexpr bool LAST = (U + k == K);
    static_assert(groups % DIRECT_THREADS == 0 || groups < DIRECT_THREADS, "group count");
#pragma unroll
    for (int gi0 = 0; gi0 < groups; gi0 += DIRECT_THREADS) {
        const int gi = gi0 + tid;
        if (groups < DIRECT_THREADS && gi >= groups) break;
        const int lane = gi & ((1 << LL) - 1);
        const int g = gi >> LL;
        const int highpart = g >> lowbits, lowpart = g & ((1 << lowbits) - 1);
        const int base = (highpart << (K - U)) + lowpart;
        const uint32_t col = col0 + lane * 4;
        const bool col_ok = col < p.width;
        Vec<4> x[R];
        if (FROM_GLOBAL) {
            const uint32_t* src = p.in + (row_base + ((uint64_t)base << L)) * p.in_pitch + col;
            const uint64_t step = ((uint64_t)p.in_pitch << L) << lowbits;
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (col_ok) x[q].load(src + q * step);
                else x[q].v[0] = x[q].v[1] = x[q].v[2] = x[q].v[3] = 0;
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) x[q].load(sm + (base + (q << lowbits)) * TILE_COLS + lane * 4);
        }
        if (fac_pre) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = fac_pre[base + (q << lowbits)];
#pragma unroll
                for (int c = 0; c < 4; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
#pragma unroll
        for (int v = 0; v < k; v++) {
            const int half = 1 << (k - 1 - v);
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (q & half) continue;
                if (LAST && (q & (half - 1)) == 0) {  // w^0
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t a = x[q].v[c], b = x[q + half].v[c];
                        x[q].v[c] = bb::add(a, b);
                        x[q + half].v[c] = bb::sub(a, b);
                    }
                    continue;
                }
                const int e = (((q & (half - 1)) << lowbits) + lowpart) << (U + v);
                const uint2 w = sm_tw[e];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t a = x[q].v[c], b = x[q + half].v[c];
                    x[q].v[c] = bb::add(a, b);
                    const uint32_t d = a - b + bb::P;
                    const uint32_t qq = __umulhi(d, w.y);
                    x[q + half].v[c] = bb::red2p(d * w.x - qq * bb::P);
                }
            }
        }
        if (fac_post) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = fac_post[base + (q << lowbits)];
#pragma unroll
                for (int c = 0; c < 4; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
        if (TO_GLOBAL) {
            if (col_ok) {
#pragma unroll
                for (int q = 0; q < R; q++) {
                    uint64_t row = row_base + ((uint64_t)(base + (q << lowbits)) << L);
                    if (p.out_natural) row = bb::bitrev((uint32_t)row, p.n);
                    x[q].store(p.out + row * p.out_pitch + col);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) x[q].store(sm + (base + (q << lowbits)) * TILE_COLS + lane * 4);
        }
    }
}

template <int K, int LC>
__global__ void __launch_bounds__(DIRECT_THREADS, NTT_DIRECT_CTAS) pass_kernel_direct(const PassParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int R = 1 << K, TILE_COLS = 1 << LC;
    uint32_t* sm = smem;                                                     // [R][TILE_COLS]
    uint2* sm_tw = reinterpret_cast<uint2*>(smem + R * TILE_COLS);           // [R/2] local roots (Shoup pairs)
    uint2* sm_fac = sm_tw + R / 2;                                           // [2][R] prescale, twist (Shoup pairs)
    const int tid = threadIdx.x;
    const int n = p.n, s0 = p.s0;
    const int L = n - s0 - K;
    const uint32_t col_tiles = (p.width + TILE_COLS - 1) >> LC;
    const uint32_t ct = blockIdx.x % col_tiles;
    const uint64_t rt = blockIdx.x / col_tiles;
    const uint64_t low = rt & ((1ull << L) - 1);
    const uint64_t high = rt >> L;
    const uint64_t row_base = (high << (n - s0)) + low;
    const uint32_t col0 = ct << LC;
    const bool need_twist = L > 0 && low != 0;
    const uint2* fac_pre = p.pre_lo ? sm_fac : nullptr;
    const uint2* fac_post = need_twist ? sm_fac + R : nullptr;

    // L2 prefetch of the tile the CTA scheduled `prefetch_dist` blocks later will load (one 128-byte line per row and 32 columns):
    // when that CTA starts, its first-round loads see L2 latency instead of HBM latency (ncu r02: long_scoreboard is this
    // kernel's top stall while DRAM runs at 45 %)
    if (p.prefetch_dist && blockIdx.x + p.prefetch_dist < gridDim.x) {
        const uint32_t fb = blockIdx.x + p.prefetch_dist;
        const uint32_t fct = fb % col_tiles;
        const uint64_t frt = fb / col_tiles;
        const uint64_t frow_base = ((frt >> L) << (n - s0)) + (frt & ((1ull << L) - 1));
        constexpr int LINES = TILE_COLS / 32 > 0 ? TILE_COLS / 32 : 1;
        for (int i = tid; i < R * LINES; i += DIRECT_THREADS) {
            const int t = i / LINES, ln = i % LINES;
            const uint32_t c = (fct << LC) + ln * 32;
            if (c < p.width) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in + (frow_base + ((uint64_t)t << L)) * p.in_pitch + c));
        }
    }
    // tables first (their L2 look-ups overlap the first round's global loads, which do not depend on them until the
    // barrier below); R <= 256 rows, one per thread
    for (int i = tid; i < R / 2; i += DIRECT_THREADS) sm_tw[i] = __ldg(p.tw_local + i);
    for (int t = tid; t < R; t += DIRECT_THREADS) {
        if (p.pre_lo) sm_fac[t] = shoup_pair(pow2level(p.pre_lo, p.pre_hi, row_base + ((uint64_t)t << L)));
        if (need_twist) {
            uint64_t e = (low * (uint64_t)bb::bitrev((uint32_t)t, K)) << s0;
            if (p.inverse) e = ((1ull << n) - e) & ((1ull << n) - 1);
            sm_fac[R + t] = shoup_pair(pow2level(p.tw_lo, p.tw_hi, e));
        }
    }
    __syncthreads();
    constexpr int REM = K % 3;
    if (K <= 3) {
        direct_round<K, LC, K, 0, true, true>(sm, sm_tw, p, row_base, col0, L, fac_pre, fac_post, tid);
    } else if (REM == 0) {
        direct_round<K, LC, 3, 0, true, false>(sm, sm_tw, p, row_base, col0, L, fac_pre, nullptr, tid);
        __syncthreads();
        if (K == 9) {
            direct_round<K, LC, 3, 3, false, false>(sm, sm_tw, p, row_base, col0, L, nullptr, nullptr, tid);
            __syncthreads();
        }
        direct_round<K, LC, 3, K - 3, false, true>(sm, sm_tw, p, row_base, col0, L, nullptr, fac_post, tid);
    } else {
        direct_round<K, LC, REM, 0, true, false>(sm, sm_tw, p, row_base, col0, L, fac_pre, nullptr, tid);
        __syncthreads();
        if (K > 6) {
            direct_round<K, LC, 3, REM, false, false>(sm, sm_tw, p, row_base, col0, L, nullptr, nullptr, tid);
            __syncthreads();
        }
        direct_round<K, LC, 3, K - 3, false, true>(sm, sm_tw, p, row_base, col0, L, nullptr, fac_post, tid);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused middle of a coset LDE (round 2).  The tile the LAST inverse pass finishes -- positions [j 2^K, (j+1) 2^K), which hold
// the coefficients  k 2^(n-K) + q,  k = bitrev_K(slot), q = bitrev_(n-K)(j)  -- is exactly the tile the FIRST forward pass of
// every coset starts from.  One CTA therefore: loads the tile once, finishes the inverse transform in shared memory, and for
// each coset c multiplies by the coset powers and runs the first K forward stages, storing straight into coset block c.
// The matrix is read once and written C times instead of being read 1 + C times and written 1 + C times (for log_blowup 1:
// 3 sweeps of 8 B/element instead of 6), and the natural-order coefficient matrix never exists in HBM.
//   prescale  s_c^(k 2^(n-K) + q) / N  =  sigma_c[k] * (s_c^q / N):  sigma is tile independent (a 2^K-entry table per coset),
//   the per-tile scalar is folded into the twist  w_N^(q * bitrev_K(slot))  that the first forward pass applies anyway.
// The forward rounds read the inverse result through bit-reversed slot numbers (whole 2^LC-column rows move, so any row
// permutation is bank-conflict free) -- no reordering pass.  Tiles are 2^13 elements; shared memory holds the inverse
// result X, a work tile Y, both local root tables and the twists: 70 KB, three CTAs per SM.
constexpr int MID_MAX_COSETS = 8;
#ifndef NTT_MID_CTAS
#define NTT_MID_CTAS 3
#endif
struct MidParams {
    const uint32_t* in;       // inverse transform after all but its last pass (2^n rows, DIF positions)
    uint32_t* out;            // coset block c = out + c * block_stride
    uint64_t block_stride;    // elements between coset blocks
    uint32_t in_pitch, out_pitch, width;
    int n, cosets;
    uint32_t prefetch_dist;   // CTAs of look-ahead of the L2 prefetch (0: none)
    const uint2* tw_inv;      // 2^(K-1) local inverse roots (Shoup pairs)
    const uint2* tw_fwd;      // 2^(K-1) local forward roots
    const uint32_t* tw_lo;    // w_N^i, two-level
    const uint32_t* tw_hi;
    const uint2* sigma;       // [cosets][2^K]: s_c^(k 2^(n-K)) as Shoup pairs
    const uint32_t* pre_lo[MID_MAX_COSETS];  // two-level tables of s_c^j / N (the ones the unfused prescale uses)
    const uint32_t* pre_hi[MID_MAX_COSETS];
};
enum MidSrc { MID_FROM_GLOBAL = 0, MID_FROM_SMEM = 1, MID_FROM_SMEM_BREV = 2 };

// one register round (k stages starting at local stage U) of a 2^K-point DIF over a 2^K x 2^LC tile, 256 threads, a thread
// owning 2^k rows x 4 adjacent columns (NT threads per CTA, 32 elements per thread and round).  g_in / g_out already point at this thread's column; *_step = elements per slot.
template <int K, int LC, int NT, int k, int U, int SRC, bool TO_GLOBAL>
__device__ __forceinline__ void mid_round(const uint32_t* sm_in, uint32_t* sm_out, const uint2* sm_tw, const uint32_t* g_in, uint64_t g_in_step, uint32_t* g_out,
                                          uint64_t g_out_step, bool col_ok, const uint2* fac_pre, const uint2* fac_post, int tid) {
    constexpr int LL = LC - 2, TILE_COLS = 1 << LC, R = 1 << k;
    constexpr int lowbits = K - U - k;
    constexpr int groups = 1 << (K - k + LL);
    constexpr bool LAST = (U + k == K);
    static_assert(groups % NT == 0, "group count");
#pragma unroll
    for (int gi0 = 0; gi0 < groups; gi0 += NT) {
        const int gi = gi0 + tid;
        const int lane = gi & ((1 << LL) - 1);
        const int g = gi >> LL;
        const int highpart = g >> lowbits, lowpart = g & ((1 << lowbits) - 1);
        const int base = (highpart << (K - U)) + lowpart;
        Vec<4> x[R];
        if (SRC == MID_FROM_GLOBAL) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (col_ok) x[q].load(g_in + (uint64_t)(base + (q << lowbits)) * g_in_step);
                else x[q].v[0] = x[q].v[1] = x[q].v[2] = x[q].v[3] = 0;
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const int slot = base + (q << lowbits);
                const int src = SRC == MID_FROM_SMEM_BREV ? (int)(__brev((uint32_t)slot) >> (32 - K)) : slot;
                x[q].load(sm_in + src * TILE_COLS + lane * 4);
            }
        }
        if (fac_pre) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = __ldg(fac_pre + base + (q << lowbits));
#pragma unroll
                for (int c = 0; c < 4; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
#pragma unroll
        for (int v = 0; v < k; v++) {
            const int half = 1 << (k - 1 - v);
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (q & half) continue;
                if (LAST && (q & (half - 1)) == 0) {  // w^0
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t a = x[q].v[c], b = x[q + half].v[c];
                        x[q].v[c] = bb::add(a, b);
                        x[q + half].v[c] = bb::sub(a, b);
                    }
                    continue;
                }
                const int e = (((q & (half - 1)) << lowbits) + lowpart) << (U + v);
                const uint2 w = sm_tw[e];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t a = x[q].v[c], b = x[q + half].v[c];
                    x[q].v[c] = bb::add(a, b);
                    const uint32_t d = a - b + bb::P;
                    const uint32_t qq = __umulhi(d, w.y);
                    x[q + half].v[c] = bb::red2p(d * w.x - qq * bb::P);
                }
            }
        }
        if (fac_post) {
#pragma unroll
            for (int q = 0; q < R; q++) {
                const uint2 f = fac_post[base + (q << lowbits)];
#pragma unroll
                for (int c = 0; c < 4; c++) x[q].v[c] = shoup_mul(x[q].v[c], f);
            }
        }
        if (TO_GLOBAL) {
            if (col_ok) {
#pragma unroll
                for (int q = 0; q < R; q++) x[q].store(g_out + (uint64_t)(base + (q << lowbits)) * g_out_step);
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; q++) x[q].store(sm_out + (base + (q << lowbits)) * TILE_COLS + lane * 4);
        }
    }
}

// the K stages of one 2^K-point DIF as register rounds (remainder round first, like the other pass kernels);
// FIRST_SRC says where round one reads (global / X / X through bit-reversed slots), the last round stores to global if
// TO_GLOBAL, else everything ends in sm_work
template <int K, int LC, int NT, int FIRST_SRC, bool TO_GLOBAL>
__device__ __forceinline__ void mid_dif(const uint32_t* sm_first, uint32_t* sm_work, const uint2* sm_tw, const uint32_t* g_in, uint64_t g_in_step, uint32_t* g_out,
                                        uint64_t g_out_step, bool col_ok, const uint2* fac_pre, const uint2* fac_post, int tid) {
    constexpr int REM = K % 3;
    constexpr int K1 = REM ? REM : 3;  // stages of the first round
    static_assert(K >= 6 && K <= 8, "fused middle supports K = 6, 7, 8");
    mid_round<K, LC, NT, K1, 0, FIRST_SRC, false>(sm_first, sm_work, sm_tw, g_in, g_in_step, nullptr, 0, col_ok, fac_pre, nullptr, tid);
    __syncthreads();
    if (K1 + 3 < K) {
        mid_round<K, LC, NT, 3, K1, MID_FROM_SMEM, false>(sm_work, sm_work, sm_tw, nullptr, 0, nullptr, 0, col_ok, nullptr, nullptr, tid);
        __syncthreads();
    }
    mid_round<K, LC, NT, 3, K - 3, MID_FROM_SMEM, TO_GLOBAL>(sm_work, sm_work, sm_tw, nullptr, 0, g_out, g_out_step, col_ok, nullptr, fac_post, tid);
}

template <int K, int LC, int NT, int CTAS>
__global__ void __launch_bounds__(NT, CTAS) lde_mid_kernel(const MidParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int R = 1 << K, TILE_COLS = 1 << LC, LL = LC - 2;
    uint32_t* X = smem;                                                      // [R][TILE_COLS] inverse result (bit-reversed slots)
    uint32_t* Y = X + R * TILE_COLS;                                         // [R][TILE_COLS] forward work tile
    uint2* tw_inv = reinterpret_cast<uint2*>(Y + R * TILE_COLS);             // [R/2]
    uint2* tw_fwd = tw_inv + R / 2;                                          // [R/2]
    uint2* twist = tw_fwd + R / 2;                                           // [cosets][R]
    const int tid = threadIdx.x;
    const int n = p.n, Lf = n - K;
    const uint32_t col_tiles = (p.width + TILE_COLS - 1) >> LC;
    const uint32_t ct = blockIdx.x % col_tiles;
    const uint32_t j = blockIdx.x / col_tiles;                               // row tile of the inverse = high index of its last pass
    const uint32_t q = bb::bitrev(j, Lf);                                    // low index of the first forward pass
    const uint32_t col = (ct << LC) + (uint32_t)(tid & ((1 << LL) - 1)) * 4;
    const bool col_ok = col < p.width;

    if (p.prefetch_dist && blockIdx.x + p.prefetch_dist < gridDim.x) {  // L2 prefetch of a later CTA's tile (see pass_kernel_direct)
        const uint32_t fb = blockIdx.x + p.prefetch_dist;
        const uint32_t fct = fb % col_tiles, fj = fb / col_tiles;
        constexpr int LINES = TILE_COLS / 32 > 0 ? TILE_COLS / 32 : 1;
        for (int i = tid; i < R * LINES; i += NT) {
            const int t = i / LINES, ln = i % LINES;
            const uint32_t c = (fct << LC) + ln * 32;
            if (c < p.width) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.in + (((uint64_t)fj << K) + t) * p.in_pitch + c));
        }
    }
    for (int i = tid; i < R / 2; i += NT) {
        tw_inv[i] = __ldg(p.tw_inv + i);
        tw_fwd[i] = __ldg(p.tw_fwd + i);
    }
    for (int t = tid; t < R; t += NT) {
        uint32_t w = bb::ONE;
        if (q != 0) w = pow2level(p.tw_lo, p.tw_hi, (uint64_t)q * bb::bitrev((uint32_t)t, K));   // < 2^(n-K) * 2^K
        for (int c = 0; c < p.cosets; c++) twist[c * R + t] = shoup_pair(bb::mul(w, pow2level(p.pre_lo[c], p.pre_hi[c], q)));
    }
    __syncthreads();

    // inverse tail: rows j 2^K + slot, contiguous
    mid_dif<K, LC, NT, MID_FROM_GLOBAL, false>(nullptr, X, tw_inv, p.in + ((uint64_t)j << K) * p.in_pitch + col, p.in_pitch, nullptr, 0, col_ok, nullptr, nullptr, tid);
    __syncthreads();
    // forward head of every coset: logical k lives in slot bitrev_K(k) of X; output slot t goes to row q + t 2^(n-K) of the block
    for (int c = 0; c < p.cosets; c++) {
        uint32_t* dst = p.out + (uint64_t)c * p.block_stride + (uint64_t)q * p.out_pitch + col;
        mid_dif<K, LC, NT, MID_FROM_SMEM_BREV, true>(X, Y, tw_fwd, nullptr, 0, dst, (uint64_t)p.out_pitch << Lf, col_ok, p.sigma + c * R, twist + c * R, tid);
        __syncthreads();  // Y is rewritten by the next coset's first round
    }
}

// sigma[k] = base^k as a Shoup pair (base = s_c^(2^(n-K)), Montgomery form)
__global__ void mid_sigma_kernel(uint2* out, uint32_t base, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = shoup_pair(bb::pow(base, i));
}

// ------------------------------------------------------------------------------------------------------------------
// TMA pass: the same K stages, as a persistent warp-specialised kernel.  One producer warp moves whole tiles between
// HBM and shared memory with cp.async.bulk.tensor (one 4-D box per tile: {columns, low, t, high} of the strided row set)
// through a 3-stage mbarrier ring; 256 consumer threads only ever touch shared memory, so butterflies of tile c overlap
// the load of tile c+1 and the store of tile c-1.  Tiles are 2^13 elements (32 KB): 3 stages x 2 CTAs per SM.
constexpr int TMA_CONSUMERS = 256;
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;
#ifndef NTT_TMA_STAGES
#define NTT_TMA_STAGES 3
#endif
#ifndef NTT_TMA_CTAS
#define NTT_TMA_CTAS 2
#endif
constexpr int TMA_STAGES = NTT_TMA_STAGES;
constexpr int TMA_TILE_LOG = 13;
constexpr int TMA_MAX_K = 8;

struct alignas(64) TensorMap {  // same size/alignment as CUtensorMap (cuda.h), kept opaque here
    uint64_t opaque[16];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const TensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const TensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3)
                 : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, NTT_TMA_CTAS)
pass_kernel_tma(const __grid_constant__ TensorMap in_map, const __grid_constant__ TensorMap out_map, const PassParams p, uint32_t n_tiles) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int K = p.K, n = p.n, s0 = p.s0, lc = p.lc;
    const int L = n - s0 - K;
    const int R = 1 << K;
    const uint32_t tile_bytes = (uint32_t)(R << lc) * 4u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                        // [3] tile landed in smem   (first 128 B: barriers)
    uint64_t* done = full + TMA_STAGES;                                            // [3] consumers finished the tile
    uint32_t* bufs = reinterpret_cast<uint32_t*>(smem_raw + 128);                  // TMA_STAGES tiles, each 128-B aligned
    uint2* sm_tw = reinterpret_cast<uint2*>(bufs + TMA_STAGES * (tile_bytes / 4));  // [R/2] local roots (Shoup pairs)
    uint2* sm_fac = sm_tw + (R > 1 ? R / 2 : 1);                                   // [2][R] prescale / twist per row (Shoup pairs)
    const int tid = threadIdx.x;
    const uint32_t col_tiles = (p.width + (1u << lc) - 1) >> lc;

    if (tid == 0) {
        for (int i = 0; i < TMA_STAGES; i++) {
            mbar_init(full + i, 1);
            mbar_init(done + i, TMA_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

    if (tid >= TMA_CONSUMERS) {
        // ===== producer warp: one elected lane drives the ring
        if (tid == TMA_CONSUMERS) {
            auto coords = [&](uint32_t i, int& c0, int& c1, int& c3) {
                const uint32_t tile = first + i * stride;
                const uint32_t ct = tile % col_tiles;
                const uint64_t rt = tile / col_tiles + p.rt_base;
                c0 = (int)(ct << lc);
                c1 = (int)(rt & ((1ull << L) - 1));
                c3 = (int)(rt >> L);
            };
            auto load = [&](uint32_t i) {
                int c0, c1, c3;
                coords(i, c0, c1, c3);
                const int b = i % TMA_STAGES;
                mbar_expect_tx(full + b, tile_bytes);
                tma_load_4d(bufs + b * (tile_bytes / 4), &in_map, full + b, c0, c1, 0, c3);
            };
            for (uint32_t i = 0; i < my_tiles && i < (uint32_t)TMA_STAGES; i++) load(i);
            for (uint32_t c = 0; c < my_tiles; c++) {
                const int b = c % TMA_STAGES;
                mbar_wait(done + b, (c / TMA_STAGES) & 1);       // consumers are finished with tile c (they fenced the async proxy)
                int c0, c1, c3;
                coords(c, c0, c1, c3);
                if (p.out_natural) {  // last pass (L = 0): the tile leaves as rows k1 * 2^(n-K) + bitrev(high), k1 = 0..2^K-1
                    c1 = (int)bb::bitrev((uint32_t)c3, n - K);
                    c3 = 0;
                } else {
                    c3 -= (int)(p.rt_base >> L);  // the destination map of a partial launch starts at its own first row tile
                }
                tma_store_4d(&out_map, bufs + b * (tile_bytes / 4), c0, c1, 0, c3);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (c + TMA_STAGES < my_tiles) {
                    // Refill THIS buffer as soon as its store has drained it from shared memory (about a microsecond), not when
                    // the consumers finish the next tile: every stage of the ring then holds a tile in flight, i.e. the loads run
                    // STAGES - 1 tiles ahead of the consumers instead of one (ncu r01: waiting for a tile was their top stall).
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    load(c + TMA_STAGES);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }

    // ===== consumers
    for (int i = tid; i < R / 2; i += TMA_CONSUMERS) sm_tw[i] = __ldg(p.tw_local + i);
    const int rem = K % 3;
    for (uint32_t c = 0; c < my_tiles; c++) {
        const int b = c % TMA_STAGES;
        uint32_t* sm = bufs + b * (tile_bytes / 4);
        const uint32_t tile = first + c * stride;
        const uint64_t rt = tile / col_tiles + p.rt_base;
        const uint64_t low = rt & ((1ull << L) - 1);
        const uint64_t high = rt >> L;
        const uint64_t row_base = (high << (n - s0)) + low;
        const bool need_twist = L > 0 && low != 0;
        // per-row factors for this tile (tables are L2/L1 resident); double-buffered by tile parity
        uint2* fac_pre = p.pre_lo ? sm_fac : nullptr;
        uint2* fac_post = need_twist ? sm_fac + R : nullptr;
        asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");  // previous tile's factors are no longer read
        for (int t = tid; t < R; t += TMA_CONSUMERS) {
            if (fac_pre) fac_pre[t] = shoup_pair(pow2level(p.pre_lo, p.pre_hi, row_base + ((uint64_t)t << L)));
            if (fac_post) {
                uint64_t e = (low * (uint64_t)bb::bitrev((uint32_t)t, K)) << s0;
                if (p.inverse) e = ((1ull << n) - e) & ((1ull << n) - 1);
                fac_post[t] = shoup_pair(pow2level(p.tw_lo, p.tw_hi, e));
            }
        }
        // one warp polls the mbarrier, the other seven sleep in the hardware barrier below: eight polling warps spent 14 % of
        // the kernel's issue slots on try_wait / branch (ncu r01), slots the other CTA of the SM needs for butterflies
        if (tid < 32) mbar_wait(full + b, (c / TMA_STAGES) & 1);
        asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");  // tile landed (observed by warp 0) and factors visible to every consumer
        int u = 0;
        const uint2* pre = fac_pre;
        auto post_if_last = [&](int stages) { return (u + stages == K) ? (const uint2*)fac_post : (const uint2*)nullptr; };
        const bool brev = p.out_natural != 0;
        if (rem == 1) {
            if (K == 1 && brev) radix_round<1, 4, true, TMA_CONSUMERS, true>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(1));
            else if (K == 1) radix_round<1, 4, true, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(1));
            else radix_round<1, 4, false, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(1));
            u += 1; pre = nullptr;
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
        }
        if (rem == 2) {
            if (K == 2 && brev) radix_round<2, 4, true, TMA_CONSUMERS, true>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(2));
            else if (K == 2) radix_round<2, 4, true, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(2));
            else radix_round<2, 4, false, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, post_if_last(2));
            u += 2; pre = nullptr;
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
        }
        for (; u + 3 < K; u += 3) {
            radix_round<3, 4, false, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, nullptr);
            pre = nullptr;
            asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
        }
        if (u < K) {
            if (brev) radix_round<3, 4, true, TMA_CONSUMERS, true>(sm, sm_tw, K, lc, u, tid, pre, fac_post);
            else radix_round<3, 4, true, TMA_CONSUMERS>(sm, sm_tw, K, lc, u, tid, pre, fac_post);
            u += 3;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
        mbar_arrive(done + b);
    }
}

// local roots for the butterflies: (w, floor(w 2^32 / p)) with w = g^i as a plain integer, i < half
__global__ void shoup_table_kernel(uint2* tw, uint32_t g_monty, uint32_t half) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const uint32_t w = bb::from_monty(bb::pow(g_monty, i));
    tw[i] = make_uint2(w, (uint32_t)((((uint64_t)w) << 32) / bb::P));
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
