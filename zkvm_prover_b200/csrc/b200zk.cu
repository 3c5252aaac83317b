// C ABI of libb200zk (see include/b200zk.h).  Host-side orchestration only: argument checking, device
// memory, pass planning and kernel launches on the context's stream.  No CPU arithmetic path exists here:
// every field operation on user data happens in the kernels of ntt.cuh / merkle.cuh / fri.cuh.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200zk.h"
#include "fri.cuh"
#include "merkle.cuh"
#include "ntt.cuh"
#include "open.cuh"

namespace {

constexpr int MAX_LOG = 27;
constexpr int KMAX_SMALL = 9;   // 64 KB tile: two CTAs per SM
constexpr int KMAX_BIG = 10;    // 128 KB tile: used when it saves a whole pass (n = 19, 20)
constexpr uint32_t TOP_LAYER = 2048;  // digest layers at or below this size are finished by one CTA

}  // namespace

struct b200zk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint2* tw_local[2][KMAX_BIG + 1] = {};  // Shoup pairs
    uint32_t* tw_lo[MAX_LOG + 1] = {};
    uint32_t* tw_hi[MAX_LOG + 1] = {};
    uint32_t* tab = nullptr;  // scratch for per-call power tables (stream ordered reuse)
    size_t tab_words = 0;
    uint2* mid_sigma = nullptr;  // [cosets][2^K] coset powers of the fused LDE middle (stream ordered reuse)
    size_t mid_sigma_pairs = 0;
    uint32_t* d_small = nullptr;  // 64 KB of small device scratch (roots, betas, flags)
    int max_smem_optin = 0;
    int num_sms = 148;
    void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled (driver entry point), null if unavailable
    cudaStream_t copy_stream = nullptr;  // host->device strip copies of b200zk_lde_commit_host (created on first use)
    // strip buffers of the host pipeline: two sets of two, dedicated (never handed to the allocator), so the copies of call
    // i + 1 can run while call i is still computing out of the other set; one (copied, consumed) event pair per buffer
    cudaStream_t hash_stream = nullptr;  // experiment (B200ZK_HASH_STREAM=1): strip absorbs run here, under the next strip's LDE
    cudaEvent_t ev_lde[8] = {}, ev_abs = nullptr;
    uint32_t* h_root_ring = nullptr;  // pinned, 64 slots of 8 words (roots of asynchronous commits)
    uint32_t root_ring_next = 0;
    uint32_t* strip_buf[4] = {};
    size_t strip_buf_bytes = 0;
    uint64_t strip_calls = 0;
    cudaEvent_t ev_copied[4] = {}, ev_consumed[4] = {};
    std::unordered_multimap<size_t, void*> cache;   // freed blocks by (rounded) size, see dev_alloc
    std::unordered_map<void*, size_t> live;         // blocks handed out by dev_alloc
    size_t cache_bytes = 0, cache_cap = 0;
};
struct b200zk_mat {
    uint32_t* d = nullptr;
    uint64_t rows = 0;
    uint32_t width = 0;
    bool owned = false;
};
struct b200zk_tree {
    std::vector<b200zk_mat*> mats;  // original order
    bool owns_mats = false;
    uint64_t max_h = 0;
    uint32_t depth = 0;
    uint64_t total_width = 0;
    uint32_t* d_digests = nullptr;  // layer 0 | layer 1 | ... | root
    std::vector<uint64_t> layer_off;  // in digests
    mk::OpenMat* d_open = nullptr;
    // asynchronous commits (b200zk_lde_commit_host_async): the root is copied to a pinned slot behind the last kernel and
    // `ev_done` marks that point, so b200zk_tree_root waits for THIS tree only, not for commits enqueued after it
    mutable cudaEvent_t ev_done = nullptr;
    mutable uint32_t* h_root_slot = nullptr;   // one of 64 pinned slots of the context: valid until 64 later asynchronous commits were issued
    mutable uint32_t root_cache[8] = {};       // the root, once collected (the slot is recycled)
    mutable bool root_cached = false;
};
struct b200zk_chal {
    fri::ChalState* d = nullptr;
};

namespace {

int fail(b200zk_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA,       \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
        }                                                                                              \
    } while (0)
#define LAUNCHED()                                                                                     \
    do {                                                                                               \
        ctx->launches++;                                                                               \
        CU(cudaGetLastError());                                                                        \
    } while (0)
#define TRY(call)                                                                                      \
    do {                                                                                               \
        int r_ = (call);                                                                               \
        if (r_ != B200ZK_OK) return r_;                                                                \
    } while (0)

inline bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
inline int log2u(uint64_t x) {
    int l = 0;
    while ((1ull << l) < x) l++;
    return l;
}

// Device memory comes from the stream-ordered pool (release threshold = never), fronted by an exact-size block cache per
// ctx: a prover repeats the same allocation sizes segment after segment, and even a pool hit costs 2-5 ms for a GB-sized
// block (50-70 ms when the pool has to grow; measured with the real 17-AIR segment shape), while a cache hit costs a
// hash lookup.  Reuse is safe because every consumer of a block is ordered on ctx->stream (the copy stream of the host
// strip pipeline synchronises with it through events before a buffer is released).  The cache holds at most half of the
// device memory, is emptied by b200zk_ctx_trim, and is flushed automatically when the pool reports out-of-memory.
void cache_flush(b200zk_ctx* ctx) {
    for (auto& kv : ctx->cache) cudaFreeAsync(kv.second, ctx->stream);
    ctx->cache.clear();
    ctx->cache_bytes = 0;
}
int dev_alloc(b200zk_ctx* ctx, size_t bytes, void** out) {
    *out = nullptr;
    if (!bytes) bytes = 16;
    bytes = (bytes + 511) & ~(size_t)511;
    auto it = ctx->cache.find(bytes);
    if (it != ctx->cache.end()) {
        *out = it->second;
        ctx->cache.erase(it);
        ctx->cache_bytes -= bytes;
        ctx->live.emplace(*out, bytes);
        return B200ZK_OK;
    }
    cudaError_t e = cudaMallocAsync(out, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation && !ctx->cache.empty()) {  // give the cached blocks back and try once more
        cudaGetLastError();
        cache_flush(ctx);
        cudaStreamSynchronize(ctx->stream);
        e = cudaMallocAsync(out, bytes, ctx->stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(ctx, B200ZK_ERR_OOM, "cudaMallocAsync(" + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    ctx->live.emplace(*out, bytes);
    return B200ZK_OK;
}
void dev_free(b200zk_ctx* ctx, void* p) {
    if (!p) return;
    if (!ctx || !ctx->stream) {
        cudaFree(p);
        return;
    }
    auto it = ctx->live.find(p);
    if (it == ctx->live.end()) {  // not from dev_alloc
        cudaFreeAsync(p, ctx->stream);
        return;
    }
    const size_t bytes = it->second;
    ctx->live.erase(it);
    if (ctx->cache_cap && ctx->cache_bytes + bytes <= ctx->cache_cap) {
        ctx->cache.emplace(bytes, p);
        ctx->cache_bytes += bytes;
    } else {
        cudaFreeAsync(p, ctx->stream);
    }
}

// ---------------------------------------------------------------------------------------------- twiddles
int ensure_tab(b200zk_ctx* ctx, size_t words) {
    if (ctx->tab_words >= words) return B200ZK_OK;
    if (ctx->tab) {
        dev_free(ctx, ctx->tab);
        ctx->tab = nullptr;
        ctx->tab_words = 0;
    }
    TRY(dev_alloc(ctx, words * 4, (void**)&ctx->tab));
    ctx->tab_words = words;
    return B200ZK_OK;
}

int pow_tables(b200zk_ctx* ctx, uint32_t* lo, uint32_t* hi, uint32_t base, uint32_t scale, uint64_t n_entries) {
    uint32_t n_lo = (uint32_t)std::min<uint64_t>(n_entries, 1u << ntt::LO_BITS);
    uint32_t n_hi = (uint32_t)std::max<uint64_t>(n_entries >> ntt::LO_BITS, 1);
    uint32_t m = std::max(n_lo, n_hi);
    ntt::pow_table_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(lo, hi, base, scale, n_lo, n_hi);
    LAUNCHED();
    return B200ZK_OK;
}
inline size_t lo_words(uint64_t n_entries) { return (size_t)std::min<uint64_t>(n_entries, 1u << ntt::LO_BITS); }
inline size_t hi_words(uint64_t n_entries) { return (size_t)std::max<uint64_t>(n_entries >> ntt::LO_BITS, 1); }

int ensure_roots(b200zk_ctx* ctx, int n) {
    if (ctx->tw_lo[n]) return B200ZK_OK;
    uint64_t N = 1ull << n;
    uint32_t *lo = nullptr, *hi = nullptr;  // published only when both tables exist and the generating kernel was enqueued
    TRY(dev_alloc(ctx, lo_words(N) * 4, (void**)&lo));
    int rc = dev_alloc(ctx, hi_words(N) * 4, (void**)&hi);
    if (rc == B200ZK_OK) rc = pow_tables(ctx, lo, hi, bb::two_adic_generator(n), bb::ONE, N);
    if (rc != B200ZK_OK) {
        dev_free(ctx, lo);
        dev_free(ctx, hi);
        return rc;
    }
    ctx->tw_lo[n] = lo;
    ctx->tw_hi[n] = hi;
    return B200ZK_OK;
}
int ensure_local(b200zk_ctx* ctx, int inverse, int K) {
    if (ctx->tw_local[inverse][K]) return B200ZK_OK;
    uint32_t half = K ? (1u << (K - 1)) : 1;
    TRY(dev_alloc(ctx, (size_t)half * 8, (void**)&ctx->tw_local[inverse][K]));
    uint32_t g = bb::two_adic_generator(K);
    if (inverse) g = bb::inv(g);
    ntt::shoup_table_kernel<<<(half + 255) / 256, 256, 0, ctx->stream>>>(ctx->tw_local[inverse][K], g, half);
    LAUNCHED();
    return B200ZK_OK;
}

// split n stages into passes
std::vector<int> make_plan(int n, int kmax_force = 0) {
    std::vector<int> plan;
    if (n <= 0) return plan;
    int kmax = (n > 2 * KMAX_SMALL && n <= 2 * KMAX_BIG) ? KMAX_BIG : KMAX_SMALL;
    if (kmax_force) kmax = kmax_force;
    int m = (n + kmax - 1) / kmax;
    for (int i = 0; i < m; i++) plan.push_back(n / m + (i < n % m ? 1 : 0));
    return plan;
}


static_assert(sizeof(ntt::TensorMap) == sizeof(CUtensorMap), "TensorMap must mirror CUtensorMap");
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D view of the rows a pass touches: {column, low, t, high}; one box = one tile
int make_pass_map(b200zk_ctx* ctx, const uint32_t* base, uint32_t width, uint32_t pitch, int n, int s0, int K, int lc, ntt::TensorMap* out) {
    const int L = n - s0 - K;
    cuuint64_t dims[4] = {width, 1ull << L, 1ull << K, 1ull << s0};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, ((cuuint64_t)pitch * 4) << L, ((cuuint64_t)pitch * 4) << (L + K)};
    cuuint32_t box[4] = {1u << lc, 1, 1u << K, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)base, dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, B200ZK_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return B200ZK_OK;
}

// destination of a scatter pass: natural-order rows k1 * 2^(n-K) + q as the box {column, q, k1}
int make_natural_map(b200zk_ctx* ctx, const uint32_t* base, uint32_t width, uint32_t pitch, int n, int K, int lc, ntt::TensorMap* out) {
    cuuint64_t dims[4] = {width, 1ull << (n - K), 1ull << K, 1};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, ((cuuint64_t)pitch * 4) << (n - K), ((cuuint64_t)pitch * 4) << n};
    cuuint32_t box[4] = {1u << lc, 1, 1u << K, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)base, dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, B200ZK_ERR_CUDA, "cuTensorMapEncodeTiled (natural) failed: " + std::to_string((int)r));
    return B200ZK_OK;
}

// dynamic shared memory of pass_kernel_tma: barriers | STAGES tiles | local roots | {prescale, twist} row factors
inline size_t tma_smem_bytes(int K, uint32_t tcols) {
    const size_t R = (size_t)1 << K;
    return 128 + (size_t)ntt::TMA_STAGES * (R * tcols * 4) + (2 * std::max<size_t>(R / 2, 1) + 4 * R) * 4;
}

// B200ZK_NTT_TRACE=1: synchronise around every pass and print its time (experiments only; serialises the stream)
bool ntt_trace() {
    static const int v = [] {
        const char* e = getenv("B200ZK_NTT_TRACE");
        return e ? atoi(e) : 0;
    }();
    return v != 0;
}
struct PassTimer {
    b200zk_ctx* ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const char* what;
    int n, s0, K;
    uint32_t width;
    PassTimer(b200zk_ctx* c, const char* w, int n_, int s0_, int K_, uint32_t width_) : ctx(c), what(w), n(n_), s0(s0_), K(K_), width(width_) {
        if (!ntt_trace()) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
    }
    ~PassTimer() {
        if (!e0) return;
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 8.0 * (double)(1ull << n) * width;
        fprintf(stderr, "[ntt] %-10s n=%d s0=%2d K=%d width=%u  %.3f ms  %.0f GB/s (read+write)\n", what, n, s0, K, width, ms, bytes / ms / 1e6);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
};

bool tma_enabled() {
    static const int v = [] {
        const char* e = getenv("B200ZK_NTT_TMA");  // experiment knob: 0 forces the plain pass kernel
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

bool direct_enabled() {
    static const int v = [] {
        const char* e = getenv("B200ZK_NTT_DIRECT");  // experiment knob: 0 keeps every pass on the TMA ring kernel
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

struct Scale {
    const uint32_t* lo = nullptr;
    const uint32_t* hi = nullptr;
};

// n-stage DIF over `rows = 2^n` rows.  src -> dst (first pass), then in place on dst; when out_natural the
// last pass scatters into dst_final (which must not alias its source).
// Destination of a sharded LDE (b200zk_coset_lde_scatter): the finished rows do not stay on this GPU, the last pass stores its
// tiles straight into the row-block owner's memory (a peer mapping) with TMA.  Rows g0 .. g0 + 2^n of the whole LDE are
// produced by this transform; rank r owns rows [r * mg, (r + 1) * mg) and keeps them at dst[r] + slot_off, pitch = width.
struct Scatter {
    uint32_t* const* dst = nullptr;
    uint64_t g0 = 0, mg = 0, slot_off = 0;
    uint32_t pitch = 0;  // row pitch at the destination in elements (0: the strip's own width -- one [mg][width] slot per sender)
    uint32_t rank = 0;   // this sender: destinations are visited starting behind it, so the ranks do not all store into the same GPU at once
};

// A caller may hand in its own pass plan and run only the passes [pass_begin, pass_end) of it (the fused LDE middle,
// lde_core below, replaces the last inverse pass and the first forward pass): the first pass that runs reads `src`, the
// coset prescale belongs to pass 0 and the `last` duties (dst_final, post scale, natural order, scatter) to the plan's final pass.
int run_transform(b200zk_ctx* ctx, const uint32_t* src, uint32_t* work, uint32_t* dst_final, int n, uint32_t width, int inverse, Scale pre,
                  Scale post, int out_natural, uint32_t src_pitch = 0, uint32_t work_pitch = 0, uint32_t dst_pitch = 0, const Scatter* scatter = nullptr,
                  const std::vector<int>* plan_in = nullptr, size_t pass_begin = 0, size_t pass_end = ~(size_t)0) {
    if (!src_pitch) src_pitch = width;
    if (!work_pitch) work_pitch = width;
    if (!dst_pitch) dst_pitch = width;
    TRY(ensure_roots(ctx, n));
    const int vec = (width % 4 == 0 && src_pitch % 4 == 0 && work_pitch % 4 == 0 && dst_pitch % 4 == 0 && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)work % 16 == 0) && ((uintptr_t)dst_final % 16 == 0))
                        ? 4
                        : 1;
    const bool tma_ok = vec == 4 && ctx->encode_tiled && tma_enabled();
    std::vector<int> plan = plan_in ? *plan_in : make_plan(n, tma_ok ? ntt::TMA_MAX_K : 0);
    pass_end = std::min(pass_end, plan.size());
    int s0 = 0;
    for (size_t i = 0; i < pass_begin && i < plan.size(); i++) s0 += plan[i];
    for (size_t i = pass_begin; i < pass_end; i++) {
        const int K = plan[i];
        const bool first = i == pass_begin, last = i + 1 == plan.size();
        TRY(ensure_local(ctx, inverse, K));
        ntt::PassParams p;
        p.in = first ? src : work;
        p.out = last ? dst_final : work;
        p.width = width;
        p.in_pitch = first ? src_pitch : work_pitch;
        p.out_pitch = last ? dst_pitch : work_pitch;
        p.n = n;
        p.s0 = s0;
        p.K = K;
        // column tile: 2^14 elements per tile, at most 128 columns, no wider than the matrix needs
        int lc = std::min(ntt::TILE_ELEMS_LOG - K, ntt::MAX_TILE_COLS_LOG);
        while (lc > (vec == 4 ? 2 : 0) && (1u << (lc - 1)) >= width) lc--;
        if (K == KMAX_BIG) lc = 5;  // the two-pass plan for n = 19, 20 keeps its 128 KB tile
        p.lc = lc;
        const uint32_t tile_cols = 1u << lc;
        const uint32_t col_tiles = (width + tile_cols - 1) / tile_cols;
        p.inverse = inverse;
        p.tw_local = ctx->tw_local[inverse][K];
        p.tw_lo = ctx->tw_lo[n];
        p.tw_hi = ctx->tw_hi[n];
        p.pre_lo = i == 0 ? pre.lo : nullptr;
        p.pre_hi = i == 0 ? pre.hi : nullptr;
        p.post_lo = last ? post.lo : nullptr;
        p.post_hi = last ? post.hi : nullptr;
        p.out_natural = last ? out_natural : 0;
        if (p.out_natural && p.in == p.out) return fail(ctx, B200ZK_ERR_ARG, "internal: a natural-order (scattering) pass cannot run in place");
        p.rt_base = 0;
        {
            const char* e = getenv("B200ZK_NTT_PREFETCH");  // experiment knob: CTAs of look-ahead (0 disables)
            p.prefetch_dist = e ? (uint32_t)atoi(e) : 148u; // measured best look-ahead of pass_kernel (profiles/ntt_tuning_r01.txt)
        }
        const uint64_t R = 1ull << K;
        // Passes of K = 6..8 stages run on the direct kernel (B200, LDE 2^23 x 256: 4.75 / 4.55 / 5.2 ms per pass with factors
        // against 5.0 / 4.9 / 6.3 ms on the TMA ring, profiles/ntt_lab_r02.txt; the factor-free K = 8 last pass 3.64 against
        // 4.16 ms, profiles/ntt_fused_mid_r02.txt).  The TMA ring keeps the other K, the natural-order scatter of the unfused
        // inverse and the peer-memory scatter of the sharded LDE.
        static const bool direct_last = [] {
            const char* e = getenv("B200ZK_DIRECT_LAST");  // experiment knob: 0 = the factor-free last pass back on the TMA ring
            return !e || atoi(e) != 0;                     // measured r02, K = 8 last pass at 2^23 x 256: direct 3.64 ms, TMA ring 4.16 ms
        }();
        const bool has_factors = p.pre_lo != nullptr || (n - s0 - K) > 0 || (direct_last && !p.out_natural);
        if (tma_ok && direct_enabled() && has_factors && K >= 6 && K <= 8 && !p.post_lo && !(last && scatter)) {
            // direct pass: 2^13-element tiles, first round from global memory, last round to global memory (ntt.cuh)
            const int lcd = 13 - K;
            p.lc = lcd;
            const uint64_t tiles = ((1ull << n) >> K) * ((width + (1u << lcd) - 1) >> lcd);
            if (tiles > 0x7fffffffull) return fail(ctx, B200ZK_ERR_SHAPE, "too many tiles for one launch");
            const size_t dsm = ((size_t)R << lcd) * 4 + (R / 2) * 8 + 2 * R * 8;
            {
                const char* e = getenv("B200ZK_DIRECT_PREFETCH");  // experiment knob: look-ahead of the direct kernel in CTAs (0 disables)
                p.prefetch_dist = e ? (uint32_t)atoi(e) : (uint32_t)ctx->num_sms;  // measured (profiles/ntt_fused_mid_r02.txt): 148: 36.3 ms, 0: 36.7, 592: 36.8, 1184: 38.8
            }
            {
                PassTimer tm(ctx, inverse ? (p.out_natural ? "inv-scat" : "inv") : (p.pre_lo ? "fwd-pre" : "fwd"), n, s0, K, width);
                if (K == 8) ntt::pass_kernel_direct<8, 5><<<(uint32_t)tiles, ntt::DIRECT_THREADS, dsm, ctx->stream>>>(p);
                else if (K == 7) ntt::pass_kernel_direct<7, 6><<<(uint32_t)tiles, ntt::DIRECT_THREADS, dsm, ctx->stream>>>(p);
                else ntt::pass_kernel_direct<6, 7><<<(uint32_t)tiles, ntt::DIRECT_THREADS, dsm, ctx->stream>>>(p);
            }
            LAUNCHED();
            s0 += K;
            continue;
        }
        if (tma_ok && K <= ntt::TMA_MAX_K && !p.post_lo) {
            // persistent warp-specialised TMA pass: 2^13-element tiles, 3-stage ring, 2 CTAs per SM
            int tl = std::min(ntt::TMA_TILE_LOG - K, ntt::MAX_TILE_COLS_LOG);
            while (tl > 2 && (1u << (tl - 1)) >= width) tl--;
            p.lc = tl;
            const uint32_t tcols = 1u << tl;
            const uint64_t tiles = ((1ull << n) >> K) * ((width + tcols - 1) / tcols);
            if (tiles > 0xffffffffull) return fail(ctx, B200ZK_ERR_SHAPE, "too many tiles for one launch");
            ntt::TensorMap in_map, out_map;
            TRY(make_pass_map(ctx, p.in, width, p.in_pitch, n, s0, K, tl, &in_map));
            if (last && scatter) {
                // one launch per destination: the row tiles [rt0, rt1) of this pass are the rows of one owner (L = 0 here)
                const uint64_t R2 = 1ull << K, rows = 1ull << n;
                if (scatter->mg % R2) return fail(ctx, B200ZK_ERR_SHAPE, "row block smaller than a tile");
                const size_t tsm2 = tma_smem_bytes(K, tcols);
                // row ranges of this pass by owner, visited in an order rotated by the sender's rank: with every rank walking the
                // owners 0, 1, 2, ... all of them store into ONE GPU at a time (its NVLink ingress, 900 GB/s, is shared by world - 1
                // senders while the other links idle; measured at 8 GPUs: 6.5 ms on rank 0, 14 ms on the last rank to finish)
                std::vector<std::pair<uint64_t, uint64_t>> segs;
                for (uint64_t g = scatter->g0; g < scatter->g0 + rows;) {
                    const uint64_t g_end = std::min(scatter->g0 + rows, (g / scatter->mg + 1) * scatter->mg);
                    segs.emplace_back(g, g_end);
                    g = g_end;
                }
                const size_t first_seg = (scatter->rank + 1) % segs.size();  // a coset block covers world / cosets owners: rank and rank + segs.size() share one
                for (size_t si = 0; si < segs.size(); si++) {
                    uint64_t g = segs[(first_seg + si) % segs.size()].first;
                    const uint64_t r = g / scatter->mg;
                    const uint64_t g_end = segs[(first_seg + si) % segs.size()].second;
                    const uint64_t n_rt = (g_end - g) >> K;                                  // a power of two
                    const uint32_t dpitch = scatter->pitch ? scatter->pitch : width;
                    uint32_t* base = scatter->dst[r] + scatter->slot_off + (g - r * scatter->mg) * dpitch;
                    const int nv = log2u(n_rt << K);
                    TRY(make_pass_map(ctx, base, width, dpitch, nv, nv - K, K, tl, &out_map));
                    p.rt_base = (uint32_t)((g - scatter->g0) >> K);
                    const uint64_t ntile = n_rt * ((width + tcols - 1) / tcols);
                    const uint32_t grid2 = (uint32_t)std::min<uint64_t>(ntile, (uint64_t)NTT_TMA_CTAS * ctx->num_sms);
                    ntt::pass_kernel_tma<<<grid2, ntt::TMA_THREADS, tsm2, ctx->stream>>>(in_map, out_map, p, (uint32_t)ntile);
                    LAUNCHED();
                }
                s0 += K;
                continue;
            }
            if (p.out_natural) TRY(make_natural_map(ctx, p.out, width, p.out_pitch, n, K, tl, &out_map));
            else TRY(make_pass_map(ctx, p.out, width, p.out_pitch, n, s0, K, tl, &out_map));
            // (dynamic shared memory limits are raised once per device in configure_kernels)
            const size_t tsm = tma_smem_bytes(K, tcols);
            const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)NTT_TMA_CTAS * ctx->num_sms);
            {
                PassTimer tm(ctx, inverse ? (p.out_natural ? "inv-scat" : "inv") : (p.pre_lo ? "fwd-pre" : "fwd"), n, s0, K, width);
                ntt::pass_kernel_tma<<<grid, ntt::TMA_THREADS, tsm, ctx->stream>>>(in_map, out_map, p, (uint32_t)tiles);
            }
            LAUNCHED();
            s0 += K;
            continue;
        }
        if (last && scatter) return fail(ctx, B200ZK_ERR_SHAPE, "sharded LDE needs the TMA pass (width % 4 == 0, 16-byte aligned buffers)");
        const size_t smem = (R * tile_cols + 2 * std::max<uint64_t>(R / 2, 1) + R) * 4;
        const uint64_t blocks = ((1ull << n) >> K) * col_tiles;
        if (blocks > 0x7fffffffull) return fail(ctx, B200ZK_ERR_SHAPE, "too many tiles for one launch");
        if (vec == 4) {
            ntt::pass_kernel<4><<<(uint32_t)blocks, ntt::THREADS, smem, ctx->stream>>>(p);
        } else {
            ntt::pass_kernel<1><<<(uint32_t)blocks, ntt::THREADS, smem, ctx->stream>>>(p);
        }
        LAUNCHED();
        s0 += K;
    }
    return B200ZK_OK;
}

int check_mat(b200zk_ctx* ctx, const b200zk_mat* m) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!m || !m->d) return fail(ctx, B200ZK_ERR_ARG, "null matrix");
    return B200ZK_OK;
}

__global__ void fill_kernel(uint32_t* out, uint64_t n, uint64_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (uint32_t)(bb::splitmix64(seed ^ i) % bb::P);
}
__global__ void checksum_kernel(const uint32_t* __restrict__ v, uint64_t n, unsigned long long* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i < n; i += stride) acc += bb::splitmix64(i ^ ((uint64_t)v[i] << 32));
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
// dst[r][c] = c < width ? src[r][c] : 0 for c < dst_width (re-pitching: pad an odd-width matrix to a multiple of 4
// columns so it can take the vectorised / TMA NTT path, and compact the result again)
__global__ void repitch_kernel(const uint32_t* __restrict__ src, uint32_t src_pitch, uint32_t width, uint32_t* __restrict__ dst, uint32_t dst_pitch,
                               uint32_t dst_width, uint64_t rows) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * dst_width) return;
    const uint64_t r = i / dst_width;
    const uint32_t col = (uint32_t)(i % dst_width);
    dst[r * dst_pitch + col] = col < width ? src[r * src_pitch + col] : 0u;
}
__global__ void replicate_row_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t rows, uint32_t width) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * width) out[i] = in[i % width];
}

}  // namespace

// ================================================================================================ context
extern "C" {

const char* b200zk_version(void) { return "b200zk 0.1 (sm_100a)"; }

// Function attributes are per device, not per ctx: they are set to the device maximum, so concurrent contexts (one per
// host thread) can only ever write the same value and no launch depends on a value another thread may be changing.
static int configure_kernels(int max_smem_optin) {
    const void* big[] = {(const void*)ntt::pass_kernel_tma, (const void*)ntt::pass_kernel<4>, (const void*)ntt::pass_kernel<1>,
                         (const void*)ntt::pass_kernel_direct<8, 5>, (const void*)ntt::pass_kernel_direct<7, 6>, (const void*)ntt::pass_kernel_direct<6, 7>,
                         (const void*)ntt::lde_mid_kernel<8, 5, 256, 3>, (const void*)ntt::lde_mid_kernel<7, 6, 256, 3>, (const void*)ntt::lde_mid_kernel<6, 7, 256, 3>,
                         (const void*)ntt::lde_mid_kernel<7, 5, 128, 6>, (const void*)ntt::lde_mid_kernel<6, 6, 128, 6>,
                         (const void*)op::dot_ext_powers_kernel<4>, (const void*)op::dot_ext_powers_kernel<1>};
    for (const void* f : big)
        if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin) != cudaSuccess) return B200ZK_ERR_CUDA;
    for (int i = 0; i < 11; i++)
        if (cudaFuncSetAttribute(big[i], cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess) return B200ZK_ERR_CUDA;
    if (const char* e = getenv("B200ZK_HASH_STREAM"))  // co-residency experiment: the absorb kernel asks for the NTT kernels' carve-out
        if (atoi(e) == 2 && cudaFuncSetAttribute((const void*)mk::leaf_absorb_strip_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess)
            return B200ZK_ERR_CUDA;
    return B200ZK_OK;
}

int b200zk_ctx_create(int device, b200zk_ctx** out) {
    if (!out) return B200ZK_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return B200ZK_ERR_CUDA;  // no CUDA device: there is no CPU fallback
    }
    b200zk_ctx* ctx = new (std::nothrow) b200zk_ctx();
    if (!ctx) return B200ZK_ERR_OOM;
    ctx->device = device;
    int prio_least = 0, prio_greatest = 0;
    cudaSetDevice(device);
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    cudaGetLastError();
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
        delete ctx;
        return B200ZK_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (configure_kernels(ctx->max_smem_optin) != B200ZK_OK) {
        cudaGetLastError();
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return B200ZK_ERR_CUDA;
    }
    if (const char* e = getenv("B200ZK_L2_FETCH")) {  // experiment knob: L2 fetch granularity hint (32/64/128)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
        cudaGetLastError();
    }
    {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) ctx->encode_tiled = fn;
        cudaGetLastError();
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) ctx->cache_cap = total_b / 2;
        if (const char* e = getenv("B200ZK_ALLOC_CACHE"))  // experiment knob: 0 disables the block cache
            if (atoi(e) == 0) ctx->cache_cap = 0;
        cudaGetLastError();
    }
    if (cudaMalloc((void**)&ctx->d_small, 65536) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return B200ZK_ERR_OOM;
    }
    *out = ctx;
    return B200ZK_OK;
}

void b200zk_ctx_destroy(b200zk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cache_flush(ctx);
    for (auto& dir : ctx->tw_local)
        for (auto& p : dir) cudaFree(p);
    for (auto& p : ctx->tw_lo) cudaFree(p);
    for (auto& p : ctx->tw_hi) cudaFree(p);
    cudaFree(ctx->tab);
    cudaFree(ctx->mid_sigma);
    cudaFree(ctx->d_small);
    if (ctx->h_root_ring) cudaFreeHost(ctx->h_root_ring);
    if (ctx->hash_stream) {
        cudaStreamDestroy(ctx->hash_stream);
        for (auto& e : ctx->ev_lde) cudaEventDestroy(e);
        cudaEventDestroy(ctx->ev_abs);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        for (int i = 0; i < 4; i++) {
            cudaEventDestroy(ctx->ev_copied[i]);
            cudaEventDestroy(ctx->ev_consumed[i]);
            cudaFree(ctx->strip_buf[i]);
        }
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char* b200zk_last_error(const b200zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int b200zk_ctx_sync(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}
int b200zk_ctx_trim(b200zk_ctx* ctx) {
    if (!ctx) return B200ZK_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    cache_flush(ctx);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
    for (int i = 0; i < 4; i++) {
        cudaFree(ctx->strip_buf[i]);
        ctx->strip_buf[i] = nullptr;
    }
    ctx->strip_buf_bytes = 0;
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    CU(cudaMemPoolTrimTo(pool, 0));
    return B200ZK_OK;
}
void* b200zk_ctx_stream(b200zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t b200zk_kernel_launches(const b200zk_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ================================================================================================ matrices
int b200zk_mat_alloc(b200zk_ctx* ctx, uint64_t rows, uint32_t width, b200zk_mat** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    if (!rows || !width) return fail(ctx, B200ZK_ERR_SHAPE, "empty matrix");
    CU(cudaSetDevice(ctx->device));
    b200zk_mat* m = new (std::nothrow) b200zk_mat();
    if (!m) return B200ZK_ERR_OOM;
    int r = dev_alloc(ctx, rows * width * 4, (void**)&m->d);
    if (r) {
        delete m;
        return r;
    }
    m->rows = rows;
    m->width = width;
    m->owned = true;
    *out = m;
    return B200ZK_OK;
}
int b200zk_mat_upload_into(b200zk_ctx* ctx, const uint32_t* h, b200zk_mat* dst) {
    TRY(check_mat(ctx, dst));
    if (!h) return fail(ctx, B200ZK_ERR_ARG, "null host pointer");
    CU(cudaMemcpyAsync(dst->d, h, dst->rows * dst->width * 4, cudaMemcpyHostToDevice, ctx->stream));
    return B200ZK_OK;
}
int b200zk_mat_upload(b200zk_ctx* ctx, const uint32_t* h, uint64_t rows, uint32_t width, b200zk_mat** out) {
    if (!h) return fail(ctx, B200ZK_ERR_ARG, "null host pointer");
    TRY(b200zk_mat_alloc(ctx, rows, width, out));
    int r = b200zk_mat_upload_into(ctx, h, *out);
    if (r == B200ZK_OK) r = b200zk_ctx_sync(ctx);
    if (r) {
        b200zk_mat_free(ctx, *out);
        *out = nullptr;
    }
    return r;
}
int b200zk_mat_wrap(b200zk_ctx* ctx, uint32_t* d, uint64_t rows, uint32_t width, b200zk_mat** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    if (!d) return fail(ctx, B200ZK_ERR_ARG, "null device pointer");
    if (!rows || !width) return fail(ctx, B200ZK_ERR_SHAPE, "empty matrix");
    b200zk_mat* m = new (std::nothrow) b200zk_mat();
    if (!m) return B200ZK_ERR_OOM;
    m->d = d;
    m->rows = rows;
    m->width = width;
    m->owned = false;
    *out = m;
    return B200ZK_OK;
}
int b200zk_mat_download_rows(b200zk_ctx* ctx, const b200zk_mat* m, uint64_t row0, uint64_t nrows, uint32_t* h) {
    TRY(check_mat(ctx, m));
    if (!h || row0 + nrows > m->rows) return fail(ctx, B200ZK_ERR_ARG, "bad row range");
    CU(cudaMemcpyAsync(h, m->d + row0 * m->width, nrows * m->width * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}
int b200zk_mat_download(b200zk_ctx* ctx, const b200zk_mat* m, uint32_t* h) {
    TRY(check_mat(ctx, m));
    return b200zk_mat_download_rows(ctx, m, 0, m->rows, h);
}
uint64_t b200zk_mat_rows(const b200zk_mat* m) { return m ? m->rows : 0; }
uint32_t b200zk_mat_width(const b200zk_mat* m) { return m ? m->width : 0; }
uint32_t* b200zk_mat_device_ptr(const b200zk_mat* m) { return m ? m->d : nullptr; }
void b200zk_mat_free(b200zk_ctx* ctx, b200zk_mat* m) {
    if (!m) return;
    if (m->owned && m->d) {
        if (ctx) cudaSetDevice(ctx->device);
        dev_free(ctx, m->d);
    }
    delete m;
}
int b200zk_mat_fill(b200zk_ctx* ctx, b200zk_mat* m, uint64_t seed) {
    TRY(check_mat(ctx, m));
    fill_kernel<<<148 * 16, 256, 0, ctx->stream>>>(m->d, m->rows * m->width, seed);
    LAUNCHED();
    return B200ZK_OK;
}
int b200zk_mat_checksum(b200zk_ctx* ctx, const b200zk_mat* m, uint64_t* h_out) {
    TRY(check_mat(ctx, m));
    if (!h_out) return fail(ctx, B200ZK_ERR_ARG, "null output");
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->d_small);
    CU(cudaMemsetAsync(acc, 0, 8, ctx->stream));
    checksum_kernel<<<148 * 16, 256, 0, ctx->stream>>>(m->d, m->rows * m->width, acc);
    LAUNCHED();
    CU(cudaMemcpyAsync(h_out, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}


namespace {
bool mid_enabled() {
    static const int v = [] {
        const char* e = getenv("B200ZK_LDE_FUSED_MID");  // experiment knob: 0 keeps the unfused pass sequence
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

// pass plan of a fused LDE: the inverse runs `rest` then the K_mid stages, every forward transform K_mid then `rest`
int mid_plan(int n, std::vector<int>* rest) {
    const int m = (n + ntt::TMA_MAX_K - 1) / ntt::TMA_MAX_K;
    int km = std::max(6, std::min(8, n - ntt::TMA_MAX_K * (m - 1)));
    static const int km_force = [] {
        const char* e = getenv("B200ZK_MID_K");  // experiment knob: force the stage count of the fused middle (6..8)
        return e ? atoi(e) : 0;
    }();
    if (km_force >= 6 && km_force <= 8 && km_force <= n) km = km_force;
    *rest = make_plan(n - km, ntt::TMA_MAX_K);
    return km;
}

// per-coset scale tables of an LDE: (shift * w'^bitrev(c))^j / N for every coset block c, in ctx->tab
int lde_tables(b200zk_ctx* ctx, int n, uint32_t added_bits, uint32_t shift) {
    const uint64_t N = 1ull << n;
    const uint32_t C = 1u << added_bits;
    const size_t per = lo_words(N) + hi_words(N);
    TRY(ensure_tab(ctx, per * C));
    const uint32_t wprime = bb::two_adic_generator(n + (int)added_bits);
    const uint32_t ninv = bb::inv(bb::to_monty((uint32_t)(N % bb::P)));
    for (uint32_t c = 0; c < C; c++) {
        uint32_t base = bb::mul(shift, bb::pow(wprime, bb::bitrev(c, (int)added_bits)));
        TRY(pow_tables(ctx, ctx->tab + per * c, ctx->tab + per * c + lo_words(N), base, ninv, N));
    }
    if (n >= 6 && C <= (uint32_t)ntt::MID_MAX_COSETS) {  // coset powers of the fused middle: sigma_c[k] = (base_c^(2^(n-K)))^k
        std::vector<int> rest;
        const int km = mid_plan(n, &rest);
        const uint32_t R = 1u << km;
        if (ctx->mid_sigma_pairs < (size_t)C * R) {
            if (ctx->mid_sigma) dev_free(ctx, ctx->mid_sigma);
            ctx->mid_sigma = nullptr;
            ctx->mid_sigma_pairs = 0;
            TRY(dev_alloc(ctx, (size_t)C * R * 8, (void**)&ctx->mid_sigma));
            ctx->mid_sigma_pairs = (size_t)C * R;
        }
        for (uint32_t c = 0; c < C; c++) {
            uint32_t base = bb::mul(shift, bb::pow(wprime, bb::bitrev(c, (int)added_bits)));
            for (int i = 0; i < n - km; i++) base = bb::mul(base, base);
            ntt::mid_sigma_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(ctx->mid_sigma + (size_t)c * R, base, R);
            LAUNCHED();
        }
    }
    return B200ZK_OK;
}

// coset LDE of `width` columns (n >= 1, bit-reversed rows out); src / dst may be column strips of wider matrices
// (pitches in elements).  Needs lde_tables() for the same (n, added_bits, shift) to have been enqueued.
//
// Fused form (vectorisable shapes, n >= 6): inverse passes but the last -> `scratch` (N x width; may alias src when the
// caller owns src, else a temporary); ntt::lde_mid_kernel finishes the inverse, applies the coset powers and runs the first
// forward pass of every coset straight into the coset blocks; the remaining forward passes run in place per block.
// Unfused form: work space is the destination itself: block 0 holds the inverse transform in flight, the natural-order
// coefficients land in the last block, and every coset block is produced from them (the last one in place).
int lde_core(b200zk_ctx* ctx, const uint32_t* src, uint32_t src_pitch, int n, uint32_t width, uint32_t added_bits, uint32_t* dst, uint32_t dst_pitch,
             const Scatter* scatter = nullptr, uint32_t* scratch = nullptr, uint32_t scratch_pitch = 0) {
    const uint64_t N = 1ull << n;
    const uint32_t C = 1u << added_bits;
    const size_t per = lo_words(N) + hi_words(N);
    const bool vec4 = width % 4 == 0 && src_pitch % 4 == 0 && dst_pitch % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0 &&
                      (!scratch || ((uintptr_t)scratch % 16 == 0 && scratch_pitch % 4 == 0));
    std::vector<int> rest;
    const int km = n >= 6 ? mid_plan(n, &rest) : 0;
    bool fused = mid_enabled() && vec4 && ctx->encode_tiled && tma_enabled() && n >= 6 && C <= (uint32_t)ntt::MID_MAX_COSETS && !(scatter && rest.empty()) &&
                 (N >> km) * ((width + (1u << (12 - std::min(km, 7))) - 1) >> (12 - std::min(km, 7))) <= 0x7fffffffull;
    uint32_t* tmp = nullptr;
    if (fused && !rest.empty() && !scratch && dev_alloc(ctx, N * width * 4, (void**)&tmp) != B200ZK_OK) {
        fused = false;  // no room for the N x W scratch (e.g. 2^24 x 512 next to other residents): the unfused form needs none
        ctx->err.clear();
    }
    if (fused) {
        const uint64_t R = 1ull << km;
        const int lcm = 13 - km;
        std::vector<int> inv_plan = rest, fwd_plan = {km};
        inv_plan.push_back(km);
        fwd_plan.insert(fwd_plan.end(), rest.begin(), rest.end());
        int rc = ensure_roots(ctx, n);
        if (rc == B200ZK_OK) rc = ensure_local(ctx, 0, km);
        if (rc == B200ZK_OK) rc = ensure_local(ctx, 1, km);
        if (rc == B200ZK_OK && ctx->mid_sigma_pairs < C * R) rc = fail(ctx, B200ZK_ERR_ARG, "internal: lde_tables was not run for this shape");
        ntt::MidParams mp{};
        const uint32_t* mid_in = src;
        uint32_t mid_pitch = src_pitch;
        if (rc == B200ZK_OK && !rest.empty()) {
            uint32_t* work = scratch;
            uint32_t wpitch = scratch_pitch ? scratch_pitch : width;
            if (!work) {
                work = tmp;
                wpitch = width;
            }
            if (rc == B200ZK_OK)
                rc = run_transform(ctx, src, work, work, n, width, /*inverse=*/1, Scale{}, Scale{}, 0, src_pitch, wpitch, wpitch, nullptr, &inv_plan, 0, inv_plan.size() - 1);
            mid_in = work;
            mid_pitch = wpitch;
        }
        if (rc == B200ZK_OK) {
            mp.in = mid_in;
            mp.out = dst;
            mp.block_stride = N * dst_pitch;
            mp.in_pitch = mid_pitch;
            mp.out_pitch = dst_pitch;
            mp.width = width;
            mp.n = n;
            mp.cosets = (int)C;
            {
                const char* e = getenv("B200ZK_MID_PREFETCH");  // experiment knob: look-ahead of the fused middle in CTAs (0 disables)
                mp.prefetch_dist = e ? (uint32_t)atoi(e) : 0u;
            }
            mp.tw_inv = ctx->tw_local[1][km];
            mp.tw_fwd = ctx->tw_local[0][km];
            mp.tw_lo = ctx->tw_lo[n];
            mp.tw_hi = ctx->tw_hi[n];
            mp.sigma = ctx->mid_sigma;
            for (uint32_t c = 0; c < C; c++) {
                mp.pre_lo[c] = ctx->tab + per * c;
                mp.pre_hi[c] = ctx->tab + per * c + lo_words(N);
            }
            static const int small_tile = [] {
                const char* e = getenv("B200ZK_MID_TILE");  // experiment knob: 12 = 2^12-element tiles, 128 threads, six CTAs per SM (K = 6, 7)
                return e ? atoi(e) == 12 : 1;   // measured: 10.06 vs 10.99 ms at 2^23 x 256 (barrier stalls span 4 warps instead of 8)
            }();
            const bool small = small_tile && km <= 7;
            const int lct = small ? 12 - km : lcm;
            const uint64_t tiles = (N >> km) * ((width + (1u << lct) - 1) >> lct);
            const size_t dsm = 2 * ((size_t)R << lct) * 4 + R * 8 + (size_t)C * R * 8;
            PassTimer tm(ctx, "mid", n, n - km, km, width);
            if (small && km == 7) ntt::lde_mid_kernel<7, 5, 128, 6><<<(uint32_t)tiles, 128, dsm, ctx->stream>>>(mp);
            else if (small) ntt::lde_mid_kernel<6, 6, 128, 6><<<(uint32_t)tiles, 128, dsm, ctx->stream>>>(mp);
            else if (km == 8) ntt::lde_mid_kernel<8, 5, 256, 3><<<(uint32_t)tiles, 256, dsm, ctx->stream>>>(mp);
            else if (km == 7) ntt::lde_mid_kernel<7, 6, 256, 3><<<(uint32_t)tiles, 256, dsm, ctx->stream>>>(mp);
            else ntt::lde_mid_kernel<6, 7, 256, 3><<<(uint32_t)tiles, 256, dsm, ctx->stream>>>(mp);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "lde_mid launch failed");
        }
        for (uint32_t c = 0; c < C && rc == B200ZK_OK && !rest.empty(); c++) {
            uint32_t* blk = dst + (uint64_t)c * N * dst_pitch;
            Scatter sc;
            if (scatter) {
                sc = *scatter;
                sc.g0 = (uint64_t)c * N;
            }
            rc = run_transform(ctx, blk, blk, blk, n, width, /*inverse=*/0, Scale{}, Scale{}, 0, dst_pitch, dst_pitch, dst_pitch, scatter ? &sc : nullptr, &fwd_plan, 1);
        }
        if (tmp) dev_free(ctx, tmp);
        return rc;
    }
    uint32_t* coef = dst + (uint64_t)(C - 1) * N * dst_pitch;
    b200zk_mat* tmp_inv = nullptr;
    uint32_t* inv_work = dst;  // block 0
    uint32_t inv_pitch = dst_pitch;
    if (C == 1) {  // no spare block: the inverse transform needs its own work area
        TRY(b200zk_mat_alloc(ctx, N, width, &tmp_inv));
        inv_work = tmp_inv->d;
        inv_pitch = width;
    }
    int rc = run_transform(ctx, src, inv_work, coef, n, width, /*inverse=*/1, Scale{}, Scale{}, /*out_natural=*/1, src_pitch, inv_pitch, dst_pitch);
    for (uint32_t c = 0; c < C && rc == B200ZK_OK; c++) {  // the block holding the coefficients goes last (in place)
        uint32_t* blk = dst + (uint64_t)c * N * dst_pitch;
        Scale pre{ctx->tab + per * c, ctx->tab + per * c + lo_words(N)};
        Scatter sc;
        if (scatter) {
            sc = *scatter;
            sc.g0 = (uint64_t)c * N;  // this coset block is rows [c N, (c + 1) N) of the LDE
        }
        rc = run_transform(ctx, coef, blk, blk, n, width, /*inverse=*/0, pre, Scale{}, 0, dst_pitch, dst_pitch, dst_pitch, scatter ? &sc : nullptr);
    }
    if (tmp_inv) b200zk_mat_free(ctx, tmp_inv);
    return rc;
}
}  // namespace

// ================================================================================================ NTT / LDE
int b200zk_coset_lde_batch_into(b200zk_ctx* ctx, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift, int bitrev_rows, b200zk_mat* out) {
    TRY(check_mat(ctx, evals));
    TRY(check_mat(ctx, out));
    const uint64_t N = evals->rows;
    const uint32_t W = evals->width;
    if (!is_pow2(N)) return fail(ctx, B200ZK_ERR_SHAPE, "height must be a power of two");
    const int n = log2u(N);
    if (n + (int)added_bits > MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "size exceeds the two-adicity of BabyBear (2^27)");
    if (out->rows != (N << added_bits) || out->width != W) return fail(ctx, B200ZK_ERR_SHAPE, "output matrix has the wrong shape");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    if (out->d == evals->d) return fail(ctx, B200ZK_ERR_ARG, "output aliases input");
    CU(cudaSetDevice(ctx->device));
    const uint32_t C = 1u << added_bits;
    if (n == 0) {  // constant polynomial: every evaluation equals the single input row
        uint64_t tot = (uint64_t)C * W;
        replicate_row_kernel<<<(uint32_t)((tot + 255) / 256), 256, 0, ctx->stream>>>(evals->d, out->d, C, W);
        LAUNCHED();
        return B200ZK_OK;
    }
    uint32_t* final_dst = out->d;
    b200zk_mat* tmp_nat = nullptr;
    if (!bitrev_rows) {  // natural order requested: build bit-reversed in a temporary, permute at the end
        TRY(b200zk_mat_alloc(ctx, out->rows, W, &tmp_nat));
        final_dst = tmp_nat->d;
    }
    int rc = lde_tables(ctx, n, added_bits, shift);
    if (rc == B200ZK_OK && W % 4 != 0 && N * W >= (1ull << 14)) {
        // ragged width (the usual case for real traces): extend a zero-padded copy whose pitch is a multiple of 4 so the
        // passes run vectorised through TMA, then compact the result (two extra streaming copies, ~2.5x faster overall)
        const uint32_t Wp = (W + 3) & ~3u;
        const uint64_t M = N << added_bits;
        uint32_t *pin = nullptr, *pout = nullptr;
        rc = dev_alloc(ctx, N * Wp * 4, (void**)&pin);
        if (rc == B200ZK_OK) rc = dev_alloc(ctx, M * Wp * 4, (void**)&pout);
        if (rc == B200ZK_OK) {
            repitch_kernel<<<(uint32_t)((N * Wp + 255) / 256), 256, 0, ctx->stream>>>(evals->d, W, W, pin, Wp, Wp, N);
            ctx->launches++;
            rc = lde_core(ctx, pin, Wp, n, Wp, added_bits, pout, Wp, nullptr, /*scratch=*/pin, Wp);
        }
        if (rc == B200ZK_OK) {
            repitch_kernel<<<(uint32_t)((M * W + 255) / 256), 256, 0, ctx->stream>>>(pout, Wp, W, final_dst, W, W, M);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "repitch launch failed");
        }
        dev_free(ctx, pin);
        dev_free(ctx, pout);
    } else if (rc == B200ZK_OK) {
        rc = lde_core(ctx, evals->d, W, n, W, added_bits, final_dst, W);
    }
    if (rc == B200ZK_OK && !bitrev_rows) {
        const int nb = n + (int)added_bits;
        const int vec = (W % 4 == 0) ? 4 : 1;
        uint64_t threads = out->rows * ((W + vec - 1) / vec);
        if (vec == 4) ntt::bitrev_rows_kernel<4><<<(uint32_t)((threads + 255) / 256), 256, 0, ctx->stream>>>(final_dst, out->d, nb, W);
        else ntt::bitrev_rows_kernel<1><<<(uint32_t)((threads + 255) / 256), 256, 0, ctx->stream>>>(final_dst, out->d, nb, W);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "bitrev_rows launch failed");
    }
    if (tmp_nat) b200zk_mat_free(ctx, tmp_nat);
    return rc;
}

int b200zk_coset_lde_batch(b200zk_ctx* ctx, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift, int bitrev_rows, b200zk_mat** out) {
    if (!out) return B200ZK_ERR_ARG;
    *out = nullptr;
    TRY(check_mat(ctx, evals));
    if (added_bits > MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "added_bits too large");
    TRY(b200zk_mat_alloc(ctx, evals->rows << added_bits, evals->width, out));
    int rc = b200zk_coset_lde_batch_into(ctx, evals, added_bits, shift, bitrev_rows, *out);
    if (rc) {
        b200zk_mat_free(ctx, *out);
        *out = nullptr;
    }
    return rc;
}

int b200zk_dft_batch(b200zk_ctx* ctx, const b200zk_mat* in, uint32_t shift, int inverse, int bitrev_rows, b200zk_mat** out) {
    if (!out) return B200ZK_ERR_ARG;
    *out = nullptr;
    TRY(check_mat(ctx, in));
    const uint64_t N = in->rows;
    const uint32_t W = in->width;
    if (!is_pow2(N)) return fail(ctx, B200ZK_ERR_SHAPE, "height must be a power of two");
    const int n = log2u(N);
    if (n > MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "size exceeds the two-adicity of BabyBear (2^27)");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    CU(cudaSetDevice(ctx->device));
    TRY(b200zk_mat_alloc(ctx, N, W, out));
    int rc = B200ZK_OK;
    if (n == 0) {
        CU(cudaMemcpyAsync((*out)->d, in->d, (size_t)W * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return rc;
    }
    Scale pre, post;
    const size_t per = lo_words(N) + hi_words(N);
    rc = ensure_tab(ctx, per);
    if (rc == B200ZK_OK && !inverse && shift != bb::ONE) {
        rc = pow_tables(ctx, ctx->tab, ctx->tab + lo_words(N), shift, bb::ONE, N);
        pre = Scale{ctx->tab, ctx->tab + lo_words(N)};
    }
    if (rc == B200ZK_OK && inverse) {
        rc = pow_tables(ctx, ctx->tab, ctx->tab + lo_words(N), bb::inv(shift), bb::inv(bb::to_monty((uint32_t)(N % bb::P))), N);
        post = Scale{ctx->tab, ctx->tab + lo_words(N)};
    }
    const bool natural = !bitrev_rows;
    // Whether the transform takes more than one pass depends on the plan run_transform picks (the TMA plan cuts at 8 stages,
    // the plain one at 9/10), so every natural-order transform with more than one stage gets its own work buffer: the
    // scattering last pass must never run in place (r01 advice: n = 9 was planned [9] here and [5,4] there).
    const bool multi = n > 1;
    b200zk_mat* tmp = nullptr;
    uint32_t* work = (*out)->d;
    if (rc == B200ZK_OK && natural && multi) {
        rc = b200zk_mat_alloc(ctx, N, W, &tmp);
        if (rc == B200ZK_OK) work = tmp->d;
    }
    if (rc == B200ZK_OK) rc = run_transform(ctx, in->d, work, (*out)->d, n, W, inverse, pre, post, natural ? 1 : 0);
    if (tmp) b200zk_mat_free(ctx, tmp);
    if (rc) {
        b200zk_mat_free(ctx, *out);
        *out = nullptr;
    }
    return rc;
}

// ================================================================================================ Poseidon2
int b200zk_poseidon2_permute_dev(b200zk_ctx* ctx, uint32_t* d_states, uint64_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_states) return fail(ctx, B200ZK_ERR_ARG, "null states");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    mk::permute_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(d_states, n, 0);
    LAUNCHED();
    return B200ZK_OK;
}
// same permutation through the straightforward formulation (cross-check of the optimised kernel)
int b200zk_poseidon2_permute_plain_dev(b200zk_ctx* ctx, uint32_t* d_states, uint64_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_states) return fail(ctx, B200ZK_ERR_ARG, "null states");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    mk::permute_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(d_states, n, 1);
    LAUNCHED();
    return B200ZK_OK;
}
int b200zk_poseidon2_permute(b200zk_ctx* ctx, uint32_t* h_states, uint64_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_states) return fail(ctx, B200ZK_ERR_ARG, "null states");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, n * 64, (void**)&d));
    int rc = B200ZK_OK;
    cudaError_t e = cudaMemcpyAsync(d, h_states, n * 64, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = b200zk_poseidon2_permute_dev(ctx, d, n);
    if (e == cudaSuccess && rc == B200ZK_OK) e = cudaMemcpyAsync(h_states, d, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// Descriptors of the matrices hashed into one sponge.  Up to MAX_GROUP of them travel as a kernel parameter; more (p3's
// MerkleTreeMmcs has no limit on matrices per height) go through a device array, returned in *d_ext for the caller to
// release with dev_free once the kernel that reads it has been enqueued (release is stream ordered).
static int make_group(b200zk_ctx* ctx, b200zk_mat* const* mats, const uint32_t* idx, uint32_t cnt, mk::Group* g, void** d_ext) {
    *d_ext = nullptr;
    g->n = (int)cnt;
    g->fast8 = cnt > 0;
    g->ext = nullptr;
    std::vector<mk::MatRef> big;
    if (cnt > (uint32_t)mk::MAX_GROUP) big.resize(cnt);
    for (uint32_t i = 0; i < cnt; i++) {
        const b200zk_mat* m = mats[idx[i]];
        mk::MatRef& r = big.empty() ? g->m[i] : big[i];
        r.ptr = m->d;
        r.width = m->width;
        if (m->width % 8 != 0 || ((uintptr_t)m->d % 32) != 0) g->fast8 = 0;
    }
    if (!big.empty()) {
        TRY(dev_alloc(ctx, big.size() * sizeof(mk::MatRef), d_ext));
        // pageable source: the runtime stages it before returning, so `big` may go out of scope
        cudaError_t e = cudaMemcpyAsync(*d_ext, big.data(), big.size() * sizeof(mk::MatRef), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            dev_free(ctx, *d_ext);
            *d_ext = nullptr;
            return fail(ctx, B200ZK_ERR_CUDA, std::string("descriptor upload: ") + cudaGetErrorString(e));
        }
        g->ext = static_cast<const mk::MatRef*>(*d_ext);
    }
    return B200ZK_OK;
}

int b200zk_hash_rows_dev(b200zk_ctx* ctx, const b200zk_mat* m, uint32_t* d_digests) {
    TRY(check_mat(ctx, m));
    if (!d_digests) return fail(ctx, B200ZK_ERR_ARG, "null digests");
    CU(cudaSetDevice(ctx->device));
    mk::Group g;
    b200zk_mat* one[1] = {const_cast<b200zk_mat*>(m)};
    uint32_t idx0 = 0;
    void* d_ext = nullptr;
    TRY(make_group(ctx, one, &idx0, 1, &g, &d_ext));
    if (g.fast8) mk::leaf_hash_fast_kernel<<<(uint32_t)((m->rows + 255) / 256), 256, 0, ctx->stream>>>(g, m->rows, d_digests);
    else mk::leaf_hash_kernel<<<(uint32_t)((m->rows + 255) / 256), 256, 0, ctx->stream>>>(g, m->rows, d_digests);
    dev_free(ctx, d_ext);
    LAUNCHED();
    return B200ZK_OK;
}
int b200zk_hash_rows(b200zk_ctx* ctx, const b200zk_mat* m, uint32_t* h_digests) {
    TRY(check_mat(ctx, m));
    if (!h_digests) return fail(ctx, B200ZK_ERR_ARG, "null digests");
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, m->rows * 32, (void**)&d));
    int rc = b200zk_hash_rows_dev(ctx, m, d);
    cudaError_t e = cudaSuccess;
    if (rc == B200ZK_OK) e = cudaMemcpyAsync(h_digests, d, m->rows * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}
int b200zk_compress_pairs_dev(b200zk_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, uint64_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_in || !d_out) return fail(ctx, B200ZK_ERR_ARG, "null buffer");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    mk::Group g;
    g.n = 0;
    g.fast8 = 0;
    g.ext = nullptr;
    mk::compress_layer_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, d_out, n, g);
    LAUNCHED();
    return B200ZK_OK;
}
int b200zk_compress_pairs(b200zk_ctx* ctx, const uint32_t* h_in, uint32_t* h_out, uint64_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_in || !h_out) return fail(ctx, B200ZK_ERR_ARG, "null buffer");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, n * 96, (void**)&d));
    cudaError_t e = cudaMemcpyAsync(d, h_in, n * 64, cudaMemcpyHostToDevice, ctx->stream);
    int rc = B200ZK_OK;
    if (e == cudaSuccess) rc = b200zk_compress_pairs_dev(ctx, d, d + 16 * n, n);
    if (e == cudaSuccess && rc == B200ZK_OK) e = cudaMemcpyAsync(h_out, d + 16 * n, n * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// ================================================================================================ Merkle
void b200zk_tree_free(b200zk_ctx* ctx, b200zk_tree* t) {
    if (!t) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (t->ev_done) cudaEventDestroy(t->ev_done);
    dev_free(ctx, t->d_digests);
    dev_free(ctx, t->d_open);
    if (t->owns_mats)
        for (auto* m : t->mats) b200zk_mat_free(ctx, m);
    delete t;
}

// descriptor array for the open kernels: built on the first open of a tree (a commit never needs it, and staging it
// from pageable memory would synchronise the stream in the middle of the FRI commit phase)
static int ensure_open(b200zk_ctx* ctx, b200zk_tree* t) {
    if (t->d_open) return B200ZK_OK;
    const uint32_t k = (uint32_t)t->mats.size();
    TRY(dev_alloc(ctx, sizeof(mk::OpenMat) * k, (void**)&t->d_open));
    std::vector<mk::OpenMat> om(k);
    uint64_t off = 0;
    for (uint32_t i = 0; i < k; i++) {
        om[i].ptr = t->mats[i]->d;
        om[i].width = t->mats[i]->width;
        om[i].shift = t->depth - (uint32_t)log2u(t->mats[i]->rows);
        om[i].out_off = off;
        off += t->mats[i]->width;
    }
    CU(cudaMemcpyAsync(t->d_open, om.data(), sizeof(mk::OpenMat) * k, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // om is a temporary
    return B200ZK_OK;
}

// builds the tree on the stream; the root stays on the device (last digest); no synchronisation
typedef int (*LeafFn)(b200zk_ctx*, void* user, uint32_t* d_leaf_digests);
static int commit_async(b200zk_ctx* ctx, b200zk_mat* const* mats, uint32_t k, int take, b200zk_tree** out, LeafFn leaf_fn = nullptr, void* leaf_user = nullptr) {
    *out = nullptr;
    if (!k || !mats) return fail(ctx, B200ZK_ERR_ARG, "no matrices");
    for (uint32_t i = 0; i < k; i++) {
        TRY(check_mat(ctx, mats[i]));
        if (!is_pow2(mats[i]->rows)) return fail(ctx, B200ZK_ERR_SHAPE, "matrix heights must be powers of two");
    }
    CU(cudaSetDevice(ctx->device));
    std::vector<uint32_t> order(k);
    for (uint32_t i = 0; i < k; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return mats[a]->rows > mats[b]->rows; });
    b200zk_tree* t = new (std::nothrow) b200zk_tree();
    if (!t) return B200ZK_ERR_OOM;
    t->max_h = mats[order[0]]->rows;
    t->depth = (uint32_t)log2u(t->max_h);
    t->mats.assign(mats, mats + k);
    t->owns_mats = false;  // set at the very end so a failed commit never frees the caller's matrices
    int rc = dev_alloc(ctx, (2 * t->max_h - 1) * 32, (void**)&t->d_digests);
    for (uint32_t i = 0; i < k; i++) t->total_width += mats[i]->width;  // open descriptors are built lazily (ensure_open)
    uint64_t off = 0;
    for (uint64_t n = t->max_h; n >= 1; n >>= 1) {
        t->layer_off.push_back(off);
        off += n;
        if (n == 1) break;
    }
    uint32_t pos = 0;
    if (rc == B200ZK_OK) {
        uint32_t g0 = pos;
        while (pos < k && mats[order[pos]]->rows == t->max_h) pos++;
        mk::Group g;
        void* d_ext = nullptr;
        rc = make_group(ctx, mats, order.data() + g0, pos - g0, &g, &d_ext);
        if (rc == B200ZK_OK && leaf_fn) {
            rc = leaf_fn(ctx, leaf_user, t->d_digests);  // the caller produces layer 0 itself (strip pipeline)
        } else if (rc == B200ZK_OK) {
            if (g.fast8) mk::leaf_hash_fast_kernel<<<(uint32_t)((t->max_h + 255) / 256), 256, 0, ctx->stream>>>(g, t->max_h, t->d_digests);
            else mk::leaf_hash_kernel<<<(uint32_t)((t->max_h + 255) / 256), 256, 0, ctx->stream>>>(g, t->max_h, t->d_digests);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "leaf_hash launch failed");
        }
        dev_free(ctx, d_ext);
    }
    uint32_t layer = 0;
    for (uint64_t len = t->max_h; len > 1 && rc == B200ZK_OK; len >>= 1, layer++) {
        const uint64_t n_next = len >> 1;
        uint32_t* prev = t->d_digests + 8 * t->layer_off[layer];
        uint32_t* next = t->d_digests + 8 * t->layer_off[layer + 1];
        if (pos == k && len <= TOP_LAYER) {  // nothing left to inject: one CTA finishes the tree
            mk::compress_top_kernel<<<1, 1024, 0, ctx->stream>>>(prev, (uint32_t)len);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "compress_top launch failed");
            break;
        }
        uint32_t g0 = pos;
        while (pos < k && mats[order[pos]]->rows == n_next) pos++;
        mk::Group g;
        void* d_ext = nullptr;
        rc = make_group(ctx, mats, order.data() + g0, pos - g0, &g, &d_ext);
        if (rc != B200ZK_OK) break;
        mk::compress_layer_kernel<<<(uint32_t)((n_next + 255) / 256), 256, 0, ctx->stream>>>(prev, next, n_next, g);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "compress_layer launch failed");
        dev_free(ctx, d_ext);
    }
    if (rc == B200ZK_OK && pos != k) rc = fail(ctx, B200ZK_ERR_SHAPE, "matrix height not reached while building the tree");
    if (rc != B200ZK_OK) {
        b200zk_tree_free(ctx, t);
        return rc;
    }
    t->owns_mats = take != 0;
    *out = t;
    return B200ZK_OK;
}

int b200zk_tree_root(b200zk_ctx* ctx, const b200zk_tree* t, uint32_t h_root[8]) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!t || !h_root) return fail(ctx, B200ZK_ERR_ARG, "null tree/root");
    if (t->root_cached) {
        memcpy(h_root, t->root_cache, 32);
        return B200ZK_OK;
    }
    if (t->ev_done) {  // asynchronous commit: wait for this tree's own completion point, then keep the root (the pinned slot is recycled)
        CU(cudaEventSynchronize(t->ev_done));
        memcpy(t->root_cache, t->h_root_slot, 32);
        t->root_cached = true;
        cudaEventDestroy(t->ev_done);
        t->ev_done = nullptr;
        t->h_root_slot = nullptr;
        memcpy(h_root, t->root_cache, 32);
        return B200ZK_OK;
    }
    CU(cudaMemcpyAsync(h_root, t->d_digests + 8 * (2 * t->max_h - 2), 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_merkle_commit(b200zk_ctx* ctx, b200zk_mat* const* mats, uint32_t k, int take, uint32_t h_root[8], b200zk_tree** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    TRY(commit_async(ctx, mats, k, take, out));
    if (h_root) {
        int rc = b200zk_tree_root(ctx, *out, h_root);
        if (rc) {
            (*out)->owns_mats = false;
            b200zk_tree_free(ctx, *out);
            *out = nullptr;
            return rc;
        }
    }
    return B200ZK_OK;
}

int b200zk_lde_commit(b200zk_ctx* ctx, b200zk_mat* const* evals, uint32_t k, uint32_t added_bits, const uint32_t* shifts, uint32_t h_root[8],
                      b200zk_tree** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    if (!k || !evals || !shifts) return fail(ctx, B200ZK_ERR_ARG, "no matrices");
    std::vector<b200zk_mat*> ldes(k, nullptr);
    int rc = B200ZK_OK;
    for (uint32_t i = 0; i < k && rc == B200ZK_OK; i++) rc = check_mat(ctx, evals[i]);
    // Real traces are many narrow matrices of a few heights (the reference fixture: 17 AIRs, widths 1..398).  The transform
    // is column-independent, so all matrices of one (height, shift) class are extended as ONE matrix: their columns are
    // gathered side by side into a work matrix (pitch a multiple of 4), extended with full-width tiles, and the result is
    // scattered into each matrix's own tightly packed LDE (the Merkle leaves stay separate row-major matrices).
    std::vector<char> done(k, 0);
    if (rc == B200ZK_OK && cudaSetDevice(ctx->device) != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "cudaSetDevice failed");
    for (uint32_t i = 0; i < k && rc == B200ZK_OK; i++) {
        if (done[i]) continue;
        std::vector<uint32_t> grp;
        uint64_t wt = 0;
        for (uint32_t j = i; j < k; j++)
            if (!done[j] && evals[j]->rows == evals[i]->rows && shifts[j] == shifts[i]) {
                grp.push_back(j);
                wt += evals[j]->width;
            }
        const uint64_t N = evals[i]->rows;
        const uint64_t M = N << added_bits;
        const uint64_t wp = (wt + 3) & ~3ull;
        static const bool group_on = [] {
            const char* e = getenv("B200ZK_LDE_GROUP");  // experiment knob: 0 extends every matrix on its own
            return !e || atoi(e) != 0;
        }();
        if (!group_on) {
            grp.assign(1, i);
            wt = evals[i]->width;
        }
        const bool grouped = group_on && (grp.size() > 1 || evals[i]->width % 4 != 0) && is_pow2(N) && N >= 2 && N * wt >= (1ull << 14) && M * wp * 4 <= (8ull << 30) &&
                             log2u(N) + (int)added_bits <= MAX_LOG && shifts[i] != 0 && shifts[i] < bb::P;
        if (!grouped) {  // a single aligned matrix (the benchmark shape), tiny inputs, and every error case: the plain entry point
            rc = b200zk_coset_lde_batch(ctx, evals[i], added_bits, shifts[i], 1, &ldes[i]);
            done[i] = 1;
            continue;
        }
        const int n = log2u(N);
        uint32_t *pin = nullptr, *pout = nullptr;
        rc = dev_alloc(ctx, N * wp * 4, (void**)&pin);
        if (rc == B200ZK_OK) rc = dev_alloc(ctx, M * wp * 4, (void**)&pout);
        if (rc == B200ZK_OK) rc = lde_tables(ctx, n, added_bits, shifts[i]);
        uint64_t off = 0;
        for (size_t g = 0; g < grp.size() && rc == B200ZK_OK; g++) {
            const b200zk_mat* m = evals[grp[g]];
            const uint32_t dw = m->width + (g + 1 == grp.size() ? (uint32_t)(wp - wt) : 0u);  // the last one also writes the zero padding
            repitch_kernel<<<(uint32_t)((N * dw + 255) / 256), 256, 0, ctx->stream>>>(m->d, m->width, m->width, pin + off, (uint32_t)wp, dw, N);
            ctx->launches++;
            off += m->width;
        }
        if (rc == B200ZK_OK) rc = lde_core(ctx, pin, (uint32_t)wp, n, (uint32_t)wp, added_bits, pout, (uint32_t)wp, nullptr, /*scratch=*/pin, (uint32_t)wp);
        off = 0;
        for (size_t g = 0; g < grp.size() && rc == B200ZK_OK; g++) {
            const uint32_t j = grp[g];
            rc = b200zk_mat_alloc(ctx, M, evals[j]->width, &ldes[j]);
            if (rc != B200ZK_OK) break;
            const uint32_t w = evals[j]->width;
            repitch_kernel<<<(uint32_t)((M * w + 255) / 256), 256, 0, ctx->stream>>>(pout + off, (uint32_t)wp, w, ldes[j]->d, w, w, M);
            ctx->launches++;
            off += w;
        }
        if (rc == B200ZK_OK && cudaGetLastError() != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, "column gather/scatter launch failed");
        dev_free(ctx, pin);
        dev_free(ctx, pout);
        for (uint32_t j : grp) done[j] = 1;
    }
    if (rc == B200ZK_OK) rc = b200zk_merkle_commit(ctx, ldes.data(), k, /*take=*/1, h_root, out);
    if (rc != B200ZK_OK)
        for (auto* m : ldes) b200zk_mat_free(ctx, m);
    return rc;
}


// ---- TwoAdicFriPcs::commit of ONE host-resident trace with the transfer hidden behind the arithmetic -------------------
// Everything before tree building is column-local (NTT) or column-sequential (sponge), so the trace is processed in
// column strips: strip s+1 crosses PCIe on the copy stream while strip s is extended and absorbed on the compute stream.
namespace {
struct StripJob {
    const uint32_t* h_values;
    uint64_t N;
    uint32_t W, added_bits, shift;
    std::vector<uint32_t> strips;  // column count of every strip, in order (multiples of 16)
    b200zk_mat* lde;
};
int strip_pipeline(b200zk_ctx* ctx, void* user, uint32_t* d_digests) {
    StripJob& j = *static_cast<StripJob*>(user);
    const int n = log2u(j.N);
    const uint64_t M = j.N << j.added_bits;
    if (!ctx->copy_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) {
            CU(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming));
        }
    }
    const uint32_t max_strip = *std::max_element(j.strips.begin(), j.strips.end());
    const size_t need = j.N * max_strip * 4;
    if (ctx->strip_buf_bytes < need) {  // grow: nothing may be in flight on the old buffers
        CU(cudaStreamSynchronize(ctx->copy_stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < 4; i++) {
            cudaFree(ctx->strip_buf[i]);
            ctx->strip_buf[i] = nullptr;
        }
        ctx->strip_buf_bytes = 0;
        for (int i = 0; i < 4; i++) {
            cudaError_t e = cudaMalloc((void**)&ctx->strip_buf[i], need);
            if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try once more
                cudaGetLastError();
                cache_flush(ctx);
                cudaStreamSynchronize(ctx->stream);
                e = cudaMalloc((void**)&ctx->strip_buf[i], need);
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int k = 0; k < 4; k++) {
                    cudaFree(ctx->strip_buf[k]);
                    ctx->strip_buf[k] = nullptr;
                }
                return fail(ctx, B200ZK_ERR_OOM, "strip buffers: " + std::string(cudaGetErrorString(e)));
            }
        }
        ctx->strip_buf_bytes = need;
    }
    static const int hash_mode = [] {
        // experiment knob: 1 = absorb strip s on a second (low-priority) stream while strip s+1 is extended; 2 = the same with a small
        // persistent absorb grid (B200ZK_HASH_CTAS per SM, default 2) so that the NTT kernels find room on every SM next to it
        const char* e = getenv("B200ZK_HASH_STREAM");
        return e ? atoi(e) : 0;
    }();
    static const int hash_ctas = [] {
        const char* e = getenv("B200ZK_HASH_CTAS");
        return e ? std::max(1, atoi(e)) : 2;
    }();
    const bool hash_side = hash_mode != 0;
    if (hash_side && !ctx->hash_stream) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        CU(cudaStreamCreateWithPriority(&ctx->hash_stream, cudaStreamNonBlocking, least));
        for (auto& e : ctx->ev_lde) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_abs, cudaEventDisableTiming));
    }
    const int set = (int)(ctx->strip_calls++ & 1) * 2;  // consecutive calls alternate between the two buffer sets
    uint32_t* cap = nullptr;
    int rc = dev_alloc(ctx, M * 32, (void**)&cap);
    if (rc == B200ZK_OK) rc = lde_tables(ctx, n, j.added_bits, j.shift);
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == B200ZK_OK) rc = fail(ctx, B200ZK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    const size_t strips = j.strips.size();
    bool copies_in_flight = false;
    uint32_t col0 = 0;
    for (size_t s = 0; s < strips && rc == B200ZK_OK; s++) {
        const int b = set + (int)(s & 1);
        uint32_t* sb = ctx->strip_buf[b];
        const uint32_t sw = j.strips[s];
        // the buffer's previous contents (two strips ago, or two calls ago) must have been consumed; a never-recorded event is a no-op
        if (!cuda_ok(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[b], 0), "stream wait")) break;
        if (!cuda_ok(cudaMemcpy2DAsync(sb, (size_t)sw * 4, j.h_values + col0, (size_t)j.W * 4, (size_t)sw * 4, j.N, cudaMemcpyHostToDevice, ctx->copy_stream), "strip copy")) break;
        copies_in_flight = true;
        if (!cuda_ok(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream), "event record")) break;
        if (!cuda_ok(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0), "stream wait")) break;
        rc = lde_core(ctx, sb, sw, n, sw, j.added_bits, j.lde->d + col0, j.W, nullptr, /*scratch=*/sb, sw);
        if (rc != B200ZK_OK) break;
        if (!cuda_ok(cudaEventRecord(ctx->ev_consumed[b], ctx->stream), "event record")) break;
        cudaStream_t hs = ctx->stream;
        if (hash_side) {
            hs = ctx->hash_stream;
            if (!cuda_ok(cudaEventRecord(ctx->ev_lde[s & 7], ctx->stream), "event record")) break;
            if (!cuda_ok(cudaStreamWaitEvent(hs, ctx->ev_lde[s & 7], 0), "stream wait")) break;
        }
        const uint32_t agrid = hash_mode == 2 ? (uint32_t)std::min<uint64_t>((M + 255) / 256, (uint64_t)hash_ctas * ctx->num_sms) : (uint32_t)((M + 255) / 256);
        mk::leaf_absorb_strip_kernel<<<agrid, 256, 0, hs>>>(j.lde->d, j.W, col0, sw, M, cap, s == 0, s + 1 == strips, d_digests);
        ctx->launches++;
        cuda_ok(cudaGetLastError(), "leaf_absorb_strip launch");
        col0 += sw;
    }
    if (hash_side && rc == B200ZK_OK) {  // the tree is built on the main stream, behind the last absorb
        cuda_ok(cudaEventRecord(ctx->ev_abs, ctx->hash_stream), "event record");
        cuda_ok(cudaStreamWaitEvent(ctx->stream, ctx->ev_abs, 0), "stream wait");
    }
    // on an error path copies from the caller's h_values may still be queued on the copy stream: drain it before the
    // buffers go back to the allocator and before the caller is told it may release the host memory
    if (rc != B200ZK_OK && copies_in_flight) cudaStreamSynchronize(ctx->copy_stream);
    dev_free(ctx, cap);
    return rc;
}
}  // namespace

static int lde_commit_host_impl(b200zk_ctx* ctx, const uint32_t* h_values, uint64_t rows, uint32_t width, uint32_t added_bits, uint32_t shift, uint32_t strip_cols,
                                uint32_t h_root[8], b200zk_tree** out, bool async) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    if (!h_values) return fail(ctx, B200ZK_ERR_ARG, "null host pointer");
    if (!rows || !width || !is_pow2(rows)) return fail(ctx, B200ZK_ERR_SHAPE, "height must be a power of two");
    const int n = log2u(rows);
    if (n + (int)added_bits > MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "size exceeds the two-adicity of BabyBear (2^27)");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    CU(cudaSetDevice(ctx->device));
    // Strip width.  A strip row is one DMA segment: 128-byte segments (32 columns) reach ~86 % of the contiguous copy rate,
    // 256-byte segments (64 columns) all of it, 64-byte segments under half.  Inside ONE blocking call the first strip's
    // transfer and the last strip's arithmetic are exposed, which favours narrow strips (measured at 2^23 x 256, r01: 16: 341 ms,
    // 32: 222 ms, 64: 233 ms, 128: 270 ms; r02: 32: 203 ms, 64: 205 ms).  A stream of asynchronous calls hides both ends under
    // the neighbouring calls, so only the copy rate matters: 64.  (A graded schedule 16 16 32 64 64 32 16 16 was tried in r02
    // and is slower than either, 227 ms: with copy and arithmetic this close to each other per column, any strip whose copy
    // is longer than the previous strip's arithmetic stalls the compute stream -- profiles/e2e_strip_schedule_r02.txt.)
    std::vector<uint32_t> sched;
    {
        uint32_t strip = strip_cols ? strip_cols : (async ? 64 : 32);
        while (strip > 16 && (width % strip || width / strip < 2)) strip >>= 1;
        if (width % strip == 0 && strip % 16 == 0 && width / strip >= 2) sched.assign(width / strip, strip);
    }
    const bool pipelined = n >= 1 && added_bits >= 1 && !sched.empty() && rows * (uint64_t)width >= (1ull << 22);
    if (!pipelined) {  // small or ragged: upload (synchronous: h_values is free again on return), extend, commit
        b200zk_mat* m = nullptr;
        TRY(b200zk_mat_upload(ctx, h_values, rows, width, &m));
        b200zk_mat* arr[1] = {m};
        int rc = b200zk_lde_commit(ctx, arr, 1, added_bits, &shift, async ? nullptr : h_root, out);
        b200zk_mat_free(ctx, m);
        return rc;
    }
    b200zk_mat* lde = nullptr;
    TRY(b200zk_mat_alloc(ctx, rows << added_bits, width, &lde));
    StripJob job{h_values, rows, width, added_bits, shift, sched, lde};
    b200zk_mat* arr[1] = {lde};
    int rc = commit_async(ctx, arr, 1, /*take=*/1, out, strip_pipeline, &job);
    if (rc == B200ZK_OK && async) {  // everything is enqueued; b200zk_tree_root waits for this tree's completion point
        b200zk_tree* t = *out;
        cudaError_t e = cudaSuccess;
        if (!ctx->h_root_ring) e = cudaMallocHost((void**)&ctx->h_root_ring, 64 * 32);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) {
            t->h_root_slot = ctx->h_root_ring + 8 * (ctx->root_ring_next++ & 63);
            e = cudaMemcpyAsync(t->h_root_slot, t->d_digests + 8 * (2 * t->max_h - 2), 32, cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(t->ev_done, ctx->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(ctx->stream);  // h_values may be released by the caller after an error
            if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
            b200zk_tree_free(ctx, t);
            *out = nullptr;
            return fail(ctx, B200ZK_ERR_CUDA, std::string("async commit: ") + cudaGetErrorString(e));
        }
        return rc;
    }
    if (rc == B200ZK_OK && h_root) rc = b200zk_tree_root(ctx, *out, h_root);
    else if (rc == B200ZK_OK) rc = b200zk_ctx_sync(ctx);  // h_values must stay valid until the copies are done
    if (rc != B200ZK_OK) {
        if (*out) {
            b200zk_tree_free(ctx, *out);  // frees the LDE with it
            *out = nullptr;
        } else {
            b200zk_mat_free(ctx, lde);
        }
    }
    return rc;
}

int b200zk_lde_commit_host(b200zk_ctx* ctx, const uint32_t* h_values, uint64_t rows, uint32_t width, uint32_t added_bits, uint32_t shift, uint32_t strip_cols,
                           uint32_t h_root[8], b200zk_tree** out) {
    return lde_commit_host_impl(ctx, h_values, rows, width, added_bits, shift, strip_cols, h_root, out, false);
}
// The same, returning as soon as the copies and kernels are enqueued: a prover that commits one trace after another (the
// segments of a chunk proof) issues call i + 1 before it reads the root of call i, so the first strip's transfer -- the one
// part of the pipeline nothing can hide inside a single call -- runs under the previous call's arithmetic.
int b200zk_lde_commit_host_async(b200zk_ctx* ctx, const uint32_t* h_values, uint64_t rows, uint32_t width, uint32_t added_bits, uint32_t shift,
                                 uint32_t strip_cols, b200zk_tree** out) {
    return lde_commit_host_impl(ctx, h_values, rows, width, added_bits, shift, strip_cols, nullptr, out, true);
}

uint32_t b200zk_tree_depth(const b200zk_tree* t) { return t ? t->depth : 0; }
uint32_t b200zk_tree_num_mats(const b200zk_tree* t) { return t ? (uint32_t)t->mats.size() : 0; }
uint64_t b200zk_tree_total_width(const b200zk_tree* t) { return t ? t->total_width : 0; }
const b200zk_mat* b200zk_tree_mat(const b200zk_tree* t, uint32_t i) { return (t && i < t->mats.size()) ? t->mats[i] : nullptr; }
int b200zk_tree_download_layer(b200zk_ctx* ctx, const b200zk_tree* t, uint32_t layer, uint32_t* h) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!t || !h || layer > t->depth) return fail(ctx, B200ZK_ERR_ARG, "bad layer");
    CU(cudaMemcpyAsync(h, t->d_digests + 8 * t->layer_off[layer], (t->max_h >> layer) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int b200zk_merkle_open(b200zk_ctx* ctx, const b200zk_tree* t, uint64_t index, uint32_t* h_rows, uint32_t* h_path) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!t || !h_rows || (!h_path && t->depth)) return fail(ctx, B200ZK_ERR_ARG, "null tree/output");
    if (index >= t->max_h) return fail(ctx, B200ZK_ERR_ARG, "index out of range");
    CU(cudaSetDevice(ctx->device));
    TRY(ensure_open(ctx, const_cast<b200zk_tree*>(t)));
    uint32_t* d = nullptr;
    const size_t words = t->total_width + 8ull * t->depth;
    TRY(dev_alloc(ctx, words * 4, (void**)&d));
    const uint32_t k = (uint32_t)t->mats.size();
    mk::open_rows_kernel<<<std::min<uint32_t>(k, 1024), 128, 0, ctx->stream>>>(t->d_open, k, index, d);
    ctx->launches++;
    if (t->depth) {
        mk::open_path_kernel<<<t->depth, 32, 0, ctx->stream>>>(t->d_digests, t->max_h, t->depth, index, d + t->total_width);
        ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_rows, d, t->total_width * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && t->depth) e = cudaMemcpyAsync(h_path, d + t->total_width, 32ull * t->depth, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return B200ZK_OK;
}

int b200zk_merkle_open_many(b200zk_ctx* ctx, const b200zk_tree* t, const uint64_t* h_indices, uint32_t n_idx, uint32_t* h_rows, uint32_t* h_paths) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!t || !h_indices || !h_rows || (!h_paths && t->depth)) return fail(ctx, B200ZK_ERR_ARG, "null tree/output");
    if (!n_idx) return B200ZK_OK;
    if (n_idx > 65535) return fail(ctx, B200ZK_ERR_ARG, "at most 65535 indices per call");
    for (uint32_t i = 0; i < n_idx; i++)
        if (h_indices[i] >= t->max_h) return fail(ctx, B200ZK_ERR_ARG, "index out of range");
    CU(cudaSetDevice(ctx->device));
    TRY(ensure_open(ctx, const_cast<b200zk_tree*>(t)));
    const size_t row_words = (size_t)t->total_width * n_idx, path_words = 8ull * t->depth * n_idx;
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, (row_words + path_words) * 4 + 8ull * n_idx + 16, (void**)&d));
    uint64_t* d_idx = reinterpret_cast<uint64_t*>(d + ((row_words + path_words + 1) & ~(size_t)1));
    cudaError_t e = cudaMemcpyAsync(d_idx, h_indices, 8ull * n_idx, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const uint32_t k = (uint32_t)t->mats.size();
        dim3 grid(std::min<uint32_t>(k, 64), n_idx);
        mk::open_many_kernel<<<grid, 128, 0, ctx->stream>>>(t->d_open, k, d_idx, n_idx, t->total_width, t->d_digests, t->max_h, t->depth, d, d + row_words);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_rows, d, row_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && t->depth) e = cudaMemcpyAsync(h_paths, d + row_words, path_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return B200ZK_OK;
}

// the FRI query phase over ALL commit-phase trees at once: tree r is opened at (index >> r) >> 1 for every query index.
// One launch per tree, one upload of the indices and one download of everything (the per-tree call pays a
// synchronisation per round: 20 rounds x 100 queries cost 4 ms of a 45 ms segment that way).
//   h_pairs: n_trees x n_idx x 8 (the opened (lo, hi) EF4 pair), h_paths: tree after tree, n_idx x depth_r x 8
int b200zk_fri_open_queries(b200zk_ctx* ctx, const b200zk_tree* const* trees, uint32_t n_trees, const uint64_t* h_indices, uint32_t n_idx, uint32_t* h_pairs,
                            uint32_t* h_paths) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!n_trees || !n_idx) return B200ZK_OK;
    if (!trees || !h_indices || !h_pairs || !h_paths) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (n_idx > 65535) return fail(ctx, B200ZK_ERR_ARG, "at most 65535 indices per call");
    CU(cudaSetDevice(ctx->device));
    std::vector<uint64_t> idx((size_t)n_trees * n_idx);
    size_t path_words = 0;
    for (uint32_t r = 0; r < n_trees; r++) {
        const b200zk_tree* t = trees[r];
        if (!t || t->mats.size() != 1 || t->total_width != 8) return fail(ctx, B200ZK_ERR_ARG, "not a FRI commit-phase tree");
        for (uint32_t q = 0; q < n_idx; q++) {
            const uint64_t i = r < 63 ? (h_indices[q] >> r) >> 1 : 0;
            if (i >= t->max_h) return fail(ctx, B200ZK_ERR_ARG, "index out of range");
            idx[(size_t)r * n_idx + q] = i;
        }
        path_words += 8ull * t->depth * n_idx;
        TRY(ensure_open(ctx, const_cast<b200zk_tree*>(t)));
    }
    const size_t pair_words = 8ull * n_trees * n_idx;
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, (pair_words + path_words) * 4 + 8ull * idx.size() + 16, (void**)&d));
    uint64_t* d_idx = reinterpret_cast<uint64_t*>(d + ((pair_words + path_words + 1) & ~(size_t)1));
    cudaError_t e = cudaMemcpyAsync(d_idx, idx.data(), 8ull * idx.size(), cudaMemcpyHostToDevice, ctx->stream);  // pageable: staged before returning
    size_t poff = 0;
    for (uint32_t r = 0; r < n_trees && e == cudaSuccess; r++) {
        const b200zk_tree* t = trees[r];
        dim3 grid(1, n_idx);
        mk::open_many_kernel<<<grid, 128, 0, ctx->stream>>>(t->d_open, 1, d_idx + (size_t)r * n_idx, n_idx, 8, t->d_digests, t->max_h, t->depth,
                                                             d + 8ull * r * n_idx, d + pair_words + poff);
        ctx->launches++;
        poff += 8ull * t->depth * n_idx;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_pairs, d, pair_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && path_words) e = cudaMemcpyAsync(h_paths, d + pair_words, path_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    return B200ZK_OK;
}

int b200zk_merkle_verify(b200zk_ctx* ctx, const uint32_t* h_rows, const uint64_t* heights, const uint32_t* widths, uint32_t k, const uint32_t* h_path,
                         uint32_t depth, uint64_t index, const uint32_t h_root[8], int* h_ok) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_rows || !heights || !widths || !k || !h_root || !h_ok || (depth && !h_path)) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    *h_ok = 0;
    std::vector<uint32_t> order(k);
    std::vector<uint64_t> offs(k);
    uint64_t tot = 0;
    for (uint32_t i = 0; i < k; i++) {
        if (!is_pow2(heights[i])) return fail(ctx, B200ZK_ERR_SHAPE, "matrix heights must be powers of two");
        order[i] = i;
        offs[i] = tot;
        tot += widths[i];
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return heights[a] > heights[b]; });
    if ((uint32_t)log2u(heights[order[0]]) != depth) return fail(ctx, B200ZK_ERR_SHAPE, "path length does not match the tallest matrix");
    // pack: rows (sorted order) | widths | log_heights | path | root | ok
    std::vector<uint32_t> pack;
    pack.reserve(tot + 2 * k + 8 * depth + 16);
    for (uint32_t i = 0; i < k; i++) pack.insert(pack.end(), h_rows + offs[order[i]], h_rows + offs[order[i]] + widths[order[i]]);
    const size_t o_w = pack.size();
    for (uint32_t i = 0; i < k; i++) pack.push_back(widths[order[i]]);
    const size_t o_h = pack.size();
    for (uint32_t i = 0; i < k; i++) pack.push_back((uint32_t)log2u(heights[order[i]]));
    const size_t o_p = pack.size();
    pack.insert(pack.end(), h_path, h_path + 8 * (size_t)depth);
    const size_t o_r = pack.size();
    pack.insert(pack.end(), h_root, h_root + 8);
    const size_t o_ok = pack.size();
    pack.push_back(0);
    CU(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    TRY(dev_alloc(ctx, pack.size() * 4, (void**)&d));
    cudaError_t e = cudaMemcpyAsync(d, pack.data(), pack.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        mk::VerifyArgs a{d, d + o_w, d + o_h, k, d + o_p, depth, index, d + o_r, reinterpret_cast<int*>(d + o_ok)};
        mk::verify_kernel<<<1, 32, 0, ctx->stream>>>(a);
        ctx->launches++;
        e = cudaGetLastError();
    }
    uint32_t ok = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ok, d + o_ok, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    *h_ok = (int)ok;
    return B200ZK_OK;
}

// ================================================================================================ challenger
int b200zk_chal_create(b200zk_ctx* ctx, b200zk_chal** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    b200zk_chal* c = new (std::nothrow) b200zk_chal();
    if (!c) return B200ZK_ERR_OOM;
    int rc = dev_alloc(ctx, sizeof(fri::ChalState), (void**)&c->d);
    if (rc) {
        delete c;
        return rc;
    }
    CU(cudaMemsetAsync(c->d, 0, sizeof(fri::ChalState), ctx->stream));
    *out = c;
    return B200ZK_OK;
}
void b200zk_chal_free(b200zk_ctx* ctx, b200zk_chal* c) {
    if (!c) return;
    dev_free(ctx, c->d);
    delete c;
}
int b200zk_chal_observe(b200zk_ctx* ctx, b200zk_chal* c, const uint32_t* h_values, uint32_t n) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!c || (!h_values && n)) return fail(ctx, B200ZK_ERR_ARG, "null challenger/values");
    if (!n) return B200ZK_OK;
    if (n > 8192) return fail(ctx, B200ZK_ERR_ARG, "observe at most 8192 elements per call");
    for (uint32_t i = 0; i < n; i++)
        if (h_values[i] >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "value is not a reduced field element");
    CU(cudaMemcpyAsync(ctx->d_small, h_values, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    fri::chal_observe_kernel<<<1, 32, 0, ctx->stream>>>(c->d, ctx->d_small, n);
    LAUNCHED();
    CU(cudaStreamSynchronize(ctx->stream));  // d_small is reused by the next call
    return B200ZK_OK;
}
static int chal_sample_impl(b200zk_ctx* ctx, b200zk_chal* c, uint32_t* h_out, uint32_t n, uint32_t bits) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!c || !h_out) return fail(ctx, B200ZK_ERR_ARG, "null challenger/output");
    if (n > 8192 || bits > 31) return fail(ctx, B200ZK_ERR_ARG, "bad sample request");
    if (!n) return B200ZK_OK;
    fri::chal_sample_kernel<<<1, 32, 0, ctx->stream>>>(c->d, ctx->d_small, n, bits);
    LAUNCHED();
    CU(cudaMemcpyAsync(h_out, ctx->d_small, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}
int b200zk_chal_sample(b200zk_ctx* ctx, b200zk_chal* c, uint32_t* h_out, uint32_t n) { return chal_sample_impl(ctx, c, h_out, n, 0); }
int b200zk_chal_sample_bits(b200zk_ctx* ctx, b200zk_chal* c, uint32_t bits, uint32_t* h_out) {
    if (bits == 0) {  // sample_bits(0) still consumes a sample in p3; result is 0
        uint32_t tmp;
        int rc = chal_sample_impl(ctx, c, &tmp, 1, 0);
        if (rc == B200ZK_OK && h_out) *h_out = 0;
        return rc;
    }
    return chal_sample_impl(ctx, c, h_out, 1, bits);
}
int b200zk_chal_grind(b200zk_ctx* ctx, b200zk_chal* c, uint32_t bits, uint32_t* h_witness) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!c || !h_witness || bits > 30) return fail(ctx, B200ZK_ERR_ARG, "bad grind request");
    if (bits == 0) {  // p3-challenger 0.4.3 (the reference's pin, Cargo.lock:5576): no proof of work, the transcript is left untouched
        *h_witness = 0;
        return B200ZK_OK;
    }
    uint32_t* best = ctx->d_small;
    uint32_t init = 0xffffffffu;
    CU(cudaMemcpyAsync(best, &init, 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t batch = 1u << 20;
    uint32_t found = init;
    for (uint64_t base = 0; base < bb::P && found == init; base += batch) {
        uint32_t cnt = (uint32_t)std::min<uint64_t>(batch, bb::P - base);
        fri::chal_grind_kernel<<<(cnt + 255) / 256, 256, 0, ctx->stream>>>(c->d, bits, (uint32_t)base, cnt, best);
        LAUNCHED();
        CU(cudaMemcpyAsync(&found, best, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (found == init) return fail(ctx, B200ZK_ERR_ARG, "no proof-of-work witness exists");
    fri::chal_observe_witness_kernel<<<1, 32, 0, ctx->stream>>>(c->d, best, bits);
    LAUNCHED();
    CU(cudaStreamSynchronize(ctx->stream));
    *h_witness = found;
    return B200ZK_OK;
}
int b200zk_chal_state(b200zk_ctx* ctx, const b200zk_chal* c, uint32_t h_state[34]) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!c || !h_state) return fail(ctx, B200ZK_ERR_ARG, "null challenger/output");
    CU(cudaMemcpyAsync(h_state, c->d, sizeof(fri::ChalState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

// the inverse of b200zk_chal_state: load a host DuplexChallenger's fields, so a transcript driven on the host can hand over to
// the device for the FRI commit phase and take the advanced state back afterwards (bindings/b200zk-p3/src/challenger.rs)
int b200zk_chal_set_state(b200zk_ctx* ctx, b200zk_chal* c, const uint32_t h_state[34]) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!c || !h_state) return fail(ctx, B200ZK_ERR_ARG, "null challenger/state");
    if (h_state[24] > 8 || h_state[33] > 8) return fail(ctx, B200ZK_ERR_ARG, "buffer fill counts must be <= 8");
    for (int i = 0; i < 34; i++)
        if (i != 24 && i != 33 && h_state[i] >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "state word is not a reduced field element");
    CU(cudaMemcpyAsync(c->d, h_state, sizeof(fri::ChalState), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // h_state may be a temporary
    return B200ZK_OK;
}

// ================================================================================================ FRI
static int fold_launch(b200zk_ctx* ctx, const uint32_t* d_in, uint64_t len, const uint32_t* d_beta, const uint32_t* d_add, int add_mode, uint32_t* tab,
                       uint32_t* d_out) {
    const uint64_t half = len / 2;
    const int lh = log2u(half);
    // powers of g^-1 (g generates the size-len subgroup) scaled by 1/2
    const uint32_t ginv = bb::inv(bb::two_adic_generator(lh + 1));
    TRY(pow_tables(ctx, tab, tab + 4096, ginv, bb::HALF, std::max<uint64_t>(half, 1)));
    fri::fold_kernel<<<(uint32_t)((half + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(d_in), half, lh, d_beta, tab, tab + 4096,
                                                                              reinterpret_cast<const uint4*>(d_add), d_add ? add_mode : 0,
                                                                              reinterpret_cast<uint4*>(d_out));
    LAUNCHED();
    return B200ZK_OK;
}

int b200zk_fri_commit_layer(b200zk_ctx* ctx, const uint32_t* d_folded, uint64_t len, uint32_t h_root[8], b200zk_tree** out) {
    if (!ctx || !out) return B200ZK_ERR_ARG;
    *out = nullptr;
    if (!d_folded) return fail(ctx, B200ZK_ERR_ARG, "null input");
    if (!is_pow2(len) || len < 2) return fail(ctx, B200ZK_ERR_SHAPE, "length must be a power of two >= 2");
    b200zk_mat* m = nullptr;
    TRY(b200zk_mat_wrap(ctx, const_cast<uint32_t*>(d_folded), len / 2, 8, &m));
    b200zk_mat* arr[1] = {m};
    int rc = b200zk_merkle_commit(ctx, arr, 1, /*take=*/1, h_root, out);  // the handle (not the memory) belongs to the tree
    if (rc) b200zk_mat_free(ctx, m);
    return rc;
}

int b200zk_fri_fold_layer(b200zk_ctx* ctx, const uint32_t* d_folded, uint64_t len, const uint32_t h_beta[4], const uint32_t* d_add, uint32_t* d_out) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_folded || !h_beta || !d_out) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (!is_pow2(len) || len < 2) return fail(ctx, B200ZK_ERR_SHAPE, "length must be a power of two >= 2");
    if ((len >> 1) > (1ull << MAX_LOG)) return fail(ctx, B200ZK_ERR_SHAPE, "length exceeds two-adicity");
    CU(cudaSetDevice(ctx->device));
    const uint64_t half = len / 2;
    TRY(ensure_tab(ctx, 4096 + hi_words(half) + 8));
    uint32_t* d_beta = ctx->tab + 4096 + hi_words(half);
    CU(cudaMemcpyAsync(d_beta, h_beta, 16, cudaMemcpyHostToDevice, ctx->stream));
    TRY(fold_launch(ctx, d_folded, len, d_beta, d_add, 1, ctx->tab, d_out));
    CU(cudaStreamSynchronize(ctx->stream));  // h_beta may be a stack temporary of the caller
    return B200ZK_OK;
}

int b200zk_fri_commit_phase(b200zk_ctx* ctx, const uint32_t* const* d_inputs, const uint64_t* lens, uint32_t n_inputs, uint32_t log_blowup,
                            uint32_t log_final_poly_len, b200zk_chal* chal, const uint32_t* h_betas_forced, uint32_t* h_roots, uint32_t* h_betas,
                            uint32_t* h_final, b200zk_tree** trees, uint32_t* h_rounds) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_inputs || !lens || !n_inputs || !h_roots || !h_final || !h_rounds) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (!chal && !h_betas_forced) return fail(ctx, B200ZK_ERR_ARG, "need a challenger or forced betas");
    for (uint32_t j = 0; j < n_inputs; j++) {
        if (!d_inputs[j] || !is_pow2(lens[j])) return fail(ctx, B200ZK_ERR_SHAPE, "input lengths must be powers of two");
        if (j && lens[j] >= lens[j - 1]) return fail(ctx, B200ZK_ERR_SHAPE, "input lengths must be strictly decreasing");
    }
    const uint64_t stop = 1ull << (log_blowup + log_final_poly_len);
    const uint64_t len0 = lens[0];
    if (len0 < stop || log2u(len0) > MAX_LOG + 1) return fail(ctx, B200ZK_ERR_SHAPE, "bad first input length");
    CU(cudaSetDevice(ctx->device));
    uint32_t max_rounds = 0;
    for (uint64_t l = len0; l > stop; l >>= 1) max_rounds++;
    *h_rounds = max_rounds;
    // scratch: per-round power tables + betas
    size_t tab_per = 4096 + hi_words(len0 / 2);
    TRY(ensure_tab(ctx, tab_per * std::max<uint32_t>(max_rounds, 1) + 12 * (size_t)max_rounds + 32));
    uint32_t* d_betas = ctx->tab + tab_per * std::max<uint32_t>(max_rounds, 1);
    uint32_t* d_roots = d_betas + 4 * (size_t)max_rounds + 8;  // roots are gathered here: one D2H at the end, no per-round sync
    if (h_betas_forced && max_rounds) CU(cudaMemcpyAsync(d_betas, h_betas_forced, 16ull * max_rounds, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<b200zk_tree*> made;
    const uint32_t* cur = d_inputs[0];
    uint32_t* owned_cur = nullptr;  // folded vectors we allocated (round >= 1 leaves live in their tree's matrix)
    uint64_t len = len0;
    uint32_t next_in = 1;
    int rc = B200ZK_OK;
    for (uint32_t r = 0; r < max_rounds && rc == B200ZK_OK; r++) {
        // commit the (len/2) x 2 EF4 matrix = (len/2) x 8 base matrix
        b200zk_mat* leaves = nullptr;
        rc = b200zk_mat_wrap(ctx, const_cast<uint32_t*>(cur), len / 2, 8, &leaves);
        if (rc) break;
        if (owned_cur) leaves->owned = true;  // the tree frees the folded vector with its leaves
        b200zk_tree* t = nullptr;
        b200zk_mat* arr[1] = {leaves};
        rc = commit_async(ctx, arr, 1, /*take=*/1, &t);
        if (rc) {
            leaves->owned = false;
            b200zk_mat_free(ctx, leaves);
            break;
        }
        owned_cur = nullptr;
        made.push_back(t);
        const uint32_t* d_root = t->d_digests + 8 * (2 * t->max_h - 2);
        cudaError_t e = cudaMemcpyAsync(d_roots + 8 * r, d_root, 32, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { rc = fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e)); break; }
        if (!h_betas_forced) {
            fri::chal_fri_round_kernel<<<1, 32, 0, ctx->stream>>>(chal->d, d_root, d_betas + 4 * r);
            ctx->launches++;
        }
        uint32_t* nxt = nullptr;
        rc = dev_alloc(ctx, (len / 2) * 16, (void**)&nxt);
        if (rc) break;
        const uint32_t* add = nullptr;
        if (next_in < n_inputs && lens[next_in] == len / 2) add = d_inputs[next_in++];
        rc = fold_launch(ctx, cur, len, d_betas + 4 * r, add, 1, ctx->tab + tab_per * r, nxt);
        if (rc) { dev_free(ctx, nxt); break; }
        cur = nxt;
        owned_cur = nxt;
        len >>= 1;
    }
    if (rc == B200ZK_OK) {
        cudaError_t e = cudaMemcpyAsync(h_final, cur, len * 16, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && max_rounds) e = cudaMemcpyAsync(h_roots, d_roots, 32ull * max_rounds, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && h_betas && max_rounds) e = cudaMemcpyAsync(h_betas, d_betas, 16ull * max_rounds, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    }
    if (owned_cur) dev_free(ctx, owned_cur);
    if (rc == B200ZK_OK && trees) {
        for (uint32_t r = 0; r < max_rounds; r++) trees[r] = made[r];
    } else {
        for (auto* t : made) b200zk_tree_free(ctx, t);
    }
    return rc;
}


// ================================================================================================ PCS open phase
int b200zk_open_denominators(b200zk_ctx* ctx, uint32_t log_m, uint32_t shift, const uint32_t h_point[4], uint32_t* d_inv_den) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_point || !d_inv_den) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (log_m > (uint32_t)MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "size exceeds the two-adicity of BabyBear (2^27)");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    CU(cudaSetDevice(ctx->device));
    TRY(ensure_roots(ctx, (int)log_m));
    uint32_t* d_z = ctx->d_small + 1024;
    CU(cudaMemcpyAsync(d_z, h_point, 16, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t threads = ((1ull << log_m) + op::INV_BATCH - 1) / op::INV_BATCH;
    op::denominators_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, ctx->stream>>>((int)log_m, shift, d_z, ctx->tw_lo[log_m], ctx->tw_hi[log_m], d_inv_den);
    LAUNCHED();
    return B200ZK_OK;
}

int b200zk_mat_dot_ext_powers(b200zk_ctx* ctx, const b200zk_mat* m, const uint32_t h_alpha[4], uint32_t* d_out) {
    TRY(check_mat(ctx, m));
    if (!h_alpha || !d_out) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    uint32_t* d_alpha = ctx->d_small + 1040;
    CU(cudaMemcpyAsync(d_alpha, h_alpha, 16, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* d_pw = nullptr;
    TRY(dev_alloc(ctx, (size_t)m->width * 16, (void**)&d_pw));
    op::ext_powers_kernel<<<(m->width + 127) / 128, 128, 0, ctx->stream>>>(d_alpha, m->width, d_pw);
    LAUNCHED();
    const bool vec4 = m->width % 4 == 0 && ((uintptr_t)m->d % 16) == 0;
    const size_t smem = (size_t)((m->width + 3) / 4 * 4) * 16;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((m->rows + 8 * op::DEP_ROWS - 1) / (8 * op::DEP_ROWS), (uint64_t)ctx->num_sms * 8);
    if (smem > (size_t)ctx->max_smem_optin) return fail(ctx, B200ZK_ERR_SHAPE, "matrix too wide for dot_ext_powers");
    const uint32_t lanes_needed = vec4 ? m->width / 4 : m->width;   // lanes a row keeps busy in the warp-per-row kernel
    if (lanes_needed <= 24) {  // narrow: one thread per row
        const uint32_t nblk = (uint32_t)((m->rows + 255) / 256);
        if (vec4) op::dot_ext_powers_narrow_kernel<4><<<nblk, 256, smem, ctx->stream>>>(m->d, m->rows, m->width, d_pw, d_out);
        else op::dot_ext_powers_narrow_kernel<1><<<nblk, 256, smem, ctx->stream>>>(m->d, m->rows, m->width, d_pw, d_out);
    } else if (vec4) op::dot_ext_powers_kernel<4><<<grid, 256, smem, ctx->stream>>>(m->d, m->rows, m->width, d_pw, d_out);
    else op::dot_ext_powers_kernel<1><<<grid, 256, smem, ctx->stream>>>(m->d, m->rows, m->width, d_pw, d_out);
    LAUNCHED();
    dev_free(ctx, d_pw);
    return B200ZK_OK;
}

// opened values of every column into DEVICE memory (d_ys: width x 4); everything is enqueued on the ctx stream
static int interpolate_coset_dev(b200zk_ctx* ctx, const b200zk_mat* lde, uint32_t log_blowup, uint32_t shift, const uint32_t h_point[4], const uint32_t* d_inv_den,
                                 uint32_t* d_ys) {
    TRY(check_mat(ctx, lde));
    if (!h_point || !d_inv_den || !d_ys) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (!is_pow2(lde->rows) || (lde->rows >> log_blowup) == 0) return fail(ctx, B200ZK_ERR_SHAPE, "bad LDE height / blow-up");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    CU(cudaSetDevice(ctx->device));
    const int lm = log2u(lde->rows);
    const uint64_t n = lde->rows >> log_blowup;
    const uint32_t W = lde->width;
    TRY(ensure_roots(ctx, lm));
    // scale = ((z / shift)^n - 1) / n  in EF4, on the host (a few dozen field operations)
    bb::ef4 z{{h_point[0], h_point[1], h_point[2], h_point[3]}};
    bb::ef4 zn = z;
    for (uint64_t k = 1; k < n; k <<= 1) zn = bb::ef_mul(zn, zn);
    const uint32_t sn_inv = bb::inv(bb::pow(shift, n));
    bb::ef4 sc = bb::ef_scale(zn, sn_inv);
    sc.c[0] = bb::sub(sc.c[0], bb::ONE);
    sc = bb::ef_scale(sc, bb::inv(bb::to_monty((uint32_t)(n % bb::P))));
    uint32_t* d_scale = ctx->d_small + 1056;
    CU(cudaMemcpyAsync(d_scale, sc.c, 16, cudaMemcpyHostToDevice, ctx->stream));
    const bool vec4 = W % 4 == 0 && ((uintptr_t)lde->d % 16) == 0;
    const uint32_t vecw = vec4 ? 4 : 1;
    uint32_t tx_n = 1;  // threads along the columns: just enough to cover the width (narrow traces leave the rest of the CTA to the rows)
    while (tx_n < 256 && tx_n * vecw < W) tx_n <<= 1;
    const uint32_t col_blocks = (W + tx_n * vecw - 1) / (tx_n * vecw);
    uint32_t row_blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, ((uint64_t)ctx->num_sms * 4 + col_blocks - 1) / col_blocks));
    const uint32_t rows_per_cta = (uint32_t)(((n + row_blocks - 1) / row_blocks + 63) / 64 * 64);
    row_blocks = (uint32_t)((n + rows_per_cta - 1) / rows_per_cta);
    uint32_t* d_partial = nullptr;
    TRY(dev_alloc(ctx, (size_t)row_blocks * W * 16, (void**)&d_partial));
    dim3 grid(row_blocks, col_blocks);
    const size_t rsm = (size_t)256 * vecw * 16;  // [ty][tx * VEC][4] words
    if (vec4) op::colwise_bary_kernel<4><<<grid, 256, rsm, ctx->stream>>>(lde->d, n, W, lm, shift, ctx->tw_lo[lm], ctx->tw_hi[lm], d_inv_den, rows_per_cta, tx_n, d_partial);
    else op::colwise_bary_kernel<1><<<grid, 256, rsm, ctx->stream>>>(lde->d, n, W, lm, shift, ctx->tw_lo[lm], ctx->tw_hi[lm], d_inv_den, rows_per_cta, tx_n, d_partial);
    LAUNCHED();
    op::bary_finish_kernel<<<(W + 127) / 128, 128, 0, ctx->stream>>>(d_partial, row_blocks, W, d_scale, d_ys);
    LAUNCHED();
    dev_free(ctx, d_partial);
    return B200ZK_OK;
}


int b200zk_interpolate_coset(b200zk_ctx* ctx, const b200zk_mat* lde, uint32_t log_blowup, uint32_t shift, const uint32_t h_point[4], const uint32_t* d_inv_den,
                             uint32_t* h_ys) {
    TRY(check_mat(ctx, lde));
    if (!h_ys) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    uint32_t* d_ys = nullptr;
    TRY(dev_alloc(ctx, (size_t)lde->width * 16, (void**)&d_ys));
    int rc = interpolate_coset_dev(ctx, lde, log_blowup, shift, h_point, d_inv_den, d_ys);
    if (rc == B200ZK_OK) {
        cudaError_t e = cudaMemcpyAsync(h_ys, d_ys, (size_t)lde->width * 16, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, B200ZK_ERR_CUDA, cudaGetErrorString(e));
    }
    dev_free(ctx, d_ys);
    return rc;
}

int b200zk_ext_powers(b200zk_ctx* ctx, const uint32_t h_alpha[4], uint32_t n, uint32_t* d_out) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_alpha || (!d_out && n)) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (!n) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t* d_alpha = nullptr;
    TRY(dev_alloc(ctx, 16, (void**)&d_alpha));
    CU(cudaMemcpyAsync(d_alpha, h_alpha, 16, cudaMemcpyHostToDevice, ctx->stream));
    op::ext_powers_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_alpha, n, d_out);
    LAUNCHED();
    dev_free(ctx, d_alpha);
    return B200ZK_OK;
}

int b200zk_open_reduce(b200zk_ctx* ctx, const b200zk_mat* lde, uint32_t log_blowup, uint32_t shift, const uint32_t h_point[4], const uint32_t* d_inv_den,
                       const uint32_t* d_reduced_row, const uint32_t* d_alpha_pows, uint32_t alpha_offset, uint32_t* d_ro, uint32_t* d_ys) {
    TRY(check_mat(ctx, lde));
    if (!d_reduced_row || !d_alpha_pows || !d_ro) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    TRY(interpolate_coset_dev(ctx, lde, log_blowup, shift, h_point, d_inv_den, d_ys));
    uint32_t* d_c = nullptr;
    TRY(dev_alloc(ctx, 32, (void**)&d_c));
    op::ef_dot_kernel<<<1, 256, 0, ctx->stream>>>(d_alpha_pows, d_ys, lde->width, alpha_offset, d_c);
    LAUNCHED();
    op::reduce_openings_kernel<<<(uint32_t)((lde->rows + 255) / 256), 256, 0, ctx->stream>>>(d_reduced_row, lde->rows, d_inv_den, d_c, d_c + 4, d_ro);
    LAUNCHED();
    dev_free(ctx, d_c);
    return B200ZK_OK;
}

int b200zk_reduce_openings(b200zk_ctx* ctx, const uint32_t* d_rr, uint64_t m, const uint32_t* d_inv_den, const uint32_t h_rys[4], const uint32_t h_apo[4],
                           uint32_t* d_ro) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_rr || !d_inv_den || !h_rys || !h_apo || !d_ro) return fail(ctx, B200ZK_ERR_ARG, "null argument");
    if (!m) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    uint32_t* d_c = nullptr;  // the two EF4 constants must outlive this call on the stream: take them from the pool
    TRY(dev_alloc(ctx, 32, (void**)&d_c));
    uint32_t tmp[8];
    memcpy(tmp, h_rys, 16);
    memcpy(tmp + 4, h_apo, 16);
    CU(cudaMemcpyAsync(d_c, tmp, 32, cudaMemcpyHostToDevice, ctx->stream));
    op::reduce_openings_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(d_rr, m, d_inv_den, d_c, d_c + 4, d_ro);
    LAUNCHED();
    dev_free(ctx, d_c);
    return B200ZK_OK;
}

// ================================================================================================ peer memory (multi-GPU)
// One process per GPU: a buffer other ranks store into is a plain cudaMalloc allocation exported with a CUDA IPC handle.
int b200zk_peer_alloc(b200zk_ctx* ctx, uint64_t bytes, void** d_out, uint8_t h_handle[64]) {
    if (!ctx || !d_out || !h_handle) return B200ZK_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    *d_out = nullptr;
    cudaError_t e = cudaMalloc(d_out, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cache_flush(ctx);
        cudaStreamSynchronize(ctx->stream);
        e = cudaMalloc(d_out, bytes ? bytes : 16);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, B200ZK_ERR_OOM, std::string("cudaMalloc (peer buffer): ") + cudaGetErrorString(e));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, *d_out));
    memcpy(h_handle, &h, 64);
    return B200ZK_OK;
}
int b200zk_peer_open(b200zk_ctx* ctx, const uint8_t h_handle[64], void** d_out) {
    if (!ctx || !d_out || !h_handle) return B200ZK_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, 64);
    CU(cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
    return B200ZK_OK;
}
int b200zk_peer_close(b200zk_ctx* ctx, void* d_peer) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_peer) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaIpcCloseMemHandle(d_peer));
    return B200ZK_OK;
}
int b200zk_peer_free(b200zk_ctx* ctx, void* d) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d) return B200ZK_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(d));
    return B200ZK_OK;
}

// Column-sharded coset LDE whose result leaves in ROW blocks: this rank extends its `evals` (N x wg columns of the trace) and the last
// NTT pass of every coset stores each finished tile with TMA directly into the memory of the rank that owns those rows
// (d_recv[r], local or a peer mapping), at slot `rank` of that rank's [world][M / world][wg] receive buffer.  The exchange
// rides on the stores of the pass: no separate all-to-all, no staging copy.  Callers synchronise (stream sync + a
// barrier across the ranks) before reading their receive buffer.
static int lde_scatter_impl(b200zk_ctx* ctx, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift, uint32_t world, uint32_t rank, uint32_t* const* d_recv,
                            bool interleaved) {
    TRY(check_mat(ctx, evals));
    if (!d_recv || !world || rank >= world || (world & (world - 1))) return fail(ctx, B200ZK_ERR_ARG, "bad world / rank / receive buffers");
    const uint64_t N = evals->rows;
    const uint32_t W = evals->width;
    if (!is_pow2(N) || N < 2) return fail(ctx, B200ZK_ERR_SHAPE, "height must be a power of two >= 2");
    const int n = log2u(N);
    if (n + (int)added_bits > MAX_LOG) return fail(ctx, B200ZK_ERR_SHAPE, "size exceeds the two-adicity of BabyBear (2^27)");
    if (shift == 0 || shift >= bb::P) return fail(ctx, B200ZK_ERR_ARG, "shift must be a non-zero field element");
    const uint64_t M = N << added_bits;
    if (M % world) return fail(ctx, B200ZK_ERR_SHAPE, "LDE height must be divisible by the number of ranks");
    if (W % 4) return fail(ctx, B200ZK_ERR_SHAPE, "sharded LDE needs width % 4 == 0");
    for (uint32_t r = 0; r < world; r++)
        if (!d_recv[r] || ((uintptr_t)d_recv[r] % 16)) return fail(ctx, B200ZK_ERR_ARG, "null / misaligned receive buffer");
    CU(cudaSetDevice(ctx->device));
    b200zk_mat* work = nullptr;  // local work space: every pass but the last of each coset runs here, exactly as in the unsharded LDE
    TRY(b200zk_mat_alloc(ctx, M, W, &work));
    int rc = lde_tables(ctx, n, added_bits, shift);
    Scatter sc;
    sc.dst = d_recv;
    sc.mg = M / world;
    sc.rank = rank;
    if (interleaved) {  // the owner's buffer is ONE [mg][world * W] matrix, this rank's columns start at rank * W
        sc.slot_off = (uint64_t)rank * W;
        sc.pitch = world * W;
    } else {            // one [mg][W] slot per sender
        sc.slot_off = (uint64_t)rank * sc.mg * W;
    }
    if (rc == B200ZK_OK) rc = lde_core(ctx, evals->d, W, n, W, added_bits, work->d, W, &sc);
    b200zk_mat_free(ctx, work);
    return rc;
}
int b200zk_coset_lde_scatter(b200zk_ctx* ctx, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift, uint32_t world, uint32_t rank,
                             uint32_t* const* d_recv) {
    return lde_scatter_impl(ctx, evals, added_bits, shift, world, rank, d_recv, false);
}
int b200zk_coset_lde_scatter_rows(b200zk_ctx* ctx, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift, uint32_t world, uint32_t rank,
                                  uint32_t* const* d_recv) {
    return lde_scatter_impl(ctx, evals, added_bits, shift, world, rank, d_recv, true);
}

// ================================================================================================ host staging memory
// Page-locked host memory for traces that will be committed with b200zk_lde_commit_host(_async).  write_combined = 1 asks for
// write-combining pages: the CPU fills them with streaming stores (reading them back on the CPU is slow) and the device's DMA reads do
// not snoop the CPU caches, which matters when several GPUs pull from host memory at once.
int b200zk_host_alloc(uint64_t bytes, int write_combined, void** h_out) {
    if (!h_out) return B200ZK_ERR_ARG;
    *h_out = nullptr;
    cudaError_t e = cudaHostAlloc(h_out, bytes ? bytes : 16, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *h_out = nullptr;
        return e == cudaErrorMemoryAllocation ? B200ZK_ERR_OOM : B200ZK_ERR_CUDA;
    }
    return B200ZK_OK;
}
void b200zk_host_free(void* h) {
    if (h) cudaFreeHost(h);
}

// ================================================================================================ raw memory
int b200zk_dev_alloc(b200zk_ctx* ctx, uint64_t bytes, void** d_out) {
    if (!ctx || !d_out) return B200ZK_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    return dev_alloc(ctx, bytes, d_out);
}
void b200zk_dev_free(b200zk_ctx* ctx, void* d) { dev_free(ctx, d); }
int b200zk_dev_zero(b200zk_ctx* ctx, void* d, uint64_t bytes) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d && bytes) return fail(ctx, B200ZK_ERR_ARG, "null pointer");
    CU(cudaSetDevice(ctx->device));
    if (bytes) CU(cudaMemsetAsync(d, 0, bytes, ctx->stream));
    return B200ZK_OK;
}
int b200zk_dev_upload(b200zk_ctx* ctx, void* d_dst, const void* h_src, uint64_t bytes) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!d_dst || !h_src) return fail(ctx, B200ZK_ERR_ARG, "null pointer");
    CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}
int b200zk_dev_download(b200zk_ctx* ctx, void* h_dst, const void* d_src, uint64_t bytes) {
    if (!ctx) return B200ZK_ERR_ARG;
    if (!h_dst || !d_src) return fail(ctx, B200ZK_ERR_ARG, "null pointer");
    CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

}  // extern "C"
