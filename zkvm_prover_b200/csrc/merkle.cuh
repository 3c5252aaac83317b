// K4/K5: PaddingFreeSponge leaf hashing and TruncatedPermutation tree compression for
// MerkleTreeMmcs<.., 8> (p3-merkle-tree MerkleTree::new: first_digest_layer + compress_and_inject; v1-era
// crate that produced the reference's proof fixtures, see tests/golden).  One permutation chain per
// thread; the kernels are integer-pipe bound, loads are issued one permutation ahead.
#pragma once
#include "poseidon2.cuh"

namespace mk {

constexpr int MAX_GROUP = 128;  // matrices of one height class hashed into one sponge per launch

struct MatRef {
    const uint32_t* ptr;
    uint32_t width;
};
struct Group {
    MatRef m[MAX_GROUP];
    int n;
    int fast8;  // every matrix: width % 8 == 0 and 32-byte aligned base  -> vector path
    const MatRef* ext;  // n > MAX_GROUP (p3 has no limit on matrices per height): the descriptors live in device memory instead
};
__device__ __forceinline__ MatRef mref(const Group& g, int i) { return g.ext ? g.ext[i] : g.m[i]; }

// row data is read through L1 (allocating): a thread consumes its 128-byte line in two 64-byte steps ~100 us apart, and
// the per-SM working set (resident threads x 128 B <= 160 KB) fits the L1, so the second half never goes back to L2/HBM
// (with L1::no_allocate the kernel moved 26.6 GB for 17.7 GB of data)
#ifndef MK_LDG_NOALLOC
#define MK_LDG_NOALLOC 0
#endif
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
#if MK_LDG_NOALLOC
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}

struct Cursor {
    int m;
    uint32_t c, w;
    const uint32_t* p;
};
__device__ __forceinline__ void cursor_next_matrix(const Group& g, uint64_t row, Cursor& cur) {
    cur.m++;
    if (cur.m < g.n) {
        const MatRef r = mref(g, cur.m);
        cur.w = r.width;
        cur.p = r.ptr + row * cur.w;
        cur.c = 0;
    }
}
// next (up to) 8 elements of the concatenated row; returns how many (0 = end of row)
__device__ __forceinline__ int gather8(const Group& g, uint64_t row, Cursor& cur, uint32_t (&buf)[8]) {
    if (cur.m >= g.n) return 0;
    if (cur.w - cur.c >= 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) buf[i] = __ldg(cur.p + cur.c + i);
        cur.c += 8;
        if (cur.c == cur.w) cursor_next_matrix(g, row, cur);
        return 8;
    }
    int cnt = 0;
    while (cnt < 8 && cur.m < g.n) {
        if (cur.c < cur.w) {
            const uint32_t x = __ldg(cur.p + cur.c);
            cur.c++;
#pragma unroll
            for (int i = 0; i < 8; i++) buf[i] = (cnt == i) ? x : buf[i];
            cnt++;
        } else {
            cursor_next_matrix(g, row, cur);
        }
    }
    return cnt;
}

// sponge over the concatenation of row `row` of every matrix of the group -> st[0..8]
__device__ __forceinline__ void sponge_rows(const Group& g, uint64_t row, uint32_t (&st)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = 0;
    if (g.fast8) {
        // 64 B (one HBM access atom) per load step: consuming a row 32 B at a time left every atom half used and
        // refetched later (ncu r01: 47.7 GB of DRAM traffic for 17.7 GB of data).  Two permutations run per step
        // while the next 64 B are in flight.
        for (int m = 0; m < g.n; m++) {
            const MatRef r = mref(g, m);
            const uint32_t w = r.width;
            const uint4* p = reinterpret_cast<const uint4*>(r.ptr + row * w);
            const uint32_t chunks = w >> 3;  // 32-byte chunks in this row
            uint4 a = ldg_stream(p), b = ldg_stream(p + 1), c = a, d = b;
            if (chunks > 1) { c = ldg_stream(p + 2); d = ldg_stream(p + 3); }
            for (uint32_t k = 0; k < chunks; k += 2) {
                st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
                st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
                const uint4 c2 = c, d2 = d;
                if (k + 2 < chunks) {  // prefetch the next 64 bytes under the two permutations
                    a = ldg_stream(p + 2 * (k + 2));
                    b = ldg_stream(p + 2 * (k + 2) + 1);
                    if (k + 3 < chunks) { c = ldg_stream(p + 2 * (k + 3)); d = ldg_stream(p + 2 * (k + 3) + 1); }
                }
                p2::permute(st);
                if (k + 1 < chunks) {
                    st[0] = c2.x; st[1] = c2.y; st[2] = c2.z; st[3] = c2.w;
                    st[4] = d2.x; st[5] = d2.y; st[6] = d2.z; st[7] = d2.w;
                    p2::permute(st);
                }
            }
        }
        return;
    }
    // general path (ragged widths, several matrices in one sponge -- the shape of real traces): a cursor walks the
    // concatenated row; 8 elements are gathered with independent loads (selects only where a chunk straddles two
    // matrices).
    Cursor cur;
    cur.m = 0;
    cur.c = 0;
    const MatRef r0 = mref(g, 0);
    cur.w = r0.width;
    cur.p = r0.ptr + row * cur.w;
    uint32_t buf[8];
    int cnt = gather8(g, row, cur, buf);
    while (cnt > 0) {
        if (cnt == 8) {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = buf[i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = (i < cnt) ? buf[i] : st[i];
        }
        p2::permute(st);
        cnt = gather8(g, row, cur, buf);   // loaded at the point of use (prefetching before the permutation measured the same)
    }
}

__device__ __forceinline__ void store_digest(uint32_t* out, const uint32_t (&st)[16]) {
    uint4* o = reinterpret_cast<uint4*>(out);
    o[0] = make_uint4(st[0], st[1], st[2], st[3]);
    o[1] = make_uint4(st[4], st[5], st[6], st[7]);
}

// first_digest_layer: digests[i] = hash_iter(concat rows i of the tallest matrices)
__global__ void __launch_bounds__(256) leaf_hash_kernel(const __grid_constant__ Group g, uint64_t rows, uint32_t* __restrict__ digests) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    uint32_t st[16];
    sponge_rows(g, i, st);
    store_digest(digests + 8 * i, st);
}

// The same for groups whose matrices all have width % 8 == 0 (the LDE commit): a kernel of its own so that the
// permutation gets the registers.  A row is read 64 B (one HBM access atom) at a time, at the point of use: a warp spends
// ~70k clocks per permutation, so the ~1k clocks of an exposed load cost ~1 % and other warps cover them, while the
// prefetch registers of the general kernel cost the permutation its scheduling freedom (3.97 -> 4.39 Gperm/s at 2^24 x 256).
#ifndef MK_FAST_MINB
#define MK_FAST_MINB 3
#endif
__global__ void __launch_bounds__(256, MK_FAST_MINB) leaf_hash_fast_kernel(const __grid_constant__ Group g, uint64_t rows, uint32_t* __restrict__ digests) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    uint32_t st[16];
#pragma unroll
    for (int k = 0; k < 16; k++) st[k] = 0;
    for (int m = 0; m < g.n; m++) {
        const MatRef r = mref(g, m);
        const uint32_t w = r.width;
        const uint4* p = reinterpret_cast<const uint4*>(r.ptr + i * w);
        const uint32_t chunks = w >> 3;
#pragma unroll 1
        for (uint32_t k = 0; k + 1 < chunks; k += 2) {
            const uint4 a = ldg_stream(p + 2 * k), b = ldg_stream(p + 2 * k + 1), c = ldg_stream(p + 2 * k + 2), d = ldg_stream(p + 2 * k + 3);
            st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
            st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
            p2::permute(st);
            st[0] = c.x; st[1] = c.y; st[2] = c.z; st[3] = c.w;
            st[4] = d.x; st[5] = d.y; st[6] = d.z; st[7] = d.w;
            p2::permute(st);
        }
        if (chunks & 1) {
            const uint4 a = ldg_stream(p + 2 * (chunks - 1)), b = ldg_stream(p + 2 * (chunks - 1) + 1);
            st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
            st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
            p2::permute(st);
        }
    }
    store_digest(digests + 8 * i, st);
}

// incremental leaf hashing for the column-strip pipeline (b200zk_lde_commit_host): absorbs columns [col0, col0+cols) of
// every row, cols % 8 == 0.  Between strips only the capacity half of the sponge (st[8..16)) has to be carried: the
// rate half is overwritten by the next chunk.  `cap` is rows x 8; the last strip writes the digest.
__global__ void __launch_bounds__(256, MK_FAST_MINB) leaf_absorb_strip_kernel(const uint32_t* __restrict__ mat, uint32_t pitch, uint32_t col0, uint32_t cols, uint64_t rows,
                                                                uint32_t* __restrict__ cap, int first, int last, uint32_t* __restrict__ digests) {
    // grid-stride over the rows: with the usual one-thread-per-row grid the loop runs once; a small persistent grid (a couple of
    // CTAs per SM) leaves registers and shared memory of every SM free for the NTT kernels of the next strip (B200ZK_HASH_STREAM=2)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t st[16];
#pragma unroll
    for (int k = 0; k < 16; k++) st[k] = 0;
    if (!first) {
        const uint4* cp = reinterpret_cast<const uint4*>(cap + 8 * i);
        uint4 a = cp[0], b = cp[1];
        st[8] = a.x; st[9] = a.y; st[10] = a.z; st[11] = a.w; st[12] = b.x; st[13] = b.y; st[14] = b.z; st[15] = b.w;
    }
    const uint4* p = reinterpret_cast<const uint4*>(mat + i * pitch + col0);
    const uint32_t chunks = cols >> 3;
#pragma unroll 1
    for (uint32_t k = 0; k + 1 < chunks; k += 2) {  // 64 B per step, loaded at the point of use (see leaf_hash_fast_kernel)
        const uint4 a = ldg_stream(p + 2 * k), b = ldg_stream(p + 2 * k + 1), c = ldg_stream(p + 2 * k + 2), d = ldg_stream(p + 2 * k + 3);
        st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
        st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
        p2::permute(st);
        st[0] = c.x; st[1] = c.y; st[2] = c.z; st[3] = c.w;
        st[4] = d.x; st[5] = d.y; st[6] = d.z; st[7] = d.w;
        p2::permute(st);
    }
    if (chunks & 1) {
        const uint4 a = ldg_stream(p + 2 * (chunks - 1)), b = ldg_stream(p + 2 * (chunks - 1) + 1);
        st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
        st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
        p2::permute(st);
    }
    if (last) {
        store_digest(digests + 8 * i, st);
    } else {
        uint4* cp = reinterpret_cast<uint4*>(cap + 8 * i);
        cp[0] = make_uint4(st[8], st[9], st[10], st[11]);
        cp[1] = make_uint4(st[12], st[13], st[14], st[15]);
    }
    }
}

__device__ __forceinline__ void compress_node(const uint32_t* __restrict__ prev, uint64_t i, uint32_t (&st)[16]) {
    const uint4* p = reinterpret_cast<const uint4*>(prev + 16 * i);
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w; st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
    st[8] = c.x; st[9] = c.y; st[10] = c.z; st[11] = c.w; st[12] = d.x; st[13] = d.y; st[14] = d.z; st[15] = d.w;
    p2::permute(st);
}

// compress_and_inject: next[i] = compress(prev[2i], prev[2i+1]); if matrices of this height exist:
// next[i] = compress(next[i], hash_iter(rows i of those))
__global__ void __launch_bounds__(256) compress_layer_kernel(const uint32_t* __restrict__ prev, uint32_t* __restrict__ next, uint64_t n_next,
                                                             const __grid_constant__ Group g) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_next) return;
    uint32_t st[16];
    compress_node(prev, i, st);
    if (g.n > 0) {
        uint32_t h[16];
        sponge_rows(g, i, h);
#pragma unroll
        for (int k = 0; k < 8; k++) st[8 + k] = h[k];
        p2::permute(st);
    }
    store_digest(next + 8 * i, st);
}

// the top of the tree (no injected matrices): layers n0 -> n0/2 -> ... -> 1 in one CTA
__global__ void __launch_bounds__(1024) compress_top_kernel(uint32_t* __restrict__ layers /* layer of n0 digests, followed by the next ones */, uint32_t n0) {
    uint32_t* prev = layers;
    for (uint32_t n = n0 >> 1; n >= 1; n >>= 1) {
        uint32_t* next = prev + 16ull * n;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t st[16];
            compress_node(prev, i, st);
            store_digest(next + 8 * i, st);
        }
        __syncthreads();
        prev = next;
        if (n == 1) break;
    }
}

// standalone K3/K5 entry points
__global__ void __launch_bounds__(256) permute_kernel(uint32_t* __restrict__ states, uint64_t n, int plain) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4* p = reinterpret_cast<uint4*>(states + 16 * i);
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    uint32_t st[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    if (plain) p2::permute_plain(st); else p2::permute(st);
    p[0] = make_uint4(st[0], st[1], st[2], st[3]);
    p[1] = make_uint4(st[4], st[5], st[6], st[7]);
    p[2] = make_uint4(st[8], st[9], st[10], st[11]);
    p[3] = make_uint4(st[12], st[13], st[14], st[15]);
}

// Mmcs::open_batch gather: rows of every matrix at index >> shift, and the sibling path
struct OpenMat {
    const uint32_t* ptr;
    uint32_t width;
    uint32_t shift;      // log2(max_height) - log2(height)
    uint64_t out_off;    // offset in rows_out
};
__global__ void open_rows_kernel(const OpenMat* __restrict__ mats, uint32_t nmats, uint64_t index, uint32_t* __restrict__ rows_out) {
    for (uint32_t m = blockIdx.x; m < nmats; m += gridDim.x) {
        OpenMat om = mats[m];
        const uint32_t* src = om.ptr + (index >> om.shift) * om.width;
        for (uint32_t c = threadIdx.x; c < om.width; c += blockDim.x) rows_out[om.out_off + c] = src[c];
    }
}
__global__ void open_path_kernel(const uint32_t* __restrict__ digests, uint64_t max_h, uint32_t depth, uint64_t index, uint32_t* __restrict__ path_out) {
    // digests: layer 0 (max_h) | layer 1 (max_h/2) | ...
    uint32_t d = blockIdx.x;
    if (d >= depth) return;
    uint64_t off = 0, n = max_h;
    for (uint32_t k = 0; k < d; k++) { off += n; n >>= 1; }
    uint64_t sib = (index >> d) ^ 1;
    if (threadIdx.x < 8) path_out[8 * d + threadIdx.x] = digests[8 * (off + sib) + threadIdx.x];
}

// batched open (query phase): block (q, m) copies row index[q] >> shift of matrix m; path blocks copy the siblings
__global__ void open_many_kernel(const OpenMat* __restrict__ mats, uint32_t nmats, const uint64_t* __restrict__ indices, uint32_t n_idx, uint64_t total_width,
                                 const uint32_t* __restrict__ digests, uint64_t max_h, uint32_t depth, uint32_t* __restrict__ rows_out,
                                 uint32_t* __restrict__ paths_out) {
    const uint32_t q = blockIdx.y;
    if (q >= n_idx) return;
    const uint64_t index = indices[q];
    for (uint32_t m = blockIdx.x; m < nmats; m += gridDim.x) {
        OpenMat om = mats[m];
        const uint32_t* src = om.ptr + (index >> om.shift) * om.width;
        for (uint32_t c = threadIdx.x; c < om.width; c += blockDim.x) rows_out[q * total_width + om.out_off + c] = src[c];
    }
    if (blockIdx.x == 0) {
        for (uint32_t i = threadIdx.x; i < depth * 8; i += blockDim.x) {
            const uint32_t d = i >> 3;
            uint64_t off = 0, n = max_h;
            for (uint32_t k = 0; k < d; k++) { off += n; n >>= 1; }
            paths_out[(uint64_t)q * depth * 8 + i] = digests[8 * (off + ((index >> d) ^ 1)) + (i & 7)];
        }
    }
}

// MerkleTreeMmcs::verify_batch, single thread (tiny; device so the host mirror has no CPU hash)
struct VerifyArgs {
    const uint32_t* rows;       // concatenated opened rows, sorted order (tallest first)
    const uint32_t* widths;     // per matrix, sorted order
    const uint32_t* log_heights;
    uint32_t k;
    const uint32_t* path;
    uint32_t depth;
    uint64_t index;
    const uint32_t* root;
    int* ok;
};
__global__ void verify_kernel(VerifyArgs a) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t pos = 0;
    uint64_t off = 0;
    uint32_t node[16], h[16];
    auto absorb_group = [&](uint32_t lh, uint32_t (&st)[16]) -> bool {
        for (int i = 0; i < 16; i++) st[i] = 0;
        int fill = 0;
        bool any = false;
        while (pos < a.k && a.log_heights[pos] == lh) {
            any = true;
            for (uint32_t c = 0; c < a.widths[pos]; c++) {
                uint32_t x = a.rows[off + c];
#pragma unroll
                for (int i = 0; i < 8; i++) st[i] = (fill == i) ? x : st[i];
                if (++fill == 8) { p2::permute(st); fill = 0; }
            }
            off += a.widths[pos];
            pos++;
        }
        if (fill) p2::permute(st);
        return any;
    };
    uint32_t lh = a.log_heights[0];
    absorb_group(lh, node);
    uint64_t index = a.index;
    for (uint32_t d = 0; d < a.depth; d++) {
        uint32_t st[16];
        const uint32_t* sib = a.path + 8 * d;
        if (index & 1) { for (int i = 0; i < 8; i++) { st[i] = sib[i]; st[8 + i] = node[i]; } }
        else { for (int i = 0; i < 8; i++) { st[i] = node[i]; st[8 + i] = sib[i]; } }
        p2::permute(st);
        for (int i = 0; i < 8; i++) node[i] = st[i];
        index >>= 1;
        lh--;
        if (pos < a.k && a.log_heights[pos] == lh) {
            absorb_group(lh, h);
            for (int i = 0; i < 8; i++) st[i] = node[i], st[8 + i] = h[i];
            p2::permute(st);
            for (int i = 0; i < 8; i++) node[i] = st[i];
        }
    }
    int ok = (pos == a.k) && (lh == 0);
    for (int i = 0; i < 8; i++) ok &= (node[i] == a.root[i]);
    *a.ok = ok;
}

}  // namespace mk
