// PCS open-phase primitives: what p3-fri's TwoAdicFriPcs::open does to the committed LDE
// matrices right after the commit -- the reason Mmcs::get_matrices wants host-readable data in the CPU prover.
//   denominators      inv_den[i] = 1 / (z - x_i),  x_i = shift * w_M^bitrev(i)           (batch_multiplicative_inverse)
//   dot_ext_powers    rr[r] = sum_c alpha^c * M[r][c]                                    (p3_matrix::Matrix::dot_ext_powers)
//   interpolate_coset y_c = p_c(z) from the low coset of the LDE, barycentric            (p3_interpolation::interpolate_coset)
//   reduce_openings   ro[i] += alpha^offset * (sum_c alpha^c y_c - rr[i]) * inv_den[i]   (the loop body of TwoAdicFriPcs::open)
// All values are field-exact, so any correct evaluation order is bit-identical to the CPU path.
// These kernels are the HBM-bound part of the path: one streaming read of the LDE each.
#pragma once
#include "bb31.cuh"

namespace op {

using bb::ef4;

__device__ __forceinline__ ef4 ef_load(const uint32_t* p) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    return ef4{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void ef_store(uint32_t* p, const ef4& a) { *reinterpret_cast<uint4*>(p) = make_uint4(a.c[0], a.c[1], a.c[2], a.c[3]); }

// a^-1 in EF4 = F[x]/(x^4 - 11): with a = (a0 + a2 x^2) + x (a1 + a3 x^2) and y = x^2 (y^2 = 11):
//   a * a' = A^2 - y B^2 =: c0 + c1 y   (a' = A - xB, A = a0 + a2 y, B = a1 + a3 y)   -- an element of F[y]/(y^2 - 11)
//   (c0 + c1 y)^-1 = (c0 - c1 y) / (c0^2 - 11 c1^2)                                   -- one base-field inversion
__device__ __forceinline__ ef4 ef_inv(const ef4& a) {
    using namespace bb;
    const uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    // A^2 = (a0^2 + 11 a2^2) + (2 a0 a2) y ;  B^2 = (a1^2 + 11 a3^2) + (2 a1 a3) y ;  y B^2 = 11 (2 a1 a3) + (a1^2 + 11 a3^2) y
    const uint32_t A0 = add(mul(a0, a0), mul(W11, mul(a2, a2))), A1 = dbl(mul(a0, a2));
    const uint32_t B0 = add(mul(a1, a1), mul(W11, mul(a3, a3))), B1 = dbl(mul(a1, a3));
    const uint32_t c0 = sub(A0, mul(W11, B1)), c1 = sub(A1, B0);
    const uint32_t nrm = sub(mul(c0, c0), mul(W11, mul(c1, c1)));
    const uint32_t ni = inv(nrm);
    const uint32_t d0 = mul(c0, ni), d1 = neg(mul(c1, ni));  // (c0 + c1 y)^-1 = d0 + d1 y
    // a^-1 = a' * (d0 + d1 y),  a' = (a0, -a1, a2, -a3)
    const uint32_t p0 = a0, p1 = neg(a1), p2 = a2, p3 = neg(a3);
    // (p0 + p1 x + p2 x^2 + p3 x^3) * (d0 + d1 x^2), x^4 = 11
    ef4 r;
    r.c[0] = add(mul(p0, d0), mul(W11, mul(p2, d1)));
    r.c[1] = add(mul(p1, d0), mul(W11, mul(p3, d1)));
    r.c[2] = add(mul(p2, d0), mul(p0, d1));
    r.c[3] = add(mul(p3, d0), mul(p1, d1));
    return r;
}

// inv_den[i] = 1 / (z - shift * w^bitrev(i)), i < 2^lm.  Each thread inverts INV_BATCH consecutive entries with one
// EF4 inversion (Montgomery's trick).  Roots come from the two-level tables of the ctx (forward roots of size 2^lm).
constexpr int INV_BATCH = 8;
__global__ void __launch_bounds__(256) denominators_kernel(int lm, uint32_t shift, const uint32_t* __restrict__ z4, const uint32_t* __restrict__ tw_lo,
                                                           const uint32_t* __restrict__ tw_hi, uint32_t* __restrict__ inv_den) {
    const uint64_t m = 1ull << lm;
    const uint64_t t0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * INV_BATCH;
    if (t0 >= m) return;
    const ef4 z = ef_load(z4);
    ef4 den[INV_BATCH], pre[INV_BATCH];
    ef4 acc{{bb::ONE, 0, 0, 0}};
#pragma unroll
    for (int k = 0; k < INV_BATCH; k++) {
        const uint64_t i = t0 + k;
        ef4 d = z;
        if (i < m) {
            const uint32_t e = bb::bitrev((uint32_t)i, lm);
            const uint32_t x = bb::mul(shift, bb::mul(__ldg(tw_lo + (e & 4095)), __ldg(tw_hi + (e >> 12))));
            d.c[0] = bb::sub(d.c[0], x);
        } else {
            d = ef4{{bb::ONE, 0, 0, 0}};
        }
        den[k] = d;
        pre[k] = acc;
        acc = bb::ef_mul(acc, d);
    }
    ef4 inv = ef_inv(acc);
#pragma unroll
    for (int k = INV_BATCH - 1; k >= 0; k--) {
        const uint64_t i = t0 + k;
        const ef4 r = bb::ef_mul(inv, pre[k]);
        inv = bb::ef_mul(inv, den[k]);
        if (i < m) ef_store(inv_den + 4 * i, r);
    }
}

// alpha powers: pw[c] = alpha^c, c < width  (one thread per entry, square-and-multiply; width is small)
__global__ void ext_powers_kernel(const uint32_t* __restrict__ alpha4, uint32_t width, uint32_t* __restrict__ pw) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    ef4 b = ef_load(alpha4), r{{bb::ONE, 0, 0, 0}};
    for (uint32_t e = c; e; e >>= 1) {
        if (e & 1) r = bb::ef_mul(r, b);
        b = bb::ef_mul(b, b);
    }
    ef_store(pw + 4 * c, r);
}

// unsigned Montgomery reduction of t < 2^32 * p (two lazily accumulated products) -> [0, p)
__device__ __forceinline__ uint32_t reduce2(uint64_t t) {
    const uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    const uint32_t q = lo * bb::MU;
    const uint32_t qh = __umulhi(q, bb::P);
    const uint32_t r = hi - qh;
    return bb::umin32(r, r + bb::P);
}

// rr[r] = sum_c alpha^c * M[r][c].  A warp walks one row at a time with lanes on adjacent columns (fully coalesced 512 B
// per request), products accumulate lazily in 64 bits (two products per Montgomery reduction), the 32 partial sums of
// 8 rows are combined by one transposed butterfly (31 shuffles per 8 rows) and leave as one coalesced 128 B store.
constexpr int DEP_ROWS = 8;
template <int VEC>
__global__ void __launch_bounds__(256) dot_ext_powers_kernel(const uint32_t* __restrict__ mat, uint64_t rows, uint32_t width, const uint32_t* __restrict__ pw_g,
                                                             uint32_t* __restrict__ out) {
    // powers in shared memory, transposed to [e][column group] so that the 32 lanes of a request read 32 consecutive
    // 16-byte entries (the natural [column] order made lanes stride by 64 B: a 4-way bank conflict on every read)
    extern __shared__ __align__(16) uint32_t pw[];  // VEC x ceil(width / VEC) x 4
    const uint32_t groups = (width + VEC - 1) / VEC;
    for (uint32_t i = threadIdx.x; i < width; i += blockDim.x) {
        const uint4 t = *reinterpret_cast<const uint4*>(pw_g + 4 * i);
        *reinterpret_cast<uint4*>(pw + 4 * ((i % VEC) * groups + i / VEC)) = t;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r0 = warp * DEP_ROWS; r0 < rows; r0 += nwarps * DEP_ROWS) {
        uint32_t part[DEP_ROWS * 4];
#pragma unroll
        for (int k = 0; k < DEP_ROWS * 4; k++) part[k] = 0;
        // column block outermost: the 8 rows' loads are issued together (8 independent 16-byte requests per lane in
        // flight) and the four power entries are read from shared memory once for all 8 rows
        for (uint32_t c = lane * VEC; c < width; c += 32 * VEC) {
            uint32_t v[DEP_ROWS][VEC];
#pragma unroll
            for (int j = 0; j < DEP_ROWS; j++) {
                const uint64_t r = r0 + j < rows ? r0 + j : rows - 1;  // clamped: the result of a clamped row is discarded
                const uint32_t* row = mat + r * width + c;
                if (VEC == 4) {
                    const uint4 t = *reinterpret_cast<const uint4*>(row);
                    v[j][0] = t.x; v[j][1 % VEC] = t.y; v[j][2 % VEC] = t.z; v[j][3 % VEC] = t.w;
                } else {
                    v[j][0] = row[0];
                }
            }
            uint4 pe[VEC];
#pragma unroll
            for (int e = 0; e < VEC; e++) pe[e] = *reinterpret_cast<const uint4*>(pw + 4 * (e * groups + c / VEC));
#pragma unroll
            for (int j = 0; j < DEP_ROWS; j++) {
                uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                for (int e = 0; e < VEC; e++) {
                    s0 += (uint64_t)pe[e].x * v[j][e];
                    s1 += (uint64_t)pe[e].y * v[j][e];
                    s2 += (uint64_t)pe[e].z * v[j][e];
                    s3 += (uint64_t)pe[e].w * v[j][e];
                    if ((e & 1) || VEC == 1) {  // two products per reduction keep the sum below 2^32 * p
                        part[4 * j] = bb::add(part[4 * j], reduce2(s0)); part[4 * j + 1] = bb::add(part[4 * j + 1], reduce2(s1));
                        part[4 * j + 2] = bb::add(part[4 * j + 2], reduce2(s2)); part[4 * j + 3] = bb::add(part[4 * j + 3], reduce2(s3));
                        s0 = s1 = s2 = s3 = 0;
                    }
                }
            }
        }
        // transposed butterfly: after step s every lane keeps half of its values, summed with its partner's copy
        // 32 values -> 16 -> 8 -> 4 -> 2 -> 1; lane l ends with value index l (row l / 4, coefficient l % 4)
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const int half = 16 >> s;          // values kept after this step
            const int bit = 16 >> s;           // lane bit that decides which half is kept
            const bool upper = lane & bit;
#pragma unroll
            for (int k = 0; k < half; k++) {
                const uint32_t keep = upper ? part[k + half] : part[k];
                const uint32_t send = upper ? part[k] : part[k + half];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, send, bit);
                part[k] = bb::add(keep, got);
            }
        }
        // lane l now holds the total of value index v(l) = sum over steps of (lane bit set ? half : 0) = l with bits mapped 16,8,4,2,1 -> same order
        const uint64_t idx = r0 * 4 + lane;
        if (idx < rows * 4) out[idx] = part[0];
    }
}

// column-wise barycentric sum over the low coset: acc_c = sum_{i < n} (x_i * inv_den[i]) * M[i][c].
// A CTA is TX x TY threads: tx sits on VEC adjacent columns (coalesced rows), ty strides over the rows of the CTA's
// slab, so narrow matrices still fill the CTA; products of two rows accumulate lazily in 64 bits per reduction.
// Partial sums of the TY row lanes are combined in shared memory; a second kernel adds the per-CTA partials.
// The same for narrow matrices (real traces: a dozen columns): one thread per row, so no lane idles beside a short row.
// Rows are adjacent in memory, a warp's loads cover one contiguous span, and every lane reads the same power entry
// (shared-memory broadcast).
template <int VEC>
__global__ void __launch_bounds__(256) dot_ext_powers_narrow_kernel(const uint32_t* __restrict__ mat, uint64_t rows, uint32_t width,
                                                                    const uint32_t* __restrict__ pw_g, uint32_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t pw[];  // width x 4, natural order
    for (uint32_t i = threadIdx.x; i < width; i += blockDim.x) *reinterpret_cast<uint4*>(pw + 4 * i) = *reinterpret_cast<const uint4*>(pw_g + 4 * i);
    __syncthreads();
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint32_t* row = mat + r * width;
    uint32_t acc[4] = {0, 0, 0, 0};
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t pending = 0;
    for (uint32_t c = 0; c < width; c += VEC) {
        uint32_t v[VEC];
        if (VEC == 4) {
            const uint4 t = *reinterpret_cast<const uint4*>(row + c);
            v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
        } else {
            v[0] = row[c];
        }
#pragma unroll
        for (int e = 0; e < VEC; e++) {
            const uint4 pe = *reinterpret_cast<const uint4*>(pw + 4 * (c + e));
            s0 += (uint64_t)pe.x * v[e];
            s1 += (uint64_t)pe.y * v[e];
            s2 += (uint64_t)pe.z * v[e];
            s3 += (uint64_t)pe.w * v[e];
            if (++pending == 2) {  // two products per reduction keep the sum below 2^32 * p
                acc[0] = bb::add(acc[0], reduce2(s0)); acc[1] = bb::add(acc[1], reduce2(s1));
                acc[2] = bb::add(acc[2], reduce2(s2)); acc[3] = bb::add(acc[3], reduce2(s3));
                s0 = s1 = s2 = s3 = 0;
                pending = 0;
            }
        }
    }
    if (pending) {
        acc[0] = bb::add(acc[0], reduce2(s0)); acc[1] = bb::add(acc[1], reduce2(s1));
        acc[2] = bb::add(acc[2], reduce2(s2)); acc[3] = bb::add(acc[3], reduce2(s3));
    }
    *reinterpret_cast<uint4*>(out + 4 * r) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
}

template <int VEC>
__global__ void __launch_bounds__(256) colwise_bary_kernel(const uint32_t* __restrict__ mat, uint64_t n, uint32_t width, int lm, uint32_t shift,
                                                           const uint32_t* __restrict__ tw_lo, const uint32_t* __restrict__ tw_hi,
                                                           const uint32_t* __restrict__ inv_den, uint32_t rows_per_cta, uint32_t tx_n, uint32_t* __restrict__ partial) {
    const uint32_t ty_n = 256 / tx_n;
    const uint32_t tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const uint32_t c0 = (blockIdx.y * tx_n + tx) * VEC;
    const uint64_t r_begin = (uint64_t)blockIdx.x * rows_per_cta;
    const uint64_t r_end = r_begin + rows_per_cta < n ? r_begin + rows_per_cta : n;
    uint32_t acc[VEC][4];
#pragma unroll
    for (int e = 0; e < VEC; e++) acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0;
    __shared__ __align__(16) uint32_t sd[64 * 4];
    extern __shared__ __align__(16) uint32_t red[];  // [ty_n][tx_n * VEC][4] for the final reduction
    for (uint64_t rb = r_begin; rb < r_end; rb += 64) {
        __syncthreads();
        if (threadIdx.x < 64 && rb + threadIdx.x < r_end) {  // d_i = x_i / (z - x_i) for the next 64 rows
            const uint64_t i = rb + threadIdx.x;
            const uint32_t e = bb::bitrev((uint32_t)i, lm);
            const uint32_t x = bb::mul(shift, bb::mul(__ldg(tw_lo + (e & 4095)), __ldg(tw_hi + (e >> 12))));
            const ef4 d = bb::ef_scale(ef_load(inv_den + 4 * i), x);
            ef_store(sd + 4 * threadIdx.x, d);
        }
        __syncthreads();
        if (c0 >= width) continue;
        const uint32_t lim = (uint32_t)(r_end - rb < 64 ? r_end - rb : 64);
        for (uint32_t k = ty; k < lim; k += 2 * ty_n) {  // two rows per reduction
            const uint32_t k2 = k + ty_n;
            const bool has2 = k2 < lim;
            uint32_t v[VEC], u[VEC];
            const uint32_t* row = mat + (rb + k) * width + c0;
            const uint32_t* row2 = mat + (rb + (has2 ? k2 : k)) * width + c0;
            if (VEC == 4) {
                const uint4 t = *reinterpret_cast<const uint4*>(row);
                const uint4 t2 = *reinterpret_cast<const uint4*>(row2);
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
                u[0] = t2.x; u[1 % VEC] = t2.y; u[2 % VEC] = t2.z; u[3 % VEC] = t2.w;
            } else {
                v[0] = row[0];
                u[0] = row2[0];
            }
            const uint4 d = *reinterpret_cast<const uint4*>(sd + 4 * k);
            uint4 g = *reinterpret_cast<const uint4*>(sd + 4 * (has2 ? k2 : k));
            if (!has2) g = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                acc[e][0] = bb::add(acc[e][0], reduce2((uint64_t)d.x * v[e] + (uint64_t)g.x * u[e]));
                acc[e][1] = bb::add(acc[e][1], reduce2((uint64_t)d.y * v[e] + (uint64_t)g.y * u[e]));
                acc[e][2] = bb::add(acc[e][2], reduce2((uint64_t)d.z * v[e] + (uint64_t)g.z * u[e]));
                acc[e][3] = bb::add(acc[e][3], reduce2((uint64_t)d.w * v[e] + (uint64_t)g.w * u[e]));
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < VEC; e++) {
        uint32_t* o = red + ((size_t)ty * tx_n * VEC + tx * VEC + e) * 4;
        o[0] = acc[e][0]; o[1] = acc[e][1]; o[2] = acc[e][2]; o[3] = acc[e][3];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tx_n * VEC * 4; i += 256) {
        const uint32_t col = blockIdx.y * tx_n * VEC + i / 4;
        if (col >= width) continue;
        uint32_t a = 0;
        for (uint32_t y = 0; y < ty_n; y++) a = bb::add(a, red[(size_t)y * tx_n * VEC * 4 + i]);
        partial[((uint64_t)blockIdx.x * width + col) * 4 + (i & 3)] = a;
    }
}
// ys[c] = scale * sum_b partial[b][c]   (scale = ((z/shift)^n - 1) / n, EF4)
__global__ void bary_finish_kernel(const uint32_t* __restrict__ partial, uint32_t nblocks, uint32_t width, const uint32_t* __restrict__ scale4,
                                   uint32_t* __restrict__ ys) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    ef4 acc{{0, 0, 0, 0}};
    for (uint32_t b = 0; b < nblocks; b++) acc = bb::ef_add(acc, ef_load(partial + ((uint64_t)b * width + c) * 4));
    ef_store(ys + 4 * c, bb::ef_mul(acc, ef_load(scale4)));
}

// ro[i] += apo * (rys - rr[i]) * inv_den[i]
__global__ void __launch_bounds__(256) reduce_openings_kernel(const uint32_t* __restrict__ rr, uint64_t m, const uint32_t* __restrict__ inv_den,
                                                              const uint32_t* __restrict__ rys4, const uint32_t* __restrict__ apo4, uint32_t* __restrict__ ro) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const ef4 rys = ef_load(rys4), apo = ef_load(apo4);
    const ef4 t = bb::ef_mul(bb::ef_mul(apo, bb::ef_sub(rys, ef_load(rr + 4 * i))), ef_load(inv_den + 4 * i));
    ef_store(ro + 4 * i, bb::ef_add(ef_load(ro + 4 * i), t));
}

// out[0..4) = sum_c pw[c] * ys[c] (the reduced opening sum_c alpha^c p_c(z)), out[4..8) = pw[offset] (alpha^offset);
// one CTA, the two constants reduce_openings_kernel needs, produced without leaving the device
__global__ void __launch_bounds__(256) ef_dot_kernel(const uint32_t* __restrict__ pw, const uint32_t* __restrict__ ys, uint32_t width, uint32_t offset,
                                                     uint32_t* __restrict__ out) {
    __shared__ uint32_t part[256 * 4];
    ef4 acc{{0, 0, 0, 0}};
    for (uint32_t c = threadIdx.x; c < width; c += 256) acc = bb::ef_add(acc, bb::ef_mul(ef_load(pw + 4 * c), ef_load(ys + 4 * c)));
    ef_store(part + 4 * threadIdx.x, acc);
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) ef_store(part + 4 * threadIdx.x, bb::ef_add(ef_load(part + 4 * threadIdx.x), ef_load(part + 4 * (threadIdx.x + s))));
        __syncthreads();
    }
    if (threadIdx.x < 4) {
        out[threadIdx.x] = part[threadIdx.x];
        out[4 + threadIdx.x] = pw[4 * (size_t)offset + threadIdx.x];
    }
}

}  // namespace op
