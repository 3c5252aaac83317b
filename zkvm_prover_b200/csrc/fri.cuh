// K6: FRI commit-phase kernels and the device-resident DuplexChallenger.
// Replaces p3_fri::prover::commit_phase / TwoAdicFriGenericConfig::fold_matrix (v1-era p3-fri that produced
// the reference's fixtures) and p3_challenger::DuplexChallenger<BabyBear,Perm,16,8> (p3-challenger 0.4.3,
// Cargo.lock:5576).  The challenger state stays in HBM so a whole commit phase (commit -> observe root ->
// sample beta -> fold, ~20 rounds) is enqueued without a host round trip.
#pragma once
#include "poseidon2.cuh"

namespace fri {

// out[i] = (lo + hi)/2 + (beta/2) * g^-bitrev(i) * (lo - hi),  (lo, hi) = in[2i], in[2i+1]   (EF4)
//   == (1/2 + beta/2 g_inv^i') lo + (1/2 - beta/2 g_inv^i') hi   of fold_matrix, evaluated in the field.
// ginv tables: two-level powers of g^-1, g = two_adic_generator(log2(len)).
// add_mode: 0 none, 1 out += add[i], 2 out += beta^2 * add[i] (later p3-fri versions)
__global__ void __launch_bounds__(256) fold_kernel(const uint4* __restrict__ in, uint64_t half_len, int log_half, const uint32_t* __restrict__ beta_dev,
                                                   const uint32_t* __restrict__ ginv_lo, const uint32_t* __restrict__ ginv_hi,
                                                   const uint4* __restrict__ add, int add_mode, uint4* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half_len) return;
    uint4 l = in[2 * i], h = in[2 * i + 1];
    bb::ef4 lo = {{l.x, l.y, l.z, l.w}}, hi = {{h.x, h.y, h.z, h.w}};
    bb::ef4 beta = {{beta_dev[0], beta_dev[1], beta_dev[2], beta_dev[3]}};
    uint64_t e = bb::bitrev((uint32_t)i, log_half);
    uint32_t pw = bb::mul(__ldg(ginv_lo + (e & 4095)), __ldg(ginv_hi + (e >> 12)));  // g^-e / 2 (1/2 folded into hi[])
    bb::ef4 hb = bb::ef_scale(beta, pw);
    bb::ef4 s = bb::ef_add(lo, hi), d = bb::ef_sub(lo, hi);
    bb::ef4 r = bb::ef_add(bb::ef_scale(s, bb::HALF), bb::ef_mul(hb, d));
    if (add_mode) {
        uint4 a4 = add[i];
        bb::ef4 a = {{a4.x, a4.y, a4.z, a4.w}};
        if (add_mode == 2) a = bb::ef_mul(bb::ef_mul(beta, beta), a);
        r = bb::ef_add(r, a);
    }
    out[i] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
}

// ---- DuplexChallenger state in device memory
struct ChalState {
    uint32_t st[16];
    uint32_t in[8];
    uint32_t nin;
    uint32_t out[8];
    uint32_t nout;
};

__device__ __forceinline__ void duplex(ChalState& c) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = c.st[i];
#pragma unroll
    for (int i = 0; i < 8; i++) if ((uint32_t)i < c.nin) s[i] = c.in[i];
    c.nin = 0;
    p2::permute(s);
#pragma unroll
    for (int i = 0; i < 16; i++) c.st[i] = s[i];
#pragma unroll
    for (int i = 0; i < 8; i++) c.out[i] = s[i];
    c.nout = 8;
}
__device__ __forceinline__ void observe(ChalState& c, uint32_t v) {
    c.nout = 0;
    c.in[c.nin++] = v;
    if (c.nin == 8) duplex(c);
}
__device__ __forceinline__ uint32_t sample(ChalState& c) {
    if (c.nin || !c.nout) duplex(c);
    return c.out[--c.nout];
}

__global__ void chal_observe_kernel(ChalState* cs, const uint32_t* __restrict__ vals, uint32_t n) {
    if (threadIdx.x || blockIdx.x) return;
    ChalState c = *cs;
    for (uint32_t i = 0; i < n; i++) observe(c, vals[i]);
    *cs = c;
}
// sample n base elements (in order); if bits > 0 the single sample is reduced to its low `bits` bits
__global__ void chal_sample_kernel(ChalState* cs, uint32_t* __restrict__ out, uint32_t n, uint32_t bits) {
    if (threadIdx.x || blockIdx.x) return;
    ChalState c = *cs;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t v = sample(c);
        out[i] = bits ? (bb::from_monty(v) & ((1u << bits) - 1u)) : v;
    }
    *cs = c;
}
// FRI round glue: observe the 8 elements of the root, sample beta (4 base elements)
__global__ void chal_fri_round_kernel(ChalState* cs, const uint32_t* __restrict__ root, uint32_t* __restrict__ beta_out) {
    if (threadIdx.x || blockIdx.x) return;
    ChalState c = *cs;
    for (int i = 0; i < 8; i++) observe(c, root[i]);
    for (int i = 0; i < 4; i++) beta_out[i] = sample(c);
    *cs = c;
}
// PoW grinding: thread t tests witness base + t; the smallest passing witness wins (atomicMin)
__global__ void __launch_bounds__(256) chal_grind_kernel(const ChalState* __restrict__ cs, uint32_t bits, uint32_t base, uint32_t count, uint32_t* best) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint32_t w = base + t;
    if (w >= bb::P || w >= *best) return;
    ChalState c = *cs;
    observe(c, bb::to_monty(w));
    uint32_t v = bb::from_monty(sample(c)) & ((1u << bits) - 1u);
    if (v == 0) atomicMin(best, w);
}
__global__ void chal_observe_witness_kernel(ChalState* cs, const uint32_t* best, uint32_t bits) {
    if (threadIdx.x || blockIdx.x) return;
    ChalState c = *cs;
    observe(c, bb::to_monty(*best));
    (void)sample(c);  // check_witness consumes the sample (sample_bits)
    (void)bits;
    *cs = c;
}

}  // namespace fri
