// Poseidon2 over BabyBear, width 16, S-box x^7, 4 + 13 + 4 rounds: the permutation behind
// openvm_stark_sdk::config::baby_bear_poseidon2::default_perm() (p3-poseidon2 0.4.3 structure with the
// Horizen-Labs RC16 round constants of zkhash-axiom 0.2.0; Cargo.lock:5708,10231 of the reference), the
// PaddingFreeSponge<_,16,8,8> leaf hash and the TruncatedPermutation<_,2,8,16> 2-to-1 compression of
// p3-symmetric 0.4.3 (Cargo.lock:5736).
//
// One permutation per thread, the 16-element state in registers, everything fully unrolled so round
// constants become immediates.  The kernel is bound by the integer pipes (about 6.5k instructions per
// permutation against 64 B of state), so the arithmetic is arranged to save instructions:
//  * S-box chain in signed Montgomery form (bb::smul: no correction between the four multiplies);
//    the round constant is pre-shifted by -p so `state + rc` lands in [-p, p) with a single IADD.
//  * internal layer: the 16-term sum is accumulated in 64 bits and reduced once; the diagonal
//    V = [-2,1,2,1/2,3,4,-1/2,-3,-4,2^-8,1/4,1/8,2^-27,-2^-8,-1/16,-2^-27] is applied with shift based
//    exact divisions (p == 1 mod 2^27) and small-integer multiply-accumulates instead of generic multiplies.
#pragma once
#include "bb31.cuh"

namespace p2 {

// canonical Horizen RC16: 4x16 initial | 13 internal | 4x16 terminal
#define P2_RC_LIST \
    0x69cbb6af, 0x46ad93f9, 0x60a00f4e, 0x6b1297cd, 0x23189afe, 0x732e7bef, 0x72c246de, 0x2c941900, 0x0557eede, 0x1580496f, 0x3a3ea77b, 0x54f3f271, 0x0f49b029, 0x47872fe1, 0x221e2e36, 0x1ab7202e, \
    0x487779a6, 0x3851c9d8, 0x38dc17c0, 0x209f8849, 0x268dcee8, 0x350c48da, 0x5b9ad32e, 0x0523272b, 0x3f89055b, 0x01e894b2, 0x13ddedde, 0x1b2ef334, 0x7507d8b4, 0x6ceeb94e, 0x52eb6ba2, 0x50642905, \
    0x05453f3f, 0x06349efc, 0x6922787c, 0x04bfff9c, 0x768c714a, 0x3e9ff21a, 0x15737c9c, 0x2229c807, 0x0d47f88c, 0x097e0ecc, 0x27eadba0, 0x2d7d29e4, 0x3502aaa0, 0x0f475fd7, 0x29fbda49, 0x018afffd, \
    0x0315b618, 0x6d4497d1, 0x1b171d9e, 0x52861abd, 0x2e5d0501, 0x3ec8646c, 0x6e5f250a, 0x148ae8e6, 0x17f5fa4a, 0x3e66d284, 0x0051aa3b, 0x483f7913, 0x2cfe5f15, 0x023427ca, 0x2cc78315, 0x1e36ea47, \
    0x5a8053c0, 0x693be639, 0x3858867d, 0x19334f6b, 0x128f0fd8, 0x4e2b1ccb, 0x61210ce0, 0x3c318939, 0x0b5b2f22, 0x2edb11d5, 0x213effdf, 0x0cac4606, 0x241af16d, \
    0x7290a80d, 0x6f7e5329, 0x598ec8a8, 0x76a859a0, 0x6559e868, 0x657b83af, 0x13271d3f, 0x1f876063, 0x0aeeae37, 0x706e9ca6, 0x46400cee, 0x72a05c26, 0x2c589c9e, 0x20bd37a7, 0x6a2d3d10, 0x20523767, \
    0x5b8fe9c4, 0x2aa501d6, 0x1e01ac3e, 0x1448bc54, 0x5ce5ad1c, 0x4918a14d, 0x2c46a83f, 0x4fcf6876, 0x61d8d5c8, 0x6ddf4ff9, 0x11fda4d3, 0x02933a8f, 0x170eaf81, 0x5a9c314f, 0x49a12590, 0x35ec52a1, \
    0x58eb1611, 0x5e481e65, 0x367125c9, 0x0eba33ba, 0x1fc28ded, 0x066399ad, 0x0cbec0ea, 0x75fd1af0, 0x50f5bf4e, 0x643d5f41, 0x6f4fe718, 0x5b3cbbde, 0x1e3afb3e, 0x296fb027, 0x45e1547b, 0x4a8db2ab, \
    0x59986d19, 0x30bcdfa3, 0x1db63932, 0x1d7c2824, 0x53b33681, 0x0673b747, 0x038a98a3, 0x2c5bce60, 0x351979cd, 0x5008fb73, 0x547bca78, 0x711af481, 0x3f93bf64, 0x644d987b, 0x3c8bcd87, 0x608758b8,
constexpr uint32_t RC_CANON[141] = {P2_RC_LIST};
#ifdef __CUDACC__
// same table addressable with run-time indices from device code (only the plain cross-check path uses it)
__device__ __constant__ uint32_t RC_CANON_DEV[141] = {P2_RC_LIST};
#endif

// round constants in Montgomery form minus p, as signed values in [-p, 0): state + rc lands in [-p, p)
struct Tables {
    int32_t ext[8][16];
    int32_t in[13];
};
constexpr Tables make_tables() {
    Tables t{};
    for (int r = 0; r < 8; r++)
        for (int i = 0; i < 16; i++) {
            const int idx = (r < 4 ? 16 * r : 77 + 16 * (r - 4)) + i;
            t.ext[r][i] = (int32_t)((((uint64_t)RC_CANON[idx]) << 32) % bb::P) - (int32_t)bb::P;
        }
    for (int r = 0; r < 13; r++) t.in[r] = (int32_t)((((uint64_t)RC_CANON[64 + r]) << 32) % bb::P) - (int32_t)bb::P;
    return t;
}
#ifdef __CUDACC__
static __device__ __constant__ Tables T_DEV = make_tables();
#endif
static constexpr Tables T_HOST = make_tables();
#ifdef __CUDA_ARCH__
#define P2_TAB p2::T_DEV
#else
#define P2_TAB p2::T_HOST
#endif

// x in [-p, p) (signed) -> x^7 canonical
BB_HD uint32_t sbox7(int32_t x) {
    int32_t x2 = bb::smul(x, x);
    int32_t x3 = bb::smul(x2, x);
    int32_t x4 = bb::smul(x2, x2);
    return bb::canon(bb::smul(x3, x4));
}

// The same with bb::smulz (three FMA-pipe instructions per product, none on the ALU pipe).  The always-zero low words
// are chained through the q of the next product of the chain (x2 -> x3 -> x4 -> x7; x4 is ordered behind x3 only for
// that), so one S-box takes a z in and hands one z out.  P2_QLEA: this S-box derives its four q's with shifts on the ALU
// pipe instead (bb::smulz_lea).
template <bool QLEA>
BB_HD uint32_t sbox7z(int32_t x, uint32_t& z) {
    if (QLEA) {
        int32_t x2 = bb::smulz_lea(x, x, z);
        int32_t x3 = bb::smulz_lea(x2, x, z);
        int32_t x4 = bb::smulz_lea(x2, x2, z);
        return bb::canon(bb::smulz_lea(x3, x4, z));
    }
    int32_t x2 = bb::smulz(x, x, z);
    int32_t x3 = bb::smulz(x2, x, z);
    int32_t x4 = bb::smulz(x2, x2, z);
    return bb::canon(bb::smulz(x3, x4, z));
}

// ---- tuning knobs (pipe assignment; see bb31.cuh "pipe model").  Defaults are the measured best.
#ifndef P2_FUSED
// S-box products: 1 = bb::smulz (IMAD.WIDE + IMAD + IMAD.WIDE with addend), 0 = bb::smul.  Measured on B200 (tools/p2_sweep_r02.sh,
// profiles/poseidon2_sweep_r02.txt): the fused form has 64 fewer ALU instructions per external round and the same FMA-pipe
// instruction count, and is 3-5 % SLOWER (4.13-4.24 vs 4.36 Gperm/s) -- the multiply-accumulate with a 64-bit addend does not issue
// at the rate of the plain IMAD.WIDE.  Kept as an experiment knob.
#define P2_FUSED 0
#endif
#ifndef P2_QLEA_MASK
#define P2_QLEA_MASK 0x0000  // external rounds, fused form: lanes whose S-box computes q with shifts (ALU pipe) instead of an IMAD
#endif
#ifndef P2_QLEA_INT
#define P2_QLEA_INT 0    // the same for the single S-box of an internal round
#endif
#ifndef P2_RC_FMA
#define P2_RC_FMA 0      // round-constant additions on the FMA pipe (1) or ALU pipe (0)
#endif
#ifndef P2_MDS_FMA_MASK
// external linear layer: which groups of additions are steered to the FMA pipe.  bit 0: t01, t23; bit 1: the doublings;
// bit 2: the 16 final `+ column sum`; bit 3: t0123; bit 4: t01123, t01233; bit 5: column sums; bits 6, 7: the M4 outputs
#define P2_MDS_FMA_MASK 0x0F
#endif
#ifndef P2_INT_SUM_FMA
#define P2_INT_SUM_FMA 3  // internal layer, ALU formulation: additions of the 16-lane sum steered to the FMA pipe (0 none, 1 first level, 2 +second/third, 3 all)
#endif
#ifndef P2_INT_LIN_FMA
#define P2_INT_LIN_FMA 1  // internal layer: doublings (1) and the 3x additions (2) of the small diagonal multiples on the FMA pipe
#endif
#ifndef P2_INT_OUT_FMA
#define P2_INT_OUT_FMA 0  // internal layer: how many of the plain `sum + x` output additions go to the FMA pipe
#endif
#ifndef P2_INT_MODE
#define P2_INT_MODE 1    // internal layer: 0 = 64-bit IMAD.WIDE formulation, 1 = ALU formulation (doublings / 32-bit shifts)
#endif

BB_HD uint32_t add_f(uint32_t a, uint32_t b) { uint32_t s = bb::fadd(a, b); return bb::umin32(s, s - bb::P); }
#define P2_ADD_LVL(lvl, a, b) (((P2_MDS_FMA_MASK >> ((lvl) - 1)) & 1) ? add_f((a), (b)) : bb::add((a), (b)))

// circ(2*M4, M4, M4, M4) with M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] (p3-poseidon2 mds_light_permutation);
// the 4x4 block uses the 11-addition schedule (t01, t23, t0123, t01123, t01233, two doublings, four sums).
BB_HD void mds_light(uint32_t (&s)[16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
        uint32_t x0 = s[c], x1 = s[c + 1], x2 = s[c + 2], x3 = s[c + 3];
        uint32_t t01 = P2_ADD_LVL(1, x0, x1);
        uint32_t t23 = P2_ADD_LVL(1, x2, x3);
        uint32_t t0123 = P2_ADD_LVL(4, t01, t23);
        uint32_t t01123 = P2_ADD_LVL(5, t0123, x1);
        uint32_t t01233 = P2_ADD_LVL(5, t0123, x3);
        uint32_t d0 = P2_ADD_LVL(2, x0, x0), d2 = P2_ADD_LVL(2, x2, x2);
        s[c + 3] = P2_ADD_LVL(7, t01233, d0);     // 3x0 + x1 + x2 + 2x3
        s[c + 1] = P2_ADD_LVL(7, t01123, d2);     // x0 + 2x1 + 3x2 + x3
        s[c] = P2_ADD_LVL(8, t01123, t01);        // 2x0 + 3x1 + x2 + x3
        s[c + 2] = P2_ADD_LVL(8, t01233, t23);    // x0 + x1 + 2x2 + 3x3
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t t = P2_ADD_LVL(6, P2_ADD_LVL(6, s[k], s[4 + k]), P2_ADD_LVL(6, s[8 + k], s[12 + k]));
#pragma unroll
        for (int j = 0; j < 16; j += 4) s[j + k] = P2_ADD_LVL(3, s[j + k], t);
    }
}

BB_HD void external_round(uint32_t (&s)[16], const int32_t* rc) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = sbox7((int32_t)(P2_RC_FMA ? bb::fadd(s[i], (uint32_t)rc[i]) : bb::aadd(s[i], (uint32_t)rc[i])));
    mds_light(s);
}
// fused-product form: z is the running always-zero word (see bb::smulz); every S-box starts from the round's incoming z
// and the sixteen outgoing ones are OR-ed together (eight 3-input LOP3) into the next round's
BB_HD void external_round_z(uint32_t (&s)[16], const int32_t* rc, uint32_t& z) {
    uint32_t zo = z;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t zi = z;
        const int32_t x = (int32_t)(P2_RC_FMA ? bb::fadd(s[i], (uint32_t)rc[i]) : bb::aadd(s[i], (uint32_t)rc[i]));
        s[i] = ((P2_QLEA_MASK >> i) & 1) ? sbox7z<true>(x, zi) : sbox7z<false>(x, zi);
        zo |= zi;
    }
    z = zo;
    mds_light(s);
}

#if P2_INT_MODE == 0
// sum + c*x for a small integer c in [-4, 4], everything canonical (64-bit multiply-accumulate formulation)
template <int C>
BB_HD uint32_t lin_small(uint32_t sum, uint32_t x) {
    constexpr uint64_t OFF = C < 0 ? (uint64_t)(-C) * bb::P : 0ull;
    uint64_t u = (uint64_t)sum + OFF + (uint64_t)((int64_t)C * (int64_t)x);  // in [0, 5p)
    uint32_t q = (uint32_t)(u >> 31);                                        // <= 4
    uint32_t r = (uint32_t)u - q * bb::P;                                    // < 2^31 + 4 * 2^27 < 2p
    return bb::red2p(r);
}
template <int K>
BB_HD uint32_t div2(uint32_t x) { return bb::div2exp<K>(x); }
BB_HD uint32_t sum16(const uint32_t (&s)[16]) {
    // pair sums fit in 32 bits, the rest accumulates in 64 bits; one reduction
    uint64_t acc = (uint64_t)(s[0] + s[1]);
    acc = bb::fadd64(acc, s[2] + s[3]);
    acc = bb::fadd64(acc, s[4] + s[5]);
    acc = bb::fadd64(acc, s[6] + s[7]);
    uint64_t acc2 = (uint64_t)(s[8] + s[9]);
    acc2 = bb::fadd64(acc2, s[10] + s[11]);
    acc2 = bb::fadd64(acc2, s[12] + s[13]);
    acc2 = bb::fadd64(acc2, s[14] + s[15]);
    acc += acc2;
    uint32_t q = (uint32_t)(acc >> 31);          // < 16
    uint32_t r = (uint32_t)acc - q * bb::P;      // < 2^31 + 15 * 2^27 < 2^32
    return bb::red2p(bb::red2p(r));              // r < 2.07 p
}
#else
// ALU formulation: the FMA pipe is saturated by the S-box multiplies, so small multiples are built from canonical
// doublings and the exact divisions by 2^K use 32-bit arithmetic only:
//   x / 2^K = ((x + 2^K - 1) >> K) + m * (15 << (27 - K)),  m = (-x) mod 2^K     (p = 15 * 2^27 + 1)
template <int C>
BB_HD uint32_t lin_small(uint32_t sum, uint32_t x) {
    constexpr int A = C < 0 ? -C : C;
    uint32_t d = P2_INT_LIN_FMA >= 1 ? add_f(x, x) : bb::dbl(x);
    uint32_t m = A == 2 ? d : (A == 3 ? (P2_INT_LIN_FMA >= 2 ? add_f(d, x) : bb::add(d, x)) : (P2_INT_LIN_FMA >= 1 ? add_f(d, d) : bb::dbl(d)));
    return C < 0 ? bb::sub(sum, m) : bb::add(sum, m);
}
template <int K>
BB_HD uint32_t div2(uint32_t x) {
    constexpr uint32_t MASK = (1u << K) - 1u;
    uint32_t m = (0u - x) & MASK;
    return ((x + MASK) >> K) + m * (15u << (27 - K));
}
#define P2_SUM_ADD(lvl, a, b) ((P2_INT_SUM_FMA >= (lvl)) ? add_f((a), (b)) : bb::add((a), (b)))
#ifndef P2_INT_SUM_WIDE
#define P2_INT_SUM_WIDE 0  // experiment: 1 = the 16-lane sum in a 64-bit accumulator (8 plain pair sums + 7 IMAD.WIDE + one reduction: 21 instructions instead of 30)
#endif
#if P2_INT_SUM_WIDE
BB_HD uint32_t sum16(const uint32_t (&s)[16]) {
    uint64_t acc = (uint64_t)(s[0] + s[1]);
    acc = bb::fadd64(acc, s[2] + s[3]);
    acc = bb::fadd64(acc, s[4] + s[5]);
    acc = bb::fadd64(acc, s[6] + s[7]);
    uint64_t acc2 = (uint64_t)(s[8] + s[9]);
    acc2 = bb::fadd64(acc2, s[10] + s[11]);
    acc2 = bb::fadd64(acc2, s[12] + s[13]);
    acc2 = bb::fadd64(acc2, s[14] + s[15]);
    acc += acc2;
    uint32_t q = (uint32_t)(acc >> 31);          // < 32
    uint32_t r = (uint32_t)acc - q * bb::P;      // < 2^31 + 31 * 2^27 < 2^32
    return bb::red2p(bb::red2p(r));              // r < 2^32 < 2.14 p
}
#else
BB_HD uint32_t sum16(const uint32_t (&s)[16]) {
    uint32_t a0 = P2_SUM_ADD(1, s[0], s[1]), a1 = P2_SUM_ADD(1, s[2], s[3]), a2 = P2_SUM_ADD(1, s[4], s[5]), a3 = P2_SUM_ADD(1, s[6], s[7]);
    uint32_t a4 = P2_SUM_ADD(1, s[8], s[9]), a5 = P2_SUM_ADD(1, s[10], s[11]), a6 = P2_SUM_ADD(1, s[12], s[13]), a7 = P2_SUM_ADD(1, s[14], s[15]);
    return P2_SUM_ADD(3, P2_SUM_ADD(2, P2_SUM_ADD(2, a0, a1), P2_SUM_ADD(2, a2, a3)), P2_SUM_ADD(2, P2_SUM_ADD(2, a4, a5), P2_SUM_ADD(2, a6, a7)));
}
#endif
#endif

#define P2_OUT_ADD(lvl, a, b) ((P2_INT_OUT_FMA >= (lvl)) ? add_f((a), (b)) : bb::add((a), (b)))
BB_HD void internal_linear(uint32_t (&s)[16]);
BB_HD void internal_round(uint32_t (&s)[16], int32_t rc) {
    s[0] = sbox7((int32_t)(P2_RC_FMA ? bb::fadd(s[0], (uint32_t)rc) : bb::aadd(s[0], (uint32_t)rc)));
    internal_linear(s);
}
BB_HD void internal_round_z(uint32_t (&s)[16], int32_t rc, uint32_t& z) {
    const int32_t x = (int32_t)(P2_RC_FMA ? bb::fadd(s[0], (uint32_t)rc) : bb::aadd(s[0], (uint32_t)rc));
    s[0] = P2_QLEA_INT ? sbox7z<true>(x, z) : sbox7z<false>(x, z);
    internal_linear(s);
}
BB_HD void internal_linear(uint32_t (&s)[16]) {
    const uint32_t sum = sum16(s);
    s[0] = lin_small<-2>(sum, s[0]);
    s[1] = P2_OUT_ADD(1, sum, s[1]);
    s[2] = lin_small<2>(sum, s[2]);
    s[3] = P2_OUT_ADD(2, sum, div2<1>(s[3]));
    s[4] = lin_small<3>(sum, s[4]);
    s[5] = lin_small<4>(sum, s[5]);
    s[6] = bb::sub(sum, div2<1>(s[6]));
    s[7] = lin_small<-3>(sum, s[7]);
    s[8] = lin_small<-4>(sum, s[8]);
    s[9] = P2_OUT_ADD(3, sum, div2<8>(s[9]));
    s[10] = P2_OUT_ADD(4, sum, div2<2>(s[10]));
    s[11] = P2_OUT_ADD(5, sum, div2<3>(s[11]));
    s[12] = P2_OUT_ADD(6, sum, div2<27>(s[12]));
    s[13] = bb::sub(sum, div2<8>(s[13]));
    s[14] = bb::sub(sum, div2<4>(s[14]));
    s[15] = bb::sub(sum, div2<27>(s[15]));
}

// canonical Montgomery state in, canonical out.  The round loops are deliberately NOT unrolled: the fully
// unrolled body (~90 KB of SASS) thrashed the instruction cache (ncu: `no_instruction` was the top stall);
// rolled, the whole permutation is ~10 KB and the round constants come from constant memory.
BB_HD void permute(uint32_t (&s)[16]) {
    mds_light(s);
#if P2_FUSED && defined(__CUDA_ARCH__)
    uint32_t z = bb::K_ZERO;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) external_round_z(s, P2_TAB.ext[4 * half + r], z);
        if (half == 0) {
#pragma unroll 1
            for (int r = 0; r < 13; r++) internal_round_z(s, P2_TAB.in[r], z);
        }
    }
    s[15] |= z;  // z == 0; this use is what keeps the low halves of the products alive (see bb::smulz)
#else
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) external_round(s, P2_TAB.ext[4 * half + r]);
        if (half == 0) {
#pragma unroll 1
            for (int r = 0; r < 13; r++) internal_round(s, P2_TAB.in[r]);
        }
    }
#endif
}

// straightforward variant (generic Montgomery multiplies everywhere); kept as an in-library cross-check
// of the optimised path above (tests compare both against the oracle)
BB_HD void permute_plain(uint32_t (&s)[16]) {
    constexpr uint32_t two = 0x1ffffffcu;  // monty(2)
    auto mds = [&]() {
        for (int c = 0; c < 16; c += 4) {
            uint32_t a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
            uint32_t t = bb::add(bb::add(a, b), bb::add(cc, d));
            s[c] = bb::add(bb::add(t, a), bb::add(b, b));
            s[c + 1] = bb::add(bb::add(t, b), bb::add(cc, cc));
            s[c + 2] = bb::add(bb::add(t, cc), bb::add(d, d));
            s[c + 3] = bb::add(bb::add(t, d), bb::add(a, a));
        }
        for (int k = 0; k < 4; k++) {
            uint32_t t = bb::add(bb::add(s[k], s[4 + k]), bb::add(s[8 + k], s[12 + k]));
            for (int j = 0; j < 16; j += 4) s[j + k] = bb::add(s[j + k], t);
        }
    };
    auto sb = [&](uint32_t x) {
        uint32_t x2 = bb::mul(x, x), x3 = bb::mul(x2, x), x4 = bb::mul(x2, x2);
        return bb::mul(x3, x4);
    };
#ifdef __CUDA_ARCH__
    auto rc = [&](int i) { return (uint32_t)((((uint64_t)RC_CANON_DEV[i]) << 32) % bb::P); };
#else
    auto rc = [&](int i) { return (uint32_t)((((uint64_t)RC_CANON[i]) << 32) % bb::P); };
#endif
    const uint32_t i2 = bb::HALF, i4 = bb::mul(i2, i2), i8 = bb::mul(i4, i2), i16 = bb::mul(i8, i2), i256 = bb::mul(i16, i16);
    const uint32_t i27 = bb::inv(bb::to_monty(1u << 27));
    const uint32_t three = bb::add(two, bb::ONE), four = bb::add(two, two);
    const uint32_t V[16] = {bb::neg(two), bb::ONE, two, i2, three, four, bb::neg(i2), bb::neg(three), bb::neg(four), i256, i4, i8, i27, bb::neg(i256), bb::neg(i16), bb::neg(i27)};
    mds();
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sb(bb::add(s[i], rc(16 * r + i)));
        mds();
    }
    for (int r = 0; r < 13; r++) {
        s[0] = sb(bb::add(s[0], rc(64 + r)));
        uint32_t t = 0;
        for (int i = 0; i < 16; i++) t = bb::add(t, s[i]);
        for (int i = 0; i < 16; i++) s[i] = bb::add(t, bb::mul(V[i], s[i]));
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sb(bb::add(s[i], rc(77 + 16 * r + i)));
        mds();
    }
}

}  // namespace p2
