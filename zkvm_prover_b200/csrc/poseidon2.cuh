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

// round constant idx in Montgomery form minus p, as a signed value in [-p, 0): state + rcs in [-p, p)
template <int IDX>
struct RC {
    static constexpr uint32_t monty = (uint32_t)((((uint64_t)RC_CANON[IDX]) << 32) % bb::P);
    static constexpr int32_t shifted = (int32_t)monty - (int32_t)bb::P;
};

// x in [-p, p) (signed) -> x^7 canonical
BB_HD uint32_t sbox7(int32_t x) {
    int32_t x2 = bb::smul(x, x);
    int32_t x3 = bb::smul(x2, x);
    int32_t x4 = bb::smul(x2, x2);
    return bb::canon(bb::smul(x3, x4));
}

// circ(2*M4, M4, M4, M4) with M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] (p3-poseidon2 mds_light_permutation)
BB_HD void mds_light(uint32_t (&s)[16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
        uint32_t x0 = s[c], x1 = s[c + 1], x2 = s[c + 2], x3 = s[c + 3];
        uint32_t t01 = bb::red2p(x0 + x1);
        uint32_t t23 = bb::red2p(x2 + x3);
        uint32_t t = bb::add(t01, t23);
        uint32_t a = bb::add(t01, x1);               // x0 + 2 x1
        uint32_t b = bb::add(t23, x3);               // x2 + 2 x3
        s[c] = bb::add(t, a);                        // 2x0 + 3x1 + x2 + x3
        s[c + 2] = bb::add(t, b);                    // x0 + x1 + 2x2 + 3x3
        s[c + 1] = bb::add(bb::add(t, x1), bb::dbl(x2));  // x0 + 2x1 + 3x2 + x3
        s[c + 3] = bb::add(bb::add(t, x3), bb::dbl(x0));  // 3x0 + x1 + x2 + 2x3
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t t = bb::add(bb::red2p(s[k] + s[4 + k]), bb::red2p(s[8 + k] + s[12 + k]));
#pragma unroll
        for (int j = 0; j < 16; j += 4) s[j + k] = bb::add(s[j + k], t);
    }
}

template <int BASE>
BB_HD void external_round(uint32_t (&s)[16]) {
#define P2_SB(i) s[i] = sbox7((int32_t)s[i] + RC<BASE + i>::shifted);
    P2_SB(0) P2_SB(1) P2_SB(2) P2_SB(3) P2_SB(4) P2_SB(5) P2_SB(6) P2_SB(7)
    P2_SB(8) P2_SB(9) P2_SB(10) P2_SB(11) P2_SB(12) P2_SB(13) P2_SB(14) P2_SB(15)
#undef P2_SB
    mds_light(s);
}

// sum + c*x for a small integer c in [-4, 4], everything canonical
template <int C>
BB_HD uint32_t lin_small(uint32_t sum, uint32_t x) {
    constexpr uint64_t OFF = C < 0 ? (uint64_t)(-C) * bb::P : 0ull;
    uint64_t u = (uint64_t)sum + OFF + (uint64_t)((int64_t)C * (int64_t)x);  // in [0, 5p)
    uint32_t q = (uint32_t)(u >> 31);                                        // <= 4
    uint32_t r = (uint32_t)u - q * bb::P;                                    // < 2^31 + 4 * 2^27 < 2p
    return bb::red2p(r);
}

template <int IDX>
BB_HD void internal_round(uint32_t (&s)[16]) {
    s[0] = sbox7((int32_t)s[0] + RC<IDX>::shifted);
    // 16-term sum: pair sums fit in 32 bits, the rest accumulates in 64 bits; one reduction
    uint64_t acc = (uint64_t)(s[0] + s[1]) + (uint64_t)(s[2] + s[3]) + (uint64_t)(s[4] + s[5]) + (uint64_t)(s[6] + s[7]) +
                   (uint64_t)(s[8] + s[9]) + (uint64_t)(s[10] + s[11]) + (uint64_t)(s[12] + s[13]) + (uint64_t)(s[14] + s[15]);
    uint32_t q = (uint32_t)(acc >> 31);          // < 16
    uint32_t r = (uint32_t)acc - q * bb::P;      // < 2^31 + 15 * 2^27 < 2^32
    uint32_t sum = bb::red2p(bb::red2p(r));      // r < 2.07 p
    s[0] = lin_small<-2>(sum, s[0]);
    s[1] = bb::add(sum, s[1]);
    s[2] = lin_small<2>(sum, s[2]);
    s[3] = bb::add(sum, bb::div2exp<1>(s[3]));
    s[4] = lin_small<3>(sum, s[4]);
    s[5] = lin_small<4>(sum, s[5]);
    s[6] = bb::sub(sum, bb::div2exp<1>(s[6]));
    s[7] = lin_small<-3>(sum, s[7]);
    s[8] = lin_small<-4>(sum, s[8]);
    s[9] = bb::add(sum, bb::div2exp<8>(s[9]));
    s[10] = bb::add(sum, bb::div2exp<2>(s[10]));
    s[11] = bb::add(sum, bb::div2exp<3>(s[11]));
    s[12] = bb::add(sum, bb::div2exp<27>(s[12]));
    s[13] = bb::sub(sum, bb::div2exp<8>(s[13]));
    s[14] = bb::sub(sum, bb::div2exp<4>(s[14]));
    s[15] = bb::sub(sum, bb::div2exp<27>(s[15]));
}

// canonical Montgomery state in, canonical out
BB_HD void permute(uint32_t (&s)[16]) {
    mds_light(s);
    external_round<0>(s);
    external_round<16>(s);
    external_round<32>(s);
    external_round<48>(s);
    internal_round<64>(s);
    internal_round<65>(s);
    internal_round<66>(s);
    internal_round<67>(s);
    internal_round<68>(s);
    internal_round<69>(s);
    internal_round<70>(s);
    internal_round<71>(s);
    internal_round<72>(s);
    internal_round<73>(s);
    internal_round<74>(s);
    internal_round<75>(s);
    internal_round<76>(s);
    external_round<77>(s);
    external_round<93>(s);
    external_round<109>(s);
    external_round<125>(s);
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
