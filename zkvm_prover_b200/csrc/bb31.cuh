// BabyBear (p = 2^31 - 2^27 + 1) in 32-bit integer lanes, Montgomery form R = 2^32.
// Same in-memory representation as p3_baby_bear::BabyBear = MontyField31<BabyBearParameters>
// (p3-monty-31 0.4.3, Cargo.lock:5685 of the reference; not vendored).  Written for sm_100a:
// IMAD.WIDE / IMAD / IMAD.HI on the fma pipe, IADD3 / IMNMX on the alu pipe.
//
// Every function is __host__ __device__ so tests/host_check.cu can run the very same arithmetic on
// the CPU against the oracle (arithmetic validation only; the library has no CPU path).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

namespace bb {

constexpr uint32_t P = 0x78000001u;
constexpr uint32_t MU = 0x88000001u;       // P^-1 mod 2^32
constexpr uint32_t ONE = 0x0ffffffeu;      // R mod P
constexpr uint32_t R2 = 0x45dddde3u;       // R^2 mod P
constexpr uint32_t HALF = 0x07ffffffu;     // monty(1/2)
constexpr int TWO_ADICITY = 27;
constexpr uint32_t GEN27 = 0x1a427a41u;    // canonical 31^15: generator of the 2^27 subgroup

BB_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// ---- pipe model (sm_100a, measured: tools/pipe_microbench.cu -> profiles/pipe_microbench_r01.txt)
//   FMA pipe : IMAD and IMAD.WIDE 64 lanes/clk/SM, IMAD.HI 32 lanes/clk/SM
//   ALU pipe : IADD3, VIADDMNMX, SHF, LOP3 64 lanes/clk/SM; the two pipes issue concurrently.
// A Montgomery product costs 10 pipe-cycles per warp however it is written (IMAD.WIDE + IMAD + IMAD.HI + IADD3;
// ptxas does not fuse a 64-bit addend into IMAD.WIDE for register operands), so the remaining lever is keeping both
// pipes equally busy.  ptxas likes to turn plain additions into IMAD.IADD, which overloads the FMA pipe that already
// carries the multiplies; `aadd` therefore pins an addition to the ALU pipe (a three-input IADD3 with an opaque zero
// from constant memory -- IMAD cannot express it), `fadd` pins one to the FMA pipe (a*1+b with an opaque one).
#ifdef __CUDACC__
static __device__ __constant__ uint32_t K_ONE = 1u;
static __device__ __constant__ uint32_t K_ZERO = 0u;
static __device__ __constant__ uint32_t K_MU = 0x88000001u;   // P^-1 mod 2^32, opaque to ptxas (see smulz)
static __device__ __constant__ uint32_t K_27 = 27u, K_31 = 31u;
#endif
#ifdef __CUDA_ARCH__
BB_HD uint32_t fadd(uint32_t a, uint32_t b) { return a * K_ONE + b; }
#ifndef BB_AADD_MODE
#define BB_AADD_MODE 0  // how an addition is pinned to the ALU pipe: 0 = three-input IADD3 with an opaque zero, 1 = add.cc (carry-out form), 2 = not pinned
#endif
#if BB_AADD_MODE == 0
BB_HD uint32_t aadd(uint32_t a, uint32_t b) { return a + b + K_ZERO; }
BB_HD uint32_t asub(uint32_t a, uint32_t b) { return a - b + K_ZERO; }
#elif BB_AADD_MODE == 1
BB_HD uint32_t aadd(uint32_t a, uint32_t b) { uint32_t r; asm("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
BB_HD uint32_t asub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#else
BB_HD uint32_t aadd(uint32_t a, uint32_t b) { return a + b; }
BB_HD uint32_t asub(uint32_t a, uint32_t b) { return a - b; }
#endif
BB_HD uint64_t fadd64(uint64_t acc, uint32_t x) { return (uint64_t)x * K_ONE + acc; }
BB_HD int32_t mulhi32(int32_t a, int32_t b) { return __mulhi(a, b); }
#else
BB_HD uint32_t fadd(uint32_t a, uint32_t b) { return a + b; }
BB_HD uint32_t aadd(uint32_t a, uint32_t b) { return a + b; }
BB_HD uint32_t asub(uint32_t a, uint32_t b) { return a - b; }
BB_HD uint64_t fadd64(uint64_t acc, uint32_t x) { return acc + x; }
BB_HD int32_t mulhi32(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 32); }
#endif

// canonical [0,p) x [0,p) -> [0,p)
BB_HD uint32_t add(uint32_t a, uint32_t b) { uint32_t s = aadd(a, b); return umin32(s, s - P); }
BB_HD uint32_t sub(uint32_t a, uint32_t b) { uint32_t d = asub(a, b); return umin32(d, d + P); }
BB_HD uint32_t neg(uint32_t a) { return a ? P - a : 0u; }
BB_HD uint32_t dbl(uint32_t a) { return add(a, a); }
// any value < 2p -> [0,p)
BB_HD uint32_t red2p(uint32_t a) { return umin32(a, a - P); }
// signed representative in (-p,p) (as int32 bits) -> [0,p)
BB_HD uint32_t canon(int32_t r) { uint32_t u = (uint32_t)r; return umin32(u, u + P); }

// signed Montgomery product: |a*b| < 2^31 * p  ->  representative in (-p, p), no correction step.
//   t = a*b ; q = lo(t) * p^-1 (mod 2^32, signed) ; (t - q*p) / 2^32 = hi(t) - hi(q*p)  (the low words cancel)
BB_HD int32_t smul(int32_t a, int32_t b) {
    int64_t t = (int64_t)a * (int64_t)b;
    int32_t q = (int32_t)((uint32_t)t * MU);
    int32_t qh = mulhi32(q, (int32_t)P);
    return (int32_t)asub((uint32_t)(t >> 32), (uint32_t)qh);
}
// The same product in three FMA-pipe instructions and no ALU instruction: IMAD.WIDE t = a*b; IMAD q = lo(t)*mu (+ z);
// IMAD.WIDE r = q*(-p) + t, whose low word is zero by construction and whose high word is the result.  ptxas only keeps
// the multiply-accumulate whole (64-bit addend next to an immediate multiplicand) if BOTH halves of r are used and if it
// cannot prove the low half is zero: mu therefore comes from constant memory, and the low word is handed back in `z`
// (always 0 at run time) for the caller to fold into a later instruction that has a free addend slot (here: the next
// product's q).  Otherwise it falls back to IMAD.HI + IADD3/IADD3.X, which is slower than smul().
#ifdef __CUDA_ARCH__
BB_HD int32_t smulz(int32_t a, int32_t b, uint32_t& z) {
    int64_t t;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    const uint32_t q = (uint32_t)t * K_MU + z;
    const int64_t r = (int64_t)(int32_t)q * (int64_t)(-(int32_t)P) + t;
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r));
    z = lo;
    return (int32_t)hi;
}
// q from shifts on the ALU pipe instead of an IMAD (mu = 2^31 + 2^27 + 1): 8 FMA-pipe clocks + 3-4 ALU instructions.
// No addend slot is free here, so the incoming z is OR-ed into the outgoing one (ptxas merges two of these into one LOP3).
BB_HD int32_t smulz_lea(int32_t a, int32_t b, uint32_t& z) {
    int64_t t;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    const uint32_t l = (uint32_t)t;
    const uint32_t q = l + (l << K_27) + (l << K_31);  // shift counts from constant memory: with immediates ptxas turns this back into IMADs
    const int64_t r = (int64_t)(int32_t)q * (int64_t)(-(int32_t)P) + t;
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(r));
    z |= lo;
    return (int32_t)hi;
}
#else
BB_HD int32_t smulz(int32_t a, int32_t b, uint32_t& z) { (void)z; return smul(a, b); }
BB_HD int32_t smulz_lea(int32_t a, int32_t b, uint32_t& z) { (void)z; return smul(a, b); }
#endif
// canonical operands (any pair with |a*b| < 2^31 p, taken as signed) -> [0,p)
BB_HD uint32_t mul(uint32_t a, uint32_t b) { return canon(smul((int32_t)a, (int32_t)b)); }
// Montgomery reduction of 0 <= t < 2^31 * p -> [0,p)
BB_HD uint32_t reduce(uint64_t t) {
    int32_t q = (int32_t)((uint32_t)t * MU);
    int32_t qh = mulhi32(q, (int32_t)P);
    return canon((int32_t)((uint32_t)(t >> 32) - (uint32_t)qh));
}

BB_HD uint32_t to_monty(uint32_t x) { return mul(x % P, R2); }
BB_HD uint32_t from_monty(uint32_t m) { return reduce((uint64_t)m); }

BB_HD uint32_t pow(uint32_t a, uint64_t e) {
    uint32_t r = ONE;
    while (e) {
        if (e & 1) r = mul(r, a);
        a = mul(a, a);
        e >>= 1;
    }
    return r;
}
BB_HD uint32_t inv(uint32_t a) { return pow(a, (uint64_t)P - 2); }

// x * 2^-k (k in 1..27) for canonical x: p == 1 mod 2^27, so (x + ((-x) mod 2^k) * p) is divisible by 2^k
// and the quotient is < p.  (Montgomery form is preserved by multiplication with a field constant.)
template <int K>
BB_HD uint32_t div2exp(uint32_t x) {
    uint32_t m = (0u - x) & ((1u << K) - 1u);
    uint64_t v = (uint64_t)m * P + x;
    return (uint32_t)(v >> K);
}

// two_adic_generator(bits), Montgomery form (p3-field TwoAdicField::two_adic_generator)
BB_HD uint32_t two_adic_generator(int bits) {
    uint32_t g = to_monty(GEN27);
    for (int i = bits; i < TWO_ADICITY; i++) g = mul(g, g);
    return g;
}

BB_HD uint32_t bitrev(uint32_t x, int bits) {
    if (bits == 0) return 0;
#ifdef __CUDA_ARCH__
    return __brev(x) >> (32 - bits);
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

// ---- EF4 = F[x]/(x^4 - 11): p3 BinomialExtensionField<BabyBear, 4>, coefficients low -> high
struct ef4 {
    uint32_t c[4];
};
constexpr uint32_t W11 = 0x37ffffe9u;      // monty(11)
BB_HD ef4 ef_add(const ef4& a, const ef4& b) {
    ef4 r;
    for (int i = 0; i < 4; i++) r.c[i] = add(a.c[i], b.c[i]);
    return r;
}
BB_HD ef4 ef_sub(const ef4& a, const ef4& b) {
    ef4 r;
    for (int i = 0; i < 4; i++) r.c[i] = sub(a.c[i], b.c[i]);
    return r;
}
BB_HD ef4 ef_scale(const ef4& a, uint32_t k) {
    ef4 r;
    for (int i = 0; i < 4; i++) r.c[i] = mul(a.c[i], k);
    return r;
}
// schoolbook with 64-bit accumulation: each partial product < p^2, up to 4 of them plus a reduced
// high part scaled by 11 would overflow, so high and low halves are reduced separately.
BB_HD ef4 ef_mul(const ef4& a, const ef4& b) {
    // c_k = sum_{i+j=k} a_i b_j ; result_k = c_k + 11 * c_{k+4}
    uint32_t lo[4], hi[3];
    // products are < p^2 < 2^62; two of them fit in 2^63, reduce pairwise
    lo[0] = mul(a.c[0], b.c[0]);
    lo[1] = add(mul(a.c[0], b.c[1]), mul(a.c[1], b.c[0]));
    lo[2] = add(add(mul(a.c[0], b.c[2]), mul(a.c[1], b.c[1])), mul(a.c[2], b.c[0]));
    lo[3] = add(add(mul(a.c[0], b.c[3]), mul(a.c[1], b.c[2])), add(mul(a.c[2], b.c[1]), mul(a.c[3], b.c[0])));
    hi[0] = add(add(mul(a.c[1], b.c[3]), mul(a.c[2], b.c[2])), mul(a.c[3], b.c[1]));
    hi[1] = add(mul(a.c[2], b.c[3]), mul(a.c[3], b.c[2]));
    hi[2] = mul(a.c[3], b.c[3]);
    ef4 r;
    r.c[0] = add(lo[0], mul(hi[0], W11));
    r.c[1] = add(lo[1], mul(hi[1], W11));
    r.c[2] = add(lo[2], mul(hi[2], W11));
    r.c[3] = lo[3];
    return r;
}

// counter-based synthetic data shared with the oracle (oracle/bb_oracle.c orc_fill / orc_checksum)
BB_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

}  // namespace bb
