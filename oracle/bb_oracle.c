/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the B200 hot path (plain C + OpenMP).
 *
 * A restatement of the algorithms the reference reaches through `sdk.prove`
 * (/root/reference/crates/prover/src/prover/mod.rs:355-357).  The arithmetic itself lives in
 * crates that are NOT vendored under /root/reference (pinned by Cargo.lock):
 *   p3-baby-bear/p3-monty-31/p3-field 0.4.3 (Cargo.lock:5545,5685,5605)   -> bb_* field ops
 *   p3-dft 0.4.3 (Cargo.lock:5590)  TwoAdicSubgroupDft / NaiveDft          -> orc_dft*, orc_coset_lde
 *   p3-poseidon2 0.4.3 + zkhash-axiom 0.2.0 (Cargo.lock:5708,10231)        -> orc_permute
 *   p3-symmetric 0.4.3 (Cargo.lock:5736) PaddingFreeSponge<16,8,8>,
 *                                        TruncatedPermutation<2,8,16>      -> orc_hash_row, orc_compress
 *   p3-merkle-tree / p3-commit (v1 era)  MerkleTreeMmcs                    -> orc_merkle_*
 *   p3-fri (v1 era) fold_matrix / commit_phase                             -> orc_fri_*
 *   p3-challenger 0.4.3 (Cargo.lock:5576) DuplexChallenger<_,_,16,8>       -> orc_chal_*
 * Pinned against the reference's own proof fixture (tests/golden/chunk_proof_phase2_kats.json,
 * mined by oracle/mine_fixture.py) and Plonky3's published Poseidon2 test vector; see
 * tests/test_oracle_kats.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (zkvm_prover_b200/) never does.
 *
 * All field elements crossing this API are Montgomery-form uint32_t (R = 2^32), i.e. the
 * in-memory representation of p3's BabyBear.  Matrices are row-major rows x width.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P 0x78000001u
#define MU 0x88000001u /* P^-1 mod 2^32 */
#define R_MOD_P 0x0ffffffeu /* 2^32 mod P  == monty(1) */
#define R2_MOD_P 0x45dddde3u /* 2^64 mod P */

typedef uint32_t u32;
typedef uint64_t u64;

static inline u32 bb_add(u32 a, u32 b) { u32 s = a + b; return s >= P ? s - P : s; }
static inline u32 bb_sub(u32 a, u32 b) { return a >= b ? a - b : a + P - b; }
static inline u32 bb_reduce(u64 t) {
    u32 q = (u32)t * MU;
    u64 qp = (u64)q * P;
    u32 hi = (u32)(t >> 32), qh = (u32)(qp >> 32);
    u32 r = hi - qh;
    return hi < qh ? r + P : r;
}
static inline u32 bb_mul(u32 a, u32 b) { return bb_reduce((u64)a * b); }
static inline u32 bb_to_monty(u32 x) { return bb_mul(x % P, R2_MOD_P); }
static inline u32 bb_from_monty(u32 m) { return bb_reduce((u64)m); }
static u32 bb_pow(u32 a, u64 e) {
    u32 r = R_MOD_P;
    while (e) { if (e & 1) r = bb_mul(r, a); a = bb_mul(a, a); e >>= 1; }
    return r;
}
static inline u32 bb_inv(u32 a) { return bb_pow(a, (u64)P - 2); }

u32 orc_to_monty(u32 x) { return bb_to_monty(x); }
u32 orc_from_monty(u32 x) { return bb_from_monty(x); }
u32 orc_mul(u32 a, u32 b) { return bb_mul(a, b); }
u32 orc_add(u32 a, u32 b) { return bb_add(a, b); }
u32 orc_sub(u32 a, u32 b) { return bb_sub(a, b); }
u32 orc_inv(u32 a) { return bb_inv(a); }
u32 orc_pow(u32 a, u64 e) { return bb_pow(a, e); }

/* two_adic_generator(bits) = (31^15)^(2^(27-bits)), Montgomery form */
u32 orc_two_adic_generator(u32 bits) {
    u32 g = bb_pow(bb_to_monty(31), 15);
    for (u32 i = bits; i < 27; i++) g = bb_mul(g, g);
    return g;
}

static inline u32 bitrev32(u32 x, u32 bits) {
    if (!bits) return 0;
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    x = (x >> 16) | (x << 16);
    return x >> (32 - bits);
}

/* ------------------------------------------------------------------ EF4 = F[x]/(x^4-11) */
typedef struct { u32 c[4]; } ef4;
static u32 W11; /* monty(11) */
static inline ef4 ef_add(ef4 a, ef4 b) { ef4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_add(a.c[i], b.c[i]); return r; }
static inline ef4 ef_sub(ef4 a, ef4 b) { ef4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_sub(a.c[i], b.c[i]); return r; }
static inline ef4 ef_scale(ef4 a, u32 k) { ef4 r; for (int i = 0; i < 4; i++) r.c[i] = bb_mul(a.c[i], k); return r; }
static ef4 ef_mul(ef4 a, ef4 b) {
    u32 t[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t[i + j] = bb_add(t[i + j], bb_mul(a.c[i], b.c[j]));
    ef4 r;
    for (int i = 0; i < 4; i++) r.c[i] = i < 3 ? bb_add(t[i], bb_mul(W11, t[i + 4])) : t[i];
    return r;
}
void orc_ef_mul(const u32* a, const u32* b, u32* out) {
    ef4 x, y; memcpy(x.c, a, 16); memcpy(y.c, b, 16);
    ef4 r = ef_mul(x, y); memcpy(out, r.c, 16);
}

/* ------------------------------------------------------------------ Poseidon2 width 16 */
/* Horizen-Labs RC16 (zkhash) in canonical form; cross-checked in tests against the Grain-LFSR
 * derivation in oracle/pyref.py.  Layout: 4x16 initial | 13 internal | 4x16 terminal. */
static const u32 RC_CANON[141] = {
    0x69cbb6af, 0x46ad93f9, 0x60a00f4e, 0x6b1297cd, 0x23189afe, 0x732e7bef, 0x72c246de, 0x2c941900, 0x0557eede, 0x1580496f, 0x3a3ea77b, 0x54f3f271, 0x0f49b029, 0x47872fe1, 0x221e2e36, 0x1ab7202e,
    0x487779a6, 0x3851c9d8, 0x38dc17c0, 0x209f8849, 0x268dcee8, 0x350c48da, 0x5b9ad32e, 0x0523272b, 0x3f89055b, 0x01e894b2, 0x13ddedde, 0x1b2ef334, 0x7507d8b4, 0x6ceeb94e, 0x52eb6ba2, 0x50642905,
    0x05453f3f, 0x06349efc, 0x6922787c, 0x04bfff9c, 0x768c714a, 0x3e9ff21a, 0x15737c9c, 0x2229c807, 0x0d47f88c, 0x097e0ecc, 0x27eadba0, 0x2d7d29e4, 0x3502aaa0, 0x0f475fd7, 0x29fbda49, 0x018afffd,
    0x0315b618, 0x6d4497d1, 0x1b171d9e, 0x52861abd, 0x2e5d0501, 0x3ec8646c, 0x6e5f250a, 0x148ae8e6, 0x17f5fa4a, 0x3e66d284, 0x0051aa3b, 0x483f7913, 0x2cfe5f15, 0x023427ca, 0x2cc78315, 0x1e36ea47,
    0x5a8053c0, 0x693be639, 0x3858867d, 0x19334f6b, 0x128f0fd8, 0x4e2b1ccb, 0x61210ce0, 0x3c318939, 0x0b5b2f22, 0x2edb11d5, 0x213effdf, 0x0cac4606, 0x241af16d,
    0x7290a80d, 0x6f7e5329, 0x598ec8a8, 0x76a859a0, 0x6559e868, 0x657b83af, 0x13271d3f, 0x1f876063, 0x0aeeae37, 0x706e9ca6, 0x46400cee, 0x72a05c26, 0x2c589c9e, 0x20bd37a7, 0x6a2d3d10, 0x20523767,
    0x5b8fe9c4, 0x2aa501d6, 0x1e01ac3e, 0x1448bc54, 0x5ce5ad1c, 0x4918a14d, 0x2c46a83f, 0x4fcf6876, 0x61d8d5c8, 0x6ddf4ff9, 0x11fda4d3, 0x02933a8f, 0x170eaf81, 0x5a9c314f, 0x49a12590, 0x35ec52a1,
    0x58eb1611, 0x5e481e65, 0x367125c9, 0x0eba33ba, 0x1fc28ded, 0x066399ad, 0x0cbec0ea, 0x75fd1af0, 0x50f5bf4e, 0x643d5f41, 0x6f4fe718, 0x5b3cbbde, 0x1e3afb3e, 0x296fb027, 0x45e1547b, 0x4a8db2ab,
    0x59986d19, 0x30bcdfa3, 0x1db63932, 0x1d7c2824, 0x53b33681, 0x0673b747, 0x038a98a3, 0x2c5bce60, 0x351979cd, 0x5008fb73, 0x547bca78, 0x711af481, 0x3f93bf64, 0x644d987b, 0x3c8bcd87, 0x608758b8,
};
static u32 RC[141];   /* Montgomery */
static u32 DIAG[16];  /* Montgomery: V = [-2,1,2,1/2,3,4,-1/2,-3,-4,1/2^8,1/4,1/8,1/2^27,-1/2^8,-1/16,-1/2^27] */
static int g_init = 0;

void orc_init(void) {
    if (g_init) return;
    for (int i = 0; i < 141; i++) RC[i] = bb_to_monty(RC_CANON[i]);
    W11 = bb_to_monty(11);
    u32 one = R_MOD_P, two = bb_add(one, one);
    u32 i2 = bb_inv(two), neg = P - 0;
    (void)neg;
#define NEG(x) bb_sub(0, (x))
    u32 i4 = bb_mul(i2, i2), i8 = bb_mul(i4, i2), i16 = bb_mul(i8, i2), i256 = bb_mul(i16, i16);
    u32 i27 = bb_inv(bb_to_monty(1u << 27));
    u32 three = bb_add(two, one), four = bb_add(two, two);
    u32 v[16] = {NEG(two), one, two, i2, three, four, NEG(i2), NEG(three), NEG(four), i256, i4, i8, i27, NEG(i256), NEG(i16), NEG(i27)};
    memcpy(DIAG, v, sizeof v);
    g_init = 1;
}
void orc_get_constants(u32* rc141, u32* diag16) { orc_init(); memcpy(rc141, RC, sizeof RC); memcpy(diag16, DIAG, sizeof DIAG); }

static inline u32 sbox7(u32 x) { u32 x2 = bb_mul(x, x), x3 = bb_mul(x2, x), x4 = bb_mul(x2, x2); return bb_mul(x3, x4); }

static inline void mds_light(u32* s) {
    /* M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] on each 4-chunk, then add column sums */
    for (int c = 0; c < 16; c += 4) {
        u32 a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
        u32 t = bb_add(bb_add(a, b), bb_add(cc, d));
        u32 o0 = bb_add(bb_add(t, a), bb_add(b, b));   /* 2a+3b+c+d */
        u32 o1 = bb_add(bb_add(t, b), bb_add(cc, cc)); /* a+2b+3c+d */
        u32 o2 = bb_add(bb_add(t, cc), bb_add(d, d));  /* a+b+2c+3d */
        u32 o3 = bb_add(bb_add(t, d), bb_add(a, a));   /* 3a+b+c+2d */
        s[c] = o0; s[c + 1] = o1; s[c + 2] = o2; s[c + 3] = o3;
    }
    for (int k = 0; k < 4; k++) {
        u32 t = bb_add(bb_add(s[k], s[4 + k]), bb_add(s[8 + k], s[12 + k]));
        for (int j = 0; j < 16; j += 4) s[j + k] = bb_add(s[j + k], t);
    }
}

void orc_permute(u32* s) {
    orc_init();
    mds_light(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox7(bb_add(s[i], RC[16 * r + i]));
        mds_light(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = sbox7(bb_add(s[0], RC[64 + r]));
        u32 t = 0;
        for (int i = 0; i < 16; i++) t = bb_add(t, s[i]);
        for (int i = 0; i < 16; i++) s[i] = bb_add(t, bb_mul(DIAG[i], s[i]));
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox7(bb_add(s[i], RC[77 + 16 * r + i]));
        mds_light(s);
    }
}
void orc_permute_many(u32* states, u64 n) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) orc_permute(states + 16 * i);
}

/* PaddingFreeSponge<16,8,8>::hash_iter over the concatenation of `nseg` row segments */
typedef struct { u32 st[16]; int fill; } sponge;
static inline void sp_init(sponge* s) { memset(s, 0, sizeof *s); }
static inline void sp_absorb(sponge* s, const u32* x, u64 n) {
    for (u64 i = 0; i < n; i++) {
        s->st[s->fill++] = x[i];
        if (s->fill == 8) { orc_permute(s->st); s->fill = 0; }
    }
}
static inline void sp_finish(sponge* s, u32* out8) {
    if (s->fill) orc_permute(s->st);
    memcpy(out8, s->st, 32);
}
void orc_hash_row(const u32* row, u64 n, u32* out8) { sponge s; sp_init(&s); sp_absorb(&s, row, n); sp_finish(&s, out8); }
void orc_compress(const u32* l, const u32* r, u32* out8) {
    u32 st[16]; memcpy(st, l, 32); memcpy(st + 8, r, 32); orc_permute(st); memcpy(out8, st, 32);
}
void orc_hash_rows(const u32* mat, u64 rows, u64 width, u32* out) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < rows; i++) orc_hash_row(mat + i * width, width, out + 8 * i);
}
void orc_compress_pairs(const u32* in, u32* out, u64 n) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) orc_compress(in + 16 * i, in + 16 * i + 8, out + 8 * i);
}

/* ------------------------------------------------------------------ MerkleTreeMmcs (power-of-two heights)
 * mats[k] row-major, heights[k] x widths[k]; any order (stable sort by height desc inside).
 * digests_out: all layers concatenated, layer0 (max_h digests) first ... root last: 8*(2*max_h-1) u32.
 * returns 0 ok, -1 bad shape. */
int orc_merkle_commit(const u32* const* mats, const u64* heights, const u64* widths, u32 k, u32* digests_out, u32* root_out) {
    if (!k) return -1;
    u32* order = malloc(sizeof(u32) * k);
    for (u32 i = 0; i < k; i++) { order[i] = i; if (!heights[i] || (heights[i] & (heights[i] - 1))) { free(order); return -1; } }
    for (u32 i = 1; i < k; i++) { /* stable insertion sort, tallest first */
        u32 v = order[i]; int j = (int)i - 1;
        while (j >= 0 && heights[order[j]] < heights[v]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    u64 max_h = heights[order[0]];
    u32 pos = 0, g0 = 0;
    while (pos < k && heights[order[pos]] == max_h) pos++;
    u32* layer = digests_out;
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < max_h; i++) {
        sponge s; sp_init(&s);
        for (u32 m = g0; m < pos; m++) sp_absorb(&s, mats[order[m]] + i * widths[order[m]], widths[order[m]]);
        sp_finish(&s, layer + 8 * i);
    }
    u64 len = max_h;
    while (len > 1) {
        u64 h = len / 2;
        u32* next = layer + 8 * len;
        g0 = pos;
        while (pos < k && heights[order[pos]] == h) pos++;
#pragma omp parallel for schedule(static)
        for (u64 i = 0; i < h; i++) {
            u32 n[8];
            orc_compress(layer + 16 * i, layer + 16 * i + 8, n);
            if (pos > g0) {
                u32 hrow[8];
                sponge s; sp_init(&s);
                for (u32 m = g0; m < pos; m++) sp_absorb(&s, mats[order[m]] + i * widths[order[m]], widths[order[m]]);
                sp_finish(&s, hrow);
                orc_compress(n, hrow, n);
            }
            memcpy(next + 8 * i, n, 32);
        }
        layer = next; len = h;
    }
    memcpy(root_out, layer, 32);
    int ok = pos == k ? 0 : -1;
    free(order);
    return ok;
}

/* MerkleTreeMmcs::verify_batch.  rows[k] are the opened rows (widths[k] elems), heights as committed.
 * path: depth x 8.  returns 1 if the recomputed root equals `root`. */
int orc_merkle_verify(const u32* const* rows, const u64* heights, const u64* widths, u32 k, const u32* path, u32 depth, u64 index, const u32* root) {
    if (!k) return 0;
    u32* order = malloc(sizeof(u32) * k);
    for (u32 i = 0; i < k; i++) order[i] = i;
    for (u32 i = 1; i < k; i++) {
        u32 v = order[i]; int j = (int)i - 1;
        while (j >= 0 && heights[order[j]] < heights[v]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    u64 h = heights[order[0]];
    u32 pos = 0, node[8];
    sponge s; sp_init(&s);
    while (pos < k && heights[order[pos]] == h) { sp_absorb(&s, rows[order[pos]], widths[order[pos]]); pos++; }
    sp_finish(&s, node);
    for (u32 d = 0; d < depth; d++) {
        const u32* sib = path + 8 * d;
        if (index & 1) orc_compress(sib, node, node); else orc_compress(node, sib, node);
        index >>= 1; h >>= 1;
        if (pos < k && heights[order[pos]] == h) {
            u32 hr[8]; sp_init(&s);
            while (pos < k && heights[order[pos]] == h) { sp_absorb(&s, rows[order[pos]], widths[order[pos]]); pos++; }
            sp_finish(&s, hr);
            orc_compress(node, hr, node);
        }
    }
    free(order);
    return pos == k && h == 1 && memcmp(node, root, 32) == 0;
}

/* ------------------------------------------------------------------ DFT
 * NaiveDft (definition): out[i] = sum_j a[j] w^(ij); natural order in and out; per column. */
void orc_naive_dft(const u32* in, u32* out, u64 n, u64 width, int inverse) {
    orc_init();
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32 w = orc_two_adic_generator(lg);
    if (inverse) w = bb_inv(w);
    u32 ninv = bb_inv(bb_to_monty((u32)(n % P)));
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) {
        u32 wi = bb_pow(w, i);
        for (u64 c = 0; c < width; c++) {
            u32 acc = 0, x = R_MOD_P;
            for (u64 j = 0; j < n; j++) { acc = bb_add(acc, bb_mul(in[j * width + c], x)); x = bb_mul(x, wi); }
            out[i * width + c] = inverse ? bb_mul(acc, ninv) : acc;
        }
    }
}

/* in-place radix-2 decimation-in-frequency over rows of a row-major matrix:
 * natural-order input -> bit-reversed-order output (physical row j = DFT value bitrev(j)). */
static void dif_rows(u32* a, u64 n, u64 width, u32 root) {
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32* tw = malloc(sizeof(u32) * (n / 2 + 1));
    tw[0] = R_MOD_P;
    for (u64 i = 1; i < n / 2; i++) tw[i] = bb_mul(tw[i - 1], root);
    for (u32 s = 0; s < lg; s++) {
        u64 m = n >> (s + 1); /* half block */
        u64 nblk = 1ull << s;
#pragma omp parallel for collapse(2) schedule(static)
        for (u64 b = 0; b < nblk; b++)
            for (u64 j = 0; j < m; j++) {
                u32 w = tw[j << s];
                u32* x = a + (b * 2 * m + j) * width;
                u32* y = x + m * width;
                for (u64 c = 0; c < width; c++) {
                    u32 u = x[c], v = y[c];
                    x[c] = bb_add(u, v);
                    y[c] = bb_mul(bb_sub(u, v), w);
                }
            }
    }
    free(tw);
}
static void bitrev_rows(u32* a, u64 n, u64 width) {
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32* tmp = malloc(4 * width);
    for (u64 i = 0; i < n; i++) {
        u64 j = bitrev32((u32)i, lg);
        if (i < j) { memcpy(tmp, a + i * width, 4 * width); memcpy(a + i * width, a + j * width, 4 * width); memcpy(a + j * width, tmp, 4 * width); }
    }
    free(tmp);
}

/* dft_batch / coset_dft_batch / idft: `a` (n x width) is transformed in place.
 * shift: Montgomery coset shift (monty(1) for none).  bitrev_out: 1 keeps the DIF (bit-reversed) order. */
void orc_dft_batch(u32* a, u64 n, u64 width, u32 shift, int inverse, int bitrev_out) {
    orc_init();
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32 w = orc_two_adic_generator(lg);
    if (!inverse) {
        if (shift != R_MOD_P) {
#pragma omp parallel for schedule(static)
            for (u64 i = 0; i < n; i++) { u32 sp = bb_pow(shift, i); for (u64 c = 0; c < width; c++) a[i * width + c] = bb_mul(a[i * width + c], sp); }
        }
        dif_rows(a, n, width, w);
        if (!bitrev_out) bitrev_rows(a, n, width);
    } else {
        /* coset_idft: evaluations on shift*H (natural order) -> coefficients (natural order) */
        dif_rows(a, n, width, bb_inv(w));
        bitrev_rows(a, n, width);
        u32 ninv = bb_inv(bb_to_monty((u32)(n % P))), sinv = bb_inv(shift);
#pragma omp parallel for schedule(static)
        for (u64 i = 0; i < n; i++) { u32 sp = bb_mul(ninv, bb_pow(sinv, i)); for (u64 c = 0; c < width; c++) a[i * width + c] = bb_mul(a[i * width + c], sp); }
        if (bitrev_out) bitrev_rows(a, n, width);
    }
}

/* TwoAdicSubgroupDft::coset_lde_batch(evals, added_bits, shift) followed by
 * .bit_reverse_rows().to_row_major_matrix() (p3-fri TwoAdicFriPcs::commit): out is (n<<added_bits) x width,
 * physical row j = evaluation at shift * w'^bitrev(j).  bitrev_out=0 gives the logical natural order. */
void orc_coset_lde_batch(const u32* evals, u64 n, u64 width, u32 added_bits, u32 shift, int bitrev_out, u32* out) {
    orc_init();
    u64 m = n << added_bits;
    memcpy(out, evals, 4 * n * width);
    orc_dft_batch(out, n, width, R_MOD_P, 1, 0);          /* idft -> coefficients */
    memset(out + n * width, 0, 4 * (m - n) * width);      /* zero-pad */
    orc_dft_batch(out, m, width, shift, 0, bitrev_out);   /* coset dft */
}

/* ------------------------------------------------------------------ FRI */
/* fold_matrix: in = len EF4 (bit-reversed domain order), out = len/2 EF4.
 * out[i] = (1/2 + beta/2 * g^-bitrev(i)) * lo + (1/2 - beta/2 * g^-bitrev(i)) * hi,  g = two_adic_generator(log2 len) */
void orc_fri_fold(const u32* in, u64 len, const u32* beta, u32* out) {
    orc_init();
    u64 h = len / 2;
    u32 lh = 0; while ((1ull << lh) < h) lh++;
    u32 ginv = bb_inv(orc_two_adic_generator(lh + 1));
    u32 half = bb_inv(bb_add(R_MOD_P, R_MOD_P));
    ef4 b; memcpy(b.c, beta, 16);
    ef4 hb = ef_scale(b, half);
    ef4 halfe = {{half, 0, 0, 0}};
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < h; i++) {
        ef4 pw = ef_scale(hb, bb_pow(ginv, bitrev32((u32)i, lh)));
        ef4 lo, hi; memcpy(lo.c, in + 8 * i, 16); memcpy(hi.c, in + 8 * i + 4, 16);
        ef4 r = ef_add(ef_mul(ef_add(halfe, pw), lo), ef_mul(ef_sub(halfe, pw), hi));
        memcpy(out + 4 * i, r.c, 16);
    }
}

/* ------------------------------------------------------------------ PCS open phase primitives
 * p3_matrix::Matrix::dot_ext_powers(alpha): out[r] = sum_c alpha^c * mat[r][c]   (EF4) */
static ef4 ef_from_base(u32 x) { ef4 r = {{x, 0, 0, 0}}; return r; }
static ef4 ef_inv(ef4 a) {
    /* a^-1 = a^(p^4-2); p^4-2 does not fit 64 bits: use the Frobenius-free route a^(p^4-2) = a^(p-2) * (a^(p(p^3-1)/(p-1))...)
       -- simpler and obviously right: square-and-multiply over the 124-bit exponent held in two words */
    unsigned __int128 e = (unsigned __int128)P * P;
    e = e * P * P - 2; /* p^4 - 2 < 2^124 */
    ef4 r = ef_from_base(R_MOD_P), b = a;
    while (e) { if (e & 1) r = ef_mul(r, b); b = ef_mul(b, b); e >>= 1; }
    return r;
}
void orc_ef_inv(const u32* a, u32* out) { orc_init(); ef4 x; memcpy(x.c, a, 16); ef4 r = ef_inv(x); memcpy(out, r.c, 16); }
void orc_dot_ext_powers(const u32* mat, u64 rows, u64 width, const u32* alpha, u32* out) {
    orc_init();
    ef4 a; memcpy(a.c, alpha, 16);
    ef4* pw = malloc(sizeof(ef4) * (width ? width : 1));
    pw[0] = ef_from_base(R_MOD_P);
    for (u64 c = 1; c < width; c++) pw[c] = ef_mul(pw[c - 1], a);
#pragma omp parallel for schedule(static)
    for (u64 r = 0; r < rows; r++) {
        ef4 acc = {{0, 0, 0, 0}};
        for (u64 c = 0; c < width; c++) acc = ef_add(acc, ef_scale(pw[c], mat[r * width + c]));
        memcpy(out + 4 * r, acc.c, 16);
    }
    free(pw);
}
/* p3_interpolation::interpolate_coset(evals on shift*H, shift, point): value of every column's interpolant at the EF4
 * point z.  `evals` holds the N evaluations in BIT-REVERSED row order (the low coset of a committed LDE).  Definition
 * level: coefficients by inverse DFT, then Horner at z. */
void orc_interpolate_coset_bitrev(const u32* evals, u64 n, u64 width, u32 shift, const u32* point, u32* out) {
    orc_init();
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32* nat = malloc(4 * n * width);
    for (u64 i = 0; i < n; i++) memcpy(nat + (u64)bitrev32((u32)i, lg) * width, evals + i * width, 4 * width);
    orc_dft_batch(nat, n, width, shift, 1, 0); /* coset idft -> coefficients */
    ef4 z; memcpy(z.c, point, 16);
    for (u64 c = 0; c < width; c++) {
        ef4 acc = {{0, 0, 0, 0}};
        for (u64 j = n; j-- > 0;) acc = ef_add(ef_mul(acc, z), ef_from_base(nat[j * width + c]));
        memcpy(out + 4 * c, acc.c, 16);
    }
    free(nat);
}
/* reduced openings of one (matrix, point) pair, p3-fri TwoAdicFriPcs::open:
 *   ro[i] += alpha_pow_offset * (reduced_ys - reduced_row[i]) / (z - x_i),   x_i = shift * w_M^bitrev(i)
 * reduced_row = dot_ext_powers(lde, alpha), reduced_ys = sum_c alpha^c * p_c(z). */
void orc_reduce_openings(const u32* reduced_row, u64 m, u32 shift, const u32* point, const u32* reduced_ys, const u32* alpha_pow_offset, u32* ro) {
    orc_init();
    u32 lg = 0; while ((1ull << lg) < m) lg++;
    u32 w = orc_two_adic_generator(lg);
    ef4 z, ys, apo; memcpy(z.c, point, 16); memcpy(ys.c, reduced_ys, 16); memcpy(apo.c, alpha_pow_offset, 16);
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < m; i++) {
        u32 x = bb_mul(shift, bb_pow(w, bitrev32((u32)i, lg)));
        ef4 den = z; den.c[0] = bb_sub(den.c[0], x);
        ef4 rr, cur; memcpy(rr.c, reduced_row + 4 * i, 16); memcpy(cur.c, ro + 4 * i, 16);
        ef4 t = ef_mul(ef_mul(apo, ef_sub(ys, rr)), ef_inv(den));
        cur = ef_add(cur, t);
        memcpy(ro + 4 * i, cur.c, 16);
    }
}

/* DuplexChallenger<BabyBear, Poseidon2, 16, 8> */
typedef struct { u32 st[16]; u32 in[8]; u32 nin; u32 out[8]; u32 nout; } orc_chal;
void orc_chal_init(orc_chal* c) { memset(c, 0, sizeof *c); }
static void chal_duplex(orc_chal* c) {
    for (u32 i = 0; i < c->nin; i++) c->st[i] = c->in[i];
    c->nin = 0;
    orc_permute(c->st);
    memcpy(c->out, c->st, 32); c->nout = 8;
}
void orc_chal_observe(orc_chal* c, const u32* v, u64 n) {
    for (u64 i = 0; i < n; i++) { c->nout = 0; c->in[c->nin++] = v[i]; if (c->nin == 8) chal_duplex(c); }
}
u32 orc_chal_sample(orc_chal* c) {
    if (c->nin || !c->nout) chal_duplex(c);
    return c->out[--c->nout];
}
void orc_chal_sample_ext(orc_chal* c, u32* out4) { for (int i = 0; i < 4; i++) out4[i] = orc_chal_sample(c); }
u32 orc_chal_sample_bits(orc_chal* c, u32 bits) { return bb_from_monty(orc_chal_sample(c)) & ((1u << bits) - 1); }
/* smallest PoW witness (canonical value returned); observes it like DuplexChallenger::grind */
u32 orc_chal_grind(orc_chal* c, u32 bits) {
    if (bits == 0) return 0; /* p3-challenger 0.4.3 GrindingChallenger: bits == 0 needs no witness and leaves the transcript alone */
    for (u32 w = 0; w < P; w++) {
        orc_chal t = *c;
        u32 wm = bb_to_monty(w);
        orc_chal_observe(&t, &wm, 1);
        if (orc_chal_sample_bits(&t, bits) == 0) { *c = t; return w; }
    }
    return 0xffffffffu;
}

/* p3-fri prover::commit_phase for one input vector (plus optional roll-ins): see oracle/pyref.py
 * fri_commit_phase.  in0: len0 EF4; betas forced if `betas_in` != NULL (else drawn from challenger).
 * roots_out: rounds x 8; final_out: (blowup*final_poly_len) EF4 = the last folded vector (bit-reversed). */
u32 orc_fri_commit_phase(const u32* in0, u64 len0, u32 log_blowup, u32 log_final_poly_len, const u32* betas_in,
                         orc_chal* chal, u32* roots_out, u32* betas_out, u32* final_out) {
    u64 len = len0, stop = 1ull << (log_blowup + log_final_poly_len);
    u32* cur = malloc(16 * len); memcpy(cur, in0, 16 * len);
    u32* nxt = malloc(16 * len / 2 + 16);
    u32* dig = malloc(32 * (len + 1));
    u32 rounds = 0;
    while (len > stop) {
        u64 rows = len / 2, w = 8;
        const u32* mp = cur;
        orc_merkle_commit(&mp, &rows, &w, 1, dig, roots_out + 8 * rounds);
        u32 beta[4];
        if (betas_in) memcpy(beta, betas_in + 4 * rounds, 16);
        else { orc_chal_observe(chal, roots_out + 8 * rounds, 8); orc_chal_sample_ext(chal, beta); }
        if (betas_out) memcpy(betas_out + 4 * rounds, beta, 16);
        orc_fri_fold(cur, len, beta, nxt);
        u32* t = cur; cur = nxt; nxt = t;
        len /= 2; rounds++;
    }
    memcpy(final_out, cur, 16 * len);
    free(cur); free(nxt); free(dig);
    return rounds;
}

/* deterministic synthetic data shared by CPU and GPU sides: element i = splitmix64(seed ^ i) mod p,
 * stored as that value interpreted as a Montgomery u32 (any value < p is a valid element). */
static inline u64 splitmix64(u64 x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
void orc_fill(u32* out, u64 n, u64 seed, u64 offset) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) out[i] = (u32)(splitmix64(seed ^ (offset + i)) % P);
}
/* order-independent checksum of a u32 buffer: sum over i of splitmix64(i ^ v[i] << 32) mod 2^64 */
u64 orc_checksum(const u32* v, u64 n, u64 offset) {
    u64 acc = 0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (u64 i = 0; i < n; i++) acc += splitmix64((offset + i) ^ ((u64)v[i] << 32));
    return acc;
}


/* ================================================================== AVX-512 fast path (bench.py CPU baseline only)
 * The same algorithms with the data-parallel structure Plonky3's CPU code uses, so that the timed CPU baseline is not a scalar
 * strawman: PackedBabyBearAVX512-style Montgomery arithmetic on 16 lanes (p3-monty-31 x86_64_avx512 packing), Poseidon2 over 16
 * rows at a time (p3-merkle-tree hashes PackedValue::WIDTH rows per permutation), butterflies vectorised along the row
 * (p3-dft Radix2DitParallel works on row-major matrices the same way) with two stages per sweep and cache-resident tails.
 * Compiled only with -march=native on an AVX-512 host (`make native`); tests/test_oracle_kats.py checks it against the scalar
 * restatement above, which stays the parity checker. */
#if defined(__AVX512F__) && defined(__AVX512DQ__)
#include <immintrin.h>
#define ORC_AVX512 1
typedef __m512i v16;
#define VP _mm512_set1_epi32((int)P)
#define VMU _mm512_set1_epi32((int)MU)
static inline v16 v_add(v16 a, v16 b) { v16 s = _mm512_add_epi32(a, b); return _mm512_min_epu32(s, _mm512_sub_epi32(s, VP)); }
static inline v16 v_sub(v16 a, v16 b) { v16 d = _mm512_sub_epi32(a, b); return _mm512_min_epu32(d, _mm512_add_epi32(d, VP)); }
static inline v16 v_mul(v16 a, v16 b) {
    v16 ao = _mm512_srli_epi64(a, 32), bo = _mm512_srli_epi64(b, 32);
    v16 pe = _mm512_mul_epu32(a, b), po = _mm512_mul_epu32(ao, bo);
    v16 qe = _mm512_mul_epu32(pe, VMU), qo = _mm512_mul_epu32(po, VMU);
    v16 qpe = _mm512_mul_epu32(qe, VP), qpo = _mm512_mul_epu32(qo, VP);
    v16 hi = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(pe, 32), po);
    v16 qh = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(qpe, 32), qpo);
    v16 t = _mm512_sub_epi32(hi, qh);
    return _mm512_min_epu32(t, _mm512_add_epi32(t, VP));
}
static inline v16 v_sbox7(v16 x) { v16 x2 = v_mul(x, x), x3 = v_mul(x2, x), x4 = v_mul(x2, x2); return v_mul(x3, x4); }
static inline void v_mds_light(v16* s) {
    for (int c = 0; c < 16; c += 4) {
        v16 a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
        v16 t = v_add(v_add(a, b), v_add(cc, d));
        s[c] = v_add(v_add(t, a), v_add(b, b));
        s[c + 1] = v_add(v_add(t, b), v_add(cc, cc));
        s[c + 2] = v_add(v_add(t, cc), v_add(d, d));
        s[c + 3] = v_add(v_add(t, d), v_add(a, a));
    }
    for (int k = 0; k < 4; k++) {
        v16 t = v_add(v_add(s[k], s[4 + k]), v_add(s[8 + k], s[12 + k]));
        for (int j = 0; j < 16; j += 4) s[j + k] = v_add(s[j + k], t);
    }
}
/* 16 independent permutations, lane l of s[i] = element i of state l */
static void v_permute(v16* s) {
    v_mds_light(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = v_sbox7(v_add(s[i], _mm512_set1_epi32((int)RC[16 * r + i])));
        v_mds_light(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = v_sbox7(v_add(s[0], _mm512_set1_epi32((int)RC[64 + r])));
        v16 t = s[0];
        for (int i = 1; i < 16; i++) t = v_add(t, s[i]);
        for (int i = 0; i < 16; i++) s[i] = v_add(t, v_mul(_mm512_set1_epi32((int)DIAG[i]), s[i]));
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = v_sbox7(v_add(s[i], _mm512_set1_epi32((int)RC[77 + 16 * r + i])));
        v_mds_light(s);
    }
}
int orc_fast_available(void) { return 1; }

/* leaf digests of one matrix, 16 rows per permutation (rows % 16 == 0) */
static void fast_hash_rows(const u32* mat, u64 rows, u64 width, u32* out) {
    const v16 lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
#pragma omp parallel for schedule(static)
    for (u64 r0 = 0; r0 < rows; r0 += 16) {
        v16 st[16];
        for (int i = 0; i < 16; i++) st[i] = _mm512_setzero_si512();
        const v16 idx = _mm512_mullo_epi32(lane, _mm512_set1_epi32((int)width));
        const u32* base = mat + r0 * width;
        int fill = 0;
        for (u64 c = 0; c < width; c++) {
            st[fill++] = _mm512_i32gather_epi32(idx, (const int*)(base + c), 4);
            if (fill == 8) { v_permute(st); fill = 0; }
        }
        if (fill) v_permute(st);
        const v16 oidx = _mm512_slli_epi32(lane, 3);
        for (int j = 0; j < 8; j++) _mm512_i32scatter_epi32((int*)(out + 8 * r0 + j), oidx, st[j], 4);
    }
}
/* next[i] = compress(prev[2i], prev[2i+1]), 16 nodes per permutation (n % 16 == 0) */
static void fast_compress_layer(const u32* prev, u32* next, u64 n) {
    const v16 lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const v16 iidx = _mm512_slli_epi32(lane, 4), oidx = _mm512_slli_epi32(lane, 3);
#pragma omp parallel for schedule(static)
    for (u64 i0 = 0; i0 < n; i0 += 16) {
        v16 st[16];
        for (int j = 0; j < 16; j++) st[j] = _mm512_i32gather_epi32(iidx, (const int*)(prev + 16 * i0 + j), 4);
        v_permute(st);
        for (int j = 0; j < 8; j++) _mm512_i32scatter_epi32((int*)(next + 8 * i0 + j), oidx, st[j], 4);
    }
}
/* MerkleTreeMmcs commit of ONE matrix (the bench shape); digests_out as orc_merkle_commit */
int orc_fast_merkle_commit_single(const u32* mat, u64 rows, u64 width, u32* digests_out, u32* root_out) {
    orc_init();
    if (!rows || (rows & (rows - 1))) return -1;
    u32* layer = digests_out;
    if (rows >= 16) fast_hash_rows(mat, rows, width, layer);
    else orc_hash_rows(mat, rows, width, layer);
    for (u64 len = rows; len > 1; len >>= 1) {
        u32* next = layer + 8 * len;
        if (len / 2 >= 16) fast_compress_layer(layer, next, len / 2);
        else orc_compress_pairs(layer, next, len / 2);
        layer = next;
    }
    memcpy(root_out, layer, 32);
    return 0;
}

/* ---- vectorised DIF over rows (natural in, bit-reversed out), width % 16 == 0 */
static inline void v_bfly(u32* x, u32* y, u64 width, u32 w) {
    const v16 vw = _mm512_set1_epi32((int)w);
    for (u64 c = 0; c < width; c += 16) {
        v16 u = _mm512_loadu_si512(x + c), v = _mm512_loadu_si512(y + c);
        _mm512_storeu_si512(x + c, v_add(u, v));
        _mm512_storeu_si512(y + c, v_mul(v_sub(u, v), vw));
    }
}
static inline void v_bfly_notw(u32* x, u32* y, u64 width) {
    for (u64 c = 0; c < width; c += 16) {
        v16 u = _mm512_loadu_si512(x + c), v = _mm512_loadu_si512(y + c);
        _mm512_storeu_si512(x + c, v_add(u, v));
        _mm512_storeu_si512(y + c, v_sub(u, v));
    }
}
/* stages s and s+1 in one sweep: rows j, j+q, j+2q, j+3q of a block of 4q rows */
static inline void v_bfly4(u32* r0, u32* r1, u32* r2, u32* r3, u64 width, u32 wa, u32 wb, u32 wc) {
    const v16 va = _mm512_set1_epi32((int)wa), vb = _mm512_set1_epi32((int)wb), vc = _mm512_set1_epi32((int)wc);
    for (u64 c = 0; c < width; c += 16) {
        v16 x0 = _mm512_loadu_si512(r0 + c), x1 = _mm512_loadu_si512(r1 + c), x2 = _mm512_loadu_si512(r2 + c), x3 = _mm512_loadu_si512(r3 + c);
        v16 a0 = v_add(x0, x2), a2 = v_mul(v_sub(x0, x2), va);   /* stage s: pairs (0,2), (1,3) */
        v16 a1 = v_add(x1, x3), a3 = v_mul(v_sub(x1, x3), vb);
        _mm512_storeu_si512(r0 + c, v_add(a0, a1));               /* stage s+1: pairs (0,1), (2,3), same twiddle */
        _mm512_storeu_si512(r1 + c, v_mul(v_sub(a0, a1), vc));
        _mm512_storeu_si512(r2 + c, v_add(a2, a3));
        _mm512_storeu_si512(r3 + c, v_mul(v_sub(a2, a3), vc));
    }
}
static void fast_dif_rows(u32* a, u64 n, u64 width, u32 root) {
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32* tw = malloc(sizeof(u32) * (n / 2 + 1));
    tw[0] = R_MOD_P;
    for (u64 i = 1; i < n / 2; i++) tw[i] = bb_mul(tw[i - 1], root);
    /* rows of a cache-resident block: ~512 KB */
    u32 blk_lg = 1; while (blk_lg < lg && ((4ull * width) << (blk_lg + 1)) <= (512u << 10)) blk_lg++;
    if (blk_lg > lg) blk_lg = lg;
    u32 s = 0;
    const u32 global = lg - blk_lg;
    for (; s + 1 < global || (s < global && ((global - s) & 1) == 0 && s + 1 < lg); s += 2) { /* two global stages per sweep */
        if (s + 1 >= lg) break;
        const u64 q = n >> (s + 2), nblk = 1ull << s;
#pragma omp parallel for collapse(2) schedule(static)
        for (u64 b = 0; b < nblk; b++)
            for (u64 j = 0; j < q; j++) {
                u32* r0 = a + (b * 4 * q + j) * width;
                v_bfly4(r0, r0 + q * width, r0 + 2 * q * width, r0 + 3 * q * width, width, tw[j << s], tw[(j + q) << s], tw[j << (s + 1)]);
            }
    }
    for (; s < global; s++) { /* an odd global stage left over */
        const u64 m = n >> (s + 1), nblk = 1ull << s;
#pragma omp parallel for collapse(2) schedule(static)
        for (u64 b = 0; b < nblk; b++)
            for (u64 j = 0; j < m; j++) { u32* x = a + (b * 2 * m + j) * width; v_bfly(x, x + m * width, width, tw[j << s]); }
    }
    /* the remaining stages stay inside blocks of n >> s rows: one sweep, block by block */
    if (s < lg) {
        const u64 brows = n >> s, nb = 1ull << s;
        const u32 s_first = s;
#pragma omp parallel for schedule(static)
        for (u64 b = 0; b < nb; b++) {
            u32* base = a + b * brows * width;
            for (u32 t = s_first; t < lg; t++) {
                const u64 m = n >> (t + 1), inner = brows / (2 * m);
                for (u64 ib = 0; ib < inner; ib++)
                    for (u64 j = 0; j < m; j++) {
                        u32* x = base + (ib * 2 * m + j) * width;
                        if (t + 1 == lg) v_bfly_notw(x, x + m * width, width);
                        else v_bfly(x, x + m * width, width, tw[j << t]);
                    }
            }
        }
    }
    free(tw);
}
static inline void v_scale_row(const u32* src, u32* dst, u64 width, u32 f) {
    const v16 vf = _mm512_set1_epi32((int)f);
    for (u64 c = 0; c < width; c += 16) _mm512_storeu_si512(dst + c, v_mul(_mm512_loadu_si512(src + c), vf));
}
/* coset_lde_batch with bit-reversed output (width % 16 == 0, n >= 2): iDFT once, then per coset block c of the output the
 * size-n DFT of coef[k] * (shift * w'^bitrev(c))^k / n -- the zero-padded stages of the size-(n << added_bits) transform
 * are never executed (bit-identical result; tests compare with orc_coset_lde_batch) */
int orc_fast_coset_lde_batch(const u32* evals, u64 n, u64 width, u32 added_bits, u32 shift, u32* out) {
    orc_init();
    if (width % 16 || n < 2 || (n & (n - 1))) return -1;
    u32 lg = 0; while ((1ull << lg) < n) lg++;
    u32* t = malloc(4 * n * width);
    if (!t) return -2;
    memcpy(t, evals, 4 * n * width);
    fast_dif_rows(t, n, width, bb_inv(orc_two_adic_generator(lg)));   /* bit-reversed coefficients, unscaled */
    const u32 ninv = bb_inv(bb_to_monty((u32)(n % P)));
    const u32 wprime = orc_two_adic_generator(lg + added_bits);
    for (u64 c = 0; c < (1ull << added_bits); c++) {
        u32* blk = out + c * n * width;
        const u32 g = bb_mul(shift, bb_pow(wprime, bitrev32((u32)c, added_bits)));
#pragma omp parallel
        {
            /* each thread walks a contiguous range of k with a running power */
            int nt = 1, id = 0;
#ifdef _OPENMP
            nt = omp_get_num_threads(); id = omp_get_thread_num();
#endif
            u64 k0 = n * (u64)id / nt, k1 = n * (u64)(id + 1) / nt;
            u32 f = bb_mul(ninv, bb_pow(g, k0));
            for (u64 k = k0; k < k1; k++) { v_scale_row(t + (u64)bitrev32((u32)k, lg) * width, blk + k * width, width, f); f = bb_mul(f, g); }
        }
        fast_dif_rows(blk, n, width, orc_two_adic_generator(lg));
    }
    free(t);
    return 0;
}
#else
int orc_fast_available(void) { return 0; }
int orc_fast_merkle_commit_single(const u32* mat, u64 rows, u64 width, u32* digests_out, u32* root_out) {
    const u32* m[1] = {mat};
    return orc_merkle_commit(m, &rows, &width, 1, digests_out, root_out);
}
int orc_fast_coset_lde_batch(const u32* evals, u64 n, u64 width, u32 added_bits, u32 shift, u32* out) {
    orc_coset_lde_batch(evals, n, width, added_bits, shift, 1, out);
    return 0;
}
#endif

/* column extraction of the synthetic matrix without materialising it: out[i] = fill value at index offset + i * stride */
void orc_fill_strided(u32* out, u64 n, u64 seed, u64 offset, u64 stride) {
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < n; i++) out[i] = (u32)(splitmix64(seed ^ (offset + i * stride)) % P);
}
/* Horner evaluation of one polynomial (n Montgomery coefficients, natural order) at cnt points: the definition-level
 * check of an LDE entry, independent of any fast transform */
void orc_eval_poly_many(const u32* coef, u64 n, const u32* xs, u64 cnt, u32* out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (u64 k = 0; k < cnt; k++) {
        u32 acc = 0, x = xs[k];
        for (u64 j = n; j-- > 0;) acc = bb_add(bb_mul(acc, x), coef[j]);
        out[k] = acc;
    }
}
