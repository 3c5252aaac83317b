"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement (canonical, non-Montgomery integers).

Slow (~100 us / permutation): used for known-answer tests, for mining the
reference's proof fixtures (oracle/mine_fixture.py) and to cross-check the C
oracle (oracle/bb_oracle.c) at small sizes.  Nothing under zkvm_prover_b200/
may import this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may.

The algorithms live in third-party crates that are NOT vendored under
/root/reference (Cargo.lock pins them):
  p3-baby-bear / p3-monty-31 / p3-field 0.4.3   (Cargo.lock:5545,5685,5605)
  p3-dft 0.4.3                                   (Cargo.lock:5590)
  p3-poseidon2 0.4.3 + zkhash-axiom 0.2.0        (Cargo.lock:5708,10231)
  p3-symmetric 0.4.3                             (Cargo.lock:5736)
  p3-challenger 0.4.3                            (Cargo.lock:5576)
  p3-merkle-tree / p3-commit / p3-fri            (v1-era; produced the fixtures)
Each function restates the published algorithm; parity is anchored on the
reference's own proof fixtures (see oracle/mine_fixture.py).
Reference call sites: crates/prover/src/prover/mod.rs:355-357 (sdk.prove),
crates/types/src/proof.rs:70-74 (legacy VmInternalStarkProof wire format).
"""
from __future__ import annotations

P = 2013265921  # 2^31 - 2^27 + 1   (also /root/reference/scripts/compress_bn254.py:10)
R = 1 << 32
RINV = pow(R, -1, P)
GEN = 31  # multiplicative generator of BabyBear
TWO_ADICITY = 27
W_EXT = 11  # EF4 = F[x]/(x^4 - 11)


def to_monty(x: int) -> int:
    return (x * R) % P


def from_monty(m: int) -> int:
    return (m * RINV) % P


def inv(a: int) -> int:
    return pow(a, -1, P)


def two_adic_generator(bits: int) -> int:
    """p3-field TwoAdicField::two_adic_generator for BabyBear: 31^15 is the 2^27-th root."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(pow(GEN, 15, P), 1 << (TWO_ADICITY - bits), P)


def bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


# --------------------------------------------------------------------------- Poseidon2
def grain_rc16():
    """Horizen-Labs Poseidon2 BabyBear t=16 round constants via the Poseidon Grain LFSR
    (n=31, t=16, R_F=8, R_P=13); identical to zkhash's RC16 consumed by OpenVM default_perm()."""
    bits = lambda v, w: [int(c) for c in bin(v)[2:].zfill(w)]
    s = bits(1, 2) + bits(0, 4) + bits(31, 12) + bits(16, 12) + bits(8, 10) + bits(13, 10) + [1] * 30

    def step():
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    for _ in range(160):
        step()

    def bit():
        while True:
            if step():
                return step()
            step()

    out = []
    while len(out) < 8 * 16 + 13:
        v = 0
        for _ in range(31):
            v = (v << 1) | bit()
        if v < P:
            out.append(v)
    ini = [out[16 * i:16 * i + 16] for i in range(4)]
    mid = out[64:77]
    fin = [out[77 + 16 * i:93 + 16 * i] for i in range(4)]
    return ini, mid, fin


RC_INIT, RC_INT, RC_TERM = grain_rc16()

# internal diagonal (p3-baby-bear poseidon2.rs, width 16): state[i] = sum + V[i]*state[i]
DIAG_V = [x % P for x in [
    -2, 1, 2, inv(2), 3, 4, -inv(2), -3, -4, inv(1 << 8), inv(4), inv(8), inv(1 << 27),
    -inv(1 << 8), -inv(16), -inv(1 << 27)]]
M4 = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]


def mds_light(s):
    o = []
    for c in range(4):
        o += [sum(M4[r][k] * s[4 * c + k] for k in range(4)) % P for r in range(4)]
    t = [sum(o[4 * j + k] for j in range(4)) % P for k in range(4)]
    return [(o[i] + t[i % 4]) % P for i in range(16)]


def permute(s, rc=None):
    """Poseidon2 width 16, x^7, 4+13+4 rounds (p3-poseidon2 Poseidon2::permute_mut)."""
    ini, mid, fin = rc if rc is not None else (RC_INIT, RC_INT, RC_TERM)
    s = mds_light([x % P for x in s])
    for r in ini:
        s = mds_light([pow((a + b) % P, 7, P) for a, b in zip(s, r)])
    for c in mid:
        s[0] = pow((s[0] + c) % P, 7, P)
        t = sum(s) % P
        s = [(t + DIAG_V[i] * s[i]) % P for i in range(16)]
    for r in fin:
        s = mds_light([pow((a + b) % P, 7, P) for a, b in zip(s, r)])
    return s


def hash_iter(row):
    """p3-symmetric PaddingFreeSponge<Perm,16,8,8>::hash_iter (overwrite mode, no padding)."""
    st = [0] * 16
    for i in range(0, len(row), 8):
        ch = row[i:i + 8]
        st[:len(ch)] = ch
        st = permute(st)
    return st[:8]


def compress(l, r):
    """p3-symmetric TruncatedPermutation<Perm,2,8,16>::compress."""
    return permute(list(l) + list(r))[:8]


# --------------------------------------------------------------------------- Merkle (MerkleTreeMmcs)
def next_pow2(h: int) -> int:
    return 1 << (h - 1).bit_length() if h > 1 else 1


def merkle_commit(mats):
    """p3-merkle-tree MerkleTree::new over matrices `mats` (list of list-of-rows), mixed heights.
    Returns (root, digest_layers)."""
    order = sorted(range(len(mats)), key=lambda i: -len(mats[i]))  # stable, tallest first
    pos = 0
    max_h = len(mats[order[0]])

    def take(h):
        nonlocal pos
        g = []
        while pos < len(order) and next_pow2(len(mats[order[pos]])) == h:
            g.append(mats[order[pos]])
            pos += 1
        return g

    # Heights are powers of two on this path (LDE domains); p3's even-length padding rule for
    # other heights is out of scope and rejected by the C ABI as B200ZK_ERR_SHAPE.
    assert all(len(m) == next_pow2(len(m)) for m in mats), "power-of-two heights only"
    tall = take(max_h)
    layer = [hash_iter([x for m in tall for x in m[i]]) for i in range(max_h)]
    layers = [layer]
    while len(layer) > 1:
        h = len(layer) // 2
        inj = take(h)
        nxt = []
        for i in range(h):
            n = compress(layer[2 * i], layer[2 * i + 1])
            if inj:
                n = compress(n, hash_iter([x for m in inj for x in m[i]]))
            nxt.append(n)
        layer = nxt
        layers.append(layer)
    assert pos == len(order)
    return layer[0], layers


def verify_batch(rows, heights, path, index, root):
    """p3-merkle-tree MerkleTreeMmcs::verify_batch (mixed heights)."""
    order = sorted(range(len(rows)), key=lambda i: -heights[i])
    pos = 0

    def take(h):
        nonlocal pos
        g = []
        hit = False
        while pos < len(order) and next_pow2(heights[order[pos]]) == h:
            g += rows[order[pos]]
            pos += 1
            hit = True
        return g, hit

    h = next_pow2(heights[order[0]])
    g, _ = take(h)
    node = hash_iter(g)
    for sib in path:
        node = compress(node, sib) if index & 1 == 0 else compress(sib, node)
        index >>= 1
        h >>= 1
        g, hit = take(h)
        if hit:
            node = compress(node, hash_iter(g))
    return node == list(root) and pos == len(order)


# --------------------------------------------------------------------------- EF4
def ef_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def ef_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def ef_mul(a, b):
    c = [0] * 7
    for i in range(4):
        for j in range(4):
            c[i + j] += a[i] * b[j]
    return [(c[i] + W_EXT * (c[i + 4] if i + 4 < 7 else 0)) % P for i in range(4)]


def ef_scale(a, k):
    return [(x * k) % P for x in a]


def ef_pow(a, e):
    r = [1, 0, 0, 0]
    while e:
        if e & 1:
            r = ef_mul(r, a)
        a = ef_mul(a, a)
        e >>= 1
    return r


def ef_inv(a):
    # a^(p^4-2)
    return ef_pow(a, P ** 4 - 2)


# --------------------------------------------------------------------------- DFT / LDE
def naive_dft(a):
    """p3-dft NaiveDft: out[i] = sum_j a[j] * w^(ij), natural order."""
    n = len(a)
    w = two_adic_generator(n.bit_length() - 1)
    return [sum(a[j] * pow(w, i * j, P) for j in range(n)) % P for i in range(n)]


def naive_idft(a):
    n = len(a)
    w = inv(two_adic_generator(n.bit_length() - 1))
    ninv = inv(n)
    return [sum(a[j] * pow(w, i * j, P) for j in range(n)) * ninv % P for i in range(n)]


def coset_lde_col(evals, added_bits, shift):
    """TwoAdicSubgroupDft::coset_lde_batch for one column; logical (natural) order out."""
    n = len(evals)
    c = naive_idft(evals) + [0] * (n * ((1 << added_bits) - 1))
    c = [x * pow(shift, j, P) % P for j, x in enumerate(c)]
    return naive_dft(c)


def coset_lde_batch_bitrev(mat, added_bits, shift):
    """Rows of the result are stored in bit-reversed order, i.e. what
    `dft.coset_lde_batch(..).bit_reverse_rows().to_row_major_matrix()` holds in p3-fri's
    TwoAdicFriPcs::commit: physical row j = evaluation at shift * w'^bitrev(j)."""
    n, w = len(mat), len(mat[0])
    cols = [coset_lde_col([mat[r][c] for r in range(n)], added_bits, shift) for c in range(w)]
    m = n << added_bits
    lb = m.bit_length() - 1
    return [[cols[c][bitrev(j, lb)] for c in range(w)] for j in range(m)]


# --------------------------------------------------------------------------- FRI
def fold_matrix(beta, folded):
    """p3-fri TwoAdicFriGenericConfig::fold_matrix: `folded` is a list of EF4 in bit-reversed
    domain order; rows of the (len/2) x 2 matrix are (lo, hi) = (f(x_i), f(-x_i)),
    x_i = g^bitrev(i), g = two_adic_generator(log2 len)."""
    n = len(folded)
    h = n // 2
    lh = h.bit_length() - 1
    g_inv = inv(two_adic_generator(lh + 1))
    half = inv(2)
    half_beta = ef_scale(beta, half)
    out = []
    for i in range(h):
        power = ef_scale(half_beta, pow(g_inv, bitrev(i, lh), P))
        lo, hi = folded[2 * i], folded[2 * i + 1]
        a = ef_mul(ef_add([half, 0, 0, 0], power), lo)
        b = ef_mul(ef_sub([half, 0, 0, 0], power), hi)
        out.append(ef_add(a, b))
    return out


class DuplexChallenger:
    """p3-challenger DuplexChallenger<BabyBear, Perm, 16, 8> (canonical ints)."""

    def __init__(self):
        self.state = [0] * 16
        self.inp = []
        self.out = []

    def clone(self):
        c = DuplexChallenger()
        c.state, c.inp, c.out = list(self.state), list(self.inp), list(self.out)
        return c

    def _duplex(self):
        assert len(self.inp) <= 8
        for i, v in enumerate(self.inp):
            self.state[i] = v
        self.inp = []
        self.state = permute(self.state)
        self.out = list(self.state[:8])

    def observe(self, v):
        self.out = []
        self.inp.append(v % P)
        if len(self.inp) == 8:
            self._duplex()

    def observe_slice(self, vs):
        for v in vs:
            self.observe(v)

    def sample(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def sample_ext(self):
        return [self.sample() for _ in range(4)]

    def sample_bits(self, bits):
        return self.sample() & ((1 << bits) - 1)

    def check_witness(self, bits, w):
        if bits == 0:   # p3-challenger 0.4.3: no proof of work requested, transcript untouched
            return True
        self.observe(w)
        return self.sample_bits(bits) == 0

    def grind(self, bits):
        """Smallest witness (the reference uses a parallel find_any, so its witness is not
        deterministic; the smallest one is always among the valid answers)."""
        if bits == 0:
            return 0
        w = 0
        while True:
            if self.clone().check_witness(bits, w):
                assert self.check_witness(bits, w)
                return w
            w += 1


def fri_commit_phase(inputs, challenger, blowup, final_poly_len, add_mode=0):
    """p3-fri prover::commit_phase.  inputs: list of EF4 vectors (strictly decreasing lengths),
    bit-reversed order.  Returns (commits, layers, final_poly, betas)."""
    inputs = list(inputs)
    folded = list(inputs.pop(0))
    commits, layer_data, betas = [], [], []
    while len(folded) > blowup * final_poly_len:
        leaves = [folded[2 * i] + folded[2 * i + 1] for i in range(len(folded) // 2)]  # width-8 rows
        root, layers = merkle_commit([leaves])
        challenger.observe_slice(root)
        beta = challenger.sample_ext()
        betas.append(beta)
        commits.append(root)
        layer_data.append((leaves, layers))
        folded = fold_matrix(beta, folded)
        if inputs and len(inputs[0]) == len(folded):
            v = inputs.pop(0)
            if add_mode == 0:
                folded = [ef_add(c, x) for c, x in zip(folded, v)]
            else:
                b2 = ef_mul(beta, beta)
                folded = [ef_add(c, ef_mul(b2, x)) for c, x in zip(folded, v)]
    n = len(folded)
    lb = n.bit_length() - 1
    nat = [folded[bitrev(i, lb)] for i in range(n)]
    # idft_algebra: coefficient-wise base-field iDFT
    coeffs = [naive_idft([nat[i][k] for i in range(n)]) for k in range(4)]
    final_poly = [[coeffs[k][i] for k in range(4)] for i in range(n)][:final_poly_len]
    for x in final_poly:
        challenger.observe_slice(x)
    return commits, layer_data, final_poly, betas
