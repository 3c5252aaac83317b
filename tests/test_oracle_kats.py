"""CPU: pins the oracle (oracle/bb_oracle.c + oracle/pyref.py) against
  * Plonky3's own Poseidon2 width-16 unit vector (p3-baby-bear test_poseidon2_width_16_random),
  * the vectors mined from the reference's proof fixture
    crates/verifier/testdata/proofs/chunk-proof-phase2.json (tests/golden/, oracle/mine_fixture.py),
  * the mathematical definitions (NaiveDft) where the reference has no test."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as R

P = R.P


def u32(x):
    return np.asarray(x, dtype=np.uint32)


# ----------------------------------------------------------------------------- field (a1)
def test_field_constants():
    assert pow(31, 15, P) == 0x1A427A41  # SURVEY B-1
    assert pow(0x1A427A41, 1 << 26, P) == P - 1
    for b in range(0, 28):
        assert O.from_monty(u32([O.two_adic_generator(b)]))[0] == R.two_adic_generator(b)
    table = [1, 0x78000000, 0x67055C21, 0x5EE99486, 0x0BB4C4E4]
    assert [R.two_adic_generator(b) for b in range(5)] == table


def test_field_ops_random():
    rng = np.random.default_rng(1)
    a = rng.integers(0, P, 2000, dtype=np.uint64)
    b = rng.integers(0, P, 2000, dtype=np.uint64)
    am, bm = O.to_monty(a), O.to_monty(b)
    L = O.lib()
    for x, y, xm, ym in zip(a[:500], b[:500], am[:500], bm[:500]):
        assert O.from_monty(u32([L.orc_mul(int(xm), int(ym))]))[0] == (int(x) * int(y)) % P
        assert O.from_monty(u32([L.orc_add(int(xm), int(ym))]))[0] == (int(x) + int(y)) % P
        assert O.from_monty(u32([L.orc_sub(int(xm), int(ym))]))[0] == (int(x) - int(y)) % P
    x = int(a[0]) or 1
    assert O.from_monty(u32([L.orc_inv(int(O.to_monty([x])[0]))]))[0] == pow(x, -1, P)
    assert np.array_equal(O.from_monty(am), a.astype(np.uint32))
    # edge values
    for x in (0, 1, P - 1):
        for y in (0, 1, P - 1):
            xm, ym = int(O.to_monty([x])[0]), int(O.to_monty([y])[0])
            assert O.from_monty(u32([L.orc_mul(xm, ym)]))[0] == x * y % P


def test_ef4_mul():
    rng = np.random.default_rng(2)
    for _ in range(50):
        a = rng.integers(0, P, 4).tolist()
        b = rng.integers(0, P, 4).tolist()
        out = np.zeros(4, np.uint32)
        O.lib().orc_ef_mul(O.to_monty(a), O.to_monty(b), out)
        assert O.from_monty(out).tolist() == R.ef_mul(a, b)
    a = [3, 1, 4, 1]
    assert R.ef_mul(a, R.ef_inv(a)) == [1, 0, 0, 0]


# ----------------------------------------------------------------------------- Poseidon2 (a3)
def _xoroshiro128plus(seed):
    """rand_xoshiro Xoroshiro128Plus::seed_from_u64 (SplitMix64 seeding), as used by Plonky3's test."""
    M = (1 << 64) - 1

    def splitmix():
        nonlocal seed
        seed = (seed + 0x9E3779B97F4A7C15) & M
        z = seed
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    s = [splitmix(), splitmix()]
    rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & M

    def next_u32():
        s0, s1 = s
        r = (s0 + s1) & M
        s1 ^= s0
        s[0] = rotl(s0, 24) ^ s1 ^ ((s1 << 16) & M)
        s[1] = rotl(s1, 37)
        return r >> 32

    return next_u32


def test_poseidon2_plonky3_unit_vector():
    """SURVEY B-2: pins M4 / diag V / round order / Montgomery sampling (random round constants)."""
    nxt = _xoroshiro128plus(1)

    def sample():  # p3 Distribution<MontyField31>: rejection-sample 31 bits as the MONTGOMERY repr
        while True:
            v = nxt() >> 1
            if v < P:
                return R.from_monty(v)

    ini = [[sample() for _ in range(16)] for _ in range(4)]
    fin = [[sample() for _ in range(16)] for _ in range(4)]
    mid = [sample() for _ in range(13)]
    inp = [894848333, 1437655012, 1200606629, 1690012884, 71131202, 1749206695, 1717947831, 120589055, 19776022,
           42382981, 1831865506, 724844064, 171220207, 1299207443, 227047920, 1783754913]
    exp = [1255099308, 941729227, 93609187, 112406640, 492658670, 1824768948, 812517469, 1055381989, 670973674,
           1407235524, 891397172, 1003245378, 1381303998, 1564172645, 1399931635, 1005462965]
    assert R.permute(inp, rc=(ini, mid, fin)) == exp


def test_poseidon2_constants_and_c_vs_python():
    rc, diag = O.constants()
    flat = [x for r in R.RC_INIT for x in r] + R.RC_INT + [x for r in R.RC_TERM for x in r]
    assert O.from_monty(rc).tolist() == flat  # embedded table == Grain LFSR derivation
    assert O.from_monty(diag).tolist() == R.DIAG_V
    assert R.RC_INIT[0][0] == 0x69CBB6AF and R.RC_INT[0] == 0x5A8053C0 and R.RC_TERM[3][15] == 0x608758B8
    rng = np.random.default_rng(3)
    st = rng.integers(0, P, (20, 16))
    got = O.from_monty(O.permute(O.to_monty(st)))
    for i in range(20):
        assert got[i].tolist() == R.permute(st[i].tolist())
    assert R.permute(list(range(16)))[:4] == [1906786279, 1737026427, 1959749225, 700325316]
    assert R.permute([0] * 16)[:4] == [1168947398, 128782440, 747404447, 883925857]


# ----------------------------------------------------------------------------- sponge/compress/merkle (a4-a7)
def test_b3_fixture_single_matrix_opening(kats):
    k = kats["b3_single_matrix"]
    row = u32(k["row"]).reshape(1, -1)
    assert np.array_equal(O.hash_rows(row)[0], u32(k["leaf_digest"]))
    assert O.merkle_verify([k["row"]], [k["height"]], k["path"], k["index"], k["root"])
    assert not O.merkle_verify([k["row"]], [k["height"]], k["path"], k["index"] ^ 1, k["root"])


def test_b3b_fixture_mixed_height_openings(kats):
    for k in kats["b3b_mixed_height"]:
        assert O.merkle_verify(k["rows"], k["heights"], k["path"], k["index"], k["root"]), k["name"]
        bad = [list(r) for r in k["rows"]]
        bad[-1][0] ^= 1
        assert not O.merkle_verify(bad, k["heights"], k["path"], k["index"], k["root"])


@pytest.mark.parametrize("shapes", [[(8, 3)], [(16, 9), (16, 8), (4, 5), (1, 2)], [(4, 17), (32, 1), (8, 8), (8, 7)], [(1, 5)], [(2, 16), (1, 1)]])
def test_merkle_commit_c_vs_python_and_open(shapes):
    rng = np.random.default_rng(len(shapes) + shapes[0][0])
    mats = [O.to_monty(rng.integers(0, P, s)) for s in shapes]
    root, layers = O.merkle_commit(mats)
    proot, players = R.merkle_commit([O.from_monty(m).tolist() for m in mats])
    assert O.from_monty(root).tolist() == proot
    assert [O.from_monty(l).tolist() for l in layers] == players
    max_h = max(s[0] for s in shapes)
    for idx in range(max_h):
        rows, path = O.merkle_open(mats, layers, idx)
        assert O.merkle_verify(rows, [s[0] for s in shapes], path, idx, root)
        assert R.verify_batch([O.from_monty(r).tolist() for r in rows], [s[0] for s in shapes],
                              [O.from_monty(p).tolist() for p in path], idx, proot)


def test_merkle_rejects_non_pow2():
    with pytest.raises(ValueError):
        O.merkle_commit([np.zeros((6, 3), np.uint32)])


# ----------------------------------------------------------------------------- DFT / LDE (a2)
@pytest.mark.parametrize("n,w", [(1, 3), (2, 2), (8, 5), (64, 3), (256, 2)])
def test_dft_matches_naive_definition(n, w):
    rng = np.random.default_rng(n)
    a = O.to_monty(rng.integers(0, P, (n, w)))
    nat = O.naive_dft(a)
    assert np.array_equal(O.dft_batch(a), nat)
    lg = n.bit_length() - 1
    br = O.dft_batch(a, bitrev_out=True)
    assert np.array_equal(br, nat[[R.bitrev(i, lg) for i in range(n)]])
    assert np.array_equal(O.dft_batch(nat, inverse=True), a)
    assert np.array_equal(O.naive_dft(nat, inverse=True), a)
    if n <= 8:
        for c in range(w):
            assert O.from_monty(nat[:, c]).tolist() == R.naive_dft(O.from_monty(a[:, c]).tolist())


def test_b6_fixture_coset_lde(kats):
    k = kats["b6_coset_lde"]
    got = O.coset_lde_batch(u32(k["trace"]), k["added_bits"], k["shift"], bitrev_out=True)
    assert np.array_equal(got, u32(k["lde_bitrev_rows"]))
    assert O.from_monty(got[0]).tolist() == [2005799821, 1942942217, 2013265906, 0, 2011755017]  # SURVEY B-6 row j=0


@pytest.mark.parametrize("n,w,b", [(4, 3, 1), (16, 2, 2), (64, 5, 1), (32, 1, 3)])
def test_coset_lde_properties(n, w, b):
    rng = np.random.default_rng(n + b)
    ev = O.to_monty(rng.integers(0, P, (n, w)))
    shift = int(O.to_monty([31])[0])
    nat = O.coset_lde_batch(ev, b, shift, bitrev_out=False)
    pr = R.coset_lde_batch_bitrev(O.from_monty(ev).tolist(), b, 31)
    br = O.coset_lde_batch(ev, b, shift, bitrev_out=True)
    assert O.from_monty(br).tolist() == pr
    # shift = 1: every 2^b-th logical row reproduces the input
    nat1 = O.coset_lde_batch(ev, b, O.MONTY_ONE, bitrev_out=False)
    assert np.array_equal(nat1[:: 1 << b], ev)
    # Horner spot check: coefficient form evaluated at shift * w'^i
    coeffs = O.from_monty(O.dft_batch(ev, inverse=True))
    m = n << b
    wp = R.two_adic_generator(m.bit_length() - 1)
    for i in (0, 1, m // 2 + 1, m - 1):
        x = 31 * pow(wp, i, P) % P
        for c in range(w):
            acc = 0
            for coef in reversed(coeffs[:, c].tolist()):
                acc = (acc * x + coef) % P
            assert O.from_monty(nat[i : i + 1, c])[0] == acc


# ----------------------------------------------------------------------------- FRI (a8)
def test_b5_fixture_fri_last_layer(kats):
    k = kats["b5_fri_last_layer"]
    layer = u32(k["layer_bitrev"])  # 8 EF4
    root, _ = O.merkle_commit([layer.reshape(4, 8)])
    assert np.array_equal(root, u32(k["root"]))
    folded = O.fri_fold(layer, u32(k["beta"]))
    assert all(np.array_equal(f, u32(k["folded_const"])) for f in folded)
    assert O.from_monty(u32(k["beta"])).tolist() == [812494840, 1225979230, 1716823227, 785360533]


def test_fri_fold_and_commit_phase_c_vs_python():
    rng = np.random.default_rng(5)
    vec = rng.integers(0, P, (64, 4))
    beta = rng.integers(0, P, 4)
    got = O.from_monty(O.fri_fold(O.to_monty(vec), O.to_monty(beta)))
    assert got.tolist() == R.fold_matrix(beta.tolist(), vec.tolist())
    ch_c, ch_p = O.Challenger(), R.DuplexChallenger()
    seed = rng.integers(0, P, 11)
    ch_c.observe(O.to_monty(seed))
    ch_p.observe_slice(seed.tolist())
    roots, betas, fin = O.fri_commit_phase(O.to_monty(vec), 1, 0, challenger=ch_c)
    pc, _, pfinal, pb = R.fri_commit_phase([vec.tolist()], ch_p, 2, 1)
    assert O.from_monty(roots).tolist() == pc and O.from_monty(betas).tolist() == pb
    # python returns the iDFT'd final poly; the C side returns the last folded vector
    assert len(fin) == 2
    nat = O.from_monty(fin).tolist()
    coeffs = [R.naive_idft([nat[i][kk] for i in range(2)]) for kk in range(4)]
    assert [coeffs[kk][0] for kk in range(4)] == pfinal[0]


def test_fri_commit_phase_low_degree_input_folds_to_constant():
    """An honest codeword (degree < n/2 evaluated on 2-blowup domain, bit-reversed) must fold to a constant."""
    rng = np.random.default_rng(7)
    ev = O.to_monty(rng.integers(0, P, (32, 4)))  # 32 evaluations of an EF4 polynomial (4 base columns)
    code = O.coset_lde_batch(ev, 1, O.MONTY_ONE, bitrev_out=True)  # 64 x 4
    betas = O.to_monty(rng.integers(0, P, (5, 4)))
    roots, bout, fin = O.fri_commit_phase(code, 1, 0, betas=betas)
    assert len(roots) == 5 and np.array_equal(bout, betas)
    assert np.array_equal(fin[0], fin[1])  # blowup-many evaluations of a constant
    roots4, _, fin4 = O.fri_commit_phase(code, 1, 2, betas=betas)  # final_poly_len = 4
    assert len(roots4) == 3 and np.array_equal(roots4, roots[:3]) and fin4.shape == (8, 4)


def test_challenger_c_vs_python():
    rng = np.random.default_rng(6)
    c, p = O.Challenger(), R.DuplexChallenger()
    for step in range(40):
        if rng.integers(0, 2):
            v = rng.integers(0, P, int(rng.integers(1, 12)))
            c.observe(O.to_monty(v))
            p.observe_slice(v.tolist())
        else:
            assert O.from_monty(u32([c.sample()]))[0] == p.sample()
    assert c.sample_bits(10) == p.sample_bits(10)
    w = c.grind(8)
    assert w == p.grind(8)
    assert O.from_monty(u32([c.sample()]))[0] == p.sample()


def test_fill_and_checksum_deterministic():
    a = O.fill(1000, 0xB2000000)
    assert a.max() < P and np.array_equal(a[10:20], O.fill(10, 0xB2000000, offset=10))
    assert O.checksum(a) == (O.checksum(a[:500]) + O.checksum(a[500:], offset=500)) & 0xFFFFFFFFFFFFFFFF


# ----------------------------------------------------------------------------- PCS open-phase primitives (8(f)-1)
def test_open_phase_primitives_definitions():
    """dot_ext_powers / interpolate_coset / reduced openings against their definitions in pure Python."""
    rng = np.random.default_rng(8)
    n, w, b = 8, 5, 1
    tr = rng.integers(0, P, (n, w))
    alpha = rng.integers(0, P, 4).tolist()
    z = rng.integers(0, P, 4).tolist()
    shift_m = int(O.to_monty([31])[0])
    lde = O.coset_lde_batch(O.to_monty(tr), b, shift_m, bitrev_out=True)          # 2n x w, bit-reversed rows
    # dot_ext_powers
    got = O.from_monty(O.dot_ext_powers(lde, O.to_monty(alpha)))
    pw = [[1, 0, 0, 0]]
    for _ in range(w - 1):
        pw.append(R.ef_mul(pw[-1], alpha))
    ldec = O.from_monty(lde).tolist()
    exp = []
    for row in ldec:
        acc = [0, 0, 0, 0]
        for c, v in enumerate(row):
            acc = R.ef_add(acc, R.ef_scale(pw[c], v))
        exp.append(acc)
    assert got.tolist() == exp
    # interpolate_coset on the low coset (first n rows of the bit-reversed LDE = evaluations on 31*H, bit-reversed)
    ys = O.from_monty(O.interpolate_coset_bitrev(lde[:n], shift_m, O.to_monty(z))).tolist()
    coeffs = [R.naive_idft(tr[:, c].tolist()) for c in range(w)]                  # p_c in coefficient form (trace domain H)
    for c in range(w):
        acc = [0, 0, 0, 0]
        for coef in reversed(coeffs[c]):
            acc = R.ef_add(R.ef_mul(acc, z), [coef, 0, 0, 0])
        assert ys[c] == acc
    # ef inverse
    assert R.ef_mul(O.from_monty(O.ef_inv(O.to_monty(z))).tolist(), z) == [1, 0, 0, 0]
    # reduced openings: (p(z) - p(x)) / (z - x) summed with alpha powers is a polynomial of degree < n-1 in x:
    rr = O.dot_ext_powers(lde, O.to_monty(alpha))
    rys = [0, 0, 0, 0]
    for c in range(w):
        rys = R.ef_add(rys, R.ef_mul(pw[c], ys[c]))
    ro = O.reduce_openings(rr, shift_m, O.to_monty(z), O.to_monty(rys), O.to_monty([1, 0, 0, 0]), np.zeros((2 * n, 4), np.uint32))
    roc = O.from_monty(ro).tolist()
    m = 2 * n
    wm = R.two_adic_generator(m.bit_length() - 1)
    for i in (0, 1, 5, m - 1):
        x = 31 * pow(wm, R.bitrev(i, m.bit_length() - 1), P) % P
        den = R.ef_sub(z, [x, 0, 0, 0])
        assert roc[i] == R.ef_mul(R.ef_sub(rys, exp[i]), R.ef_inv(den))
    # low-degree check: the quotient's 4 coefficient columns, un-bit-reversed and iDFT'd on the coset, vanish above degree n-2
    nat = np.zeros((m, 4), np.uint32)
    for i in range(m):
        nat[R.bitrev(i, m.bit_length() - 1)] = ro[i]
    co = O.from_monty(O.dft_batch(nat, shift=shift_m, inverse=True))
    assert not co[n - 1:].any() and co[: n - 1].any()


def test_avx512_fast_path_matches_scalar_restatement():
    """bench.py's CPU baseline uses the AVX-512 path of the oracle (packed Poseidon2, vectorised butterflies) when the host
    has AVX-512; it must equal the scalar restatement (the parity checker) bit for bit."""
    try:
        O.build(native=True)
        O.use_native(True)
        if not O.fast_available():
            pytest.skip("host CPU has no AVX-512: the fast path is not compiled in")
        rng = np.random.default_rng(512)
        shift = int(O.to_monty([31])[0])
        for n, w in [(1, 16), (4, 16), (9, 32), (12, 64), (13, 48), (15, 256)]:
            m = rng.integers(0, P, (1 << n, w), dtype=np.uint64).astype(np.uint32)
            for ab in (1, 2):
                ref = O.coset_lde_batch(m, ab, shift, bitrev_out=True)
                assert np.array_equal(O.fast_coset_lde_batch(m, ab, shift), ref), (n, w, ab)
            r1, layers = O.merkle_commit([ref])
            r2, dig = O.fast_merkle_commit_single(ref)
            assert np.array_equal(r1, r2) and np.array_equal(np.concatenate(layers), dig), (n, w)
        # widths the vector path does not take fall back to the scalar code inside the same entry points
        m = rng.integers(0, P, (64, 9), dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(O.fast_coset_lde_batch(m, 1, shift), O.coset_lde_batch(m, 1, shift, bitrev_out=True))
        r1, _ = O.merkle_commit([m])
        r2, _ = O.fast_merkle_commit_single(m)
        assert np.array_equal(r1, r2)
    finally:
        O.use_native(False)


def test_horner_and_column_helpers():
    """oracle helpers used by the headline GPU tests: a column of the synthetic matrix, and Horner evaluation == DFT entry"""
    n, w, seed = 6, 5, 77
    full = O.fill((1 << n) * w, seed).reshape(1 << n, w)
    for c in range(w):
        assert np.array_equal(O.fill_column(1 << n, w, c, seed), full[:, c])
    col = full[:, 2].reshape(-1, 1)
    coef = O.dft_batch(col, inverse=True).reshape(-1)
    g = O.two_adic_generator(n)
    xs = np.array([O.lib().orc_pow(g, i) for i in range(1 << n)], np.uint32)
    assert np.array_equal(O.eval_poly_many(coef, xs), col.reshape(-1))
