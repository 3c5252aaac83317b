"""Value parity at the HEADLINE shapes: the bench.py default workload -- coset LDE of a 2^23 x 256
trace (log_blowup 1, shift 31, bit-reversed rows) + Poseidon2 MerkleTreeMmcs commit of the 2^24 x 256 LDE matrix
(BASELINE.json configs[1], [2]) -- against the oracle, by value:

  * Merkle root and every 2^16-row block checksum of the LDE (both coset halves) against tests/golden/headline_2p23x256.json,
    which oracle/bb_oracle.c produced on the CPU (tests/golden/make_headline_golden.py);
  * >= 1024 definition-level spot checks per shape at 2^22 x 256, 2^24 x 128 and 2^24 x 512: the oracle interpolates single
    columns (iDFT) and evaluates them with Horner's rule at 31 * w'^bitrev(row) for random rows of BOTH coset halves -- no
    fast transform on the checking side of the forward direction;
  * the strip-pipelined host entry point (b200zk_lde_commit_host) yields the same root.
All calls go through the C ABI; the oracle is the checker only."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

P = O.P
GOLD = os.path.join(os.path.dirname(__file__), "golden", "headline_2p23x256.json")


@pytest.fixture(scope="module")
def z():
    import zkvm_prover_b200 as zz
    return zz


@pytest.fixture(scope="module")
def ctx(z):
    return z.default_context(0)


def bitrev(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def free_gb():
    import torch
    return torch.cuda.mem_get_info()[0] / (1 << 30)


def test_headline_2p23x256_lde_and_root_match_oracle_golden(z, ctx):
    g = json.load(open(GOLD))
    n, w = g["log_rows"], g["width"]
    ctx.trim()
    if free_gb() < 64:
        pytest.skip("needs ~60 GB of free device memory")
    trace = ctx.alloc(1 << n, w).fill(g["seed"])
    assert trace.checksum() == g["input_checksum"]
    shift = int(O.to_monty([g["shift_canonical"]])[0])
    lde = z.B200Dft(ctx).coset_lde_batch(trace, g["log_blowup"], shift, bit_reversed=True)
    assert lde.checksum() == g["lde_checksum"]
    blk = g["block_log_rows"]
    nb = lde.rows >> blk
    assert nb == len(g["block_checksums"])
    base = lde.device_ptr
    for b in range(nb):                       # every 2^16-row block of both coset halves, by value (checksum of (index, value) pairs)
        sub = ctx.wrap(base + (b << blk) * w * 4, 1 << blk, w, keepalive=lde)
        assert sub.checksum() == g["block_checksums"][b], f"LDE block {b} differs from the oracle"
        sub.free()
    for j, row in g["sample_rows"].items():
        assert lde.rows_to_host(int(j), 1)[0].tolist() == row
    root, pd = z.MerkleTreeMmcs(ctx).commit([lde])
    assert [int(x) for x in root] == g["root"], "Merkle root of the 2^24 x 256 LDE matrix differs from the oracle's"
    pd.free()
    # the same through the PCS mirror (one call: LDE + commit) and through the host strip pipeline (e2e path of bench.py)
    pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=g["log_blowup"]), ctx)
    root2, pd2 = pcs.commit([trace])
    assert [int(x) for x in root2] == g["root"]
    pd2.free()
    host = trace.to_host()
    trace.free()
    root3, pd3 = pcs.commit_host(host)
    assert [int(x) for x in root3] == g["root"]
    pd3.free()
    lde.free()
    ctx.trim()


@pytest.mark.parametrize("n,w,need_gb", [(22, 256, 16), (24, 128, 30), (24, 512, 110)])
def test_lde_horner_spot_checks(z, ctx, n, w, need_gb):
    """1024 (row, column) entries of the shift-31 LDE equal the oracle's definition: interpolate the column, evaluate it at
    31 * w'^bitrev(row).  Rows are drawn from both coset halves (physical row < N: block 0, >= N: block 1)."""
    ctx.trim()
    if free_gb() < need_gb:
        pytest.skip(f"needs ~{need_gb} GB of free device memory")
    N = 1 << n
    seed = 0xB2000000 + (n << 16) + w
    trace = ctx.alloc(N, w).fill(seed)
    shift = int(O.to_monty([31])[0])
    lde = z.B200Dft(ctx).coset_lde_batch(trace, 1, shift, bit_reversed=True)
    trace.free()
    rng = np.random.default_rng(n * 1000 + w)
    cols = sorted(set([0, w - 1] + [int(c) for c in rng.integers(0, w, 6)]))[:8]
    while len(cols) < 8:
        cols = sorted(set(cols + [int(rng.integers(0, w))]))
    rows = np.concatenate([rng.integers(0, N, 62), rng.integers(N, 2 * N, 62), [0, N - 1, N, 2 * N - 1]]).astype(np.int64)
    wp = O.two_adic_generator(n + 1)
    g31 = int(O.to_monty([31])[0])
    xs = np.array([O.lib().orc_mul(g31, O.lib().orc_pow(wp, bitrev(int(j), n + 1))) for j in rows], np.uint32)
    got = np.stack([lde.rows_to_host(int(j), 1)[0] for j in rows])          # (128, w)
    checked = 0
    for c in cols:
        col = O.fill_column(N, w, c, seed).reshape(N, 1)
        coef = O.dft_batch(col, inverse=True).reshape(-1)                   # natural-order coefficients of the column
        exp = O.eval_poly_many(coef, xs)
        assert np.array_equal(got[:, c], exp), f"column {c}"
        checked += len(rows)
    assert checked >= 1024
    lde.free()
    ctx.trim()


def test_dft_natural_order_wide_n9(z, ctx):
    """a 512-row natural-order DFT / iDFT wide enough that its tiles are not all co-resident (the scattering
    last pass used to run in place when the two planners disagreed on the pass count)."""
    m = np.random.default_rng(9).integers(0, P, (512, 4096), dtype=np.uint64).astype(np.uint32)
    dft = z.B200Dft(ctx)
    dm = ctx.upload(m)
    assert np.array_equal(dft.dft_batch(dm).to_host(), O.dft_batch(m, inverse=False, bitrev_out=False))
    assert np.array_equal(dft.idft_batch(dm).to_host(), O.dft_batch(m, inverse=True, bitrev_out=False))
    shift = int(O.to_monty([31])[0])
    assert np.array_equal(dft.coset_dft_batch(dm, shift).to_host(), O.dft_batch(m, shift=shift, inverse=False, bitrev_out=False))


def test_more_than_128_matrices_of_one_height(z, ctx):
    """p3's MerkleTreeMmcs has no limit on matrices per height; 150 of them go through the device-side descriptor array."""
    rng = np.random.default_rng(150)
    mats = [rng.integers(0, P, (64, int(rng.integers(1, 12))), dtype=np.uint64).astype(np.uint32) for _ in range(150)]
    mats += [rng.integers(0, P, (16, 3), dtype=np.uint64).astype(np.uint32) for _ in range(140)]    # an inject level with > 128 too
    mmcs = z.MerkleTreeMmcs(ctx)
    root, pd = mmcs.commit([ctx.upload(m) for m in mats])
    oroot, layers = O.merkle_commit(mats)
    assert np.array_equal(root, oroot)
    rows, path = mmcs.open_batch(37, pd)
    orows, opath = O.merkle_open(mats, layers, 37)
    assert all(np.array_equal(a, b) for a, b in zip(rows, orows)) and np.array_equal(path, opath)
    pd.free()


def test_grind_zero_bits_leaves_transcript_alone(z, ctx):
    """p3-challenger 0.4.3 (the reference's pin): proof_of_work_bits == 0 needs no witness and does not touch the transcript"""
    c, o = z.DuplexChallenger(ctx), O.Challenger()
    v = np.arange(1, 12, dtype=np.uint32)
    c.observe(v)
    o.observe(v)
    before = c.state().copy()
    assert c.grind(0) == 0 and o.grind(0) == 0
    assert np.array_equal(c.state(), before) and np.array_equal(c.state(), o.state())
    assert c.grind(6) == o.grind(6) and np.array_equal(c.state(), o.state())


def test_challenger_state_handover(z, ctx):
    """b200zk_chal_set_state: a transcript can move between a host challenger and the device one at any point"""
    a, o = z.DuplexChallenger(ctx), O.Challenger()
    v = np.arange(3, 16, dtype=np.uint32)
    a.observe(v)
    o.observe(v)
    assert a.sample() == o.sample()
    b = z.DuplexChallenger(ctx).set_state(o.state())          # hand the ORACLE's (host) state to a fresh device challenger
    for _ in range(11):
        assert b.sample() == a.sample()
    b.observe(v[:5])
    a.observe(v[:5])
    assert np.array_equal(a.state(), b.state())
    bad = a.state().copy()
    bad[24] = 9
    with pytest.raises(z.B200zkError):
        b.set_state(bad)
