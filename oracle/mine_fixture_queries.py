#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- second mining pass over the reference's proof fixture (transcript-free part of SURVEY.md
section 8(f)2): everything about the QUERY phase that can be derived from the proof bytes and its own commitments alone.

    python oracle/mine_fixture_queries.py          (build container: needs /root/reference)
-> tests/golden/chunk_proof_phase2_queries.json (committed; nothing at test time reads /root/reference)

What is derived, each item checked with the oracle before it is written:
  Q-1  the query index of EVERY one of the 42 queries: the 19-level opening of main_trace[0] fixes the top 19 bits (all 2^19
       left/right patterns are walked level by level, 2^20 compressions per query, exactly one pattern reaches the root); the
       22-level mixed-height opening of main_trace[1] fixes the remaining three.
  Q-2  all 42 x 4 input openings (batches 2..5: cached main, common main, after-challenge, quotient) verify at those
       indices against the proof's own commitments -- MerkleTreeMmcs::verify_batch with 1 / 17 / 17 / 62 matrices.
  Q-3  the FRI commit-phase layers of the last rounds, reassembled from the sibling values the queries expose (round r,
       position (index >> r) ^ 1); every commit-phase opening whose own value is exposed by another query is verified
       against commit_phase_commits[r] (leaf = [value at 2i | value at 2i+1], EF4 coefficients low..high).
  Q-4  beta_r for the rounds where it is determined without the transcript: round r folds layer r (length 2^(22-r)) into
       layer r+1 and adds the reduced opening of the matrices of that height, so beta_r is computable exactly when no
       committed matrix has LDE height 2^(21-r) and some pair of layer r plus its image in layer r+1 are both exposed.
       Every further exposed pair must give the SAME beta (over-determined).
What is NOT derivable without the verifying key's pre-hash (absent from the checkout): the challenger state, hence alpha,
zeta, the betas of the other rounds, and whether the roll-in of a lower reduced opening is `+ ro` or `+ beta^2 * ro`.
"""
from __future__ import annotations

import base64
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
from oracle import pyref as R  # noqa: E402
from oracle.mine_fixture import FIXTURE, SHA256, QUERY0_INDEX, parse, canon  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "chunk_proof_phase2_queries.json")
KEEP_FULL = (0, 1, 20, 41)     # queries whose complete openings are stored (the others: index only)


def find_top_bits(row, path, root):
    """all left/right patterns of a single-matrix opening, level by level; returns the unique index reaching `root`"""
    node = O.hash_rows(np.asarray([row], np.uint32))            # (1, 8)
    cand = node                                                   # candidates after l levels, id = bits chosen so far
    for lvl, sib in enumerate(path):
        sib = np.asarray(sib, np.uint32)
        n = cand.shape[0]
        left = np.concatenate([cand, np.broadcast_to(sib, (n, 8))], axis=1)      # bit = 0: compress(node, sib)
        right = np.concatenate([np.broadcast_to(sib, (n, 8)), cand], axis=1)     # bit = 1: compress(sib, node)
        both = np.concatenate([left, right], axis=0)                               # id + bit * 2^lvl
        cand = O.compress_pairs(both)
    hits = np.nonzero((cand == np.asarray(root, np.uint32)).all(axis=1))[0]
    assert len(hits) == 1, hits
    return int(hits[0])


def main():
    raw = open(FIXTURE, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == SHA256
    pr = parse(base64.b64decode(json.loads(raw)["proof"]["proofs"]))
    degrees = [a["degree"] for a in pr["per_air"]]
    quot_chunks = [len(x) for x in pr["opened"]["quotient"]]
    log_blowup = 2
    heights = [d << log_blowup for d in degrees]
    quot_heights = [h for h, c in zip(heights, quot_chunks) for _ in range(c)]
    log_max = max(heights).bit_length() - 1
    nq = len(pr["queries"])
    roots = {2: pr["main_trace"][0], 3: pr["main_trace"][1], 4: pr["after_challenge"][0], 5: pr["quotient"]}
    hs = {2: [1 << 19], 3: heights, 4: heights, 5: quot_heights}

    # ---- Q-1 / Q-2
    indices = []
    for qi, q in enumerate(pr["queries"]):
        b2 = q["batches"][2]
        top = find_top_bits(b2["rows"][0], b2["path"], roots[2])
        full = None
        for low in range(1 << (log_max - 19)):
            idx = (top << (log_max - 19)) | low
            if O.merkle_verify(q["batches"][3]["rows"], hs[3], q["batches"][3]["path"], idx, roots[3]):
                assert full is None
                full = idx
        assert full is not None, qi
        for bi in (2, 3, 4, 5):
            idx_b = full >> (log_max - 19) if bi == 2 else full
            assert O.merkle_verify(q["batches"][bi]["rows"], hs[bi], q["batches"][bi]["path"], idx_b, roots[bi]), (qi, bi)
        indices.append(full)
        print(f"query {qi:2d}: index {full}", flush=True)
    assert indices[0] == QUERY0_INDEX

    # ---- Q-3: exposed commit-phase values.  Round r works on layer r (length 2^(log_max - r)); a query exposes position
    # ((index >> r) ^ 1) of layer r; its own value sits at (index >> r).
    n_rounds = len(pr["commit_phase_commits"])
    layers = [dict() for _ in range(n_rounds + 1)]
    for qi, q in enumerate(pr["queries"]):
        for r, step in enumerate(q["cp"]):
            pos = (indices[qi] >> r) ^ 1
            assert layers[r].setdefault(pos, step["sibling"]) == step["sibling"], "two queries disagree on an exposed value"
    # the final layer (length 2^(log_max - n_rounds) = 4 = blowup * final_poly_len): evaluations of the constant final polynomial
    final_len = 1 << (log_max - n_rounds)
    assert all(x == [0, 0, 0, 0] for x in pr["final_poly"][1:])
    for i in range(final_len):
        layers[n_rounds][i] = pr["final_poly"][0]
    verified_cp = 0
    for qi, q in enumerate(pr["queries"]):
        for r, step in enumerate(q["cp"]):
            own_pos = indices[qi] >> r
            if own_pos not in layers[r]:
                continue
            pair = [layers[r][own_pos & ~1], layers[r][own_pos | 1]]
            leaf = [pair[0] + pair[1]]
            assert O.merkle_verify(leaf, [1 << (log_max - r - 1)], step["path"], own_pos >> 1, pr["commit_phase_commits"][r]), (qi, r)
            verified_cp += 1
    print("commit-phase openings verified (own value exposed by another query):", verified_cp)

    # ---- Q-4: betas determined without the transcript
    input_log_heights = sorted({h.bit_length() - 1 for h in heights})
    half = R.inv(2)
    betas = {}
    for r in range(n_rounds):
        out_log = log_max - r - 1                         # length of layer r+1
        if out_log in input_log_heights:
            continue                                       # a reduced opening is rolled in here: beta_r needs alpha and zeta
        g = R.two_adic_generator(log_max - r)
        sols = []
        for i in range(1 << out_log):
            if 2 * i in layers[r] and 2 * i + 1 in layers[r] and i in layers[r + 1]:
                lo, hi, tgt = canon(layers[r][2 * i]), canon(layers[r][2 * i + 1]), canon(layers[r + 1][i])
                x = pow(g, R.bitrev(i, out_log), R.P)
                even = R.ef_scale(R.ef_add(lo, hi), half)
                odd = R.ef_scale(R.ef_sub(lo, hi), half * R.inv(x) % R.P)
                sols.append((i, R.ef_mul(R.ef_sub(tgt, even), R.ef_inv(odd))))
        if sols:
            assert all(s[1] == sols[0][1] for s in sols), (r, sols)
            betas[r] = {"beta": [R.to_monty(x) for x in sols[0][1]], "pairs": [s[0] for s in sols]}
            print(f"round {r}: beta determined by {len(sols)} exposed pair(s), all equal")
    assert n_rounds - 1 in betas            # B-5's beta (last round) must be among them

    gold = {
        "source": FIXTURE.replace("/root/reference/", ""), "sha256": SHA256, "log_blowup": log_blowup, "log_max_height": log_max,
        "encoding": "all field elements are Montgomery-form u32 exactly as on the wire",
        "degrees": degrees, "quotient_chunks": quot_chunks, "query_indices": indices,
        "roots": {"cached_main": roots[2], "common_main": roots[3], "after_challenge": roots[4], "quotient": roots[5]},
        "commit_phase_commits": pr["commit_phase_commits"], "final_poly": pr["final_poly"],
        "full_queries": {str(qi): {"batches": {str(bi): pr["queries"][qi]["batches"][bi] for bi in (2, 3, 4, 5)},
                                   "commit_phase": pr["queries"][qi]["cp"]} for qi in KEEP_FULL},
        "exposed_layers": {str(r): {str(p): v for p, v in sorted(layers[r].items())} for r in range(n_rounds - 6, n_rounds + 1)},
        "betas_without_transcript": {str(r): b for r, b in betas.items()},
        "commit_phase_openings_verified": verified_cp,
    }
    with open(OUT, "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes; betas for rounds", sorted(betas))


if __name__ == "__main__":
    main()
