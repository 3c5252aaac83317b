"""Query-phase known answers mined from the reference's own proof fixture without the transcript
(oracle/mine_fixture_queries.py -> tests/golden/chunk_proof_phase2_queries.json):
all 42 query indices, complete input openings of four queries against the proof's commitments (1 / 17 / 17 / 62 matrices of
mixed heights), commit-phase openings, and the betas of the five FRI rounds that are determined by exposed values alone.
CPU part: the oracle reproduces them.  GPU part: the device MMCS verifier and the device fold reproduce them."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chunk_proof_phase2_queries.json")))
LOG_MAX = G["log_max_height"]
HEIGHTS = [d << G["log_blowup"] for d in G["degrees"]]
QUOT_HEIGHTS = [h for h, c in zip(HEIGHTS, G["quotient_chunks"]) for _ in range(c)]
BATCHES = {"2": ([1 << 19], "cached_main"), "3": (HEIGHTS, "common_main"), "4": (HEIGHTS, "after_challenge"), "5": (QUOT_HEIGHTS, "quotient")}


def u32(x):
    return np.asarray(x, dtype=np.uint32)


def layer_vectors(r):
    """layer r as a dense EF4 vector (zeros where nothing is exposed) plus the pair indices beta_r was derived from"""
    lay = G["exposed_layers"][str(r)]
    vec = np.zeros((1 << (LOG_MAX - r), 4), np.uint32)
    for p, v in lay.items():
        vec[int(p)] = v
    return vec


def test_query_indices_shape_and_known_answers():
    idx = G["query_indices"]
    assert len(idx) == 42 and idx[0] == 1879182 and all(0 <= i < (1 << LOG_MAX) for i in idx)      # SURVEY B-3b: query 0
    assert G["commit_phase_openings_verified"] >= 100 and sorted(map(int, G["betas_without_transcript"])) == [14, 15, 16, 17, 19]


def test_input_openings_verify_with_oracle():
    for qi, q in G["full_queries"].items():
        index = G["query_indices"][int(qi)]
        for bi, (hs, root_name) in BATCHES.items():
            b = q["batches"][bi]
            i = index >> (LOG_MAX - 19) if bi == "2" else index
            assert O.merkle_verify(b["rows"], hs, b["path"], i, G["roots"][root_name]), (qi, bi)
            assert not O.merkle_verify(b["rows"], hs, b["path"], i ^ 1, G["roots"][root_name])


def test_commit_phase_openings_and_betas_with_oracle():
    n_rounds = len(G["commit_phase_commits"])
    for qi, q in G["full_queries"].items():
        index = G["query_indices"][int(qi)]
        for r in range(n_rounds - 6, n_rounds):
            lay = G["exposed_layers"][str(r)]
            own = index >> r
            assert lay[str(own ^ 1)] == q["commit_phase"][r]["sibling"]
            if str(own) in lay:        # own value exposed by another query: the whole opening can be checked
                leaf = [lay[str(own & ~1)] + lay[str(own | 1)]]
                assert O.merkle_verify(leaf, [1 << (LOG_MAX - r - 1)], q["commit_phase"][r]["path"], own >> 1, G["commit_phase_commits"][r])
    for r, b in G["betas_without_transcript"].items():
        r = int(r)
        folded = O.fri_fold(layer_vectors(r), u32(b["beta"]))
        nxt = layer_vectors(r + 1) if r + 1 < n_rounds else np.tile(u32(G["final_poly"][0]), (1 << (LOG_MAX - n_rounds), 1))
        for i in b["pairs"]:
            assert folded[i].tolist() == nxt[i].tolist(), (r, i)


@pytest.mark.gpu
def test_device_verifier_and_fold_reproduce_the_fixture():
    import zkvm_prover_b200 as z
    ctx = z.default_context(0)
    mmcs = z.MerkleTreeMmcs(ctx)
    for qi, q in G["full_queries"].items():
        index = G["query_indices"][int(qi)]
        for bi, (hs, root_name) in BATCHES.items():
            b = q["batches"][bi]
            i = index >> (LOG_MAX - 19) if bi == "2" else index
            dims = [(len(row), h) for row, h in zip(b["rows"], hs)]
            mmcs.verify_batch(u32(G["roots"][root_name]), dims, i, [u32(r) for r in b["rows"]], u32(b["path"]))
            with pytest.raises(Exception):
                mmcs.verify_batch(u32(G["roots"][root_name]), dims, i ^ 2, [u32(r) for r in b["rows"]], u32(b["path"]))
    n_rounds = len(G["commit_phase_commits"])
    for r, b in G["betas_without_transcript"].items():
        r = int(r)
        folded = z.fold_matrix(u32(b["beta"]), layer_vectors(r), ctx)
        nxt = layer_vectors(r + 1) if r + 1 < n_rounds else np.tile(u32(G["final_poly"][0]), (1 << (LOG_MAX - n_rounds), 1))
        for i in b["pairs"]:
            assert folded[i].tolist() == nxt[i].tolist(), (r, i)
