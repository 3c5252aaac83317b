#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- mines golden vectors for the hot path out of the reference's own
checked-in proof fixture and writes them to tests/golden/.

Run in the build container (where /root/reference exists):
    python oracle/mine_fixture.py
The resulting tests/golden/*.json are committed; nothing at test/bench time reads /root/reference.

Source: /root/reference/crates/verifier/testdata/proofs/chunk-proof-phase2.json
  (sha256 e514b529004b99978d8e6cb2bf9eee792559ee25d978d8f9a4357588455a2fd7), whose `proof.proofs`
  is base64(bincode v1 Vec<Proof<SC>>) -- the legacy VmInternalStarkProof wire format of
  /root/reference/crates/types/src/proof.rs:70-74.  All u32 are Montgomery-form BabyBear.

What is mined (SURVEY.md Appendix B numbering):
  B-3   one single-matrix Merkle opening (width 9, 19-digest path) against main_trace[0]
  B-3b  three mixed-height MerkleTreeMmcs openings (17 / 17 / 62 matrices) against
        main_trace[1], after_challenge[0], quotient
  B-5   the last FRI commit-phase layer (8 EF4 values reassembled from 42 queries), its Merkle
        root, the final polynomial, and the beta the fold relation implies
  B-6   a complete 8x5 coset LDE (AIR 1: trace height 2, log_blowup 2) in bit-reversed row order
Every vector is verified with oracle/pyref.py before it is written: a mismatch aborts.
"""
from __future__ import annotations

import base64
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pyref as R  # noqa: E402

FIXTURE = "/root/reference/crates/verifier/testdata/proofs/chunk-proof-phase2.json"
SHA256 = "e514b529004b99978d8e6cb2bf9eee792559ee25d978d8f9a4357588455a2fd7"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "chunk_proof_phase2_kats.json")

QUERY0_INDEX = 1879182  # found once by exhausting all 2^19 left/right patterns (SURVEY B-3); verified below


class Rd:
    def __init__(self, b):
        self.b, self.o = b, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.o)[0]
        self.o += 4
        return v

    def u32s(self, n):
        v = list(struct.unpack_from("<%dI" % n, self.b, self.o))
        self.o += 4 * n
        return v

    def digest(self):
        return self.u32s(8)

    def ef(self):
        return self.u32s(4)


def canon(v):
    return [R.from_monty(x) for x in v]


def parse(blob):
    r = Rd(blob)
    assert r.u64() == 1, "one Proof in the Vec"
    pr = {}
    n_main = r.u64()
    pr["main_trace"] = [r.digest() for _ in range(n_main)]
    n_after = r.u64()
    pr["after_challenge"] = [r.digest() for _ in range(n_after)]
    pr["quotient"] = r.digest()
    n_cp = r.u64()
    pr["commit_phase_commits"] = [r.digest() for _ in range(n_cp)]
    n_q = r.u64()
    queries = []
    for _ in range(n_q):
        q = {"batches": [], "cp": []}
        for _ in range(r.u64()):
            mats = []
            for _ in range(r.u64()):
                w = r.u64()
                mats.append(r.u32s(w))
            path = [r.digest() for _ in range(r.u64())]
            q["batches"].append({"rows": mats, "path": path})
        for _ in range(r.u64()):
            sib = r.ef()
            path = [r.digest() for _ in range(r.u64())]
            q["cp"].append({"sibling": sib, "path": path})
        queries.append(q)
    pr["queries"] = queries
    pr["final_poly"] = [r.ef() for _ in range(r.u64())]
    pr["pow_witness"] = r.u32()

    def adj():  # AdjacentOpenedValues {local, next}
        loc = [r.ef() for _ in range(r.u64())]
        nxt = [r.ef() for _ in range(r.u64())]
        return {"local": loc, "next": nxt}

    ov = {}
    ov["preprocessed"] = [adj() for _ in range(r.u64())]
    ov["main"] = [[adj() for _ in range(r.u64())] for _ in range(r.u64())]
    ov["after_challenge"] = [[adj() for _ in range(r.u64())] for _ in range(r.u64())]
    ov["quotient"] = [[[r.ef() for _ in range(r.u64())] for _ in range(r.u64())] for _ in range(r.u64())]
    pr["opened"] = ov
    per_air = []
    for _ in range(r.u64()):
        a = {"air_id": r.u64(), "degree": r.u64()}
        a["exposed"] = [[r.ef() for _ in range(r.u64())] for _ in range(r.u64())]
        a["public_values"] = r.u32s(r.u64())
        per_air.append(a)
    pr["per_air"] = per_air
    tag = r.b[r.o]
    r.o += 1
    if tag:
        pr["logup_pow_witness"] = r.u32()
    assert r.o == len(blob), (r.o, len(blob))
    return pr


def main():
    raw = open(FIXTURE, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == SHA256
    blob = base64.b64decode(json.loads(raw)["proof"]["proofs"])
    pr = parse(blob)
    degrees = [a["degree"] for a in pr["per_air"]]
    n_air = len(degrees)
    quot_chunks = [len(x) for x in pr["opened"]["quotient"]]
    log_blowup = 2
    heights = [d << log_blowup for d in degrees]
    log_max = max(heights).bit_length() - 1
    print("AIRs", n_air, "degrees", degrees, "quotient chunks", quot_chunks, "log_max_height", log_max)
    q0 = pr["queries"][0]
    gold = {"source": FIXTURE.replace("/root/reference/", ""), "sha256": SHA256,
            "encoding": "all field elements are Montgomery-form u32 exactly as on the wire",
            "log_blowup": log_blowup, "degrees": degrees}

    # ---- B-3: single matrix opening against main_trace[0] (cached-main commitment of one AIR)
    b2 = q0["batches"][2]
    assert len(b2["rows"]) == 1 and len(b2["path"]) == 19
    idx_b2 = QUERY0_INDEX >> (log_max - 19)
    assert R.verify_batch([canon(b2["rows"][0])], [1 << 19], [canon(p) for p in b2["path"]], idx_b2,
                          canon(pr["main_trace"][0])), "B-3 failed"
    leaf = R.hash_iter(canon(b2["rows"][0]))
    gold["b3_single_matrix"] = {"index": idx_b2, "height": 1 << 19, "row": b2["rows"][0], "path": b2["path"],
                                "root": pr["main_trace"][0], "leaf_digest": [R.to_monty(x) for x in leaf]}
    print("B-3 ok: index", idx_b2)

    # ---- B-3b: mixed-height batches
    gold["b3b_mixed_height"] = []
    quot_heights = [h for h, c in zip(heights, quot_chunks) for _ in range(c)]
    for name, bi, hs, root in [("common_main", 3, heights, pr["main_trace"][1]),
                               ("after_challenge", 4, heights, pr["after_challenge"][0]),
                               ("quotient", 5, quot_heights, pr["quotient"])]:
        b = q0["batches"][bi]
        assert len(b["rows"]) == len(hs), (name, len(b["rows"]), len(hs))
        ok = R.verify_batch([canon(x) for x in b["rows"]], hs, [canon(p) for p in b["path"]], QUERY0_INDEX, canon(root))
        assert ok, "B-3b failed for " + name
        gold["b3b_mixed_height"].append({"name": name, "index": QUERY0_INDEX, "heights": hs, "rows": b["rows"],
                                         "path": b["path"], "root": root})
        print("B-3b ok:", name, "mats", len(hs), "widths", [len(x) for x in b["rows"]][:8], "...")

    # ---- B-5: last FRI commit-phase layer, reassembled without the transcript
    n_rounds = len(pr["commit_phase_commits"])
    last_root = canon(pr["commit_phase_commits"][-1])
    sibs = [q["cp"][-1] for q in pr["queries"]]
    assert all(len(s["path"]) == 2 for s in sibs)
    layer = {}  # position (0..7) -> EF4 (monty)
    tops = [None] * len(sibs)  # top-3 index bits of every query
    uniq = []
    for s in sibs:
        if s["sibling"] not in uniq:
            uniq.append(s["sibling"])
    for qi, s in enumerate(sibs):
        path = [canon(p) for p in s["path"]]
        for own in uniq:
            for own_pos in (0, 1):
                pair = (own, s["sibling"]) if own_pos == 0 else (s["sibling"], own)
                leaf_row = canon(pair[0]) + canon(pair[1])
                for row in range(4):
                    if R.verify_batch([leaf_row], [4], path, row, last_root):
                        top = row * 2 + own_pos
                        assert tops[qi] in (None, top)
                        tops[qi] = top
                        layer[2 * row] = pair[0]
                        layer[2 * row + 1] = pair[1]
    assert all(t is not None for t in tops) and len(layer) == 8, (tops, sorted(layer))
    assert tops[0] == QUERY0_INDEX >> (log_max - 3), (tops[0], QUERY0_INDEX >> (log_max - 3))
    last_layer = [layer[i] for i in range(8)]
    root8, _ = R.merkle_commit([[canon(last_layer[2 * i]) + canon(last_layer[2 * i + 1]) for i in range(4)]])
    assert root8 == last_root
    # beta from the fold relation: final evaluations are those of a constant c0
    fp = pr["final_poly"]
    assert all(x == [0, 0, 0, 0] for x in fp[1:])
    c0 = canon(fp[0])
    half = R.inv(2)
    betas = []
    for i in range(4):
        lo, hi = canon(last_layer[2 * i]), canon(last_layer[2 * i + 1])
        x = pow(R.two_adic_generator(3), R.bitrev(i, 2), R.P)
        even = R.ef_scale(R.ef_add(lo, hi), half)
        odd = R.ef_scale(R.ef_sub(lo, hi), half * R.inv(x) % R.P)
        betas.append(R.ef_mul(R.ef_sub(c0, even), R.ef_inv(odd)))
    assert all(b == betas[0] for b in betas), betas
    beta = betas[0]
    folded = R.fold_matrix(beta, [canon(x) for x in last_layer])
    assert all(f == c0 for f in folded)
    gold["b5_fri_last_layer"] = {"layer_bitrev": last_layer, "root": pr["commit_phase_commits"][-1],
                                 "beta": [R.to_monty(x) for x in beta], "folded_const": fp[0],
                                 "final_poly": fp, "n_rounds": n_rounds, "query_top3_bits": tops}
    print("B-5 ok: beta", beta, "tops", tops)

    # ---- B-6: 8x5 coset LDE of AIR 1 (degree 2) in bit-reversed row order; shift = 31
    air = degrees.index(2)
    assert heights[air] == 8
    rows8 = {}
    pre8 = {}
    for qi, q in enumerate(pr["queries"]):
        rows8.setdefault(tops[qi], q["batches"][3]["rows"][air])
        assert rows8[tops[qi]] == q["batches"][3]["rows"][air]
    assert sorted(rows8) == list(range(8))
    lde = [rows8[j] for j in range(8)]
    # recover the trace: interpolate each column on the coset 31*<w8> in bit-reversed order
    width = len(lde[0])
    trace = [[0] * width for _ in range(2)]
    for c in range(width):
        nat = [0] * 8
        for j in range(8):
            nat[R.bitrev(j, 3)] = R.from_monty(lde[j][c])
        coeffs = R.naive_idft(nat)
        coeffs = [x * pow(R.inv(31), k, R.P) % R.P for k, x in enumerate(coeffs)]
        assert all(x == 0 for x in coeffs[2:]), ("degree >= 2; wrong convention", c, coeffs)
        ev = R.naive_dft(coeffs[:2])
        trace[0][c], trace[1][c] = ev
    back = R.coset_lde_batch_bitrev(trace, 2, 31)
    assert back == [canon(r) for r in lde]
    gold["b6_coset_lde"] = {"trace": [[R.to_monty(x) for x in r] for r in trace], "added_bits": 2,
                            "shift": R.to_monty(31), "lde_bitrev_rows": lde}
    print("B-6 ok: trace", trace)

    gold["commitments"] = {"main_trace": pr["main_trace"], "after_challenge": pr["after_challenge"],
                           "quotient": pr["quotient"], "commit_phase_commits": pr["commit_phase_commits"]}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
