#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden root for a REAL trace shape.

Shape = the 17 AIRs of the reference's aggregation-layer proof fixture (heights 2 ... 2^20, common-main widths
1 ... 398, log_blowup 2; mined by oracle/mine_fixture.py into tests/golden/chunk_proof_phase2_kats.json).  The traces are
synthetic (orc_fill, seed 0x5EA1 + air index) because the real witness is not part of the fixture; the commitment
structure (mixed heights, ragged widths, injection levels) is the real one.  The oracle computes
    TwoAdicFriPcs::commit = coset LDE (shift 31, bit-reversed rows) of every trace + one MerkleTreeMmcs commit
and the result is stored in tests/golden/real_shape_commit.json.  Takes ~2 minutes on 8 cores.
"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from oracle import oracle as O

k = json.load(open(os.path.join(HERE, "..", "tests", "golden", "chunk_proof_phase2_kats.json")))
main = [b for b in k["b3b_mixed_height"] if b["name"] == "common_main"][0]
degrees = k["degrees"]
widths = [len(r) for r in main["rows"]]
log_blowup = k["log_blowup"]
shift = int(O.to_monty([31])[0])
t0 = time.time()
ldes = []
for i, (d, w) in enumerate(zip(degrees, widths)):
    tr = O.fill(d * w, 0x5EA1 + i).reshape(d, w)
    ldes.append(O.coset_lde_batch(tr, log_blowup, shift, bitrev_out=True))
root, layers = O.merkle_commit(ldes)
index = 1879182  # the fixture's own query index
rows, path = O.merkle_open(ldes, layers, index)
assert O.merkle_verify(rows, [m.shape[0] for m in ldes], path, index, root)
out = {"degrees": degrees, "widths": widths, "log_blowup": log_blowup, "seed_base": 0x5EA1, "root": root.tolist(), "index": index,
       "opened_rows": [r.tolist() for r in rows], "path": path.tolist(), "lde_checksums": [int(O.checksum(m)) for m in ldes]}
json.dump(out, open(os.path.join(HERE, "..", "tests", "golden", "real_shape_commit.json"), "w"), separators=(",", ":"))
print("root", root.tolist(), "elements", sum(d * w for d, w in zip(degrees, widths)), "seconds", round(time.time() - t0, 1))
