#!/usr/bin/env python
"""Generates tests/golden/headline_2p23x256.json with the ORACLE (oracle/bb_oracle.c, CPU): the bench.py default workload
(BASELINE.json configs[1] and [2]) -- coset LDE of a 2^23 x 256 trace (log_blowup 1, shift 31, bit-reversed rows) and the
Poseidon2 MerkleTreeMmcs commit of the 2^24 x 256 LDE matrix.

Input: trace[r][c] = splitmix64(seed ^ (r * 256 + c)) mod p as the Montgomery word (b200zk_mat_fill / orc_fill), seed as in
bench.py.  Recorded: the Merkle root, the order-independent checksum of every 2^16-row block of the LDE (256 blocks, both coset
halves), and four sample rows.  Needs ~30 GB of host memory and a few minutes of CPU; run where that is available:
    python tests/golden/make_headline_golden.py [log_rows width]
"""
import json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
SEED = 0xB2000000 + (n << 16) + w          # bench.py: seed of the synthetic trace
BLOCK = 16
t0 = time.time()
trace = O.fill((1 << n) * w, SEED).reshape(1 << n, w)
shift = int(O.to_monty([31])[0])
lde = O.coset_lde_batch(trace, 1, shift, bitrev_out=True)
t1 = time.time()
root, _ = O.merkle_commit([lde])
t2 = time.time()
rows = lde.shape[0]
nb = rows >> BLOCK
out = {
    "what": "oracle LDE + MerkleTreeMmcs commit of the bench.py default workload", "log_rows": n, "width": w, "log_blowup": 1, "shift_canonical": 31,
    "seed": SEED, "input_checksum": int(O.checksum(trace)), "root": [int(x) for x in root],
    "lde_checksum": int(O.checksum(lde)), "block_log_rows": BLOCK,
    "block_checksums": [int(O.checksum(lde[b << BLOCK:(b + 1) << BLOCK])) for b in range(nb)],
    "sample_rows": {str(j): [int(x) for x in lde[j]] for j in (0, 1, rows // 2 + 12345, rows - 1)},
    "oracle_seconds": {"lde": round(t1 - t0, 1), "commit": round(t2 - t1, 1)},
}
name = f"headline_2p{n}x{w}.json"
json.dump(out, open(os.path.join(ROOT, "tests", "golden", name), "w"), indent=1)
print(name, "root", out["root"], "lde %.1fs commit %.1fs" % (t1 - t0, t2 - t1))
