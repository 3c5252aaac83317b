#!/usr/bin/env python
"""Generates tests/golden/proof_codec.json: run in the build container, where /root/reference exists.

For every STARK proof fixture of the reference (crates/verifier/testdata/proofs/{chunk,batch}-proof-*.json) the
`proof.proofs` / `proof.public_values` blobs are decoded with zkvm_prover_b200.proof and re-encoded; the script aborts
unless the result is byte-identical.  What is committed: the SHA-256 of each blob and a structural summary -- not the
fixtures themselves.  tests/test_proof_codec.py repeats the byte-for-byte check whenever the reference is mounted and
checks the summary against the hashes otherwise."""
import base64, glob, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from zkvm_prover_b200 import proof as W  # noqa: E402

REF = "/root/reference/crates/verifier/testdata/proofs"
out = {"source_dir": "crates/verifier/testdata/proofs", "fixtures": {}}
for path in sorted(glob.glob(os.path.join(REF, "*-proof-*.json"))):
    name = os.path.basename(path)
    d = json.load(open(path))
    pr = d.get("proof")
    if not isinstance(pr, dict) or "proofs" not in pr:
        continue  # bundle proofs are EVM proofs, a different format
    blob, pv = base64.b64decode(pr["proofs"]), base64.b64decode(pr.get("public_values", ""))
    v = W.VmInternalStarkProof.decode(blob, pv)
    assert v.encode_proofs() == blob, name
    assert v.encode_public_values() == pv, name
    assert v.to_json_fields()["proofs"] == pr["proofs"], name
    p = v.proofs[0]
    out["fixtures"][name] = {
        "file_sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
        "proofs_sha256": hashlib.sha256(blob).hexdigest(), "proofs_bytes": len(blob),
        "public_values_sha256": hashlib.sha256(pv).hexdigest(), "n_public_values": int(v.public_values.size),
        "n_proofs": len(v.proofs), "n_airs": len(p.per_air), "degrees": [a.degree for a in p.per_air],
        "n_main_commits": len(p.main_trace_commits), "fri_rounds": len(p.fri.commit_phase_commits),
        "n_queries": len(p.fri.query_proofs), "final_poly_len": len(p.fri.final_poly), "pow_witness": p.fri.pow_witness, "pow_witness_wire": int(W.monty_scalar(p.fri.pow_witness)),
        "quotient_commit": p.quotient_commit.tolist(), "has_logup_pow": p.logup_pow_witness is not None,
    }
    print(name, "ok:", len(blob), "bytes,", len(p.per_air), "AIRs,", len(p.fri.query_proofs), "queries")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "proof_codec.json"), "w"), indent=1)
