"""TEST INFRASTRUCTURE: an independent verifier for the output of TwoAdicFriPcs.open (p3-fri verifier logic restated
with the pure-Python oracle arithmetic).  It replays the transcript, checks every Merkle opening against the
commitments, recomputes the reduced openings from the opened rows, follows the folding chain of every query down to
the final polynomial and checks the proof-of-work witness.  If any stage of the device pipeline (LDE ordering, opened
values, denominators, alpha offsets, fold twiddles, roll-ins, tree layout) were inconsistent, some query would fail."""
from oracle import pyref as R

P = R.P


def canon(a):
    return [R.from_monty(int(x)) for x in a]


def verify_pcs_open(roots, dims, points, opened, proof, ch, log_blowup, log_final_poly_len, pow_bits):
    """roots[r]: commitment (monty); dims[r]: list of (width, lde_height); points[r][i]: list of EF4 (monty);
    opened[r][i][k]: (width, 4) monty; ch: pyref DuplexChallenger in the state the prover's challenger had before open."""
    alpha = ch.sample_ext()
    if "alpha" in proof:  # prover-side debugging aids; a proof decoded from the wire format carries neither
        assert alpha == canon(proof["alpha"]), "alpha diverged"
    commits = [canon(c) for c in proof["commit_phase_commits"]]
    betas = []
    for c in commits:
        ch.observe_slice(c)
        betas.append(ch.sample_ext())
    if "betas" in proof:
        assert betas == [canon(b) for b in proof["betas"]]
    final_poly = [canon(c) for c in proof["final_poly"]]
    for c in final_poly:
        ch.observe_slice(c)
    assert ch.check_witness(pow_bits, proof["pow_witness"]), "proof of work rejected"
    log_max = proof["log_max_height"]
    n_rounds = len(commits)
    assert n_rounds == log_max - log_blowup - log_final_poly_len
    for q in range(len(proof["input_openings"][0])):
        index = ch.sample_bits(log_max)
        if "query_indices" in proof:
            assert index == proof["query_indices"][q], "query index diverged"
        ro, num_reduced = {}, {}
        for r, root in enumerate(roots):
            vals, path = proof["input_openings"][r][q]
            heights = [h for _, h in dims[r]]
            lmh = max(heights).bit_length() - 1
            rows = [canon(v) for v in vals]
            assert R.verify_batch(rows, heights, [canon(p) for p in path], index >> (log_max - lmh), canon(root)), "input opening rejected"
            for i, (w, h) in enumerate(dims[r]):
                lh = h.bit_length() - 1
                x = 31 * pow(R.two_adic_generator(lh), R.bitrev(index >> (log_max - lh), lh), P) % P
                for k, z in enumerate(points[r][i]):
                    zc = canon(z)
                    ys = [canon(y) for y in opened[r][i][k]]
                    s, apw = [0, 0, 0, 0], [1, 0, 0, 0]
                    for c in range(w):
                        s = R.ef_add(s, R.ef_mul(apw, R.ef_sub(ys[c], [rows[i][c], 0, 0, 0])))
                        apw = R.ef_mul(apw, alpha)
                    apo = R.ef_pow(alpha, num_reduced.get(lh, 0))
                    term = R.ef_mul(R.ef_mul(apo, s), R.ef_inv(R.ef_sub(zc, [x, 0, 0, 0])))
                    ro[lh] = R.ef_add(ro.get(lh, [0, 0, 0, 0]), term)
                    num_reduced[lh] = num_reduced.get(lh, 0) + w
        folded = ro[log_max]
        for i in range(n_rounds):
            lfh = log_max - i - 1
            idx_i = index >> i
            pair, path = proof["commit_phase_openings"][i][q]
            if proof.get("sibling_only"):  # p3-fri's CommitPhaseProofStep: only the sibling travels, the Merkle check binds the folded value
                evals = [None, None]
                evals[idx_i & 1] = folded
                evals[1 - (idx_i & 1)] = canon(pair)
            else:
                evals = [canon(pair[0]), canon(pair[1])]
                assert evals[idx_i & 1] == folded, f"query {q}: folded value does not match the committed layer {i}"
            assert R.verify_batch([evals[0] + evals[1]], [1 << lfh], [canon(p) for p in path], idx_i >> 1, commits[i]), "commit-phase opening rejected"
            x = pow(R.two_adic_generator(lfh + 1), R.bitrev(idx_i >> 1, lfh), P)
            e0, e1 = evals
            t = R.ef_mul(R.ef_sub(betas[i], [x, 0, 0, 0]), R.ef_scale(R.ef_sub(e1, e0), R.inv((-2 * x) % P)))
            folded = R.ef_add(e0, t)
            if lfh in ro:
                folded = R.ef_add(folded, ro[lfh])
        lfin = log_blowup + log_final_poly_len
        x = pow(R.two_adic_generator(lfin), R.bitrev(index >> n_rounds, lfin), P)
        ev = [0, 0, 0, 0]
        for coef in reversed(final_poly):
            ev = R.ef_add(R.ef_scale(ev, x), coef)
        assert ev == folded, f"query {q}: final polynomial mismatch"
    return True
