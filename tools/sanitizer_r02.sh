#!/bin/bash
# compute-sanitizer over the small GPU parity tests (SURVEY.md section 5 names the tool).
# memcheck: out-of-bounds / misaligned accesses of every kernel family; racecheck: shared-memory hazards of the NTT, fold and
# Merkle kernels; synccheck: barrier misuse.  Full-size tests are excluded (the tools slow kernels 10-100x).
cd "$(dirname "$0")/.."
SEL='permute_matches or hash_rows_matches or merkle_commit_open_verify or dft_batch_matches or coset_lde_matches or coset_lde_fixture or fold_matrix_matches or commit_phase_matches or challenger_matches or open_phase_primitives or pcs_commit_lde or more_than_128 or dft_natural_order or commit_host_strip or commit_host_async'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -x -m gpu -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|=========" | tail -12
  echo "exit code: ${PIPESTATUS[0]}"
done
