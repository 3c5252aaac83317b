#!/bin/bash
# Run ON THE GPU BOX (under gpurun): captures the per-round evidence the bench numbers are judged against.
#   1. launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised: compare SHARES)
#   2. one `--set full` capture of the dominant kernel (leaf hash) and of the NTT passes at bench size
# Usage: tools/capture_profiles.sh r01
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-grid > gpurun_out/bench_under_ncu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"leaf_hash" -s 3 -c 1 -f -o gpurun_out/leaf_hash_$R \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-grid > /dev/null 2>&1
# one whole LDE of the timed step: 2 inverse passes, the fused middle, 2 x 2 forward passes (7 launches per step since r02)
ncu --set full --clock-control none --import-source on -k regex:"pass_kernel|lde_mid" -s 21 -c 7 -f -o gpurun_out/ntt_passes_$R \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-grid > /dev/null 2>&1
ls -la gpurun_out/*_$R*
