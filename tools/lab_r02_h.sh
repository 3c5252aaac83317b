#!/bin/bash
# round-2 lab H: fused LDE middle variants (tile size / K of the middle)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- $*"; env "$@" python tools/ntt_bench.py 22x256 23x256 2>&1 | grep -v "^\[ntt\]"; env "$@" B200ZK_NTT_TRACE=1 python tools/ntt_bench.py 23x256 2>&1 | grep "^\[ntt\]" | tail -7; }
{
  run X=0
  run B200ZK_MID_TILE=12
  run B200ZK_MID_K=8
  run B200ZK_MID_K=6 B200ZK_MID_TILE=12
} > gpurun_out/lab_r02_h.txt 2>&1
cat gpurun_out/lab_r02_h.txt
