#!/bin/bash
# round-2 lab F: fused LDE middle -- parity suite, LDE timing with per-pass trace, and the Poseidon2 pipe-balance sweep B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
  echo "=== ntt_bench fused"; python tools/ntt_bench.py 20x64 22x256 23x256 24x64 2>&1 | grep -v "^\[ntt\]"
  B200ZK_NTT_TRACE=1 python tools/ntt_bench.py 23x256 2>&1 | grep "^\[ntt\]" | tail -7
  echo "=== ntt_bench unfused"; B200ZK_LDE_FUSED_MID=0 python tools/ntt_bench.py 20x64 22x256 23x256 24x64 2>&1 | grep -v "^\[ntt\]"
  echo "=== p2 sweep B"; tools/p2_sweep_r02b.sh run
} > gpurun_out/lab_r02_f.txt 2>&1
cat gpurun_out/lab_r02_f.txt
