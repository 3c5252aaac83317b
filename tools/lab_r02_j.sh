#!/bin/bash
# round-2 lab J: L2 prefetch look-ahead of the direct pass kernel and of the fused middle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "--- $*"; env "$@" python tools/ntt_bench.py 23x256 2>&1 | grep -v "^\[ntt\]"; env "$@" B200ZK_NTT_TRACE=1 python tools/ntt_bench.py 23x256 2>&1 | grep "^\[ntt\]" | tail -7 | awk '{printf "%s %s K=%s %s ms | ", $2, $4, $5, $7} END{print ""}'; }
{
  for d in 0 148 296 444 592 888 1184; do run B200ZK_DIRECT_PREFETCH=$d; done
  for d in 444 888 1776; do run B200ZK_MID_PREFETCH=$d; done
} > gpurun_out/lab_r02_j.txt 2>&1
cat gpurun_out/lab_r02_j.txt
