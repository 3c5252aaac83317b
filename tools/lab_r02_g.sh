#!/bin/bash
# round-2 lab G: ncu --set full of the three LDE kernels (one launch each) on LDE 2^22 x 256
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in lde_mid_kernel pass_kernel_direct pass_kernel_tma; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/ntt_r02_$k -f python tools/ntt_bench.py 22x256 > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
