#!/bin/bash
# round-2 Poseidon2 sweep B: how many additions to steer to the FMA pipe (non-fused products).  ncu on the shipped setting:
# fmaheavy 92 % busy, alu 63 % -> the FMA pipe carries too many of the additions.
# build: tools/p2_sweep_r02b.sh build (here)     run: tools/p2_sweep_r02b.sh run (GPU box)
cd "$(dirname "$0")"
OUT=bin/p2r2b
CFGS=()
for mds in 0x00 0x01 0x02 0x03 0x04 0x08 0x0F; do
  for isum in 0 2 3; do
    for ilin in 0 1; do
      CFGS+=("m${mds}_s${isum}_l${ilin}:-DP2_MDS_FMA_MASK=${mds} -DP2_INT_SUM_FMA=${isum} -DP2_INT_LIN_FMA=${ilin}")
    done
  done
done
CFGS+=("m0x00_s3_l2_o6:-DP2_MDS_FMA_MASK=0x00 -DP2_INT_SUM_FMA=3 -DP2_INT_LIN_FMA=2 -DP2_INT_OUT_FMA=6")
CFGS+=("m0x00_s3_l1_o3:-DP2_MDS_FMA_MASK=0x00 -DP2_INT_SUM_FMA=3 -DP2_INT_LIN_FMA=1 -DP2_INT_OUT_FMA=3")
CFGS+=("m0x01_s3_l1_o3:-DP2_MDS_FMA_MASK=0x01 -DP2_INT_SUM_FMA=3 -DP2_INT_LIN_FMA=1 -DP2_INT_OUT_FMA=3")
CFGS+=("m0x00_s3_l1_i0:-DP2_MDS_FMA_MASK=0x00 -DP2_INT_SUM_FMA=3 -DP2_INT_LIN_FMA=1 -DP2_INT_MODE=0")
if [ "$1" = build ]; then
  mkdir -p $OUT
  i=0
  for c in "${CFGS[@]}"; do
    name=${c%%:*}; flags=${c#*:}
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DVARIANT=1 $flags -o $OUT/t3_${name} p2_tune3.cu &
    i=$((i+1)); if [ $((i % 14)) = 0 ]; then wait; fi
  done
  wait
  ls $OUT | wc -l
else
  for c in "${CFGS[@]}"; do
    name=${c%%:*}
    printf "%-18s " $name; $OUT/t3_${name}
  done
fi
