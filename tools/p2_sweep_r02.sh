#!/bin/bash
# round-2 Poseidon2 sweep: fused Montgomery products (bb::smulz) x pipe steering x q-on-ALU lanes.
# build: tools/p2_sweep_r02.sh build   (here, cross-compiles)     run: tools/p2_sweep_r02.sh run   (on the GPU box)
cd "$(dirname "$0")"
OUT=bin/p2r2
CFGS=(
 "base0:-DP2_FUSED=0"
 "f1_m0f:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x0F"
 "f1_m00:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00"
 "f1_m01:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x01"
 "f1_m04:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x04"
 "f1_m00_q1:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_QLEA_MASK=0x1111"
 "f1_m00_q2:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_QLEA_MASK=0x5555"
 "f1_m00_q3:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_QLEA_MASK=0x7777"
 "f1_m00_i1:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_INT_OUT_FMA=3"
 "f1_m00_i2:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_INT_OUT_FMA=6"
 "f1_m00_i3:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_INT_OUT_FMA=6 -DP2_INT_LIN_FMA=2"
 "f1_m00_rc:-DP2_FUSED=1 -DP2_MDS_FMA_MASK=0x00 -DP2_RC_FMA=1"
)
if [ "$1" = build ]; then
  mkdir -p $OUT
  for c in "${CFGS[@]}"; do
    name=${c%%:*}; flags=${c#*:}
    for v in 1 2; do
      nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DVARIANT=$v $flags -o $OUT/t3_${name}_v$v p2_tune3.cu &
    done
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a $flags -o $OUT/t4_${name} p2_tune4.cu &
    wait
  done
  ls $OUT | wc -l
else
  for c in "${CFGS[@]}"; do
    name=${c%%:*}
    for v in 1 2; do printf "%-12s " $name; $OUT/t3_${name}_v$v; done
  done
  for c in "${CFGS[@]}"; do
    name=${c%%:*}; echo "--- phases $name"; $OUT/t4_${name}
  done
fi
