#!/bin/bash
# round-2 NTT experiments: LDE 2^23 x 256 per library variant, plus a per-pass trace (B200ZK_NTT_TRACE=1 serialises the passes)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  echo "=== $lib"
  B200ZK_LIB_PATH=$lib python tools/ntt_bench.py 22x256 23x256 2>&1 | grep -v "^\[ntt\]"
  B200ZK_LIB_PATH=$lib B200ZK_NTT_TRACE=1 python tools/ntt_bench.py 23x256 2>&1 | grep "^\[ntt\]" | tail -9
done
