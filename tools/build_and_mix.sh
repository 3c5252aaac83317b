#!/bin/bash
# build libb200zk.so and print the SASS opcode mix of one kernel (default: compress_top = one inlined permutation)
set -e
R=/root/repo
K=${1:-compress_top}
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -Xcompiler -fPIC -shared -o $R/zkvm_prover_b200/libb200zk.so $R/zkvm_prover_b200/csrc/b200zk.cu 2>&1 | grep -E "error|warning|leaf_hash|pass_kernelILi4|compress_top" -A2 | grep -E "error|warning|Used|Compiling" | head -8
cuobjdump -sass $R/zkvm_prover_b200/libb200zk.so | awk -v k="$K" '/Function : /{f=($0 ~ k)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sort | uniq -c | sort -rn | head -${2:-16}
