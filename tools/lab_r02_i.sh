#!/bin/bash
# round-2 lab I: host pipeline -- parity of the async entry point, then the bench line (serial + streamed e2e)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  python -m pytest tests -x -q -m gpu -k "host or strip or e2e or headline" 2>&1 | tail -3
  python bench.py --steps 5 --warmup 3 --no-grid --no-cpu
} > gpurun_out/lab_r02_i.txt 2>&1
cat gpurun_out/lab_r02_i.txt
