#!/bin/bash
# round-2 lab E (run on the GPU box): IADD3 operand forms and which pipe they execute on; tile-shape copy ceilings
cd "$(dirname "$0")"
mkdir -p ../gpurun_out
M=sm__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum
{
  echo "=== iadd3_forms"; bin/iadd3_forms
  echo "=== tile_copy_lab"; bin/tile_copy_lab 23
} > ../gpurun_out/lab_r02_e.txt 2>&1
ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/iadd3_forms_pipes.csv -k regex:k -c 18 bin/iadd3_forms > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file ../gpurun_out/p2_base_pipes.csv -c 2 bin/p2r3/t3_a0 > /dev/null 2>&1
cat ../gpurun_out/lab_r02_e.txt
