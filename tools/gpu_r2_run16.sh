#!/bin/bash
# LayerNorm kernel variants: stand-alone microbenchmark on the three row widths of the configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_16_smi.txt
for shape in "32768 768" "16384 2048" "9600 4096"; do
  timeout 120 tools/ln_bench $shape >> gpurun_out/r2_16_ln_bench.jsonl 2>> gpurun_out/r2_16_ln_bench.err
done
cat gpurun_out/r2_16_ln_bench.jsonl
