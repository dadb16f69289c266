#!/bin/bash
# round-2 GPU call 24: is the scan bound by the L2->SM operand stream?  Fewer queries = fewer A-operand bytes per tile
# (rows beyond nq are zero-filled by TMA without being fetched), everything else unchanged; idle vs post-GEMM clock state
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for pre in none burn; do
  for nq in 128 64 8; do
    ( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 --pre $pre --queries $nq ) >> gpurun_out/r2_24_phases.jsonl 2>> gpurun_out/r2_24_phases.err
  done
done
cat gpurun_out/r2_24_phases.jsonl | cut -c1-330; tail -3 gpurun_out/r2_24_phases.err
