#!/bin/bash
# round-2 GPU call 22: phase timeline of the final selection kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768,125000x4096 ) > gpurun_out/r2_22_phases.jsonl 2> gpurun_out/r2_22_phases.err
( SGPT_SEARCH_K_HI=1001 timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 ) > gpurun_out/r2_22_phases_nosplit.jsonl 2>> gpurun_out/r2_22_phases.err
cat gpurun_out/r2_22_phases*.jsonl; tail -3 gpurun_out/r2_22_phases.err
