#!/bin/bash
# round-2 GPU call 23: rank-merge kernel for sorted packed lists, one-round list lengths in the selection; what the state
# an encode step leaves behind costs the search (dirty L2 vs power state); encoder sub-batch sizes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late.py tests/test_gpu_multi.py -x -q -k "search or topk or dres or shard or select or merge or packed" ) > gpurun_out/r2_23_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_23_pytest.log
for pre in none dirty burn; do
  ( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 --pre $pre ) >> gpurun_out/r2_23_phases.jsonl 2>> gpurun_out/r2_23_phases.err
done
for b in 256 128 64; do
  ( timeout 600 python tools/bench_models.py --models sgpt-125m --batch $b --steps 150 --no-profile ) >> gpurun_out/r2_23_subbatch.jsonl 2>> gpurun_out/r2_23_subbatch.err
done
tail -3 gpurun_out/r2_23_pytest.log; cat gpurun_out/r2_23_phases.jsonl | cut -c1-420; cut -c1-300 gpurun_out/r2_23_subbatch.jsonl; tail -3 gpurun_out/r2_23_subbatch.err
