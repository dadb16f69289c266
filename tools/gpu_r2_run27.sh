#!/bin/bash
# round-2 GPU call 27: CTA-pair scan (M = 256 queries per pass) for query batches above 128
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late.py tests/test_gpu_multi.py tests/test_gpu_raster.py -x -q -k "search or topk or dres or shard or select or merge or packed or tile_order" ) > gpurun_out/r2_27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_27_pytest.log
( timeout 300 tests/native/native_tests ) > gpurun_out/r2_27_native.log 2>&1
echo "native rc=$?" >> gpurun_out/r2_27_native.log
for nq in 128 256 1024; do
  ( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 --queries $nq ) >> gpurun_out/r2_27_phases.jsonl 2>> gpurun_out/r2_27_phases.err
done
( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x2048,1250000x4096 --queries 256 ) >> gpurun_out/r2_27_phases.jsonl 2>> gpurun_out/r2_27_phases.err
tail -3 gpurun_out/r2_27_pytest.log; tail -2 gpurun_out/r2_27_native.log; cut -c1-420 gpurun_out/r2_27_phases.jsonl; tail -3 gpurun_out/r2_27_phases.err
