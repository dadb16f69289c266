#!/bin/bash
# round-2 GPU call 21: search with sampled block maxima -> two thresholds -> full rescan with two-sided candidate lists ->
# front-only final selection; A/B against the round-2 head build on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late.py -x -q -k "search or topk or dres or shard or select" ) > gpurun_out/r2_21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_21_pytest.log
( timeout 300 tests/native/native_tests ) > gpurun_out/r2_21_native.log 2>&1
echo "native rc=$?" >> gpurun_out/r2_21_native.log
( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768,1250000x768,1250000x4096,125000x4096,1000000x2048 ) > gpurun_out/r2_21_phases_new.jsonl 2> gpurun_out/r2_21_phases_new.err
( SGPT_B200_LIB=$PWD/build/base/libsgpt_b200_r2head.so timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768,1250000x768,1250000x4096,125000x4096,1000000x2048 ) > gpurun_out/r2_21_phases_base.jsonl 2> gpurun_out/r2_21_phases_base.err
( SGPT_SEARCH_K_HI=1001 timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 ) > gpurun_out/r2_21_phases_new_nosplit.jsonl 2>> gpurun_out/r2_21_phases_new.err
B="python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( timeout 600 $B ) > gpurun_out/r2_21_bench.json 2> gpurun_out/r2_21_bench.err
tail -3 gpurun_out/r2_21_pytest.log; tail -3 gpurun_out/r2_21_native.log; cat gpurun_out/r2_21_phases_*.jsonl | cut -c1-330; cut -c1-1500 gpurun_out/r2_21_bench.json
