#!/bin/bash
# round-2 GPU call 29: front-list selection kernel + branch-free filter epilogue (tests, phase timeline, A/B by env switch),
# then the profile captures (tools/gpu_r2_run28_profiles.sh)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late.py tests/test_gpu_multi.py -x -q -k "search or topk or dres or shard or select or merge or packed" ) > gpurun_out/r2_29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_29_pytest.log
( timeout 300 tests/native/native_tests ) > gpurun_out/r2_29_native.log 2>&1
echo "native rc=$?" >> gpurun_out/r2_29_native.log
( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768,1250000x768,125000x4096 ) > gpurun_out/r2_29_phases.jsonl 2> gpurun_out/r2_29_phases.err
( SGPT_FRONT_SELECT=0 timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 ) >> gpurun_out/r2_29_phases.jsonl 2>> gpurun_out/r2_29_phases.err
( timeout 600 python tools/search_phases.py --scores cos_sim --shapes 1000000x768 --queries 256 ) >> gpurun_out/r2_29_phases.jsonl 2>> gpurun_out/r2_29_phases.err
tail -3 gpurun_out/r2_29_pytest.log; tail -2 gpurun_out/r2_29_native.log; cut -c1-420 gpurun_out/r2_29_phases.jsonl
bash tools/gpu_r2_run28_profiles.sh
