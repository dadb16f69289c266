#!/bin/bash
# round-2 GPU call 20: filter epilogue with per-tile constants prefetched before the accumulator wait (A/B against the
# previous build on the same box), per-phase search times for cos_sim and dot
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_late.py -x -q -k "search or topk or dres or shard or select" ) > gpurun_out/r2_20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_20_pytest.log
( timeout 600 python tools/search_phases.py ) > gpurun_out/r2_20_phases_new.jsonl 2> gpurun_out/r2_20_phases_new.err
( SGPT_B200_LIB=$PWD/build/base/libsgpt_b200_r2head.so timeout 600 python tools/search_phases.py ) > gpurun_out/r2_20_phases_base.jsonl 2> gpurun_out/r2_20_phases_base.err
( timeout 600 python tools/search_phases.py --shapes 1250000x4096,125000x4096 --scores cos_sim ) > gpurun_out/r2_20_phases_new_4096.jsonl 2>> gpurun_out/r2_20_phases_new.err
( SGPT_B200_LIB=$PWD/build/base/libsgpt_b200_r2head.so timeout 600 python tools/search_phases.py --shapes 1250000x4096,125000x4096 --scores cos_sim ) > gpurun_out/r2_20_phases_base_4096.jsonl 2>> gpurun_out/r2_20_phases_base.err
tail -3 gpurun_out/r2_20_pytest.log; cat gpurun_out/r2_20_phases_*.jsonl | cut -c1-330
