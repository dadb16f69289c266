#!/bin/bash
# round-2 GPU call 15: ws attention v3 (v1 softmax + single-tile 4-slot mode + conditional rescale): parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time SGPT_ATTN_WS_SINGLE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py -x -q ) > gpurun_out/r2_15_pytest_wssingle.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_15_pytest_wssingle.log
( SGPT_ATTN_WS_SINGLE=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_15_bench_wssingle.json 2> gpurun_out/r2_15_bench_wssingle.err
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_15_bench_default.json 2> gpurun_out/r2_15_bench_default.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_15_models.jsonl 2> gpurun_out/r2_15_models.err
( SGPT_ATTN_IMPL=legacy timeout 600 python tools/bench_models.py --steps 5 --models sgpt-1.3b,sgpt-bloom-7b1 ) > gpurun_out/r2_15_models_legacy.jsonl 2>> gpurun_out/r2_15_models.err
