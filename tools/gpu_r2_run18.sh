#!/bin/bash
# round-2 GPU call 18: split-K off by default, persistent LayerNorm, head-major attention grid, one-round sample pass
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2_18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_18_pytest.log
B="python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( timeout 600 $B ) > gpurun_out/r2_18_bench_default.json 2> gpurun_out/r2_18_bench_default.err
( SGPT_ATTN_HEAD_MAJOR=0 timeout 600 $B ) > gpurun_out/r2_18_bench_attn_old_order.json 2> gpurun_out/r2_18_bench_attn_old_order.err
( timeout 600 $B ) > gpurun_out/r2_18_bench_default_again.json 2> gpurun_out/r2_18_bench_default_again.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_18_models_default.jsonl 2> gpurun_out/r2_18_models.err
( SGPT_ATTN_HEAD_MAJOR=0 timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_18_models_attn_old_order.jsonl 2>> gpurun_out/r2_18_models.err
tail -4 gpurun_out/r2_18_pytest.log
