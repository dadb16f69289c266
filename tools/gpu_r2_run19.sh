#!/bin/bash
# round-2 GPU call 19: tile width 128 vs 256 on every encoder GEMM shape; attention early loads + L2 prefetch of the successor
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python tools/bench_bn.py --iters 20 ) > gpurun_out/r2_19_bn.jsonl 2> gpurun_out/r2_19_bn.err
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py -x -q ) > gpurun_out/r2_19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_19_pytest.log
B="python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( timeout 600 $B ) > gpurun_out/r2_19_bench_default.json 2> gpurun_out/r2_19_bench_default.err
( SGPT_ATTN_PREFETCH=0 timeout 600 $B ) > gpurun_out/r2_19_bench_noprefetch.json 2> gpurun_out/r2_19_bench_noprefetch.err
( timeout 600 $B ) > gpurun_out/r2_19_bench_default_again.json 2> gpurun_out/r2_19_bench_default_again.err
( timeout 600 python tools/bench_models.py --steps 5 --models sgpt-5.8b ) > gpurun_out/r2_19_models_default.jsonl 2> gpurun_out/r2_19_models.err
( SGPT_ATTN_PREFETCH=0 timeout 600 python tools/bench_models.py --steps 5 --models sgpt-5.8b ) > gpurun_out/r2_19_models_noprefetch.jsonl 2>> gpurun_out/r2_19_models.err
cat gpurun_out/r2_19_bn.jsonl; tail -3 gpurun_out/r2_19_pytest.log
