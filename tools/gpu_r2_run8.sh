#!/bin/bash
# round-2 GPU call 8: bf16 residual stream experiment (parity incl. full depth, speed), restored v1 attention / select loader
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_lnfold.py -x -q ) > gpurun_out/r2_8_lnfold.log 2>&1
echo "lnfold rc=$?" >> gpurun_out/r2_8_lnfold.log
( time SGPT_RESID_BF16=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py tests/test_gpu_cross_encoder.py tests/test_gpu_heads.py -q ) > gpurun_out/r2_8_pytest_bf16resid.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_8_pytest_bf16resid.log
( SGPT_RESID_BF16=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_8_bench_bf16resid.json 2> gpurun_out/r2_8_bench_bf16resid.err
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_8_bench_f32resid.json 2> gpurun_out/r2_8_bench_f32resid.err
( SGPT_RESID_BF16=1 timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_8_models_bf16resid.jsonl 2> gpurun_out/r2_8_models.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_8_models_f32resid.jsonl 2>> gpurun_out/r2_8_models.err
