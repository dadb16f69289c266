#!/bin/bash
# round-2 GPU call 9: bf16 residual default + dedicated bf16 LayerNorm, radix-4 select: full suite, fp32-stream subset, full bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -s --durations=5 ) > gpurun_out/r2_9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_9_pytest.log
( time SGPT_RESID_BF16=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py -q -s ) > gpurun_out/r2_9_pytest_f32resid.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_9_pytest_f32resid.log
( time timeout 900 python bench.py --steps 30 --warmup 3 ) > gpurun_out/r2_9_bench.json 2> gpurun_out/r2_9_bench.err
echo "bench rc=$?" >> gpurun_out/r2_9_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_9_smoke.log 2>&1
