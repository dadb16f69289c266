#!/bin/bash
# round-2 GPU call 2: LayerNorm-fold unit tests + whole suite + bench (no extra legs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_lnfold.py -x -q ) > gpurun_out/r2_2_lnfold.log 2>&1
echo "lnfold rc=$?" >> gpurun_out/r2_2_lnfold.log
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2_2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_2_pytest.log
( time timeout 900 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m ) > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_2_bench.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_2_models.jsonl 2> gpurun_out/r2_2_models.err
