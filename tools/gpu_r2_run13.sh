#!/bin/bash
# round-2 GPU call 13: bulk-copy staged LayerNorm: parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lnfold.py tests/test_gpu_full_depth.py -x -q ) > gpurun_out/r2_13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_13_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_13_bench_bulk.json 2> gpurun_out/r2_13_bench_bulk.err
( SGPT_LN_BULK=0 timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_13_bench_nobulk.json 2> gpurun_out/r2_13_bench_nobulk.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_13_models_bulk.jsonl 2> gpurun_out/r2_13_models.err
