#!/bin/bash
# round-2 GPU call 12: fused ln_f+pool kernel, LayerNorm occupancy tweak: full suite + bench + model sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2_12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_12_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_12_bench.json 2> gpurun_out/r2_12_bench.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_12_models.jsonl 2> gpurun_out/r2_12_models.err
