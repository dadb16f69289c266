#!/bin/bash
# round-2 GPU call 17: persistent LayerNorm (gamma/beta in registers) + split-K residual GEMMs: parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2_17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_17_pytest.log
( timeout 600 python tools/bench_splitk.py --iters 20 ) > gpurun_out/r2_17_splitk.jsonl 2> gpurun_out/r2_17_splitk.err
B="python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( timeout 600 $B ) > gpurun_out/r2_17_bench_default.json 2> gpurun_out/r2_17_bench_default.err
( SGPT_LN_PERSIST=0 timeout 600 $B ) > gpurun_out/r2_17_bench_ln_onerow.json 2> gpurun_out/r2_17_bench_ln_onerow.err
( SGPT_GEMM_SPLITK=2 timeout 600 $B ) > gpurun_out/r2_17_bench_splitk2.json 2> gpurun_out/r2_17_bench_splitk2.err
( SGPT_GEMM_SPLITK=3 timeout 600 $B ) > gpurun_out/r2_17_bench_splitk3.json 2> gpurun_out/r2_17_bench_splitk3.err
( timeout 600 $B ) > gpurun_out/r2_17_bench_default_again.json 2> gpurun_out/r2_17_bench_default_again.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_17_models_default.jsonl 2> gpurun_out/r2_17_models.err
( SGPT_LN_PERSIST=0 SGPT_GEMM_SPLITK=0 timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_17_models_old.jsonl 2>> gpurun_out/r2_17_models.err
( SGPT_GEMM_SPLITK=2 timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_17_models_splitk2.jsonl 2>> gpurun_out/r2_17_models.err
tail -4 gpurun_out/r2_17_pytest.log
