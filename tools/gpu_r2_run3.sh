#!/bin/bash
# round-2 GPU call 3: GEMM epilogue microbench (L2 prefetch on/off), new attention + select kernels through the test suite, bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 300 python tools/bench_gemm_epilogues.py --model 125m ) > gpurun_out/r2_3_epi_125m_pf1.jsonl 2>&1
( SGPT_RESID_L2_PREFETCH=0 timeout 300 python tools/bench_gemm_epilogues.py --model 125m ) > gpurun_out/r2_3_epi_125m_pf0.jsonl 2>&1
( timeout 300 python tools/bench_gemm_epilogues.py --model 1.3b --iters 10 ) > gpurun_out/r2_3_epi_1.3b.jsonl 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2_3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_3_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_3_bench_new.json 2> gpurun_out/r2_3_bench_new.err
( SGPT_ATTN_IMPL=legacy timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_3_bench_legacyattn.json 2> gpurun_out/r2_3_bench_legacyattn.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_3_models.jsonl 2> gpurun_out/r2_3_models.err
