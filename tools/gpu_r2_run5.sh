#!/bin/bash
# round-2 GPU call 5: 4-CTA multicast clusters (CL=4) correctness + microbench A/B, hoisted LN row stats
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_lnfold.py -x -q ) > gpurun_out/r2_5_lnfold.log 2>&1
echo "lnfold rc=$?" >> gpurun_out/r2_5_lnfold.log
( timeout 300 python tools/bench_gemm_epilogues.py --model 125m ) > gpurun_out/r2_5_epi_125m_cl4.jsonl 2>&1
( SGPT_GEMM_CL4=0 timeout 300 python tools/bench_gemm_epilogues.py --model 125m ) > gpurun_out/r2_5_epi_125m_cl2.jsonl 2>&1
( timeout 300 python tools/bench_gemm_epilogues.py --model 1.3b --iters 10 ) > gpurun_out/r2_5_epi_1.3b_cl4.jsonl 2>&1
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/r2_5_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_5_parity.log
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_5_bench_cl4.json 2> gpurun_out/r2_5_bench_cl4.err
( SGPT_GEMM_CL4=0 timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_5_bench_cl2.json 2> gpurun_out/r2_5_bench_cl2.err
