#!/bin/bash
# round-2 GPU call 19: tile width 128 vs 256 on every encoder GEMM shape (is the last, partly filled round what costs?)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python tools/bench_bn.py --iters 20 ) > gpurun_out/r2_19_bn.jsonl 2> gpurun_out/r2_19_bn.err
cat gpurun_out/r2_19_bn.jsonl
