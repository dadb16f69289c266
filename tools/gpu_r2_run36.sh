#!/bin/bash
# round-2 GPU call 36: the 16-warp early-release TMA epilogue (SGPT_GEMM_EPI16=1) — bit-identity tests, whole-model parity
# with it switched on, alternating bench runs on one box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_raster.py -x -q ) > gpurun_out/r2_36_pytest_raster.log 2>&1
echo "raster rc=$?" >> gpurun_out/r2_36_pytest_raster.log
( SGPT_GEMM_EPI16=1 timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "not search and not dres and not semantic" ) > gpurun_out/r2_36_pytest_parity_epi16.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_36_pytest_parity_epi16.log
B="python bench.py --steps 20 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( SGPT_GEMM_EPI16=1 timeout 100 $B ) > gpurun_out/r2_36_bench_epi16.json 2> gpurun_out/r2_36_bench.err
( timeout 100 $B ) > gpurun_out/r2_36_bench_base.json 2>> gpurun_out/r2_36_bench.err
( SGPT_GEMM_EPI16=1 timeout 100 $B ) > gpurun_out/r2_36_bench_epi16_2.json 2>> gpurun_out/r2_36_bench.err
tail -2 gpurun_out/r2_36_pytest_raster.log; tail -2 gpurun_out/r2_36_pytest_parity_epi16.log
for f in epi16 base epi16_2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_36_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items() if k in ("linear_gemm","attention","layernorm")}, round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "failed", e)
PY
done
tail -3 gpurun_out/r2_36_bench.err
