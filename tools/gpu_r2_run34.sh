#!/bin/bash
# round-2 GPU call 34: software-pipelined box loop in the TMA epilogue of the linear GEMMs (next tcgen05.ld in flight under the
# stores of the current box) — a second build of the library (build/exp/) against the in-tree one on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
EXP=$PWD/build/exp/libsgpt_b200_pipelined_epilogue.so
( SGPT_B200_LIB=$EXP timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_splitk.py tests/test_gpu_parity.py tests/test_gpu_lnfold.py -x -q ) > gpurun_out/r2_34_pytest_exp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_34_pytest_exp.log
B="python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline"
( timeout 600 $B ) > gpurun_out/r2_34_bench_base.json 2> gpurun_out/r2_34_bench_base.err
( SGPT_B200_LIB=$EXP timeout 600 $B ) > gpurun_out/r2_34_bench_exp.json 2> gpurun_out/r2_34_bench_exp.err
( timeout 600 $B ) > gpurun_out/r2_34_bench_base2.json 2>> gpurun_out/r2_34_bench_base.err
( SGPT_B200_LIB=$EXP timeout 600 $B ) > gpurun_out/r2_34_bench_exp2.json 2>> gpurun_out/r2_34_bench_exp.err
( SGPT_B200_LIB=$EXP timeout 600 python tools/bench_models.py --steps 5 --models sgpt-1.3b,sgpt-5.8b ) > gpurun_out/r2_34_models_exp.jsonl 2> gpurun_out/r2_34_models.err
( timeout 600 python tools/bench_models.py --steps 5 --models sgpt-1.3b,sgpt-5.8b ) > gpurun_out/r2_34_models_base.jsonl 2>> gpurun_out/r2_34_models.err
tail -3 gpurun_out/r2_34_pytest_exp.log
for f in base exp base2 exp2; do python - <<PY
import json
d=json.loads(open("gpurun_out/r2_34_bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items() if k in ("linear_gemm","attention","layernorm")}, round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
PY
done
cut -c1-330 gpurun_out/r2_34_models_base.jsonl gpurun_out/r2_34_models_exp.jsonl
