#!/bin/bash
# round-2 GPU call 35: last sanity pass over the final in-tree build — smoke(), the search / merge / GEMM-order tests, a short bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_35_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_35_smoke.log
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_raster.py tests/test_gpu_zz_late.py -x -q ) > gpurun_out/r2_35_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_35_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_35_bench.json 2> gpurun_out/r2_35_bench.err
tail -2 gpurun_out/r2_35_smoke.log; tail -3 gpurun_out/r2_35_pytest.log; cut -c1-300 gpurun_out/r2_35_bench.json
