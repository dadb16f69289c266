#!/bin/bash
# round-2 GPU call 30: the round's final single-GPU validation — smoke(), the whole GPU test suite, the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_30_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_30_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2_30_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_30_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/r2_30_bench.json 2> gpurun_out/r2_30_bench.err
echo "bench rc=$?" >> gpurun_out/r2_30_bench.err
tail -2 gpurun_out/r2_30_smoke.log; tail -14 gpurun_out/r2_30_pytest.log; tail -4 gpurun_out/r2_30_bench.err; cut -c1-900 gpurun_out/r2_30_bench.json
