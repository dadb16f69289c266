#!/bin/bash
# round-2 GPU call 25: banded tile rasterisation of the linear-layer GEMMs (weights larger than the L2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_splitk.py -x -q ) > gpurun_out/r2_25_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_25_pytest.log
( timeout 600 python tools/bench_band.py --iters 10 ) > gpurun_out/r2_25_band.jsonl 2> gpurun_out/r2_25_band.err
( timeout 600 python tools/bench_models.py --steps 5 --models sgpt-5.8b,sgpt-bloom-7b1 ) > gpurun_out/r2_25_models_auto.jsonl 2> gpurun_out/r2_25_models.err
( SGPT_GEMM_BAND=0 timeout 600 python tools/bench_models.py --steps 5 --models sgpt-5.8b,sgpt-bloom-7b1 ) > gpurun_out/r2_25_models_band0.jsonl 2>> gpurun_out/r2_25_models.err
tail -3 gpurun_out/r2_25_pytest.log; cut -c1-420 gpurun_out/r2_25_band.jsonl; cut -c1-420 gpurun_out/r2_25_models_auto.jsonl gpurun_out/r2_25_models_band0.jsonl; tail -2 gpurun_out/r2_25_band.err
