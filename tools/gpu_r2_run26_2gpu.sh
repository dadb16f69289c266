#!/bin/bash
# round-2 GPU call 26 (2 GPUs): cross-GPU exchange with the reworked search (two-sided lists, rank-merge kernel): NCCL and
# peer-gather correctness tests, 2-rank bench with merge verification and the 10M strong-scaling legs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_raster.py -x -q ) > gpurun_out/r2_26_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2_26_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2_26_bench2.json 2> gpurun_out/r2_26_bench2.err
echo "bench2 rc=$?" >> gpurun_out/r2_26_bench2.err
tail -4 gpurun_out/r2_26_multi.log; tail -3 gpurun_out/r2_26_bench2.err; cut -c1-600 gpurun_out/r2_26_bench2.json
