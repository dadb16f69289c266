#!/bin/bash
# round-2 GPU call 32 (2 GPUs): cross-GPU exchange with the final search kernels (two-sided lists, rank-merge, pushes from
# the selection kernel): NCCL + peer-gather correctness tests, 2-rank bench with merge verification and the 10M legs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2_32_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2_32_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2_32_bench2.json 2> gpurun_out/r2_32_bench2.err
echo "bench2 rc=$?" >> gpurun_out/r2_32_bench2.err
tail -4 gpurun_out/r2_32_multi.log; tail -3 gpurun_out/r2_32_bench2.err; cut -c1-500 gpurun_out/r2_32_bench2.json
