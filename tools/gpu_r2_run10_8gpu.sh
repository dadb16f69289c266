#!/bin/bash
# round-2 GPU call 10 (8 GPUs): bench at N=8 and N=4 with merge verification and the 10M strong-scaling legs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2_10_bench8.json 2> gpurun_out/r2_10_bench8.err
echo "bench8 rc=$?" >> gpurun_out/r2_10_bench8.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 ) > gpurun_out/r2_10_bench4.json 2> gpurun_out/r2_10_bench4.err
echo "bench4 rc=$?" >> gpurun_out/r2_10_bench4.err
