#!/bin/bash
# round-2 GPU call 33 (8 GPUs): final bench line at N = 8 with merge verification and the 10M strong-scaling legs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2_33_bench8.json 2> gpurun_out/r2_33_bench8.err
echo "bench8 rc=$?" >> gpurun_out/r2_33_bench8.err
tail -3 gpurun_out/r2_33_bench8.err; cut -c1-400 gpurun_out/r2_33_bench8.json
