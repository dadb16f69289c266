#!/bin/bash
# round-2 GPU call 4 (2 GPUs): NCCL + peer-gather correctness test, 2-rank bench with merge verification
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_4_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/r2_4_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2_4_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2_4_bench2.json 2> gpurun_out/r2_4_bench2.err
echo "bench2 rc=$?" >> gpurun_out/r2_4_bench2.err
( SGPT_BENCH_TRANSPORT=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-corpus-10m ) > gpurun_out/r2_4_bench2_nccl.json 2> gpurun_out/r2_4_bench2_nccl.err
