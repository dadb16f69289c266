#!/bin/bash
# round-2 GPU call 1: full GPU test suite (incl. full-depth configs 3-5), new bench line, ncu of the radix select
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_1_gpu.txt 2>&1
free -g >> gpurun_out/r2_1_gpu.txt; nproc >> gpurun_out/r2_1_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2_1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_1_pytest.log
( time timeout 900 python bench.py --steps 30 --warmup 3 ) > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err
echo "bench rc=$?" >> gpurun_out/r2_1_bench.err
cat > /tmp/one_search.py << 'PY'
import torch, sys
sys.path.insert(0, ".")
from sgpt_b200 import CorpusShard
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
n, D = 1_000_000, 768
sh = CorpusShard(D, n, device=dev)
for s0 in range(0, n, 250_000):
    sh.add(torch.randn(250_000, D, generator=g, device=dev))
q = torch.randn(128, D, generator=g, device=dev)
for _ in range(3):
    sh.search(q, 1001, "cos_sim")
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_select -s 2 -c 2 -o gpurun_out/r2_1_topk python /tmp/one_search.py > gpurun_out/r2_1_ncu_topk.log 2>&1
echo "done" >> gpurun_out/r2_1_gpu.txt
