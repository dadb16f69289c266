#!/bin/bash
# round-2 GPU call 7: attention ws v1.1 (64-key tiles at hd 128) + segment-table select loader: tests, model sweep, ncu of select and attention
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py tests/test_gpu_multi.py tests/test_gpu_lnfold.py -x -q ) > gpurun_out/r2_7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_7_pytest.log
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_7_models.jsonl 2> gpurun_out/r2_7_models.err
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_7_bench.json 2> gpurun_out/r2_7_bench.err
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_select -s 2 -c 2 -o gpurun_out/r2_7_topk python /tmp/one_search.py > gpurun_out/r2_7_ncu_topk.log 2>&1
cat > /tmp/one_attn.py << 'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from sgpt_b200 import Encoder, preset
from tools.bench_models import rand_weights
dev = torch.device("cuda:0")
cfg = preset("sgpt-1.3b", n_layer=2)
enc = Encoder(cfg, rand_weights(cfg, dev), device=dev, max_tokens=64 * 256, max_batch=64)
ids = torch.randint(0, cfg.vocab, (64, 256)).numpy()
mask = np.ones((64, 256), dtype=np.int8)
for _ in range(2):
    enc.encode_tokens(ids, mask)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_ws -s 2 -c 1 -o gpurun_out/r2_7_attn python /tmp/one_attn.py > gpurun_out/r2_7_ncu_attn.log 2>&1
