#!/bin/bash
# round-2 GPU call 11: evidence for profiles/ — launch list + DRAM traffic of a bench window, full ncu captures of the top kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 2 --no-other-configs --no-corpus-10m --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/r2_11_launches.csv $B > gpurun_out/r2_11_launches_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/r2_11_traffic.csv $B > gpurun_out/r2_11_traffic_bench.log 2>&1
cat > /tmp/enc125.py << 'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
from bench import synthetic_weights, token_batches
from sgpt_b200 import CorpusShard, Encoder, preset
dev = torch.device("cuda:0")
enc = Encoder(preset("sgpt-125m"), synthetic_weights(0), device=dev, max_tokens=256 * 128, max_batch=256)
ids = token_batches(1, 5)[0].numpy()
mask = np.ones((256, 128), dtype=np.int8)
g = torch.Generator(device=dev).manual_seed(7)
sh = CorpusShard(768, 1_000_000, device=dev)
for s0 in range(0, 1_000_000, 250_000):
    sh.add(torch.randn(250_000, 768, generator=g, device=dev))
q = torch.randn(128, 768, generator=g, device=dev)
for _ in range(3):
    enc.encode_tokens(ids, mask)
    sh.search(q, 1001, "cos_sim")
torch.cuda.synchronize()
PY
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*OpTmaBiasActBF16.*1' -s 24 -c 1 -o gpurun_out/r2_11_gemm_gelu python /tmp/enc125.py > gpurun_out/r2_11_ncu_a.log 2>&1
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*OpTmaResidAddBF16' -s 48 -c 1 -o gpurun_out/r2_11_gemm_resid python /tmp/enc125.py > gpurun_out/r2_11_ncu_b.log 2>&1
timeout 600 $N -k regex:attention_tc -s 24 -c 1 -o gpurun_out/r2_11_attn_single python /tmp/enc125.py > gpurun_out/r2_11_ncu_c.log 2>&1
timeout 600 $N -k regex:layernorm_bf16 -s 48 -c 1 -o gpurun_out/r2_11_layernorm python /tmp/enc125.py > gpurun_out/r2_11_ncu_d.log 2>&1
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*EpiFilterRows' -s 5 -c 1 -o gpurun_out/r2_11_simfilter python /tmp/enc125.py > gpurun_out/r2_11_ncu_e.log 2>&1
timeout 600 $N -k regex:topk_select -s 5 -c 1 -o gpurun_out/r2_11_topk python /tmp/enc125.py > gpurun_out/r2_11_ncu_f.log 2>&1
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
timeout 600 $N -k regex:attention_ws -s 2 -c 1 -o gpurun_out/r2_11_attn_ws python /tmp/one_attn.py > gpurun_out/r2_11_ncu_g.log 2>&1
( timeout 900 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err
