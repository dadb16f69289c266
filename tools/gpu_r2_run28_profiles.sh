#!/bin/bash
# round-2 GPU call 28 (run from call 29 / 31): evidence for profiles/ with the final kernels — launch list + DRAM traffic of a bench window, full ncu
# captures of the top kernels (GEMM with gelu epilogue, residual GEMM, attention, LayerNorm, both search scans, selections)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 2 --no-other-configs --no-corpus-10m --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 440 --csv --log-file gpurun_out/r2_28_launches.csv $B > gpurun_out/r2_28_launches_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 600 -c 440 --csv --log-file gpurun_out/r2_28_traffic.csv $B > gpurun_out/r2_28_traffic_bench.log 2>&1
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
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*OpTmaBiasActBF16.*1' -s 24 -c 1 -o gpurun_out/r2_28_gemm_gelu python /tmp/enc125.py > gpurun_out/r2_28_ncu_a.log 2>&1
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*OpTmaResidAddBF16' -s 48 -c 1 -o gpurun_out/r2_28_gemm_resid python /tmp/enc125.py > gpurun_out/r2_28_ncu_b.log 2>&1
timeout 600 $N -k regex:attention_tc -s 24 -c 1 -o gpurun_out/r2_28_attn_single python /tmp/enc125.py > gpurun_out/r2_28_ncu_c.log 2>&1
timeout 600 $N -k regex:layernorm_bf16 -s 48 -c 1 -o gpurun_out/r2_28_layernorm python /tmp/enc125.py > gpurun_out/r2_28_ncu_d.log 2>&1
# the two launches of the similarity GEMM of one search: sample pass (even index), filter pass (odd index)
timeout 600 $N -k regex:'gemm_bf16_tn_kernel.*EpiFilterRows' -s 4 -c 2 -o gpurun_out/r2_28_simscan python /tmp/enc125.py > gpurun_out/r2_28_ncu_e.log 2>&1
timeout 600 $N -k regex:'topk_select|tau_select' -s 4 -c 2 -o gpurun_out/r2_28_select python /tmp/enc125.py > gpurun_out/r2_28_ncu_f.log 2>&1
ls -la gpurun_out/*.ncu-rep | head -12; tail -2 gpurun_out/r2_28_ncu_e.log
