#!/bin/bash
# round-2 GPU call 6: ncu source-level captures of the fc GEMM (plain gelu vs lnfold gelu) + new attention kernel tests/bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cat > /tmp/one_gemm.py << 'PY'
import sys, torch
sys.path.insert(0, ".")
from sgpt_b200 import _lib
L, lib = _lib, _lib.lib()
M, d, ff = 32768, 768, 3072
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16
x = torch.randn(M, d, generator=g, device=dev).to(bf)
w = (torch.randn(ff, d, generator=g, device=dev) * 0.02).to(bf)
bias = torch.randn(ff, generator=g, device=dev) * 0.1
cs = torch.randn(ff, generator=g, device=dev)
resid = torch.randn(M, d, generator=g, device=dev)
P = (d + 127) // 128
stats = torch.zeros(M, P, 2, device=dev)
xb = torch.empty(M, d, dtype=bf, device=dev)
out = torch.empty(M, ff, dtype=bf, device=dev)
st = L.current_stream()
L.check(lib.sgpt_resid_stats(resid.data_ptr(), xb.data_ptr(), stats.data_ptr(), M, d, st))
for _ in range(3):
    L.check(lib.sgpt_linear(x.data_ptr(), d, w.data_ptr(), d, bias.data_ptr(), out.data_ptr(), ff, None, M, ff, d, 1, st))
    L.check(lib.sgpt_linear_lnfold(x.data_ptr(), d, w.data_ptr(), d, bias.data_ptr(), cs.data_ptr(), stats.data_ptr(), P, 1e-5,
                                   out.data_ptr(), ff, M, ff, d, 1, st))
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 2 -c 2 -o gpurun_out/r2_6_gemm python /tmp/one_gemm.py > gpurun_out/r2_6_ncu_gemm.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_depth.py -x -q ) > gpurun_out/r2_6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_6_pytest.log
( timeout 600 python bench.py --steps 30 --warmup 3 --no-other-configs --no-corpus-10m --no-cpu-baseline ) > gpurun_out/r2_6_bench_new.json 2> gpurun_out/r2_6_bench_new.err
( timeout 600 python tools/bench_models.py --steps 5 ) > gpurun_out/r2_6_models.jsonl 2> gpurun_out/r2_6_models.err
