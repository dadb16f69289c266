"""Unit parity of the kernels that replace the stand-alone LayerNorm passes (include/sgpt_b200.h "LayerNorm without a
pass of its own"), each against fp64 torch arithmetic on the CPU from the same inputs, called through the C ABI:
sgpt_resid_stats, sgpt_fold_layernorm + sgpt_linear_lnfold (== act(LayerNorm(x) W^T + b), HF:gpt_neo:332-345 + :84-87 /
:304-305), sgpt_linear_resid_ln (== x W^T + b + residual, HF:gpt_neo:342,348, plus the statistics of the result)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from sgpt_b200 import _lib

    return _lib, _lib.lib()


def _gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


@pytest.mark.parametrize("T,d", [(5, 128), (77, 192), (1000, 768), (300, 2048)])
def test_resid_stats(T, d):
    L, lib = _lib()
    g = torch.Generator().manual_seed(T + d)
    x = (torch.randn(T, d, generator=g) * 3 + 0.7)
    xd = x.cuda()
    P = (d + 127) // 128
    xb = torch.zeros(T, d, dtype=torch.bfloat16, device="cuda")
    st = torch.full((T, P, 2), float("nan"), device="cuda")
    L.check(lib.sgpt_resid_stats(xd.data_ptr(), xb.data_ptr(), st.data_ptr(), T, d, L.current_stream()))
    assert torch.equal(xb.cpu(), x.to(torch.bfloat16))
    pad = torch.nn.functional.pad(x.double(), (0, P * 128 - d)).view(T, P, 128)
    assert torch.allclose(st[..., 0].cpu().double(), pad.sum(-1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st[..., 1].cpu().double(), (pad * pad).sum(-1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,K,N,gelu", [(77, 128, 384, 0), (300, 768, 2304, 0), (1000, 768, 3072, 1), (130, 192, 576, 1)])
def test_linear_lnfold_equals_layernorm_then_linear(M, K, N, gelu):
    L, lib = _lib()
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g) * 2.0 + 0.5          # residual stream rows (non-zero mean)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    gamma, beta = torch.randn(K, generator=g) * 0.1 + 1.0, torch.randn(K, generator=g) * 0.05
    bias = torch.randn(N, generator=g) * 0.1
    eps = 1e-5
    dev = "cuda"
    wf = torch.empty(N, K, dtype=torch.bfloat16, device=dev)
    cs, bf = torch.empty(N, device=dev), torch.empty(N, device=dev)
    wd, gd, bd, biasd = w.to(dev), gamma.to(dev), beta.to(dev), bias.to(dev)
    L.check(lib.sgpt_fold_layernorm(wd.data_ptr(), gd.data_ptr(), bd.data_ptr(), biasd.data_ptr(), wf.data_ptr(),
                                    cs.data_ptr(), bf.data_ptr(), N, K, L.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(wf.cpu(), (w.float() * gamma).to(torch.bfloat16))
    assert torch.allclose(cs.cpu().double(), wf.cpu().double().sum(1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(bf.cpu().double(), bias.double() + w.double() @ beta.double(), rtol=1e-5, atol=1e-4)
    P = (K + 127) // 128
    xd = x.to(dev)
    xb = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
    st = torch.empty(M, P, 2, device=dev)
    L.check(lib.sgpt_resid_stats(xd.data_ptr(), xb.data_ptr(), st.data_ptr(), M, K, L.current_stream()))
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    L.check(lib.sgpt_linear_lnfold(xb.data_ptr(), K, wf.data_ptr(), K, bf.data_ptr(), cs.data_ptr(), st.data_ptr(), P, eps,
                                   out.data_ptr(), N, M, N, K, gelu, L.current_stream()), "sgpt_linear_lnfold")
    xx = x.double()
    ln = (xx - xx.mean(1, keepdim=True)) / torch.sqrt(xx.var(1, unbiased=False, keepdim=True) + eps) * gamma.double() + beta.double()
    want = ln @ w.double().T + bias.double()
    if gelu:
        want = _gelu_new(want)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert err < 0.02 * scale, (err, scale)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1).min().item()
    assert cos > 1 - 2e-4, cos


@pytest.mark.parametrize("M,K,N,with_bias", [(77, 64, 128, True), (300, 768, 768, True), (1000, 3072, 768, True),
                                             (130, 256, 192, False), (513, 128, 2048, True)])
def test_linear_resid_ln(M, K, N, with_bias):
    L, lib = _lib()
    g = torch.Generator().manual_seed(M * 7 + K + N)
    x = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, generator=g) * 0.1 if with_bias else None
    resid = torch.randn(M, N, generator=g) * 2 + 0.3
    dev = "cuda"
    xd, wd, rd = x.to(dev), w.to(dev), resid.to(dev).clone()
    bd = bias.to(dev) if with_bias else None
    P = (N + 127) // 128
    xb = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    st = torch.full((M, P, 2), float("nan"), device=dev)
    L.check(lib.sgpt_linear_resid_ln(xd.data_ptr(), K, wd.data_ptr(), K, L.ptr(bd), rd.data_ptr(), xb.data_ptr(),
                                     st.data_ptr(), M, N, K, L.current_stream()), "sgpt_linear_resid_ln")
    want = resid.double() + x.double() @ w.double().T + (bias.double() if with_bias else 0.0)
    got = rd.cpu()
    assert torch.isfinite(got).all()
    assert (got.double() - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
    assert torch.equal(xb.cpu(), got.to(torch.bfloat16))  # the bf16 copy is the rounding of exactly what was stored
    pad = torch.nn.functional.pad(got.double(), (0, P * 128 - N)).view(M, P, 128)
    assert torch.allclose(st[..., 0].cpu().double(), pad.sum(-1), rtol=1e-5, atol=2e-3)
    assert torch.allclose(st[..., 1].cpu().double(), (pad * pad).sum(-1), rtol=1e-5, atol=2e-3)
    # a second application accumulates on top (in place) — exercises both buffer parities again
    L.check(lib.sgpt_linear_resid_ln(xd.data_ptr(), K, wd.data_ptr(), K, L.ptr(bd), rd.data_ptr(), xb.data_ptr(),
                                     st.data_ptr(), M, N, K, L.current_stream()))
    want2 = want + x.double() @ w.double().T + (bias.double() if with_bias else 0.0)
    assert (rd.cpu().double() - want2).abs().max().item() < 4e-3 * max(1.0, want2.abs().max().item())


def test_ln_fold_block_flow_matches_reference_fixture(golden_dir, monkeypatch):
    """The opt-in block flow without LayerNorm passes (SGPT_LN_FOLD=1, csrc/model.cu) against the same HF fixture as the
    default flow: pooled embeddings within 1e-3 cosine, all-layer pooling modes, and the two flows agree with each other."""
    import os

    import numpy as np

    from oracle import gpt_neo
    from sgpt_b200 import Encoder
    from tests.helpers import min_row_cosine
    from tests.test_gpu_parity import _cfg_from_spec, _spec_from

    z = np.load(os.path.join(golden_dir, "neo_tiny.npz"))
    spec = _spec_from(z)
    w = gpt_neo.init_weights(spec, seed=int(z["weight_seed"]))
    ids, mask = z["input_ids"], z["attention_mask"]
    outs = {}
    for fold in ("0", "1"):
        monkeypatch.setenv("SGPT_LN_FOLD", fold)
        enc = Encoder(_cfg_from_spec(spec), w, device="cuda:0", max_tokens=4096, max_batch=64)
        outs[fold] = {m: enc.encode_tokens(ids, mask, method=m).cpu() for m in ("weightedmean", "mean", "meanmean", "lasttoken")}
        outs[fold]["mid"] = enc.encode_tokens(ids, mask, method="weightedmean", layer_idx=int(z["mid_layer"])).cpu()
        outs[fold]["norm"] = enc.encode_tokens(ids, mask, method="weightedmean", normalize=True).cpu()
        enc.close()
    assert min_row_cosine(outs["1"]["weightedmean"], z["pooled_weightedmean"]) > 1 - 1e-3
    assert min_row_cosine(outs["1"]["mean"], z["pooled_mean"]) > 1 - 1e-3
    assert min_row_cosine(outs["1"]["mid"], z["pooled_weightedmean_mid"]) > 1 - 1e-3
    for k in outs["0"]:
        assert min_row_cosine(outs["1"][k], outs["0"][k]) > 1 - 2e-4, k
    assert torch.allclose(outs["1"]["norm"].norm(dim=1), torch.ones(len(ids)), atol=1e-5)


@pytest.mark.parametrize("resid_bf16", ["1", "0"])
def test_residual_stream_dtype_flows_match_reference_fixture(golden_dir, monkeypatch, resid_bf16):
    """The residual stream is stored in bf16 by default (what HF does for a bf16 checkpoint) and in fp32 with
    SGPT_RESID_BF16=0: both against the HF fp32 fixture — pooled embeddings within the 1e-3 cosine bar, per-token residual
    tap readable."""
    import os

    import numpy as np

    from oracle import gpt_neo
    from sgpt_b200 import Encoder
    from tests.helpers import min_row_cosine
    from tests.test_gpu_parity import _cfg_from_spec, _spec_from

    z = np.load(os.path.join(golden_dir, "neo_tiny.npz"))
    spec = _spec_from(z)
    w = gpt_neo.init_weights(spec, seed=int(z["weight_seed"]))
    ids, mask = z["input_ids"], z["attention_mask"]
    monkeypatch.setenv("SGPT_RESID_BF16", resid_bf16)
    enc = Encoder(_cfg_from_spec(spec), w, device="cuda:0", max_tokens=4096, max_batch=64)
    for method, key in (("weightedmean", "pooled_weightedmean"), ("mean", "pooled_mean")):
        got = enc.encode_tokens(ids, mask, method=method).cpu()
        assert min_row_cosine(got, z[key]) > 1 - 1e-3, method
    got = enc.encode_tokens(ids, mask, method="weightedmean", layer_idx=int(z["mid_layer"])).cpu()
    assert min_row_cosine(got, z["pooled_weightedmean_mid"]) > 1 - 1e-3
    sp = np.load(os.path.join(golden_dir, "script_pooling_neo_tiny.npz"))
    for method in ("meanmean", "lasttokenmean", "lasttoken"):
        assert min_row_cosine(enc.encode_tokens(ids, mask, method=method).cpu(), sp["pooled_" + method]) > 1 - 1e-3, method
    enc.encode_tokens(ids, mask)
    resid = enc.last_residual().cpu()
    h = gpt_neo.layer_norm(resid, w["ln_f.weight"], w["ln_f.bias"], spec.ln_eps)
    ref = torch.from_numpy(z["hidden_states"][-1])[torch.from_numpy(mask).bool()]
    assert torch.nn.functional.cosine_similarity(h.double(), ref.double(), dim=1).min().item() > 1 - 1e-3
    enc.close()
