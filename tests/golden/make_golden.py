"""Generate the golden fixtures under tests/golden/ by running the REFERENCE code in the authoring container.

    python tests/golden/make_golden.py          # needs /root/reference and `transformers`; writes *.npz here

What is executed (nothing from this repo's product path, and the oracle only supplies the synthetic weights):
  * encoder  : HuggingFace ``GPTNeoModel`` — the class ``AutoModel.from_pretrained`` resolves to at
               biencoder/beir/beir_dense_retriever.py:123 — with ``output_hidden_states=True`` (BDR:205), fp32, CPU.
  * pooling  : the reference's own ``sentence_transformers/models/Pooling.py`` loaded by file path
               (pooling_mode_weightedmean_tokens / mean / lasttoken), cross-checked here against the inline formula of
               beir_dense_retriever.py:258-270.
  * scoring  : the reference's own ``sentence_transformers/util.py`` (cos_sim, dot_score, semantic_search) loaded by
               file path with the two dead ``huggingface_hub`` imports stubbed (they are only used by its hub helpers).
The GPU box has no /root/reference, so the fixtures are committed; tests only read them.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_ST = "/root/reference/biencoder/nli_msmarco/sentence-transformers/sentence_transformers"

from oracle.gpt_neo import NeoSpec, init_weights  # noqa: E402  (synthetic weights only)


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_util():
    import huggingface_hub

    if not hasattr(huggingface_hub, "cached_download"):
        huggingface_hub.cached_download = None  # removed upstream; only used by util.snapshot_download
    if "huggingface_hub.snapshot_download" not in sys.modules:
        stub = types.ModuleType("huggingface_hub.snapshot_download")
        stub.REPO_ID_SEPARATOR = "--"
        sys.modules["huggingface_hub.snapshot_download"] = stub
    return load_by_path("ref_st_util", os.path.join(REF_ST, "util.py"))


def hf_model(spec: NeoSpec, weights):
    from transformers import GPTNeoConfig, GPTNeoModel

    cfg = GPTNeoConfig(
        vocab_size=spec.vocab, max_position_embeddings=spec.max_pos, hidden_size=spec.d_model,
        num_layers=spec.n_layer, num_heads=spec.n_head, intermediate_size=spec.d_ff, window_size=spec.window,
        attention_types=[[["global", "local"], spec.n_layer // 2]], layer_norm_epsilon=spec.ln_eps,
        embed_dropout=0.0, attention_dropout=0.0, resid_dropout=0.0, activation_function="gelu_new")
    model = GPTNeoModel(cfg)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    assert all(("attn.attention.bias" in k) or ("masked_bias" in k) for k in missing), missing
    assert model.config._attn_implementation == "eager", model.config._attn_implementation
    return model.float().eval()


def hf_gptj(spec, weights):
    from transformers import GPTJConfig, GPTJModel

    cfg = GPTJConfig(vocab_size=spec.vocab, n_positions=spec.max_pos, n_embd=spec.d_model, n_layer=spec.n_layer,
                     n_head=spec.n_head, rotary_dim=spec.rotary_dim, n_inner=spec.d_ff, activation_function="gelu_new",
                     resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=spec.ln_eps)
    model = GPTJModel(cfg)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected, unexpected
    assert all("embed_positions" in k or "attn.bias" in k or "masked_bias" in k for k in missing), missing
    assert model.config._attn_implementation == "eager"
    return model.float().eval()


def hf_bloom(spec, weights):
    from transformers import BloomConfig, BloomModel

    cfg = BloomConfig(vocab_size=spec.vocab, hidden_size=spec.d_model, n_layer=spec.n_layer, n_head=spec.n_head,
                      layer_norm_epsilon=spec.ln_eps, hidden_dropout=0.0, attention_dropout=0.0)
    model = BloomModel(cfg)
    missing, unexpected = model.load_state_dict(weights, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    assert model.config._attn_implementation == "eager"
    return model.float().eval()


def run_family_case(name, model, spec_array, ids, mask, d_model):
    """HF model (GPT-J / BLOOM) + the reference's Pooling.py -> fixture with hidden states and pooled embeddings."""
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
    hs = out.hidden_states
    assert torch.equal(hs[-1], out.last_hidden_state)
    Pooling = load_by_path("ref_pooling", os.path.join(REF_ST, "models", "Pooling.py")).Pooling
    pooled = {}
    for mode, kw in (("weightedmean", dict(pooling_mode_weightedmean_tokens=True)),
                     ("mean", dict(pooling_mode_mean_tokens=True))):
        kwargs = dict(pooling_mode_cls_token=False, pooling_mode_max_tokens=False, pooling_mode_mean_tokens=False)
        kwargs.update(kw)
        pooled[mode] = Pooling(d_model, **kwargs)({"token_embeddings": hs[-1].clone(), "attention_mask": mask})[
            "sentence_embedding"].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), input_ids=ids.numpy().astype(np.int32),
                        attention_mask=mask.numpy().astype(np.int8), pooled_weightedmean=pooled["weightedmean"],
                        pooled_mean=pooled["mean"], hidden_states=np.stack([h.numpy() for h in hs]),
                        weight_seed=np.int32(0), spec=spec_array)
    print(name, hs[-1].shape)


def ragged_batch(B, S, vocab, seed, pad_id):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    lens[0] = 1
    lens[1] = S
    ids = torch.randint(0, vocab, (B, S), generator=g)
    mask = (torch.arange(S).unsqueeze(0) < lens.unsqueeze(1)).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, pad_id))
    return ids, mask


def run_case(name, spec, B, S, seed, keep_hidden):
    w = init_weights(spec, seed=0)
    model = hf_model(spec, w)
    ids, mask = ragged_batch(B, S, spec.vocab, seed, pad_id=min(50256, spec.vocab - 1))
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=mask, output_hidden_states=True)
    hs = out.hidden_states
    assert torch.equal(hs[-1], out.last_hidden_state)
    Pooling = load_by_path("ref_pooling", os.path.join(REF_ST, "models", "Pooling.py")).Pooling
    res = {}
    for mode, kw in (("weightedmean", dict(pooling_mode_weightedmean_tokens=True)),
                     ("mean", dict(pooling_mode_mean_tokens=True)),
                     ("lasttoken", dict(pooling_mode_lasttoken=True))):
        kwargs = dict(pooling_mode_cls_token=False, pooling_mode_max_tokens=False, pooling_mode_mean_tokens=False)
        kwargs.update(kw)
        pool = Pooling(spec.d_model, **kwargs)
        feats = {"token_embeddings": hs[-1].clone(), "attention_mask": mask}
        res[mode] = pool(feats)["sentence_embedding"].numpy()
    # the script-path formula (beir_dense_retriever.py:258-270) must agree with Pooling.py
    m = mask.unsqueeze(-1).expand(hs[-1].size()).float()
    wts = torch.arange(1, S + 1).unsqueeze(0).unsqueeze(-1).expand(hs[-1].size()).float()
    script = (torch.sum(hs[-1] * m * wts, dim=1) / torch.sum(m * wts, dim=1)).numpy()
    assert np.abs(script - res["weightedmean"]).max() < 1e-6
    # layer-0 .. layer-(L-1) pooled weightedmean too (layeridx != -1, BDR:233)
    mid = spec.n_layer // 2
    pool = Pooling(spec.d_model, pooling_mode_cls_token=False, pooling_mode_max_tokens=False,
                   pooling_mode_mean_tokens=False, pooling_mode_weightedmean_tokens=True)
    mid_pooled = pool({"token_embeddings": hs[mid].clone(), "attention_mask": mask})["sentence_embedding"].numpy()
    payload = dict(input_ids=ids.numpy().astype(np.int32), attention_mask=mask.numpy().astype(np.int8),
                   pooled_weightedmean=res["weightedmean"], pooled_mean=res["mean"], pooled_lasttoken=res["lasttoken"],
                   pooled_weightedmean_mid=mid_pooled, mid_layer=np.int32(mid), weight_seed=np.int32(0),
                   spec=np.array([spec.n_layer, spec.d_model, spec.n_head, spec.d_ff, spec.vocab, spec.max_pos,
                                  spec.window], dtype=np.int32))
    if keep_hidden:
        payload["hidden_states"] = np.stack([h.numpy() for h in hs])  # [L+1, B, S, d]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(name, {k: getattr(v, "shape", v) for k, v in payload.items()})


def run_scoring():
    util = load_reference_util()
    g = torch.Generator().manual_seed(7)
    q = torch.randn(20, 100, generator=g)
    c = torch.randn(1000, 100, generator=g)
    c[::50] = q[torch.arange(20)] + 0.3 * c[::50]  # planted neighbours
    c[3] = 0.0  # zero vector -> cos = 0 via the 1e-12 clamp
    cos = util.cos_sim(q, c).numpy()
    dot = util.dot_score(q, c).numpy()
    # the reference's own chunked search (tests/test_util.py:33-53 uses exactly these chunk sizes)
    hits = util.semantic_search(q, c, query_chunk_size=5, corpus_chunk_size=17, top_k=10)
    hit_ids = np.array([[h["corpus_id"] for h in row] for row in hits], dtype=np.int64)
    hit_scores = np.array([[h["score"] for h in row] for row in hits], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "scoring.npz"), queries=q.numpy(), corpus=c.numpy(), cos=cos, dot=dot,
                        hit_ids=hit_ids, hit_scores=hit_scores)
    print("scoring", cos.shape, hit_ids.shape)


if __name__ == "__main__":
    torch.set_num_threads(8)
    tiny = NeoSpec(n_layer=4, d_model=128, n_head=2, d_ff=512, vocab=1000, max_pos=128, window=16)
    run_case("neo_tiny", tiny, B=6, S=48, seed=11, keep_hidden=True)
    # BASELINE.json configs[0]: SGPT-125M-weightedmean, 32 sentences, seq_len 64
    run_case("neo_125m_b32_s64", NeoSpec(), B=32, S=64, seed=1234, keep_hidden=False)
    run_scoring()
    # GPT-J (SGPT-5.8B family) and BLOOM (sgpt-bloom-7b1 family), tiny configs
    from oracle import bloom as obloom
    from oracle import gptj as ogptj

    js = ogptj.GPTJSpec(n_layer=3, d_model=256, n_head=2, d_ff=1024, vocab=1000, max_pos=128, rotary_dim=32)
    ids, mask = ragged_batch(5, 40, js.vocab, seed=21, pad_id=999)
    run_family_case("gptj_tiny", hf_gptj(js, ogptj.init_weights(js, 0)),
                    np.array([js.n_layer, js.d_model, js.n_head, js.d_ff, js.vocab, js.max_pos, js.rotary_dim], np.int32),
                    ids, mask, js.d_model)
    bs = obloom.BloomSpec(n_layer=3, d_model=256, n_head=4, vocab=1000)
    ids, mask = ragged_batch(5, 40, bs.vocab, seed=22, pad_id=3)
    run_family_case("bloom_tiny", hf_bloom(bs, obloom.init_weights(bs, 0)),
                    np.array([bs.n_layer, bs.d_model, bs.n_head, bs.vocab], np.int32), ids, mask, bs.d_model)
