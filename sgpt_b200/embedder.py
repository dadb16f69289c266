"""Embedder plug-ins with the reference's call surface, running on the B200 encoder.

* ``CustomEmbedder``  — drop-in for biencoder/beir/beir_dense_retriever.py:106-348 (script / HF path): same constructor
  keywords, ``embed_batcher``, ``encode_queries(List[(qid, text)], ...)``, ``encode_corpus(List[(cid, {title,text})], ...)``.
* ``SentenceEncoder`` — the ``SentenceTransformer.encode`` signature of the vendored fork
  (sentence_transformers/SentenceTransformer.py:107-215) incl. length-sorted batching and specb bracket rules of
  models/Transformer.py:131-153; ``SentenceBERTBOSEOS`` mirrors custommodels/sentence_bert_asym.py:21-79 on top of it.

Tokenisation stays on the host (it is outside the measured path); anything with the HF tokenizer
methods used by the reference (``tokenize``, ``convert_tokens_to_ids``, ``encode``) works, so tests can plug a tiny
deterministic tokenizer in where no pretrained files are reachable.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .config import ModelConfig
from .encoder import Encoder

logger = logging.getLogger(__name__)

SPECB_QUE_BOS, SPECB_QUE_EOS = "[", "]"   # beir_dense_retriever.py:100-101
SPECB_DOC_BOS, SPECB_DOC_EOS = "{", "}"   # beir_dense_retriever.py:103-104

METHODS = ("mean", "weightedmean", "lasttoken", "meanmean", "lasttokenmean")  # BDR:238-301
ST_POOLING = ("mean", "weightedmean", "lasttoken")  # single-hidden-state modes of ST/models/Pooling.py


def _pad_batch(seqs: Sequence[Sequence[int]], pad_id: int) -> Tuple[np.ndarray, np.ndarray]:
    """tokenizer.pad(..., padding=True) (BDR:201): right-pad with pad_id, mask 1 on real tokens."""
    S = max(len(s) for s in seqs)
    ids = np.full((len(seqs), S), pad_id, dtype=np.int64)
    mask = np.zeros((len(seqs), S), dtype=np.int64)
    for i, s in enumerate(seqs):
        ids[i, :len(s)] = s
        mask[i, :len(s)] = 1
    return ids, mask


class CustomEmbedder:
    def __init__(self, model_name: str = "EleutherAI/gpt-neo-1.3B", batch_size: int = 250, device: str = "cuda:0",
                 save_emb: bool = False, reinit: bool = False, layeridx: int = -1, method: str = "mean",
                 dataset: str = "scifact", specb: bool = False, maxseqlen: Optional[int] = None, *,
                 config: Optional[ModelConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 tokenizer=None, max_tokens: Optional[int] = None, **kwargs):
        """Keywords up to ``maxseqlen`` are the reference's (BDR:107-120).  ``config``/``state_dict``/``tokenizer`` let a
        caller hand the model over directly (no hub access offline); otherwise they are loaded with HF
        ``AutoConfig``/``AutoModel``/``AutoTokenizer`` from ``model_name`` exactly like BDR:123,138."""
        if method not in METHODS:
            raise NotImplementedError(f"pooling method {method!r}: built: {METHODS} (poolout needs a pooler head the "
                                      "GPT models of the reference do not have, BDR:303-304)")
        if save_emb:
            logger.warning("save_emb pickle cache (BDR:311-323) is not implemented; embeddings are recomputed")
        if state_dict is None:
            from transformers import AutoModel  # real checkpoint path

            hf = AutoModel.from_pretrained(model_name, **kwargs)
            if reinit:
                hf.init_weights()  # BDR:124-126
            config = ModelConfig.from_hf(hf.config)
            state_dict = hf.state_dict()
        if config is None:
            raise ValueError("config is required when state_dict is given")
        if tokenizer is None:
            from transformers import AutoTokenizer

            tokenizer = AutoTokenizer.from_pretrained(model_name)
            if "gpt" in model_name.lower():
                tokenizer.pad_token = tokenizer.eos_token  # BDR:140-141
        self.tokenizer = tokenizer
        self.config = config
        self.max_token_len = maxseqlen if maxseqlen else config.max_pos  # BDR:128
        if "bert" in model_name:
            self.max_token_len -= 2  # BDR:131-133
        if specb:
            self.max_token_len -= 2  # leave two tokens for the brackets, BDR:134-136
        self.batch_size = batch_size
        self.layeridx = layeridx
        self.method = method
        self.specb = specb
        if specb:
            self.bos_token_q = list(tokenizer.encode(SPECB_QUE_BOS))  # BDR:150-153
            self.eos_token_q = list(tokenizer.encode(SPECB_QUE_EOS))
            self.bos_token_d = list(tokenizer.encode(SPECB_DOC_BOS))
            self.eos_token_d = list(tokenizer.encode(SPECB_DOC_EOS))
        pad = getattr(tokenizer, "pad_token_id", None)
        self.pad_id = int(pad) if pad is not None else 0
        seq_cap = self.max_token_len + (2 if specb else 0)
        self.encoder = Encoder(config, state_dict, device=device,
                               max_tokens=max_tokens or max(batch_size * min(seq_cap, 512), 8192),
                               max_batch=max(batch_size, 1))
        self.device = self.encoder.device

    # ------------------------------------------------------------------------------------------------------------
    def tokenize_batch(self, batch: Sequence[str], is_query: bool) -> Tuple[np.ndarray, np.ndarray]:
        """Host token preparation of ``embed`` (BDR:164-201): newline -> space, tokenize, truncate to max_token_len,
        optional specb brackets with mask 1, right-pad."""
        seqs = []
        docs_truncated = toks_truncated = total = 0
        for txt in batch:
            txt = txt.replace("\n", " ")  # BDR:166
            tokens = self.tokenizer.convert_tokens_to_ids(self.tokenizer.tokenize(txt))  # BDR:169-170
            n = len(tokens)
            total += n
            if n > self.max_token_len:
                docs_truncated += 1
                toks_truncated += n - self.max_token_len
            elif n == 0:
                raise ValueError("Empty items should be cleaned prior to running")  # BDR:180-181
            ids = list(tokens[: self.max_token_len])  # GPT tokenizers add no special tokens in prepare_for_model
            if self.specb:
                ids = (self.bos_token_q + ids + self.eos_token_q) if is_query else (self.bos_token_d + ids + self.eos_token_d)
            seqs.append(ids)
        if docs_truncated:
            logger.warning(f"Truncated {docs_truncated} out of {len(batch)} documents by {toks_truncated} out of {total}.")
        return _pad_batch(seqs, self.pad_id)

    def embed_texts(self, sentences: Sequence[str], is_query: bool) -> torch.Tensor:
        """All sentences -> fp32 [n, D] on the device, in the given order (batches of self.batch_size, BDR:227-229)."""
        out = torch.empty((len(sentences), self.config.d_model), dtype=torch.float32, device=self.device)
        for i in range(0, len(sentences), self.batch_size):
            ids, mask = self.tokenize_batch(sentences[i:i + self.batch_size], is_query)
            out[i:i + len(ids)] = self.encoder.encode_tokens(ids, mask, method=self.method, layer_idx=self.layeridx)
        return out

    def embed_batcher(self, texts: List[Tuple[str, str]], is_query: bool, out_name=None, **kwargs) -> Dict[str, np.ndarray]:
        """{id: embedding} like BDR:225-314 (one D2H copy for the whole list instead of per-row .numpy())."""
        ids, sentences = zip(*texts) if texts else ((), ())
        emb = self.embed_texts(list(sentences), is_query).cpu().numpy()
        return {i: e for i, e in zip(ids, emb)}

    def encode_queries(self, queries: List[Tuple[str, str]], batch_size: int = None, convert_to_tensor: bool = False,
                       **kwargs) -> Union[np.ndarray, torch.Tensor]:
        """BDR:316-330: rows in the order given.  convert_to_tensor=True keeps the result on the device."""
        emb = self.embed_texts([t for (_, t) in queries], is_query=True)
        logger.info(f"Produced embeddings of shape {tuple(emb.shape)}")
        return emb if convert_to_tensor else emb.cpu().numpy()

    def encode_corpus(self, corpus: List[Tuple[str, Dict[str, str]]], batch_size: int = None, batch_num="",
                      convert_to_tensor: bool = False, **kwargs) -> Union[np.ndarray, torch.Tensor]:
        """BDR:332-348: text = (title + " " + text).strip() when a title key exists (BDR:341)."""
        texts = [((d["title"] + " " + d["text"]).strip() if "title" in d else d["text"].strip()) for (_, d) in corpus]
        emb = self.embed_texts(texts, is_query=False)
        logger.info(f"Produced embeddings of shape {tuple(emb.shape)}")
        return emb if convert_to_tensor else emb.cpu().numpy()


class SentenceEncoder:
    """``SentenceTransformer.encode`` for a [Transformer -> Pooling(-> Normalize)] SGPT model on the B200 encoder."""

    def __init__(self, config: ModelConfig, state_dict: Dict[str, torch.Tensor], tokenizer, device: str = "cuda:0",
                 pooling: str = "weightedmean", max_seq_length: int = 300, batch_capacity: int = 256,
                 max_tokens: Optional[int] = None):
        if pooling not in ST_POOLING:
            raise NotImplementedError(f"pooling mode {pooling!r} not in {ST_POOLING}")
        self.config, self.tokenizer, self.pooling = config, tokenizer, pooling
        self.max_seq_length = max_seq_length
        self.encoder = Encoder(config, state_dict, device=device,
                               max_tokens=max_tokens or batch_capacity * max_seq_length, max_batch=batch_capacity)
        self.device = self.encoder.device
        pad = getattr(tokenizer, "pad_token_id", None)
        self.pad_id = int(pad) if pad is not None else 0
        # specb/speca state installed by SentenceBERTBOSEOS (models/Transformer.py attributes of the same names)
        self.bos_spec_token_q = self.bos_spec_token_d = self.eos_spec_token_q = self.eos_spec_token_d = None
        self.bos_spec_token_q_rep = self.bos_spec_token_d_rep = None
        self.replace_bos = False

    def _text_length(self, text) -> int:
        return len(text)  # SentenceTransformer.py:600-614 for plain strings

    def tokenize(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """models/Transformer.py:90-153: strip, tokenize with truncation, bracket rules, right-pad."""
        texts = [str(s).strip() for s in texts]
        spec = None not in (self.bos_spec_token_q, self.eos_spec_token_q, self.bos_spec_token_d, self.eos_spec_token_d)
        limit = self.max_seq_length - 2 if spec else self.max_seq_length  # :135
        seqs = []
        for t in texts:
            ids = list(self.tokenizer.encode(t))[:limit]
            if spec:
                if ids and ids[0] == self.bos_spec_token_d:
                    if self.replace_bos:
                        ids[0] = self.bos_spec_token_d_rep
                    ids.append(self.eos_spec_token_d)
                elif ids and ids[0] == self.bos_spec_token_q:
                    if self.replace_bos:
                        ids[0] = self.bos_spec_token_q_rep
                    ids.append(self.eos_spec_token_q)
                else:
                    raise ValueError(f"Did not find BOS Token in sequence: {t[:40]!r}")  # :148
            seqs.append(ids)
        return _pad_batch(seqs, self.pad_id)

    def encode(self, sentences: Union[str, List[str]], batch_size: int = 32, show_progress_bar: bool = None,
               output_value: str = "sentence_embedding", convert_to_numpy: bool = True, convert_to_tensor: bool = False,
               device: str = None, normalize_embeddings: bool = False, num_proc=None):
        """SentenceTransformer.py:107-215: sort by -len, encode in batches, undo the sort; str in -> 1-D out."""
        if output_value != "sentence_embedding":
            raise NotImplementedError("only output_value='sentence_embedding' is built")
        if convert_to_tensor:
            convert_to_numpy = False
        input_was_string = isinstance(sentences, str) or not hasattr(sentences, "__len__")
        if input_was_string:
            sentences = [sentences]
        order = np.argsort([-self._text_length(s) for s in sentences], kind="stable")
        out = torch.empty((len(sentences), self.config.d_model), dtype=torch.float32, device=self.device)
        for start in range(0, len(sentences), batch_size):
            idx = order[start:start + batch_size]
            ids, mask = self.tokenize([sentences[i] for i in idx])
            emb = self.encoder.encode_tokens(ids, mask, method=self.pooling, clamp=True, normalize=normalize_embeddings)
            out[torch.as_tensor(idx, device=self.device)] = emb
        if convert_to_tensor:
            res = out
        elif convert_to_numpy:
            res = out.cpu().numpy()
        else:
            res = list(out)
        return res[0] if input_was_string else res


class SentenceBERTBOSEOS:
    """custommodels/sentence_bert_asym.py:21-79 on top of SentenceEncoder (specb only: '[SOS]'/'{SOS}' markers are
    replaced by the bracket ids and the closing bracket is appended)."""

    def __init__(self, model: SentenceEncoder, sep: str = " ", specb: bool = False, sos_q: int = None, sos_d: int = None):
        self.model, self.sep, self.specb = model, sep, specb
        if specb:
            tok = model.tokenizer
            model.bos_spec_token_q = sos_q if sos_q is not None else tok.encode("[SOS]")[0]
            model.bos_spec_token_d = sos_d if sos_d is not None else tok.encode("{SOS}")[0]
            model.bos_spec_token_q_rep = tok.encode("[")[0]
            model.eos_spec_token_q = tok.encode("]")[0]
            model.bos_spec_token_d_rep = tok.encode("{")[0]
            model.eos_spec_token_d = tok.encode("}")[0]
            model.replace_bos = True

    def encode_queries(self, queries: List[str], batch_size: int = 16, **kwargs):
        if self.specb:
            queries = ["[SOS]" + q for q in queries]
        return self.model.encode(queries, batch_size=batch_size, **kwargs)

    def encode_corpus(self, corpus: List[Dict[str, str]], batch_size: int = 8, **kwargs):
        pre = "{SOS}" if self.specb else ""
        sentences = [(pre + doc["title"] + self.sep + doc["text"]).strip() if "title" in doc else pre + doc["text"].strip()
                     for doc in corpus]
        return self.model.encode(sentences, batch_size=batch_size, **kwargs)
