"""Shared test helpers: a tiny deterministic tokenizer (no pretrained files are reachable offline) and builders."""
import re
import zlib

import numpy as np
import torch


class ToyTokenizer:
    """Whitespace/bracket tokenizer with the HF methods the reference calls (tokenize, convert_tokens_to_ids, encode)."""

    SPECIAL = {"[": 58, "]": 60, "{": 90, "}": 92, "[SOS]": 997, "{SOS}": 998}  # GPT-2 BPE ids of the brackets

    def __init__(self, vocab=1000, pad_token_id=999):
        self.vocab = vocab
        self.pad_token_id = pad_token_id
        self.eos_token = "<|endoftext|>"
        self.pad_token = None

    def tokenize(self, text):
        return re.findall(r"\[SOS\]|\{SOS\}|[\[\]{}]|[^\s\[\]{}]+", text)

    def convert_tokens_to_ids(self, tokens):
        return [self.SPECIAL[t] if t in self.SPECIAL else 100 + zlib.crc32(t.encode()) % (self.vocab - 200) for t in tokens]

    def encode(self, text, add_special_tokens=False):
        return self.convert_tokens_to_ids(self.tokenize(text))


def ragged_batch(B, S, vocab, seed, pad_id=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    lens[0] = 1
    if B > 1:
        lens[1] = S
    ids = torch.randint(0, vocab, (B, S), generator=g)
    mask = (torch.arange(S).unsqueeze(0) < lens.unsqueeze(1)).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, pad_id))
    return ids, mask


def planted_corpus(n, D, nq, seed):
    """Random corpus with ~1% planted near-duplicates of the queries so the top of the ranking is meaningful."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(nq, D, generator=g)
    c = torch.randn(n, D, generator=g)
    idx = torch.arange(0, n, 97)
    c[idx] = q[idx // 97 % nq] + 0.5 * c[idx]
    return q, c


def min_row_cosine(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return torch.nn.functional.cosine_similarity(a, b, dim=1).min().item()
