"""Golden vectors for the script-path pooling modes (mean / meanmean / weightedmean / lasttoken / lasttokenmean).

    python tests/golden/make_script_pooling.py      # needs /root/reference; writes script_pooling_*.npz here

The pooling block of the reference is inline code in a script that cannot be imported offline
(biencoder/beir/beir_dense_retriever.py:238-304 sits inside CustomEmbedder.embed_batcher, and the module imports beir).
This generator therefore reads exactly those source lines from /root/reference at run time, wraps them in a function
and EXECUTES them — the reference's own statements, not a restatement — on the HF hidden states already stored in the
family fixtures (neo_tiny / gptj_tiny / bloom_tiny .npz, written by make_golden.py).  Nothing of the reference is
written into this repo; only the resulting embeddings are committed.
"""
import os
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
BDR = "/root/reference/biencoder/beir/beir_dense_retriever.py"
FIRST, LAST = 238, 304  # `if self.method == "mean":` ... `embedding = embedded_batch.pooler_output.cpu()`


def reference_pooling_fn():
    lines = open(BDR).read().splitlines()[FIRST - 1:LAST]
    assert lines[0].strip().startswith('if self.method == "mean"'), lines[0]
    assert "pooler_output" in lines[-1], lines[-1]
    body = textwrap.indent(textwrap.dedent("\n".join(lines)), "    ")
    src = ("def pool(self, hidden_state, all_hidden_states, input_mask_expanded, gather_indices, embedded_batch):\n"
           + body + "\n    return embedding\n")
    ns = {"torch": torch}
    exec(compile(src, BDR + f":{FIRST}-{LAST}", "exec"), ns)
    return ns["pool"]


def main():
    pool = reference_pooling_fn()
    for name in ("neo_tiny", "gptj_tiny", "bloom_tiny"):
        fx = np.load(os.path.join(HERE, name + ".npz"))
        hs = [torch.from_numpy(h) for h in fx["hidden_states"]]  # L+1 x [B,S,d] fp32 (HF output_hidden_states)
        mask = torch.from_numpy(fx["attention_mask"].astype(np.int64))
        # BDR:209-214 / BDR:198: the mask expansion and the last-token indices the block expects
        input_mask_expanded = mask.unsqueeze(-1).expand(hs[-1].size()).float()
        gather_indices = [int(n) - 1 for n in mask.sum(dim=1)]
        out = {}
        for method in ("mean", "meanmean", "weightedmean", "lasttoken", "lasttokenmean"):
            me = types.SimpleNamespace(method=method)
            emb = pool(me, hs[-1], hs, input_mask_expanded, gather_indices, None)
            out["pooled_" + method] = emb.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"script_pooling_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
