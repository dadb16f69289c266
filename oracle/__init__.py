"""CPU oracle for the SGPT bi-encoder hot path — TEST INFRASTRUCTURE ONLY.

This package restates, in plain CPU torch/numpy fp32/fp64 arithmetic, the algorithm the reference
(Muennighoff/sgpt @ 37c8bf09) runs for the path  token ids -> GPT forward -> position-weighted mean pool ->
cosine/dot scoring -> top-k merge.  It exists so the CUDA path can be checked against something independent.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product package ``sgpt_b200`` never imports it and has no CPU fallback.

Where the algorithm lives
-------------------------
* encoder arithmetic (GPT-Neo forward): in the reference's third-party dependency HuggingFace ``transformers``
  (pinned ``>=4.6.0,<5.0.0`` by biencoder/nli_msmarco/sentence-transformers/setup.py:21; call sites
  biencoder/beir/beir_dense_retriever.py:123,205 and sentence_transformers/models/Transformer.py:38,72).  Its source
  is not under /root/reference, so ``oracle.gpt_neo`` restates the published algorithm of
  ``transformers/models/gpt_neo/modeling_gpt_neo.py`` (line numbers of the installed 5.5.0 are cited per function).
* pooling: biencoder/beir/beir_dense_retriever.py:238-282 and sentence_transformers/models/Pooling.py:85-168.
* scoring / top-k / merge: sentence_transformers/util.py:24-63 and biencoder/beir/custommodels/exact_search.py:34-134.

Pinning
-------
The reference has no tests for the encoder or the pooling and no pretrained weights are
reachable offline, so the oracle is pinned against outputs of the reference code itself, run in the authoring
container: ``tests/golden/make_golden.py`` executes HF ``GPTNeoModel`` (the class the reference instantiates through
``AutoModel``) + the reference's own ``Pooling.py`` (loaded by file path from /root/reference) + the reference's own
``util.cos_sim`` / ``semantic_search`` and stores small input/output fixtures under ``tests/golden/``.
``tests/test_oracle.py`` checks every oracle function against those fixtures; the scoring stage is additionally
pinned by re-running the reference's property tests (sentence-transformers/tests/test_util.py:9-53) with fixed seeds.
"""
