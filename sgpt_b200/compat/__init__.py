"""Stand-ins for the two import roots the reference's retrieval script needs besides torch/transformers, so that
``biencoder/beir/beir_dense_retriever.py`` runs UNMODIFIED on top of this library:

* ``beir``          — ``util``, ``LoggingHandler``, ``datasets.data_loader.GenericDataLoader``,
                      ``retrieval.evaluation.EvaluateRetrieval``, ``retrieval.search.dense.DenseRetrievalExactSearch``
                      (beir==0.2.3 is pinned by the reference but absent offline; the real package wins when installed);
* ``custommodels``  — the reference's own side package (``from custommodels import DenseRetrievalExactSearch,
                      SentenceBERTAsym, SentenceBERTBOSEOS``, BDR:18), re-exported from ``sgpt_b200``.

``sgpt_b200.compat.run_reference`` puts this directory on ``sys.path``, loads the script by path and replaces its
HF-based ``CustomEmbedder`` class (BDR:98-348) by ``sgpt_b200.CustomEmbedder`` — the only two substitutions.
"""
import os

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))
