"""beir.retrieval.evaluation.EvaluateRetrieval (BDR:16, 440-446)."""
from sgpt_b200.beir_compat import EvaluateRetrieval  # noqa: F401
