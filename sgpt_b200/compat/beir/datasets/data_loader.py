"""beir.datasets.data_loader.GenericDataLoader (BDR:15, 379)."""
from sgpt_b200.beir_compat import GenericDataLoader  # noqa: F401
