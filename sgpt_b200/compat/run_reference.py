"""Run the reference's retrieval script unmodified on this library.

    python -m sgpt_b200.compat.run_reference /path/to/sgpt/biencoder/beir/beir_dense_retriever.py \
        --modelname Muennighoff/SGPT-125M-weightedmean-msmarco-specb-bitfit --method weightedmean --dataset scifact --specb

The script is loaded by path with ``sgpt_b200/compat`` first on ``sys.path`` (stand-in ``beir`` unless the real one is
installed, and ``custommodels``); its ``CustomEmbedder`` class (HF AutoModel + CPU pooling, BDR:98-348) is replaced by
``sgpt_b200.CustomEmbedder`` (same constructor keywords and encode_queries / encode_corpus protocol).  Nothing else of
the script is touched: argument parsing, dataset loading, the retrieve/evaluate flow and the result files are its own.
"""
from __future__ import annotations

import importlib.util
import sys

from . import COMPAT_DIR


def load_reference_script(path: str, embedder_cls=None, module_name: str = "sgpt_reference_beir_dense_retriever"):
    """Import beir_dense_retriever.py from `path` and swap its CustomEmbedder for `embedder_cls`
    (default: sgpt_b200.CustomEmbedder).  Returns the module (``module.main(module.parse_args())`` runs it)."""
    if COMPAT_DIR not in sys.path:
        try:
            import beir  # noqa: F401  (a real installation wins for `beir`; `custommodels` still comes from here)
            sys.path.append(COMPAT_DIR)
        except ImportError:
            sys.path.insert(0, COMPAT_DIR)
    spec = importlib.util.spec_from_file_location(module_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[module_name] = mod
    spec.loader.exec_module(mod)
    if embedder_cls is None:
        from sgpt_b200.embedder import CustomEmbedder as embedder_cls
    mod.CustomEmbedder = embedder_cls
    return mod


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0].startswith("-"):
        raise SystemExit("usage: python -m sgpt_b200.compat.run_reference /path/to/beir_dense_retriever.py [script args]")
    script, rest = argv[0], argv[1:]
    mod = load_reference_script(script)
    old = sys.argv
    sys.argv = [script] + rest
    try:
        mod.main(mod.parse_args())
    finally:
        sys.argv = old


if __name__ == "__main__":
    main()
