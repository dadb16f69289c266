"""beir.util subset used by biencoder/beir/beir_dense_retriever.py: `download_and_unzip` (BDR:364-371).  The scoring helpers
`beir.util.cos_sim` / `dot_score` (imported by the reference's own custommodels/exact_search.py:9) are deliberately absent:
the stand-in `custommodels.DenseRetrievalExactSearch` scores on the GPU (`sgpt_search`), there is no CPU scoring path."""
import os


def download_and_unzip(url: str, out_dir: str, chunk_size: int = 1024) -> str:
    """Datasets must already be on disk: there is no network where this stand-in is used."""
    dataset = url.split("/")[-1].replace(".zip", "")
    path = os.path.join(out_dir, dataset)
    if os.path.isdir(path):
        return path
    raise RuntimeError(f"dataset {dataset!r} not found under {out_dir!r} and downloading is unavailable offline ({url})")
