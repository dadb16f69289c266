"""beir.util subset: cos_sim / dot_score (imported by custommodels/exact_search.py:9) and download_and_unzip (BDR:368)."""
import os

import torch


def cos_sim(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    a_norm = torch.nn.functional.normalize(a, p=2, dim=1)
    b_norm = torch.nn.functional.normalize(b, p=2, dim=1)
    return torch.mm(a_norm, b_norm.transpose(0, 1))


def dot_score(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    return torch.mm(a, b.transpose(0, 1))


def download_and_unzip(url: str, out_dir: str, chunk_size: int = 1024) -> str:
    """Datasets must already be on disk: there is no network where this stand-in is used."""
    dataset = url.split("/")[-1].replace(".zip", "")
    path = os.path.join(out_dir, dataset)
    if os.path.isdir(path):
        return path
    raise RuntimeError(f"dataset {dataset!r} not found under {out_dir!r} and downloading is unavailable offline ({url})")
