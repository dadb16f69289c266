"""beir.retrieval.search.dense.DenseRetrievalExactSearch (BDR:404): upstream beir's calling convention — plain
``List[str]`` queries and ``List[{title, text}]`` documents."""
from sgpt_b200.exact_search import DenseRetrievalExactSearch as _DRES


class DenseRetrievalExactSearch(_DRES):
    def __init__(self, model, batch_size: int = 128, corpus_chunk_size: int = 50000, **kwargs):
        super().__init__(model, batch_size=batch_size, corpus_chunk_size=corpus_chunk_size, plain_lists=True, **kwargs)
