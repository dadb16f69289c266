"""beir.retrieval.models (BDR:403, 412: ``models.SentenceBERT`` for plain sentence-transformers checkpoints)."""


class SentenceBERT:
    def __init__(self, model_path=None, sep: str = " ", **kwargs):
        from sgpt_b200.embedder import SentenceBERTBOSEOS, SentenceEncoder

        self._impl = SentenceBERTBOSEOS(SentenceEncoder.from_pretrained(model_path, **kwargs), sep=sep)

    def encode_queries(self, queries, batch_size: int = 16, **kwargs):
        return self._impl.encode_queries(queries, batch_size=batch_size, **kwargs)

    def encode_corpus(self, corpus, batch_size: int = 8, **kwargs):
        return self._impl.encode_corpus(corpus, batch_size=batch_size, **kwargs)
