"""Minimal ``beir`` (0.2.3 API subset used by biencoder/beir/beir_dense_retriever.py:14-16, 364-446)."""
import logging

from . import util  # noqa: F401


class LoggingHandler(logging.Handler):
    """beir.LoggingHandler: writes records through tqdm.write so progress bars are not broken (BDR:14, 24)."""

    def __init__(self, level=logging.NOTSET):
        super().__init__(level)

    def emit(self, record):
        try:
            msg = self.format(record)
            try:
                import tqdm

                tqdm.tqdm.write(msg)
            except Exception:
                print(msg)
            self.flush()
        except (KeyboardInterrupt, SystemExit):
            raise
        except Exception:
            self.handleError(record)
