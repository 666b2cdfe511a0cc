"""Drop-in for the reference's ``src/retrieval/retrieval_engine.py``: copy this file over it (and
put ``legal_rag_engine_b200`` on the path); ``orchestrator.py`` keeps its
``from .retrieval_engine import RetrievalEngine`` (orchestrator.py:6,10) and its
``self.engine.search(q, k=k, hybrid_weight=q_weight)`` (orchestrator.py:57) unchanged.

Same constructor (``RetrievalEngine(store_dir="data/vector_store")``), same ``search`` signature and
result dicts, same ``self.chunks``; the model directory comes from ``EMBEDDING_MODEL_DIR`` or the
project's ``.hf_cache`` as in the reference (retrieval_engine.py:8-9,27-33).  Under ``torchrun`` the
engine shards itself over the ranks (rank 0 serves, the others run ``worker_loop()``).
"""
from legal_rag_engine_b200.engine import RetrievalEngine  # noqa: F401

__all__ = ["RetrievalEngine"]
