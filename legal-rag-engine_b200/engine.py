"""Drop-in for the reference's ``src/retrieval/retrieval_engine.py`` and
``create_vector_store.py``: same class, constructor, ``search`` signature, result dicts and
store directory; everything numeric runs in the CUDA library behind include/lrx.h.

    from legal_rag_engine_b200.engine import RetrievalEngine, create_vector_store

Additive surface only: ``search_batch`` (the orchestrator's concept-expansion fan-out in one
launch chain, orchestrator.py:38-62), ``fusion="linear"|"rrf"`` (default "linear" = the
reference's behaviour, retrieval_engine.py:71-96), ``encode``.

There is no CPU fallback: without a CUDA device or the built library the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import json
import logging
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import store as _store
from ._lib import LRX_MAX_BATCH, LRX_MAX_QUERY_TERMS
from .bm25_index import BM25Index, tokenize
from .device_index import FUSION, DeviceIndex
from .encoder import SentenceEncoder

logger = logging.getLogger("LegalRAG-RetrievalEngine")


def _resolve_model_dir(model_dir: Optional[str]) -> Optional[str]:
    """EMBEDDING_MODEL_DIR, or a snapshot of all-MiniLM-L6-v2 inside the project's
    ``.hf_cache`` (where the reference's build phase caches it, retrieval_engine.py:8-9)."""
    if model_dir:
        return model_dir
    env = os.getenv("EMBEDDING_MODEL_DIR")
    if env:
        return env
    for root in (Path(os.getcwd()) / ".hf_cache", Path(os.getenv("HF_HOME", "")) if os.getenv("HF_HOME") else None):
        if root and root.exists():
            for p in root.rglob("model.safetensors"):
                if "MiniLM-L6" in str(p):
                    return str(p.parent)
    return None


class RetrievalEngine:
    def __init__(self, store_dir: str = "data/vector_store", *, device: int = 0,
                 model_dir: Optional[str] = None, encoder_state_dict: Optional[Dict] = None,
                 tokenizer=None, fusion: str = "linear"):
        self.store_dir = Path(store_dir)
        self.fusion = fusion
        self.dev = DeviceIndex(device)
        # 1. model (retrieval_engine.py:27-33)
        if encoder_state_dict is None:
            model_dir = _resolve_model_dir(model_dir)
            if model_dir is None:
                raise FileNotFoundError(
                    "all-MiniLM-L6-v2 weights not found: set EMBEDDING_MODEL_DIR to a directory "
                    "holding model.safetensors and vocab.txt (or pass encoder_state_dict=)")
        self.model = SentenceEncoder(self.dev, state_dict=encoder_state_dict, model_dir=model_dir,
                                     tokenizer=tokenizer)
        # 2-4. index, BM25, metadata (retrieval_engine.py:35-56)
        self.chunks, xh, self.bm25 = _store.load_store(self.store_dir)
        self._x = torch.from_numpy(xh).to(self.dev.device)
        self.dev.set_corpus(self._x, 0)
        self.dev.set_postings(self.bm25.term_ptr, self.bm25.postings, self.bm25.doc_len, self.bm25.idf,
                              self.bm25.avgdl, self.bm25.k1, self.bm25.b)
        logger.info("Store resident on GPU %d: %d chunks, %d postings", device, len(self.chunks),
                    self.bm25.nnz)

    # ------------------------------------------------------------------ API
    def encode(self, texts: Sequence[str]) -> np.ndarray:
        return self.model.encode(texts)

    def search(self, query: str, k: int = 5, hybrid_weight: float = 0.5, fusion: Optional[str] = None):
        return self.search_batch([query], k, [hybrid_weight], fusion)[0]

    def search_batch(self, queries: Sequence[str], k: int = 5,
                     hybrid_weights: Optional[Sequence[float]] = None, fusion: Optional[str] = None):
        """All queries through one encoder pass and one K2 -> K3 -> K4 chain.  Returns one
        result list per query, each exactly what ``search`` returns."""
        queries = list(queries)
        if hybrid_weights is None:
            hybrid_weights = [0.5] * len(queries)
        mode = fusion or self.fusion
        out: List[List[dict]] = []
        for s in range(0, len(queries), LRX_MAX_BATCH):
            out.extend(self._search_block(queries[s:s + LRX_MAX_BATCH],
                                          list(hybrid_weights[s:s + LRX_MAX_BATCH]), k, mode))
        return out

    def _search_block(self, queries, weights, k, mode):
        B = len(queries)
        enc = [self.model.tokenizer.encode(q, self.model.MAX_SEQ) for q in queries]
        S = max(len(e) for e in enc)
        ids = np.zeros((B, S), dtype=np.int32)
        lens = np.empty(B, dtype=np.int32)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = e
            lens[i] = len(e)
        # BM25 side: query.lower().split() -> term ids, unknown -> -1 (retrieval_engine.py:67)
        term_lists = [self.bm25.term_ids(tokenize(q))[:LRX_MAX_QUERY_TERMS] for q in queries]
        ptr = np.zeros(B + 1, dtype=np.int32)
        for i, t in enumerate(term_lists):
            ptr[i + 1] = ptr[i] + len(t)
        terms = np.fromiter((x for t in term_lists for x in t), dtype=np.int32, count=int(ptr[-1]))
        if terms.size == 0:
            terms = np.zeros(1, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        kk = max(1, min(int(k), 128))
        o_ids = np.empty((B, kk), dtype=np.int64)
        o_score, o_sem, o_kw = (np.empty((B, kk), dtype=np.float64) for _ in range(3))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.dev._ck(self.dev.lib.lrx_search_text_host(
            self.dev.h, vp(ids), vp(lens), S, vp(terms), vp(ptr), vp(w), B, kk, FUSION[mode],
            vp(o_ids), vp(o_score), vp(o_sem), vp(o_kw)))
        results = []
        for b in range(B):
            rows = []
            for j in range(min(kk, int(k))):
                idx = int(o_ids[b, j])
                if idx < 0:
                    continue
                rows.append({"chunk": self.chunks[idx], "score": float(o_score[b, j]),
                             "semantic": float(o_sem[b, j]), "keyword": float(o_kw[b, j])})
            results.append(rows)
        return results

    def close(self):
        self.dev.close()


def fanout_queries(query: str, user_context: str, key_entities: Sequence[str], category: str):
    """The orchestrator's concept expansion (orchestrator.py:38-56) as data: the search
    strings and their hybrid weights, ready for ``search_batch``."""
    queries = [query]
    if user_context == "victim_distress":
        offence = next((e for e in key_entities if e.lower() in ["robbery", "assault", "rape", "theft"]),
                       "crime")
        queries += [f"How to file FIR for {offence} BNSS procedure",
                    f"Victim compensation rights for {offence} NALSA scheme",
                    "Zero FIR registration procedure BNSS"]
    weights = [0.6 if category == "procedure" or "procedure" in q.lower() else 0.5 for q in queries]
    return queries, weights


def merge_fanout(result_lists: Sequence[List[dict]]) -> List[dict]:
    """orchestrator.py:54-62: concatenate in query order, first occurrence of each truthy
    ``canonical_header`` wins."""
    out, seen = [], set()
    for results in result_lists:
        for r in results:
            cid = r["chunk"].get("canonical_header")
            if cid and cid not in seen:
                out.append(r)
                seen.add(cid)
    return out


def create_vector_store(chunks_path: str = "legal_chunks.json", save_dir: str = "data/vector_store",
                        *, device: int = 0, model_dir: Optional[str] = None,
                        encoder_state_dict: Optional[Dict] = None, tokenizer=None,
                        batch_size: int = 1024):
    """Index build (create_vector_store.py:14-83): embed every chunk text on the GPU, build the
    BM25 statistics/postings, write the store."""
    chunks_path = Path(chunks_path)
    if not chunks_path.exists():
        print(f"Error: {chunks_path} not found. Run ingest_legal_docs.py first.")
        return
    with open(chunks_path, "r", encoding="utf-8") as f:
        chunks = json.load(f)
    if not chunks:
        print("No chunks to process.")
        return
    if encoder_state_dict is None:
        model_dir = _resolve_model_dir(model_dir)
        if model_dir is None:
            raise FileNotFoundError("all-MiniLM-L6-v2 weights not found: set EMBEDDING_MODEL_DIR")
    dev = DeviceIndex(device)
    try:
        enc = SentenceEncoder(dev, state_dict=encoder_state_dict, model_dir=model_dir, tokenizer=tokenizer)
        texts = [c["text"] for c in chunks]
        x = enc.encode(texts, batch_size=batch_size)          # unit float32 rows
        bm25 = BM25Index.from_texts(texts)                     # text.lower().split()
        _store.save_store(save_dir, chunks, x, bm25)
    finally:
        dev.close()
    print(f"Vector store created: {save_dir}  ({len(chunks)} chunks, dim {x.shape[1]})")
    return save_dir
