"""Oracle: score fusion of ``RetrievalEngine.search`` -- the reference's linear
fusion (parity target) and Reciprocal Rank Fusion (README-only upstream).

Test infrastructure, see ``oracle/__init__.py``.  Parity unpinned.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

RRF_K0 = 60.0


def linear_fuse(D_row, I_row, bm25_scores, max_bm25: float, k: int, hybrid_weight: float):
    """``src/retrieval/retrieval_engine.py:71-96`` line by line.

    ``D_row``/``I_row``: the flat-IP hits of ONE query in index order (float32 /
    int64, ``-1`` padded).  ``bm25_scores``: anything indexable by global chunk id
    returning float64.  Returns a list of (id, score, semantic, keyword) of
    length <= k, sorted by ``score`` descending with Python's stable sort (ties
    keep the flat-IP order)."""
    combined = []
    for dist, idx in zip(D_row, I_row):
        if idx == -1:
            continue                                      # :80
        semantic_score = float(dist)                      # :81 (float32 -> Python float)
        bm25_score = bm25_scores[int(idx)] / max_bm25     # :82
        combined_score = (semantic_score * (1 - hybrid_weight)) + (bm25_score * hybrid_weight)  # :84
        combined.append((int(idx), float(combined_score), semantic_score, float(bm25_score)))
    combined.sort(key=lambda r: r[1], reverse=True)       # :95 (stable)
    return combined[:k]                                   # :96


def rrf_fuse(dense: List[Tuple[int, float, float]], sparse: List[Tuple[int, float, float]],
             max_bm25: float, k: int):
    """Reciprocal Rank Fusion (README.md:39,82-83
    advertise it, no upstream code exists -> build definition).

    ``dense``: [(id, dense_exact_f64, bm25_f64)] best first (flat-IP order);
    ``sparse``: same triples for the BM25 list (score > 0 only, best first, ties
    by id).  ``rrf(d) = [d in dense] 1/(60+rank_dense) + [d in sparse] 1/(60+rank_bm25)``,
    ranks 1-based, the dense term added first.  Order: (rrf desc, id asc).
    Returns [(id, rrf, semantic float32-rounded, keyword = bm25/max_bm25)]."""
    acc = {}
    info = {}
    for r, (i, de, bm) in enumerate(dense, start=1):
        if i < 0:
            continue
        acc[i] = 0.0 + 1.0 / (RRF_K0 + r)
        info[i] = (de, bm)
    for r, (i, de, bm) in enumerate(sparse, start=1):
        if i < 0:
            continue
        acc[i] = acc.get(i, 0.0) + 1.0 / (RRF_K0 + r)
        info[i] = (de, bm)
    items = sorted(acc.items(), key=lambda kv: (-kv[1], kv[0]))[:k]
    out = []
    for i, s in items:
        de, bm = info[i]
        out.append((int(i), float(s), float(np.float32(de)), float(bm / max_bm25)))
    return out
