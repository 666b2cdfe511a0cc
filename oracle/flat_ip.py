"""Oracle: flat inner-product index (FAISS ``IndexFlatIP`` semantics) over the
fp16-quantised chunk matrix.

Test infrastructure, see ``oracle/__init__.py``.  Parity unpinned.

Follows ``src/retrieval/retrieval_engine.py:62-64`` (``normalize_L2`` then
``index.search(query_vector, k*2)``) and ``create_vector_store.py:48-56``
(``astype('float32')``, ``normalize_L2``, ``IndexFlatIP(d).add``), restating the
published behaviour of a flat IP index: every row is scored, the best ``K`` are
returned best first as ``(D float32[B,K], I int64[B,K])``, and when ``K`` exceeds
the number of rows the tail is padded with id ``-1`` (guarded at
``retrieval_engine.py:80``).

Arithmetic definition (what the CUDA path must reproduce bit for bit):

* the stored matrix and the query are the fp16 roundings of the L2-normalised
  fp32 vectors (the B200 build keeps the chunk matrix in fp16, 768 B/row);
* ``score = sum_i fp16(q_i) * fp16(x_i)`` evaluated EXACTLY.  Every fp16 value is
  a multiple of 2**-24, so every product is a multiple of 2**-48; for
  L2-normalised vectors every partial sum is below 2 in magnitude, hence needs
  at most 49 bits and is exact in float64 IN ANY SUMMATION ORDER.  A float64
  dot product (here, and in the CUDA re-score kernel) is therefore the exact
  value, independent of BLAS blocking or reduction trees;
* ranking is the total order (exact score descending, id ascending) -- the
  documented tie-break (a sequential scan with a strict ``>`` heap test never
  lets a later equal score displace an earlier id);
* ``D`` is the exact score rounded once to float32 (FAISS's output type).
"""
from __future__ import annotations

import numpy as np

PAD_SCORE = np.float32(-3.4028234663852886e38)   # -FLT_MAX, FAISS's empty-slot value


def normalize_l2(v: np.ndarray) -> np.ndarray:
    """``faiss.normalize_L2``: row-wise x / ||x||_2 in float32, zero rows untouched."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=1))
    out = v.copy()
    nz = n > 0
    out[nz] = (v[nz] / n[nz, None]).astype(np.float32)
    return out


def quantize_fp16(v: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(v, dtype=np.float32).astype(np.float16)


def exact_scores(Xh: np.ndarray, qh: np.ndarray, chunk: int = 262144) -> np.ndarray:
    """float64 [B, N] exact inner products of fp16 rows (see module docstring)."""
    assert Xh.dtype == np.float16 and qh.dtype == np.float16
    q64 = qh.astype(np.float64)
    out = np.empty((qh.shape[0], Xh.shape[0]), dtype=np.float64)
    for s in range(0, Xh.shape[0], chunk):
        out[:, s:s + chunk] = q64 @ Xh[s:s + chunk].astype(np.float64).T
    return out


def topk_from_scores(scores: np.ndarray, K: int, id_base: int = 0):
    """(exact float64 [B,K], D float32 [B,K], I int64 [B,K]) in (score desc, id asc)
    order; pads with (-FLT_MAX, -1)."""
    Bq, N = scores.shape
    E = np.full((Bq, K), -np.inf, dtype=np.float64)
    D = np.full((Bq, K), PAD_SCORE, dtype=np.float32)
    I = np.full((Bq, K), -1, dtype=np.int64)
    kk = min(K, N)
    ids = np.arange(N, dtype=np.int64)
    for b in range(Bq):
        s = scores[b]
        if kk < N:
            # preselect generously so boundary ties are resolved by id, not by argpartition
            kth = np.partition(s, N - kk)[N - kk]
            cand = np.nonzero(s >= kth)[0]
        else:
            cand = ids
        order = np.lexsort((cand, -s[cand]))[:kk]
        sel = cand[order]
        E[b, :kk] = s[sel]
        D[b, :kk] = s[sel].astype(np.float32)
        I[b, :kk] = sel + id_base
    return E, D, I


def flat_ip_search(Xh: np.ndarray, qh: np.ndarray, K: int, id_base: int = 0):
    """``IndexFlatIP.search`` on fp16 data: returns (D float32 [B,K], I int64 [B,K])."""
    _, D, I = topk_from_scores(exact_scores(Xh, qh), K, id_base)
    return D, I
