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


# ---------------------------------------------------------------------------
# Checker for the int8 pre-filter of the CUDA scan (include/lrx.h,
# lrx_build_dense_prefilter).  Not part of the reference path: a flat IP index
# has no such stage; the stage must not change any result, which is what the
# parity tests assert.  These functions restate the shadow's DEFINITION so that
# the device-built bytes and the two error bounds can be checked on the CPU.
def int8_shadow(Xh: np.ndarray):
    """(xi int8 [n,384], scale float32 [n], E, X): per-row scaled int8 image of the
    fp16 matrix -- scale = max|x| / 127 and xi = rint(x * (127 / max|x|)), both in
    float32 -- with E = max_r ||x_r - scale_r xi_r||_2 and X = max_r ||scale_r xi_r||_2
    in float64 (every term is exact in float64)."""
    assert Xh.dtype == np.float16
    x = Xh.astype(np.float32)
    mx = np.abs(x).max(axis=1) if x.shape[0] else np.zeros(0, np.float32)
    scale = (mx / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(mx > 0, np.float32(127.0) / mx, np.float32(0.0)).astype(np.float32)
    xi = np.clip(np.rint(x * inv[:, None]), -127, 127).astype(np.int8)
    xh = scale.astype(np.float64)[:, None] * xi.astype(np.float64)
    err = np.sqrt(((x.astype(np.float64) - xh) ** 2).sum(axis=1))
    nrm = np.sqrt((xh ** 2).sum(axis=1))
    E = float(err.max()) if err.size else 0.0
    X = float(nrm.max()) if nrm.size else 0.0
    return xi, scale, E, X


def int8_query_digits(qh: np.ndarray):
    """(hi int [384], lo int [384], cq float32): q ~= cq * (256 * hi + lo), the two-digit
    int8 image of one fp16 query the scan multiplies the shadow with."""
    assert qh.dtype == np.float16 and qh.ndim == 1
    q = qh.astype(np.float32)
    mx = np.float32(np.abs(q).max())
    inv = np.float32(127.0) / mx if mx > 0 else np.float32(0.0)
    s = (q * inv).astype(np.float32)
    hi = np.clip(np.rint(s), -127, 127)
    lo = np.clip(np.rint(((s - hi.astype(np.float32)).astype(np.float32) * np.float32(256.0)).astype(np.float32)),
                 -127, 127)
    cq = np.float32(mx * np.float32(1.0 / 32512.0))
    return hi.astype(np.int64), lo.astype(np.int64), cq


def int8_guard_band(qh: np.ndarray, E: float, X: float) -> float:
    """The rigorous bound |x.q - A(x, q)| <= E |q| + X |q - qhat| + rounding the exactness
    guard of the pre-filtered scan uses (csrc/dense.cu:dense_merge_rescore_kernel)."""
    hi, lo, cq = int8_query_digits(qh)
    q = qh.astype(np.float64)
    d = q - float(cq) * (256 * hi + lo).astype(np.float64)
    nq, ne = float(np.sqrt((q * q).sum())), float(np.sqrt((d * d).sum()))
    return (E * nq + X * ne + 2.0e-7 * X * (nq + ne)) * (1.0 + 1.0e-6) + 1.0e-12


def int8_fast_scores(Xh: np.ndarray, qh: np.ndarray):
    """float32 [n] scores as the pre-filtered scan computes them for one query:
    fl(fl(scale_r * cq) * fl(256 * <xi_r, hi> + <xi_r, lo>))."""
    xi, scale, _, _ = int8_shadow(Xh)
    hi, lo, cq = int8_query_digits(qh)
    D = xi.astype(np.int64) @ (256 * hi + lo)
    return ((scale * cq).astype(np.float32) * D.astype(np.float32)).astype(np.float32)
