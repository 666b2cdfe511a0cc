"""ctypes access to oracle/c/liboracle.so (test infrastructure; CPU baseline legs)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "c"
_lib = None


def load():
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(so))
        _lib.oracle_flat_ip_search.restype = None
        _lib.oracle_flat_ip_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_bm25_scores.restype = None
        _lib.oracle_bm25_scores.argtypes = [C.c_void_p] * 5 + [C.c_double, C.c_void_p, C.c_int,
                                                               C.c_void_p]
    return _lib


def flat_ip_search_f32(x: np.ndarray, q: np.ndarray, k: int, nthreads: int = 0):
    """FAISS-style sequential fp32 scan + heap; one core per query."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    D = np.empty((q.shape[0], k), dtype=np.float32)
    I = np.empty((q.shape[0], k), dtype=np.int64)
    lib.oracle_flat_ip_search(x.ctypes.data, x.shape[0], x.shape[1], q.ctypes.data, q.shape[0], k,
                              D.ctypes.data, I.ctypes.data, nthreads)
    return D, I


def bm25_scores(csr, term_ids) -> np.ndarray:
    """oracle.bm25.BM25OkapiCSR.get_scores_ids in C (bit-identical, ~100x faster)."""
    lib = load()
    t = np.ascontiguousarray(term_ids, dtype=np.int32)
    score = np.zeros(csr.corpus_size, dtype=np.float64)
    tp = np.ascontiguousarray(csr.term_ptr, dtype=np.int64)
    pd = np.ascontiguousarray(csr.post_doc, dtype=np.int64)
    pt = np.ascontiguousarray(csr.post_tf, dtype=np.int64)
    idf = np.ascontiguousarray(csr.idf, dtype=np.float64)
    dn = np.ascontiguousarray(csr.doc_norm, dtype=np.float64)
    lib.oracle_bm25_scores(tp.ctypes.data, pd.ctypes.data, pt.ctypes.data, idf.ctypes.data,
                           dn.ctypes.data, float(csr.k1), t.ctypes.data, len(t), score.ctypes.data)
    return score
