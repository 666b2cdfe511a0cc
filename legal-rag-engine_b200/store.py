"""On-disk vector store: what ``create_vector_store()`` writes and ``RetrievalEngine`` loads
(reference: create_vector_store.py:63-78, retrieval_engine.py:35-56).

Native layout (``data/vector_store/``)
  vectors.f16.npy   float16 [N,384] unit rows (the HBM-resident chunk matrix, 768 B/row)
  bm25.npz          term_ptr u64[V+1], postings u32[nnz,2] (doc, tf), doc_len u32[N],
                    idf f64[V], avgdl, k1, b, epsilon, vocab (term strings in id order)
  metadata.json     the chunk list, ``json.dump(chunks, indent=2)`` exactly as the reference
                    (orchestrator.py:15-16 re-reads this file on its own)
  index.faiss       float32 IndexFlatIP image ("IxFI") so the reference can open our store
  bm25.pkl          the pickled ``rank_bm25.BM25Okapi`` the reference loads (retrieval_engine.py:45)

Interchange: a store written by the reference (index.faiss + bm25.pkl +
metadata.json) loads too -- the flat index is parsed directly and the pickled
``rank_bm25.BM25Okapi`` is read with a stub unpickler (rank_bm25 need not be installed).
The IxFI field layout is restated from FAISS's published index_write.cpp from memory and has
not been checked against a file written by faiss itself (none is available offline).
"""
from __future__ import annotations

import io
import json
import pickle
import struct
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

from .bm25_index import BM25Index, okapi_idf

DIM = 384


# ---------------------------------------------------------------- FAISS flat
def write_faiss_flat_ip(path, x: np.ndarray) -> None:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, 0))   # d, ntotal, 2 dummies, is_trained, METRIC_INNER_PRODUCT
        f.write(struct.pack("<Q", n * d))
        f.write(x.tobytes())


def read_faiss_flat(path) -> np.ndarray:
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise ValueError(f"{path}: not a flat FAISS index (fourcc {fourcc!r})")
        d, n, _, _, _trained, metric = struct.unpack("<iqqqBi", f.read(4 + 8 * 3 + 1 + 4))
        if metric > 1:
            f.read(4)                                  # metric_arg
        (count,) = struct.unpack("<Q", f.read(8))
        if count != n * d:
            raise ValueError(f"{path}: vector block holds {count} floats, expected {n}*{d}")
        x = np.frombuffer(f.read(count * 4), dtype=np.float32).reshape(n, d)
    return x


# ------------------------------------------------------------- bm25.pkl stub
class _StubBM25:
    """Receives the attributes of a pickled rank_bm25.BM25Okapi (corpus_size, avgdl,
    doc_freqs, idf, doc_len, k1, b, epsilon, average_idf)."""


# Exactly the globals a pickled rank_bm25.BM25Okapi needs (its attributes are ints, floats, lists,
# dicts and -- depending on the numpy in use when it was written -- numpy scalars / arrays).
# Anything else (builtins.eval, os.system, numpy.load ...) is refused: a crafted bm25.pkl must not
# be able to run code.
_PICKLE_ALLOW = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "str"), ("builtins", "bool"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float64"), ("numpy", "int64"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
}


class _BM25Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "rank_bm25":
            return _StubBM25
        if (module, name) in _PICKLE_ALLOW:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")


def write_reference_bm25_pickle(path, bm25: BM25Index) -> None:
    """``bm25.pkl`` as the reference's ``pickle.dump(bm25, f)`` writes it
    (create_vector_store.py:73-74): an object of class ``rank_bm25.BM25Okapi`` whose ``__dict__``
    holds rank_bm25 0.2.2's attributes, so that the reference's ``pickle.load``
    (retrieval_engine.py:45-46) -- with rank_bm25 installed -- gets a working scorer.  rank_bm25 is
    not importable here, so the class reference is written through a stand-in module of that name
    that exists only while pickling."""
    import sys
    import types
    terms = [None] * bm25.n_terms
    for w, i in (bm25.vocab or {}).items():
        terms[i] = w
    tp = bm25.term_ptr.astype(np.int64)
    doc_freqs = [dict() for _ in range(bm25.n_docs)]
    for t in range(bm25.n_terms):
        w = terms[t]
        for d, f in bm25.postings[tp[t]:tp[t + 1]].tolist():
            doc_freqs[d][w] = int(f)
    raw_idf = bm25.idf.tolist()
    mod = types.ModuleType("rank_bm25")

    class BM25Okapi:                     # noqa: N801 -- the pickled class path is rank_bm25.BM25Okapi
        pass
    BM25Okapi.__module__ = "rank_bm25"
    BM25Okapi.__qualname__ = "BM25Okapi"
    mod.BM25Okapi = BM25Okapi
    obj = BM25Okapi()
    obj.__dict__.update(
        corpus_size=int(bm25.n_docs), avgdl=float(bm25.avgdl), doc_freqs=doc_freqs,
        idf={terms[t]: float(raw_idf[t]) for t in range(bm25.n_terms)},
        doc_len=[int(v) for v in bm25.doc_len.tolist()], tokenizer=None,
        k1=float(bm25.k1), b=float(bm25.b), epsilon=float(bm25.epsilon),
        average_idf=float(okapi_idf(np.diff(tp), bm25.n_docs, bm25.epsilon)[1]))
    had = sys.modules.get("rank_bm25")
    sys.modules["rank_bm25"] = mod
    try:
        with open(path, "wb") as f:
            pickle.dump(obj, f)
    finally:
        if had is not None:
            sys.modules["rank_bm25"] = had
        else:
            del sys.modules["rank_bm25"]


def bm25_from_reference_pickle(path) -> BM25Index:
    with open(path, "rb") as f:
        obj = _BM25Unpickler(io.BytesIO(f.read())).load()
    vocab: Dict[str, int] = {w: i for i, w in enumerate(obj.idf.keys())}   # insertion order
    idf = np.fromiter(obj.idf.values(), dtype=np.float64, count=len(vocab))
    t_l, d_l, f_l = [], [], []
    for d, freqs in enumerate(obj.doc_freqs):
        for w, c in freqs.items():
            t_l.append(vocab[w]); d_l.append(d); f_l.append(c)
    t = np.asarray(t_l, dtype=np.int64); d = np.asarray(d_l, dtype=np.int64)
    f_ = np.asarray(f_l, dtype=np.int64)
    order = np.lexsort((d, t))
    t, d, f_ = t[order], d[order], f_[order]
    term_ptr = np.zeros(len(vocab) + 1, dtype=np.int64)
    np.add.at(term_ptr, t + 1, 1)
    term_ptr = np.cumsum(term_ptr).astype(np.uint64)
    postings = np.stack([d, f_], axis=1).astype(np.uint32)
    return BM25Index(int(obj.corpus_size), np.asarray(obj.doc_len, dtype=np.uint32), term_ptr,
                     postings, idf, float(obj.avgdl), vocab, float(obj.k1), float(obj.b),
                     float(obj.epsilon))


# -------------------------------------------------------------------- native
def save_store(save_dir, chunks: List[dict], x_f32: np.ndarray, bm25: BM25Index) -> None:
    save_dir = Path(save_dir)
    save_dir.mkdir(parents=True, exist_ok=True)
    np.save(save_dir / "vectors.f16.npy", np.ascontiguousarray(x_f32, dtype=np.float32).astype(np.float16))
    write_faiss_flat_ip(save_dir / "index.faiss", x_f32)
    terms = [None] * bm25.n_terms
    for w, i in (bm25.vocab or {}).items():
        terms[i] = w
    np.savez_compressed(save_dir / "bm25.npz", term_ptr=bm25.term_ptr, postings=bm25.postings,
                        doc_len=bm25.doc_len, idf=bm25.idf,
                        scalars=np.array([bm25.avgdl, bm25.k1, bm25.b, bm25.epsilon, bm25.n_docs]),
                        vocab=np.array(json.dumps(terms)))
    if bm25.vocab is not None:
        write_reference_bm25_pickle(save_dir / "bm25.pkl", bm25)   # create_vector_store.py:73-74
    with open(save_dir / "metadata.json", "w", encoding="utf-8") as f:
        json.dump(chunks, f, indent=2)                    # create_vector_store.py:77-78


def load_store(store_dir, mmap: bool = False) -> Tuple[List[dict], np.ndarray, BM25Index]:
    """-> (chunks, float16 [N,384] matrix, BM25Index).  Native files win; otherwise the
    reference's index.faiss / bm25.pkl are read.  mmap: map vectors.f16.npy instead of reading
    it (a shard of a multi-GPU engine slices its own row range out of it)."""
    store_dir = Path(store_dir)
    with open(store_dir / "metadata.json", "r", encoding="utf-8") as f:
        chunks = json.load(f)                             # retrieval_engine.py:53-55
    if (store_dir / "vectors.f16.npy").exists():
        xh = np.load(store_dir / "vectors.f16.npy", mmap_mode="r" if mmap else None)
    else:
        xh = read_faiss_flat(store_dir / "index.faiss").astype(np.float16)
    if (store_dir / "bm25.npz").exists():
        z = np.load(store_dir / "bm25.npz")
        avgdl, k1, b, eps, n_docs = (float(v) for v in z["scalars"])
        terms = json.loads(str(z["vocab"]))
        vocab = {w: i for i, w in enumerate(terms) if w is not None}
        bm25 = BM25Index(int(n_docs), z["doc_len"], z["term_ptr"], z["postings"], z["idf"], avgdl,
                         vocab, k1, b, eps)
    else:
        bm25 = bm25_from_reference_pickle(store_dir / "bm25.pkl")
    if xh.shape != (len(chunks), DIM) or bm25.n_docs != len(chunks):
        raise ValueError(f"store {store_dir} is inconsistent: {xh.shape} vectors, {bm25.n_docs} BM25 "
                         f"documents, {len(chunks)} chunks")
    return chunks, (xh if mmap else np.ascontiguousarray(xh)), bm25
