"""Oracle: rank_bm25 0.2.2 ``BM25Okapi`` restated (float64, same operation order).

Test infrastructure, see ``oracle/__init__.py``.  Parity unpinned (no reference
goldens): two independent forms are kept and must agree BIT FOR BIT.

Follows the reference call sites
  build : ``create_vector_store.py:60-61``  (``BM25Okapi([t.lower().split() ...])``)
  query : ``src/retrieval/retrieval_engine.py:67-68,74``
and the published rank_bm25 0.2.2 algorithm (``BM25.__init__/_initialize``,
``BM25Okapi._calc_idf/get_scores``; k1=1.5, b=0.75, epsilon=0.25):

  idf[t]   = ln(N - df + 0.5) - ln(df + 0.5);  negative idfs are replaced by
             epsilon * mean(raw idf over the vocabulary)
  score[d] = sum over query tokens IN ORDER, repeats included, of
             idf[t] * ( tf*(k1+1) / (tf + k1*(1 - b + b*len(d)/avgdl)) )
  tokens outside the vocabulary, and tokens whose idf is exactly 0, add 0
  (``self.idf.get(q) or 0``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

K1 = 1.5
B = 0.75
EPSILON = 0.25


def tokenize(text: str) -> List[str]:
    """``text.lower().split()`` -- create_vector_store.py:60, retrieval_engine.py:67."""
    return text.lower().split()


class BM25OkapiLiteral:
    """Dict-per-document form, the shape rank_bm25 itself has.  O(|q|*N) Python
    work per query: use on small corpora (and as the 'reference-literal' CPU
    baseline sample in bench.py)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1=K1, b=B, epsilon=EPSILON):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.doc_len: List[int] = []
        self.idf: Dict[str, float] = {}
        nd: Dict[str, int] = {}
        num_tokens = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_tokens += len(document)
            freqs: Dict[str, int] = {}
            for word in document:
                freqs[word] = freqs.get(word, 0) + 1
            self.doc_freqs.append(freqs)
            for word in freqs:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_tokens / self.corpus_size
        self.nd = nd
        # _calc_idf: sequential sum in vocabulary insertion order
        idf_sum = 0.0
        negative = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (
                q_freq * (self.k1 + 1)
                / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl))
            )
        return score


class BM25OkapiCSR:
    """Vectorised form over term-major CSR postings.  Same float64 operations in
    the same order per (token, document), so scores are bit-identical to
    ``BM25OkapiLiteral`` (checked in tests/test_oracle_bm25.py).

    Built either from token lists (``from_corpus``) or straight from integer
    postings (``from_postings``; used for the synthetic Zipf corpora where
    there are no strings).  Term ids follow vocabulary insertion order, i.e.
    first appearance scanning documents in order -- the order rank_bm25's
    ``nd`` dict has, which fixes the summation order of ``average_idf``.
    """

    def __init__(self, n_docs, doc_len, term_ptr, post_doc, post_tf, vocab=None,
                 k1=K1, b=B, epsilon=EPSILON):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = int(n_docs)
        self.doc_len = np.asarray(doc_len, dtype=np.int64)
        self.term_ptr = np.asarray(term_ptr, dtype=np.int64)
        self.post_doc = np.asarray(post_doc, dtype=np.int64)
        self.post_tf = np.asarray(post_tf, dtype=np.int64)
        self.vocab = vocab  # dict str -> term id, or None
        self.avgdl = int(self.doc_len.sum()) / self.corpus_size
        df = np.diff(self.term_ptr)
        n = self.corpus_size
        raw = np.array([math.log(n - int(f) + 0.5) - math.log(int(f) + 0.5) for f in df],
                       dtype=np.float64)
        idf_sum = 0.0
        for v in raw:          # sequential, insertion order
            idf_sum += float(v)
        self.average_idf = idf_sum / len(raw)
        eps = self.epsilon * self.average_idf
        self.idf = np.where(raw < 0, eps, raw)
        # query-independent per-document term k1*(1 - b + b*len/avgdl)
        self.doc_norm = self.k1 * (1 - self.b + self.b * self.doc_len / self.avgdl)

    # -- builders ---------------------------------------------------------
    @classmethod
    def from_corpus(cls, corpus: Sequence[Sequence[str]], **kw):
        vocab: Dict[str, int] = {}
        rows_t, rows_d, rows_f = [], [], []
        doc_len = []
        for d, document in enumerate(corpus):
            doc_len.append(len(document))
            freqs: Dict[str, int] = {}
            for word in document:
                freqs[word] = freqs.get(word, 0) + 1
            for word, f in freqs.items():
                t = vocab.get(word)
                if t is None:
                    t = vocab[word] = len(vocab)
                rows_t.append(t); rows_d.append(d); rows_f.append(f)
        t = np.asarray(rows_t, dtype=np.int64)
        d = np.asarray(rows_d, dtype=np.int64)
        f = np.asarray(rows_f, dtype=np.int64)
        order = np.lexsort((d, t))           # term-major, doc id ascending
        t, d, f = t[order], d[order], f[order]
        term_ptr = np.zeros(len(vocab) + 1, dtype=np.int64)
        np.add.at(term_ptr, t + 1, 1)
        term_ptr = np.cumsum(term_ptr)
        return cls(len(corpus), doc_len, term_ptr, d, f, vocab=vocab, **kw)

    @classmethod
    def from_postings(cls, n_docs, doc_len, term_ptr, post_doc, post_tf, **kw):
        return cls(n_docs, doc_len, term_ptr, post_doc, post_tf, vocab=None, **kw)

    # -- queries ----------------------------------------------------------
    def term_ids(self, query: Sequence[str]) -> List[int]:
        """Tokens -> term ids, -1 for out-of-vocabulary (adds 0, as ``idf.get(q) or 0``)."""
        return [self.vocab.get(q, -1) for q in query]

    def get_scores_ids(self, term_ids: Sequence[int]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        for t in term_ids:
            if t < 0:
                continue
            idf = self.idf[t]
            if not idf:          # `or 0`: an exact-zero idf contributes nothing
                continue
            lo, hi = self.term_ptr[t], self.term_ptr[t + 1]
            docs = self.post_doc[lo:hi]
            tf = self.post_tf[lo:hi]
            contrib = idf * (tf * (self.k1 + 1) / (tf + self.doc_norm[docs]))
            score[docs] += contrib          # doc ids unique within one term
        return score

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        return self.get_scores_ids(self.term_ids(query))


def max_or_one(scores: np.ndarray) -> float:
    """``max(bm25_scores) if max(bm25_scores) > 0 else 1.0`` -- retrieval_engine.py:74."""
    m = float(scores.max()) if len(scores) else 0.0
    return m if m > 0 else 1.0


def topk_positive(scores: np.ndarray, K: int, id_base: int = 0):
    """BM25 ranked list for RRF: documents with score > 0, best first, ties by
    ascending id, at most K.  (Build definition)"""
    idx = np.nonzero(scores > 0)[0]
    order = np.lexsort((idx, -scores[idx]))[:K]
    sel = idx[order]
    return scores[sel], sel + id_base
