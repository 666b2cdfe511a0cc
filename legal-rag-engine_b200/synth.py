"""Seeded synthetic corpora of the benchmark shapes.

Two generators with the same distributions:

* ``host_*``   numpy, exactly reproducible -- parity runs (N <= ~1 M) generate here
  and upload, so the CPU oracle and the GPU see identical bits;
* ``device_*`` torch on the GPU -- throughput runs at 10 M / 100 M rows, where the
  corpus would not fit (or take minutes to build) on the host.

Vectors : x = fp16(normalize(N(0, I_384))); 0.1 % of the rows are exact duplicates
          of earlier rows (exercises the tie-break).
BM25    : V = 50 000 terms, term rank r drawn with P ~ 1/r (Zipf s = 1); document
          length clip(round(lognormal(ln 80, 0.73)), 8, 512).
Queries : unit Gaussian vectors; 8 Zipf term draws (duplicates allowed).
"""
from __future__ import annotations

import numpy as np

from .bm25_index import BM25Index

DIM = 384
VOCAB = 50_000


# --------------------------------------------------------------------- host
def host_vectors(n: int, seed: int = 1234, dup_frac: float = 0.001, dim: int = DIM) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    xh = x.astype(np.float16)
    n_dup = int(n * dup_frac)
    if n_dup and n > 1:
        dst = rng.integers(1, n, size=n_dup)
        src = (rng.random(n_dup) * dst).astype(np.int64)      # an earlier row
        xh[dst] = xh[src]
    return xh


def host_queries(b: int, seed: int = 4321, dim: int = DIM) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((b, dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float16)


def host_planted_queries(xh: np.ndarray, rows, seed: int = 4321, noise: float = 0.1) -> np.ndarray:
    """Queries = a corpus row + N(0, noise^2/dim) noise, renormalised: top-1 is known."""
    rng = np.random.default_rng(seed)
    q = xh[np.asarray(rows)].astype(np.float32)
    q += rng.standard_normal(q.shape, dtype=np.float32) * (noise / np.sqrt(q.shape[1]))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float16)


def _zipf_cdf(vocab: int) -> np.ndarray:
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    c = np.cumsum(w)
    return c / c[-1]


def host_doc_lengths(n: int, rng) -> np.ndarray:
    return np.clip(np.rint(rng.lognormal(np.log(80.0), 0.73, size=n)), 8, 512).astype(np.int64)


def host_bm25(n_docs: int, seed: int = 777, vocab: int = VOCAB) -> BM25Index:
    rng = np.random.default_rng(seed)
    lens = host_doc_lengths(n_docs, rng)
    cdf = _zipf_cdf(vocab)
    total = int(lens.sum())
    terms = np.searchsorted(cdf, rng.random(total), side="right").astype(np.int64)
    np.minimum(terms, vocab - 1, out=terms)
    docs = np.repeat(np.arange(n_docs, dtype=np.int64), lens)
    return BM25Index.from_token_arrays(docs, terms, n_docs, vocab)


def host_query_terms(b: int, n_terms: int = 8, seed: int = 999, vocab: int = VOCAB):
    """b queries x n_terms Zipf draws -> (flat int32 term ids, int32 ptr [b+1])."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(vocab)
    t = np.searchsorted(cdf, rng.random(b * n_terms), side="right")
    t = np.minimum(t, vocab - 1).astype(np.int32)
    ptr = (np.arange(b + 1) * n_terms).astype(np.int32)
    return t, ptr


# ------------------------------------------------------------------ encoder
BERT = dict(vocab=30522, hidden=384, layers=6, heads=12, ffn=1536, max_pos=512)


def bert_state_dict(seed: int = 42, std: float = 0.02, vocab: int = BERT["vocab"],
                    ln_jitter: float = 0.0):
    """Seeded random weights in HuggingFace ``BertModel`` state_dict naming (float32 numpy):
    N(0, std^2) matrices/biases, LayerNorm gamma 1 / beta 0.  ``ln_jitter`` > 0
    perturbs gamma/beta so the affine part of every LayerNorm is exercised too."""
    rng = np.random.default_rng(seed)
    H, F = BERT["hidden"], BERT["ffn"]
    n = lambda *shape: (rng.standard_normal(shape, dtype=np.float32) * np.float32(std))

    def ln(prefix, sd):
        sd[prefix + ".weight"] = (1.0 + ln_jitter * rng.standard_normal(H)).astype(np.float32)
        sd[prefix + ".bias"] = (ln_jitter * rng.standard_normal(H)).astype(np.float32)

    sd = {"embeddings.word_embeddings.weight": n(vocab, H),
          "embeddings.position_embeddings.weight": n(BERT["max_pos"], H),
          "embeddings.token_type_embeddings.weight": n(2, H)}
    ln("embeddings.LayerNorm", sd)
    for l in range(BERT["layers"]):
        p = f"encoder.layer.{l}."
        for name, (o, i) in (("attention.self.query", (H, H)), ("attention.self.key", (H, H)),
                             ("attention.self.value", (H, H)), ("attention.output.dense", (H, H)),
                             ("intermediate.dense", (F, H)), ("output.dense", (H, F))):
            sd[p + name + ".weight"] = n(o, i)
            sd[p + name + ".bias"] = n(o)
        ln(p + "attention.output.LayerNorm", sd)
        ln(p + "output.LayerNorm", sd)
    return sd


def token_batch(b: int, s: int, seed: int = 7, vocab: int = BERT["vocab"], full: bool = False):
    """ids int32 [b,s] ([CLS]=101 ... [SEP]=102, 0-padded, body ids uniform in [1000, vocab))
    and lens int32 [b]: full length when ``full`` else uniform in [2, s]."""
    rng = np.random.default_rng(seed)
    lens = np.full(b, s, dtype=np.int32) if full else rng.integers(2, s + 1, size=b).astype(np.int32)
    ids = rng.integers(1000, vocab, size=(b, s)).astype(np.int32)
    ids[:, 0] = 101
    for i, n_ in enumerate(lens):
        ids[i, n_ - 1] = 102
        ids[i, n_:] = 0
    return ids, lens


# ------------------------------------------------------------------- device
def device_vectors(n: int, device, seed: int = 1234, dup_frac: float = 0.001, dim: int = DIM,
                   chunk: int = 1 << 20):
    """fp16 [n, dim] unit rows generated on the GPU in chunks (Philox)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float16, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.randn((e - s, dim), generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x.to(torch.float16)
    n_dup = int(n * dup_frac)
    if n_dup and n > 1:
        dst = torch.randint(1, n, (n_dup,), generator=g, device=device)
        src = (torch.rand(n_dup, generator=g, device=device) * dst).long()
        out[dst] = out[src]
    return out


def device_bm25(n_docs: int, device, seed: int = 777, vocab: int = VOCAB, doc_chunk: int = 1 << 20):
    """Zipf postings built on the GPU.  Returns a dict of device tensors
    (term_ptr u64-as-int64 [V+1], postings int32 [nnz,2], doc_len int32 [n]) plus the
    host-side statistics (idf float64 [V], avgdl)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cdf = torch.from_numpy(_zipf_cdf(vocab)).to(device)
    lens_all, keys_all, tf_all = [], [], []
    for s in range(0, n_docs, doc_chunk):
        e = min(n_docs, s + doc_chunk)
        m = e - s
        ln = torch.empty(m, device=device, dtype=torch.float32).log_normal_(float(np.log(80.0)), 0.73,
                                                                            generator=g)
        lens = ln.round().clamp_(8, 512).long()
        total = int(lens.sum().item())
        u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
        terms = torch.searchsorted(cdf, u, right=True).clamp_(max=vocab - 1)
        docs = torch.repeat_interleave(torch.arange(s, e, device=device), lens)
        key = terms * n_docs + docs                       # term-major, doc ascending
        uniq, tf = torch.unique(key, return_counts=True)  # sorted
        lens_all.append(lens); keys_all.append(uniq); tf_all.append(tf)
        del u, terms, docs, key
    doc_len = torch.cat(lens_all)
    key = torch.cat(keys_all); tf = torch.cat(tf_all)
    if len(keys_all) > 1:
        key, order = torch.sort(key)
        tf = tf[order]
        del order
    t = key // n_docs
    d = key - t * n_docs
    df = torch.bincount(t, minlength=vocab)
    term_ptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    term_ptr[1:] = torch.cumsum(df, 0)
    postings = torch.stack([d.to(torch.int32), tf.to(torch.int32)], dim=1).contiguous()
    from .bm25_index import okapi_idf
    idf, _ = okapi_idf(df.cpu().numpy(), n_docs)
    avgdl = int(doc_len.sum().item()) / n_docs
    return {"term_ptr": term_ptr, "postings": postings, "doc_len": doc_len.to(torch.int32),
            "idf": idf, "avgdl": avgdl, "n_terms": vocab, "nnz": int(postings.shape[0]),
            "df": df}
