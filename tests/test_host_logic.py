"""CPU tests of the host-side product code (index builder, synthetic generators,
C-ABI symbol coverage) against the oracle.  No compute on the GPU here."""
import ctypes

import numpy as np
import pytest

from oracle import bm25 as obm25
from oracle import flat_ip, fusion
from oracle.search import OracleIndex, fanout_queries, fanout_dedup

from legal_rag_engine_b200 import _lib, synth
from legal_rag_engine_b200.bm25_index import BM25Index, tokenize


def test_library_exports_every_declared_symbol():
    declared = _lib.declared_symbols()
    assert "lrx_search_batch_host" in declared and "lrx_dense_topk" in declared
    assert sorted(_lib._SIGNATURES) == declared, "ctypes table out of sync with include/lrx.h"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"liblrx.so does not export {name}"
    _lib.load()
    assert b"sm_100a" in _lib.load().lrx_version()


def test_open_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.lrx_config(0, 384, 0, 1)
    h = ctypes.c_void_p()
    rc = lib.lrx_open(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2                       # LRX_E_DEVICE: no CPU fallback
    assert b"no CPU fallback" in lib.lrx_last_error(None)
    from legal_rag_engine_b200.device_index import DeviceIndex
    with pytest.raises(RuntimeError):
        DeviceIndex(0)


def test_builder_matches_oracle_on_real_corpus(legal_texts, reference_queries):
    idx = BM25Index.from_texts(legal_texts)
    csr = obm25.BM25OkapiCSR.from_corpus([obm25.tokenize(t) for t in legal_texts])
    assert idx.n_docs == csr.corpus_size and idx.nnz == len(csr.post_doc)
    assert idx.vocab == csr.vocab
    assert idx.avgdl == csr.avgdl
    np.testing.assert_array_equal(idx.idf, csr.idf)           # bit for bit
    np.testing.assert_array_equal(idx.term_ptr.astype(np.int64), csr.term_ptr)
    np.testing.assert_array_equal(idx.postings[:, 0].astype(np.int64), csr.post_doc)
    np.testing.assert_array_equal(idx.postings[:, 1].astype(np.int64), csr.post_tf)
    for q in reference_queries:
        assert idx.term_ids(tokenize(q)) == csr.term_ids(obm25.tokenize(q))


def test_shard_keeps_global_statistics():
    idx = synth.host_bm25(3000, seed=5, vocab=500)
    a, b = idx.shard(0, 1700), idx.shard(1700, 3000)
    assert a.nnz + b.nnz == idx.nnz
    assert a.avgdl == idx.avgdl and b.idf is idx.idf
    # a shard's postings are the global ones restricted and re-based
    d = idx.postings[:, 0].astype(np.int64)
    np.testing.assert_array_equal(b.postings[:, 0].astype(np.int64), d[d >= 1700] - 1700)


def test_synth_shapes_and_duplicates():
    x = synth.host_vectors(5000, seed=3)
    assert x.dtype == np.float16 and x.shape == (5000, 384)
    n = np.linalg.norm(x.astype(np.float64), axis=1)
    assert np.all(np.abs(n - 1) < 2e-3)
    # 0.1 % duplicated rows exist
    _, counts = np.unique(x.view(np.uint16).reshape(5000, -1), axis=0, return_counts=True)
    assert (counts > 1).sum() >= 1
    idx = synth.host_bm25(2000, seed=7, vocab=1000)
    assert idx.n_terms == 1000 and 8 <= idx.doc_len.min() and idx.doc_len.max() <= 512
    t, p = synth.host_query_terms(4, 8, vocab=1000)
    assert t.shape == (32,) and p.tolist() == [0, 8, 16, 24, 32]


def test_flat_ip_oracle_exact_and_tiebreak():
    x = synth.host_vectors(4000, seed=11, dup_frac=0.01)
    q = synth.host_queries(3, seed=12)
    s = flat_ip.exact_scores(x, q)
    # exactness: any summation order gives the same float64
    s2 = np.stack([sum((float(a) * float(b) for a, b in zip(q[0][::-1], r[::-1])), 0.0)
                   for r in x[:50]])
    np.testing.assert_array_equal(s[0, :50], s2)
    E, D, I = flat_ip.topk_from_scores(s, 20)
    for b in range(3):
        order = np.lexsort((np.arange(4000), -s[b]))[:20]
        np.testing.assert_array_equal(I[b], order)
        assert D.dtype == np.float32
    # K > N pads with -1
    E, D, I = flat_ip.topk_from_scores(s[:, :5], 8)
    assert (I[:, 5:] == -1).all() and (I[:, :5] >= 0).all()


def test_linear_fusion_follows_reference_semantics():
    D = np.array([0.9, 0.8, 0.8, 0.1], dtype=np.float32)
    I = np.array([7, 3, 5, -1], dtype=np.int64)
    bm = np.zeros(10); bm[3] = 2.0; bm[5] = 2.0; bm[9] = 4.0      # 9 is BM25-only: never a candidate
    out = fusion.linear_fuse(D, I, bm, 4.0, 2, 0.5)
    # 3 and 5 tie on the fused score -> stable sort keeps flat-IP order (3 before 5)
    assert [r[0] for r in out] == [3, 5]
    assert out[0][1] == float(np.float32(0.8)) * 0.5 + 0.5 * 0.5
    out = fusion.linear_fuse(D, I, bm, 4.0, 5, 0.5)
    assert [r[0] for r in out] == [3, 5, 7]                        # -1 skipped


def test_rrf_oracle():
    dense = [(4, 0.9, 1.0), (2, 0.8, 0.0), (9, 0.7, 3.0)]
    sparse = [(9, 0.7, 3.0), (5, 0.1, 2.0), (4, 0.9, 1.0)]
    out = fusion.rrf_fuse(dense, sparse, 3.0, 3)
    assert [r[0] for r in out] == [4, 9, 5] or [r[0] for r in out][:2] == [4, 9]
    s4 = 1.0 / 61 + 1.0 / 63
    s9 = 1.0 / 63 + 1.0 / 61
    assert out[0][1] == s4 and out[1][1] == s9 and out[0][0] == 4   # tie -> lower id first


def test_sharded_oracle_equals_unsharded():
    """Property the multi-GPU path relies on: per-shard top-2k + global statistics,
    merged by (score desc, id asc), reproduce the unsharded search exactly."""
    n = 6000
    x = synth.host_vectors(n, seed=21)
    idx = synth.host_bm25(n, seed=22, vocab=2000)
    csr = obm25.BM25OkapiCSR.from_postings(n, idx.doc_len, idx.term_ptr.astype(np.int64),
                                           idx.postings[:, 0], idx.postings[:, 1])
    np.testing.assert_array_equal(csr.idf, idx.idf)
    full = OracleIndex(x, csr)
    q = synth.host_queries(2, seed=23)
    terms, ptr = synth.host_query_terms(2, 8, seed=24, vocab=2000)
    k = 10
    for b in range(2):
        tl = terms[ptr[b]:ptr[b + 1]].tolist()
        ref = full.search_vec(q[b], tl, k, 0.6, "linear")
        # shards
        cuts = [0, 2500, 6000]
        recs, maxes = [], []
        bm_full = csr.get_scores_ids(tl)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            s = flat_ip.exact_scores(x[lo:hi], q[b][None])
            E, D, I = flat_ip.topk_from_scores(s, 2 * k, id_base=lo)
            recs += [(int(i), float(e)) for e, i in zip(E[0], I[0]) if i >= 0]
            maxes.append(float(bm_full[lo:hi].max()))
        recs.sort(key=lambda r: (-r[1], r[0]))
        recs = recs[:2 * k]
        D = np.array([np.float32(e) for _, e in recs], dtype=np.float32)
        I = np.array([i for i, _ in recs], dtype=np.int64)
        mx = max(maxes); mx = mx if mx > 0 else 1.0
        got = fusion.linear_fuse(D, I, bm_full, mx, k, 0.6)
        assert got == ref


def test_fanout_restatement():
    qs, ws = fanout_queries("I was robbed", "victim_distress", ["Robbery"], "offence")
    assert len(qs) == 4 and qs[1] == "How to file FIR for Robbery BNSS procedure"
    assert ws == [0.5, 0.6, 0.5, 0.6]
    qs, ws = fanout_queries("what is the procedure for bail", "informational", [], "general")
    assert qs == ["what is the procedure for bail"] and ws == [0.6]
    res = fanout_dedup([[(0, 1.0), (1, 0.9)], [(1, 0.8), (2, 0.7), (3, 0.6)]],
                       {0: "A", 1: "B", 2: "", 3: "A"})
    assert [r[0] for r in res] == [0, 1]
