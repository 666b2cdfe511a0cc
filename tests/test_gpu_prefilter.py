"""The int8 pre-filter of K2a (include/lrx.h: lrx_build_dense_prefilter): the scan streams an int8
shadow of the matrix, the exact re-score reads the fp16 rows, and NOTHING in the results may
change -- ids and float64 scores stay bit-identical to the oracle (oracle/flat_ip.py) and to the
plain fp16 scan.  Also: the shadow's bytes and its two error bounds against their CPU
restatement, and the rigour of the guard band on data built to defeat it."""
import numpy as np
import pytest
import torch

from oracle import bm25 as obm25
from oracle import flat_ip
from oracle.search import OracleIndex

from legal_rag_engine_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from legal_rag_engine_b200.device_index import DeviceIndex
    d = DeviceIndex(0)
    yield d
    d.close()


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000])
def test_shadow_bytes_and_bounds_match_the_restatement(dev, n):
    x = synth.host_vectors(n, seed=300 + n, dup_frac=0.0)
    if n > 2:
        x[1] = 0                                            # an all-zero row: scale 0, digits 0
    dev.set_corpus(_cuda(x), 0, prefilter=True)
    buf, nbytes, (E, X) = dev._q8
    n_pad = (n + 127) // 128 * 128
    assert nbytes == n_pad * 388
    raw = buf.cpu().numpy()
    rows = raw[:n_pad * 384].view(np.int8).reshape(n_pad, 384).copy()
    # device layout: odd rows keep their 64-byte chunks swapped pairwise (bank-conflict-free scan)
    rows[1::2] = rows[1::2].reshape(-1, 3, 2, 64)[:, :, ::-1, :].reshape(-1, 384)
    scales = raw[n_pad * 384:n_pad * 388].view(np.float32)
    xi, sc, Eo, Xo = flat_ip.int8_shadow(x)
    np.testing.assert_array_equal(rows[:n], xi)
    np.testing.assert_array_equal(scales[:n], sc)
    assert not rows[n:].any() and not scales[n:].any()      # padding of the last tile
    assert Eo <= E <= Eo * (1 + 1e-6) + 1e-300
    assert Xo <= X <= Xo * (1 + 1e-6) + 1e-300


def test_guard_band_covers_every_row(dev):
    """|exact - fast| <= band for every (row, query) of a sample -- the inequality the exactness
    argument rests on -- with the bounds the DEVICE reported."""
    n = 20000
    x = synth.host_vectors(n, seed=11, dup_frac=0.0)
    q = synth.host_queries(4, seed=12)
    q[0] = synth.host_planted_queries(x, [77], seed=1)[0]
    dev.set_corpus(_cuda(x), 0, prefilter=True)
    E, X = dev.prefilter_bounds
    s = flat_ip.exact_scores(x, q)
    for b in range(4):
        band = flat_ip.int8_guard_band(q[b], E, X)
        fast = flat_ip.int8_fast_scores(x, q[b]).astype(np.float64)
        assert np.abs(s[b] - fast).max() <= band
        assert band < 0.02                                   # ~0.0085 for unit vectors


@pytest.mark.parametrize("n", [65, 129, 1000, 20011, 300000])
@pytest.mark.parametrize("B", [1, 3, 4, 5, 8])            # 5..8: two digit tiles per pass
def test_prefilter_on_and_off_give_the_oracle_result(dev, n, B):
    x = synth.host_vectors(n, seed=500 + n, dup_frac=0.01)
    q = synth.host_queries(B, seed=17 + B)
    q[0] = synth.host_planted_queries(x, [n // 2], seed=2)[0]
    s = flat_ip.exact_scores(x, q)
    xd, qd = _cuda(x), _cuda(q)
    for on in (True, False):
        dev.set_corpus(xd, id_base=31, prefilter=on)
        assert (dev.prefilter_bounds is not None) == on
        for K in (1, 10, 20, 64, 200):                      # 200 on >= 32768 rows: lists of 64, merged list of 2048
            Eo, Do, Io = flat_ip.topk_from_scores(s, K, id_base=31)
            E, D, I, flags = dev.dense_topk(qd, K)
            assert flags.cpu().numpy().sum() == 0, (on, K)
            np.testing.assert_array_equal(I.cpu().numpy(), Io)
            np.testing.assert_array_equal(E.cpu().numpy(), Eo)
            np.testing.assert_array_equal(D.cpu().numpy(), Do)


def _near_duplicates(n, n_close, seed, noise=0.03):
    """n rows of which n_close are row 0 plus N(0, noise^2 / dim) noise, renormalised: against a query
    planted on row 0 their exact scores spread by ~1.5e-3 -- far apart for the fp16 scan's 1e-5
    band, all inside the int8 band (~8.5e-3)."""
    rng = np.random.default_rng(seed)
    x = synth.host_vectors(n, seed=seed, dup_frac=0.0)
    where = rng.choice(np.arange(1, n), size=n_close, replace=False)
    r = x[0].astype(np.float32)[None, :] + rng.standard_normal((n_close, 384), dtype=np.float32) * (noise / np.sqrt(384.0))
    r /= np.linalg.norm(r, axis=1, keepdims=True)
    x[where] = r.astype(np.float16)
    return x


def test_unseparable_candidates_set_the_flag_and_the_search_falls_back(dev):
    """700 near-duplicates of the best row sit inside the int8 band: the pre-filtered scan must SAY
    so (flag 1 at its widths) and the host search must still return the exact result -- through the
    widening retry, which ends on the fp16 scan (band 1e-5)."""
    n, k = 30000, 10
    x = _near_duplicates(n, 700, seed=21)
    q = synth.host_planted_queries(x, [0], seed=4)
    dev.set_corpus(_cuda(x), 0, prefilter=True)
    for width in (64, 128):
        _, _, _, flags = dev.dense_topk(_cuda(q), 2 * k, width=width)
        assert flags.item() == 1, width                      # 700 rows within the band of the 20th score
    E, D, I, flags = dev.dense_topk(_cuda(q), 2 * k, width=256)   # fp16 scan: separable
    Eo, Do, Io = flat_ip.topk_from_scores(flat_ip.exact_scores(x, q), 2 * k)
    assert flags.item() == 0
    np.testing.assert_array_equal(I.cpu().numpy(), Io)
    np.testing.assert_array_equal(E.cpu().numpy(), Eo)
    # the whole search, host buffers: retried inside the call
    idx = synth.host_bm25(n, seed=22, vocab=3000)
    dev.set_postings(idx.term_ptr, idx.postings, idx.doc_len, idx.idf, idx.avgdl)
    csr = obm25.BM25OkapiCSR.from_postings(n, idx.doc_len, idx.term_ptr.astype(np.int64),
                                           idx.postings[:, 0], idx.postings[:, 1])
    terms, ptr = synth.host_query_terms(1, 8, seed=23, vocab=3000)
    lists = [terms[ptr[0]:ptr[1]].tolist()]
    for fusion in ("linear", "rrf"):
        ids, score, sem, kw = dev.search_batch_host(q, lists, k, [0.5], fusion)
        want = OracleIndex(x, csr).search_batch_vec(q, lists, k, [0.5], fusion)
        assert ids[0].tolist() == [r[0] for r in want[0]]
        assert score[0].tolist() == [r[1] for r in want[0]]


@pytest.mark.parametrize("fusion", ["linear", "rrf"])
@pytest.mark.parametrize("B", [4, 8])
def test_search_with_and_without_prefilter_bit_identical(dev, B, fusion):
    n, k = 150000, 10
    x = synth.host_vectors(n, seed=31)
    idx = synth.host_bm25(n, seed=32, vocab=4000)
    q = synth.host_queries(B, seed=33)
    terms, ptr = synth.host_query_terms(B, 8, seed=34, vocab=4000)
    lists = [terms[ptr[b]:ptr[b + 1]].tolist() for b in range(B)]
    w = [0.5, 0.6, 0.5, 0.6] * (B // 4)
    xd = _cuda(x)
    out = {}
    for on in (True, False):
        dev.set_corpus(xd, 0, prefilter=on)
        dev.set_postings(idx.term_ptr, idx.postings, idx.doc_len, idx.idf, idx.avgdl)
        res = [dev.search_batch_host(q, lists, k, w, fusion) for _ in range(3)]   # direct, capture, replay
        for r in res[1:]:
            for a, b in zip(res[0], r):
                np.testing.assert_array_equal(a, b)
        out[on] = res[0]
    for a, b in zip(out[True], out[False]):
        np.testing.assert_array_equal(a, b)
    csr = obm25.BM25OkapiCSR.from_postings(n, idx.doc_len, idx.term_ptr.astype(np.int64),
                                           idx.postings[:, 0], idx.postings[:, 1])
    want = OracleIndex(x, csr).search_batch_vec(q, lists, k, w, fusion)
    for b in range(B):
        assert out[True][0][b].tolist() == [r[0] for r in want[b]]
        assert out[True][1][b].tolist() == [r[1] for r in want[b]]


def test_clone_view_shares_the_shadow(dev):
    n = 40000
    x = synth.host_vectors(n, seed=41)
    q = synth.host_queries(4, seed=42)
    dev.set_corpus(_cuda(x), 5, prefilter=True)
    other = dev.clone_view()
    try:
        assert other._q8 is dev._q8
        a = dev.dense_topk(_cuda(q), 20)
        b = other.dense_topk(_cuda(q), 20)
        torch.cuda.synchronize()
        for u, v in zip(a, b):
            assert torch.equal(u, v)
        Eo, Do, Io = flat_ip.topk_from_scores(flat_ip.exact_scores(x, q), 20, id_base=5)
        np.testing.assert_array_equal(a[2].cpu().numpy(), Io)
    finally:
        other.close()
