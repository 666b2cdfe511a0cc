"""Store formats (host logic, CPU): native round trip, FAISS flat image, the reference's
pickled BM25Okapi read through the stub unpickler."""
import json
import pickle
import sys
import types

import numpy as np

from legal_rag_engine_b200 import store
from legal_rag_engine_b200.bm25_index import BM25Index
from legal_rag_engine_b200.engine import fanout_queries, merge_fanout
from oracle import bm25 as obm25
from oracle import search as osearch


def _mini():
    texts = ["zero fir can be lodged at any police station", "victim compensation scheme nalsa",
             "procedure after arrest of a suspect", "zero zero fir fir procedure"]
    chunks = [{"text": t, "metadata": {"law": "BNSS", "section": str(i)}, "canonical_header": f"H{i}"}
              for i, t in enumerate(texts)]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 384)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return texts, chunks, x


def test_native_round_trip(tmp_path):
    texts, chunks, x = _mini()
    idx = BM25Index.from_texts(texts)
    store.save_store(tmp_path, chunks, x, idx)
    c2, xh, b2 = store.load_store(tmp_path)
    assert c2 == chunks
    np.testing.assert_array_equal(xh, x.astype(np.float16))
    np.testing.assert_array_equal(b2.term_ptr, idx.term_ptr)
    np.testing.assert_array_equal(b2.postings, idx.postings)
    np.testing.assert_array_equal(b2.idf, idx.idf)
    assert b2.vocab == idx.vocab and b2.avgdl == idx.avgdl
    # the flat image parses back to the float32 rows
    np.testing.assert_array_equal(store.read_faiss_flat(tmp_path / "index.faiss"), x)
    # metadata.json is the reference's exact serialisation (indent=2)
    assert (tmp_path / "metadata.json").read_text(encoding="utf-8") == json.dumps(chunks, indent=2)


def test_reference_store_interchange(tmp_path, monkeypatch):
    """A store as the reference writes it: index.faiss + bm25.pkl (pickled rank_bm25.BM25Okapi)."""
    texts, chunks, x = _mini()
    lit = obm25.BM25OkapiLiteral([t.lower().split() for t in texts])
    mod = types.ModuleType("rank_bm25")

    class BM25Okapi:            # stands in for the real class at pickling time only
        pass
    BM25Okapi.__module__ = "rank_bm25"
    BM25Okapi.__qualname__ = "BM25Okapi"
    mod.BM25Okapi = BM25Okapi
    monkeypatch.setitem(sys.modules, "rank_bm25", mod)
    obj = BM25Okapi()
    obj.corpus_size, obj.avgdl, obj.doc_freqs = lit.corpus_size, lit.avgdl, lit.doc_freqs
    obj.idf, obj.doc_len, obj.k1, obj.b, obj.epsilon = lit.idf, lit.doc_len, lit.k1, lit.b, lit.epsilon
    obj.average_idf, obj.tokenizer = lit.average_idf, None
    with open(tmp_path / "bm25.pkl", "wb") as f:
        pickle.dump(obj, f)
    monkeypatch.delitem(sys.modules, "rank_bm25")           # loading must not need the module
    store.write_faiss_flat_ip(tmp_path / "index.faiss", x)
    (tmp_path / "metadata.json").write_text(json.dumps(chunks, indent=2), encoding="utf-8")
    c2, xh, b2 = store.load_store(tmp_path)
    assert c2 == chunks and xh.dtype == np.float16 and xh.shape == (4, 384)
    csr = obm25.BM25OkapiCSR(b2.n_docs, b2.doc_len, b2.term_ptr.astype(np.int64), b2.postings[:, 0],
                             b2.postings[:, 1])
    for q in ("zero fir procedure", "nalsa victim", "nothing here"):
        toks = q.split()
        np.testing.assert_array_equal(csr.get_scores_ids(b2.term_ids(toks)), lit.get_scores(toks))
    np.testing.assert_array_equal(b2.idf, csr.idf)


def test_written_store_opens_through_the_reference_files(tmp_path):
    """save_store also writes what the reference reads (index.faiss + bm25.pkl): with the native
    files removed the store loads from those two and gives the same index; the pickle names
    rank_bm25.BM25Okapi and carries rank_bm25 0.2.2's attributes."""
    texts, chunks, x = _mini()
    idx = BM25Index.from_texts(texts)
    store.save_store(tmp_path, chunks, x, idx)
    raw = (tmp_path / "bm25.pkl").read_bytes()
    assert b"rank_bm25" in raw and b"BM25Okapi" in raw
    (tmp_path / "vectors.f16.npy").unlink()
    (tmp_path / "bm25.npz").unlink()
    c2, xh, b2 = store.load_store(tmp_path)
    np.testing.assert_array_equal(xh, x.astype(np.float16))
    np.testing.assert_array_equal(b2.term_ptr, idx.term_ptr)
    np.testing.assert_array_equal(b2.postings, idx.postings)
    np.testing.assert_array_equal(b2.idf, idx.idf)
    assert b2.vocab == idx.vocab and b2.avgdl == idx.avgdl
    lit = obm25.BM25OkapiLiteral([t.lower().split() for t in texts])
    obj = store._BM25Unpickler(__import__("io").BytesIO(raw)).load()
    assert obj.doc_freqs == lit.doc_freqs and obj.doc_len == lit.doc_len and obj.idf == lit.idf
    assert obj.average_idf == lit.average_idf and obj.corpus_size == 4


def test_crafted_pickle_cannot_run_code(tmp_path):
    import pytest

    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('os').getpid()",))
    for payload in (Evil(), {"x": Evil()}):
        (tmp_path / "bm25.pkl").write_bytes(pickle.dumps(payload))
        with pytest.raises(pickle.UnpicklingError, match="refusing"):
            store.bm25_from_reference_pickle(tmp_path / "bm25.pkl")
    import numpy
    (tmp_path / "bm25.pkl").write_bytes(pickle.dumps(numpy.load))
    with pytest.raises(pickle.UnpicklingError, match="refusing"):
        store.bm25_from_reference_pickle(tmp_path / "bm25.pkl")


def test_fanout_helpers_match_oracle():
    for ctx, ents, cat in (("victim_distress", ["Robbery"], "procedure"), ("victim_distress", [], "x"),
                           ("informational", ["theft"], "general")):
        assert fanout_queries("I was robbed, what is the procedure?", ctx, ents, cat) == \
            osearch.fanout_queries("I was robbed, what is the procedure?", ctx, ents, cat)
    a = [{"chunk": {"canonical_header": "A"}, "score": 1}, {"chunk": {"canonical_header": ""}, "score": 2}]
    b = [{"chunk": {"canonical_header": "A"}, "score": 3}, {"chunk": {"canonical_header": "B"}, "score": 4}]
    assert [r["score"] for r in merge_fanout([a, b])] == [1, 4]
