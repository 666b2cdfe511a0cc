"""Oracle self-consistency: the literal dict-loop restatement of rank_bm25 and the
vectorised CSR form must agree bit for bit; golden numbers of the real corpus."""
import json
import math

import numpy as np
import pytest

from oracle import bm25 as obm25


def test_literal_vs_csr_small():
    corpus = [["a", "b", "a", "c"], ["b", "b", "d"], [], ["a"], ["e", "a", "b", "c", "d", "e"]]
    lit = obm25.BM25OkapiLiteral(corpus)
    csr = obm25.BM25OkapiCSR.from_corpus(corpus)
    assert lit.avgdl == csr.avgdl
    for q in (["a"], ["b", "a", "a"], ["zzz"], [], ["e", "d", "c", "b", "a"]):
        np.testing.assert_array_equal(lit.get_scores(q), csr.get_scores(q))


def test_negative_idf_floor():
    # 'x' is in every document -> raw idf < 0 -> epsilon * average idf
    corpus = [["x", "a"], ["x", "b"], ["x", "c"], ["x", "d", "e"]]
    lit = obm25.BM25OkapiLiteral(corpus)
    raw = math.log(4 - 4 + 0.5) - math.log(4 + 0.5)
    assert raw < 0
    assert lit.idf["x"] == 0.25 * lit.average_idf
    csr = obm25.BM25OkapiCSR.from_corpus(corpus)
    assert csr.idf[csr.vocab["x"]] == lit.idf["x"]
    np.testing.assert_array_equal(lit.get_scores(["x", "a"]), csr.get_scores(["x", "a"]))


def test_real_corpus_stats_and_parity(legal_texts, reference_queries):
    corpus = [obm25.tokenize(t) for t in legal_texts]
    csr = obm25.BM25OkapiCSR.from_corpus(corpus)
    # SURVEY.md section 8a row A6 (probed on the reference corpus)
    assert csr.corpus_size == 2620
    assert len(csr.vocab) == 12630
    assert len(csr.post_doc) == 168899
    assert abs(csr.avgdl - 104.3557) < 1e-3
    assert abs(csr.average_idf - 6.6722) < 1e-3
    lit = obm25.BM25OkapiLiteral(corpus)
    assert lit.avgdl == csr.avgdl and lit.average_idf == csr.average_idf
    for q in reference_queries[:6]:
        toks = obm25.tokenize(q)
        np.testing.assert_array_equal(lit.get_scores(toks), csr.get_scores(toks))


def test_golden_bm25(legal_texts):
    """Committed known answers (tests/golden/make_golden.py) for the real corpus."""
    from conftest import GOLDEN
    gold = json.loads((GOLDEN / "bm25_real_corpus.json").read_text())
    csr = obm25.BM25OkapiCSR.from_corpus([obm25.tokenize(t) for t in legal_texts])
    for case in gold["cases"]:
        s = csr.get_scores(obm25.tokenize(case["query"]))
        assert float(s.max()).hex() == case["max_hex"]
        top = np.lexsort((np.arange(len(s)), -s))[:10]
        assert [int(i) for i in top] == case["top10_ids"]
        assert [float(s[i]).hex() for i in top] == case["top10_scores_hex"]
