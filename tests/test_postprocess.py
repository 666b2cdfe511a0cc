"""the vectorised result post-processing equals the literal restatement of
orchestrator.py:75-139 on the reference's own corpus, result for result and bit for bit."""
import copy
import random

import pytest

from legal_rag_engine_b200.postprocess import ResultPostProcessor
from oracle import postprocess as opp

INTENTS = [
    {"category": "procedure", "user_context": "victim_distress", "key_entities": ["robbery", "BNSS"], "sub_intent": None},
    {"category": "rights_of_victim", "user_context": "victim_distress", "key_entities": ["NALSA"], "sub_intent": "file FIR"},
    {"category": "rights_of_victim", "user_context": "victim_distress", "key_entities": [], "sub_intent": "compensation"},
    {"category": "definition", "user_context": "informational", "key_entities": ["bns", "sop"], "sub_intent": None},
    {"category": "punishment", "user_context": "professional", "key_entities": ["BSA", "BSA"], "sub_intent": "x"},
    {"category": "general_explanation", "user_context": "informational", "key_entities": [], "sub_intent": None},
]


def _results(chunks, rng, n):
    rows = [rng.randrange(len(chunks)) for _ in range(n)]
    rows += rows[:3]                                             # duplicates: expansion must skip them
    out = []
    for i in rows:
        s = rng.choice([0.25, 0.5, rng.random()])               # tied scores: the sort is stable
        out.append({"chunk": chunks[i], "score": s, "semantic": s, "keyword": 0.0})
    return out


@pytest.fixture(scope="module")
def pp(legal_chunks):
    return ResultPostProcessor(legal_chunks)


@pytest.mark.parametrize("intent", INTENTS)
def test_prioritize_and_expand_match_reference_restatement(legal_chunks, pp, intent):
    rng = random.Random(hash(intent["category"]) & 0xffff)
    lookup = opp.section_lookup(legal_chunks)
    for trial in range(20):
        base = _results(legal_chunks, rng, rng.choice([0, 1, 5, 20, 40]))
        a, b = [dict(r) for r in base], [dict(r) for r in base]
        want = opp.expand_results(opp.prioritize_results(a, intent)[:10], lookup)
        got = pp.expand(pp.prioritize(b, intent)[:10])
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g["chunk"] is w["chunk"]
            assert g["score"] == w["score"]                      # same float64 operations
            assert g.get("parent_context") == w.get("parent_context")


def test_parent_table_covers_every_sub_unit(legal_chunks, pp):
    lookup = opp.section_lookup(legal_chunks)
    n_sub = 0
    for i, c in enumerate(legal_chunks):
        meta = c.get("metadata", {})
        if meta.get("unit_type") in ["illustration", "explanation", "sub_section"]:
            n_sub += 1
            parent = lookup.get((meta.get("law"), meta.get("section")))
            assert (pp.parent[i] >= 0) == (parent is not None)
            if parent is not None:
                assert legal_chunks[pp.parent[i]] is parent
        else:
            assert pp.parent[i] == -1
    assert n_sub > 0 or True


def test_reference_quirk_sub_intent_none_raises(legal_chunks, pp):
    intent = {"category": "rights_of_victim", "user_context": "victim_distress", "key_entities": [],
              "sub_intent": None}
    res = [{"chunk": legal_chunks[0], "score": 1.0}]
    with pytest.raises(TypeError):
        opp.prioritize_results(copy.copy(res), intent)
    with pytest.raises(TypeError):
        pp.prioritize(copy.copy(res), intent)


def test_finish_merges_fanout_first(legal_chunks, pp):
    a = [{"chunk": legal_chunks[5], "score": 0.9}, {"chunk": legal_chunks[6], "score": 0.8}]
    b = [{"chunk": legal_chunks[6], "score": 0.95}, {"chunk": legal_chunks[7], "score": 0.1}]
    intent = INTENTS[-1]
    out = pp.finish([a, b], intent, k=5)
    heads = [r["chunk"].get("canonical_header") for r in out]
    assert len(heads) == len(set(heads))
    assert out[0]["chunk"] is legal_chunks[5]
