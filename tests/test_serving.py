"""Micro-batching front (legal-rag-engine_b200/serving.py): host logic on CPU with a recording
stand-in engine; the GPU test runs it over the real engine with concurrent client threads."""
import threading
import time

import pytest

from legal_rag_engine_b200.serving import MicroBatchingEngine


class RecordingEngine:
    def __init__(self, delay=0.002):
        self.calls = []
        self.delay = delay
        self.chunks = ["c0", "c1"]
        self.closed = False

    def search_batch(self, queries, k=5, hybrid_weights=None, fusion=None):
        self.calls.append((list(queries), k, list(hybrid_weights), fusion, threading.get_ident()))
        if any(q == "boom" for q in queries):
            raise ValueError("engine failure")
        time.sleep(self.delay)
        return [[{"chunk": {"q": q}, "score": w, "k": k, "fusion": fusion}] for q, w in zip(queries, hybrid_weights)]

    def close(self):
        self.closed = True


def test_concurrent_searches_are_coalesced_and_routed_back():
    eng = RecordingEngine()
    mb = MicroBatchingEngine(eng, max_batch=16, max_wait_ms=20.0)
    out = {}

    def client(i):
        out[i] = mb.search(f"q{i}", k=5, hybrid_weight=i / 100.0)
    ts = [threading.Thread(target=client, args=(i,)) for i in range(40)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(40):
        assert out[i] == [{"chunk": {"q": f"q{i}"}, "score": i / 100.0, "k": 5, "fusion": None}]
    assert sum(len(c[0]) for c in eng.calls) == 40
    assert len(eng.calls) < 40 and max(len(c[0]) for c in eng.calls) <= 16
    assert len({c[4] for c in eng.calls}) == 1            # the engine is only ever touched by one thread
    assert mb.requests == 40 and mb.batches == len(eng.calls) and mb.largest_batch > 1
    assert mb.chunks == ["c0", "c1"]                      # attribute passthrough
    mb.close()
    assert eng.closed
    with pytest.raises(RuntimeError):
        mb.search("late")


def test_groups_by_k_and_fusion_and_keeps_fanout_together():
    eng = RecordingEngine(delay=0.0)
    mb = MicroBatchingEngine(eng, max_batch=64, max_wait_ms=30.0)
    res = {}
    ts = [threading.Thread(target=lambda: res.setdefault("a", mb.search("a", k=5))),
          threading.Thread(target=lambda: res.setdefault("b", mb.search("b", k=10))),
          threading.Thread(target=lambda: res.setdefault("c", mb.search("c", k=5, fusion="rrf"))),
          threading.Thread(target=lambda: res.setdefault("f", mb.search_batch(["f1", "f2", "f3", "f4"], 5,
                                                                               [0.5, 0.6, 0.5, 0.6])))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert res["b"][0]["k"] == 10 and res["c"][0]["fusion"] == "rrf"
    assert [r[0]["chunk"]["q"] for r in res["f"]] == ["f1", "f2", "f3", "f4"]
    for queries, k, w, fusion, _ in eng.calls:
        assert len({k}) == 1
        if "f1" in queries:                               # the fan-out went out in one launch chain
            assert {"f1", "f2", "f3", "f4"} <= set(queries)
    keys = {(c[1], c[3]) for c in eng.calls}
    assert keys == {(5, None), (10, None), (5, "rrf")}
    mb.close()


def test_engine_failure_reaches_every_waiter_and_the_front_survives():
    eng = RecordingEngine(delay=0.0)
    mb = MicroBatchingEngine(eng, max_batch=8, max_wait_ms=20.0)
    errs = []

    def client(q):
        try:
            mb.search(q)
        except ValueError as e:
            errs.append(str(e))
    ts = [threading.Thread(target=client, args=(q,)) for q in ("x", "boom", "y")]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert errs and all(e == "engine failure" for e in errs)
    assert mb.search("fine")[0]["chunk"]["q"] == "fine"
    mb.close()
