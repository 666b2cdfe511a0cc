"""The drop-in boundary on the GPU: create_vector_store + RetrievalEngine.search /
search_batch on the reference's own corpus (legal_chunks.json) and query strings, checked
bit for bit against the CPU oracle downstream of the embeddings (the embeddings themselves
are checked against the fp32 oracle here on a sample and in test_gpu_encoder.py)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def built(tmp_path_factory, legal_chunks):
    from legal_rag_engine_b200 import synth
    from legal_rag_engine_b200.engine import RetrievalEngine, create_vector_store
    d = tmp_path_factory.mktemp("store")
    src = d / "legal_chunks.json"
    src.write_text(json.dumps(legal_chunks), encoding="utf-8")
    from legal_rag_engine_b200.tokenizer import HashTokenizer
    sd = synth.bert_state_dict(42, 0.05, ln_jitter=0.1)
    tok = HashTokenizer(30522)          # seeded synthetic weights: the stand-in tokenizer, passed explicitly
    create_vector_store(str(src), str(d / "vs"), encoder_state_dict=sd, tokenizer=tok)
    eng = RetrievalEngine(str(d / "vs"), encoder_state_dict=sd, tokenizer=tok)
    yield eng, sd, d / "vs"
    eng.close()


def _oracle(eng):
    from oracle import bm25 as obm25
    from oracle.search import OracleIndex
    b = eng.bm25
    csr = obm25.BM25OkapiCSR(b.n_docs, b.doc_len, b.term_ptr.astype(np.int64), b.postings[:, 0],
                             b.postings[:, 1], vocab=b.vocab)
    return OracleIndex(eng._x.cpu().numpy(), csr), csr


def test_store_embeddings_match_fp32_oracle(built, legal_texts):
    eng, sd, vs = built
    from oracle import encoder as oenc
    xh = np.load(vs / "vectors.f16.npy")
    assert xh.shape == (len(legal_texts), 384)
    tok = eng.model.tokenizer
    for i in [0, 1, 17, 500, 2619]:
        ids = np.asarray([tok.encode(legal_texts[i], 256)], dtype=np.int32)
        want = oenc.encode_ids(sd, ids, np.array([ids.shape[1]]))[0].astype(np.float64)
        got = xh[i].astype(np.float64)
        assert got @ want / np.linalg.norm(got) / np.linalg.norm(want) >= 0.9999


@pytest.mark.parametrize("fusion", ["linear", "rrf"])
def test_search_matches_oracle(built, reference_queries, fusion):
    eng, sd, _ = built
    oracle, csr = _oracle(eng)
    from oracle.bm25 import tokenize
    for q in reference_queries:
        for k, w in ((5, 0.5), (10, 0.6)):
            got = eng.search(q, k=k, hybrid_weight=w, fusion=fusion)
            qh = eng.encode([q]).astype(np.float16)[0]
            want = oracle.search_vec(qh, csr.term_ids(tokenize(q)), k, w, fusion)
            assert len(got) == len(want) <= k
            for r, (i, score, sem, kw) in zip(got, want):
                assert set(r) == {"chunk", "score", "semantic", "keyword"}
                assert r["chunk"] is eng.chunks[i]          # the metadata dict object itself
                assert r["score"] == score and r["semantic"] == sem and r["keyword"] == kw
                assert isinstance(r["score"], float) and "text" in r["chunk"]
            assert [r["score"] for r in got] == sorted((r["score"] for r in got), reverse=True)


def test_search_batch_equals_sequential_search(built):
    eng, _, _ = built
    from legal_rag_engine_b200.engine import fanout_queries, merge_fanout
    qs, ws = fanout_queries("I was robbed at knife point, what should I do?", "victim_distress",
                            ["robbery"], "procedure")
    assert len(qs) == 4
    batch = eng.search_batch(qs, k=5, hybrid_weights=ws)
    single = [eng.search(q, k=5, hybrid_weight=w) for q, w in zip(qs, ws)]
    strip = lambda rs: [(r["chunk"]["canonical_header"], r["score"], r["semantic"], r["keyword"]) for r in rs]
    assert [strip(r) for r in batch] == [strip(r) for r in single]
    merged = merge_fanout(batch)
    heads = [r["chunk"]["canonical_header"] for r in merged]
    assert len(heads) == len(set(heads)) and len(merged) <= 20


def test_text_path_begin_end_equals_blocking_call(built):
    """lrx_search_text_host_begin / lrx_search_host_end (encoder in the call) on two handles over the
    one index, two batches in flight: the same results, bit for bit, as the blocking call."""
    import torch
    from legal_rag_engine_b200.bm25_index import tokenize
    from legal_rag_engine_b200.encoder import SentenceEncoder
    eng, sd, _ = built
    tok = eng.model.tokenizer
    batches = [["What is the procedure for Zero FIR?", "punishment for murder"],
               ["Compensation for victims of acid attack", "bail for non-bailable offence", "theft"],
               ["How to file FIR for robbery BNSS procedure"]]

    def host_form(qs):
        enc = [tok.encode(q, 256) for q in qs]
        S = max(len(e) for e in enc)
        ids = np.zeros((len(qs), S), dtype=np.int32)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = e
        lens = np.array([len(e) for e in enc], dtype=np.int32)
        lists = [[t for t in eng.bm25.term_ids(tokenize(q)) if t >= 0] for q in qs]
        return ids, lens, lists
    forms = [host_form(qs) for qs in batches]
    want = [eng.dev.search_text_host(i, l, t, 5, [0.5] * len(t), "rrf") for i, l, t in forms]
    view = eng.dev.clone_view()
    venc = SentenceEncoder(view, state_dict=sd, tokenizer=tok)      # the second handle's own encoder state
    try:
        hs = [eng.dev, view]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        for d, s in zip(hs, streams):
            with torch.cuda.stream(s):
                d.use_current_stream()
        got = [None] * len(forms)
        order = list(range(len(forms))) * 3                      # slots and captured chains are reused
        pend = {}
        for n, j in enumerate(order):
            d = hs[n % 2]
            if d in pend:
                got[pend.pop(d)] = d.search_host_end()
            i, l, t = forms[j]
            d.search_text_host_begin(i, l, t, 5, [0.5] * len(t), "rrf")
            pend[d] = j
        for d, j in pend.items():
            got[j] = d.search_host_end()
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                np.testing.assert_array_equal(a, b)
    finally:
        del venc
        view.close()
        eng.dev.use_current_stream()


def test_small_k_and_bad_fusion(built):
    eng, _, _ = built
    assert len(eng.search("zero fir", k=3)) == 3
    with pytest.raises(KeyError):
        eng.search("zero fir", fusion="nope")


def test_k_beyond_depth_raises_and_long_query_scores_every_token(built):
    eng, _, _ = built
    oracle, csr = _oracle(eng)
    from oracle.bm25 import tokenize
    with pytest.raises(ValueError):
        eng.search("zero fir", k=129)
    assert len(eng.search("zero fir", k=128)) == 128
    # a pasted paragraph: 200+ whitespace tokens, every one of them scored (retrieval_engine.py:67-68)
    longest = max(eng.chunks, key=lambda c: len(c["text"].split()))
    long_q = " ".join(longest["text"].split()[:230])
    assert len(tokenize(long_q)) > 200
    got = eng.search(long_q, k=10, hybrid_weight=0.5)
    qh = eng.encode([long_q]).astype(np.float16)[0]
    want = oracle.search_vec(qh, csr.term_ids(tokenize(long_q)), 10, 0.5, "linear")
    assert [(r["score"], r["keyword"]) for r in got] == [(w[1], w[3]) for w in want]
    assert [r["chunk"] is eng.chunks[w[0]] for r, w in zip(got, want)] == [True] * len(want)


def test_state_dict_without_tokenizer_is_refused(built):
    eng, sd, vs = built
    from legal_rag_engine_b200.engine import RetrievalEngine
    with pytest.raises(ValueError):
        RetrievalEngine(str(vs), encoder_state_dict=sd)


def test_orchestrate_flow_on_gpu_equals_oracle_flow(built):
    """orchestrator.py:28-70 on the GPU engine: fan-out -> search_batch -> merge -> priority boosts ->
    parent expansion (postprocess.ResultPostProcessor) against the literal restatement over the CPU
    oracle (oracle/search.py + oracle/postprocess.py).  (The reference's own orchestrator.py runs
    over the engine's host code in tests/test_orchestrator_dropin.py, where the reference tree is.)"""
    import copy
    eng, _, _ = built
    oracle, csr = _oracle(eng)
    from legal_rag_engine_b200.engine import fanout_queries
    from legal_rag_engine_b200.postprocess import ResultPostProcessor
    from oracle import postprocess as opost
    from oracle.bm25 import tokenize
    pp = ResultPostProcessor(eng.chunks)
    lookup = opost.section_lookup(eng.chunks)
    cases = [("I was robbed at knife point, what should I do?",
              dict(category="procedure", sub_intent="report FIR", key_entities=["robbery", "BNSS"],
                   user_context="victim_distress", confidence=0.9)),
             ("What is the punishment for murder?",
              dict(category="punishment", sub_intent=None, key_entities=["BNS"], user_context="informational",
                   confidence=0.8))]
    for query, intent in cases:
        qs, ws = fanout_queries(query, intent["user_context"], intent["key_entities"], intent["category"])
        got = pp.finish(eng.search_batch(qs, k=5, hybrid_weights=ws), intent, 5)
        lists = []
        for q, w in zip(qs, ws):
            qh = eng.encode([q]).astype(np.float16)[0]
            rows = oracle.search_vec(qh, csr.term_ids(tokenize(q)), 5, w, "linear")
            lists.append([{"chunk": copy.deepcopy(eng.chunks[i]), "score": s, "semantic": sem, "keyword": kw}
                          for i, s, sem, kw in rows])
        flat, seen = [], set()
        for rs in lists:                                        # orchestrator.py:54-62
            for r in rs:
                cid = r["chunk"].get("canonical_header")
                if cid and cid not in seen:
                    flat.append(r)
                    seen.add(cid)
        want = opost.expand_results(opost.prioritize_results(flat, dict(intent))[:5], lookup)
        key = lambda rs: [(r["chunk"]["canonical_header"], r["score"], r["semantic"], r["keyword"],
                           r.get("parent_context")) for r in rs]
        assert key(got) == key(want) and 1 <= len(got) <= 5


def test_micro_batching_front_under_concurrent_clients(built):
    """serving.MicroBatchingEngine over the real engine: 24 client threads calling search() get
    exactly what sequential search() returns, and their calls were coalesced into fewer launches."""
    import threading
    from legal_rag_engine_b200.serving import MicroBatchingEngine
    eng, _, _ = built
    queries = ["What is the procedure for Zero FIR?", "Compensation for victims of acid attack",
               "Definition of a public servant under BNS", "Procedure after arrest of a suspect in rape case",
               "How to file FIR for robbery BNSS procedure", "What is the punishment for murder?"]
    strip = lambda rs: [(r["chunk"]["canonical_header"], r["score"], r["semantic"], r["keyword"]) for r in rs]
    want = {q: strip(eng.search(q, k=5, hybrid_weight=0.6)) for q in queries}
    mb = MicroBatchingEngine(eng, max_batch=32, max_wait_ms=5.0)
    real_close, eng.close = eng.close, (lambda: None)       # the fixture owns the engine
    try:
        got = {}

        def client(i):
            q = queries[i % len(queries)]
            got[i] = (q, strip(mb.search(q, k=5, hybrid_weight=0.6)))
        ts = [threading.Thread(target=client, args=(i,)) for i in range(24)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert len(got) == 24
        for q, res in got.values():
            assert res == want[q]
        assert mb.requests == 24 and mb.batches < 24 and mb.largest_batch > 1
    finally:
        mb.close()
        eng.close = real_close
