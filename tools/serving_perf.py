#!/usr/bin/env python
"""Queries/s through RetrievalEngine.search on the reference's corpus (2 620 chunks, seeded encoder
weights): one client calling search() in a loop against N client threads through the micro-batching
front (serving.MicroBatchingEngine).  python tools/serving_perf.py -> one JSON line."""
import gzip, json, sys, tempfile, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.engine import RetrievalEngine, create_vector_store
from legal_rag_engine_b200.serving import MicroBatchingEngine
from legal_rag_engine_b200.tokenizer import HashTokenizer
chunks = json.load(gzip.open(ROOT / "tests" / "golden" / "legal_chunks.json.gz", "rt", encoding="utf-8"))
d = Path(tempfile.mkdtemp())
(d / "legal_chunks.json").write_text(json.dumps(chunks), encoding="utf-8")
sd, tok = synth.bert_state_dict(42, 0.05, ln_jitter=0.1), HashTokenizer(30522)
create_vector_store(str(d / "legal_chunks.json"), str(d / "vs"), encoder_state_dict=sd, tokenizer=tok)
eng = RetrievalEngine(str(d / "vs"), encoder_state_dict=sd, tokenizer=tok)
qs = ["What is the procedure for Zero FIR?", "Compensation for victims of acid attack",
      "Definition of a public servant under BNS", "Procedure after arrest of a suspect in rape case",
      "How to file FIR for robbery BNSS procedure", "Victim compensation rights for robbery NALSA scheme",
      "Zero FIR registration procedure BNSS", "What is the punishment for murder?"]
for q in qs: eng.search(q, k=5)
n = 400
t0 = time.perf_counter()
for i in range(n): eng.search(qs[i % len(qs)], k=5)
seq = n / (time.perf_counter() - t0)
out = {"chunks": len(chunks), "sequential_search_per_s": seq}
for threads in (8, 32):
    mb = MicroBatchingEngine(eng, max_batch=64, max_wait_ms=0.2)
    close, eng.close = eng.close, (lambda: None)
    def client(j):
        for i in range(n // 8): mb.search(qs[(i + j) % len(qs)], k=5)
    ts = [threading.Thread(target=client, args=(j,)) for j in range(threads)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0
    out[f"micro_batched_{threads}_clients_per_s"] = threads * (n // 8) / dt
    out[f"mean_batch_{threads}"] = mb.requests / max(mb.batches, 1)
    mb.close(); eng.close = close
print(json.dumps(out))
eng.close()
