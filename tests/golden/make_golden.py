"""Generates the committed golden fixtures from the oracle on the reference's real
corpus.  Run once in the build container:  python tests/golden/make_golden.py
(reads tests/golden/legal_chunks.json.gz, itself a gzip of /root/reference/legal_chunks.json).

The reference ships no golden vectors for this path and its libraries
cannot be imported here, so these are outputs of the CPU restatement (oracle/), used to
pin the oracle against accidental change and to check the CUDA path on the GPU box."""
import gzip
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import REFERENCE_QUERIES  # noqa: E402
from oracle import bm25 as obm25  # noqa: E402

GOLDEN = Path(__file__).resolve().parent


def main():
    with gzip.open(GOLDEN / "legal_chunks.json.gz", "rt", encoding="utf-8") as f:
        chunks = json.load(f)
    texts = [c["text"] for c in chunks]
    csr = obm25.BM25OkapiCSR.from_corpus([obm25.tokenize(t) for t in texts])
    cases = []
    for q in REFERENCE_QUERIES:
        s = csr.get_scores(obm25.tokenize(q))
        top = np.lexsort((np.arange(len(s)), -s))[:10]
        cases.append({"query": q, "max_hex": float(s.max()).hex(),
                      "top10_ids": [int(i) for i in top],
                      "top10_scores_hex": [float(s[i]).hex() for i in top]})
    out = {"n_docs": csr.corpus_size, "vocab": len(csr.vocab), "nnz": int(len(csr.post_doc)),
           "avgdl_hex": float(csr.avgdl).hex(), "average_idf_hex": float(csr.average_idf).hex(),
           "cases": cases}
    (GOLDEN / "bm25_real_corpus.json").write_text(json.dumps(out, indent=1))
    print("wrote", GOLDEN / "bm25_real_corpus.json")


if __name__ == "__main__":
    main()
