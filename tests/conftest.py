import gzip
import json
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def legal_chunks():
    """The reference's own corpus (legal_chunks.json, 2 620 chunks), shipped as a
    compressed fixture because /root/reference does not exist on the GPU box."""
    with gzip.open(GOLDEN / "legal_chunks.json.gz", "rt", encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def legal_texts(legal_chunks):
    return [c["text"] for c in legal_chunks]


# query strings the reference itself uses (test_retrieval.py:79-84, orchestrator.py:46-48,
# test_api.py:36, test_quality.py:17-28)
REFERENCE_QUERIES = [
    "What is the procedure for Zero FIR?",
    "Compensation for victims of acid attack",
    "Definition of a public servant under BNS",
    "Procedure after arrest of a suspect in rape case",
    "How to file FIR for robbery BNSS procedure",
    "Victim compensation rights for robbery NALSA scheme",
    "Zero FIR registration procedure BNSS",
    "I was robbed at knife point, what should I do?",
    "What is the punishment for murder?",
    "the of and section 2023 bharatiya",
    "zzzz-not-in-vocabulary qqqq",
    "",
]


@pytest.fixture(scope="session")
def reference_queries():
    return list(REFERENCE_QUERIES)
