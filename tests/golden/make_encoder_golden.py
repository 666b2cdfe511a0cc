#!/usr/bin/env python
"""Generates tests/golden/encoder_golden.npz: outputs of HuggingFace ``transformers.BertModel``
(the module sentence-transformers wraps for all-MiniLM-L6-v2) for seeded random weights, used to
PIN oracle/encoder.py.  Run in the build container (transformers 5.5, torch CPU):

    python tests/golden/make_encoder_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth  # noqa: E402

CASES = [  # (weight seed, std, ln_jitter, batch, seq, token seed)
    (42, 0.02, 0.0, 3, 24, 7),
    (43, 0.06, 0.2, 4, 40, 8),     # larger weights: softmax / GELU far from linear
]


def main():
    from transformers import BertConfig, BertModel
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=512, type_vocab_size=2,
                     layer_norm_eps=1e-12, hidden_act="gelu", hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    out = {}
    for ci, (wseed, std, jit, b, s, tseed) in enumerate(CASES):
        sd = synth.bert_state_dict(wseed, std, ln_jitter=jit)
        model = BertModel(cfg, add_pooling_layer=False).eval()
        missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not [k for k in missing.missing_keys if "position_ids" not in k], missing
        ids, lens = synth.token_batch(b, s, tseed)
        mask = (np.arange(s)[None] < lens[:, None]).astype(np.int64)
        with torch.no_grad():
            hid = model(input_ids=torch.from_numpy(ids.astype(np.int64)),
                        attention_mask=torch.from_numpy(mask)).last_hidden_state
        out[f"hidden_{ci}"] = hid.numpy().astype(np.float32)
        out[f"case_{ci}"] = np.array([wseed, std, jit, b, s, tseed], dtype=np.float64)
    np.savez_compressed(ROOT / "tests" / "golden" / "encoder_golden.npz", **out)
    print("wrote encoder_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
