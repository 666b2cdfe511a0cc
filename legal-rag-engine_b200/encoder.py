"""Host side of K1: uploads the all-MiniLM-L6-v2 weights through the C ABI and runs
``lrx_encode``.  Mirrors what the reference gets from
``SentenceTransformer("all-MiniLM-L6-v2").encode(texts)`` + ``faiss.normalize_L2``
(retrieval_engine.py:28,61-62; create_vector_store.py:33-34,45,51).  No arithmetic here."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, Sequence

import numpy as np
import torch

from . import _lib
from .tokenizer import load_tokenizer

_LAYER_KEYS = (("wq", "attention.self.query.weight"), ("bq", "attention.self.query.bias"),
               ("wk", "attention.self.key.weight"), ("bk", "attention.self.key.bias"),
               ("wv", "attention.self.value.weight"), ("bv", "attention.self.value.bias"),
               ("wo", "attention.output.dense.weight"), ("bo", "attention.output.dense.bias"),
               ("ln1_g", "attention.output.LayerNorm.weight"), ("ln1_b", "attention.output.LayerNorm.bias"),
               ("w1", "intermediate.dense.weight"), ("b1", "intermediate.dense.bias"),
               ("w2", "output.dense.weight"), ("b2", "output.dense.bias"),
               ("ln2_g", "output.LayerNorm.weight"), ("ln2_b", "output.LayerNorm.bias"))


def _strip(sd: Dict[str, object]) -> Dict[str, object]:
    """Accept ``bert.``-prefixed or bare HuggingFace names."""
    out = {}
    for k, v in sd.items():
        for pre in ("bert.", "0.auto_model.", "auto_model."):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v
    return out


def upload_weights(dev, state_dict: Dict[str, object]) -> None:
    """``dev``: DeviceIndex.  Copies the float32 tensors to the GPU, lets the library pack its
    fp16 copies (lrx_set_encoder_weights synchronises), then drops the float32 copies."""
    sd = _strip(state_dict)
    keep = []

    def up(name):
        a = sd[name]
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.detach().to(device=dev.device, dtype=torch.float32).contiguous()
        keep.append(t)
        return C.c_void_p(t.data_ptr())

    w = _lib.lrx_bert_weights()
    word = sd["embeddings.word_embeddings.weight"]
    w.vocab_size = int(word.shape[0])
    w.max_positions = int(sd["embeddings.position_embeddings.weight"].shape[0])
    w.word_emb = up("embeddings.word_embeddings.weight")
    w.pos_emb = up("embeddings.position_embeddings.weight")
    w.type_emb = up("embeddings.token_type_embeddings.weight")
    w.emb_ln_g = up("embeddings.LayerNorm.weight")
    w.emb_ln_b = up("embeddings.LayerNorm.bias")
    for l in range(6):
        for field, key in _LAYER_KEYS:
            setattr(w.layers[l], field, up(f"encoder.layer.{l}.{key}"))
    dev._ck(dev.lib.lrx_set_encoder_weights(dev.h, C.byref(w)))
    dev.vocab_size = w.vocab_size
    del keep


def load_checkpoint(model_dir) -> Dict[str, np.ndarray]:
    """``model.safetensors`` (or ``pytorch_model.bin``) of all-MiniLM-L6-v2 -> state_dict."""
    model_dir = Path(model_dir)
    for sub in ("", "0_Transformer"):
        st = model_dir / sub / "model.safetensors"
        if st.exists():
            from safetensors.numpy import load_file
            return load_file(str(st))
        pt = model_dir / sub / "pytorch_model.bin"
        if pt.exists():
            return {k: v.numpy() for k, v in torch.load(str(pt), map_location="cpu").items()}
    raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {model_dir}")


class SentenceEncoder:
    """texts -> unit float32 [n,384] embeddings on one GPU (the reference's
    ``self.model.encode`` + ``faiss.normalize_L2``)."""

    MAX_SEQ = 256          # sentence-transformers max_seq_length of all-MiniLM-L6-v2

    def __init__(self, dev, state_dict=None, model_dir=None, tokenizer=None):
        self.dev = dev
        if state_dict is None:
            state_dict = load_checkpoint(model_dir)
        upload_weights(dev, state_dict)
        # a tokenizer is only needed for text input; it is never guessed: either the caller's, or
        # the checkpoint's vocab.txt (FileNotFoundError when model_dir has none)
        self.tokenizer = tokenizer if tokenizer is not None else \
            (load_tokenizer(model_dir, dev.vocab_size) if model_dir is not None else None)

    def require_tokenizer(self):
        if self.tokenizer is None:
            raise ValueError("this encoder was built from a state_dict without a tokenizer: pass "
                             "tokenizer= (e.g. WordPieceTokenizer.from_file(vocab.txt); HashTokenizer "
                             "only for seeded synthetic weights)")
        return self.tokenizer

    def encode_ids(self, ids: np.ndarray, lens: np.ndarray) -> np.ndarray:
        """Host buffers in/out through lrx_encode_host."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        B, S = ids.shape
        out = np.empty((B, 384), dtype=np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.dev._ck(self.dev.lib.lrx_encode_host(self.dev.h, vp(ids), vp(lens), B, S, vp(out)))
        return out

    def encode_ids_device(self, ids: torch.Tensor, lens: torch.Tensor, want_f32=True, want_f16=True):
        """Device tensors in/out through lrx_encode (no synchronisation)."""
        B, S = int(ids.shape[0]), int(ids.shape[1])
        o32 = torch.empty((B, 384), dtype=torch.float32, device=self.dev.device) if want_f32 else None
        o16 = torch.empty((B, 384), dtype=torch.float16, device=self.dev.device) if want_f16 else None
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        self.dev._ck(self.dev.lib.lrx_encode(self.dev.h, p(ids), p(lens), B, S, p(o32), p(o16)))
        return o32, o16

    def encode(self, texts: Sequence[str], batch_size: int = 1024) -> np.ndarray:
        """Tokenise (truncate to 256, pad to the longest of the batch) and embed.  Texts are
        processed longest-first so batches are dense, results returned in input order."""
        texts = list(texts)
        self.require_tokenizer()
        out = np.empty((len(texts), 384), dtype=np.float32)
        enc = (self.tokenizer.encode_batch(texts, self.MAX_SEQ) if hasattr(self.tokenizer, "encode_batch")
               else [self.tokenizer.encode(t, self.MAX_SEQ) for t in texts])
        order = sorted(range(len(texts)), key=lambda i: -len(enc[i]))
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            S = max(len(enc[i]) for i in idx)
            ids = np.zeros((len(idx), S), dtype=np.int32)
            lens = np.empty(len(idx), dtype=np.int32)
            for r, i in enumerate(idx):
                ids[r, :len(enc[i])] = enc[i]
                lens[r] = len(enc[i])
            out[idx] = self.encode_ids(ids, lens)
        return out
