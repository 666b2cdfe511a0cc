"""Query/chunk text -> WordPiece ids for K1 (host side).

The reference tokenises inside ``SentenceTransformer.encode`` (retrieval_engine.py:61) with
the BertTokenizer of all-MiniLM-L6-v2: lower-case, strip accents, split punctuation,
greedy longest-match WordPiece against ``vocab.txt``, ``[CLS]`` ... ``[SEP]``, truncate to 256.

* ``WordPieceTokenizer``: that algorithm over a real ``vocab.txt`` (used whenever the model
  directory holds one).
* ``HashTokenizer``: seeded stand-in for boxes without the checkpoint (this build container
  and the benchmark box have no network): lower-case whitespace tokens hashed into the id
  range [1000, vocab).  Same framing and truncation, so every downstream shape is identical.
"""
from __future__ import annotations

import hashlib
import unicodedata
from pathlib import Path
from typing import Dict, List, Optional

CLS, SEP, UNK, PAD = 101, 102, 100, 0


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def basic_tokenize(text: str) -> List[str]:
    """BertTokenizer's BasicTokenizer with do_lower_case=True: clean, lower, strip accents,
    split on whitespace and punctuation (CJK characters spaced out)."""
    out_chars = []
    for ch in text:
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD or (unicodedata.category(ch) in ("Cc", "Cf") and ch not in "\t\n\r"):
            continue
        if ch in " \t\n\r" or unicodedata.category(ch) == "Zs":
            out_chars.append(" ")
        elif (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or
              0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F):
            out_chars.append(f" {ch} ")
        else:
            out_chars.append(ch)
    tokens = []
    for tok in "".join(out_chars).split():
        tok = unicodedata.normalize("NFD", tok.lower())
        tok = "".join(c for c in tok if unicodedata.category(c) != "Mn")
        cur = []
        for ch in tok:
            if _is_punct(ch):
                if cur:
                    tokens.append("".join(cur))
                    cur = []
                tokens.append(ch)
            else:
                cur.append(ch)
        if cur:
            tokens.append("".join(cur))
    return tokens


class WordPieceTokenizer:
    def __init__(self, vocab: Dict[str, int], max_chars: int = 100):
        self.vocab = vocab
        self.max_chars = max_chars
        self.cls, self.sep = vocab.get("[CLS]", CLS), vocab.get("[SEP]", SEP)
        self.unk = vocab.get("[UNK]", UNK)

    @classmethod
    def from_file(cls, path):
        with open(path, encoding="utf-8") as f:
            return cls({line.rstrip("\n"): i for i, line in enumerate(f)})

    def wordpiece(self, token: str) -> List[int]:
        if len(token) > self.max_chars:
            return [self.unk]
        ids, start = [], 0
        while start < len(token):
            end, cur = len(token), None
            while start < end:
                sub = token[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [self.unk]
            ids.append(cur)
            start = end
        return ids

    def encode(self, text: str, max_len: int = 256) -> List[int]:
        ids = [i for tok in basic_tokenize(text) for i in self.wordpiece(tok)]
        return [self.cls] + ids[:max_len - 2] + [self.sep]

    def encode_batch(self, texts, max_len: int = 256) -> List[List[int]]:
        """Index-build path: many texts at once.  When the `tokenizers` package
        is importable its multi-threaded BertWordPieceTokenizer does the work over the same
        vocabulary (tests/test_oracle_encoder.py pins this class to it on the reference's corpus);
        otherwise the pure-Python loop."""
        fast = self._fast(max_len)
        if fast is None:
            return [self.encode(t, max_len) for t in texts]
        return [e.ids for e in fast.encode_batch(list(texts))]

    def _fast(self, max_len: int):
        cache = self.__dict__.setdefault("_fast_cache", {})
        if max_len in cache:
            return cache[max_len]
        tok = None
        try:
            from tokenizers import BertWordPieceTokenizer
            # the ids must be the line numbers of a vocab.txt: only possible for a dense vocabulary
            if sorted(self.vocab.values()) == list(range(len(self.vocab))) and \
                    all(s in self.vocab for s in ("[UNK]", "[CLS]", "[SEP]", "[PAD]", "[MASK]")):
                tok = BertWordPieceTokenizer(dict(self.vocab), lowercase=True)
                tok.enable_truncation(max_len)
        except Exception:
            tok = None
        cache[max_len] = tok
        return tok


class HashTokenizer:
    """Stand-in when no vocab.txt is available: md5(token) -> id in [1000, vocab)."""

    def __init__(self, vocab_size: int = 30522):
        self.vocab_size = vocab_size

    def encode(self, text: str, max_len: int = 256) -> List[int]:
        span = max(self.vocab_size - 1000, 1)
        ids = [1000 % self.vocab_size + int.from_bytes(hashlib.md5(t.encode("utf-8")).digest()[:4],
                                                     "little") % span
               if self.vocab_size > 1000 else 1 for t in text.lower().split()]
        return [CLS % self.vocab_size] + ids[:max_len - 2] + [SEP % self.vocab_size]

    def encode_batch(self, texts, max_len: int = 256) -> List[List[int]]:
        return [self.encode(t, max_len) for t in texts]


def load_tokenizer(model_dir: Optional[str], vocab_size: int = 30522):
    """The checkpoint's WordPiece vocabulary (``vocab.txt`` beside the weights, or under
    ``0_Transformer``).  A checkpoint without it is an error: feeding real weights hashed token ids
    would give garbage embeddings without any sign of it.  ``HashTokenizer`` is only ever used when
    a caller passes it explicitly (seeded synthetic weights in tests and the bench)."""
    if model_dir is None:
        raise FileNotFoundError("no model directory: a tokenizer cannot be loaded (pass tokenizer= explicitly "
                                "when using a state_dict of your own)")
    for sub in ("", "0_Transformer"):
        p = Path(model_dir) / sub / "vocab.txt"
        if p.exists():
            return WordPieceTokenizer.from_file(p)
    raise FileNotFoundError(f"{model_dir} holds no vocab.txt: the all-MiniLM-L6-v2 checkpoint needs its "
                            f"WordPiece vocabulary (there is no fallback tokenizer)")
