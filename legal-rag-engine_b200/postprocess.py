"""Orchestrator post-processing of the fused results: the priority boosts and
the parent expansion of the reference's ``LegalOrchestrator`` (``src/retrieval/orchestrator.py:
75-139``), as one vectorised pass over per-chunk tables built once per store, so that a batch of
user queries is finished without per-result dictionary walks.

    pp = ResultPostProcessor(engine.chunks)
    merged = merge_fanout(engine.search_batch(queries, k, weights))      # engine.py, A9
    ranked = pp.prioritize(merged, intent)                                # :75-114
    final = pp.expand(ranked[:k])                                         # :116-139

Same semantics as the reference, result for result: the boost is accumulated with the same
float64 additions in the same order (a rule that does not apply adds 0.0, which changes no bit),
``score`` is multiplied in place, the sort is Python's stable descending sort, ``parent_context``
is attached to sub-units whose section chunk exists.  ``intent`` is anything with the fields of
the reference's ``QueryIntent`` (attributes or dict keys: ``category``, ``user_context``,
``key_entities``, ``sub_intent``).  One reference quirk is kept on purpose: with
``user_context == "victim_distress"``, a category outside police_duty/procedure and
``sub_intent`` None, the reference's ``w in None`` raises ``TypeError`` -- so does this.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import numpy as np

_SUB_UNITS = ("illustration", "explanation", "sub_section")


def _field(intent: Any, name: str, default=None):
    if isinstance(intent, dict):
        return intent.get(name, default)
    return getattr(intent, name, default)


class ResultPostProcessor:
    def __init__(self, chunks: Sequence[Dict]):
        self.chunks = chunks
        n = len(chunks)
        # identity of the chunk dict -> row (engine results carry the original dict objects)
        self._row_of = {id(c): i for i, c in enumerate(chunks)}
        self.law_upper: List[str] = []
        sub_unit = np.zeros(n, dtype=bool)
        section_of: Dict[tuple, int] = {}
        keys = []
        for i, c in enumerate(chunks):
            meta = c.get("metadata", {}) if isinstance(c, dict) else {}
            self.law_upper.append(str(meta.get("law", "")).upper())
            law, section, unit = meta.get("law"), meta.get("section"), meta.get("unit_type")
            keys.append((law, section))
            sub_unit[i] = unit in _SUB_UNITS
            if law and section and unit == "section":
                section_of[(law, section)] = i          # later chunks overwrite (dict semantics)
        law_arr = np.array(self.law_upper, dtype=object)
        has = lambda s: np.fromiter((s in x for x in law_arr), dtype=bool, count=n)
        self.f_bnss, self.f_sop, self.f_nalsa = has("BNSS"), has("SOP"), has("NALSA")
        self.f_bns_only = has("BNS") & ~self.f_bnss
        self.sub_unit = sub_unit
        self.parent = np.full(n, -1, dtype=np.int64)
        for i, key in enumerate(keys):
            if sub_unit[i]:
                self.parent[i] = section_of.get(key, -1)
        self._law_match_cache: Dict[str, np.ndarray] = {}

    def _rows(self, results: Sequence[Dict]) -> np.ndarray:
        return np.fromiter((self._row_of[id(r["chunk"])] for r in results), dtype=np.int64,
                           count=len(results))

    def _law_contains(self, needle: str) -> np.ndarray:
        m = self._law_match_cache.get(needle)
        if m is None:
            m = np.fromiter((needle in x for x in self.law_upper), dtype=bool, count=len(self.law_upper))
            self._law_match_cache[needle] = m
        return m

    # ------------------------------------------------------------ orchestrator.py:75-114
    def boosts(self, rows: np.ndarray, intent: Any) -> np.ndarray:
        category = _field(intent, "category")
        boost = np.ones(len(rows), dtype=np.float64)
        if _field(intent, "user_context") == "victim_distress":
            police = category in ["police_duty", "procedure"]
            if not police:
                sub = _field(intent, "sub_intent", "")
                police = any((w in sub) or "" for w in ["FIR", "report", "police"])   # TypeError on None
            boost = boost + np.where(self.f_bnss[rows] | self.f_sop[rows], 0.5 if police else 0.3, 0.0)
            boost = boost + np.where(self.f_nalsa[rows], 0.2 if police else 0.4, 0.0)
            boost = boost - np.where(self.f_bns_only[rows], 0.2, 0.0)
        for entity in _field(intent, "key_entities", []) or []:
            boost = boost + np.where(self._law_contains(entity.upper())[rows], 0.2, 0.0)
        if category in ["definition", "punishment"]:
            boost = boost - np.where(self.f_sop[rows], 0.3, 0.0)
        return boost

    def prioritize(self, results: List[Dict], intent: Any) -> List[Dict]:
        if results:
            rows = self._rows(results)
            boost = self.boosts(rows, intent)
            for r, b in zip(results, boost.tolist()):
                r["score"] *= b
        results.sort(key=lambda x: x["score"], reverse=True)
        return results

    # ----------------------------------------------------------- orchestrator.py:116-139
    def expand(self, results: Sequence[Dict]) -> List[Dict]:
        final, seen = [], set()
        rows = self._rows(results) if results else np.zeros(0, dtype=np.int64)
        parents = self.parent[rows]
        for res, p in zip(results, parents.tolist()):
            header = res["chunk"].get("canonical_header")
            if header in seen:
                continue
            seen.add(header)
            if p >= 0:
                parent = self.chunks[p]
                if parent.get("canonical_header") != header:
                    res["parent_context"] = parent["text"]
            final.append(res)
        return final

    def finish(self, result_lists: Sequence[List[Dict]], intent: Any, k: int) -> List[Dict]:
        """orchestrator.py:54-70 after the searches: merge the fan-out, prioritise, expand."""
        from .engine import merge_fanout
        return self.expand(self.prioritize(merge_fanout(result_lists), intent)[:k])
