"""TEST INFRASTRUCTURE ONLY (imported by tests/): literal restatement of the reference's result
post-processing, ``src/retrieval/orchestrator.py`` -- section lookup (:18-26), priority boosts
(:75-114), parent expansion (:116-139) -- written as the reference writes it, result by result,
to check legal-rag-engine_b200/postprocess.py against.  Parity unpinned: the reference holds no
tests or golden outputs for these functions."""
from __future__ import annotations


def section_lookup(all_chunks):
    lookup = {}
    for chunk in all_chunks:                                   # orchestrator.py:19-26
        meta = chunk.get("metadata", {})
        law, section, unit_type = meta.get("law"), meta.get("section"), meta.get("unit_type")
        if law and section and unit_type == "section":
            lookup[(law, section)] = chunk
    return lookup


def prioritize_results(results, intent):
    """intent: dict with category / user_context / key_entities / sub_intent (model_dump())."""
    for res in results:                                        # orchestrator.py:76-111
        meta = res["chunk"].get("metadata", {})
        law = str(meta.get("law", "")).upper()
        boost = 1.0
        if intent["user_context"] == "victim_distress":
            is_police_task = intent["category"] in ["police_duty", "procedure"] or any(
                w in intent.get("sub_intent", "") or "" for w in ["FIR", "report", "police"])
            if "BNSS" in law or "SOP" in law:
                boost += 0.5 if is_police_task else 0.3
            if "NALSA" in law:
                boost += 0.2 if is_police_task else 0.4
            if "BNS" in law and "BNSS" not in law:
                boost -= 0.2
        for entity in intent["key_entities"]:
            if entity.upper() in law:
                boost += 0.2
        if intent["category"] in ["definition", "punishment"] and "SOP" in law:
            boost -= 0.3
        res["score"] *= boost
    results.sort(key=lambda x: x["score"], reverse=True)      # orchestrator.py:113
    return results


def expand_results(results, lookup):
    final_results, seen_headers = [], set()
    for res in results:                                        # orchestrator.py:120-137
        chunk = res["chunk"]
        meta = chunk.get("metadata", {})
        header = chunk.get("canonical_header")
        if header in seen_headers:
            continue
        seen_headers.add(header)
        if meta.get("unit_type") in ["illustration", "explanation", "sub_section"]:
            parent = lookup.get((meta.get("law"), meta.get("section")))
            if parent and parent.get("canonical_header") != header:
                res["parent_context"] = parent["text"]
        final_results.append(res)
    return final_results
