import json
from pathlib import Path

import numpy as np

from youtu_rag_b200.base import Chunk

GOLDEN = json.loads((Path(__file__).parent / "golden" / "reference_glue.json").read_text())


def golden_chunks():
    x = np.asarray(GOLDEN["corpus"]["embeddings"], np.float32)
    metas = GOLDEN["corpus"]["metadatas"]
    return [Chunk(id=f"doc{i // 8}_chunk_{i % 8}", document_id=f"doc{i // 8}", content=f"text {i}", chunk_index=i % 8,
                  metadata={**metas[i], "none_field": None}, embedding=x[i].tolist()) for i in range(len(metas))]


def compare_with_golden(results, want, tol, tie_eps=5e-7):
    """results: list[(Chunk, score)]; want: golden result dicts.  Ids must agree except inside groups of
    scores tied within tie_eps (the reference's engine breaks such ties by insertion order)."""
    assert len(results) == len(want)
    got_ids = [c.id for c, _ in results]
    got_scores = np.array([s for _, s in results], np.float64)
    want_scores = np.array([w["score"] for w in want], np.float64)
    np.testing.assert_allclose(got_scores, want_scores, rtol=tol, atol=tol)
    for i, (g, w) in enumerate(zip(got_ids, want)):
        if g != w["id"]:
            # accept only if the golden item sits in a tie group that also contains ours
            group = [x["id"] for x in want if abs(x["score"] - w["score"]) <= max(tie_eps, tol)]
            assert g in group or i == len(want) - 1, f"pos {i}: {g} vs {w['id']}"
    for (c, _), w in zip(results, want):
        if "metadata" in w and c.id == w["id"]:
            assert c.document_id == w["document_id"] and c.chunk_index == w["chunk_index"]
            assert c.content == w["content"] and c.metadata == w["metadata"]
