"""Host logic of the product's stores on the CPU: B200VectorStore / B200MemoryVectorStore / VectorRetriever run
UNCHANGED over tests/fake_index.py (the oracle behind native.Index's interface), against the same golden outputs
of the reference's glue that the GPU tests use.  Covers id/row bookkeeping, metadata columns and their upload,
filter compilation, tombstones, upsert, result shaping — everything in store.py above the C ABI."""

import asyncio
import json
from pathlib import Path

import numpy as np
import pytest

from tests.fake_index import FakeIndex
from tests.golden_util import replay_memory_scenario, check_memory_outputs
from youtu_rag_b200 import B200MemoryVectorStore, Chunk, VectorStoreConfig, native


@pytest.fixture(autouse=True)
def _oracle_behind_the_index(monkeypatch):
    monkeypatch.setattr(native, "Index", FakeIndex)


# the GPU tests of the store, collected here without the gpu mark: same bodies, fake index underneath
from tests.test_gpu_store import (  # noqa: E402,F401
    test_persistence_roundtrip_is_bit_identical,
    test_retriever_over_b200_store_matches_reference_retriever,
    test_store_bf16_within_north_star_tolerance,
    test_store_edge_cases,
    test_store_mutations_match_reference_glue,
    test_store_reproduces_reference_glue_f32,
)


from tests.test_gpu_batched import test_per_query_filters  # noqa: E402,F401  (search_batch with one filter per query)
from tests.test_gpu_memory_store import test_memory_store_matches_reference_semantics, test_rescoring_formulas  # noqa: E402,F401


def test_memory_store_replays_the_reference_scenario():
    """every step of tests/golden/memory_store.json (the reference's MemoryVectorStore driven by
    tests/golden/make_memory_store_golden.py) gives the same outputs on B200MemoryVectorStore."""
    g = json.loads((Path(__file__).parent / "golden" / "memory_store.json").read_text())
    store = B200MemoryVectorStore(VectorStoreConfig(backend="b200", collection_name="agent_memory",
                                                    index_params={"storage_dtype": "f32"}))
    got = asyncio.run(replay_memory_scenario(store, Chunk, g["specs"], g["steps"]))
    check_memory_outputs(g, got, tol=2e-6, emb_atol=1e-6)
