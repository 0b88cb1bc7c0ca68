"""On-disk form of one collection: what replaces Chroma's `chroma.sqlite3` + segment
directories (utu/rag/storage/implementations/chroma_store.py:41-44,274-329).

    <persist_directory>/<collection_name>.b200/
        manifest.json            {"version", "dim", "metric", "dtype", "ld", "segments": n}
        seg_000000.rows.npy      stored rows of one add_chunks call, bit-exact (uint16 bf16 bits or float32) [n, ld]
        seg_000000.sqnorm.npy    float32 [n]
        seg_000000.meta.jsonl    one line per row: {"id", "document", "metadata"}
        deleted.json             row numbers tombstoned since (rows keep their numbers across reloads)

Append-only like the write path it mirrors: every add_chunks writes one segment, deletes rewrite the
small tombstone list, clear() removes the directory.  Loading replays the segments through
`yrb_index_append_raw`, so a reloaded collection searches bit-identically (no re-normalisation).
"""

from __future__ import annotations

import json
import re
import shutil
from pathlib import Path

import numpy as np

VERSION = 1


_NAME_RE = re.compile(r"^[a-zA-Z0-9][a-zA-Z0-9._-]{1,510}[a-zA-Z0-9]$")


def validate_collection_name(name: str) -> str:
    """Chroma's collection-name rule (the store this one replaces rejects the same names): 3-512 characters
    from [a-zA-Z0-9._-], starting and ending with an alphanumeric, no "..".  Memory collections are named
    `memory_<user_id>` (memory_store.py:209-223) with a caller-supplied user id, so a separator or ".." here
    would let mkdir / rmtree leave `persist_directory`."""
    if not isinstance(name, str) or not _NAME_RE.match(name) or ".." in name:
        raise ValueError(
            f"Expected collection name that (1) contains 3-512 characters from [a-zA-Z0-9._-], (2) starts and "
            f"ends with a character in [a-zA-Z0-9] and (3) contains no two consecutive periods, got {name!r}")
    return name


class CollectionDir:
    def __init__(self, persist_directory: str, collection_name: str):
        validate_collection_name(collection_name)
        root = Path(persist_directory).resolve()
        self.path = root / f"{collection_name}.b200"
        if self.path.resolve().parent != root:   # belt and braces: a symlinked component cannot escape either
            raise ValueError(f"collection {collection_name!r} resolves outside {persist_directory!r}")

    def exists(self) -> bool:
        return (self.path / "manifest.json").exists()

    def manifest(self) -> dict:
        return json.loads((self.path / "manifest.json").read_text())

    def _write_manifest(self, m: dict) -> None:
        tmp = self.path / "manifest.json.tmp"
        tmp.write_text(json.dumps(m))
        tmp.replace(self.path / "manifest.json")

    def create(self, dim: int, metric: str, dtype: str, ld: int) -> None:
        self.path.mkdir(parents=True, exist_ok=True)
        self._write_manifest({"version": VERSION, "dim": dim, "metric": metric, "dtype": dtype, "ld": ld, "segments": 0})

    def append_segment(self, rows: np.ndarray, sqnorm: np.ndarray, ids, documents, metadatas) -> None:
        m = self.manifest()
        stem = self.path / f"seg_{m['segments']:06d}"
        # serialise first: a value json cannot encode fails before anything touches the disk
        lines = [json.dumps({"id": i, "document": d, "metadata": md}, ensure_ascii=False)
                 for i, d, md in zip(ids, documents, metadatas)]
        np.save(f"{stem}.rows.npy", rows)
        np.save(f"{stem}.sqnorm.npy", sqnorm)
        with open(f"{stem}.meta.jsonl", "w", encoding="utf-8") as f:
            f.write("\n".join(lines) + "\n")
        m["segments"] += 1
        self._write_manifest(m)  # the manifest is written last: a torn segment is simply not listed

    def truncate_segments(self, n: int) -> None:
        """Forget segments >= n (rollback of an add whose device append failed after the segment was written)."""
        m = self.manifest()
        if m["segments"] > n:
            m["segments"] = n
            self._write_manifest(m)

    def segments(self):
        for s in range(self.manifest()["segments"]):
            stem = self.path / f"seg_{s:06d}"
            with open(f"{stem}.meta.jsonl", encoding="utf-8") as f:
                recs = [json.loads(line) for line in f]
            yield np.load(f"{stem}.rows.npy"), np.load(f"{stem}.sqnorm.npy"), recs

    def write_deleted(self, ids: list[int]) -> None:
        tmp = self.path / "deleted.json.tmp"
        tmp.write_text(json.dumps(sorted(ids), ensure_ascii=False))
        tmp.replace(self.path / "deleted.json")

    def deleted(self) -> list[int]:
        p = self.path / "deleted.json"
        return json.loads(p.read_text()) if p.exists() else []

    def remove(self) -> None:
        if self.path.exists():
            shutil.rmtree(self.path)
