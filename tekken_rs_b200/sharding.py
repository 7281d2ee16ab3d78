"""Multi-GPU sharding of a batch: documents are independent (one `encode` call per string in the
reference, src/tekkenizer.rs:378), so a batch is cut into contiguous document ranges balanced by
bytes, one per GPU, with no data-path collective.  Only the per-shard token counts are exchanged
(a handful of integers) to turn shard-local token offsets into global ones."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from .tekkenizer import shard_plan


def local_shard(doc_off: np.ndarray, rank: int, world: int) -> Tuple[int, int]:
    """Document range [begin, end) of `rank` among `world` byte-balanced shards."""
    plan = shard_plan(doc_off, world)
    return int(plan[rank]), int(plan[rank + 1])


def rebase_offsets(doc_off: np.ndarray, begin: int, end: int) -> np.ndarray:
    """Shard-local byte offsets (starting at 0) of documents [begin, end)."""
    off = np.ascontiguousarray(doc_off[begin:end + 1], dtype=np.uint64)
    return off - off[0]


def stitch_token_offsets(local_tok_off: np.ndarray, shard_token_counts: Sequence[int], rank: int) -> np.ndarray:
    """Global token offsets of this shard's documents given every shard's total token count."""
    base = int(np.sum(np.asarray(shard_token_counts[:rank], dtype=np.uint64))) if rank else 0
    return np.ascontiguousarray(local_tok_off, dtype=np.uint64) + np.uint64(base)
