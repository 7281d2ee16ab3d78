"""CPU baseline on the UPSTREAM ENGINE (test / measurement infrastructure, never on the product path).

The reference's arithmetic lives in tiktoken-rs 0.7.0 (`Cargo.toml:40`), which vendors openai/tiktoken's Rust
`CoreBPE`; Python `tiktoken` is the same Rust core.  BASELINE.md section 3: the engine configured like
`Tekkenizer::new` configures it (src/tekkenizer.rs:122-126), in N worker PROCESSES (its thread batch API is GIL-bound
on short documents), each encoding a contiguous range of documents with one `encode_ordinary` call per document --
the reference makes one `encode` call per string -- on a bounded sample, reported as a rate.

Only bench.py's cpu_baseline leg imports this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_ENC = None
_DOCS = None


def _init(path: str):
    global _ENC
    from oracle import tekken_oracle as TO
    orc = TO.OracleTekkenizer.from_file(path)
    _ENC = TO.tiktoken_engine(orc.ranks)
    _ENC.encode_ordinary("warm up the regex and the thread-local state")


def _work(texts):
    n = 0
    for t in texts:
        n += len(_ENC.encode_ordinary(t))
    return n


def measure(path: str, data: np.ndarray, off: np.ndarray, procs: int, budget_s: float = 12.0, long_pieces: bool = False) -> dict:
    """Rate of the engine over documents data[off[d]:off[d+1]] with `procs` processes.  The sample is cut so that
    the run takes about `budget_s` (the engine does roughly 4 MB/s per core on short pieces; a 64 KiB single
    pre-token takes seconds: quadratic merge loop)."""
    import tiktoken  # noqa: F401  (ImportError -> the caller reports the engine as unavailable)
    procs = max(1, int(procs))
    off = np.asarray(off, dtype=np.int64)
    n_docs = len(off) - 1
    est = 0.35 if long_pieces else 4.0e6            # 64 KiB pieces: ~0.4 per second per core; else bytes per second per core
    if long_pieces:
        n_take = int(min(n_docs, max(procs, est * budget_s * procs)))
    else:
        want = est * budget_s * procs
        n_take = int(min(n_docs, max(procs, np.searchsorted(off, want))))
    raw = data[:int(off[n_take])].tobytes()
    texts = [raw[int(off[d]):int(off[d + 1])].decode("utf-8") for d in range(n_take)]
    # contiguous ranges with equal byte counts, several per process so that a slow range does not set the time
    n_tasks = min(n_take, procs * 4)
    cuts = np.searchsorted(off[:n_take + 1], np.linspace(0, off[n_take], n_tasks + 1)).tolist()
    cuts[0], cuts[-1] = 0, n_take
    tasks = [texts[cuts[i]:cuts[i + 1]] for i in range(n_tasks) if cuts[i + 1] > cuts[i]]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_init, initargs=(path,)) as pool:
        pool.map(_work, [["warm"]] * procs)
        t0 = time.perf_counter()
        n_tok = sum(pool.map(_work, tasks, chunksize=1))
        dt = time.perf_counter() - t0
    nb = int(off[n_take])
    return {"value": nb / dt / 1e9, "unit": "GB/s", "cores": procs, "kind": "engine", "tokens_per_s": n_tok / dt, "seconds": dt,
            "sample": "%d documents, %.1f MB, %d processes, one encode_ordinary call per document" % (n_take, nb / 1e6, procs),
            "note": "tiktoken %s (the Rust CoreBPE tiktoken-rs vendors) with the pattern of src/tekkenizer.rs:123; ordinary ids only (no +1000, no BOS/EOS)" % _version()}


def _version() -> str:
    try:
        import tiktoken
        return getattr(tiktoken, "__version__", "?")
    except Exception:
        return "?"


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tekken_rs_b200 import assets, corpus
    d, o = corpus.mixed_script_docs(16384, 42)
    print(measure(assets.ensure_tekken_json(), d, o, len(os.sched_getaffinity(0)), 5.0))
