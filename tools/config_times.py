#!/usr/bin/env python
"""Device-resident encode / decode times of the BASELINE configs other than the bench line
(configs 1, 3, 4), with the per-stage breakdown.  Developer tool; results quoted in DESIGN.md."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tekken_rs_b200 import Tekkenizer, assets, corpus  # noqa: E402


def run(tk, name, data, off, reps=3):
    n, nd = len(data), len(off) - 1
    d_data = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    d_data[:n] = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    cap = n + 2 * nd + 2
    d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
    d_toff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
    d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    d_boff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    tk.set_stage_timing(True)
    best, stages, ntok = None, None, 0
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ntok = tk.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), nd, n, True, True, d_tok.data_ptr(), cap, d_toff.data_ptr(), st)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, stages = dt, tk.last_stage_times()
    tk.set_stage_timing(False)
    dbest = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nb = tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), nd, ntok, 0, d_out.data_ptr(), n + 64, d_boff.data_ptr(), 0, st)
        dt = time.perf_counter() - t0
        dbest = dt if dbest is None or dt < dbest else dbest
    ok = nb == n and bool(torch.equal(d_out[:n], d_data[:n]))
    print("%-34s %8.1f MB %9d docs %10d ids | encode %8.2f ms = %7.2f GB/s | decode %7.2f ms = %7.2f GB/s | roundtrip %s" % (
        name, n / 1e6, nd, ntok, best * 1e3, n / best / 1e9, dbest * 1e3, n / dbest / 1e9, ok))
    print("    stages ms: " + ", ".join("%s %.2f" % (k, v) for k, v in stages.items()), flush=True)


def main():
    tk = Tekkenizer.from_file(assets.ensure_tekken_json(), device=0)
    one = lambda raw: (np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64))
    run(tk, "config1 english 1 MiB, 1 doc", *one(corpus.english_like(1 << 20)))
    run(tk, "config3 single doc 256 MiB", *one(corpus.single_long_document(1 << 28)))
    run(tk, "config3 single doc 1 GiB", *one(corpus.single_long_document(1 << 30)))
    run(tk, "config4 64 x 64 KiB pieces", *one(corpus.adversarial_pieces(64, 1 << 16)))
    run(tk, "config4 256 x 64 KiB pieces", *one(corpus.adversarial_pieces(256, 1 << 16)), reps=1)


if __name__ == "__main__":
    main()
