#!/usr/bin/env python
"""Experiment: does running the encode pipeline of SEVERAL sub-batches concurrently on one GPU (one handle, stream and
host thread per sub-batch) beat one pipeline over the whole batch?  The stages have different bottlenecks (split /
lookup / emit: instruction issue; short merge classes: the L1TEX pipe; long merge classes: latency at low occupancy),
so kernels of different stages might fill each other's idle resources.  Prints one JSON object."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tekken_rs_b200 import Tekkenizer, assets, corpus, shard_plan  # noqa: E402


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    data, off = corpus.mixed_script_docs(n_docs, 42)
    path = assets.ensure_tekken_json()
    res = {"docs": n_docs, "bytes": int(len(data)), "ways": {}}
    for ways in (1, 2, 3, 4):
        plan = shard_plan(off, ways)
        parts = []
        for s in range(ways):
            b, e = int(plan[s]), int(plan[s + 1])
            d = data[int(off[b]):int(off[e])]
            o = (off[b:e + 1] - off[b]).astype(np.int64)
            st = torch.cuda.Stream()
            dd = torch.from_numpy(np.concatenate([d, np.zeros(64, np.uint8)])).cuda()
            do = torch.from_numpy(o).cuda()
            cap = len(d) + 2 * (e - b) + 2
            parts.append(dict(tk=Tekkenizer.from_file(path, device=0), st=st, dd=dd, do=do, n=len(d), nd=e - b, cap=cap,
                              tok=torch.empty(cap, dtype=torch.int32, device="cuda"), toff=torch.empty(e - b + 1, dtype=torch.int64, device="cuda")))
        torch.cuda.synchronize()

        def work(p, reps):
            for _ in range(reps):
                p["tk"].encode_batch_device(p["dd"].data_ptr(), p["do"].data_ptr(), p["nd"], p["n"], True, True, p["tok"].data_ptr(), p["cap"],
                                            p["toff"].data_ptr(), p["st"].cuda_stream)

        def run(reps):
            th = [threading.Thread(target=work, args=(p, reps)) for p in parts]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps
        run(3)
        best = min(run(5) for _ in range(3))
        res["ways"][str(ways)] = {"ms_per_batch": best * 1e3, "GBs": len(data) / best / 1e9}
        for p in parts:
            p["tk"].close()
        del parts
    print(json.dumps(res))


if __name__ == "__main__":
    main()
