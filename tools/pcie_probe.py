#!/usr/bin/env python
"""Host <-> device copy bandwidth of this box, measured the way the host-buffer encode call uses it: pinned
buffers of the bench step's sizes (0.97 GB up, 1.64 GB down per GPU), each direction alone and both at once, on
1, 2, 4, ... GPUs concurrently (one CUDA stream pair per GPU, all started together, wall clock from the first
launch to the last completion).  Prints one JSON object; bench.py's e2e number cannot beat
`both.effective_GBs_of_text` of the same GPU count.

  python tools/pcie_probe.py [--gpus N] [--up BYTES] [--down BYTES]
"""
import argparse
import json
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--up", type=int, default=973059346)
    ap.add_argument("--down", type=int, default=1643251828)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    res = {"up_bytes": a.up, "down_bytes": a.down, "gpus_visible": torch.cuda.device_count(), "by_gpus": {}}
    try:
        import pynvml
        pynvml.nvmlInit()
        res["numa_of_gpu"] = []
        for g in range(torch.cuda.device_count()):
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            try:
                res["numa_of_gpu"].append(int(pynvml.nvmlDeviceGetNumaNodeId(h)))
            except Exception:
                res["numa_of_gpu"].append(None)
        res["pcie"] = {"gen": int(pynvml.nvmlDeviceGetCurrPcieLinkGeneration(pynvml.nvmlDeviceGetHandleByIndex(0))),
                       "width": int(pynvml.nvmlDeviceGetCurrPcieLinkWidth(pynvml.nvmlDeviceGetHandleByIndex(0)))}
    except Exception as e:
        res["nvml"] = "unavailable: %s" % type(e).__name__
    bufs = []
    for g in range(a.gpus):
        with torch.cuda.device(g):
            bufs.append(dict(hu=torch.empty(a.up, dtype=torch.uint8).pin_memory(), du=torch.empty(a.up, dtype=torch.uint8, device="cuda"),
                             hd=torch.empty(a.down, dtype=torch.uint8).pin_memory(), dd=torch.empty(a.down, dtype=torch.uint8, device="cuda"),
                             su=torch.cuda.Stream(), sd=torch.cuda.Stream()))

    def run(n, up, down):
        best = None
        for _ in range(a.reps):
            for g in range(n):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for g in range(n):
                b = bufs[g]
                with torch.cuda.device(g):
                    if up:
                        with torch.cuda.stream(b["su"]):
                            b["du"].copy_(b["hu"], non_blocking=True)
                    if down:
                        with torch.cuda.stream(b["sd"]):
                            b["hd"].copy_(b["dd"], non_blocking=True)
            for g in range(n):
                torch.cuda.synchronize(g)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return best

    n = 1
    while n <= a.gpus:
        run(n, True, True)
        tu, td, tb = run(n, True, False), run(n, False, True), run(n, True, True)
        res["by_gpus"][str(n)] = {
            "h2d_alone": {"ms": tu * 1e3, "GBs_total": n * a.up / tu / 1e9},
            "d2h_alone": {"ms": td * 1e3, "GBs_total": n * a.down / td / 1e9},
            "both": {"ms": tb * 1e3, "GBs_total_both_directions": n * (a.up + a.down) / tb / 1e9,
                     "effective_GBs_of_text": n * a.up / tb / 1e9,
                     "note": "the floor of a host-buffer encode step of this size: text up and ids down at the same time"},
        }
        n *= 2
    print(json.dumps(res))


if __name__ == "__main__":
    main()
