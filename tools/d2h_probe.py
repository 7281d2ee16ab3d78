#!/usr/bin/env python
"""Experiment: device -> host copy rate into pinned memory by allocation flag, copy size and destination alignment
(the host-buffer encode call is bound by this copy).  ctypes on libcudart; prints one JSON object."""
import ctypes
import json
import time

import torch

rt = ctypes.CDLL("libcudart.so.12")
torch.cuda.init()
torch.zeros(1, device="cuda")
N = 1643251828


def check(rc):
    assert rc == 0, rc


def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    check(rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags)))
    ctypes.memset(p, 1, nbytes)          # touch
    return p


def timed(fn, reps=4):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best


d = torch.empty(N + 4096, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
res = {}
D2H = 2
for name, flags in (("default", 0), ("portable", 1), ("mapped", 2), ("write_combined", 4)):
    h = host_alloc(N + 4096, flags)

    def one():
        check(rt.cudaMemcpyAsync(h, ctypes.c_void_p(d.data_ptr()), ctypes.c_size_t(N), D2H, ctypes.c_void_p(st.cuda_stream)))

    def pieces(k, skew):
        step = N // k // 4 * 4
        for i in range(k):
            check(rt.cudaMemcpyAsync(ctypes.c_void_p(h.value + i * step + skew), ctypes.c_void_p(d.data_ptr() + i * step), ctypes.c_size_t(step), D2H,
                                     ctypes.c_void_p(st.cuda_stream)))
    res[name] = {"one_copy_GBs": N / timed(one) / 1e9, "8_copies_aligned_GBs": N / timed(lambda: pieces(8, 0)) / 1e9,
                 "8_copies_dst_plus_4_bytes_GBs": N / timed(lambda: pieces(8, 4)) / 1e9,
                 "32_copies_dst_plus_1028_bytes_GBs": N / timed(lambda: pieces(32, 1028)) / 1e9}
    check(rt.cudaFreeHost(h))
t = torch.empty(N, dtype=torch.uint8).pin_memory()
res["torch_pin_memory"] = {"one_copy_GBs": N / timed(lambda: t.copy_(d[:N], non_blocking=True)) / 1e9}
print(json.dumps(res))
