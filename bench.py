#!/usr/bin/env python
"""bench.py -- encode throughput of the B200 Tekkenizer path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--docs D]

A step is one pass of the hot path (Tekkenizer::encode over a batch, src/tekkenizer.rs:378-405)
over one batch of synthetic documents.  Workload at every N: BASELINE.json configs[1] -- 1,000,000
mixed-script UTF-8 documents of <= 1 KiB per GPU (rank r encodes documents [r*D, (r+1)*D) of the
seed-42 corpus: documents are independent, so ranks share nothing and no collective is on the data
path: weak scaling).

value  = input GB/s with documents and ids resident in HBM (tk_encode_batch_device), CUDA events.
e2e    = the same metric through the host-buffer C ABI call (tk_encode_batch): pinned host text
         in, pinned host ids out, H2D and D2H copies inside the timed region.
roofline / cpu_baseline: see DESIGN.md (sections "Measurement" and "CPU baseline").

--impl reference times the CPU arm: the oracle's C restatement of the reference engine on all
host cores (the reference is Rust and cannot be built in this image), same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode_input_throughput"
UNIT = "GB/s"
SEED = 42


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.02)

    def start(self):
        if self.nv is None:
            return
        self._stop.clear()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._once()
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def algorithmic_bytes(n_bytes: int, n_tokens: int, n_docs: int) -> int:
    """read every text byte once, write every id once, read the document offsets,
    write the token offsets.  Vocabulary tables are L2-resident and not counted."""
    return n_bytes + 4 * n_tokens + 16 * (n_docs + 1)


# ---------------------------------------------------------------------------------------------- CPU arm

def cpu_rate(orc, data, off, threads: int, reps: int = 1):
    """Best-of-reps all-thread encode of the sample; returns (GB/s, tokens/s, seconds)."""
    best = None
    n_tok = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        n_tok = orc.encode_count_mt(data, off, True, True, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return len(data) / best / 1e9, n_tok / best, best


WORKLOAD = "mixed"


def make_docs(n_docs: int, first_doc: int = 0):
    """The step's documents: (bytes uint8, doc_off uint64[n_docs+1])."""
    from tekken_rs_b200 import corpus
    if WORKLOAD == "mixed":
        return corpus.mixed_script_docs(n_docs, SEED, first_doc=first_doc)
    # English-like ASCII (config 1's generator), 1 KiB per document; the 64 MiB base text repeats
    base = np.frombuffer(corpus.english_like(1 << 26, 1234 + first_doc), dtype=np.uint8)
    reps = -(-n_docs * 1024 // len(base))
    data = np.tile(base, reps)[:n_docs * 1024]
    return data, np.arange(n_docs + 1, dtype=np.uint64) * np.uint64(1024)


def cpu_sample(n_docs_total: int, threads: int):
    n = int(min(n_docs_total, max(16384, 32768 * threads)))
    data, off = make_docs(n)
    return data, off, "first %d of %d documents of the workload (%.0f MB), all ids computed, %d threads" % (
        n, n_docs_total, len(data) / 1e6, threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import tekken_oracle as TO
    from tekken_rs_b200 import assets
    threads = host_threads()
    orc = TO.OracleTekkenizer.from_file(assets.ensure_tekken_json())
    data, off, desc = cpu_sample(args.docs, threads)
    for _ in range(args.warmup):
        orc.encode_count_mt(data, off, True, True, threads)
    t0 = time.perf_counter()
    n_tok = 0
    for _ in range(args.steps):
        n_tok = orc.encode_count_mt(data, off, True, True, threads)
    dt = time.perf_counter() - t0
    gbs = len(data) * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8->u32", "data": "synthetic",
        "tokens_per_s": n_tok * args.steps / dt,
        "config": workload_config(args, len(off) - 1, len(data)),
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "note": "C restatement of tiktoken-rs CoreBPE + Tekkenizer glue (oracle/tekken_oracle_core.c); the Rust reference cannot be built here"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_docs, n_bytes):
    name = ("BASELINE configs[1]: batch encode of %d synthetic mixed-script UTF-8 documents x <=1 KiB per GPU (seed %d), add_bos+add_eos" % (args.docs, SEED)
            if WORKLOAD == "mixed" else
            "context only (not the bench line): %d documents x 1 KiB of configs[0]'s English-like ASCII text per GPU, add_bos+add_eos" % args.docs)
    return {"workload": name,
            "docs_per_step": int(n_docs), "bytes_per_step": int(n_bytes), "parallelism": "documents sharded by rank, no collective",
            "l2": ("input per GPU per step (%.0f MB) exceeds the 126 MB L2; no flush needed" % (args_bytes_per_gpu(args, n_bytes) / 1e6))
            if args_bytes_per_gpu(args, n_bytes) > 2 * L2_BYTES else "L2 flushed between steps (256 MiB write), outside the per-step events"}


L2_BYTES = 126 << 20


def args_bytes_per_gpu(args, n_bytes):
    return n_bytes / max(1, args.gpus)


# ---------------------------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist

    from tekken_rs_b200 import Tekkenizer, assets, corpus, kernel_launch_count
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the encode path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = None
    if world > 1:
        # one process per GPU: run on (and allocate the pinned staging buffers from) the CPUs next to this GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = "cpu affinity set to GPU %d's NUMA node (%d cpus)" % (local, len(os.sched_getaffinity(0)))
        except Exception as e:   # not fatal: the numbers are then just measured without the binding
            numa = "cpu affinity not set (%s)" % type(e).__name__
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    path = assets.ensure_tekken_json() if rank == 0 else None
    if world > 1:
        dist.barrier()
        path = assets.ensure_tekken_json()
    tk = Tekkenizer.from_file(path, device=local)

    # this rank's shard of the corpus, in pinned host memory (the e2e arm reads it from there)
    data_np, off_np = make_docs(args.docs, first_doc=rank * args.docs)
    n_docs, n_bytes = len(off_np) - 1, len(data_np)
    h_data = torch.empty(n_bytes + 64, dtype=torch.uint8).pin_memory()
    h_data[:n_bytes].numpy()[:] = data_np
    h_off = torch.from_numpy(off_np.astype(np.int64)).pin_memory()
    del data_np

    stream = torch.cuda.current_stream()
    d_data = torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda")
    d_data.copy_(h_data, non_blocking=True)
    d_off = h_off.cuda()
    cap = n_bytes + 2 * n_docs + 2
    d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
    d_toff = torch.empty(n_docs + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def step_device():
        return tk.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), n_docs, n_bytes, True, True, d_tok.data_ptr(), cap,
                                      d_toff.data_ptr(), stream.cuda_stream)

    # ---- device-resident arm -----------------------------------------------------------------
    n_tokens = 0
    for _ in range(max(args.warmup, 3)):
        n_tokens = step_device()
    tk.set_stage_timing(True)
    clocks = ClockSampler(local)
    stage_ms = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    need_flush = n_bytes <= 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = kernel_launch_count()
    clocks.start()
    for a, b in pairs:
        if need_flush:
            flush_buf.fill_(1)
        a.record(stream)
        step_device()
        b.record(stream)
        for k, v in tk.last_stage_times().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
    torch.cuda.synchronize()
    clocks.stop()
    launches = kernel_launch_count() - launches0
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))
    barrier()
    tk.set_stage_timing(False)
    total_bytes = sum_over_ranks(float(n_bytes))
    total_tokens = sum_over_ranks(float(n_tokens))
    total_docs = sum_over_ranks(float(n_docs))
    ms_per_step = dev_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end arm: host buffers through tk_encode_batch --------------------------------
    h_view = h_data[:n_bytes].numpy()
    h_off_u64 = h_off.numpy().view(np.uint64)
    import ctypes

    from tekken_rs_b200 import _lib
    lib = _lib.load()

    def step_host():
        tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
        rc = lib.tk_encode_batch(tk._h, h_view.ctypes.data, h_off_u64.ctypes.data, n_docs, 1, 1, ctypes.byref(tok), ctypes.byref(toff))
        if rc != 0:
            raise RuntimeError(lib.tk_last_error().decode())
        last = ctypes.c_uint64.from_address(toff.value + 8 * n_docs).value      # read the result on the host
        first = ctypes.c_uint32.from_address(tok.value).value
        lib.tk_buffer_free(tok)
        lib.tk_buffer_free(toff)
        return last, first

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        n_e2e, first = step_host()
    assert n_e2e == n_tokens and first == 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total_bytes * e2e_steps / e2e_s / 1e9

    # ---- decode (secondary line: Tekkenizer::decode, src/tekkenizer.rs:436-560) ---------------
    d_out = torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda")
    d_boff = torch.empty(n_docs + 1, dtype=torch.int64, device="cuda")

    def step_decode():
        return tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), n_docs, n_tokens, 0, d_out.data_ptr(), n_bytes + 64,
                                      d_boff.data_ptr(), 0, stream.cuda_stream)
    for _ in range(2):
        nb = step_decode()
    roundtrip_ok = bool(nb == n_bytes and torch.equal(d_out[:n_bytes], d_data[:n_bytes]))
    dsteps = max(1, min(args.steps, 10))
    barrier()
    ev0.record(stream)
    for _ in range(dsteps):
        step_decode()
    ev1.record(stream)
    torch.cuda.synchronize()
    dec_ms = max_over_ranks(ev0.elapsed_time(ev1)) / dsteps
    barrier()

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(n_bytes, n_tokens, n_docs)
    stage_avg = {k: v / args.steps for k, v in stage_ms.items()}
    kern = {k: v for k, v in stage_avg.items() if k not in ("setup",)}
    top = max(kern, key=kern.get) if kern else None
    roof = None
    if top:
        ach = alg / (kern[top] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "kernel_ms": kern[top], "algorithmic_bytes_per_launch": alg,
                "whole_path_achieved": alg / (ms_per_step / 1e0 * 1e-3) / 1e9 if world == 1 else None,
                "whole_path_frac": alg / (ms_per_step * 1e-3) / 1e9 / peak if world == 1 else None,
                "stage_ms": {k: round(v, 4) for k, v in stage_avg.items()}}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                roof["traffic"] = json.load(open(tpath)).get(top)
            except Exception:
                pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->u32",
        "data": "synthetic", "tokens_per_s": total_tokens / (ms_per_step * 1e-3),
        "config": workload_config(args, int(total_docs), int(total_bytes)),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_bytes + 8 * (n_docs + 1)),
                "d2h_bytes_per_step": int(4 * n_tokens + 8 * (n_docs + 1)), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "tokens_per_s": total_tokens * e2e_steps / e2e_s, "api": "tk_encode_batch (pinned host text in, pinned host ids out)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "decode": {"value": total_bytes / (dec_ms * 1e-3) / 1e9, "unit": "GB/s of text out", "ms_per_step": dec_ms,
                   "tokens_per_s": total_tokens / (dec_ms * 1e-3), "roundtrip_byte_exact": roundtrip_ok,
                   "hbm_frac": (4 * n_tokens + n_bytes + 16 * (n_docs + 1)) / (dec_ms * 1e-3) / 1e9 / peak if world == 1 else None},
        "tokens_per_step": int(total_tokens), "bytes_per_token": total_bytes / max(total_tokens, 1.0),
    }
    if numa:
        line["host_binding"] = numa

    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import tekken_oracle as TO
        threads = host_threads()
        orc = TO.OracleTekkenizer.from_file(path)
        sdata, soff, desc = cpu_sample(args.docs, threads)
        gbs, tps, secs = cpu_rate(orc, sdata, soff, threads, reps=2)
        # the sample doubles as a parity check of this very run
        ns = len(soff) - 1
        rid, roff = orc.encode_batch_np(sdata, soff, True, True, n_threads=threads)
        got = d_tok[:int(roff[-1])].cpu().numpy().view(np.uint32)
        line["cpu_baseline"] = {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc, "tokens_per_s": tps,
                                "seconds": secs, "ids_match_gpu_on_sample": bool(np.array_equal(got, rid)), "sample_docs": ns}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000, help="documents per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="mixed", choices=["mixed", "english"],
                    help="mixed = BASELINE configs[1] (the bench line); english = configs[0]'s English-like text cut into 1 KiB documents (context only)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch one rank per GPU the way the driver does
        import socket
        import subprocess
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        return subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                                "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
