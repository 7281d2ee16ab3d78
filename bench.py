#!/usr/bin/env python
"""bench.py -- encode throughput of the B200 Tekkenizer path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one pass of the hot path (Tekkenizer::encode over a batch, src/tekkenizer.rs:378-405)
over one batch of synthetic input.  Workloads (BASELINE.json `configs`):

  mixed         configs[1]  1,000,000 mixed-script documents x <= 1 KiB per GPU     (the bench line, default)
  english1m     configs[0]  one 1 MiB English-like document, encode(true,true) + decode(Keep)
  single1g      configs[2]  one 1 GiB document with cross-tile whitespace / digit / CR-LF runs
  adversarial   configs[3]  256 single pre-tokens x 64 KiB (repeated bytes, random letters, CJK, emoji)
  roundtrip64g  configs[4]  64 shards x 1 Mi documents: encode + decode round trip, shards dealt to the ranks
  english       context only: configs[0]'s text cut into 1 KiB documents

At N > 1 every rank works on its own documents (documents are independent: no collective on the
data path, weak scaling); single-document workloads run as N replicas with different seeds.

value  = input GB/s with text and ids resident in HBM (tk_encode_batch_device), CUDA events per step.
e2e    = the same metric through the host-buffer C ABI call (tk_encode_batch / tk_encode): pinned
         host text in, host ids out, H2D and D2H copies inside the timed region.
roofline / cpu_baseline: see DESIGN.md section 6.

--impl reference times the CPU arm: the oracle's C restatement of the reference engine on all host
cores (the reference is Rust and cannot be built in this image), same metric and config.
"""
from __future__ import annotations

import argparse
import ctypes
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
L2_BYTES = 126 << 20
SHARD_DOCS = (1 << 20)            # roundtrip64g: documents per shard


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


def family(stage: str) -> str:
    """Stage names of the library's per-launch timer -> kernel family (the nine lane-merge launches are one)."""
    if stage.startswith("lanemerge"):
        return "lanemerge"
    if stage.startswith("pretok"):
        return "pretok"
    if stage in ("longmark", "longmerge"):
        return "longpiece"
    return stage


# ---------------------------------------------------------------------------------------------- workloads

class Workload:
    def __init__(self, name, args):
        self.name, self.args = name, args
        self.single_doc = name in ("english1m", "single1g", "adversarial")

    def docs(self, rank: int, n_docs=None):
        """(bytes uint8, doc_off uint64[n+1]) of this rank's step."""
        from tekken_rs_b200 import corpus
        a = self.args
        one = lambda raw: (np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64))
        if self.name == "mixed":
            n = a.docs if n_docs is None else n_docs
            return corpus.mixed_script_docs(n, SEED, first_doc=rank * a.docs)
        if self.name == "english":
            n = a.docs if n_docs is None else n_docs
            base = np.frombuffer(corpus.english_like(1 << 26, 1234 + rank), dtype=np.uint8)
            data = np.tile(base, -(-n * 1024 // len(base)))[:n * 1024]
            return data, np.arange(n + 1, dtype=np.uint64) * np.uint64(1024)
        if self.name == "english1m":
            return one(corpus.english_like(1 << 20, 1234 + rank))
        if self.name == "single1g":
            return one(corpus.single_long_document(a.bytes or (1 << 30), 7 + rank))
        if self.name == "adversarial":
            return one(corpus.adversarial_pieces(a.pieces, 1 << 16, 11 + rank))
        raise ValueError(self.name)

    def label(self, n_docs, n_bytes):
        a = self.args
        return {
            "mixed": "BASELINE configs[1]: batch encode of %d synthetic mixed-script UTF-8 documents x <=1 KiB per GPU (seed %d), add_bos+add_eos" % (a.docs, SEED),
            "english": "context only (not a BASELINE config): %d documents x 1 KiB of configs[0]'s English-like ASCII text per GPU, add_bos+add_eos" % a.docs,
            "english1m": "BASELINE configs[0]: Tekkenizer::encode(text, true, true) + decode(Keep) of one 1 MiB English-like document (examples/basic_tokenizer_test.rs path)",
            "single1g": "BASELINE configs[2]: one %.2f GiB synthetic document (whitespace / newline / digit runs across tile boundaries, seed 7), add_bos+add_eos" % (n_bytes / 2**30),
            "adversarial": "BASELINE configs[3]: %d adversarial single pre-tokens x 64 KiB in one document (repeated bytes, random letters, CJK, emoji, mixed scripts)" % a.pieces,
            "roundtrip64g": "BASELINE configs[4]: %d shards x %d mixed-script documents (%.1f GB in all), encode + decode round trip, shards dealt to the ranks" % (
                a.shards, SHARD_DOCS, n_bytes / 1e9),
        }[self.name]

    def config(self, world, n_docs, n_bytes, flushed):
        return {"workload": self.label(n_docs, n_bytes), "name": self.name, "docs_per_step": int(n_docs), "bytes_per_step": int(n_bytes),
                "parallelism": "documents sharded by rank, no collective" if not self.single_doc else "one document per GPU (replicas at N > 1), no collective",
                "l2": "L2 flushed between steps (256 MiB write), outside the per-step events" if flushed else
                      "input per GPU per step (%.0f MB) exceeds the 126 MB L2; no flush needed" % (n_bytes / max(1, world) / 1e6)}

    def cpu_sample(self, threads):
        """A bounded sample of the workload for the CPU arm: (data, off, description)."""
        a = self.args
        if self.name in ("mixed", "english"):
            n = int(min(a.docs, max(16384, 32768 * threads)))
            data, off = self.docs(0, n)
            return data, off, "first %d of %d documents of the workload (%.0f MB)" % (n, a.docs, len(data) / 1e6)
        if self.name == "english1m":
            data, off = self.docs(0)
            return data, off, "the whole 1 MiB document (one encode call: single-threaded by construction)"
        if self.name == "single1g":
            from tekken_rs_b200 import corpus
            raw = corpus.single_long_document(1 << 25, 7)
            return np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64), \
                "a 32 MiB document from the same generator (one encode call: single-threaded by construction)"
        if self.name == "adversarial":
            from tekken_rs_b200 import corpus
            k = max(7, min(a.pieces, threads))
            raw = corpus.adversarial_pieces(k, 1 << 16, 11).split(b" ")
            off = np.zeros(len(raw) + 1, dtype=np.uint64)
            np.cumsum([len(x) for x in raw], out=off[1:])
            return np.frombuffer(b"".join(raw), dtype=np.uint8), off, \
                "%d of the %d pieces, one per document so that the host threads share them" % (k, a.pieces)
        raise ValueError(self.name)


def cpu_baselines(wl: Workload, path: str, engine_budget_s: float = 12.0):
    """cpu_baseline object: the C restatement (kind "port") on all host threads, and beside it the upstream engine
    (tiktoken's Rust CoreBPE, what tiktoken-rs vendors) in N worker processes, both on a bounded sample."""
    from oracle import engine_baseline
    from oracle import tekken_oracle as TO
    threads = 1 if wl.name in ("english1m", "single1g") else host_threads()
    orc = TO.OracleTekkenizer.from_file(path)
    data, off, desc = wl.cpu_sample(threads)
    best, n_tok = None, 0
    for _ in range(2):
        t0 = time.perf_counter()
        n_tok = orc.encode_count_mt(data, off, True, True, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    out = {"value": len(data) / best / 1e9, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc + ", all ids computed, %d threads" % threads,
           "tokens_per_s": n_tok / best, "seconds": best,
           "note": "C restatement of tiktoken-rs CoreBPE + Tekkenizer glue (oracle/tekken_oracle_core.c; pieces > 2 KiB use a heap instead of the engine's quadratic loop)"}
    try:
        out["engine"] = engine_baseline.measure(path, data, off, host_threads() if not wl.single_doc or wl.name == "adversarial" else 1,
                                                engine_budget_s, long_pieces=wl.name == "adversarial")
    except Exception as e:      # tiktoken missing on the box: say so, keep the port
        out["engine"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return out, orc, (data, off)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import tekken_oracle as TO
    from tekken_rs_b200 import assets
    wl = Workload(args.workload if args.workload != "roundtrip64g" else "mixed", args)
    threads = 1 if wl.name in ("english1m", "single1g") else host_threads()
    orc = TO.OracleTekkenizer.from_file(assets.ensure_tekken_json())
    data, off, desc = wl.cpu_sample(threads)
    for _ in range(args.warmup):
        orc.encode_count_mt(data, off, True, True, threads)
    t0 = time.perf_counter()
    n_tok = 0
    for _ in range(args.steps):
        n_tok = orc.encode_count_mt(data, off, True, True, threads)
    dt = time.perf_counter() - t0
    gbs = len(data) * args.steps / dt / 1e9
    real = Workload(args.workload, args)
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8->u32", "data": "synthetic",
        "tokens_per_s": n_tok * args.steps / dt,
        "config": real.config(args.gpus, len(off) - 1, len(data), False) if args.workload != "roundtrip64g" else
                  {"workload": real.label(args.shards * SHARD_DOCS, 0), "name": real.name},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc + ", all ids computed, %d threads" % threads,
                         "note": "C restatement of tiktoken-rs CoreBPE + Tekkenizer glue (oracle/tekken_oracle_core.c); the Rust reference cannot be built here"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


# ---------------------------------------------------------------------------------------------- GPU arm

_JSON_FD = None


def emit_line(line):
    """The one JSON line, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, text.encode())


class Ctx:
    """torch / torch.distributed plumbing shared by the GPU legs."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the encode path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.numa = None
        if self.world > 1:
            # one process per GPU: run on (and allocate the pinned staging buffers from) the CPUs next to this GPU
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local))
                self.numa = "cpu affinity set to GPU %d's NUMA node (%d cpus)" % (self.local, len(os.sched_getaffinity(0)))
            except Exception as e:   # not fatal: the numbers are then just measured without the binding
                self.numa = "cpu affinity not set (%s)" % type(e).__name__
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def pinned_copy(torch, a: np.ndarray, pad: int = 64):
    h = torch.empty(len(a) + pad, dtype=torch.uint8).pin_memory()
    h[:len(a)].numpy()[:] = a
    return h


def run_ours(args):
    cx = Ctx()
    torch = cx.torch
    from tekken_rs_b200 import Tekkenizer, _lib, assets, kernel_launch_count
    path = assets.ensure_tekken_json() if cx.rank == 0 else None
    if cx.world > 1:
        cx.dist.barrier()
        path = assets.ensure_tekken_json()
    tk = Tekkenizer.from_file(path, device=cx.local, split=1 if args.split == "config" else 0)
    lib = _lib.load()
    if args.workload == "roundtrip64g":
        return run_roundtrip(args, cx, tk, lib, path)
    wl = Workload(args.workload, args)
    stream = cx.stream

    # this rank's input, in pinned host memory (the e2e arm reads it from there) and in HBM
    data_np, off_np = wl.docs(cx.rank)
    n_docs, n_bytes = len(off_np) - 1, len(data_np)
    h_data = pinned_copy(torch, data_np)
    h_off = torch.from_numpy(off_np.astype(np.int64)).pin_memory()
    del data_np
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
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        n_tokens = step_device()
    need_flush = n_bytes <= 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None
    clocks = ClockSampler(cx.local)
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    cx.barrier()
    launches0 = kernel_launch_count()
    clocks.start()
    for a, b in pairs:
        if need_flush:
            flush_buf.fill_(1)
        a.record(stream)
        step_device()
        b.record(stream)
    torch.cuda.synchronize()
    clocks.stop()
    launches = kernel_launch_count() - launches0
    dev_ms = cx.max(sum(a.elapsed_time(b) for a, b in pairs))
    cx.barrier()
    # per-stage times: a separate pass with the library's per-launch events on (kept out of the timed steps)
    tk.set_stage_timing(True)
    stage_ms, stage_steps = {}, max(1, min(args.steps, 5))
    for _ in range(stage_steps):
        if need_flush:
            flush_buf.fill_(1)
        step_device()
        for k, v in tk.last_stage_times().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / stage_steps
    tk.set_stage_timing(False)
    counters = tk.last_encode_counters()
    total_bytes, total_tokens, total_docs = cx.sum(float(n_bytes)), cx.sum(float(n_tokens)), cx.sum(float(n_docs))
    ms_per_step = dev_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------
    h_view = h_data[:n_bytes].numpy()
    h_off_u64 = h_off.numpy().view(np.uint64)

    def step_host():
        tok, toff, n_out = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        if wl.single_doc:       # the reference's own call: one text (Tekkenizer::encode)
            rc = lib.tk_encode(tk._h, h_view.ctypes.data, n_bytes, 1, 1, ctypes.byref(tok), ctypes.byref(n_out))
            if rc != 0:
                raise RuntimeError(lib.tk_last_error().decode())
            last = n_out.value
        else:
            rc = lib.tk_encode_batch(tk._h, h_view.ctypes.data, h_off_u64.ctypes.data, n_docs, 1, 1, ctypes.byref(tok), ctypes.byref(toff))
            if rc != 0:
                raise RuntimeError(lib.tk_last_error().decode())
            last = ctypes.c_uint64.from_address(toff.value + 8 * n_docs).value      # read the result on the host
            lib.tk_buffer_free(toff)
        first = ctypes.c_uint32.from_address(tok.value).value
        lib.tk_buffer_free(tok)
        return last, first

    e2e_steps = max(1, min(args.steps, 10)) if not args.no_e2e else 1
    for _ in range(2 if not args.no_e2e else 0):
        n_e2e, first = step_host()
        assert n_e2e == n_tokens and first == 1, (n_e2e, n_tokens, first)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = cx.max(time.perf_counter() - t0)
    cx.barrier()
    e2e_value = total_bytes * e2e_steps / e2e_s / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_bytes + 8 * (n_docs + 1)),
           "d2h_bytes_per_step": int(4 * n_tokens + 8 * (n_docs + 1)), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "tokens_per_s": total_tokens * e2e_steps / e2e_s,
           "api": "tk_encode (one text, pinned host in, host ids out)" if wl.single_doc else "tk_encode_batch (pinned host text in, pinned host ids out)"}
    if args.no_e2e:
        e2e["note"] = "--no-e2e (profiling run): one unwarmed call, not a measurement"
    if not wl.single_doc and not args.quick and not args.no_e2e:
        # the same call on PAGEABLE caller memory (what a Rust &str / Vec<u8> is): the library stages it itself
        pg = np.empty(n_bytes + 64, dtype=np.uint8)
        pg[:n_bytes] = h_view
        pg_off = off_np.copy()

        def step_pageable():
            tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
            rc = lib.tk_encode_batch(tk._h, pg.ctypes.data, pg_off.ctypes.data, n_docs, 1, 1, ctypes.byref(tok), ctypes.byref(toff))
            if rc != 0:
                raise RuntimeError(lib.tk_last_error().decode())
            lib.tk_buffer_free(tok)
            lib.tk_buffer_free(toff)
        step_pageable()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, e2e_steps // 2)):
            step_pageable()
        pg_s = cx.max(time.perf_counter() - t0) / max(1, e2e_steps // 2)
        cx.barrier()
        e2e["pageable_input"] = {"value": total_bytes / pg_s / 1e9, "unit": UNIT, "ms_per_step": pg_s * 1e3,
                                 "note": "caller's text and offsets in ordinary (pageable) host memory"}
        del pg

    # ---- decode (secondary: Tekkenizer::decode, src/tekkenizer.rs:436-560) --------------------
    d_out = torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda")
    d_boff = torch.empty(n_docs + 1, dtype=torch.int64, device="cuda")

    def step_decode():
        return tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), n_docs, n_tokens, 0, d_out.data_ptr(), n_bytes + 64,
                                      d_boff.data_ptr(), 0, stream.cuda_stream)
    for _ in range(2):
        nb = step_decode()
    roundtrip_ok = bool(nb == n_bytes and torch.equal(d_out[:n_bytes], d_data[:n_bytes]))
    dsteps = max(1, min(args.steps, 10))
    dpairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(dsteps)]
    cx.barrier()
    for a, b in dpairs:
        if need_flush:
            flush_buf.fill_(1)
        a.record(stream)
        step_decode()
        b.record(stream)
    torch.cuda.synchronize()
    dec_ms = cx.max(sum(a.elapsed_time(b) for a, b in dpairs)) / dsteps
    cx.barrier()
    peak, peak_src = measured_peak_gbs()
    dec_alg = 4 * n_tokens + n_bytes + 16 * (n_docs + 1)
    decode = {"value": total_bytes / (dec_ms * 1e-3) / 1e9, "unit": "GB/s of text out", "ms_per_step": dec_ms,
              "tokens_per_s": total_tokens / (dec_ms * 1e-3), "roundtrip_byte_exact": roundtrip_ok,
              "algorithmic_bytes": int(dec_alg), "hbm_frac": dec_alg / (dec_ms * 1e-3) / 1e9 / peak if cx.world == 1 else None}
    if wl.single_doc and cx.world == 1:
        # the reference-facing call of configs[0]: decode(ids, Keep) host to host
        ids_h = d_tok[:n_tokens].cpu().numpy().view(np.uint32)
        t0 = time.perf_counter()
        for _ in range(3):
            txt = tk.decode_bytes(ids_h, 1)
        decode["e2e_ms_tk_decode_keep"] = (time.perf_counter() - t0) / 3 * 1e3
        decode["keep_ok"] = bool(txt == b"<s>" + h_view.tobytes() + b"</s>")

    # ---- roofline of the dominant kernel family --------------------------------------------------
    alg = algorithmic_bytes(n_bytes, n_tokens, n_docs)
    fam = {}
    for k, v in stage_ms.items():
        if k != "setup":
            fam[family(k)] = fam.get(family(k), 0.0) + v
    top = max(fam, key=fam.get) if fam else None
    roof = None
    if top:
        ach = alg / (fam[top] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "kernel_ms": fam[top], "algorithmic_bytes_per_launch": alg,
                "note": "dominant kernel FAMILY by CUDA events on the launch stream (the nine lane-merge launches count as one kernel)",
                "whole_path_achieved": alg / (ms_per_step * 1e-3) / 1e9 if cx.world == 1 else None,
                "whole_path_frac": alg / (ms_per_step * 1e-3) / 1e9 / peak if cx.world == 1 else None,
                "family_ms": {k: round(v, 4) for k, v in fam.items()},
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and wl.name == "mixed":
            try:
                tj = json.load(open(tpath))
                roof["traffic"] = tj.get("families", {}).get(top)
                roof["traffic_whole_step"] = tj.get("step_total")
                roof["traffic_source"] = tj.get("source")
            except Exception:
                pass
    merge = None
    if counters and fam.get("lanemerge"):
        s = fam["lanemerge"] * 1e-3
        merge = {"pieces_merged": counters["pieces_queued"], "pieces_by_class": counters["by_class"], "long_pieces": counters["n_long"],
                 "pair_table_lookups": counters["pair_lookups"], "byte_pair_lookups": counters["byte_pair_lookups"],
                 "hash_probes_per_s": (counters["pair_lookups"] + counters["byte_pair_lookups"]) / s,
                 "pieces_per_s": counters["pieces_queued"] / s, "lanemerge_ms": fam["lanemerge"],
                 "l2": "see profiles/ (lts__t_sector_hit_rate of the lane-merge launches)"}
    if counters and counters.get("long_piece_rounds") and counters["n_huge"]:
        line_extra = {"pieces_over_512_bytes": counters["n_huge"], **counters["long_piece_rounds"]}
    else:
        line_extra = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": cx.world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->u32",
        "data": "synthetic", "tokens_per_s": total_tokens / (ms_per_step * 1e-3),
        "config": wl.config(cx.world, int(total_docs), int(total_bytes), need_flush),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "merge_stage": merge,
        "long_piece_stage": line_extra, "decode": decode,
        "tokens_per_step": int(total_tokens), "bytes_per_token": total_bytes / max(total_tokens, 1.0),
    }
    if cx.numa:
        line["host_binding"] = cx.numa

    # ---- one call, all GPUs (rank 0 drives every device of the box through tk_encode_batch_multi) ----
    if cx.world > 1 and not wl.single_doc and not args.quick:
        line["e2e_one_call_all_gpus"] = one_call_all_gpus(cx, lib, path, h_view, h_off_u64, n_docs, n_bytes, n_tokens)

    if cx.rank == 0 and cx.world == 1 and not args.no_cpu:
        cb, orc, (sdata, soff) = cpu_baselines(wl, path)
        # the sample doubles as a parity check of this very run
        if wl.name in ("mixed", "english", "english1m"):
            rid, roff = orc.encode_batch_np(sdata, soff, True, True, n_threads=host_threads(), mode=4 if args.split == "config" else 0)
            got = d_tok[:int(roff[-1])].cpu().numpy().view(np.uint32)
            cb["ids_match_gpu_on_sample"] = bool(np.array_equal(got, rid))
            cb["sample_docs"] = len(soff) - 1
        line["cpu_baseline"] = cb
    if args.split == "config":
        line["config"]["split"] = "TK_SPLIT_CONFIG: the pattern stored in tekken.json (not what the reference computes; SURVEY 8f rank 1)"
    if cx.rank == 0 and cx.world == 1 and args.latency:
        line["latency"] = latency_table(tk, lib, path)
    cx.close()
    if cx.rank == 0:
        emit_line(line)
    return 0


def one_call_all_gpus(cx, lib, path, h_view, h_off_u64, n_docs, n_bytes, n_tokens):
    """Strong scaling of ONE host call: rank 0's batch through tk_encode_batch_multi over all the box's GPUs
    (the other ranks idle at the barrier; their contexts stay alive).  Returns the object for the bench line."""
    res = None
    cx.barrier()
    if cx.rank == 0:
        from tekken_rs_b200 import Tekkenizer
        handles = [Tekkenizer.from_file(path, device=g) for g in range(cx.world)]
        arr = (ctypes.c_void_p * cx.world)(*[h._h for h in handles])

        def step(k):
            tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
            rc = lib.tk_encode_batch_multi(arr, k, h_view.ctypes.data, h_off_u64.ctypes.data, n_docs, 1, 1, ctypes.byref(tok), ctypes.byref(toff))
            if rc != 0:
                raise RuntimeError(lib.tk_last_error().decode())
            last = ctypes.c_uint64.from_address(toff.value + 8 * n_docs).value
            lib.tk_buffer_free(tok)
            lib.tk_buffer_free(toff)
            return last
        res = {"api": "tk_encode_batch_multi (one call, one host thread + pipeline per GPU, stitched offsets)", "scaling": "strong",
               "bytes_per_call": int(n_bytes), "by_gpus": {}}
        k = 1
        while k <= cx.world:
            for _ in range(2):
                assert step(k) == n_tokens
            t0 = time.perf_counter()
            for _ in range(3):
                step(k)
            dt = (time.perf_counter() - t0) / 3
            res["by_gpus"][str(k)] = {"ms_per_call": dt * 1e3, "GB/s": n_bytes / dt / 1e9}
            k *= 2
        for h in handles:
            h.close()
    cx.barrier()
    return res


def latency_table(tk, lib, path):
    """Host-to-host latency of the reference's own call shape -- one string per call (src/tekkenizer.rs:378-405) --
    at four sizes, next to the CPU port and the upstream engine on one core."""
    from oracle import tekken_oracle as TO
    from tekken_rs_b200 import corpus
    orc = TO.OracleTekkenizer.from_file(path)
    try:
        enc = TO.tiktoken_engine(orc.ranks)
    except Exception:
        enc = None
    rows = []
    for label, raw in (("13 B (\"Hello, world!\")", b"Hello, world!"), ("1 KiB", corpus.english_like(1 << 10)),
                       ("64 KiB", corpus.english_like(1 << 16)), ("1 MiB", corpus.english_like(1 << 20))):
        a = np.frombuffer(raw, dtype=np.uint8)
        reps = 2000 if len(raw) <= 1024 else 200 if len(raw) <= 65536 else 20

        def gpu_once():
            out, n = ctypes.c_void_p(), ctypes.c_size_t()
            rc = lib.tk_encode(tk._h, a.ctypes.data, len(a), 1, 1, ctypes.byref(out), ctypes.byref(n))
            assert rc == 0
            lib.tk_buffer_free(out)
            return n.value
        for _ in range(10):
            gpu_once()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            gpu_once()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        row = {"input": label, "bytes": len(raw), "gpu_us_median": ts[len(ts) // 2] * 1e6, "gpu_us_p90": ts[int(len(ts) * 0.9)] * 1e6}
        text = raw.decode()
        creps = max(3, reps // 20)
        t0 = time.perf_counter()
        for _ in range(creps):
            orc.encode_np(a, True, True)
        row["cpu_port_us"] = (time.perf_counter() - t0) / creps * 1e6
        if enc is not None:
            t0 = time.perf_counter()
            for _ in range(creps):
                enc.encode_ordinary(text)
            row["cpu_engine_us"] = (time.perf_counter() - t0) / creps * 1e6
        # ... and the way back: Tekkenizer::decode of that string's ids (src/tekkenizer.rs:436-443)
        ids = np.asarray(orc.encode_np(a, True, True), dtype=np.uint32)

        def gpu_dec():
            out, n = ctypes.c_void_p(), ctypes.c_size_t()
            rc = lib.tk_decode(tk._h, ids.ctypes.data, len(ids), 0, ctypes.byref(out), ctypes.byref(n))
            assert rc == 0 and n.value == len(raw)
            lib.tk_buffer_free(out)
        for _ in range(10):
            gpu_dec()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            gpu_dec()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        row["decode_ids"] = int(len(ids))
        row["decode_gpu_us_median"] = ts[len(ts) // 2] * 1e6
        t0 = time.perf_counter()
        for _ in range(creps):
            orc.decode_bytes(ids, "Ignore")
        row["decode_cpu_port_us"] = (time.perf_counter() - t0) / creps * 1e6
        if enc is not None:
            plain = [int(i) - 1000 for i in ids.tolist() if i >= 1000]
            t0 = time.perf_counter()
            for _ in range(creps):
                enc.decode_bytes(plain)
            row["decode_cpu_engine_us"] = (time.perf_counter() - t0) / creps * 1e6
        rows.append(row)
    return rows


# ---------------------------------------------------------------------------------------------- configs[4]

def _gen_chunk(first):
    from tekken_rs_b200 import corpus
    return corpus.mixed_script_docs(1 << 14, SEED, first_doc=first)


def run_roundtrip(args, cx, tk, lib, path):
    """BASELINE configs[4]: `--shards` shards of 2^20 mixed-script documents (64 shards = 64.8 GB), dealt round-robin
    to the ranks.  Per shard, device-resident: encode, decode, byte compare of decode(encode(x)) with x on the
    device, 64-bit order-sensitive checksum of the ids (xor over shards: equal for every GPU count).  Then the same
    shard host-to-host through tk_encode_batch + tk_decode_batch (the e2e number).  Shards are generated on the
    host by worker processes while the GPU works on the previous one."""
    import multiprocessing as mp
    torch = cx.torch
    stream = cx.stream
    mine = list(range(cx.rank, args.shards, cx.world))
    procs = max(1, min(16, host_threads() // max(1, cx.world)))
    tot_b = tot_t = tot_d = 0
    enc_ms = dec_ms = 0.0
    e2e_s = 0.0
    chk = 0
    ok = True
    oracle_docs = oracle_ok = 0
    clocks = ClockSampler(cx.local)
    launches0 = None
    from tekken_rs_b200 import kernel_launch_count
    orc = None
    if cx.rank == 0 and not args.no_cpu:
        from oracle import tekken_oracle as TO
        orc = TO.OracleTekkenizer.from_file(path)

    def gen(pool, s):
        return pool.map_async(_gen_chunk, [s * SHARD_DOCS + i for i in range(0, SHARD_DOCS, 1 << 14)])

    def assemble(parts):
        data = np.concatenate([p[0] for p in parts])
        lens = np.concatenate([np.diff(p[1].astype(np.int64)) for p in parts])
        off = np.zeros(len(lens) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        return data, off

    with mp.get_context("fork").Pool(procs) as pool:
        pending = gen(pool, mine[0]) if mine else None
        cx.barrier()
        clocks.start()
        launches0 = kernel_launch_count()
        for i, s in enumerate(mine):
            data, off = assemble(pending.get())
            pending = gen(pool, mine[i + 1]) if i + 1 < len(mine) else None
            n, nd = len(data), len(off) - 1
            h_data = pinned_copy(torch, data)
            d_data = h_data.cuda()
            d_off = torch.from_numpy(off.astype(np.int64)).cuda()
            cap = n + 2 * nd + 2
            d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
            d_toff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
            d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            d_boff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            if i == 0:      # warm-up of the first shard (workspace allocation, first launch of every kernel)
                nt0 = tk.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), nd, n, True, True, d_tok.data_ptr(), cap, d_toff.data_ptr(), stream.cuda_stream)
                tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), nd, nt0, 0, d_out.data_ptr(), n + 64, d_boff.data_ptr(), 0, stream.cuda_stream)
            torch.cuda.synchronize()
            ev[0].record(stream)
            ntok = tk.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), nd, n, True, True, d_tok.data_ptr(), cap, d_toff.data_ptr(), stream.cuda_stream)
            ev[1].record(stream)
            nb = tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), nd, ntok, 0, d_out.data_ptr(), n + 64, d_boff.data_ptr(), 0, stream.cuda_stream)
            ev[2].record(stream)
            torch.cuda.synchronize()
            good = nb == n and bool(torch.equal(d_out[:n], d_data[:n])) and bool(torch.equal(d_boff, d_off))
            ok &= good
            ids = d_tok[:ntok].to(torch.int64)
            idx = torch.arange(1, ntok + 1, device="cuda", dtype=torch.int64)
            chk ^= (int(((ids + 0x9E3779B9) * (idx * 0x85EBCA6B + 1)).sum().item()) + 0x1000003 * s) & 0xFFFFFFFFFFFFFFFF
            del ids, idx
            enc_ms += ev[0].elapsed_time(ev[1])
            dec_ms += ev[1].elapsed_time(ev[2])
            # oracle on every 64th document of the shard (rank 0)
            if orc is not None:
                sel = np.arange(0, nd, 64)
                sub = [data[int(off[d]):int(off[d + 1])] for d in sel]
                soff = np.zeros(len(sub) + 1, dtype=np.uint64)
                np.cumsum([len(x) for x in sub], out=soff[1:])
                rid, roff = orc.encode_batch_np(np.concatenate(sub), soff, True, True, n_threads=host_threads())
                toff_h = d_toff.cpu().numpy()
                tok_h = d_tok[:ntok].cpu().numpy().view(np.uint32)
                got = np.concatenate([tok_h[int(toff_h[d]):int(toff_h[d + 1])] for d in sel])
                oracle_docs += len(sel)
                oracle_ok += int(np.array_equal(got, rid))
            # host-to-host round trip of the same shard
            del d_tok, d_out
            hv = h_data[:n].numpy()
            t0 = time.perf_counter()
            tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
            rc = lib.tk_encode_batch(tk._h, hv.ctypes.data, off.ctypes.data, nd, 1, 1, ctypes.byref(tok), ctypes.byref(toff))
            assert rc == 0, lib.tk_last_error()
            out, boff = ctypes.c_void_p(), ctypes.c_void_p()
            rc = lib.tk_decode_batch(tk._h, tok.value, toff.value, nd, 0, ctypes.byref(out), ctypes.byref(boff), None)
            assert rc == 0, lib.tk_last_error()
            e2e_s += time.perf_counter() - t0
            back = np.frombuffer((ctypes.c_uint8 * n).from_address(out.value), dtype=np.uint8)
            ok &= bool(np.array_equal(back, hv))
            for p in (tok, toff, out, boff):
                lib.tk_buffer_free(p)
            tot_b += n; tot_t += ntok; tot_d += nd
            del d_data, h_data
        clocks.stop()
    launches = kernel_launch_count() - launches0
    t_enc, t_dec, t_e2e = cx.max(enc_ms), cx.max(dec_ms), cx.max(e2e_s)
    B, T, D = cx.sum(float(tot_b)), cx.sum(float(tot_t)), cx.sum(float(tot_d))
    all_ok = cx.sum(0.0 if ok else 1.0) == 0.0
    # xor of the per-rank checksums
    if cx.world > 1:
        t = torch.tensor([chk >> 32, chk & 0xFFFFFFFF], dtype=torch.int64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(cx.world)]
        cx.dist.all_gather(parts, t)
        chk = 0
        for p in parts:
            chk ^= (int(p[0].item()) << 32) | int(p[1].item())
    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(int(B), int(T), int(D)) + 4 * int(T) + int(B) + 16 * int(D)
    rt_ms = t_enc + t_dec
    line = {
        "metric": "encode_decode_roundtrip_input_throughput", "value": B / (rt_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": cx.world,
        "steps": len(mine), "warmup": 1, "ms_per_step": rt_ms / max(1, len(mine)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8->u32->u8", "data": "synthetic", "tokens_per_s": T / (rt_ms * 1e-3),
        "config": {"workload": Workload("roundtrip64g", args).label(int(D), int(B)), "name": "roundtrip64g", "shards": args.shards,
                   "docs": int(D), "bytes": int(B), "ids": int(T), "parallelism": "shards dealt round-robin to the ranks, no collective",
                   "l2": "every shard (about 1 GB) exceeds the 126 MB L2; no flush needed"},
        "encode": {"value": B / (t_enc * 1e-3) / 1e9, "unit": UNIT, "ms_total": t_enc},
        "decode": {"value": B / (t_dec * 1e-3) / 1e9, "unit": "GB/s of text out", "ms_total": t_dec,
                   "hbm_frac": (4 * T + B + 16 * D) / (t_dec * 1e-3) / 1e9 / peak / cx.world},
        "e2e": {"value": B / t_e2e / 1e9, "unit": UNIT, "h2d_bytes_per_step": int((B + 4 * T + 16 * D) / max(1, args.shards)),
                "d2h_bytes_per_step": int((B + 4 * T + 16 * D) / max(1, args.shards)), "seconds": t_e2e,
                "api": "tk_encode_batch then tk_decode_batch per shard (pinned host text in, host text out, compared on the host)"},
        "roofline": {"bound": "hbm", "kernel": "whole round trip (encode + decode kernels of every shard)", "achieved": alg / (rt_ms * 1e-3) / 1e9 / cx.world,
                     "peak": peak, "unit": "GB/s", "frac": alg / (rt_ms * 1e-3) / 1e9 / peak / cx.world, "traffic": None, "peak_source": peak_src,
                     "note": "per GPU; algorithmic bytes = encode (text in, ids out) + decode (ids in, text out)"},
        "checks": {"roundtrip_byte_exact_all_shards": bool(all_ok), "ids_checksum64_xor_over_shards": "%016x" % chk,
                   "oracle_sample_docs": oracle_docs, "oracle_sample_shards_equal": oracle_ok, "oracle_sample_shards": len(mine) if orc is not None else 0},
        "clocks": clocks.summary(), "gpu_launches": int(launches),
    }
    cx.close()
    if cx.rank == 0:
        emit_line(line)
    return 0 if all_ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000, help="documents per GPU per step (mixed / english)")
    ap.add_argument("--bytes", type=int, default=0, help="single1g: document size (default 1 GiB)")
    ap.add_argument("--pieces", type=int, default=256, help="adversarial: number of 64 KiB pieces")
    ap.add_argument("--shards", type=int, default=64, help="roundtrip64g: shards of 2^20 documents in all")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the pageable-input and one-call-all-GPUs legs")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: a single host-buffer call instead of the timed e2e loop")
    ap.add_argument("--latency", action="store_true", help="add the single-string latency table (13 B .. 1 MiB)")
    ap.add_argument("--split", default="reference", choices=["reference", "config"],
                    help="reference = the pattern the reference hard-codes (the bench line); config = the pattern stored in tekken.json (SURVEY 8f-1)")
    ap.add_argument("--workload", default="mixed", choices=["mixed", "english1m", "single1g", "adversarial", "roundtrip64g", "english"])
    args = ap.parse_args()
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
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # libraries below us write to stdout (NCCL prints its version banner there): keep fd 1 for the ONE JSON line
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
