#!/usr/bin/env python
"""BASELINE config 5 at reduced host cost: S shards of 1,000,000 mixed-script documents each (the seed-42 corpus,
documents [shard*1M, (shard+1)*1M)), every shard encoded and decoded on the device with a byte-exact round-trip
check and a running 64-bit checksum of the ids.  Shards are independent, so the 64 GB configuration is the same
per-GPU loop with more shards (8 shards per GPU on 8 GPUs); under torchrun every rank takes its own shards.

  python tools/roundtrip_corpus.py --shards 8
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tekken_rs_b200 import corpus  # noqa: E402

CH = 1 << 14


def _gen(first):
    return corpus.mixed_script_docs(CH, 42, first_doc=first)


def shard(pool, index, n_docs):
    parts = pool.map(_gen, [index * n_docs + i for i in range(0, n_docs, CH)])
    data = np.concatenate([p[0] for p in parts])
    lens = np.concatenate([np.diff(p[1].astype(np.int64)) for p in parts])
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return data, off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--docs", type=int, default=(1 << 20) - (1 << 20) % CH)
    args = ap.parse_args()
    import torch

    from tekken_rs_b200 import Tekkenizer, assets
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    tk = Tekkenizer.from_file(assets.ensure_tekken_json(), device=local)
    st = torch.cuda.current_stream().cuda_stream
    tot_b = tot_t = 0
    t_enc = t_dec = 0.0
    chk = 0
    ok = True
    with mp.get_context("fork").Pool(min(16, len(os.sched_getaffinity(0)))) as pool:
        for s in range(args.shards):
            data, off = shard(pool, rank * args.shards + s, args.docs)
            n, nd = len(data), len(off) - 1
            d_data = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            d_data[:n] = torch.from_numpy(data).cuda()
            d_off = torch.from_numpy(off.astype(np.int64)).cuda()
            cap = n + 2 * nd + 2
            d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
            d_toff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
            d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            d_boff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ntok = tk.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), nd, n, True, True, d_tok.data_ptr(), cap, d_toff.data_ptr(), st)
            t1 = time.perf_counter()
            nb = tk.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), nd, ntok, 0, d_out.data_ptr(), n + 64, d_boff.data_ptr(), 0, st)
            t2 = time.perf_counter()
            good = nb == n and bool(torch.equal(d_out[:n], d_data[:n])) and bool(torch.equal(d_boff, d_off))
            ok &= good
            # order-sensitive checksum of the ids, on the device
            ids = d_tok[:ntok].to(torch.int64)
            idx = torch.arange(1, ntok + 1, device="cuda", dtype=torch.int64)
            chk ^= int(((ids + 0x9E3779B9) * (idx * 0x85EBCA6B + 1)).sum().item()) & 0xFFFFFFFFFFFFFFFF
            tot_b += n; tot_t += ntok; t_enc += t1 - t0; t_dec += t2 - t1
            print("rank %d shard %d: %.1f MB, %d ids, encode %.1f ms, decode %.1f ms, round trip %s" % (
                rank, s, n / 1e6, ntok, (t1 - t0) * 1e3, (t2 - t1) * 1e3, "byte-exact" if good else "MISMATCH"), flush=True)
            del d_data, d_tok, d_out
    print("rank %d of %d: %d shards, %.2f GB, %d ids: encode %.1f GB/s, decode %.1f GB/s, all round trips byte-exact: %s, id checksum %016x" % (
        rank, world, args.shards, tot_b / 1e9, tot_t, tot_b / t_enc / 1e9, tot_b / t_dec / 1e9, ok, chk))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
