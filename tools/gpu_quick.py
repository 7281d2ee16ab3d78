#!/usr/bin/env python
"""Developer smoke on a GPU box: CUDA path vs oracle on goldens, fuzz and the synthetic corpora."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tekken_oracle as TO  # noqa: E402
from tekken_rs_b200 import Tekkenizer, TokenizerError, assets, corpus  # noqa: E402


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if len(d) else n


def compare_batch(tk, orc, data, off, bos, eos, name):
    t0 = time.time()
    ids, toff = tk.encode_batch_np(data, off, bos, eos)
    t1 = time.time()
    rid, roff = orc.encode_batch_np(data, off, bos, eos)
    t2 = time.time()
    ok = np.array_equal(ids, rid) and np.array_equal(toff, roff)
    print("%-28s %s bytes=%d docs=%d tokens=%d gpu=%.3fs oracle=%.3fs" % (name, "OK" if ok else "MISMATCH", len(data), len(off) - 1,
                                                                        len(rid), t1 - t0, t2 - t1), flush=True)
    if not ok:
        if not np.array_equal(toff, roff):
            d = first_diff(toff, roff)
            print("   tok_off differ at doc", d, toff[max(0, d - 1):d + 2], roff[max(0, d - 1):d + 2])
            d = max(0, d - 1)
        else:
            i = first_diff(ids, rid)
            d = int(np.searchsorted(roff, i, side="right") - 1)
        a, b = int(off[d]), int(off[d + 1])
        print("   doc", d, "bytes", a, b, repr(bytes(data[a:b])[:200]))
        print("   gpu   ", ids[int(toff[d]):int(toff[d]) + 40].tolist())
        print("   oracle", rid[int(roff[d]):int(roff[d]) + 40].tolist())
    return ok


def main():
    path = assets.ensure_tekken_json()
    t0 = time.time()
    tk = Tekkenizer.from_file(path, device=0)
    print("load %.2fs" % (time.time() - t0))
    orc = TO.OracleTekkenizer.from_file(path)
    g = json.load(open(os.path.join(ROOT, "tests/golden/reference_goldens.json")))
    bad = 0
    for e in g["encode"]:
        ids = tk.encode(e["text"], False, False)
        if ids != e["ids"]:
            bad += 1
            print("GOLDEN MISMATCH", repr(e["text"]), ids, e["ids"])
    print("goldens: %d/%d ok" % (len(g["encode"]) - bad, len(g["encode"])))
    for d in g["decode"]:
        s = tk.decode(d["ids"], d["policy"])
        print("decode golden", s == d["text"], repr(s[:50]))
    allok = bad == 0
    # fuzz
    alpha = list("aAbsStTrReEvVmMlLdD'’ſ  \t\n\r\n 　 !?.,;-_()[]{}0123456789٣९①é́üßçñöКириллица中文日本語かなカナ한국어ไทยคำकाि्ाअ😀👍🏽‍👩‍💻€£∑√\x00\x7f\x85 ")
    rng = random.Random(5)
    texts = []
    for _ in range(30000):
        L = rng.choice([0, 1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100, 200])
        texts.append("".join(rng.choice(alpha) for _ in range(L)).encode())
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in texts], out=off[1:])
    data = np.frombuffer(b"".join(texts), dtype=np.uint8)
    allok &= compare_batch(tk, orc, data, off, True, True, "fuzz 30k docs bos+eos")
    allok &= compare_batch(tk, orc, data, off, False, False, "fuzz 30k docs")
    allok &= compare_batch(tk, orc, data, np.array([0, len(data)], dtype=np.uint64), False, True, "fuzz as one doc")
    eng = np.frombuffer(corpus.english_like(1 << 20), dtype=np.uint8)
    allok &= compare_batch(tk, orc, eng, np.array([0, len(eng)], dtype=np.uint64), True, True, "config1 english 1MiB")
    d2, o2 = corpus.mixed_script_docs(50000, 42)
    allok &= compare_batch(tk, orc, d2, o2, True, True, "config2 50k docs")
    c3 = np.frombuffer(corpus.single_long_document(1 << 24), dtype=np.uint8)
    allok &= compare_batch(tk, orc, c3, np.array([0, len(c3)], dtype=np.uint64), False, False, "config3 16MiB")
    c4 = np.frombuffer(corpus.adversarial_pieces(14, 1 << 12), dtype=np.uint8)
    allok &= compare_batch(tk, orc, c4, np.array([0, len(c4)], dtype=np.uint64), False, False, "config4 14x4KiB")
    c4 = np.frombuffer(corpus.adversarial_pieces(7, 1 << 16), dtype=np.uint8)
    allok &= compare_batch(tk, orc, c4, np.array([0, len(c4)], dtype=np.uint64), False, False, "config4 7x64KiB")
    # decode round trip
    ids, toff = tk.encode_batch_np(d2, o2, True, True)
    t0 = time.time()
    raw, boff = tk.decode_batch_np(ids, toff, "Ignore")
    print("decode round trip", np.array_equal(raw, d2) and np.array_equal(boff, o2), "%.3fs" % (time.time() - t0))
    allok &= np.array_equal(raw, d2) and np.array_equal(boff, o2)
    raw, boff = tk.decode_batch_np(ids[:2000], np.array([0, 2000], dtype=np.uint64), "Keep")
    ref = orc.decode_bytes(ids[:2000], "Keep") if True else b""
    print("decode keep vs oracle", raw.tobytes() == ref)
    allok &= raw.tobytes() == ref
    for bad_ids, pol in (([1, 22177], "Raise"), ([1000 + 0xE4, 1, 1000 + 0xB8], "Ignore"), ([200000], "Ignore")):
        try:
            tk.decode(bad_ids, pol)
            print("expected error missing", bad_ids, pol)
            allok = False
        except TokenizerError as e:
            try:
                orc.decode(bad_ids, pol)
                okk = False
            except TO.TokenizerError as oe:
                okk = oe.kind == e.kind
            print("decode error", bad_ids, pol, e.kind, "matches oracle:", okk)
            allok &= okk
    try:
        tk.encode(b"abc\xff", False, False)
        allok = False
    except TokenizerError as e:
        print("invalid utf8:", e)
    print("ALL OK" if allok else "FAILURES")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
