#!/usr/bin/env python
"""Small parity run meant to be wrapped in compute-sanitizer (memcheck / racecheck): every kernel of
the encode and decode paths on a few thousand documents, checked against the oracle."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tekken_oracle as TO  # noqa: E402
from tekken_rs_b200 import Tekkenizer, assets, corpus  # noqa: E402


def main():
    path = assets.ensure_tekken_json()
    tk = Tekkenizer.from_file(path, device=0)
    orc = TO.OracleTekkenizer.from_file(path)
    ok = True
    from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET
    rng = random.Random(1)
    texts = ["".join(rng.choice(FUZZ_ALPHABET) for _ in range(rng.choice([0, 1, 3, 9, 33, 70, 200]))).encode() for _ in range(1500)]
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in texts], out=off[1:])
    data = np.frombuffer(b"".join(texts), dtype=np.uint8)
    for d, o in ((data, off), corpus.mixed_script_docs(600, 3),
                 (np.frombuffer(corpus.adversarial_pieces(7, 3000), dtype=np.uint8), None),
                 (np.frombuffer(corpus.single_long_document(1 << 18), dtype=np.uint8), None)):
        if o is None:
            o = np.array([0, len(d)], dtype=np.uint64)
        ids, toff = tk.encode_batch_np(d, o, True, True)
        rid, roff = orc.encode_batch_np(d, o, True, True)
        good = np.array_equal(ids, rid) and np.array_equal(toff, roff)
        raw, boff = tk.decode_batch_np(ids, toff, "Ignore")
        good &= np.array_equal(raw, d) and np.array_equal(boff, o)
        print("case bytes=%d docs=%d ids=%d %s" % (len(d), len(o) - 1, len(ids), "OK" if good else "MISMATCH"), flush=True)
        ok &= good
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
