"""The bounds-checked build (-DTK_DEBUG_BOUNDS: every store of the encode and decode kernels into a workspace / staging /
output array is checked against the array's size) run over a parity corpus in a child process.  compute-sanitizer is closed on the GPU
pool this library is developed on; this is its substitute: zero refused stores, and results equal to the oracle's."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes, json, random, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import tekken_oracle as TO
from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET
from tekken_rs_b200 import Tekkenizer, _lib, assets, corpus, set_chunk_bytes
path = assets.ensure_tekken_json()
tk, orc = Tekkenizer.from_file(path, device=0), TO.OracleTekkenizer.from_file(path)
lib = _lib.load()
def pack(texts):
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in texts], out=off[1:])
    return np.frombuffer(b"".join(texts), dtype=np.uint8), off
rng = random.Random(41)
fuzz = ["".join(rng.choice(FUZZ_ALPHABET) for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100, 200, 700]))).encode() for _ in range(20000)]
cases = [pack(fuzz), corpus.mixed_script_docs(30000, 42),
         pack([corpus.adversarial_pieces(7, 1 << 12), corpus.single_long_document(1 << 22, 5), b"", b"x" * 5000, b" " * 70000 + b"\n", b"7" * 100001])]
ok = True
for chunk in (0, 256 << 10):
    set_chunk_bytes(chunk)
    for data, off in cases:
        for bos, eos in ((True, True), (False, False)):
            ids, toff = tk.encode_batch_np(data, off, bos, eos)
            rid, roff = orc.encode_batch_np(data, off, bos, eos, n_threads=8)
            ok &= bool(np.array_equal(ids, rid) and np.array_equal(toff, roff))
for t in fuzz[:500]:
    ok &= tk.encode(t, True, True) == orc.encode(t, True, True)
# ... and the decode kernels: the batches back to their bytes (Ignore) and with the specials kept, oversized tiles
# (long tokens), the single-sequence kernel, invalid runs
set_chunk_bytes(0)
for data, off in cases:
    ids, toff = tk.encode_batch_np(data, off, True, True)
    back, boff = tk.decode_batch_np(ids, toff, "Ignore")
    ok &= bool(np.array_equal(back, data) and np.array_equal(boff, off))
    kept, koff = tk.decode_batch_np(ids, toff, "Keep")
    ok &= len(kept) == len(data) + 7 * (len(off) - 1)
long_tok = max(orc.encode("internationalization " * 3, False, False), key=lambda i: len(orc.decode_bytes([i], "Ignore")))
for seq in ([long_tok] * 5000, [long_tok, 1, 1044] * 3000, orc.encode(" " * 3000 + "x", True, True) * 40):
    raw, _ = tk.decode_batch_np(np.asarray(seq, dtype=np.uint32), np.array([0, len(seq)], dtype=np.uint64), "Keep")
    ok &= raw.tobytes() == orc.decode_bytes(seq, "Keep")
for t in fuzz[:300]:
    ids = orc.encode(t, True, True)
    ok &= tk.decode_bytes(ids, "Keep") == orc.decode_bytes(ids, "Keep")
for bad in ([1228, 1184], [1228, 2, 1184, 1173], [1000 + 0xFF] * 70, [200000]):
    try:
        tk.decode_bytes(bad, "Ignore")
        ok = False
    except Exception:
        pass
d4 = (ctypes.c_uint64 * 4)()
n = lib.tk_debug_bounds_violations(tk._h, d4)
print(json.dumps({"parity": ok, "violations": n, "detail": list(d4)}))
"""


def test_debug_bounds_build_refuses_no_store():
    from tekken_rs_b200 import _build
    dbg = _build.build_debug()
    env = dict(os.environ, TEKKEN_B200_LIB=dbg, TEKKEN_B200_NO_BUILD="1")
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["violations"] == 0, "out-of-range store refused at line %d (tk_kernels.cu; + 1,000,000: tk_decode.cu), index %d, limit %d" % tuple(res["detail"][1:])
    assert res["parity"]


def test_regular_build_has_no_checks(gpu_tok):
    from tekken_rs_b200 import _lib
    assert _lib.load().tk_debug_bounds_violations(gpu_tok._h, None) == -1
