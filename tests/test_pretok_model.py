"""The product's pre-tokeniser logic (tekken_rs_b200/csrc/tk_pretok.h -- the code each CUDA thread
runs on its 32-byte window) compiled for the host and compared with the oracle's regex split."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import tekken_oracle as TO
from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libpretok_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "native", "pretok_host.cpp"),
                           os.path.join(ROOT, "tekken_rs_b200", "csrc", "tk_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.pretok_host.restype = ctypes.c_int64
    lib.pretok_host.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
    return lib


def model_starts(lib, b, offs):
    a = np.concatenate([np.frombuffer(b, dtype=np.uint8), np.zeros(8, np.uint8)])
    n = len(b)
    mask = np.zeros(n // 32 + 1, dtype=np.uint32)
    do = np.array(offs, dtype=np.uint64)
    rc = lib.pretok_host(a.ctypes.data, n, do.ctypes.data, len(do) - 1, mask.ctypes.data)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n + 1]
    return rc, np.nonzero(bits)[0].tolist()


def oracle_starts(b, offs):
    core = TO.core()
    out = set([len(b)])
    for d in range(len(offs) - 1):
        s, e = offs[d], offs[d + 1]
        if e > s:
            seg = np.frombuffer(b[s:e], dtype=np.uint8)
            ends = np.zeros(len(seg), dtype=np.uint64)
            k = core.orc_split(seg.ctypes.data, len(seg), ends.ctypes.data)
            assert k > 0
            out.add(s)
            out.update(int(x) + s for x in ends[:k - 1])
        else:
            out.add(s)   # an empty document still marks its position
    return sorted(out)


def gen_case(rng):
    mode = rng.random()
    if mode < 0.5:
        s = "".join(rng.choice(FUZZ_ALPHABET) for _ in range(rng.choice([1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100, 200])))
    elif mode < 0.8:
        parts = []
        for _ in range(rng.randint(1, 8)):
            k = rng.choice("dwnplm")
            ln = rng.choice([1, 2, 3, 4, 31, 32, 33, 63, 64, 65, 70, 130])
            parts.append("".join(rng.choice({"d": "0123456789", "w": " \t  ", "n": "\n\r", "p": "!?.'",
                                            "l": "abc'sST", "m": " \n"}[k]) for _ in range(ln)))
        s = "".join(parts)
    else:
        s = "".join(rng.choice("ab '\n!1 ") for _ in range(rng.choice([40, 80, 160])))
    nd = rng.choice([1, 1, 1, 2, 3, 5])
    cuts = sorted(rng.randint(0, len(s)) for _ in range(nd - 1))
    offs = [0] + [len(s[:c].encode()) for c in cuts] + [len(s.encode())]
    return s.encode(), offs


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_piece_starts_match_oracle(model, seed):
    rng = random.Random(seed)
    for _ in range(12000):
        b, offs = gen_case(rng)
        rc, m = model_starts(model, b, offs)
        assert rc >= 0
        assert m == oracle_starts(b, offs), (b, offs)


def test_invalid_utf8_is_flagged(model):
    for bad, pos in ((b"abc\xff", 3), (b"\xe4\xb8", 0), (b"a\x80b", 1), (b"\xc0\xaf", 0), (b"\xed\xa0\x80", 0),
                     (b"\xf4\x90\x80\x80", 0), (b"x" * 31 + b"\xe4\xb8" + b"y" * 40, 31), (b"x" * 30 + b"\xe4\xb8\xad" + b"\xad", 33)):
        rc, _ = model_starts(model, bad, [0, len(bad)])
        assert rc == -1 - pos, (bad, rc)
    # a document boundary inside a char
    b = "中".encode()
    rc, _ = model_starts(model, b, [0, 1, 3])
    assert rc < 0


def test_digit_runs_at_every_offset(model):
    # \p{N}{1,3}: every third digit of a run starts a piece, counted from the run's start -- across windows, across
    # document starts inside the run, with one-byte and multi-byte digits (the window evaluator has a bit-arithmetic
    # path for windows whose digits are all one byte and a per-digit loop for the rest)
    rng = random.Random(11)
    for lead in list(range(0, 70, 1)):
        for run in (1, 2, 3, 4, 5, 29, 30, 31, 32, 33, 34, 35, 64, 65, 66, 97, 200):
            body = "".join(rng.choice("0123456789") for _ in range(run))
            for tail in ("", "a", " 12", "٣٤٥٦", "x٣" + "7" * 40):
                s = ("a" * lead + body + tail).encode()
                offs = [0, len(s)]
                rc, m = model_starts(model, s, offs)
                assert rc >= 0 and m == oracle_starts(s, offs), (s, offs)
                cut = lead + rng.randint(0, run)
                offs = [0, cut, len(s)]
                rc, m = model_starts(model, s, offs)
                assert rc >= 0 and m == oracle_starts(s, offs), (s, offs)
    s = ("٣" * 50 + "1234567" + "٣" * 3 + "89" * 40).encode()
    for cut in range(0, len(s), 2):
        try:
            s[:cut].decode()
        except UnicodeDecodeError:
            continue
        offs = [0, cut, len(s)]
        rc, m = model_starts(model, s, offs)
        assert rc >= 0 and m == oracle_starts(s, offs), (cut,)
