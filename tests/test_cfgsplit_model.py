"""Groundwork for the config-driven pattern: the device-side split functions of
tekken_rs_b200/csrc/tk_pretok_cfg.h (local "safe start" rules + one sequential matcher walk per segment), compiled for
the host and compared with the oracle's restatement of the pattern stored in tekken.json.  No kernel uses them yet."""
import bisect
import ctypes
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import tekken_oracle as TO
from oracle.tools.make_config_pattern_fixtures import ALPHABET
from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC = {"O": 0, "U": 1, "LO": 2, "C": 3, "M": 4, "N": 5, "W": 6}


def flat_classes():
    base = json.load(open(os.path.join(ROOT, "oracle", "unicode_tables.json")))
    sub = json.load(open(os.path.join(ROOT, "oracle", "unicode_subclasses.json")))
    flat = np.zeros(0x110000, dtype=np.uint8)
    for name, rs in (("W", base["S"]), ("N", base["N"]), ("U", sub["UPPER"]), ("LO", sub["LOWER"]), ("C", sub["BOTH"]),
                     ("M", sub["MARK"])):
        for a, b in rs:
            flat[a:b + 1] = CC[name]
    return flat


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("native") / "libcfgsplit_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "native", "cfgsplit_host.cpp"),
                           os.path.join(ROOT, "tekken_rs_b200", "csrc", "tk_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.cfgsplit_host.restype = ctypes.c_int64
    lib.cfgsplit_host.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
    lib.cfgsplit_set_classes.argtypes = [ctypes.c_void_p]
    lib.cfgsplit_use_library_tables.restype = ctypes.c_int64
    lib.cfgsplit_use_library_tables.argtypes = [ctypes.c_void_p]
    flat = flat_classes()
    # the library's own table builder (csrc/unicode_subclasses.inc) against the classes measured from the engine
    assert lib.cfgsplit_use_library_tables(flat.ctypes.data) == 0
    return lib


def model_starts(lib, b, offs):
    a = np.concatenate([np.frombuffer(b, dtype=np.uint8), np.zeros(8, np.uint8)])
    n = len(b)
    mask = np.zeros(n // 32 + 1, dtype=np.uint32)
    do = np.array(offs, dtype=np.uint64)
    rc = lib.cfgsplit_host(a.ctypes.data, n, do.ctypes.data, len(do) - 1, mask.ctypes.data)
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little")[:n]
    return rc, np.nonzero(bits)[0].tolist()


def oracle_starts(orc, b, offs):
    out = []
    for d in range(len(offs) - 1):
        pos = offs[d]
        for p in orc.split_config(b[offs[d]:offs[d + 1]]):
            out.append(pos)
            pos += len(p)
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_piece_starts_match_oracle(model, oracle, seed):
    rng = random.Random(seed)
    alpha = ALPHABET + FUZZ_ALPHABET
    for _ in range(10000):
        s = "".join(rng.choice(alpha) for _ in range(rng.choice([1, 2, 3, 5, 8, 13, 30, 64, 100, 200])))
        nd = rng.choice([1, 1, 1, 2, 3])
        cuts = sorted(rng.randint(0, len(s)) for _ in range(nd - 1))
        offs = [0] + [len(s[:c].encode()) for c in cuts] + [len(s.encode())]
        b = s.encode()
        rc, got = model_starts(model, b, offs)
        assert rc >= 0, (b, offs, rc)
        assert got == oracle_starts(oracle, b, offs), (b, offs)
