"""Seeded synthetic UTF-8 corpora of the shapes BASELINE.json names (configs 1-5).

Used by the parity tests and bench.py.  Everything is numpy / pure Python and deterministic in
the seed, so the oracle and the GPU path see identical bytes on every machine.
"""
from __future__ import annotations

import random
from typing import Tuple

import numpy as np

COMMON_WORDS = (
    "the of and to a in is that it was for on are as with his they at be this from have or by one had not but what "
    "all were when we there can an your which their said if do will each about how up out them then she many some so "
    "these would other into has more her two like him see time could no make than first been its who now people my "
    "made over did down only way find use may water long little very after words called just where most know"
).split()


# ------------------------------------------------------------------------------------------ config 1
def english_like(n_bytes: int = 1 << 20, seed: int = 1234) -> bytes:
    """Config 1: one English-like ASCII document."""
    r = random.Random(seed)
    parts, sz = [], 0
    while sz < n_bytes:
        w = r.choice(COMMON_WORDS)
        if r.random() < 0.10:
            w = w.capitalize()
        if r.random() < 0.03:
            w = str(r.randint(0, 99999))
        if r.random() < 0.02:
            w = w + r.choice(["'s", "'t", "'re", "'ll", "'ve", "'d", "'m"])
        x = r.random()
        sep = " " if x < 0.85 else ", " if x < 0.92 else ". " if x < 0.97 else ".\n\n"
        parts.append(w + sep)
        sz += len(w) + len(sep)
    return "".join(parts).encode("ascii")[:n_bytes]


# ------------------------------------------------------------------------------------------ config 2
def _words(r: random.Random, alphabet: str, lo: int, hi: int) -> str:
    return "".join(r.choice(alphabet) for _ in range(r.randint(lo, hi)))


_CYR = "абвгдежзийклмнопрстуфхцчшщъыьэюяАБВГДЕЖЗИКЛМНОПРСТ"
_ARA = "ابتثجحخدذرزسشصضطظعغفقكلمنهوي"
_HARAKAT = "ًٌٍَُِّْ"
_DEV_C = "कखगघचछजझटठडढणतथदधनपफबभमयरलवशषसह"
_DEV_M = "ािीुूेैोौ्ं"
_THAI = "กขคงจฉชซญดตถทธนบปผฝพฟภมยรลวศษสหอะาิีึืุูเแโใไ็่้๊๋"
_EMOJI = ["😀", "😂", "🚀", "👍", "👍🏽", "❤️", "🔥", "👩‍💻", "👨‍👩‍👧‍👦", "🎉", "✨", "🇫🇷", "🤖", "🙏", "💯"]
_PUNCT = [".", ",", "!", "?", ";", ":", " -", "...", "\"", "(", ")", "'"]


def _snippet(r: random.Random, script: str) -> str:
    """A short phrase (always ends in a separator so snippets concatenate naturally)."""
    if script == "latin":
        n = r.randint(3, 12)
        ws = []
        for _ in range(n):
            w = r.choice(COMMON_WORDS) if r.random() < 0.7 else _words(r, "abcdefghijklmnopqrstuvwxyzéèüöñçß", 2, 11)
            if r.random() < 0.12:
                w = w.capitalize()
            if r.random() < 0.03:
                w += r.choice(["'s", "'t", "'re", "'LL", "'ve", "'d", "'M", "’s"])
            ws.append(w)
        return " ".join(ws) + r.choice([". ", ", ", "! ", "? ", ".\n", "\n\n", " ", ": "])
    if script == "cyrillic":
        return " ".join(_words(r, _CYR, 2, 10) for _ in range(r.randint(3, 9))) + r.choice([". ", ", ", "\n", " — ", " "])
    if script == "cjk":
        s = "".join(chr(r.randint(0x4E00, 0x9FA5)) if r.random() < 0.8 else chr(r.randint(0x3041, 0x3096))
                    for _ in range(r.randint(4, 24)))
        return s + r.choice(["。", "，", "、", "！", "\n", " ", "？"])
    if script == "arabic":
        ws = []
        for _ in range(r.randint(3, 8)):
            w = ""
            for _ in range(r.randint(2, 7)):
                w += r.choice(_ARA)
                if r.random() < 0.3:
                    w += r.choice(_HARAKAT)
            ws.append(w)
        return " ".join(ws) + r.choice(["، ", ". ", "؟ ", "\n", " "])
    if script == "devanagari":
        ws = []
        for _ in range(r.randint(3, 8)):
            w = ""
            for _ in range(r.randint(1, 5)):
                w += r.choice(_DEV_C)
                if r.random() < 0.6:
                    w += r.choice(_DEV_M)
            ws.append(w)
        return " ".join(ws) + r.choice(["। ", ", ", "\n", " ", "? "])
    if script == "thai":
        return " ".join(_words(r, _THAI, 4, 30) for _ in range(r.randint(1, 4))) + r.choice([" ", "\n", " ๆ ", ". "])
    if script == "emoji":
        s = ""
        for _ in range(r.randint(1, 6)):
            s += r.choice(_EMOJI) if r.random() < 0.7 else (" " + r.choice(COMMON_WORDS) + " ")
        return s + r.choice([" ", "\n", "!! ", " "])
    # "code": digits, punctuation, code-like
    k = r.random()
    if k < 0.25:
        return str(r.randint(0, 10 ** r.randint(1, 12))) + r.choice([" ", ", ", ".", "\n", "; ", " + "])
    if k < 0.5:
        return "%s_%s(%d, %s) %s " % (r.choice(COMMON_WORDS), r.choice(COMMON_WORDS), r.randint(0, 999),
                                      r.choice(COMMON_WORDS), r.choice(["{", "}", ";", "=>", "//", "==", "!="]))
    if k < 0.7:
        return "    " * r.randint(0, 3) + "%s = %s[%d];\r\n" % (r.choice(COMMON_WORDS), r.choice(COMMON_WORDS), r.randint(0, 99))
    if k < 0.85:
        return "".join(r.choice(_PUNCT) for _ in range(r.randint(1, 6))) + " "
    return "%d.%02d%% $%d,%03d %s\t" % (r.randint(0, 99), r.randint(0, 99), r.randint(0, 999), r.randint(0, 999),
                                       r.choice(["\n", "\n\n", "  ", "\t\t"]))


SCRIPTS = ["latin", "cyrillic", "cjk", "arabic", "devanagari", "thai", "emoji", "code"]
SCRIPT_P = [0.50, 0.10, 0.10, 0.05, 0.05, 0.03, 0.02, 0.15]
_SNIP_MAX = 128
_SNIP_PER_DOC = 40
_DOC_MAX = 1024
_N_SNIP = 2048

_table_cache = {}


def _snippet_tables(seed: int):
    """Per script: flat byte table + offsets of _N_SNIP snippets, each <= _SNIP_MAX bytes."""
    if seed in _table_cache:
        return _table_cache[seed]
    r = random.Random(seed * 7919 + 13)
    flats, offs = [], []
    for s in SCRIPTS:
        blobs = []
        while len(blobs) < _N_SNIP:
            b = _snippet(r, s).encode("utf-8")
            while len(b) > _SNIP_MAX:                      # cut back to a char boundary
                b = b[:_SNIP_MAX]
                while b and (b[-1] & 0xC0) == 0x80:
                    b = b[:-1]
                if b and b[-1] >= 0xC0:
                    b = b[:-1]
            if b:
                blobs.append(b)
        off = np.zeros(_N_SNIP + 1, dtype=np.int64)
        np.cumsum([len(b) for b in blobs], out=off[1:])
        flats.append(np.frombuffer(b"".join(blobs), dtype=np.uint8))
        offs.append(off)
    _table_cache[seed] = (flats, offs)
    return flats, offs


def mixed_script_docs(n_docs: int, seed: int = 42, first_doc: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Config 2: n_docs mixed-script documents of at most 1 KiB (snippets of <= 128 bytes of the
    document's script, cut back to the limit; 30 % of snippets are swapped for Latin/code ones).  Document i depends only
    on (seed, first_doc + i), so shards of one corpus can be generated independently.
    Returns (bytes uint8, doc_off uint64[n_docs+1])."""
    flats, offs = _snippet_tables(seed)
    flat = np.concatenate(flats)
    base = np.zeros(len(SCRIPTS) + 1, dtype=np.int64)
    np.cumsum([len(f) for f in flats], out=base[1:])
    g_off = np.concatenate([offs[s][:-1] + base[s] for s in range(len(SCRIPTS))])       # start of every snippet in flat
    g_len = np.concatenate([np.diff(offs[s]) for s in range(len(SCRIPTS))])
    out_parts, lens_parts = [], []
    CH = 1 << 14
    lo, hi = first_doc, first_doc + n_docs
    for ci in range(lo // CH, (hi + CH - 1) // CH if n_docs else 0):
        # counter-based: one PCG stream per absolute chunk of CH documents, always drawn in full
        rng = np.random.Generator(np.random.PCG64([seed, ci, 0x5EED]))
        script = rng.choice(len(SCRIPTS), size=CH, p=SCRIPT_P)
        swap = rng.random((CH, _SNIP_PER_DOC))
        alt = rng.choice([0, 7], size=(CH, _SNIP_PER_DOC))
        pick = rng.integers(0, _N_SNIP, size=(CH, _SNIP_PER_DOC))
        a, b = max(lo, ci * CH) - ci * CH, min(hi, (ci + 1) * CH) - ci * CH
        script, swap, alt, pick = script[a:b], swap[a:b], alt[a:b], pick[a:b]
        m = b - a
        sc = np.where(swap < 0.3, alt, script[:, None])
        gid2 = sc * _N_SNIP + pick
        ln2 = g_len[gid2]
        ln2 = ln2 * (np.cumsum(ln2, axis=1) <= _DOC_MAX)          # cut the document back to <= 1 KiB
        gid = gid2.reshape(-1)
        ln = ln2.reshape(-1)
        src = g_off[gid]
        total = int(ln.sum())
        dst_start = np.zeros(len(ln) + 1, dtype=np.int64)
        np.cumsum(ln, out=dst_start[1:])
        idx = np.repeat(src - dst_start[:-1], ln) + np.arange(total, dtype=np.int64)
        out_parts.append(flat[idx])
        lens_parts.append(ln2.sum(axis=1))
    data = np.concatenate(out_parts) if out_parts else np.zeros(0, np.uint8)
    doc_off = np.zeros(n_docs + 1, dtype=np.uint64)
    if n_docs:
        np.cumsum(np.concatenate(lens_parts), out=doc_off[1:])
    return data, doc_off


# ------------------------------------------------------------------------------------------ config 3
def single_long_document(n_bytes: int, seed: int = 7, max_ws_run: int = 1 << 14, max_digit_run: int = 1 << 22,
                         align: int = 1 << 13) -> bytes:
    """Config 3: one document: English-like base with injected whitespace runs (space/tab/LF/CRLF
    mixes, <= max_ws_run bytes: each such run is ONE pre-token, and the reference's merge loop is
    quadratic in it) and digit runs (log-uniform up to max_digit_run), centred on multiples of
    `align` so that they straddle the kernels' tile boundaries; contractions at tile edges; CR-only
    and CRLF line endings; trailing whitespace at the end."""
    r = random.Random(seed)
    base = english_like(min(n_bytes, 1 << 22), seed + 1)
    reps = -(-n_bytes // len(base))
    buf = bytearray((base * reps)[:n_bytes])
    n_inj = max(8, n_bytes >> 15)
    for k in range(n_inj):
        kind = r.choice(["ws", "ws", "ws", "digit", "crlf", "crlf", "contr", "contr", "cr", "cr", "ws", "contr"])
        centre = (r.randint(1, max(1, n_bytes // align - 1))) * align
        if kind == "digit":
            ln = int(2 ** r.uniform(0, np.log2(max_digit_run)))
            ln = min(ln, n_bytes // 16)
            run = bytes(r.choice(b"0123456789") for _ in range(min(ln, 4096)))
            run = (run * (ln // len(run) + 1))[:ln]
        elif kind == "ws":
            ln = int(2 ** r.uniform(0, np.log2(max_ws_run)))
            alphabet = r.choice([b" ", b" \t", b" \n", b" \r\n", b"\n", b" \t\n\r"])
            run = bytes(r.choice(alphabet) for _ in range(ln))
        elif kind == "crlf":
            run = b"\r\n" * r.randint(1, 40)
        elif kind == "cr":
            run = b"\r" * r.randint(1, 5) + b" " * r.randint(0, 3)
        else:
            run = r.choice([b"it's", b"we'LL", b"you're", b"I'd", b"they've", b" 's", b"!'s"])
        start = max(0, min(n_bytes - len(run), centre - r.randint(0, len(run))))
        buf[start:start + len(run)] = run
    tail = b" \n \t  "
    buf[-len(tail):] = tail
    return bytes(buf)


# ------------------------------------------------------------------------------------------ config 4
def adversarial_pieces(n_pieces: int = 256, piece_bytes: int = 1 << 16, seed: int = 11) -> bytes:
    """Config 4: long single pre-tokens separated by single spaces."""
    r = random.Random(seed)
    fams = ["a", "lower", "ab", "cjk", "cjk1", "emoji", "mixed"]
    out = []
    for i in range(n_pieces):
        f = fams[i % len(fams)]
        if f == "a":
            s = "a" * piece_bytes
        elif f == "lower":
            s = "".join(r.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(piece_bytes))
        elif f == "ab":
            s = "ab" * (piece_bytes // 2)
        elif f == "cjk":
            s = "".join(chr(r.randint(0x4E00, 0x9FA5)) for _ in range(piece_bytes // 3))
        elif f == "cjk1":
            s = "中" * (piece_bytes // 3)
        elif f == "emoji":
            s = "".join(r.choice(["😀", "👩‍💻", "🔥", "👍🏽"]) for _ in range(piece_bytes // 8))
        else:
            s = "".join(r.choice("abcdeКирилلعربية中文ñü") for _ in range(piece_bytes // 2))
        out.append(s)
    return " ".join(out).encode("utf-8")


def checksum64(a: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of a uint32/uint8 array (for cross-GPU-count comparisons)."""
    a = np.ascontiguousarray(a)
    x = a.astype(np.uint64)
    idx = np.arange(1, len(x) + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = (x + np.uint64(0x9E3779B97F4A7C15)) * (idx * np.uint64(0xBF58476D1CE4E5B9) + np.uint64(1))
        h ^= h >> np.uint64(29)
        return int(np.bitwise_xor.reduce(h) ^ np.uint64(len(x))) if len(x) else 0
