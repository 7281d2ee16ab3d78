"""GPU parity of the TK_SPLIT_CONFIG split (the pattern STORED in tekken.json) with the oracle's
restatement of that pattern (`oracle.encode_config` / encode mode bit 4), which tests/test_oracle.py pins to the
upstream engine (656 committed fixtures, live fuzz, every scalar value)."""
import json
import os
import random

import numpy as np
import pytest

from tekken_rs_b200 import SpecialTokenPolicy, SplitMode, Tekkenizer, TokenizerError, corpus

pytestmark = pytest.mark.gpu
CFG = 4      # oracle encode mode bit: split with the stored pattern


@pytest.fixture(scope="module")
def cfg_tok(tekken_json):
    tk = Tekkenizer.from_file(tekken_json, device=0, split=SplitMode.Config)
    assert tk.split_mode() == SplitMode.Config
    return tk


def _pack(texts):
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in texts], out=off[1:])
    return np.frombuffer(b"".join(texts), dtype=np.uint8), off


def same(cfg_tok, oracle, data, off, bos, eos, n_threads=8):
    ids, toff = cfg_tok.encode_batch_np(data, off, bos, eos)
    rid, roff = oracle.encode_batch_np(data, off, bos, eos, mode=CFG, n_threads=n_threads)
    if not (np.array_equal(toff, roff) and np.array_equal(ids, rid)):
        bad = np.nonzero(toff != roff)[0]
        d = int(bad[0]) - 1 if len(bad) else int(np.searchsorted(roff, np.nonzero(ids != rid)[0][0], side="right") - 1)
        a, b = int(off[d]), int(off[d + 1])
        raise AssertionError("doc %d differs: %r\n gpu    %s\n oracle %s" % (
            d, bytes(data[a:b])[:160], ids[int(toff[d]):int(toff[d]) + 40].tolist(), rid[int(roff[d]):int(roff[d]) + 40].tolist()))
    return ids, toff


def test_committed_fixtures(cfg_tok, oracle):
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config_pattern_fixtures.json")))
    for c in fx["cases"]:
        assert cfg_tok.encode(c["text"], False, False) == c["ids"], repr(c["text"])
    data, off = _pack([c["text"].encode("utf-8") for c in fx["cases"]])
    same(cfg_tok, oracle, data, off, True, True)


def test_differs_from_reference_pattern_where_it_should(cfg_tok, gpu_tok):
    # the stored pattern keeps a Devanagari consonant and its matra together
    assert cfg_tok.encode("का", False, False) == [15729]
    assert gpu_tok.encode("का", False, False) == [2622, 1658]
    assert cfg_tok.decode(cfg_tok.encode("HelloWorld 12345 it's", True, True), SpecialTokenPolicy.Ignore) == "HelloWorld 12345 it's"


def test_fuzz(cfg_tok, oracle):
    from oracle.tools.make_config_pattern_fixtures import ALPHABET
    rng = random.Random(31)
    texts = ["".join(rng.choice(ALPHABET) for _ in range(rng.choice([0, 1, 2, 3, 5, 9, 17, 33, 40, 64, 90, 200]))).encode("utf-8") for _ in range(30000)]
    data, off = _pack(texts)
    for bos, eos in ((False, False), (True, True)):
        same(cfg_tok, oracle, data, off, bos, eos)
    same(cfg_tok, oracle, data, np.array([0, len(data)], dtype=np.uint64), False, True, n_threads=1)     # one document: runs join


def test_every_unicode_scalar_value(cfg_tok, oracle):
    cps = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF]
    docs = [("A%sb %sB %s 1%s/" % (chr(c), chr(c), chr(c), chr(c))).encode("utf-8") for c in cps]
    data, off = _pack(docs)
    same(cfg_tok, oracle, data, off, False, False)
    rng = random.Random(4)
    text = "".join(chr(rng.choice(cps)) if rng.random() < 0.5 else rng.choice("aB 1\n/.") for _ in range(300000)).encode("utf-8")
    same(cfg_tok, oracle, np.frombuffer(text, dtype=np.uint8), np.array([0, len(text)], dtype=np.uint64), False, False, n_threads=1)


def test_window_and_tile_edges(cfg_tok, oracle):
    cons = ["HelloWorld", "ABCʰDE", ".́a", "..́a", " .́a", ".\n/.́", "12345678", "  \n\n  x", "!!!\r\n//\r\n", " \t\n ", "中文字", "😀😀", "XMLHttpRequest", "é́É́"]
    docs = []
    for c in cons:
        cb = c.encode()
        for edge in (32, 64, 8192, 16384):
            for shift in range(-len(cb) - 1, 3):
                pad = edge + shift
                docs.append(b"ab " * (pad // 3) + b"x" * (pad % 3) + cb + b" tail")
    data, off = _pack(docs)
    same(cfg_tok, oracle, data, off, True, True)
    same(cfg_tok, oracle, data, np.array([0, len(data)], dtype=np.uint64), False, False, n_threads=1)


def test_configs_1_and_2(cfg_tok, oracle):
    raw = corpus.english_like(1 << 20)
    ids, _ = same(cfg_tok, oracle, np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64), True, True, n_threads=1)
    assert cfg_tok.decode_bytes(ids, SpecialTokenPolicy.Ignore) == raw
    data, off = corpus.mixed_script_docs(100000, 42)
    ids, toff = same(cfg_tok, oracle, data, off, True, True)
    back, boff = cfg_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert np.array_equal(back, data) and np.array_equal(boff, off)


def test_long_runs_and_documents(cfg_tok, oracle):
    # walks as long as the distance between safe starts: camelCase of 64 KiB, a 200 kB whitespace run, digit runs,
    # ragged / empty documents
    rng = random.Random(9)
    camel = "".join(rng.choice("ABCDEFGH") + "".join(rng.choice("abcdefgh") for _ in range(rng.randint(0, 6))) for _ in range(16000))
    docs = [camel.encode(), b" " * 200001 + b"x", b"7" * 100003, b"", b"a", b"\n" * 50000 + b" z", ("é" * 40000).encode(), b"", b"/" * 3000 + b"\r\n" * 100]
    data, off = _pack(docs)
    same(cfg_tok, oracle, data, off, True, True)
    same(cfg_tok, oracle, data, np.array([0, len(data)], dtype=np.uint64), False, False, n_threads=1)


def test_invalid_utf8_and_wrong_pattern(cfg_tok, tekken_json):
    with pytest.raises(TokenizerError) as e:
        cfg_tok.encode(b"abc\xff def", False, False)
    assert e.value.kind == "InvalidUtf8"
    import base64
    vocab = [{"rank": i, "token_bytes": base64.b64encode(bytes([i])).decode()} for i in range(256)]
    with pytest.raises(TokenizerError) as e:
        Tekkenizer.new(vocab, [], "\\w+", 260, 4, "v7", device=0, split=SplitMode.Config)
    assert e.value.kind == "InvalidConfig"
