"""The CPU oracle against (a) the reference's own golden vectors, (b) fixtures generated from the
upstream engine, (c) the engine itself when importable.  No GPU."""
import base64
import random

import numpy as np
import pytest

from oracle import tekken_oracle as TO
from tekken_rs_b200 import corpus


def test_reference_golden_encode(oracle, goldens):
    assert len(goldens["encode"]) == 20
    for e in goldens["encode"]:
        assert oracle.encode(e["text"], e["add_bos"], e["add_eos"]) == e["ids"], e["source"]
        assert oracle.decode(e["ids"], "Ignore") == e["text"], e["source"]   # the tests also round-trip


def test_reference_golden_decode(oracle, goldens):
    for d in goldens["decode"]:
        assert oracle.decode(d["ids"], d["policy"]) == d["text"], d["source"]


def test_sizes(oracle):
    # tests/test_tekken.rs:45-46
    assert oracle.vocab_size() == 131072
    assert oracle.num_special_tokens() == 1000
    assert oracle.bos_id() == 1 and oracle.eos_id() == 2


def test_engine_fixture_cases(oracle, fixtures):
    for c in fixtures["cases"]:
        assert oracle.encode(c["text"], c["add_bos"], c["add_eos"]) == c["ids"], repr(c["text"])


def test_engine_fixture_pieces(oracle, fixtures):
    for p in fixtures["pieces"]:
        b = base64.b64decode(p["bytes_b64"])
        for mode in (1, 2):   # literal restatement and the heap form
            assert oracle.encode_piece(b, mode) == p["ranks"]


def test_engine_fixture_corpora(oracle, fixtures):
    for c in fixtures["corpora"]:
        if c["generator"] == "mixed_script_docs":
            data, off = corpus.mixed_script_docs(**c["kwargs"])
        else:
            raw = getattr(corpus, c["generator"])(**c["kwargs"])
            data, off = np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64)
        assert len(data) == c["n_bytes"] and corpus.checksum64(data) == c["bytes_checksum"], "generator drifted"
        ids, _ = oracle.encode_batch_np(data, off, True, True)
        assert len(ids) == c["n_ids"]
        assert corpus.checksum64(ids) == c["ids_checksum"]
        assert ids[:32].tolist() == c["ids_head"] and ids[-32:].tolist() == c["ids_tail"]


def test_live_engine_fuzz(oracle):
    tiktoken = pytest.importorskip("tiktoken")
    del tiktoken
    from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET
    enc = TO.tiktoken_engine(oracle.ranks)
    rng = random.Random(99)
    texts = ["".join(rng.choice(FUZZ_ALPHABET) for _ in range(rng.choice([1, 2, 4, 9, 17, 40, 90]))) for _ in range(8000)]
    ref = enc.encode_ordinary_batch(texts, num_threads=4)
    for t, r in zip(texts, ref):
        assert oracle.encode(t, False, False) == [x + 1000 for x in r], repr(t)


def test_live_engine_every_unicode_scalar_value(oracle):
    # pins the oracle's Unicode class tables (and its UTF-8 handling) to the engine for EVERY scalar value: each one
    # between letters, before a digit, doubled after a space, and followed by a line feed
    tiktoken = pytest.importorskip("tiktoken")
    del tiktoken
    enc = TO.tiktoken_engine(oracle.ranks)
    cps = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF]
    texts = ["".join("x%sy1 %s%s\n" % (chr(c), chr(c), chr(c)) for c in cps[i:i + 64]) for i in range(0, len(cps), 64)]
    ref = enc.encode_ordinary_batch(texts, num_threads=8)
    for t, r in zip(texts, ref):
        assert oracle.encode(t, False, False) == [x + 1000 for x in r], repr(t[:40])


# ---- groundwork for SURVEY 8(f) rank 1: the pattern stored in tekken.json (the reference ignores it) ----

def test_config_pattern_fixtures(oracle):
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "config_pattern_fixtures.json")
    fx = json.load(open(path))
    for c in fx["cases"]:
        assert oracle.encode_config(c["text"], False, False) == c["ids"], repr(c["text"])
        assert b"".join(oracle.split_config(c["text"])) == c["text"].encode("utf-8")


def test_config_pattern_live_engine_fuzz(oracle, tekken_json):
    tiktoken = pytest.importorskip("tiktoken")
    import json
    from oracle.tools.make_config_pattern_fixtures import ALPHABET
    pattern = json.load(open(tekken_json))["config"]["pattern"]
    enc = tiktoken.Encoding("tekken-config-pattern", pat_str=pattern,
                            mergeable_ranks={b: r for r, b in enumerate(oracle.ranks)}, special_tokens={})
    rng = random.Random(5)
    texts = ["".join(rng.choice(ALPHABET) for _ in range(rng.choice([1, 2, 3, 5, 9, 17, 40, 90]))) for _ in range(8000)]
    ref = enc.encode_ordinary_batch(texts, num_threads=4)
    for t, r in zip(texts, ref):
        assert oracle.encode_config(t, False, False) == [x + 1000 for x in r], repr(t)


def test_config_pattern_every_unicode_scalar_value(oracle, tekken_json):
    # pins the sub-class tables (Lu|Lt, Ll, Lm|Lo, M) to the engine for EVERY scalar value: each one after an upper-case
    # and before a lower-case letter, the other way round, alone after a space, after a digit and before a slash
    tiktoken = pytest.importorskip("tiktoken")
    import json
    pattern = json.load(open(tekken_json))["config"]["pattern"]
    enc = tiktoken.Encoding("tekken-config-pattern", pat_str=pattern,
                            mergeable_ranks={b: r for r, b in enumerate(oracle.ranks)}, special_tokens={})
    cps = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF]
    texts = ["".join("A%sb %sB %s 1%s/\n" % (chr(c), chr(c), chr(c), chr(c)) for c in cps[i:i + 64]) for i in range(0, len(cps), 64)]
    ref = enc.encode_ordinary_batch(texts, num_threads=8)
    for t, r in zip(texts, ref):
        assert oracle.encode_config(t, False, False) == [x + 1000 for x in r], repr(t[:40])


def test_config_pattern_run_predicate(oracle):
    # the run-based "is a piece start" predicate a GPU split for the stored pattern would use (research prototype,
    # oracle/research/): no mismatch against the sequential restatement
    from oracle.research import config_pattern_positionwise as pw
    assert pw.check(20000, seed=7) == 0
    assert pw.check(2000, seed=8, lengths=(60, 200)) == 0
    assert pw.check_safe(5000) == 0          # the purely local "safe start" rules only ever mark real starts


def test_config_pattern_differs_from_reference_pattern(oracle):
    # the hard-coded pattern splits the Devanagari matra off, the stored pattern keeps it
    assert oracle.encode("\u0915\u093e", False, False) == [2622, 1658]
    assert oracle.encode_config("\u0915\u093e", False, False) == [15729]
    assert [p.decode() for p in oracle.split_config("HelloWORLD helloWorld 123")] == ["Hello", "WORLD", " hello", "World", " ", "1", "2", "3"]


def test_split_examples(oracle):
    # SURVEY.md section 3.2 consequences
    def sp(s):
        return [p.decode() for p in oracle.split(s)]
    assert sp("!!word") == ["!!", "word"] and sp("!word") == ["!word"]
    assert sp(" !b") == [" !", "b"]
    assert sp("1234567") == ["123", "456", "7"]
    assert sp("x!!!\n\ny") == ["x", "!!!\n\n", "y"]
    assert sp("a \n\n  b") == ["a", " \n\n", " ", " b"]
    assert sp("a'ſb") == ["a", "'ſ", "b"]
    assert sp("x'sy") == ["x", "'s", "y"]


def test_heap_matches_literal_on_long_pieces(oracle):
    rng = random.Random(3)
    for n in (700, 3000):
        for kind in ("lower", "a", "ab", "cjk"):
            s = {"lower": "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(n)), "a": "a" * n,
                 "ab": "ab" * (n // 2), "cjk": "".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(n // 3))}[kind]
            b = s.encode()
            assert oracle.encode_piece(b, 1) == oracle.encode_piece(b, 2)


def test_decode_policies_and_errors(oracle):
    ids = oracle.encode("Hello, world", True, True)
    assert ids == [1, 22177, 1044, 4304, 2]                      # src/tekkenizer.rs:427
    assert oracle.decode(ids, "Keep") == "<s>Hello, world</s>"
    assert oracle.decode(ids, "Ignore") == "Hello, world"
    assert oracle.decode_all(ids, "Keep") == ["<s>", "Hello, world", "</s>"]
    assert oracle.decode_all([1, 1, 22177, 2], "Keep") == ["<s>", "<s>", "Hello", "</s>"]
    with pytest.raises(TO.TokenizerError) as e:
        oracle.decode(ids, "Raise")
    assert e.value.kind == "SpecialTokenPolicy"
    with pytest.raises(TO.TokenizerError) as e:
        oracle.decode([131072 + 5], "Ignore")
    assert e.value.kind == "Tokenizers"
    # a char split across a special id: each ordinary run must be valid on its own
    with pytest.raises(TO.TokenizerError):
        oracle.decode([1000 + 0xE4, 1, 1000 + 0xB8, 1000 + 0xAD], "Ignore")
    assert oracle.decode([1000 + 0xE4, 1000 + 0xB8, 1000 + 0xAD], "Ignore") == "中"
    assert oracle.decode([], "Keep") == ""


def test_load_validations(tekken_json):
    import json
    md = json.load(open(tekken_json, encoding="utf-8"))
    vocab = md["vocab"][:300]
    sp = [{"rank": 0, "token_str": "<unk>", "is_control": True}, {"rank": 1, "token_str": "<s>", "is_control": True},
          {"rank": 2, "token_str": "</s>", "is_control": True}]
    ok = TO.OracleTekkenizer(vocab, sp, "ignored", 310, 10, "v7")
    assert ok.encode("ab", True, True)[0] == 1 and ok.special_tokens[5]["token_str"] == "<SPECIAL_5>"
    with pytest.raises(TO.TokenizerError) as e:
        TO.OracleTekkenizer(vocab, sp, "", 311, 10, "v7")         # :80-87
    assert e.value.kind == "InvalidConfig"
    with pytest.raises(TO.TokenizerError):
        TO.OracleTekkenizer(vocab, sp + [sp[0]], "", 310, 10, "v7")   # duplicate :90-98
    with pytest.raises(TO.TokenizerError):
        TO.OracleTekkenizer(vocab, sp, "", 302, 2, "v7")          # len(special) > num_special :100-106
    bad = [dict(v) for v in vocab]
    bad[65]["token_bytes"] = base64.b64encode(b"B").decode()     # rank 65 must be byte 65
    with pytest.raises(TO.TokenizerError):
        TO.OracleTekkenizer(bad, sp, "", 310, 10, "v7")
    bad = [dict(v) for v in vocab]
    bad[299]["rank"] = 400                                       # non-contiguous :806-813
    with pytest.raises(TO.TokenizerError):
        TO.OracleTekkenizer(bad, sp, "", 310, 10, "v7")
