"""C-ABI library on a machine without a GPU: it loads, exports every symbol the header declares,
the host-side logic (loader validations, accessors, single-token helpers, shard plan) behaves like
the reference, and the compute entry points refuse to run instead of falling back to a CPU path."""
import base64
import ctypes
import json
import os
import re

import numpy as np
import pytest

from tekken_rs_b200 import SpecialTokenPolicy, Tekkenizer, TokenizerError, TokenizerVersion, _lib, shard_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tekken_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libtekken_b200.so does not export " + n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for " + n
    assert sorted(_lib.PROTOTYPES) == names


def test_library_is_the_in_tree_cuda_build():
    p = _lib.library_path()
    assert os.path.dirname(p) == os.path.join(ROOT, "tekken_rs_b200")
    blob = open(p, "rb").read()
    assert b"sm_100a" in blob or b"sm_100" in blob    # device code for B200 is embedded


def test_accessors(host_tok):
    t = host_tok
    assert t.vocab_size() == 131072 and t.num_special_tokens() == 1000          # tests/test_tekken.rs:45
    assert t.version() == TokenizerVersion.V7 and t.version().as_str() == "v7"  # :46
    assert (t.bos_id(), t.eos_id(), t.unk_id(), t.pad_id()) == (1, 2, 0, 11)
    assert t.get_control_token("[INST]") == 3 and t.get_control_token("[AUDIO]") == 24
    with pytest.raises(TokenizerError) as e:                                    # tests/test_tokenizer_detailed.rs:470-521
        t.get_control_token("<nope>")
    assert e.value.kind == "TokenNotFound" and "<nope>" in e.value.msg and "Available special tokens" in e.value.msg
    assert t.is_special_token(999) and not t.is_special_token(1000)
    assert t.is_byte(1000) and t.is_byte(1255) and not t.is_byte(1256) and not t.is_byte(5)
    assert t.vocab_piece(1) == "<s>" and t.vocab_piece(500) == "<SPECIAL_500>" and t.vocab_piece(22177) == "Hello"


def test_single_token_helpers(host_tok, oracle):
    t = host_tok
    assert t.id_to_piece(22177) == "Hello" and t.id_to_piece(2) == "</s>"
    with pytest.raises(TokenizerError) as e:
        t.id_to_piece(131072)
    assert e.value.kind == "InvalidConfig"
    with pytest.raises(TokenizerError) as e:      # byte tokens >= 0x80 are not UTF-8 on their own (:617-628)
        t.id_to_piece(1000 + 0x80)
    assert e.value.kind == "Tokenizers"
    assert t.id_to_byte_piece(22177) == b"Hello"
    assert t.id_to_byte_piece(1000 + 0x80) == b"\xef\xbf\xbd"   # lossy fallback quirk (:683-687)
    assert t.id_to_byte_piece(1, SpecialTokenPolicy.Keep) == b"<s>"
    assert t.id_to_byte_piece(1, SpecialTokenPolicy.Ignore) == b""
    with pytest.raises(TokenizerError) as e:
        t.id_to_byte_piece(1, SpecialTokenPolicy.Raise)
    assert e.value.kind == "SpecialTokenPolicy"
    # same answers as decoding one id with the oracle, over a sample of the vocabulary
    for i in list(range(990, 1300)) + list(range(5000, 131072, 997)):
        try:
            want = oracle.decode([i], "Keep")
        except Exception:
            want = None
        try:
            got = t.id_to_piece(i)
        except TokenizerError:
            got = None
        assert got == want, i


def test_compute_refuses_without_device(host_tok):
    for call in (lambda: host_tok.encode("hi", False, False), lambda: host_tok.decode([22177]),
                 lambda: host_tok.encode_batch(["a", "b"], True, True)):
        with pytest.raises(TokenizerError) as e:
            call()
        assert e.value.kind == "Cuda" and "no CPU fallback" in e.value.msg


def _mini_vocab():
    # tests/test_small_vocab.rs:11-67: 256 bytes + "hello", "world"; 10 specials
    vocab = [{"rank": i, "token_bytes": base64.b64encode(bytes([i])).decode(), "token_str": None} for i in range(256)]
    vocab.append({"rank": 256, "token_bytes": base64.b64encode(b"hello").decode(), "token_str": "hello"})
    vocab.append({"rank": 257, "token_bytes": base64.b64encode(b"world").decode(), "token_str": "world"})
    sp = [{"rank": i, "token_str": s, "is_control": True} for i, s in enumerate(["<unk>", "<s>", "</s>"])]
    return vocab, sp


def test_new_and_validations():
    vocab, sp = _mini_vocab()
    t = Tekkenizer.new(vocab, sp, r"ignored pattern", 268, 10, "v7", device=-1)
    assert t.vocab_size() == 268 and t.num_special_tokens() == 10 and t.vocab_piece(5) == "<SPECIAL_5>"
    assert t.vocab_piece(10 + 256) == "hello"

    def err(**kw):
        a = dict(vocab=vocab, special_tokens=sp, pattern="", vocab_size=268, num_special_tokens=10, version="v7", device=-1)
        a.update(kw)
        with pytest.raises(TokenizerError) as e:
            Tekkenizer.new(**a)
        return e.value
    assert err(vocab_size=269).kind == "InvalidConfig"                       # src/tekkenizer.rs:80-87
    assert "Duplicate special token" in err(special_tokens=sp + [sp[1]]).msg  # :90-98
    assert err(num_special_tokens=2, vocab_size=260).kind == "InvalidConfig"  # :100-106
    assert err(version="v9").kind == "InvalidConfig"                         # :226-232
    bad = [dict(v) for v in vocab]
    bad[65]["token_bytes"] = base64.b64encode(b"B").decode()
    assert "Expected byte token at rank 65" in err(vocab=bad).msg            # :793-798
    bad = [dict(v) for v in vocab]
    bad[257]["rank"] = 300
    assert "not contiguous" in err(vocab=bad).msg                            # :809-813
    bad = [dict(v) for v in vocab]
    bad[256]["token_bytes"] = "aGVsbG8"                                       # missing padding
    assert err(vocab=bad).kind == "Base64"                                   # :789
    bad[256]["token_bytes"] = "aGV$bG8="
    assert err(vocab=bad).kind == "Base64"


def test_from_file_errors(tmp_path, tekken_json):
    with pytest.raises(TokenizerError) as e:
        Tekkenizer.from_file(str(tmp_path / "missing.json"), device=-1)
    assert e.value.kind == "Io"
    p = tmp_path / "bad.json"
    p.write_text("{ not json")
    with pytest.raises(TokenizerError) as e:
        Tekkenizer.from_file(str(p), device=-1)
    assert e.value.kind == "Json"
    vocab, sp = _mini_vocab()
    cfg = {"pattern": "x", "num_vocab_tokens": 258, "default_vocab_size": 268, "default_num_special_tokens": 10, "version": "v7"}
    p.write_text(json.dumps({"vocab": vocab, "config": {k: v for k, v in cfg.items() if k != "version"}}))
    with pytest.raises(TokenizerError) as e:
        Tekkenizer.from_file(str(p), device=-1)
    assert e.value.kind == "Json" and "version" in e.value.msg
    p.write_text(json.dumps({"vocab": vocab, "config": dict(cfg, version="v2")}))
    with pytest.raises(TokenizerError) as e:
        Tekkenizer.from_file(str(p), device=-1)
    assert e.value.kind == "InvalidConfig" and "Unknown version: v2" in e.value.msg
    # no special_tokens -> the 20 built-ins (:234-237); unknown keys ignored; escapes decoded
    with pytest.raises(TokenizerError) as e:      # 20 built-ins do not fit 10 special slots (:100-106)
        p.write_text(json.dumps({"vocab": vocab, "config": cfg}))
        Tekkenizer.from_file(str(p), device=-1)
    assert "special_tokens.len() (20)" in e.value.msg
    cfg30 = dict(cfg, default_vocab_size=288, default_num_special_tokens=30)
    p.write_text(json.dumps({"vocab": vocab, "config": cfg30, "image": {"x": [1, 2, {"y": None}]}, "special_tokens": None}))
    t = Tekkenizer.from_file(str(p), device=-1)
    assert t.get_control_token("[TOOL_CONTENT]") == 19 and t.bos_id() == 1 and t.vocab_piece(25) == "<SPECIAL_25>"
    assert t.vocab_piece(30 + 0x41) == "A" and t.vocab_piece(30 + 0xFF) == "\ufffd"


def test_shard_plan_balances_bytes():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 2000, size=10000)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    for n in (1, 2, 3, 4, 8):
        plan = shard_plan(off, n)
        assert plan[0] == 0 and plan[-1] == len(lens) and np.all(np.diff(plan.astype(np.int64)) >= 0)
        sizes = [int(off[int(plan[i + 1])] - off[int(plan[i])]) for i in range(n)]
        assert max(sizes) - min(sizes) <= 2 * 2000
    # degenerate: fewer documents than shards, empty documents
    off = np.array([0, 0, 10, 10], dtype=np.uint64)
    plan = shard_plan(off, 8)
    assert plan[0] == 0 and plan[-1] == 3 and np.all(np.diff(plan.astype(np.int64)) >= 0)


def test_status_names():
    lib = _lib.load()
    assert lib.tk_status_name(0) == b"Ok" and lib.tk_status_name(-4) == b"Tokenizers" and lib.tk_status_name(-20) == b"InvalidUtf8"
    assert isinstance(lib.tk_kernel_launch_count(), int)
    p = ctypes.POINTER(_lib.SpecialEntry)()
    n = lib.tk_deprecated_special_tokens(ctypes.byref(p))
    assert n == 20 and p[1].token_str == b"<s>" and p[19].token_str == b"[TOOL_CONTENT]"


def test_audio_token_counting(host_tok, tekken_json):
    # AudioEncoder::encode's token count (src/audio.rs:555-591) -- host arithmetic, no GPU involved
    import json
    import random
    from oracle import tekken_oracle as TO
    from tekken_rs_b200 import audio_token_count
    cfg = json.load(open(tekken_json))["audio"]
    assert host_tok.has_audio_support() and host_tok.audio_config() == cfg
    # by hand: 16 kHz, 12.5 frames/s, hop 160, chunks of 30 s = 480,000 samples -> 3,000 frames / 8 per token = 375
    assert audio_token_count(cfg, 16000) == (480000, 375)
    assert audio_token_count(cfg, 480001) == (960000, 750)
    assert audio_token_count(cfg, 0) == (0, 0)
    toks = host_tok.encode_audio_tokens(16000)
    assert toks[0] == host_tok.get_control_token("[BEGIN_AUDIO]") and toks[1:] == [host_tok.get_control_token("[AUDIO]")] * 375
    # no chunking: short clips are padded to the window; lengths that are not a multiple of the hop lose a frame
    nochunk = dict(cfg, chunk_length_s=None)
    assert audio_token_count(nochunk, 100) == (400, 1)          # 400 / 160 = 2.5 -> ceil(1.5) = 2 frames -> ceil(2 / 8) = 1
    assert audio_token_count(nochunk, 16000) == (16000, 13)     # 100 frames -> ceil(12.5)
    assert audio_token_count(nochunk, 16001) == (16001, 13)     # ceil(100.006 - 1) = 100 frames
    rng = random.Random(3)
    for _ in range(3000):
        c = {"sampling_rate": rng.choice([8000, 16000, 22050, 44100, 48000]), "frame_rate": rng.choice([12.5, 25.0, 50.0, 7.5, 100.0]),
             "chunk_length_s": rng.choice([None, None, 30.0, 0.5, 10.25]),
             "audio_encoding_config": {"num_mel_bins": 128, "hop_length": rng.choice([128, 160, 200, 256]), "window_size": rng.choice([400, 512, 1024])}}
        if int(c["sampling_rate"] / c["frame_rate"] / c["audio_encoding_config"]["hop_length"]) == 0:
            continue
        n = rng.choice([0, 1, 399, 400, 401, rng.randrange(1, 10**7)])
        assert audio_token_count(c, n) == TO.audio_token_count(c, n), (c, n)
    # a file without an `audio` block: Audio error, as src/tekkenizer.rs:731-734; with one but without the tokens: TokenNotFound (:158-170)
    src = json.load(open(tekken_json))
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        a = dict(src); a.pop("audio")
        p = os.path.join(d, "noaudio.json"); json.dump(a, open(p, "w"))
        t = Tekkenizer.from_file(p, device=-1)
        assert not t.has_audio_support() and t.audio_config() is None
        with pytest.raises(TokenizerError) as e:
            t.encode_audio_tokens(16000)
        assert e.value.kind == "Audio"
        b = dict(src); b["special_tokens"] = [x for x in src["special_tokens"] if x["token_str"] != "[AUDIO]"]
        p = os.path.join(d, "notoken.json"); json.dump(b, open(p, "w"))
        with pytest.raises(TokenizerError) as e:
            Tekkenizer.from_file(p, device=-1)
        assert e.value.kind == "TokenNotFound"


def _pack_ids_reference(ids: np.ndarray, bits: int) -> np.ndarray:
    """The layout tkk::pack_ids writes: a little-endian bit stream, `bits` bits per id, padded to groups of 16 ids."""
    n = len(ids)
    groups = (n + 15) // 16
    padded = np.zeros(groups * 16, dtype=np.uint64)
    padded[:n] = ids
    bit_matrix = ((padded[:, None] >> np.arange(bits, dtype=np.uint64)[None, :]) & 1).astype(np.uint8)   # id x bit, LSB first
    return np.packbits(bit_matrix.reshape(-1), bitorder="little")


@pytest.mark.parametrize("bits", [18, 24])
def test_packed_id_stream_is_widened_exactly(bits):
    """Host half of the packed download: every length (vector body, ragged head and tail), every destination
    alignment, several threads' worth of ids."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for n in [0, 1, 3, 7, 8, 15, 16, 17, 31, 33, 1000, (1 << 18) - 1, (1 << 18) + 5, 3 * (1 << 18) + 77]:
        ids = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
        if n:
            ids[0] = (1 << bits) - 1
            ids[-1] = (1 << bits) - 1
        stream = np.concatenate([_pack_ids_reference(ids, bits), np.full(64, 0xAB, dtype=np.uint8)])    # readable slack, not zero
        for shift in (0, 1, 5):                                  # destination not 32-byte aligned
            out = np.full(n + 16, 0xDEADBEEF, dtype=np.uint32)
            dst = out[shift:shift + n]
            assert lib.tk_debug_unpack_ids(stream.ctypes.data, n, bits, dst.ctypes.data) == 0
            assert np.array_equal(dst, ids), (bits, n, shift)
            assert (out[:shift] == 0xDEADBEEF).all() and (out[shift + n:] == 0xDEADBEEF).all()
    assert lib.tk_debug_unpack_ids(None, 1, 17, None) != 0
