"""Parity of the CUDA path (through the C ABI) with the oracle, the reference's golden vectors and
the committed engine fixtures.  Token ids must be bit-exact, decoded text byte-exact.

Every test here needs a B200 (`-m gpu`).  The gpu_tok fixture fails -- it never skips -- when the
CUDA library or the device is missing, so these tests cannot pass on a fallback."""
import base64
import ctypes
import json
import random

import numpy as np
import pytest

from tekken_rs_b200 import SpecialTokenPolicy, Tekkenizer, TokenizerError, corpus, kernel_launch_count, shard_plan, sharding

pytestmark = pytest.mark.gpu


def one_doc(raw):
    a = np.frombuffer(raw, dtype=np.uint8) if isinstance(raw, (bytes, bytearray)) else raw
    return a, np.array([0, len(a)], dtype=np.uint64)


def assert_same_batch(gpu_tok, oracle, data, off, bos, eos, n_threads=8):
    ids, toff = gpu_tok.encode_batch_np(data, off, bos, eos)
    rid, roff = oracle.encode_batch_np(data, off, bos, eos, n_threads=n_threads)
    if not np.array_equal(toff, roff):
        d = int(np.nonzero(toff != roff)[0][0]) - 1
        a, b = int(off[d]), int(off[d + 1])
        raise AssertionError("doc %d differs: %r\n gpu    %s\n oracle %s" % (
            d, bytes(data[a:b])[:120], ids[int(toff[d]):int(toff[d + 1])][:40].tolist(), rid[int(roff[d]):int(roff[d + 1])][:40].tolist()))
    if not np.array_equal(ids, rid):
        i = int(np.nonzero(ids != rid)[0][0])
        d = int(np.searchsorted(roff, i, side="right") - 1)
        a, b = int(off[d]), int(off[d + 1])
        raise AssertionError("doc %d id %d differs: %r\n gpu    %s\n oracle %s" % (
            d, i, bytes(data[a:b])[:120], ids[max(0, i - 8):i + 8].tolist(), rid[max(0, i - 8):i + 8].tolist()))
    return ids, toff


# ------------------------------------------------------------------------------------------ known answers

def test_native_library_ran(gpu_tok):
    before = kernel_launch_count()
    assert gpu_tok.device() == 0
    assert gpu_tok.encode("Hello, world!", False, False) == [22177, 1044, 4304, 1033]   # tests/test_tokenizer_output.rs:26
    assert kernel_launch_count() > before          # our kernels, not a fallback


def test_reference_golden_encode(gpu_tok, goldens):
    # tests/test_tokenizer_output.rs:22-373
    for e in goldens["encode"]:
        assert gpu_tok.encode(e["text"], e["add_bos"], e["add_eos"]) == e["ids"], e["source"]
        assert gpu_tok.decode(e["ids"], SpecialTokenPolicy.Ignore) == e["text"], e["source"]


def test_reference_golden_decode(gpu_tok, goldens):
    # tests/test_rust_tokenizer.rs:16-19,80 ; src/tekkenizer.rs:427
    for d in goldens["decode"]:
        assert gpu_tok.decode(d["ids"], d["policy"]) == d["text"], d["source"]


def test_engine_fixture_cases(gpu_tok, fixtures):
    for c in fixtures["cases"]:
        assert gpu_tok.encode(c["text"], c["add_bos"], c["add_eos"]) == c["ids"], repr(c["text"])
    texts = [c["text"] for c in fixtures["cases"]]
    got = gpu_tok.encode_batch(texts, False, False)
    for c, g in zip(fixtures["cases"], got):
        want = c["ids"][1 if c["add_bos"] else 0:len(c["ids"]) - (1 if c["add_eos"] else 0)]
        assert g == want, repr(c["text"])


def test_engine_fixture_pieces(gpu_tok, fixtures):
    # single pre-tokens isolating the merge loop: a piece made only of letters of one script, or an
    # arbitrary byte string, is not always ONE pre-token, so compare through the oracle's split
    for p in fixtures["pieces"]:
        b = base64.b64decode(p["bytes_b64"])
        try:
            b.decode("utf-8")
        except UnicodeDecodeError:
            continue
        ids = gpu_tok.encode(b, False, False)
        if len(set(b)) > 0 and all(0x61 <= x <= 0x7A for x in b):     # pure lowercase = one piece
            assert ids == [r + 1000 for r in p["ranks"]]


def test_engine_fixture_corpora(gpu_tok, fixtures):
    for c in fixtures["corpora"]:
        if c["generator"] == "mixed_script_docs":
            data, off = corpus.mixed_script_docs(**c["kwargs"])
        else:
            data, off = one_doc(getattr(corpus, c["generator"])(**c["kwargs"]))
        ids, _ = gpu_tok.encode_batch_np(data, off, True, True)
        assert len(ids) == c["n_ids"] and corpus.checksum64(ids) == c["ids_checksum"]
        assert ids[:32].tolist() == c["ids_head"] and ids[-32:].tolist() == c["ids_tail"]


def test_basic_tokenizer_example(gpu_tok):
    # examples/basic_tokenizer_test.rs:7-18
    ids = gpu_tok.encode("Hello, world! This is a test.", True, True)
    assert ids == [1, 22177, 1044, 4304, 1033, 2409, 1395, 1261, 2688, 1046, 2]
    assert gpu_tok.decode(ids, SpecialTokenPolicy.Keep) == "<s>Hello, world! This is a test.</s>"
    assert gpu_tok.decode(ids, SpecialTokenPolicy.Ignore) == "Hello, world! This is a test."
    assert gpu_tok.encode("", True, True) == [1, 2] and gpu_tok.encode("", False, False) == []
    assert gpu_tok.encode("का", False, False) == [2622, 1658]      # the hard-coded pattern, not Mistral's


def test_single_text_latency_path(gpu_tok, oracle):
    # tk_encode of a text <= 8128 bytes runs the single-block kernel (tk_small.cuh), longer texts and texts with a piece
    # > 512 bytes the batch pipeline: both must give the oracle's ids, at every length around the switch
    rng = random.Random(17)
    from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET
    texts = ["", " ", "a", "\n", "  \n  ", "x \n \t", "tail ws   ", "12345678901", "it's we'LL", "\r\n\r\n", "é" * 300, "q" * 512, "q" * 513,
             " " * 700, "ab" * 3000, "😀" * 1500, "z" * 8128, "word " * 1625, "word " * 1626, ("x" * 31 + " ") * 254]
    for L in (1, 31, 32, 33, 95, 96, 97, 511, 512, 513, 4095, 4096, 4097, 8100, 8127, 8128, 8129, 8192, 9000):
        texts.append("".join(rng.choice("abc de.\n1'") for _ in range(L)))
        t = "".join(rng.choice(FUZZ_ALPHABET) for _ in range(L))
        while len(t.encode()) > L:
            t = t[:-1]
        texts.append(t)
    for t in texts:
        for bos, eos in ((True, True), (False, False), (True, False)):
            assert gpu_tok.encode(t, bos, eos) == oracle.encode(t, bos, eos), (len(t), t[:60])
    # random short strings: single-text path == batch path == oracle
    fz = _fuzz_texts(4000, 23, [0, 1, 2, 3, 5, 8, 13, 30, 64, 200, 700, 3000])
    data, off = _pack(fz)
    ids, toff = assert_same_batch(gpu_tok, oracle, data, off, True, True)
    for k in range(0, len(fz), 7):
        assert gpu_tok.encode(fz[k], True, True) == ids[int(toff[k]):int(toff[k + 1])].tolist(), fz[k][:60]
    for bad in (b"ab\xff", b"x" * 5000 + b"\xe4\xb8", b"\x80"):
        with pytest.raises(TokenizerError) as e:
            gpu_tok.encode(bad, True, True)
        assert e.value.kind == "InvalidUtf8"


def test_single_sequence_decode_latency_path(gpu_tok, oracle):
    # tk_decode of at most 2,048 ids that decode to at most 24 KiB runs the single-block kernel (decode_small_kernel),
    # anything larger the batch pipeline: the same bytes and the same errors as the oracle on both sides of each switch
    from oracle import tekken_oracle as TO
    rng = random.Random(29)
    long_word = gpu_tok.encode("internationalization " * 3, False, False)
    cases = [[], [1], [2], [1, 2], gpu_tok.encode("Hello, world!", True, True), gpu_tok.encode("日本語 😀 é", True, True)]
    for n in (1, 7, 8, 9, 255, 256, 257, 2047, 2048, 2049, 3000):
        cases.append([rng.choice(long_word) for _ in range(n)])                      # long tokens: up to > 24 KiB of text
        cases.append(gpu_tok.encode("".join(rng.choice("abc de.\n1'é日") for _ in range(n)), True, True)[:n])
        cases.append([1000 + rng.randrange(256) for _ in range(n)])                 # raw bytes: mostly invalid UTF-8
        cases.append([rng.randrange(0, 1300) for _ in range(n)])                    # specials and bytes
    cases += [[1000 + 0xE4, 1000 + 0xB8, 1000 + 0xAD], [1000 + 0xE4, 1000 + 0xB8], [1000 + 0xE4, 2, 1000 + 0xB8, 1000 + 0xAD], [200000], [131071]]
    longest = max(long_word, key=lambda i: len(oracle.decode_bytes([i], "Ignore")))
    assert len(oracle.decode_bytes([longest], "Ignore")) >= 13
    cases += [[longest] * 2048, [longest] * 1800 + [1] * 248]                       # <= 2,048 ids but more than 24 KiB of text
    for ids in cases:
        for pol in ("Ignore", "Keep", "Raise"):
            try:
                want = ("ok", oracle.decode_bytes(ids, pol))
            except TO.TokenizerError as e:
                want = ("err", e.kind)
            try:
                got = ("ok", gpu_tok.decode_bytes(ids, pol))
            except TokenizerError as e:
                got = ("err", e.kind)
            assert got == want, (len(ids), ids[:8], pol)


def test_host_engine_chunks_slices_and_devices(gpu_tok, oracle, tekken_json):
    # the host-buffer engine (tk_api.cu): chunks cut at document boundaries, documents larger than a chunk sliced at
    # context-free piece boundaries, chunks dealt to several handles with the ids landing in document order, pageable
    # and page-locked callers.  Small chunk sizes make all of that happen on a few MB.
    import torch
    from tekken_rs_b200 import encode_batch_multi, set_chunk_bytes, set_pack_ids
    rng = random.Random(5)
    big = corpus.english_like(3 << 20, 99)                                   # sliced (many cut points)
    nocut = ("x" * 700000).encode()                                          # no cut point at all: one piece of work
    cjk = "".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(120000)).encode()   # no ASCII: no cut point
    mixed, moff = corpus.mixed_script_docs(3000, 3)
    docs = [bytes(mixed[int(moff[i]):int(moff[i + 1])]) for i in range(3000)]
    docs[10:10] = [big, b"", nocut, b"tiny", cjk, corpus.single_long_document(1 << 20, 3)]
    data, off = _pack(docs)
    want, woff = oracle.encode_batch_np(data, off, True, True, n_threads=8)
    others = [Tekkenizer.from_file(tekken_json, device=0) for _ in range(2)]
    try:
        # ... and the ids come back as uint32 (0), as an 18- or 24-bit stream widened on the host, or however the
        # library decides for a call of this size (-1)
        for chunk, pack in ((64 << 10, 18), (300 << 10, 24), (1 << 20, 0), (0, -1), (0, 18)):
            set_chunk_bytes(chunk)
            set_pack_ids(pack)
            ids, toff = gpu_tok.encode_batch_np(data, off, True, True)                       # pageable caller memory
            assert np.array_equal(toff, woff) and np.array_equal(ids, want), "chunk %d" % chunk
            ids, toff = encode_batch_multi([gpu_tok] + others, data, off, True, True)       # three pipelines, one result
            assert np.array_equal(toff, woff) and np.array_equal(ids, want), "multi, chunk %d" % chunk
            pinned = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).pin_memory()
            ids, toff = gpu_tok.encode_batch_np(pinned.numpy()[:len(data)], off, True, True)   # page-locked caller memory
            assert np.array_equal(toff, woff) and np.array_equal(ids, want), "pinned, chunk %d" % chunk
            for bos, eos in ((False, False), (True, False)):
                one, _ = gpu_tok.encode_batch_np(np.frombuffer(big, dtype=np.uint8), np.array([0, len(big)], dtype=np.uint64), bos, eos)
                assert one.tolist() == oracle.encode(big, bos, eos)
            # the decode engine on the same ids: chunks of ids, sequences larger than a chunk sliced before an id that
            # starts with an ASCII byte (the CJK document has no such id for 100 k ids), Keep and Ignore
            back, boff = gpu_tok.decode_batch_np(want, woff, SpecialTokenPolicy.Ignore)
            assert np.array_equal(boff, off) and np.array_equal(back, data), "decode, chunk %d" % chunk
            kept, koff = gpu_tok.decode_batch_np(want, woff, SpecialTokenPolicy.Keep)
            assert len(kept) == len(data) + 7 * len(docs) and bytes(kept[:3]) == b"<s>" and bytes(kept[-4:]) == b"</s>"
            assert np.array_equal(koff, off + np.arange(len(off), dtype=np.uint64) * np.uint64(7))
            broken = want.copy()
            k = int(woff[len(docs) - 5]) + 1
            broken[k], broken[k + 1] = 1000 + 0xE4, 1000 + 0x41     # a lead byte followed by 'A' in a late sequence
            with pytest.raises(TokenizerError) as e:
                gpu_tok.decode_batch_np(broken, woff, SpecialTokenPolicy.Ignore)
            assert e.value.kind == "Tokenizers" and "sequence %d" % (len(docs) - 5) in e.value.msg
            with pytest.raises(TokenizerError) as e:
                gpu_tok.decode_batch_np(want, woff, SpecialTokenPolicy.Raise)
            assert e.value.kind == "SpecialTokenPolicy" and "sequence 0)" in e.value.msg
        # a tiny estimate for the result buffer: the first chunk has few ids per byte, the rest many (forces the repeat)
        set_chunk_bytes(64 << 10)
        skew = [b" " * 60000 + b"x"] + ["".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(300)).encode() for _ in range(2000)]
        sdata, soff = _pack(skew)
        for pack in (0, 18):
            set_pack_ids(pack)
            assert_same_batch(gpu_tok, oracle, sdata, soff, True, True)
    finally:
        set_chunk_bytes(0)
        set_pack_ids(-1)
        for t in others:
            t.close()


def test_one_call_over_all_visible_gpus(gpu_tok, oracle, tekken_json):
    # tk_encode_batch_multi with one handle per device of the box (skipped on a one-GPU box: the same code path runs
    # there with several handles on device 0, test_host_engine_chunks_slices_and_devices)
    import torch
    from tekken_rs_b200 import encode_batch_multi, set_chunk_bytes
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("one GPU visible")
    handles = [gpu_tok] + [Tekkenizer.from_file(tekken_json, device=g) for g in range(1, n_gpu)]
    try:
        data, off = corpus.mixed_script_docs(60000, 9)
        want, woff = oracle.encode_batch_np(data, off, True, True, n_threads=8)
        for chunk in (1 << 20, 0):
            set_chunk_bytes(chunk)
            ids, toff = encode_batch_multi(handles, data, off, True, True)
            assert np.array_equal(toff, woff) and np.array_equal(ids, want)
    finally:
        set_chunk_bytes(0)
        for t in handles[1:]:
            t.close()


def test_encode_file_streams_shards(gpu_tok, oracle, tmp_path, monkeypatch):
    # text file in, u32 id shard + u64 offsets out, window by window (1 MiB windows here), raw and .npy
    import os
    from tekken_rs_b200 import encode_file
    monkeypatch.setenv("TEKKEN_B200_FILE_WINDOW_MB", "1")
    data, off = corpus.mixed_script_docs(6000, 11)
    lines = [bytes(data[int(off[i]):int(off[i + 1])]).replace(b"\n", b" ").replace(b"\r", b" ") + b"\n" for i in range(6000)]
    lines.append(b"last line without a newline")
    text = b"".join(lines)
    src = tmp_path / "corpus.txt"
    src.write_bytes(text)
    ldata, loff = _pack(lines)
    want, woff = oracle.encode_batch_np(ldata, loff, True, True, n_threads=8)
    st = encode_file([gpu_tok], str(src), str(tmp_path / "ids.bin"), str(tmp_path / "ids.idx"), delimiter=10)
    assert st["n_docs"] == len(lines) and st["n_bytes"] == len(text) and st["n_tokens"] == len(want)
    assert np.array_equal(np.fromfile(tmp_path / "ids.bin", dtype=np.uint32), want)
    assert np.array_equal(np.fromfile(tmp_path / "ids.idx", dtype=np.uint64), woff)
    st = encode_file([gpu_tok], str(src), str(tmp_path / "ids.npy"), str(tmp_path / "off.npy"), delimiter=10, npy=True)
    assert np.array_equal(np.load(tmp_path / "ids.npy"), want) and np.array_equal(np.load(tmp_path / "off.npy"), woff)
    # the whole file as one document, no offsets file; an empty file
    st = encode_file([gpu_tok], str(src), str(tmp_path / "one.bin"), None, delimiter=None, add_bos=False, add_eos=False)
    assert np.fromfile(tmp_path / "one.bin", dtype=np.uint32).tolist() == oracle.encode(text, False, False) and st["n_docs"] == 1
    (tmp_path / "empty.txt").write_bytes(b"")
    st = encode_file([gpu_tok], str(tmp_path / "empty.txt"), str(tmp_path / "e.bin"), str(tmp_path / "e.idx"))
    assert st["n_tokens"] == 0 and os.path.getsize(tmp_path / "e.bin") == 0 and np.fromfile(tmp_path / "e.idx", dtype=np.uint64).tolist() == [0]
    with pytest.raises(TokenizerError) as e:
        encode_file([gpu_tok], str(tmp_path / "missing.txt"), str(tmp_path / "x.bin"))
    assert e.value.kind == "Io"


# ------------------------------------------------------------------------------------------ differential fuzz

def _fuzz_texts(n, seed, lengths):
    from oracle.tools.make_golden_fixtures import FUZZ_ALPHABET
    rng = random.Random(seed)
    return ["".join(rng.choice(FUZZ_ALPHABET) for _ in range(rng.choice(lengths))).encode() for _ in range(n)]


def _pack(texts):
    off = np.zeros(len(texts) + 1, dtype=np.uint64)
    np.cumsum([len(t) for t in texts], out=off[1:])
    return np.frombuffer(b"".join(texts), dtype=np.uint8), off


@pytest.mark.parametrize("bos,eos", [(False, False), (True, False), (False, True), (True, True)])
def test_fuzz_batch(gpu_tok, oracle, bos, eos):
    data, off = _pack(_fuzz_texts(30000, 5, [0, 1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100, 200]))
    assert_same_batch(gpu_tok, oracle, data, off, bos, eos)


def test_fuzz_as_one_document(gpu_tok, oracle):
    # the same bytes as ONE text: document boundaries disappear, runs join across them
    data, _ = _pack(_fuzz_texts(30000, 6, [1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100, 200]))
    assert_same_batch(gpu_tok, oracle, data, np.array([0, len(data)], dtype=np.uint64), False, True)


def test_fuzz_window_and_tile_edges(gpu_tok, oracle):
    # every interesting construct slid across the 32-byte window and 8 KiB tile boundaries
    cons = ["it's", "we'LL", "a'ſb", "  \n\n  x", "!!!\r\n\r\n", "12345678", " \t\n ", "中文字", "😀😀", " !x", "é́", "\r \n"]
    docs = []
    for c in cons:
        cb = c.encode()
        for edge in (32, 64, 8192, 16384):
            for shift in range(-len(cb) - 1, 3):
                pad = edge + shift
                docs.append(b"ab " * (pad // 3) + b"x" * (pad % 3) + cb + b" tail")   # construct starts at byte `pad`
    data, off = _pack(docs)
    assert_same_batch(gpu_tok, oracle, data, off, True, True)
    assert_same_batch(gpu_tok, oracle, data, np.array([0, len(data)], dtype=np.uint64), False, False)


def test_every_unicode_scalar_value(gpu_tok, oracle):
    # the UTF-8 decoder and the two-stage class table of the CUDA path against the oracle for EVERY scalar value:
    # each one between letters, before a digit, doubled, and after a space
    cps = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF]
    docs = [("x%sy1 %s%s" % (chr(c), chr(c), chr(c))).encode("utf-8") for c in cps]
    data, off = _pack(docs)
    assert_same_batch(gpu_tok, oracle, data, off, False, False)
    # and as running text (classes of neighbouring characters interact across the 32-byte windows)
    rng = random.Random(4)
    text = "".join(chr(rng.choice(cps)) if rng.random() < 0.5 else rng.choice("ab 1\n'.") for _ in range(400000)).encode("utf-8")
    assert_same_batch(gpu_tok, oracle, *one_doc(text), False, False, n_threads=1)


def test_ragged_and_empty_documents(gpu_tok, oracle):
    texts = [b"", b"", b"a", b"", " ".encode(), b"", "日本".encode(), b"", b""]
    data, off = _pack(texts)
    for bos, eos in ((True, True), (False, False), (True, False)):
        ids, toff = assert_same_batch(gpu_tok, oracle, data, off, bos, eos)
    assert gpu_tok.encode_batch([], True, True) == []
    assert gpu_tok.encode_batch(["", "", ""], True, True) == [[1, 2]] * 3
    assert gpu_tok.encode_batch(["", ""], False, False) == [[], []]
    ids, toff = gpu_tok.encode_batch_np(np.zeros(0, np.uint8), np.zeros(1, np.uint64), True, True)
    assert len(ids) == 0 and toff.tolist() == [0]


def test_many_tiny_and_empty_documents(gpu_tok, oracle):
    # thousands of document starts inside one 4 KiB tile: more BOS/EOS than the emit kernel stages in shared
    # memory (its slow path), several documents per 32-byte window, runs of empty documents
    rng = random.Random(21)
    texts = []
    for _ in range(30000):
        r = rng.random()
        texts.append(b"" if r < 0.5 else rng.choice([b"a", b"ab", b" x", "é".encode(), b"12", b"\n", b"hi there", b"!"]) if r < 0.97
                     else ("word " * rng.randint(1, 300)).encode())
    texts += [b""] * 5000 + [b"tail"] + [b""] * 3000
    data, off = _pack(texts)
    for bos, eos in ((True, True), (False, True), (False, False)):
        ids, toff = assert_same_batch(gpu_tok, oracle, data, off, bos, eos)
    raw, boff = gpu_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert np.array_equal(boff, off)


def test_document_starts_at_every_window_position(gpu_tok, oracle):
    # the emit kernel works on 32-byte windows (one lane per position), 512-byte warp ranges and 4 KiB tiles: put
    # document starts -- single ones, several per window, empty documents, starts next to pieces longer than
    # 96 bytes -- at every position of those units, with every BOS/EOS combination
    rng = random.Random(77)
    texts = []
    for rep in range(3):
        for gap in list(range(0, 70)) + [95, 96, 97, 127, 128, 129, 480, 511, 512, 513, 4000, 4095, 4096, 4097]:
            filler = "".join(rng.choice("abc de.\n") for _ in range(gap)).encode()
            texts.append(filler)
            kind = rng.randrange(5)
            if kind == 0:
                texts += [b"", b""]                                   # empty documents at this position
            elif kind == 1:
                texts += [b"x", b"", b"yz"]                           # several starts in one window
            elif kind == 2:
                texts.append(("q" * rng.randint(97, 300)).encode())   # a long piece right at a document start
            elif kind == 3:
                texts.append(("é" * rng.randint(20, 60)).encode())    # a merge-class piece right at a document start
    data, off = _pack(texts)
    for bos, eos in ((True, True), (True, False), (False, True), (False, False)):
        ids, toff = assert_same_batch(gpu_tok, oracle, data, off, bos, eos)
    raw, boff = gpu_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert np.array_equal(boff, off) and raw.tobytes() == data.tobytes()


# ------------------------------------------------------------------------------------------ the five configs

def test_config1_english_1mib(gpu_tok, oracle):
    # BASELINE config 1: encode(text, true, true) then decode(Keep) (examples/basic_tokenizer_test.rs:7-18)
    raw = corpus.english_like(1 << 20)
    data, off = one_doc(raw)
    ids, _ = assert_same_batch(gpu_tok, oracle, data, off, True, True)
    assert gpu_tok.decode_bytes(ids, SpecialTokenPolicy.Keep) == b"<s>" + raw + b"</s>"
    assert gpu_tok.decode_bytes(ids, SpecialTokenPolicy.Ignore) == raw


def test_config2_mixed_script_docs(gpu_tok, oracle):
    data, off = corpus.mixed_script_docs(100000, 42)
    ids, toff = assert_same_batch(gpu_tok, oracle, data, off, True, True)
    raw, boff = gpu_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert np.array_equal(raw, data) and np.array_equal(boff, off)


def test_config2_full_size_roundtrip_and_sample(gpu_tok, oracle):
    # 1,000,000 documents: encode -> decode round trip over all of them, shard invariance, and the
    # oracle on every 16th chunk of 16,384 documents
    n_docs = 1_000_000
    data, off = corpus.mixed_script_docs(n_docs, 42)
    ids, toff = gpu_tok.encode_batch_np(data, off, True, True)
    assert len(toff) == n_docs + 1 and int(toff[-1]) == len(ids)
    raw, boff = gpu_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert np.array_equal(boff, off) and np.array_equal(raw, data)
    assert np.all(ids[toff[:-1].astype(np.int64)] == 1) and np.all(ids[toff[1:].astype(np.int64) - 1] == 2)
    for c in range(0, n_docs >> 14, 16):
        a, b = c << 14, min(n_docs, (c + 1) << 14)
        rid, roff = oracle.encode_batch_np(data[int(off[a]):int(off[b])], off[a:b + 1] - off[a], True, True, n_threads=8)
        assert np.array_equal(rid, ids[int(toff[a]):int(toff[b])])
        assert np.array_equal(roff + toff[a], toff[a:b + 1])
    # shard invariance (the multi-GPU path on one device): 8 byte-balanced shards, stitched
    plan = shard_plan(off, 8)
    parts, counts, offs = [], [], []
    for s in range(8):
        b, e = int(plan[s]), int(plan[s + 1])
        loc_off = sharding.rebase_offsets(off, b, e)
        sid, stoff = gpu_tok.encode_batch_np(data[int(off[b]):int(off[e])], loc_off, True, True)
        parts.append(sid); counts.append(len(sid)); offs.append(stoff)
    assert np.array_equal(np.concatenate(parts), ids)
    stitched = np.concatenate([sharding.stitch_token_offsets(offs[s], counts, s)[:-1] for s in range(8)] + [toff[-1:]])
    assert np.array_equal(stitched, toff)


def test_config3_single_long_document(gpu_tok, oracle):
    # cross-tile digit / whitespace / CR-LF runs; 64 MiB against the oracle
    raw = corpus.single_long_document(1 << 26)
    data, off = one_doc(raw)
    ids, _ = assert_same_batch(gpu_tok, oracle, data, off, False, False, n_threads=1)
    assert gpu_tok.decode_bytes(ids, SpecialTokenPolicy.Ignore) == raw


def test_config3_full_size_1gib(gpu_tok, oracle):
    raw = corpus.single_long_document(1 << 30)
    data, off = one_doc(raw)
    ids, toff = gpu_tok.encode_batch_np(data, off, True, True)
    out, boff = gpu_tok.decode_batch_np(ids, toff, SpecialTokenPolicy.Ignore)
    assert int(boff[-1]) == len(raw) and np.array_equal(out, data)
    del out
    # the oracle is linear on this text except inside the injected whitespace runs: check it all
    rid, _ = oracle.encode_batch_np(data, off, True, True)
    assert np.array_equal(rid, ids)


def test_long_runs_across_many_tiles(gpu_tok, oracle):
    # digit and whitespace runs spanning hundreds of 8 KiB tiles, with every alignment mod 3
    for lead in (0, 1, 2, 31, 8191):
        for run in (b"7" * 1_000_003, b" " * 300_001, b"\n" * 100_000, b" \t" * 70_001, b"\r\n" * 50_001, b" " * 99_999 + b"\n" + b" " * 99_999):
            raw = b"a" * lead + run + b"z" + run[:4097] + b"!" + run[:9000]
            data, off = one_doc(raw)
            assert_same_batch(gpu_tok, oracle, data, off, False, False, n_threads=1)


def test_config4_adversarial_long_pieces(gpu_tok, oracle):
    for n_pieces, size in ((14, 1 << 9), (14, 1 << 12), (14, 1 << 14), (7, 1 << 16)):
        data, off = one_doc(corpus.adversarial_pieces(n_pieces, size))
        ids, _ = assert_same_batch(gpu_tok, oracle, data, off, False, False, n_threads=1)
        assert gpu_tok.decode_bytes(ids) == data.tobytes()


def test_config4_full_256_pieces(gpu_tok, oracle):
    # BASELINE configs[3] at its named size: 256 single pre-tokens x 64 KiB (16 MiB, one document); the oracle's
    # heap variant of the merge loop (cross-checked against the literal loop in tests/test_oracle.py) takes seconds
    data, off = one_doc(corpus.adversarial_pieces(256, 1 << 16))
    ids, _ = assert_same_batch(gpu_tok, oracle, data, off, False, False, n_threads=1)
    assert gpu_tok.decode_bytes(ids) == data.tobytes()


def test_config5_sharded_roundtrip_reduced_scale(gpu_tok, oracle):
    # BASELINE configs[4] at reduced scale: 8 shards x 65,536 documents of the 64-shard corpus' generator.  Every shard:
    # decode(encode(x)) == x compared ON THE DEVICE; ids identical whether the corpus goes through as 1 shard plan or 8
    # (checksum of checksums + concatenation); oracle on every 16th document.
    import torch
    n_shards, per = 8, 1 << 16
    datas, offs = [], []
    for s in range(n_shards):
        d, o = corpus.mixed_script_docs(per, 42, first_doc=s * (1 << 20))     # the first 65,536 documents of bench.py's shard s
        datas.append(d); offs.append(o)
    st = torch.cuda.current_stream().cuda_stream
    shard_ids, shard_chk = [], 0
    for s in range(n_shards):
        d, o = datas[s], offs[s]
        n, nd = len(d), per
        d_data = torch.from_numpy(np.concatenate([d, np.zeros(64, np.uint8)])).cuda()
        d_off = torch.from_numpy(o.astype(np.int64)).cuda()
        cap = n + 2 * nd + 2
        d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
        d_toff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
        ntok = gpu_tok.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), nd, n, True, True, d_tok.data_ptr(), cap, d_toff.data_ptr(), st)
        d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        d_boff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
        nb = gpu_tok.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), nd, ntok, SpecialTokenPolicy.Ignore, d_out.data_ptr(), n + 64,
                                         d_boff.data_ptr(), 0, st)
        assert nb == n and bool(torch.equal(d_out[:n], d_data[:n])) and bool(torch.equal(d_boff, d_off)), "shard %d round trip" % s
        ids = d_tok[:ntok].cpu().numpy().view(np.uint32)
        toff = d_toff.cpu().numpy().view(np.uint64)
        shard_ids.append(ids)
        shard_chk ^= corpus.checksum64(ids) + 0x1000003 * s
        sel = np.arange(0, nd, 16)
        sub = [d[int(o[k]):int(o[k + 1])] for k in sel]
        soff = np.zeros(len(sub) + 1, dtype=np.uint64)
        np.cumsum([len(x) for x in sub], out=soff[1:])
        rid, _ = oracle.encode_batch_np(np.concatenate(sub), soff, True, True, n_threads=8)
        assert np.array_equal(np.concatenate([ids[int(toff[k]):int(toff[k + 1])] for k in sel]), rid), "shard %d oracle sample" % s
    # the same corpus as ONE batch cut by a 1-shard and an 8-shard byte-balanced plan through the host API
    data = np.concatenate(datas)
    off = np.zeros(n_shards * per + 1, dtype=np.uint64)
    np.cumsum(np.concatenate([np.diff(o.astype(np.int64)) for o in offs]), out=off[1:])
    whole, wtoff = gpu_tok.encode_batch_np(data, off, True, True)
    assert np.array_equal(whole, np.concatenate(shard_ids))
    plan = shard_plan(off, 8)
    parts, chk8 = [], 0
    for s in range(8):
        b, e = int(plan[s]), int(plan[s + 1])
        sid, _ = gpu_tok.encode_batch_np(data[int(off[b]):int(off[e])], sharding.rebase_offsets(off, b, e), True, True)
        parts.append(sid)
    assert corpus.checksum64(np.concatenate(parts)) == corpus.checksum64(whole)
    chk1 = 0
    for s in range(n_shards):
        a, b = int(wtoff[s * per]), int(wtoff[(s + 1) * per])
        chk1 ^= corpus.checksum64(whole[a:b]) + 0x1000003 * s
    assert chk1 == shard_chk


def test_bad_offsets_fail_cleanly_on_the_device_path(gpu_tok):
    # offsets that do not cover the text must come back as InvalidArgument, never as a device fault --
    # also when the text ends in a long run of punctuation / whitespace (no natural piece start at the end)
    import torch
    raw = (b"word " * 50 + b"." * 5000 + b" " * 3000)
    d_data = torch.from_numpy(np.frombuffer(raw + b"\0" * 64, dtype=np.uint8).copy()).cuda()
    for bad in ([0, len(raw) - 7], [5, len(raw)], [0, 900, 100, len(raw)], [0, len(raw) + 9]):
        d_off = torch.tensor(bad, dtype=torch.int64, device="cuda")
        d_tok = torch.empty(len(raw) + 16, dtype=torch.int32, device="cuda")
        d_toff = torch.empty(len(bad), dtype=torch.int64, device="cuda")
        with pytest.raises(TokenizerError) as e:
            gpu_tok.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), len(bad) - 1, len(raw), True, True, d_tok.data_ptr(),
                                        len(raw) + 16, d_toff.data_ptr(), 0)
        assert e.value.kind == "InvalidArgument"
    torch.cuda.synchronize()
    assert gpu_tok.encode("still alive", False, False) == gpu_tok.encode(b"still alive", False, False)
    data = np.frombuffer(raw, dtype=np.uint8)
    for bad in ([0, 900, 100, len(raw)], [0, 10**12, 5, len(raw)], [3, len(raw)]):
        with pytest.raises(TokenizerError) as e:
            gpu_tok.encode_batch_np(data, np.array(bad, dtype=np.uint64), False, False)
        assert e.value.kind == "InvalidArgument"


def test_config4_piece_length_sweep(gpu_tok, oracle):
    # every length across the lane-merge classes (one lane per piece <= 96 B, class limits 4, 8, 12, 16, 24, 32, 48,
    # 64, 96) and around the later escalation thresholds (warp <= 512 B, block beyond)
    rng = random.Random(3)
    docs = []
    for L in list(range(1, 132)) + [255, 256, 257, 500, 511, 512, 513, 514, 600, 1023, 1024, 1025, 3000, 5000]:
        docs.append("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(L)).encode())
        docs.append(("é" * L).encode()[:L - (L % 2)] or b"e")
        docs.append(("." * L).encode())
        docs.append((" " * L).encode())
        docs.append(("\n" * L).encode())
    data, off = _pack(docs)
    assert_same_batch(gpu_tok, oracle, data, off, False, False)
    one = b" ".join(docs)
    assert_same_batch(gpu_tok, oracle, *one_doc(one), False, False)


# ------------------------------------------------------------------------------------------ decode

def test_decode_policies_and_grouping(gpu_tok, oracle):
    # tests/test_tekken.rs:53-86, tests/test_tokenizer_detailed.rs:140-180, 326-370
    ids = gpu_tok.encode("Hello world", True, True)
    assert gpu_tok.decode(ids, SpecialTokenPolicy.Keep) == "<s>Hello world</s>"
    assert gpu_tok.decode(ids, SpecialTokenPolicy.Ignore) == "Hello world"
    with pytest.raises(TokenizerError) as e:
        gpu_tok.decode(ids, SpecialTokenPolicy.Raise)
    assert e.value.kind == "SpecialTokenPolicy"
    assert gpu_tok.decode(ids[1:-1], SpecialTokenPolicy.Raise) == "Hello world"
    mixed = [1, 3] + gpu_tok.encode("ab cd", False, False) + [4, 4] + gpu_tok.encode("ef", False, False) + [2]
    for pol in ("Keep", "Ignore"):
        assert gpu_tok.decode_all(mixed, pol) == oracle.decode_all(mixed, pol)
        assert gpu_tok.decode(mixed, pol) == oracle.decode(mixed, pol)
    assert gpu_tok.decode_all(mixed, "Keep") == ["<s>", "[INST]", "ab cd", "[/INST]", "[/INST]", "ef", "</s>"]
    assert gpu_tok.decode([], "Keep") == "" and gpu_tok.decode_all([], "Keep") == []     # :470-521
    assert gpu_tok.decode([500], "Keep") == "<SPECIAL_500>"


def test_decode_errors_match_oracle(gpu_tok, oracle):
    from oracle import tekken_oracle as TO
    e4 = 1000 + 0xE4
    cases = [([131072], "Ignore"), ([200000], "Keep"), ([e4], "Ignore"), ([e4, 1, 1000 + 0xB8, 1000 + 0xAD], "Ignore"),
             ([e4, 1000 + 0xB8, 1000 + 0xAD], "Ignore"), ([1000 + 0x80], "Keep"), ([22177, 1000 + 0xC3], "Keep"),
             ([1, 22177], "Raise"), ([e4, 1], "Raise"), ([1, e4], "Raise"), ([1000 + 0xC3, 1000 + 0xA9], "Raise"),
             ([1000 + 0xF0, 1000 + 0x9F, 1000 + 0x98, 1000 + 0x80], "Keep"), ([1000 + 0xF0, 1000 + 0x9F, 1000 + 0x98], "Keep"),
             ([1000 + 0xED, 1000 + 0xA0, 1000 + 0x80], "Keep"), ([1000 + 0xC0, 1000 + 0x80], "Keep"), ([131071], "Keep")]
    for ids, pol in cases:
        try:
            want = ("ok", oracle.decode_bytes(ids, pol))
        except TO.TokenizerError as e:
            want = ("err", e.kind)
        try:
            got = ("ok", gpu_tok.decode_bytes(ids, pol))
        except TokenizerError as e:
            got = ("err", e.kind)
        assert got == want, (ids, pol)


def test_decode_random_ids_match_oracle(gpu_tok, oracle):
    from oracle import tekken_oracle as TO
    rng = np.random.default_rng(12)
    seqs = []
    for _ in range(3000):
        n = int(rng.integers(0, 40))
        kind = rng.integers(0, 4)
        if kind == 0:
            s = rng.integers(1000, 131072, size=n)            # arbitrary ordinary ids: often invalid UTF-8 as a run
        elif kind == 1:
            s = rng.integers(0, 1300, size=n)                 # specials and raw bytes
        elif kind == 2:
            s = rng.integers(1256, 40000, size=n)
        else:
            s = rng.integers(130000, 132000, size=n)          # around the end of the vocabulary
        seqs.append(s.astype(np.uint32))
    for pol in ("Ignore", "Keep", "Raise"):
        want = []
        for s in seqs:
            try:
                want.append(("ok", oracle.decode_bytes(s, pol)))
            except TO.TokenizerError as e:
                want.append(("err", e.kind))
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        flat = np.concatenate(seqs)
        first_bad = next((i for i, w in enumerate(want) if w[0] == "err"), None)
        # batch call: Result<Vec<_>> semantics -- the error of the first failing sequence
        try:
            gpu_tok.decode_batch_np(flat, off, pol)
            assert first_bad is None
        except TokenizerError as e:
            assert first_bad is not None and e.kind == want[first_bad][1]
        # per-sequence
        for s, w in list(zip(seqs, want))[:600]:
            try:
                got = ("ok", gpu_tok.decode_bytes(s, pol))
            except TokenizerError as e:
                got = ("err", e.kind)
            assert got == w, (s.tolist(), pol)
        # the valid ones as one batch
        good = [s for s, w in zip(seqs, want) if w[0] == "ok"]
        goff = np.zeros(len(good) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in good], out=goff[1:])
        raw, boff = gpu_tok.decode_batch_np(np.concatenate(good) if good else np.zeros(0, np.uint32), goff, pol)
        b = raw.tobytes()
        k = 0
        for s, w in zip(seqs, want):
            if w[0] == "ok":
                assert b[int(boff[k]):int(boff[k + 1])] == w[1]
                k += 1


# ------------------------------------------------------------------------------------------ errors and custom vocabularies

def test_invalid_utf8_is_rejected(gpu_tok):
    for bad in (b"abc\xff", b"\x80", b"\xc3", b"ab\xe4\xb8", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xc0\xaf", b"x" * 100 + b"\xe4\xb8" + b"y" * 100):
        with pytest.raises(TokenizerError) as e:
            gpu_tok.encode(bad, False, False)
        assert e.value.kind == "InvalidUtf8", bad
    # a character may not straddle a document boundary
    data = np.frombuffer("日本".encode(), dtype=np.uint8)
    with pytest.raises(TokenizerError) as e:
        gpu_tok.encode_batch_np(data, np.array([0, 2, 6], dtype=np.uint64), False, False)
    assert e.value.kind == "InvalidUtf8"
    with pytest.raises(TokenizerError) as e:
        gpu_tok.encode_batch_np(data, np.array([0, 7, 6], dtype=np.uint64), False, False)
    assert e.value.kind == "InvalidArgument"


def _mini_vocab():
    # tests/test_small_vocab.rs:11-67
    vocab = [{"rank": i, "token_bytes": base64.b64encode(bytes([i])).decode()} for i in range(256)]
    vocab.append({"rank": 256, "token_bytes": base64.b64encode(b"hello").decode()})
    vocab.append({"rank": 257, "token_bytes": base64.b64encode(b"world").decode()})
    return vocab


def test_small_vocab_whole_piece_shortcut():
    # "hello" is reachable only through the whole-piece lookup (no merge path leads to it)
    sp = [{"rank": i, "token_str": s, "is_control": True} for i, s in enumerate(["<unk>", "<s>", "</s>"])]
    t = Tekkenizer.new(_mini_vocab(), sp, "ignored", 268, 10, "v7", device=0)
    assert t.encode("hello", False, False) == [266]
    assert t.encode("hello world", True, True) == [1, 266] + [10 + b for b in b" world"] + [2]
    assert t.encode("helloworld", False, False) == [10 + b for b in b"helloworld"]
    assert t.decode([266, 10 + 32, 267], "Keep") == "hello world"
    with pytest.raises(TokenizerError) as e:
        t.decode([268], "Keep")
    assert e.value.kind == "Tokenizers"


def test_missing_bos_eos_is_token_not_found():
    # src/tekkenizer.rs:335-340 via :394-402
    sp = [{"rank": 0, "token_str": "<unk>", "is_control": True}]
    t = Tekkenizer.new(_mini_vocab(), sp, "", 268, 10, "v7", device=0)
    assert t.encode("hi", False, False) == [10 + ord("h"), 10 + ord("i")]
    for bos, eos in ((True, False), (False, True)):
        with pytest.raises(TokenizerError) as e:
            t.encode("hi", bos, eos)
        assert e.value.kind == "TokenNotFound"


def test_synthetic_vocab_non_monotone_ranks(oracle):
    # merge order on a vocabulary whose ranks are shuffled (rank(merged) < rank(part) is common):
    # the oracle's literal loop is the reference definition
    from oracle import tekken_oracle as TO
    rng = random.Random(8)
    toks = set()
    while len(toks) < 600:
        toks.add(bytes(rng.choice(b"abcd") for _ in range(rng.randint(2, 5))))
    toks = list(toks)
    rng.shuffle(toks)
    vocab = [{"rank": i, "token_bytes": base64.b64encode(bytes([i])).decode()} for i in range(256)]
    vocab += [{"rank": 256 + i, "token_bytes": base64.b64encode(t).decode()} for i, t in enumerate(toks)]
    sp = [{"rank": i, "token_str": s, "is_control": True} for i, s in enumerate(["<unk>", "<s>", "</s>"])]
    n = len(vocab)
    t = Tekkenizer.new(vocab, sp, "", n + 3, 3, "v3", device=0)
    o = TO.OracleTekkenizer(vocab, sp, "", n + 3, 3, "v3")
    docs = ["".join(rng.choice("abcd") for _ in range(rng.choice([3, 9, 40, 64, 65, 200, 600, 2000]))).encode() for _ in range(400)]
    data, off = _pack(docs)
    assert_same_batch(t, o, data, off, True, True)


# ------------------------------------------------------------------------------------------ device-pointer ABI

def test_device_pointer_entry_points(gpu_tok, oracle):
    import torch
    data, off = corpus.mixed_script_docs(4096, 7)
    d_data = torch.from_numpy(data.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    cap = len(data) + 2 * 4096 + 2
    d_tok = torch.empty(cap, dtype=torch.int32, device="cuda")
    d_toff = torch.empty(4097, dtype=torch.int64, device="cuda")
    n = gpu_tok.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), 4096, len(data), True, True, d_tok.data_ptr(), cap,
                                    d_toff.data_ptr(), torch.cuda.current_stream().cuda_stream)
    rid, roff = oracle.encode_batch_np(data, off, True, True)
    assert n == len(rid)
    assert np.array_equal(d_tok[:n].cpu().numpy().view(np.uint32), rid)
    assert np.array_equal(d_toff.cpu().numpy().view(np.uint64), roff)
    # too small an output buffer is reported, with the size that would have been enough
    with pytest.raises(TokenizerError) as e:
        gpu_tok.encode_batch_device(d_data.data_ptr(), d_off.data_ptr(), 4096, len(data), True, True, d_tok.data_ptr(), 1000,
                                    d_toff.data_ptr(), 0)
    assert e.value.kind == "BufferTooSmall" and str(n) in e.value.msg
    # decode on the device
    d_out = torch.empty(len(data) + 64, dtype=torch.uint8, device="cuda")
    d_boff = torch.empty(4097, dtype=torch.int64, device="cuda")
    d_st = torch.empty(4096, dtype=torch.int32, device="cuda")
    nb = gpu_tok.decode_batch_device(d_tok.data_ptr(), d_toff.data_ptr(), 4096, n, SpecialTokenPolicy.Ignore, d_out.data_ptr(),
                                     len(data) + 64, d_boff.data_ptr(), d_st.data_ptr(), 0)
    assert nb == len(data) and np.array_equal(d_out[:nb].cpu().numpy(), data)
    assert np.array_equal(d_boff.cpu().numpy().view(np.uint64), off) and int(d_st.abs().sum()) == 0


def test_concurrent_host_threads(gpu_tok, oracle):
    # the handle is usable from many host threads at once (&self in the reference)
    import threading
    texts = ["thread %d says: %s" % (i, "héllo wörld 123 " * (i % 7 + 1)) for i in range(64)]
    want = [oracle.encode(t, True, True) for t in texts]
    got = [None] * len(texts)

    def work(k):
        for i in range(k, len(texts), 8):
            got[i] = gpu_tok.encode(texts[i], True, True)
    th = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert got == want


# ------------------------------------------------------------------------------------------ decode kernels: every path of tk_decode.cu

def _vocab_bytes(tekken_json):
    cfg = json.load(open(tekken_json))
    ns, vs = cfg["config"]["default_num_special_tokens"], cfg["config"]["default_vocab_size"]
    return ns, [base64.b64decode(v["token_bytes"]) for v in cfg["vocab"][:vs - ns]]


def _is_utf8(b):
    try:
        b.decode("utf-8")
        return True
    except UnicodeDecodeError:
        return False


def test_decode_token_lengths_and_tile_shapes(gpu_tok, oracle, tekken_json):
    # The gather reads a token from the first half of its table cell (<= 7 bytes), from both halves (8..15) or from the
    # byte table (>= 16); a tile whose text exceeds the staging buffer (> 7 bytes per id over 2,048 ids) takes the
    # direct path; the copy-out shifts by the phase of the tile's first byte.  Drive every combination, with sequence
    # starts at every position of a thread's group of ids and outputs that are not word aligned.
    import torch
    ns, voc = _vocab_bytes(tekken_json)
    by_len = {}
    for r, b in enumerate(voc):
        if _is_utf8(b):
            by_len.setdefault(len(b), []).append(ns + r)
    pick = lambda L, k=0: by_len[L][k % len(by_len[L])]
    long_len = max(L for L in by_len if L >= 16)
    rng = np.random.default_rng(3)
    seqs = []
    seqs.append([pick(long_len)] * 5000)                                         # > 7 B/id: whole tiles on the direct path
    seqs.append([pick(16, i) for i in range(4100)])                              # byte-table tokens, tiles on the direct path
    seqs.append([x for i in range(3000) for x in (pick(16, i), pick(1, i), pick(2, i), pick(1, i + 1))])   # ... inside staged tiles
    for L in (7, 8, 11, 12, 15):
        seqs.append([x for i in range(2500) for x in (pick(L, i), pick(1, i), pick(3, i))])                 # cell halves
        seqs.append([pick(L, i) for i in range(2100)])
    lens = sorted(by_len)
    for n in list(range(0, 20)) + [2047, 2048, 2049, 4096, 5000]:                # ragged: starts at every phase of a tile
        seqs.append([pick(int(rng.choice(lens[:24])), int(rng.integers(0, 1000))) for _ in range(n)])
    seqs.append([1] + [pick(5, 3)] * 10 + [4, 3, 2])                             # specials between ordinary runs
    flat = np.concatenate([np.asarray(s, dtype=np.uint32) for s in seqs])
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    sp = {1: b"<s>", 2: b"</s>", 3: b"[INST]", 4: b"[/INST]"}
    for pol in ("Ignore", "Keep"):
        want = [b"".join((sp[i] if pol == "Keep" else b"") if i < ns else voc[i - ns] for i in s) for s in seqs]
        raw, boff = gpu_tok.decode_batch_np(flat, off, pol)
        b = raw.tobytes()
        assert [b[int(boff[i]):int(boff[i + 1])] for i in range(len(seqs))] == want, pol
    # device pointers, output buffer at every byte phase
    want = b"".join(voc[i - ns] for s in seqs for i in s if i >= ns)
    d_ids = torch.from_numpy(flat.view(np.int32)).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_boff = torch.empty(len(seqs) + 1, dtype=torch.int64, device="cuda")
    buf = torch.zeros(len(want) + 80, dtype=torch.uint8, device="cuda")
    for phase in range(0, 5):
        buf.fill_(0xEE)
        nb = gpu_tok.decode_batch_device(d_ids.data_ptr(), d_off.data_ptr(), len(seqs), len(flat), "Ignore", buf.data_ptr() + phase,
                                         len(want), d_boff.data_ptr(), 0, 0)
        got = buf.cpu().numpy()
        assert nb == len(want) and got[phase:phase + nb].tobytes() == want, phase
        assert (got[:phase] == 0xEE).all() and (got[phase + nb:] == 0xEE).all(), phase   # nothing outside the caller's range


def test_decode_utf8_verdicts_at_every_boundary(gpu_tok, oracle):
    # Strict UTF-8 per ordinary run (tk_decode.cu, validate): a character that is complete, cut off by a special id, by
    # the end of its sequence, by a following ASCII byte, or missing its lead, slid over every position of a 32-byte
    # window, a 4-byte word and a 2,048-id tile; plus the ranges a lead byte restricts its second byte to.
    import torch
    from oracle import tekken_oracle as TO
    B = lambda *bs: [1000 + b for b in bs]
    tails = [B(0xE4, 0xB8, 0xAD), B(0xE4, 0xB8), B(0xE4), B(0xE4, 0xB8) + [2] + B(0xAD), B(0xE4, 0xB8, 0x41), B(0xB8, 0xAD),
             B(0xF0, 0x9F, 0x98, 0x80), B(0xF0, 0x9F, 0x98), B(0xF0, 0x9F) + [3] + B(0x98, 0x80), B(0xC3, 0xA9), B(0xC3), B(0xC3, 0x41),
             B(0xE0, 0x80, 0x80), B(0xE0, 0xA0, 0x80), B(0xED, 0x9F, 0xBF), B(0xED, 0xA0, 0x80), B(0xF0, 0x8F, 0x80, 0x80),
             B(0xF4, 0x8F, 0xBF, 0xBF), B(0xF4, 0x90, 0x80, 0x80), B(0xC1, 0x81), B(0xF5, 0x80, 0x80, 0x80), B(0xFF)]
    seqs = []
    for k in list(range(0, 70)) + [2040, 2045, 2046, 2047, 2048, 2049, 4095]:
        for tail in tails:
            seqs.append(B(*([0x61] * k)) + tail)
            seqs.append(B(*([0x61] * k)) + tail + B(0x62, 0x63))
    flat = np.concatenate([np.asarray(s, dtype=np.uint32) for s in seqs])
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    codes = {"Tokenizers": -4, "SpecialTokenPolicy": -8}
    d_ids = torch.from_numpy(flat.view(np.int32)).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_boff = torch.empty(len(seqs) + 1, dtype=torch.int64, device="cuda")
    d_st = torch.empty(len(seqs), dtype=torch.int32, device="cuda")
    d_out = torch.empty(len(flat) * 8 + 64, dtype=torch.uint8, device="cuda")
    for pol in ("Ignore", "Keep", "Raise"):
        want = []
        for s in seqs:
            try:
                oracle.decode_bytes(s, pol)
                want.append(0)
            except TO.TokenizerError as e:
                want.append(codes[e.kind])
        try:
            gpu_tok.decode_batch_device(d_ids.data_ptr(), d_off.data_ptr(), len(seqs), len(flat), pol, d_out.data_ptr(), len(flat) * 8 + 64,
                                        d_boff.data_ptr(), d_st.data_ptr(), 0)
            assert not any(want)
        except TokenizerError as e:
            first = next(i for i, w in enumerate(want) if w)
            assert "sequence %d" % first in e.msg, (pol, e.msg, first)
        got = d_st.cpu().numpy().tolist()
        bad = [i for i in range(len(seqs)) if got[i] != want[i]]
        assert not bad, (pol, bad[:5], [(seqs[i][-6:], got[i], want[i]) for i in bad[:5]])
