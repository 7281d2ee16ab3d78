"""Host-side mirror of the reference's ``Tekkenizer`` API over the C ABI (include/tekken_b200.h).

Same method names, argument meaning and error behaviour as ``struct Tekkenizer``
(src/tekkenizer.rs:34-44; methods :71-760) for the text path, plus ``encode_batch`` /
``decode_batch``.  This module holds no tokenisation logic: every encode/decode call goes to the
CUDA library, and raises if that library or a CUDA device is unavailable.
"""
from __future__ import annotations

import ctypes
import enum
import json
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


class TokenizerError(Exception):
    """src/errors.rs:23-59.  ``kind`` is the variant name, ``code`` the C ABI status."""

    def __init__(self, code: int, msg: str):
        lib = _lib.load()
        self.code = code
        self.kind = lib.tk_status_name(code).decode()
        self.msg = msg
        super().__init__("%s: %s" % (self.kind, msg))


class SpecialTokenPolicy(enum.IntEnum):
    """src/special_tokens.rs:129-136"""
    Ignore = 0
    Keep = 1
    Raise = 2


class SplitMode(enum.IntEnum):
    """include/tekken_b200.h tk_split_mode"""
    Reference = 0      # the literal of src/tekkenizer.rs:123 (what the reference computes)
    Config = 1         # the pattern stored in tekken.json


class TokenizerVersion(enum.IntEnum):
    """src/config.rs:97-103"""
    V3 = 3
    V7 = 7
    V11 = 11
    V13 = 13

    @classmethod
    def from_string(cls, s: str) -> Optional["TokenizerVersion"]:
        return {"v3": cls.V3, "v7": cls.V7, "v11": cls.V11, "v13": cls.V13}.get(s)

    def as_str(self) -> str:
        return "v%d" % int(self)


def _check(rc: int):
    if rc != 0:
        raise TokenizerError(rc, _lib.load().tk_last_error().decode("utf-8", "replace"))


def _policy(p) -> int:
    if isinstance(p, str):
        return int(SpecialTokenPolicy[p])
    return int(p)


def _u8(text) -> np.ndarray:
    if isinstance(text, str):
        text = text.encode("utf-8")
    if isinstance(text, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(text), dtype=np.uint8)
        return a
    return np.ascontiguousarray(text, dtype=np.uint8)


def _take(ptr: int, nbytes: int, dtype) -> np.ndarray:
    """Copy a library-owned buffer into a numpy array and release it."""
    lib = _lib.load()
    if nbytes:
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        out = np.frombuffer(buf, dtype=dtype).copy()
    else:
        out = np.zeros(0, dtype=dtype)
    lib.tk_buffer_free(ptr)
    return out


class Tekkenizer:
    """Drop-in for ``tekken::tekkenizer::Tekkenizer`` (text path) running on one B200."""

    def __init__(self, handle: int):
        self._h = handle
        self._lib = _lib.load()

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_file(cls, path, device: int = 0, split: "SplitMode" = 0) -> "Tekkenizer":
        """Tekkenizer::from_file (src/tekkenizer.rs:222-248).  ``split`` = SplitMode.Config honours the pattern stored
        in the file (Mistral's own) instead of the literal the reference hard-codes."""
        lib = _lib.load()
        h = ctypes.c_void_p()
        _check(lib.tk_load_file_ex(str(path).encode(), device, int(split), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def new(cls, vocab: Sequence[dict], special_tokens: Sequence[dict], pattern: str, vocab_size: int,
            num_special_tokens: int, version, audio_config=None, device: int = 0, split: "SplitMode" = 0) -> "Tekkenizer":
        """Tekkenizer::new (src/tekkenizer.rs:71-191).  ``pattern`` is ignored as in the reference;
        ``audio_config`` is outside the text path and must be None."""
        if audio_config is not None:
            raise TokenizerError(-5, "audio is outside the accelerated text path")
        lib = _lib.load()
        if isinstance(version, str):
            v = TokenizerVersion.from_string(version)
            if v is None:
                raise TokenizerError(-6, "Unknown version: %s" % version)
            version = v
        keep = []
        va = (_lib.VocabEntry * max(len(vocab), 1))()
        for i, e in enumerate(vocab):
            b = e["token_bytes"].encode() if isinstance(e["token_bytes"], str) else bytes(e["token_bytes"])
            keep.append(b)
            va[i].rank = int(e["rank"])
            va[i].token_bytes_b64 = b
        sa = (_lib.SpecialEntry * max(len(special_tokens), 1))()
        for i, e in enumerate(special_tokens):
            b = e["token_str"].encode("utf-8")
            keep.append(b)
            sa[i].rank = int(e["rank"])
            sa[i].token_str = b
            sa[i].is_control = 1 if e.get("is_control", True) else 0
        h = ctypes.c_void_p()
        _check(lib.tk_new_ex(va, len(vocab), sa, len(special_tokens), pattern.encode("utf-8"), vocab_size,
                             num_special_tokens, int(version), device, int(split), ctypes.byref(h)))
        return cls(h.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tk_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- accessors (src/tekkenizer.rs:261-350, 574-600) ------------------------------------
    def vocab_size(self) -> int:
        return self._lib.tk_vocab_size(self._h)

    def num_special_tokens(self) -> int:
        return self._lib.tk_num_special_tokens(self._h)

    def version(self) -> TokenizerVersion:
        return TokenizerVersion(self._lib.tk_version_of(self._h))

    def device(self) -> int:
        return self._lib.tk_device_of(self._h)

    def split_mode(self) -> SplitMode:
        return SplitMode(self._lib.tk_split_mode_of(self._h))

    def get_control_token(self, token_str: str) -> int:
        v = ctypes.c_uint32()
        _check(self._lib.tk_get_control_token(self._h, token_str.encode("utf-8"), ctypes.byref(v)))
        return v.value

    def bos_id(self) -> int:
        return self.get_control_token("<s>")

    def eos_id(self) -> int:
        return self.get_control_token("</s>")

    def pad_id(self) -> int:
        return self.get_control_token("<pad>")

    def unk_id(self) -> int:
        return self.get_control_token("<unk>")

    def is_special_token(self, token_id: int) -> bool:
        return bool(self._lib.tk_is_special_token(self._h, token_id))

    def is_byte(self, token_id: int) -> bool:
        return bool(self._lib.tk_is_byte(self._h, token_id))

    def vocab_piece(self, token_id: int) -> str:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_vocab_piece(self._h, token_id, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value).decode("utf-8")

    def vocab(self) -> List[str]:
        return [self.vocab_piece(i) for i in range(self.vocab_size())]

    def id_to_piece(self, token_id: int) -> str:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_id_to_piece(self._h, token_id, ctypes.byref(p), ctypes.byref(n)))
        return _take(p.value, n.value, np.uint8).tobytes().decode("utf-8")

    def id_to_byte_piece(self, token_id: int, special_token_policy=SpecialTokenPolicy.Keep) -> bytes:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_id_to_byte_piece(self._h, token_id, _policy(special_token_policy), ctypes.byref(p),
                                             ctypes.byref(n)))
        return _take(p.value, n.value, np.uint8).tobytes()

    # ---- encode (src/tekkenizer.rs:378-405) ---------------------------------------------------
    def encode_np(self, text, add_bos: bool, add_eos: bool) -> np.ndarray:
        a = _u8(text)
        out, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_encode(self._h, a.ctypes.data if len(a) else None, len(a), int(add_bos), int(add_eos),
                                   ctypes.byref(out), ctypes.byref(n)))
        return _take(out.value, n.value * 4, np.uint32)

    def encode(self, text, add_bos: bool, add_eos: bool) -> List[int]:
        return self.encode_np(text, add_bos, add_eos).tolist()

    def encode_batch_np(self, data, doc_off, add_bos: bool, add_eos: bool):
        """Flat form: ``data`` = concatenated UTF-8 bytes, ``doc_off`` = n_docs+1 byte offsets.
        Returns (ids uint32, tok_off uint64[n_docs+1])."""
        a = _u8(data)
        off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        n_docs = len(off) - 1
        tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
        _check(self._lib.tk_encode_batch(self._h, a.ctypes.data if len(a) else None, off.ctypes.data, n_docs,
                                         int(add_bos), int(add_eos), ctypes.byref(tok), ctypes.byref(toff)))
        tok_off = _take(toff.value, (n_docs + 1) * 8, np.uint64)
        return _take(tok.value, int(tok_off[-1]) * 4, np.uint32), tok_off

    def encode_batch(self, texts: Sequence, add_bos: bool, add_eos: bool) -> List[List[int]]:
        bs = [t.encode("utf-8") if isinstance(t, str) else bytes(t) for t in texts]
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            np.cumsum(np.fromiter((len(b) for b in bs), dtype=np.uint64, count=len(bs)), out=off[1:])
        ids, toff = self.encode_batch_np(b"".join(bs), off, add_bos, add_eos)
        return [ids[int(toff[i]):int(toff[i + 1])].tolist() for i in range(len(bs))]

    def encode_batch_device(self, d_data: int, d_doc_off: int, n_docs: int, total_bytes: int, add_bos: bool,
                            add_eos: bool, d_tokens: int, capacity: int, d_tok_off: int, stream: int = 0) -> int:
        """Zero-copy form on raw device pointers; returns the total id count."""
        n = ctypes.c_uint64()
        _check(self._lib.tk_encode_batch_device(self._h, d_data, d_doc_off, n_docs, total_bytes, int(add_bos), int(add_eos),
                                                d_tokens, capacity, d_tok_off, ctypes.byref(n), stream))
        return n.value

    # ---- decode (src/tekkenizer.rs:436-560) ---------------------------------------------------
    def decode_bytes(self, tokens, special_token_policy=SpecialTokenPolicy.Ignore) -> bytes:
        ids = np.ascontiguousarray(tokens, dtype=np.uint32)
        out, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_decode(self._h, ids.ctypes.data if len(ids) else None, len(ids), _policy(special_token_policy),
                                   ctypes.byref(out), ctypes.byref(n)))
        return _take(out.value, n.value, np.uint8).tobytes()

    def decode(self, tokens, special_token_policy=SpecialTokenPolicy.Ignore) -> str:
        return self.decode_bytes(tokens, special_token_policy).decode("utf-8")

    def decode_all(self, tokens, special_token_policy=SpecialTokenPolicy.Ignore) -> List[str]:
        ids = np.ascontiguousarray(tokens, dtype=np.uint32)
        out, pe, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_decode_all(self._h, ids.ctypes.data if len(ids) else None, len(ids), _policy(special_token_policy),
                                       ctypes.byref(out), ctypes.byref(pe), ctypes.byref(n)))
        ends = _take(pe.value, n.value * 8, np.uint64)
        total = int(ends[-1]) if len(ends) else 0
        raw = _take(out.value, total, np.uint8).tobytes()
        res, s = [], 0
        for e in ends:
            res.append(raw[s:int(e)].decode("utf-8"))
            s = int(e)
        return res

    def decode_batch_np(self, ids, tok_off, special_token_policy=SpecialTokenPolicy.Ignore):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        off = np.ascontiguousarray(tok_off, dtype=np.uint64)
        n_docs = len(off) - 1
        out, boff = ctypes.c_void_p(), ctypes.c_void_p()
        bad = ctypes.c_uint64(0)
        _check(self._lib.tk_decode_batch(self._h, ids.ctypes.data if len(ids) else None, off.ctypes.data, n_docs,
                                         _policy(special_token_policy), ctypes.byref(out), ctypes.byref(boff), ctypes.byref(bad)))
        byte_off = _take(boff.value, (n_docs + 1) * 8, np.uint64)
        return _take(out.value, int(byte_off[-1]), np.uint8), byte_off

    def decode_batch(self, token_lists: Sequence[Sequence[int]], special_token_policy=SpecialTokenPolicy.Ignore) -> List[str]:
        off = np.zeros(len(token_lists) + 1, dtype=np.uint64)
        if len(token_lists):
            np.cumsum(np.fromiter((len(t) for t in token_lists), dtype=np.uint64, count=len(token_lists)), out=off[1:])
        flat = np.concatenate([np.asarray(t, dtype=np.uint32) for t in token_lists]) if len(token_lists) else np.zeros(0, np.uint32)
        raw, boff = self.decode_batch_np(flat, off, special_token_policy)
        b = raw.tobytes()
        return [b[int(boff[i]):int(boff[i + 1])].decode("utf-8") for i in range(len(token_lists))]

    def decode_batch_device(self, d_ids: int, d_tok_off: int, n_docs: int, total_ids: int, special_token_policy, d_out: int,
                            capacity: int, d_byte_off: int, d_status: int = 0, stream: int = 0) -> int:
        n, bad = ctypes.c_uint64(), ctypes.c_uint64()
        _check(self._lib.tk_decode_batch_device(self._h, d_ids, d_tok_off, n_docs, total_ids, _policy(special_token_policy),
                                                d_out, capacity, d_byte_off, d_status, ctypes.byref(n), ctypes.byref(bad), stream))
        return n.value

    # ---- audio token counting (src/audio.rs:555-591, src/tekkenizer.rs:728-760) ----------------------
    def has_audio_support(self) -> bool:
        return bool(self._lib.tk_has_audio_support(self._h))

    def audio_config(self) -> Optional[dict]:
        c = _lib.AudioConfig()
        if self._lib.tk_audio_config_of(self._h, ctypes.byref(c)) != 0:
            return None
        return {"sampling_rate": c.sampling_rate, "frame_rate": c.frame_rate, "chunk_length_s": c.chunk_length_s if c.chunk_length_s > 0 else None,
                "audio_encoding_config": {"num_mel_bins": c.num_mel_bins, "hop_length": c.hop_length, "window_size": c.window_size}}

    def encode_audio_tokens(self, n_samples: int) -> List[int]:
        """AudioEncoding.tokens for a clip of n_samples samples at the configured sampling rate."""
        out, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self._lib.tk_encode_audio_tokens(self._h, int(n_samples), ctypes.byref(out), ctypes.byref(n)))
        return _take(out.value, n.value * 4, np.uint32).tolist()

    # ---- instrumentation ----------------------------------------------------------------------
    def set_stage_timing(self, enabled: bool):
        self._lib.tk_set_stage_timing(self._h, int(enabled))

    def last_stage_times(self) -> dict:
        names = (ctypes.c_char_p * 32)()
        ms = (ctypes.c_float * 32)()
        n = self._lib.tk_last_stage_times(self._h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def last_encode_counters(self) -> dict:
        """Work counters of the last device-pointer encode call (pieces per merge class, table lookups)."""
        v = (ctypes.c_uint64 * 20)()
        n = self._lib.tk_last_encode_counters(self._h, v, 20)
        if n < 13:
            return {}
        by_class = [int(v[i]) for i in range(9)]
        out = {"by_class": by_class, "pieces_queued": sum(by_class), "n_long": int(v[9]), "n_huge": int(v[10]),
               "pair_lookups": int(v[11]), "byte_pair_lookups": int(v[12])}
        if n >= 17:
            out["long_piece_rounds"] = {"single_rank": int(v[13]), "multi_rank": int(v[14]), "multi_rank_cut": int(v[15]),
                                        "merges_by_multi_rank": int(v[16])}
        return out


def shard_plan(doc_off, n_shards: int) -> np.ndarray:
    """Byte-balanced contiguous document ranges for n_shards GPUs (no collective needed)."""
    lib = _lib.load()
    off = np.ascontiguousarray(doc_off, dtype=np.uint64)
    out = np.zeros(n_shards + 1, dtype=np.uint64)
    _check(lib.tk_shard_plan(off.ctypes.data, len(off) - 1, n_shards, out.ctypes.data))
    return out


def encode_batch_multi(tokenizers: Sequence["Tekkenizer"], data, doc_off, add_bos: bool, add_eos: bool):
    """One call over several handles of the same tokenizer (one per GPU): chunks of the batch are dealt to the devices,
    ids and offsets come back in document order in one buffer.  Returns (ids uint32, tok_off uint64[n_docs+1])."""
    lib = _lib.load()
    a = _u8(data)
    off = np.ascontiguousarray(doc_off, dtype=np.uint64)
    n_docs = len(off) - 1
    arr = (ctypes.c_void_p * len(tokenizers))(*[t._h for t in tokenizers])
    tok, toff = ctypes.c_void_p(), ctypes.c_void_p()
    _check(lib.tk_encode_batch_multi(arr, len(tokenizers), a.ctypes.data if len(a) else None, off.ctypes.data, n_docs,
                                     int(add_bos), int(add_eos), ctypes.byref(tok), ctypes.byref(toff)))
    tok_off = _take(toff.value, (n_docs + 1) * 8, np.uint64)
    return _take(tok.value, int(tok_off[-1]) * 4, np.uint32), tok_off


def encode_file(tokenizers: Sequence["Tekkenizer"], text_path: str, tokens_path: str, offsets_path: Optional[str] = None,
                delimiter: Optional[int] = 10, add_bos: bool = True, add_eos: bool = True, npy: bool = False) -> dict:
    """Text file -> shard of u32 ids (+ u64 token offsets), streamed through the GPU(s) window by window."""
    lib = _lib.load()
    arr = (ctypes.c_void_p * len(tokenizers))(*[t._h for t in tokenizers])
    st = _lib.FileStats()
    _check(lib.tk_encode_file(arr, len(tokenizers), str(text_path).encode(), -1 if delimiter is None else int(delimiter), int(add_bos), int(add_eos),
                              str(tokens_path).encode(), str(offsets_path).encode() if offsets_path else None, 1 if npy else 0, ctypes.byref(st)))
    return {"n_docs": st.n_docs, "n_bytes": st.n_bytes, "n_tokens": st.n_tokens, "seconds": st.seconds}


def set_chunk_bytes(n: int):
    """Tuning knob: chunk size of the host-buffer calls (0 = default 128 MB)."""
    _lib.load().tk_set_chunk_bytes(int(n))


def set_pack_ids(mode: int):
    """Tuning knob: how ids cross PCIe in host-buffer encodes (-1 default: 18/24-bit stream for large calls; 0 never; 18 / 24 always)."""
    _lib.load().tk_set_pack_ids(int(mode))


def audio_token_count(cfg: dict, n_samples: int):
    """(padded samples, number of [AUDIO] ids) for a clip under an AudioConfig given as the tekken.json `audio` dict."""
    lib = _lib.load()
    e = cfg["audio_encoding_config"]
    c = _lib.AudioConfig(int(cfg["sampling_rate"]), float(cfg["frame_rate"]), int(e["num_mel_bins"]), int(e["hop_length"]), int(e["window_size"]),
                         float(cfg["chunk_length_s"]) if cfg.get("chunk_length_s") else -1.0)
    p, n = ctypes.c_uint64(), ctypes.c_uint64()
    _check(lib.tk_audio_token_count(ctypes.byref(c), int(n_samples), ctypes.byref(p), ctypes.byref(n)))
    return p.value, n.value


def kernel_launch_count() -> int:
    return int(_lib.load().tk_kernel_launch_count())
