"""ctypes view of include/tekken_b200.h.  Fails loudly if the CUDA library is missing."""
from __future__ import annotations

import ctypes
import os

from . import _build

_lib = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_vp = ctypes.c_void_p


class VocabEntry(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_uint64), ("token_bytes_b64", ctypes.c_char_p)]


class AudioConfig(ctypes.Structure):
    _fields_ = [("sampling_rate", ctypes.c_uint64), ("frame_rate", ctypes.c_double), ("num_mel_bins", ctypes.c_uint64),
                ("hop_length", ctypes.c_uint64), ("window_size", ctypes.c_uint64), ("chunk_length_s", ctypes.c_double)]


class FileStats(ctypes.Structure):
    _fields_ = [("n_docs", ctypes.c_uint64), ("n_bytes", ctypes.c_uint64), ("n_tokens", ctypes.c_uint64), ("seconds", ctypes.c_double)]


class SpecialEntry(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_uint64), ("token_str", ctypes.c_char_p), ("is_control", ctypes.c_int)]


# every symbol include/tekken_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "tk_load_file": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "tk_new": (ctypes.c_int, [ctypes.POINTER(VocabEntry), ctypes.c_size_t, ctypes.POINTER(SpecialEntry), ctypes.c_size_t,
                              ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                              ctypes.POINTER(c_vp)]),
    "tk_load_file_ex": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "tk_new_ex": (ctypes.c_int, [ctypes.POINTER(VocabEntry), ctypes.c_size_t, ctypes.POINTER(SpecialEntry), ctypes.c_size_t,
                                 ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.POINTER(c_vp)]),
    "tk_split_mode_of": (ctypes.c_int, [c_vp]),
    "tk_deprecated_special_tokens": (ctypes.c_size_t, [ctypes.POINTER(ctypes.POINTER(SpecialEntry))]),
    "tk_free": (None, [c_vp]),
    "tk_vocab_size": (ctypes.c_size_t, [c_vp]),
    "tk_num_special_tokens": (ctypes.c_size_t, [c_vp]),
    "tk_version_of": (ctypes.c_int, [c_vp]),
    "tk_device_of": (ctypes.c_int, [c_vp]),
    "tk_get_control_token": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_u32p]),
    "tk_bos_id": (ctypes.c_int, [c_vp, c_u32p]),
    "tk_eos_id": (ctypes.c_int, [c_vp, c_u32p]),
    "tk_pad_id": (ctypes.c_int, [c_vp, c_u32p]),
    "tk_unk_id": (ctypes.c_int, [c_vp, c_u32p]),
    "tk_is_special_token": (ctypes.c_int, [c_vp, ctypes.c_uint32]),
    "tk_is_byte": (ctypes.c_int, [c_vp, ctypes.c_uint32]),
    "tk_vocab_piece": (ctypes.c_int, [c_vp, ctypes.c_uint32, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "tk_id_to_piece": (ctypes.c_int, [c_vp, ctypes.c_uint32, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "tk_id_to_byte_piece": (ctypes.c_int, [c_vp, ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(c_vp),
                                           ctypes.POINTER(ctypes.c_size_t)]),
    "tk_encode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp),
                                 ctypes.POINTER(ctypes.c_size_t)]),
    "tk_encode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp),
                                       ctypes.POINTER(c_vp)]),
    "tk_encode_batch_multi": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t, c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)]),
    "tk_encode_batch_device": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                              c_vp, ctypes.c_uint64, c_vp, c_u64p, c_vp]),
    "tk_decode": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_vp),
                                 ctypes.POINTER(ctypes.c_size_t)]),
    "tk_decode_all": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                     ctypes.POINTER(ctypes.c_size_t)]),
    "tk_decode_batch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_vp),
                                       ctypes.POINTER(c_vp), c_u64p]),
    "tk_decode_batch_device": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int, c_vp,
                                              ctypes.c_uint64, c_vp, c_vp, c_u64p, c_u64p, c_vp]),
    "tk_encode_file": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(FileStats)]),
    "tk_has_audio_support": (ctypes.c_int, [c_vp]),
    "tk_audio_config_of": (ctypes.c_int, [c_vp, ctypes.POINTER(AudioConfig)]),
    "tk_audio_token_count": (ctypes.c_int, [ctypes.POINTER(AudioConfig), ctypes.c_uint64, c_u64p, c_u64p]),
    "tk_encode_audio_tokens": (ctypes.c_int, [c_vp, ctypes.c_uint64, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "tk_shard_plan": (ctypes.c_int, [c_vp, ctypes.c_size_t, ctypes.c_size_t, c_vp]),
    "tk_buffer_free": (None, [c_vp]),
    "tk_last_error": (ctypes.c_char_p, []),
    "tk_status_name": (ctypes.c_char_p, [ctypes.c_int]),
    "tk_set_chunk_bytes": (None, [ctypes.c_uint64]),
    "tk_debug_bounds_violations": (ctypes.c_longlong, [c_vp, c_u64p]),
    "tk_set_pack_ids": (None, [ctypes.c_int]),
    "tk_debug_unpack_ids": (ctypes.c_int, [c_vp, ctypes.c_size_t, ctypes.c_int, c_vp]),
    "tk_kernel_launch_count": (ctypes.c_uint64, []),
    "tk_set_stage_timing": (None, [c_vp, ctypes.c_int]),
    "tk_last_encode_counters": (ctypes.c_size_t, [c_vp, c_u64p, ctypes.c_size_t]),
    "tk_last_stage_times": (ctypes.c_size_t, [c_vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float),
                                              ctypes.c_size_t]),
}


def library_path() -> str:
    return _build.LIB


def load():
    """Load libtekken_b200.so (building it with nvcc first if it is stale or missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("TEKKEN_B200_NO_BUILD") != "1":
        try:
            _build.build()
        except Exception as e:  # keep going only if a previously built library exists
            if not os.path.exists(_build.LIB):
                raise RuntimeError("libtekken_b200.so is missing and could not be built: %s "
                                   "(the encode/decode path has no CPU fallback)" % e)
    if not os.path.exists(_build.LIB):
        raise RuntimeError("libtekken_b200.so is missing (the encode/decode path has no CPU fallback)")
    lib = ctypes.CDLL(_build.LIB)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
