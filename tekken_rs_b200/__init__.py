"""tekken-rs_b200: the Tekkenizer text encode/decode path of jorge-menjivar/tekken-rs, rebuilt for
NVIDIA B200 (sm_100a) behind a C ABI (include/tekken_b200.h)."""
from .tekkenizer import (audio_token_count, SpecialTokenPolicy, SplitMode, Tekkenizer, encode_batch_multi, encode_file, set_chunk_bytes, set_pack_ids, TokenizerError, TokenizerVersion, kernel_launch_count,
                         shard_plan)
from ._build import build
from ._lib import library_path

__all__ = ["audio_token_count", "Tekkenizer", "SpecialTokenPolicy", "SplitMode", "encode_batch_multi", "encode_file", "set_chunk_bytes", "set_pack_ids", "TokenizerError", "TokenizerVersion", "shard_plan",
           "kernel_launch_count", "build", "library_path"]
