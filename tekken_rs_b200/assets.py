"""Locate / reconstruct the vocabulary file.

The reference's tests/assets/tekken.json (v7) is not shipped with the reference mount
(/root/reference/.MISSING_LARGE_BLOBS).  ``mistral_common`` (in this image) ships the same Tekken
base vocabulary as tekken_240911.json (version "v3", no special_tokens list); SURVEY.md section 8c
describes how a v7 file differs: version string, an explicit special_tokens list (the 20 built-ins
of src/tekkenizer.rs:827-930 plus [AUDIO]/[BEGIN_AUDIO] at ranks 24/25 and fillers) and an audio
block.  Token ids on the text path do not depend on any of that."""
from __future__ import annotations

import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEFAULT_ASSET = os.path.join(ROOT, "tests", "assets", "tekken.json")

BUILTIN_SPECIALS = [
    "<unk>", "<s>", "</s>", "[INST]", "[/INST]", "[AVAILABLE_TOOLS]", "[/AVAILABLE_TOOLS]", "[TOOL_RESULTS]",
    "[/TOOL_RESULTS]", "[TOOL_CALLS]", "[IMG]", "<pad>", "[IMG_BREAK]", "[IMG_END]", "[PREFIX]", "[MIDDLE]", "[SUFFIX]",
    "[SYSTEM_PROMPT]", "[/SYSTEM_PROMPT]", "[TOOL_CONTENT]",
]


def mistral_common_vocab() -> str:
    import mistral_common
    p = os.path.join(os.path.dirname(mistral_common.__file__), "data", "tekken_240911.json")
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    return p


def ensure_tekken_json(path: str = DEFAULT_ASSET) -> str:
    """Write a v7-shaped tekken.json (same vocab as mistral_common's file) if it is missing."""
    env = os.environ.get("TEKKEN_JSON")
    if env and os.path.exists(env):
        return env
    if os.path.exists(path) and os.path.getsize(path) > 1 << 20:
        return path
    src = json.load(open(mistral_common_vocab(), "r", encoding="utf-8"))
    src["config"]["version"] = "v7"
    specials = [{"rank": i, "token_str": s, "is_control": True} for i, s in enumerate(BUILTIN_SPECIALS)]
    for i in range(len(specials), 24):
        specials.append({"rank": i, "token_str": "<SPECIAL_%d>" % i, "is_control": True})
    specials.append({"rank": 24, "token_str": "[AUDIO]", "is_control": True})
    specials.append({"rank": 25, "token_str": "[BEGIN_AUDIO]", "is_control": True})
    src["special_tokens"] = specials
    src["audio"] = {"sampling_rate": 16000, "frame_rate": 12.5, "chunk_length_s": 30.0,
                    "audio_encoding_config": {"num_mel_bins": 128, "hop_length": 160, "window_size": 400}}
    os.makedirs(os.path.dirname(path), exist_ok=True)
    tmp = path + ".%d.tmp" % os.getpid()
    with open(tmp, "w", encoding="utf-8") as f:
        json.dump(src, f, ensure_ascii=False)
    os.replace(tmp, path)
    return path
