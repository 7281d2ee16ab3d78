#!/usr/bin/env python
"""Generate tests/golden/engine_fixtures.json from the upstream engine itself.

TEST INFRASTRUCTURE.  Runs in the build container.  The engine is Python `tiktoken` (the
openai/tiktoken Rust CoreBPE that tiktoken-rs 0.7.0 -- the reference's dependency, Cargo.toml:40 --
vendors), configured exactly as the reference configures it (src/tekkenizer.rs:122-126): the
hard-coded pattern, the first `vocab_size - num_special` ranks of the vocabulary file, an EMPTY
special-token map; ids get the reference's glue (+num_special, BOS=1, EOS=2; :390-402).

The fixtures pin (a) the oracle restatement and (b) the CUDA path on machines where the engine
is not importable.  Contents: SURVEY.md Appendix A cases, adversarial split/merge cases, seeded
fuzz strings, and checksums + token counts of the synthetic corpora (configs 1-4, reduced sizes).
"""
import base64
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tekken_oracle as TO  # noqa: E402
from tekken_rs_b200 import corpus  # noqa: E402

# adversarial alphabet for the fuzz strings (escapes only, so the file stays plain ASCII)
FUZZ_ALPHABET = list(
    "aAbsStTrReEvVmMlLdD'’ſ  \t\n\r\n 　 !?.,;-_()[]{}0123456789٣९①"
    "éüßçñöКириллица中文日本語"
    "かなカナ한국어ไทยคำकाि्अ"
    "\U0001f600\U0001f44d\U0001f3fd‍\U0001f469\U0001f4bb€£∑√\x00\x7f\x85 ")

HAND = [
    "", " ", "\n", "a", "Hello, world! This is a test.", "का", "नमस्ते दुनिया",
    "I'm, you'LL, it'ſ", "a  \n\n  b\t c", "x = 12345;\r\n", "\U0001f680\U0001f680 go", "北京欢迎你", "a" * 16,
    "Hello\x00World", "!!word !word  !b", "1234567", "x!!!\n\ny", "a \n\n  b", "café", "a'ſb", "'s", " 's", "!'s",
    "''s", "'ll's", "x'sy", "\t's", "a\n  ", "!\n \n", "!\n\n\n x", "  \n  \n  x", "\r\r\n\n a", "   x", "　　",
    "1٣९①2", "<s>[INST] hi [/INST]</s>", "tab\there", "trailing   ", "   leading", "a b cd",
    "don't won'T I'D he'S we'VE i'M", "\U0001f44d\U0001f3fd\U0001f469‍\U0001f4bb emoji‍zwj", "ＡＢＣ１２３",
    "x" * 100, "ab" * 50, " " * 70, "\n" * 70, "1" * 100, "0" * 64 + "a" + "9" * 65,
    "mixed 中文 and English и русский 123 ٤٥٦",
    "https://example.com/a_b-c?d=1&e=2#frag", "e=mc^2; f(x)=x**2",
]


def main():
    tk = TO.OracleTekkenizer.from_file(TO.find_tekken_json())
    enc = TO.tiktoken_engine(tk.ranks)
    ns = tk.num_special

    def ref(text, bos, eos):
        ids = [x + ns for x in enc.encode_ordinary(text)]
        return ([1] if bos else []) + ids + ([2] if eos else [])

    cases = []
    for i, t in enumerate(HAND):
        bos, eos = bool(i & 1), bool(i & 2)
        cases.append({"text": t, "add_bos": bos, "add_eos": eos, "ids": ref(t, bos, eos)})
    rng = random.Random(2026)
    for _ in range(400):
        L = rng.choice([1, 2, 3, 5, 8, 13, 30, 33, 64, 70, 100])
        t = "".join(rng.choice(FUZZ_ALPHABET) for _ in range(L))
        cases.append({"text": t, "add_bos": False, "add_eos": False, "ids": ref(t, False, False)})
    # single pieces isolating the merge loop
    pieces = []
    for _ in range(150):
        L = rng.choice([2, 3, 4, 7, 12, 20, 31, 32, 33, 63, 64, 65, 90, 200, 513, 600])
        kind = rng.choice(["lower", "cjk", "bytes", "ab", "emoji", "cyr"])
        if kind == "lower":
            b = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(L)).encode()
        elif kind == "cjk":
            b = "".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(max(1, L // 3))).encode()
        elif kind == "bytes":
            b = bytes(rng.randrange(256) for _ in range(L))
        elif kind == "ab":
            b = (b"ab" * L)[:L]
        elif kind == "emoji":
            b = "".join(rng.choice("\U0001f600\U0001f44d\U0001f3fd‍\U0001f469\U0001f4bb\U0001f680") for _ in range(max(1, L // 4))).encode()
        else:
            b = "".join(rng.choice("абвгдежзиклмнопрст") for _ in range(max(1, L // 2))).encode()
        pieces.append({"bytes_b64": base64.b64encode(b).decode(), "ranks": enc._encode_single_piece(b)})
    corp = []
    specs = [("english_like", {"n_bytes": 1 << 18, "seed": 1234}), ("mixed_script_docs", {"n_docs": 2000, "seed": 42}),
             ("single_long_document", {"n_bytes": 1 << 20, "seed": 7}),
             ("adversarial_pieces", {"n_pieces": 7, "piece_bytes": 2048, "seed": 11})]
    for name, kw in specs:
        if name == "mixed_script_docs":
            data, off = corpus.mixed_script_docs(**kw)
            raw = data.tobytes()
            ids = []
            for d in range(len(off) - 1):
                ids.extend(ref(raw[int(off[d]):int(off[d + 1])].decode(), True, True))
        else:
            raw = getattr(corpus, name)(**kw)
            ids = ref(raw.decode(), True, True)
        a = np.asarray(ids, dtype=np.uint32)
        corp.append({"generator": name, "kwargs": kw, "n_bytes": len(raw), "n_ids": len(a),
                     "bytes_checksum": corpus.checksum64(np.frombuffer(raw, dtype=np.uint8)),
                     "ids_checksum": corpus.checksum64(a), "ids_head": a[:32].tolist(), "ids_tail": a[-32:].tolist()})
    import tiktoken
    out = {"engine": "tiktoken " + tiktoken.__version__, "pattern": TO.REFERENCE_PATTERN, "num_special": ns,
           "vocab": "mistral_common tekken_240911.json (first %d ranks)" % len(tk.ranks),
           "cases": cases, "pieces": pieces, "corpora": corp}
    p = os.path.join(ROOT, "tests", "golden", "engine_fixtures.json")
    json.dump(out, open(p, "w"), ensure_ascii=True)
    print(len(cases), "cases", len(pieces), "pieces", len(corp), "corpora ->", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
