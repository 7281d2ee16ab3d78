#!/usr/bin/env python
"""Generate tests/golden/config_pattern_fixtures.json: splits and ids under the pattern STORED in tekken.json.

TEST INFRASTRUCTURE, groundwork for SURVEY section 8(f) rank 1 (config-driven pattern).  The reference ignores
config.pattern (src/tekkenizer.rs:74,123); Mistral's own stack compiles it.  The engine is Python `tiktoken` (the
same Rust CoreBPE) built with that pattern, the first `vocab_size - num_special` ranks and an empty special map; ids
get the reference's glue (+num_special).  Pins oracle.split_config / encode_config where the engine is not importable."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tekken_oracle as TO  # noqa: E402

# case transitions, caseless letters, marks, digits, the slash of the punctuation tail, whitespace kinds
ALPHABET = list("aAbBzZéÉßǅǲʰˀ中あアकाि्ًَ́̈ 　\t\n\r\n/\\.,;:!?-_'\"()0123٣९①€😀👍🏽‍ ")

HAND = [
    "", " ", "a", "A", "Hello", "HELLO", "helloWorld", "HELLOworld", "HelloWORLD", "hELLO", "ABCdefGHI", "iPhone",
    "McDonald's", "NASA's", "don't", "I'M", "x1y", "12345", "3.14", "a/b/c", "http://x.y/z", "end.\n/next", "!!!\r\n//",
    " /usr/bin", "का", "काि्", "é", "é", "a" + "́" * 3, "́", " ́a", ".́", "́.", "中文ABC", "ABC中文", "ǅa", "Aǅ", "ʰA", "Aʰ",
    "Aʰb", "AʰB", "AB ", " AB", "\tAB", "\nAB", "A B", "A  B", "a \n\n  b", "x = 12345;\r\n", "   leading", "trailing   ",
    "Hello, world! This is a test.", "ＡＢＣ１２３ａｂｃ", "ÀÉÎõü", "ǅǅǅ", "ØRSTED ørsted Ørsted øRSTED",
]


def main():
    path = TO.find_tekken_json()
    cfg = json.load(open(path))["config"]
    orc = TO.OracleTekkenizer.from_file(path)
    import tiktoken
    enc = tiktoken.Encoding("tekken-config-pattern", pat_str=cfg["pattern"],
                            mergeable_ranks={b: r for r, b in enumerate(orc.ranks)}, special_tokens={})
    rng = random.Random(2024)
    texts = list(HAND)
    for _ in range(600):
        texts.append("".join(rng.choice(ALPHABET) for _ in range(rng.choice([1, 2, 3, 5, 9, 17, 40]))))
    cases = []
    for t in texts:
        ids = enc.encode_ordinary(t)
        cases.append({"text": t, "ids": [i + orc.num_special_tokens() for i in ids]})
    out = {"engine": "tiktoken " + tiktoken.__version__, "pattern": cfg["pattern"], "cases": cases}
    dst = os.path.join(ROOT, "tests", "golden", "config_pattern_fixtures.json")
    json.dump(out, open(dst, "w"), ensure_ascii=True, indent=0)
    print(dst, len(cases), "cases")


if __name__ == "__main__":
    main()
