#!/usr/bin/env python
"""Extract the Unicode class tables the reference's hard-coded split regex uses.

TEST/BUILD INFRASTRUCTURE (oracle side).  The reference (src/tekkenizer.rs:123) compiles
its pattern with tiktoken-rs 0.7.0 -> fancy-regex -> regex-syntax, whose Unicode tables
are not under /root/reference.  The same upstream engine is importable here as the
Python package `tiktoken` (0.12.0), so membership of every scalar value in \\p{L},
\\p{N}, \\s and the case-insensitive contraction letters is *measured* from that engine
(SURVEY.md section 8c recipe) instead of being taken from a Unicode library of a
different version.

Output: a JSON file with sorted inclusive ranges per class, plus generated C headers
(one for oracle/, one for the CUDA library) holding the same ranges.
"""
import json
import sys
import os
import tiktoken

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def members(pat):
    enc = tiktoken.Encoding("probe", pat_str=pat,
                            mergeable_ranks={bytes([i]): i for i in range(256)},
                            special_tokens={})
    cps = [c for c in range(0x110000) if not (0xD800 <= c <= 0xDFFF)]
    out = []
    B = 1 << 15
    for i in range(0, len(cps), B):
        chunk = cps[i:i + B]
        res = enc.encode_ordinary_batch([chr(c) for c in chunk], num_threads=8)
        out.extend(c for c, r in zip(chunk, res) if len(r) > 0)
    return out


def to_ranges(cps):
    rs = []
    for c in cps:
        if rs and rs[-1][1] == c - 1:
            rs[-1][1] = c
        else:
            rs.append([c, c])
    return rs


def main():
    tables = {}
    for name, pat in (("L", r"\p{L}"), ("N", r"\p{N}"), ("S", r"\s")):
        m = members(pat)
        tables[name] = to_ranges(m)
        print(name, len(m), "code points", len(tables[name]), "ranges", file=sys.stderr)
    fold = {}
    for ch in "stmdrevl":
        m = members("(?i:%s)" % ch)
        fold[ch] = m
        print("fold", ch, [hex(x) for x in m], file=sys.stderr)
    tables["fold"] = fold
    tables["engine"] = "tiktoken " + tiktoken.__version__
    with open(os.path.join(ROOT, "oracle", "unicode_tables.json"), "w") as f:
        json.dump(tables, f)


if __name__ == "__main__":
    main()
