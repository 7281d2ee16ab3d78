#!/usr/bin/env python
"""Collect the reference's own known-answer vectors for the encode/decode path.

TEST INFRASTRUCTURE.  Runs in the build container only (reads /root/reference, which does
not exist on the GPU box) and writes tests/golden/reference_goldens.json, which is
committed.  Sources: tests/test_tokenizer_output.rs (20 exact encode vectors, file:line
recorded per vector), tests/test_rust_tokenizer.rs:16-19,80 (JFK decode), and the doc
example src/tekkenizer.rs:427.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                   "tests", "golden", "reference_goldens.json")


def unescape(s):
    return (s.replace('\\"', '"').replace("\\n", "\n").replace("\\t", "\t")
             .replace("\\\\", "\\"))


def main():
    src = open(os.path.join(REF, "tests/test_tokenizer_output.rs"), encoding="utf-8").read()
    encode = []
    for m in re.finditer(r'let input = "((?:[^"\\]|\\.)*)";\s*let expected_tokens = vec!\[([^\]]*)\];', src):
        line = src[:m.start()].count("\n") + 1
        ids = [int(x) for x in re.findall(r"\d+", m.group(2))]
        encode.append({"text": unescape(m.group(1)), "ids": ids, "add_bos": False, "add_eos": False,
                       "source": "tests/test_tokenizer_output.rs:%d" % line})
    jfk = open(os.path.join(REF, "tests/test_rust_tokenizer.rs"), encoding="utf-8").read()
    m = re.search(r"vec!\[([^\]]*)\]", jfk)
    jfk_ids = [int(x) for x in re.findall(r"\d+", m.group(1))]
    m2 = re.search(r'"(And so, my fellow Americans[^"]*)"', jfk)
    decode = [
        {"ids": jfk_ids, "policy": "Ignore", "text": m2.group(1),
         "source": "tests/test_rust_tokenizer.rs:16-19,80"},
        {"ids": [1, 22177, 1044, 4304, 2], "policy": "Keep", "text": "<s>Hello, world</s>",
         "source": "src/tekkenizer.rs:427"},
    ]
    json.dump({"encode": encode, "decode": decode}, open(OUT, "w"), ensure_ascii=False, indent=1)
    print(len(encode), "encode vectors,", len(decode), "decode vectors ->", OUT)


if __name__ == "__main__":
    main()
