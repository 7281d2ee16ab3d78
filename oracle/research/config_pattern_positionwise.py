#!/usr/bin/env python
"""RESEARCH, groundwork for SURVEY 8(f) rank 1 -- NOT product code, NOT used by any test of the product path.

A run-based "is a piece start" predicate for the pattern STORED in tekken.json (Mistral's own Tekken regex), the form a
GPU pre-tokeniser needs: every decision uses the classes of a bounded neighbourhood plus properties of maximal runs
(what segmented scans provide), except for one small left-to-right carry along chains of punctuation runs separated
by CR/LF (the `[\\r\\n/]*` tail of a punctuation piece can eat the leading slashes of the next punctuation run).
SURVEY Appendix B is the same exercise for the reference's hard-coded pattern.

`python oracle/research/config_pattern_positionwise.py [n_strings]` fuzzes the predicate against the sequential
restatement in the oracle (oracle.split_config, itself pinned to the engine) and prints the first mismatch.

Classes per code point: U = Lu|Lt, l = Ll, C = Lm|Lo, M = mark, N = digit, R = CR/LF, W = other whitespace
(sp = U+0020), O = everything else (sl = '/')."""
import bisect
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


class Classifier:
    def __init__(self):
        base = json.load(open(os.path.join(ROOT, "oracle", "unicode_tables.json")))
        sub = json.load(open(os.path.join(ROOT, "oracle", "unicode_subclasses.json")))
        self.tabs = []
        for name, key, src in (("U", "UPPER", sub), ("l", "LOWER", sub), ("C", "BOTH", sub), ("M", "MARK", sub),
                               ("N", "N", base), ("W", "S", base)):
            rs = src[key]
            self.tabs.append((name, [r[0] for r in rs], [r[1] for r in rs]))

    def cls(self, ch):
        c = ord(ch)
        if c == 0x0A or c == 0x0D:
            return "R"
        for name, lo, hi in self.tabs:
            i = bisect.bisect_right(lo, c) - 1
            if i >= 0 and c <= hi[i]:
                return name
        return "O"


def starts_by_runs(text, K):
    """K[i] = class of code point i.  Returns the list of booleans 'a piece starts at code point i'."""
    n = len(K)
    S = [False] * n
    if n == 0:
        return S
    # ------------------------------------------------------------------ 1. punctuation-and-mark runs (T-runs)
    # wordy[i]: mark i acts as a word character (C-like).  absorbed[i]: CR/LF or '/' eaten by the tail of a
    # punctuation piece.  prefix_o[i]: single punctuation char that starts a piece and is the prefix of the word
    # piece that follows it (marks and/or letters).
    wordy = [False] * n
    absorbed = [False] * n
    tail_open = False                      # the carry: a punctuation piece ended just before i and its tail is running
    i = 0
    while i < n:
        k = K[i]
        if k not in "OM":
            if tail_open and k == "R":
                absorbed[i] = True
            else:
                tail_open = False
            i += 1
            continue
        a = i
        b = i
        while b < n and K[b] in "OM":
            b += 1
        # leading slashes eaten by a running tail
        a2 = a
        while tail_open and a2 < b and text[a2] == "/":
            absorbed[a2] = True
            a2 += 1
        if a2 == b:                        # the whole run was slashes of the tail: the tail keeps running
            i = b
            continue
        tail_open = False
        entered_by_space = a2 == a and a > 0 and text[a - 1] == " " and K[a] == "O"
        # trigger: first punctuation char at the start of a block (run start or after marks) that is followed by another
        # punctuation char, or the first char when the run is entered through a space
        g = None
        p = a2
        while p < b:
            if K[p] == "O" and (p == a2 or K[p - 1] == "M"):
                if (p + 1 < b and K[p + 1] == "O") or (p == a2 and entered_by_space):
                    g = p
                    break
            p += 1
        lim = g if g is not None else b
        for p in range(a2, lim):
            if K[p] == "O":
                S[p] = True                # a single punctuation char: starts a piece (prefix of marks/letters, or alone)
            else:
                wordy[p] = True
        if g is not None:
            S[g] = not (g == a2 and entered_by_space)
            ends_b4 = True
        else:
            last = b - 1
            ends_b4 = K[last] == "O" and not (b < n and K[b] in "UlC")
        tail_open = ends_b4
        i = b
    # ------------------------------------------------------------------ 2. digits: every digit is a piece
    for i in range(n):
        if K[i] == "N":
            S[i] = True
    # ------------------------------------------------------------------ 3. whitespace runs (R|W), minus the absorbed CR/LF
    i = 0
    while i < n:
        if K[i] not in "RW":
            i += 1
            continue
        a = i
        e = i
        while e < n and K[e] in "RW":
            e += 1
        s = a
        while s < e and absorbed[s]:
            s += 1
        if s < e:
            last_r = -1
            for p in range(s, e):
                if K[p] == "R":
                    last_r = p
            for p in range(s, e):
                if p == s or p == last_r + 1 or (p == e - 1 and e < n and p > last_r):
                    S[p] = True
        i = e
    # ------------------------------------------------------------------ 4. word runs: letters and wordy marks
    def is_word(p):
        return K[p] in "UlC" or (K[p] == "M" and wordy[p])

    def kind(p):                           # U, l or C (marks are caseless)
        return "C" if K[p] == "M" else K[p]

    i = 0
    while i < n:
        if not is_word(i):
            i += 1
            continue
        s = i
        e = i
        while e < n and is_word(e):
            e += 1
        # first piece: starts at s unless the char before is a start that can be a prefix: whitespace other than
        # CR/LF (the last char of a whitespace run is always a start), or a single punctuation char that starts a piece
        has_prefix = s > 0 and ((K[s - 1] == "W") or (K[s - 1] == "O" and S[s - 1]))
        if not has_prefix:
            S[s] = True
        # later pieces start at upper-case letters only
        for j in range(s + 1, e):
            if kind(j) != "U":
                continue
            q = j - 1
            while q >= s and kind(q) == "C":
                q -= 1
            y_rule = q >= s and kind(q) == "l"
            tail_rule = kind(j - 1) == "C" and all(kind(t) == "U" for t in range(j, e))
            if y_rule or tail_rule:
                S[j] = True
        i = e
    return S


def check(n_strings=200000, seed=1, lengths=(1, 2, 3, 4, 5, 7, 9, 14, 25)):
    from oracle import tekken_oracle as TO
    orc = TO.OracleTekkenizer.from_file(TO.find_tekken_json())
    cl = Classifier()
    alphabet = list("aAbBzZǅʰ中あकाि्ً́̈ 　\t\n\r/\\.,;!?-'\"(0123٣①€😀‍") + ["\u00a0", "\u0085", "\u000b", "\u2028", "\u0301", "\u20dd", "/", "/", "\n", " ", "A", "a"]
    cache = {ch: cl.cls(ch) for ch in alphabet}
    rng = random.Random(seed)
    bad = 0
    for it in range(n_strings):
        text = "".join(rng.choice(alphabet) for _ in range(rng.choice(lengths)))
        K = [cache[ch] for ch in text]
        S = starts_by_runs(text, K)
        got = [i for i, v in enumerate(S) if v]
        pieces = orc.split_config(text)
        want, pos = [], 0
        for p in pieces:
            want.append(pos)
            pos += len(p.decode("utf-8"))
        if got != want:
            bad += 1
            if bad <= 5:
                print("MISMATCH", repr(text), "".join(K), "\n   predicate", got, "\n   oracle   ", want)
    print("%d strings, %d mismatches" % (n_strings, bad))
    return bad


if __name__ == "__main__":
    sys.exit(1 if check(int(sys.argv[1]) if len(sys.argv) > 1 else 200000) else 0)


def safe_starts(text, K):
    """Positions that are piece starts by a purely LOCAL rule (own class + previous class).  A first GPU split for the
    stored pattern can cut the text at these and let one lane run the sequential matcher over each segment."""
    n = len(K)
    out = []
    for i in range(n):
        k = K[i]
        p = K[i - 1] if i else None
        if i == 0:
            out.append(i)
        elif k == "N" or p == "N":
            out.append(i)                               # every digit is a piece; whatever follows a digit starts one
        elif k == "W" and p not in "RW":
            out.append(i)                               # a whitespace run that does not begin with CR/LF starts a piece
        elif k == "O" and p in "UlC":
            out.append(i)                               # punctuation after a letter is never absorbed
    return out


def check_safe(n_strings=100000, seed=3):
    from oracle import tekken_oracle as TO
    orc = TO.OracleTekkenizer.from_file(TO.find_tekken_json())
    cl = Classifier()
    alphabet = list("aAbBzZ\u01c5\u02b0\u4e2d\u3042\u0915\u093e\u093f\u094d\u064b\u0301\u0308 \u3000\t\n\r/\\.,;!?-'\"(0123\u0663\u2460\u20ac\U0001f600\u200d") + ["\u00a0", "\u0085", "/", "\n", " ", "A", "a"]
    cache = {ch: cl.cls(ch) for ch in alphabet}
    rng = random.Random(seed)
    bad = 0
    n_safe = n_all = 0
    for it in range(n_strings):
        text = "".join(rng.choice(alphabet) for _ in range(rng.choice([2, 3, 5, 9, 14, 25, 60])))
        K = [cache[ch] for ch in text]
        want, pos = set(), 0
        for p in orc.split_config(text):
            want.add(pos)
            pos += len(p.decode("utf-8"))
        got = safe_starts(text, K)
        n_safe += len(got)
        n_all += len(want)
        if not set(got) <= want:
            bad += 1
            if bad <= 5:
                print("NOT A START", repr(text), "".join(K), sorted(set(got) - want))
    print("%d strings, %d with a wrong safe start; %d of %d starts are safe starts" % (n_strings, bad, n_safe, n_all))
    return bad
