"""Research prototype (test infrastructure): an exact PARALLEL byte_pair_merge for long pieces that applies merges of
MANY different ranks per round -- the generalisation of "rank-rounds with hazard cut" (SURVEY.md Appendix C), which
handles one rank per round and degenerates to thousands of rounds on text without repetition (BASELINE config 4).

Round (compact arrays: id[i] = part i, rk[i] = rank of the pair (id[i], id[i+1]) or INF):
  1. candidates: pairs with rank <= T (T >= the smallest rank present).  Keys order them as the sequential loop would
     take them: (rank, position).
  2. selection = what the sequential loop would merge if no merge created a pair of rank <= T: going through the
     candidates by key, a candidate merges unless a neighbouring candidate (i-1 or i+1: they share a part with it) with
     a SMALLER key merged.  Decided by a few passes of a local rule; candidates still undecided after the passes cut
     the round at their key.
  3. every selected pair computes the two pairs its merge creates AT ITS TIME: its neighbours two positions away are
     already merged iff they are selected with a smaller key.  A created pair of rank <= T is a hazard: the sequential
     loop might take it before a later candidate, so the round applies the selected merges up to and including the
     smallest hazard key and drops the rest.
  4. apply, rebuild the compact arrays (the final rank between two adjacent merged parts is the one computed by the
     LATER of the two).
Checked here against the literal loop (the definition) on the Tekken vocabulary and on shuffled-rank synthetic
vocabularies where hazards are frequent.  oracle/research/ is not imported by anything in the product path."""
import random
import sys

INF = 1 << 40


def literal(parts, rank):
    """tiktoken's _byte_pair_merge: global minimum rank, leftmost on ties."""
    parts = list(parts)
    while len(parts) > 1:
        best, bi = INF, -1
        for i in range(len(parts) - 1):
            r = rank(parts[i], parts[i + 1])
            if r < best:
                best, bi = r, i
        if bi < 0:
            break
        parts[bi:bi + 2] = [best]
    return parts


def rounds(parts, rank, passes=3, stats=None):
    idv = list(parts)
    m = len(idv)
    rk = [rank(idv[i], idv[i + 1]) if i + 1 < m else INF for i in range(m)]
    delta = 0
    nrounds = 0
    while True:
        mn = min(rk) if rk else INF
        if mn >= INF:
            break
        nrounds += 1
        T = mn + delta
        m = len(idv)
        key = lambda i: (rk[i], i)
        cand = [rk[i] <= T for i in range(m)]
        # --- selection by passes of the local rule
        UND, SEL, NOT = 1, 2, 3
        st = [UND if cand[i] else 0 for i in range(m)]
        for _ in range(passes):
            new = list(st)
            for i in range(m):
                if st[i] != UND:
                    continue
                lower = [j for j in (i - 1, i + 1) if 0 <= j < m and cand[j] and key(j) < key(i)]
                if any(st[j] == SEL for j in lower):
                    new[i] = NOT
                elif all(st[j] == NOT for j in lower):
                    new[i] = SEL
            st = new
        cut = (INF, INF)            # keys >= cut are dropped
        for i in range(m):
            if st[i] == UND:
                cut = min(cut, key(i))
        # --- created pairs and hazards
        x = [INF] * m
        y = [INF] * m
        hazard = (INF, INF)
        for i in range(m):
            if st[i] != SEL or key(i) >= cut:
                continue
            N = rk[i]
            if i >= 1:
                left = rk[i - 2] if (i >= 2 and st[i - 2] == SEL and key(i - 2) < key(i)) else idv[i - 1]
                x[i] = rank(left, N)
            if i + 2 < m:
                right = rk[i + 2] if (st[i + 2] == SEL and key(i + 2) < key(i)) else idv[i + 2]
                y[i] = rank(N, right)
            if x[i] <= T or y[i] <= T:
                hazard = min(hazard, key(i))
        # apply: selected, key < cut, key <= hazard
        app = [st[i] == SEL and key(i) < cut and key(i) <= hazard for i in range(m)]
        napp = sum(app)
        nid, nrk = [], []
        i = 0
        while i < m:
            if app[i]:
                nid.append(rk[i])
                if i + 2 >= m:
                    r = INF
                elif app[i + 2]:
                    r = x[i + 2] if key(i + 2) > key(i) else y[i]
                else:
                    r = y[i]
                nrk.append(r)
                i += 2
            else:
                nid.append(idv[i])
                if i + 1 >= m:
                    r = INF
                elif app[i + 1]:
                    r = x[i + 1]
                else:
                    r = rk[i]
                nrk.append(r)
                i += 1
        nsel = sum(1 for i in range(m) if st[i] == SEL)
        idv, rk = nid, nrk
        # adapt the window of ranks: widen while rounds go through, narrow after a cut
        if napp == nsel and napp > 0:
            delta = max(1, delta * 2) if delta else 64
        else:
            delta //= 4
        if stats is not None:
            stats.append((m, napp))
    if stats is not None:
        stats.append(("rounds", nrounds))
    return idv


def main():
    sys.path.insert(0, ".")
    from oracle import tekken_oracle as TO
    from tekken_rs_b200 import assets
    orc = TO.OracleTekkenizer.from_file(assets.ensure_tekken_json())
    table = {b: r for r, b in enumerate(orc.ranks)}
    inv = orc.ranks

    def rank_tekken(a, b):
        return table.get(inv[a] + inv[b], INF)
    rng = random.Random(5)
    bad = 0
    tot_rounds, tot_merges = 0, 0
    for trial in range(300):
        kind = trial % 6
        n = rng.choice([5, 17, 64, 200, 700, 2000])
        if kind == 0:
            s = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(n)).encode()
        elif kind == 1:
            s = "".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(n // 3 + 1)).encode()
        elif kind == 2:
            s = "".join(rng.choice("abcdeКирилلعربية中文ñü") for _ in range(n // 2 + 1)).encode()
        elif kind == 3:
            s = (rng.choice(["ab", "a", "abc", "中", " "]) * n).encode()[:n + 3]
        elif kind == 4:
            s = bytes(rng.randrange(256) for _ in range(n))
        else:
            s = "".join(rng.choice(["the", "ing", "tion", "a", "e", "xq", "zz"]) for _ in range(n // 3 + 1)).encode()
        st = []
        got = rounds(list(s), rank_tekken, stats=st)
        want = literal(list(s), rank_tekken)
        tot_rounds += st[-1][1]
        tot_merges += len(s) - len(want)
        if got != want:
            bad += 1
            print("MISMATCH tekken", trial, len(s))
    print("tekken vocab: %d mismatches / 300; %d merges in %d rounds (%.1f per round)" % (bad, tot_merges, tot_rounds, tot_merges / max(1, tot_rounds)))
    # synthetic vocabularies with shuffled ranks: rank(merged) < rank(part) is common -> hazards and cuts are frequent
    bad = 0
    for v in range(40):
        vr = random.Random(100 + v)
        toks = set()
        while len(toks) < 400:
            toks.add(bytes(vr.choice(b"abcd") for _ in range(vr.randint(2, 5))))
        toks = list(toks)
        vr.shuffle(toks)
        tab = {bytes([i]): i for i in range(256)}
        for i, t in enumerate(toks):
            tab[t] = 256 + i
        invs = {r: b for b, r in tab.items()}
        rk = lambda a, b: tab.get(invs[a] + invs[b], INF)
        for trial in range(60):
            n = vr.choice([3, 9, 40, 65, 200, 600])
            s = bytes(vr.choice(b"abcd") for _ in range(n))
            if rounds(list(s), rk) != literal(list(s), rk):
                bad += 1
                print("MISMATCH synthetic", v, trial, s[:60])
    print("synthetic shuffled-rank vocabularies: %d mismatches / 2400" % bad)


if __name__ == "__main__":
    main()
