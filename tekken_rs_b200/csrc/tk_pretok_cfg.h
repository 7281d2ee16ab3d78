// tk_pretok_cfg.h -- the split for the pattern STORED in tekken.json (Mistral's own Tekken regex; the reference
// ignores it, src/tekkenizer.rs:74,123), used by handles created with TK_SPLIT_CONFIG:
//
//   [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+
//   | [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*
//   | \p{N} | ?[^\s\p{L}\p{N}]+[\r\n/]* | \s*[\r\n]+ | \s+(?!\S) | \s+
//
// Four purely local rules mark positions that are always piece starts ("safe starts": tk_cfg_safe_start /
// tk_cfg_safe_mask); between two safe starts one lane runs the sequential matcher (tk_cfg_match_end, tk_cfg_walk) --
// the lane-per-irregular-unit scheme of the merge stage.  Everything is __host__ __device__: cfg_mask_kernel and
// cfg_walk_kernel (tk_kernels.cu) run it on the GPU, tests/native/cfgsplit_host.cpp runs the same code on the CPU
// against the oracle (tests/test_cfgsplit_model.py).
#pragma once
#include "tk_common.h"
#include "tk_pretok.h"      // SWAR helpers, byte sources, TK_FFS

// classes of the stored pattern: U = Lu|Lt, LO = Ll, C = Lm|Lo (caseless letters), M = \p{M} (a word character, a
// punctuation character and a possible prefix at once), N = \p{N}, W = \s minus CR/LF, R = CR/LF, O = the rest
enum { TK_CC_O = 0, TK_CC_U = 1, TK_CC_LO = 2, TK_CC_C = 3, TK_CC_M = 4, TK_CC_N = 5, TK_CC_W = 6, TK_CC_R = 7 };

// two-stage table, 4 bits per code point: stage1[cp >> 7] -> block; a block holds 128 nibbles (64 bytes)
struct TkCfgTables {
    const uint16_t* stage1;
    const uint8_t* stage2;
};

// ASCII without the table (the matcher walks character by character: two dependent table loads per ASCII
// character were most of its time).  \s in ASCII is 9..13 and 32; tests/test_cfgsplit_model.py checks this
// function against the table for all 128 values.
TK_HD uint32_t tk_cfg_class_ascii(uint32_t c) {
    if (c - 0x61u < 26u) return TK_CC_LO;
    if (c - 0x41u < 26u) return TK_CC_U;
    if (c - 0x30u < 10u) return TK_CC_N;
    if (c == 0x0Au || c == 0x0Du) return TK_CC_R;
    if (c == 0x20u || c - 9u < 5u) return TK_CC_W;
    return TK_CC_O;
}

TK_HD uint32_t tk_cfg_class(const TkCfgTables& T, uint32_t cp) {
    if (cp < 0x80u) return tk_cfg_class_ascii(cp);
    const uint32_t blk = T.stage1[cp >> 7];
    const uint32_t b = T.stage2[blk * 64u + ((cp & 127u) >> 1)];
    return (b >> ((cp & 1u) * 4u)) & 15u;
}

TK_HD bool tk_cfg_in_U(uint32_t c) { return c == TK_CC_U || c == TK_CC_C || c == TK_CC_M; }    // [\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]
TK_HD bool tk_cfg_in_L(uint32_t c) { return c == TK_CC_LO || c == TK_CC_C || c == TK_CC_M; }   // [\p{Ll}\p{Lm}\p{Lo}\p{M}]
TK_HD bool tk_cfg_is_letter(uint32_t c) { return c == TK_CC_U || c == TK_CC_LO || c == TK_CC_C; }
TK_HD bool tk_cfg_is_ws(uint32_t c) { return c == TK_CC_W || c == TK_CC_R; }
// [^\s\p{L}\p{N}]: punctuation and marks
TK_HD bool tk_cfg_is_punct(uint32_t c) { return c == TK_CC_O || c == TK_CC_M; }
// [^\r\n\p{L}\p{N}]: what can be the optional first character of a word piece
TK_HD bool tk_cfg_is_prefix(uint32_t c) { return c == TK_CC_O || c == TK_CC_M || c == TK_CC_W; }

// Scalar value, class and length of the character whose lead byte is at pos (the text is valid UTF-8: K1 checks that).
template <class B>
TK_HD int tk_cfg_char(const B& src, int64_t pos, const TkCfgTables& T, uint32_t* cls, uint32_t* cp_out) {
    const uint32_t b0 = src.at(pos);
    uint32_t cp;
    int len;
    if (b0 < 0x80u) { cp = b0; len = 1; }
    else if (b0 < 0xE0u) { cp = ((b0 & 0x1Fu) << 6) | (src.at(pos + 1) & 0x3Fu); len = 2; }
    else if (b0 < 0xF0u) { cp = ((b0 & 0x0Fu) << 12) | ((src.at(pos + 1) & 0x3Fu) << 6) | (src.at(pos + 2) & 0x3Fu); len = 3; }
    else { cp = ((b0 & 0x07u) << 18) | ((src.at(pos + 1) & 0x3Fu) << 12) | ((src.at(pos + 2) & 0x3Fu) << 6) | (src.at(pos + 3) & 0x3Fu); len = 4; }
    *cls = tk_cfg_class(T, cp);
    *cp_out = cp;
    return len;
}

// Where a document ends.  The matcher never looks past it.  TkCfgEndAt: a known end offset (host tests).
// TkCfgEndMask: the end of the text or the next document-start bit (what a kernel has: the K0 bitmask), so a lane
// needs no search for its document's end.
struct TkCfgEndAt {
    int64_t end;
    TK_HD bool at_end(int64_t pos) const { return pos >= end; }
};
struct TkCfgEndMask {
    const uint32_t* ds_mask;     // bit = a document starts at this byte
    int64_t n;                   // end of the text
    int64_t own_start;           // the start of the walk's own document segment is not an end
    // the matcher asks about consecutive positions: the last mask word read is kept (one load per 32 bytes walked
    // instead of one per question)
    mutable int64_t word_idx = -1;
    mutable uint32_t word = 0;
    TK_HD TkCfgEndMask(const uint32_t* m, int64_t n_, int64_t own) : ds_mask(m), n(n_), own_start(own) {}
    TK_HD bool at_end(int64_t pos) const {
        if (pos >= n) return true;
        if (pos <= own_start) return false;
        const int64_t w = pos >> 5;
        if (w != word_idx) { word_idx = w; word = ds_mask[w]; }
        return (word >> (pos & 31)) & 1u;
    }
};

// U*L+ at s: end of the match, or -1.  Greedy U* takes the whole U run, L+ the L run that follows; if nothing in L
// follows, U* gives characters back until the one it gives back is in L: the match then ends after the LAST
// character of the U run that is in both sets (Lm, Lo, M).
template <class B, class E>
TK_HD int64_t tk_cfg_match_UL(const B& src, int64_t s, const E& end, const TkCfgTables& T) {
    int64_t q = s, last_both_end = -1;
    uint32_t c, cp;
    while (!end.at_end(q)) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_U(c)) break;
        q += l;
        if (tk_cfg_in_L(c)) last_both_end = q;
    }
    int64_t r = q;
    while (!end.at_end(r)) {
        const int l = tk_cfg_char(src, r, T, &c, &cp);
        if (!tk_cfg_in_L(c)) break;
        r += l;
    }
    return r > q ? r : last_both_end;
}

// U+L* at s: end of the match, or -1
template <class B, class E>
TK_HD int64_t tk_cfg_match_UpL(const B& src, int64_t s, const E& end, const TkCfgTables& T) {
    int64_t q = s;
    uint32_t c, cp;
    while (!end.at_end(q)) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_U(c)) break;
        q += l;
    }
    if (q == s) return -1;
    while (!end.at_end(q)) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_L(c)) break;
        q += l;
    }
    return q;
}

// End of the leftmost-first match of the stored pattern that starts at q (q < end, q is a character boundary).
// `end` says where the document ends: matches never cross it.
template <class B, class E>
TK_HD int64_t tk_cfg_match_end(const B& src, int64_t q, const E& end, const TkCfgTables& T) {
    uint32_t k0, c0, k1 = 0xFFu, c1 = 0;
    const int l0 = tk_cfg_char(src, q, T, &k0, &c0);
    const bool has1 = !end.at_end(q + l0);
    if (has1) tk_cfg_char(src, q + l0, T, &k1, &c1);
    int64_t e;
    // B1: P?U*L+ (greedy optional prefix first, then without it), B2: P?U+L*.  Which of the four attempts can match at
    // all follows from the first two classes: a word alternative needs a word character (U, Ll, C or M) at q, or a
    // prefix character at q and a word character after it -- so a digit, and punctuation or whitespace that no word
    // character follows, skip all four, and only a mark (prefix AND word character) has to try them all.
    const bool w0 = tk_cfg_in_U(k0) || tk_cfg_in_L(k0);
    const bool w1 = has1 && (tk_cfg_in_U(k1) || tk_cfg_in_L(k1));
    const bool pre = tk_cfg_is_prefix(k0) && w1;
    if (pre && (e = tk_cfg_match_UL(src, q + l0, end, T)) >= 0) return e;
    if (w0 && (e = tk_cfg_match_UL(src, q, end, T)) >= 0) return e;
    if (pre && (e = tk_cfg_match_UpL(src, q + l0, end, T)) >= 0) return e;
    if (w0 && (e = tk_cfg_match_UpL(src, q, end, T)) >= 0) return e;
    // B3: \p{N}
    if (k0 == TK_CC_N) return q + l0;
    // B4:  ?[^\s\p{L}\p{N}]+[\r\n/]*
    {
        int64_t s = -1;
        if (c0 == 0x20u && has1 && tk_cfg_is_punct(k1)) s = q + l0;
        else if (tk_cfg_is_punct(k0)) s = q;
        if (s >= 0) {
            uint32_t c, cp;
            while (!end.at_end(s)) {
                const int l = tk_cfg_char(src, s, T, &c, &cp);
                if (!tk_cfg_is_punct(c)) break;
                s += l;
            }
            while (!end.at_end(s)) {
                const uint32_t b = src.at(s);
                if (b != 0x0Au && b != 0x0Du && b != 0x2Fu) break;
                ++s;
            }
            return s;
        }
    }
    // whitespace run [q, e2): B5 \s*[\r\n]+ backs off to the last CR/LF; B6 \s+(?!\S) leaves the last char unless the
    // document ends; B7 \s+
    {
        int64_t e2 = q, last_r_end = -1, last_char = q;
        uint32_t c, cp;
        while (!end.at_end(e2)) {
            const int l = tk_cfg_char(src, e2, T, &c, &cp);
            if (!tk_cfg_is_ws(c)) break;
            last_char = e2;
            e2 += l;
            if (c == TK_CC_R) last_r_end = e2;
        }
        if (last_r_end >= 0) return last_r_end;
        if (end.at_end(e2)) return e2;
        if (last_char > q) return last_char;
        return e2;
    }
}

// Purely local rules that only ever mark real piece starts (verified against the oracle,
// oracle/research/config_pattern_positionwise.py: safe_starts): the first character of a document; a digit; whatever
// follows a digit; whitespace other than CR/LF after non-whitespace; punctuation (not a mark) after a letter.
TK_HD bool tk_cfg_safe_start(bool doc_start, uint32_t prev, uint32_t cur) {
    if (doc_start) return true;
    if (cur == TK_CC_N || prev == TK_CC_N) return true;
    if (cur == TK_CC_W && !tk_cfg_is_ws(prev)) return true;
    if (cur == TK_CC_O && tk_cfg_is_letter(prev)) return true;
    return false;
}

// ---- the same rules as bit logic over a 32-byte window (what K1 would run: one thread per window) ------------
// Class bits are spread over all bytes of a char (as in TkWin), so `mask << 1` with a carry from the previous window
// is "class of the previous char" at a lead byte.
struct TkCfgWin {
    uint32_t lead;                       // byte starts a char
    uint32_t mU, mLO, mC, mM, mN, mW, mR;  // classes; O = none of them
    uint32_t ds;                         // a document starts at this byte
    uint32_t bad;                        // invalid UTF-8 detected at this byte
};

// strict decode of the char whose lead byte is at pos: length (0 = invalid UTF-8) and class
template <class B>
TK_HD int tk_cfg_decode_strict(const B& src, int64_t pos, const TkCfgTables& T, uint32_t* cls) {
    const uint32_t b0 = src.at(pos);
    uint32_t cp;
    int len;
    if (b0 < 0x80u) { cp = b0; len = 1; }
    else if (b0 < 0xC2u) return 0;
    else if (b0 < 0xE0u) {
        const uint32_t b1 = src.at(pos + 1);
        if ((b1 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
        len = 2;
    } else if (b0 < 0xF0u) {
        const uint32_t b1 = src.at(pos + 1), b2 = src.at(pos + 2);
        if ((b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
        if (cp < 0x800u || (cp >= 0xD800u && cp <= 0xDFFFu)) return 0;
        len = 3;
    } else if (b0 < 0xF5u) {
        const uint32_t b1 = src.at(pos + 1), b2 = src.at(pos + 2), b3 = src.at(pos + 3);
        if ((b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u || (b3 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
        if (cp < 0x10000u || cp > 0x10FFFFu) return 0;
        len = 4;
    } else return 0;
    if (src.past_end(pos + len)) return 0;
    *cls = tk_cfg_class(T, cp);
    return len;
}

TK_HD void tk_cfg_set(TkCfgWin& r, uint32_t cls, uint32_t m) {
    if (cls == TK_CC_U) r.mU |= m;
    else if (cls == TK_CC_LO) r.mLO |= m;
    else if (cls == TK_CC_C) r.mC |= m;
    else if (cls == TK_CC_M) r.mM |= m;
    else if (cls == TK_CC_N) r.mN |= m;
    else if (cls == TK_CC_W) r.mW |= m;
    else if (cls == TK_CC_R) r.mR |= m;
}

// Classify the 32-byte window at byte offset `pos` (a multiple of 32).  w[0..7] = its bytes as little-endian words,
// zero padded beyond the text.
template <class B>
TK_HD TkCfgWin tk_cfg_classify_window(const B& src, uint64_t pos, const uint32_t* w, uint32_t ds_word, const TkCfgTables& T) {
    TkCfgWin r;
    r.mU = r.mLO = r.mC = r.mM = r.mN = r.mW = r.mR = 0;
    uint32_t hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t x = w[j];
        const uint32_t h = x & TK_H;
        const uint32_t w7 = x & 0x7F7F7F7Fu;
        const uint32_t ok = ~h;
        const uint32_t upper = tk_swar_ge(w7, 0x41u) & tk_swar_le(w7, 0x5Au) & ok;
        const uint32_t lower = tk_swar_ge(w7, 0x61u) & tk_swar_le(w7, 0x7Au) & ok;
        const uint32_t digit = tk_swar_ge(w7, 0x30u) & tk_swar_le(w7, 0x39u) & ok;
        const uint32_t c9_13 = tk_swar_ge(w7, 0x09u) & tk_swar_le(w7, 0x0Du) & ok;
        const uint32_t rr = (tk_swar_eq(w7, 0x0Au) | tk_swar_eq(w7, 0x0Du)) & ok;
        const uint32_t e20 = tk_swar_eq(w7, 0x20u) & ok;
        const int s = 4 * j;
        r.mU |= tk_swar_nib(upper) << s;
        r.mLO |= tk_swar_nib(lower) << s;
        r.mN |= tk_swar_nib(digit) << s;
        r.mR |= tk_swar_nib(rr) << s;
        r.mW |= tk_swar_nib((c9_13 & ~rr) | e20) << s;
        hi |= tk_swar_nib(h) << s;
    }
    uint32_t lead = 0xFFFFFFFFu, bad = 0;
    if (hi) {
        uint32_t cont = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) cont |= tk_swar_nib((w[j] & TK_H) & ~((w[j] << 1) & TK_H)) << (4 * j);
        lead = ~cont;
        uint32_t covered = ~hi;
        uint32_t lc = (uint32_t)(TK_FFS(~cont) - 1);          // leading continuation bytes: a char that starts in the previous window
        if (cont == 0xFFFFFFFFu) lc = 32;
        if (lc > 0) {
            int back = 0;
            int64_t q = (int64_t)pos - 1;
            while (back < 3 && q >= 0 && (src.at(q) & 0xC0u) == 0x80u) { --q; ++back; }
            uint32_t cls = TK_CC_O;
            const int len = (q >= 0) ? tk_cfg_decode_strict(src, q, T, &cls) : 0;
            const int64_t over = len ? q + len - (int64_t)pos : 0;
            if (over > 0) {
                const uint32_t m = (1u << (over < (int64_t)lc ? (uint32_t)over : lc)) - 1u;
                covered |= m;
                tk_cfg_set(r, cls, m);
            }
        }
        uint32_t todo = hi & lead;
        while (todo) {
            const int i = TK_FFS(todo) - 1;
            todo &= todo - 1;
            uint32_t cls = TK_CC_O;
            const int len = tk_cfg_decode_strict(src, (int64_t)pos + i, T, &cls);
            if (len == 0) { bad |= 1u << i; continue; }
            const uint32_t m = (len >= 32 - i) ? (0xFFFFFFFFu << i) : (((1u << len) - 1u) << i);
            covered |= m;
            tk_cfg_set(r, cls, m);
        }
        bad |= ~covered;
    }
    bad |= ds_word & ~lead;
    r.lead = lead; r.ds = ds_word; r.bad = bad;
    return r;
}

// Safe starts of window c (p = the window before it; all-zero masks for the window before the text)
TK_HD uint32_t tk_cfg_safe_mask(const TkCfgWin& p, const TkCfgWin& c) {
    const uint32_t P_N = tk_shl(c.mN, p.mN, 1), P_W = tk_shl(c.mW, p.mW, 1), P_R = tk_shl(c.mR, p.mR, 1);
    const uint32_t P_letter = tk_shl(c.mU | c.mLO | c.mC, p.mU | p.mLO | p.mC, 1);
    const uint32_t mO = ~(c.mU | c.mLO | c.mC | c.mM | c.mN | c.mW | c.mR);
    return c.lead & (c.ds | c.mN | P_N | (c.mW & ~(P_W | P_R)) | (mO & P_letter));
}

// The text as a kernel's walk sees it: a shared-memory copy of the tile around the walk's start (zero outside the
// text), the text itself beyond it (a walk is as long as the distance to the next safe start: usually a word).
struct TkBytesTileOrGlobal {
    const uint8_t* sm;        // sm[pos - origin] for lo <= pos < hi
    int64_t origin, lo, hi;
    const uint8_t* data;
    uint64_t n;
    TK_HD uint32_t at(int64_t pos) const {
        if (pos >= lo && pos < hi) return (uint32_t)sm[pos - origin];
        return (pos >= 0 && (uint64_t)pos < n) ? (uint32_t)data[pos] : 0u;
    }
    TK_HD bool past_end(int64_t end) const { return (uint64_t)end > n; }
};
// a bitmask (document starts, safe starts) with the words [w_lo, w_hi) staged in shared memory
struct TkMaskTileOrGlobal {
    const uint32_t* sm;       // sm[w - w_lo]
    int64_t w_lo, w_hi;
    const uint32_t* g;
    TK_HD uint32_t bit(int64_t pos) const {
        const int64_t w = pos >> 5;
        const uint32_t m = (w >= w_lo && w < w_hi) ? sm[w - w_lo] : g[w];
        return (m >> (pos & 31)) & 1u;
    }
};
struct TkCfgEndMaskTile {
    TkMaskTileOrGlobal ds;
    int64_t n, own_start;
    TK_HD bool at_end(int64_t pos) const { return pos >= n || (pos > own_start && ds.bit(pos)); }
};

// The walk in its general form: `stop` says where the document ends, `is_safe(pos)` whether pos is a safe start.
template <class B, class E, class S, class F>
TK_HD void tk_cfg_walk_from(const B& src, int64_t q0, const S& is_safe, const E& stop, int64_t n, const TkCfgTables& T, F set_bit) {
    int64_t q = q0;
    for (;;) {
        const int64_t e = tk_cfg_match_end(src, q, stop, T);
        if (e >= n || is_safe(e)) return;                                   // the next segment's lane takes over
        set_bit(e);
        q = e;
    }
}

// What one lane of the walk kernel does for one safe start at byte q0: run the matcher from piece to piece and mark
// every piece start until the walk reaches the next safe start (its bit is set in safe_mask), the next document or
// the end of the text.  set_bit(pos) records a piece start (an atomicOr into the piece-start mask in a kernel).
template <class B, class F>
TK_HD void tk_cfg_walk(const B& src, int64_t q0, const uint32_t* safe_mask, const uint32_t* ds_mask, int64_t n,
                       const TkCfgTables& T, F set_bit) {
    const TkCfgEndMask stop{ds_mask, n, q0};
    int64_t q = q0;
    for (;;) {
        const int64_t e = tk_cfg_match_end(src, q, stop, T);
        if (e >= n || ((safe_mask[e >> 5] >> (e & 31)) & 1u)) return;      // the next segment's lane takes over
        set_bit(e);
        q = e;
    }
}
