// tk_pretok_cfg.h -- GROUNDWORK (not yet part of the library): the split for the pattern STORED in
// tekken.json (Mistral's own Tekken regex; the reference ignores it, src/tekkenizer.rs:74,123):
//
//   [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+
//   | [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*
//   | \p{N} | ?[^\s\p{L}\p{N}]+[\r\n/]* | \s*[\r\n]+ | \s+(?!\S) | \s+
//
// Plan: four purely local rules mark positions that are always piece starts ("safe starts":
// tk_cfg_safe_start); between two safe starts one lane runs the sequential matcher (tk_cfg_match_end) -- the
// lane-per-irregular-unit scheme of the merge stage.  Both functions are __host__ __device__;
// tests/native/cfgsplit_host.cpp runs them on the CPU against the oracle (tests/test_cfgsplit_model.py).
// No kernel includes this header yet.
#pragma once
#include "tk_common.h"

// classes of the stored pattern: U = Lu|Lt, LO = Ll, C = Lm|Lo (caseless letters), M = \p{M} (a word character, a
// punctuation character and a possible prefix at once), N = \p{N}, W = \s minus CR/LF, R = CR/LF, O = the rest
enum { TK_CC_O = 0, TK_CC_U = 1, TK_CC_LO = 2, TK_CC_C = 3, TK_CC_M = 4, TK_CC_N = 5, TK_CC_W = 6, TK_CC_R = 7 };

// two-stage table, 4 bits per code point: stage1[cp >> 7] -> block; a block holds 128 nibbles (64 bytes)
struct TkCfgTables {
    const uint16_t* stage1;
    const uint8_t* stage2;
};

TK_HD uint32_t tk_cfg_class(const TkCfgTables& T, uint32_t cp) {
    if (cp == 0x0Au || cp == 0x0Du) return TK_CC_R;
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

// U*L+ at s: end of the match, or -1.  Greedy U* takes the whole U run, L+ the L run that follows; if nothing in L
// follows, U* gives characters back until the one it gives back is in L: the match then ends after the LAST
// character of the U run that is in both sets (Lm, Lo, M).
template <class B>
TK_HD int64_t tk_cfg_match_UL(const B& src, int64_t s, int64_t end, const TkCfgTables& T) {
    int64_t q = s, last_both_end = -1;
    uint32_t c, cp;
    while (q < end) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_U(c)) break;
        q += l;
        if (tk_cfg_in_L(c)) last_both_end = q;
    }
    int64_t r = q;
    while (r < end) {
        const int l = tk_cfg_char(src, r, T, &c, &cp);
        if (!tk_cfg_in_L(c)) break;
        r += l;
    }
    return r > q ? r : last_both_end;
}

// U+L* at s: end of the match, or -1
template <class B>
TK_HD int64_t tk_cfg_match_UpL(const B& src, int64_t s, int64_t end, const TkCfgTables& T) {
    int64_t q = s;
    uint32_t c, cp;
    while (q < end) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_U(c)) break;
        q += l;
    }
    if (q == s) return -1;
    while (q < end) {
        const int l = tk_cfg_char(src, q, T, &c, &cp);
        if (!tk_cfg_in_L(c)) break;
        q += l;
    }
    return q;
}

// End of the leftmost-first match of the stored pattern that starts at q (q < end, q is a character boundary).
// `end` is the end of the document: matches never cross it.
template <class B>
TK_HD int64_t tk_cfg_match_end(const B& src, int64_t q, int64_t end, const TkCfgTables& T) {
    uint32_t k0, c0, k1 = 0xFFu, c1 = 0;
    const int l0 = tk_cfg_char(src, q, T, &k0, &c0);
    const bool has1 = q + l0 < end;
    if (has1) tk_cfg_char(src, q + l0, T, &k1, &c1);
    int64_t e;
    // B1: P?U*L+ (greedy optional prefix first, then without it), B2: P?U+L*
    if (tk_cfg_is_prefix(k0) && has1 && (e = tk_cfg_match_UL(src, q + l0, end, T)) >= 0) return e;
    if ((e = tk_cfg_match_UL(src, q, end, T)) >= 0) return e;
    if (tk_cfg_is_prefix(k0) && has1 && (e = tk_cfg_match_UpL(src, q + l0, end, T)) >= 0) return e;
    if ((e = tk_cfg_match_UpL(src, q, end, T)) >= 0) return e;
    // B3: \p{N}
    if (k0 == TK_CC_N) return q + l0;
    // B4:  ?[^\s\p{L}\p{N}]+[\r\n/]*
    {
        int64_t s = -1;
        if (c0 == 0x20u && has1 && tk_cfg_is_punct(k1)) s = q + l0;
        else if (tk_cfg_is_punct(k0)) s = q;
        if (s >= 0) {
            uint32_t c, cp;
            while (s < end) {
                const int l = tk_cfg_char(src, s, T, &c, &cp);
                if (!tk_cfg_is_punct(c)) break;
                s += l;
            }
            while (s < end) {
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
        while (e2 < end) {
            const int l = tk_cfg_char(src, e2, T, &c, &cp);
            if (!tk_cfg_is_ws(c)) break;
            last_char = e2;
            e2 += l;
            if (c == TK_CC_R) last_r_end = e2;
        }
        if (last_r_end >= 0) return last_r_end;
        if (e2 == end) return e2;
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
