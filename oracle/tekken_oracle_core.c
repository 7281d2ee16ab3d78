/*
 * tekken_oracle_core.c -- CPU restatement of the reference's text encode/decode arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file; the CUDA library never links or calls it.
 *
 * What it restates (citations are into /root/reference unless noted):
 *   - Tekkenizer::encode glue: +num_special_tokens, BOS/EOS insertion   src/tekkenizer.rs:378-405
 *   - Tekkenizer::decode_all / decode_group: special/ordinary run split,
 *     SpecialTokenPolicy, strict-UTF-8 per ordinary run                  src/tekkenizer.rs:463-560
 *   - the engine those wrap: tiktoken-rs 0.7.0 `CoreBPE` (Cargo.toml:40; NOT vendored under
 *     /root/reference), constructed with the hard-coded pattern of src/tekkenizer.rs:123
 *       (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}|
 *        ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
 *     Its published algorithm (openai/tiktoken src/lib.rs, which tiktoken-rs vendors) is:
 *       for every leftmost-first regex match (a "piece"): if the piece is a vocabulary entry emit
 *       its rank, else byte_pair_merge: start from single bytes, repeatedly merge the adjacent
 *       pair whose concatenated BYTES have the lowest rank (leftmost on ties) until no adjacent
 *       pair is in the vocabulary, then emit the rank of every part.
 *
 * Parity pinning: checked in tests/ against the reference's 20 golden encode vectors
 * (tests/test_tokenizer_output.rs:22-373), its decode vectors, and against the upstream engine
 * itself (Python `tiktoken` 0.12.0 = the same Rust CoreBPE) run in the build container with the
 * same pattern and ranks (fixtures under tests/golden/, generator committed).
 *
 * The regex is restated as a literal per-alternative matcher (match_at) with the backtracking
 * outcomes written out; Unicode classes come from tables measured from the engine
 * (oracle/tools/extract_unicode_tables.py -> unicode_ranges.h).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "unicode_ranges.h"
#include "unicode_subclasses.h"

#define ORC_INF 0xFFFFFFFFu

enum { CL_O = 0, CL_L = 1, CL_N = 2, CL_R = 3, CL_W = 4 };

static uint8_t g_cls[0x110000];
static uint8_t g_sub[0x110000];      /* sub-class for the pattern stored in tekken.json: SUB_* */
static int g_cls_ready = 0;

enum { SUB_NONE = 0, SUB_UPPER = 1, SUB_LOWER = 2, SUB_BOTH = 3, SUB_MARK = 4 };

static void fill_ranges(const uint32_t (*r)[2], int n, uint8_t v) {
    for (int i = 0; i < n; i++)
        for (uint32_t c = r[i][0]; c <= r[i][1]; c++) g_cls[c] = v;
}

static void init_classes(void) {
    if (g_cls_ready) return;
    memset(g_cls, CL_O, sizeof g_cls);
    fill_ranges(UNI_L_RANGES, UNI_L_COUNT, CL_L);
    fill_ranges(UNI_N_RANGES, UNI_N_COUNT, CL_N);
    fill_ranges(UNI_S_RANGES, UNI_S_COUNT, CL_W);
    g_cls['\r'] = CL_R;
    g_cls['\n'] = CL_R;
    memset(g_sub, SUB_NONE, sizeof g_sub);
#define FILL_SUB(T, v) for (size_t i = 0; i < sizeof(T) / sizeof(T[0]); i++) for (uint32_t c = T[i][0]; c <= T[i][1]; c++) g_sub[c] = v
    FILL_SUB(ORC_SUB_UPPER, SUB_UPPER);
    FILL_SUB(ORC_SUB_LOWER, SUB_LOWER);
    FILL_SUB(ORC_SUB_BOTH, SUB_BOTH);
    FILL_SUB(ORC_SUB_MARK, SUB_MARK);
#undef FILL_SUB
    g_cls_ready = 1;
}

/* ------------------------------------------------------------------ vocabulary */

typedef struct {
    uint32_t n_vocab;          /* inner vocabulary size (ranks 0..n_vocab-1)         */
    uint32_t num_special;      /* ids below this are special                          */
    const uint8_t *bytes;      /* concatenated token bytes (owned copy)               */
    const uint64_t *off;       /* n_vocab+1 offsets into bytes (owned copy)           */
    /* open-addressing map bytes -> rank */
    uint32_t cap;              /* power of two                                        */
    uint32_t *slot_rank;       /* ORC_INF = empty                                     */
    uint64_t *slot_hash;
    /* special token strings for decode(Keep): positional index == id                */
    const uint8_t *sp_bytes;
    const uint64_t *sp_off;    /* num_special+1                                       */
    int bos_id, eos_id;        /* -1 when "<s>" / "</s>" are absent                   */
} orc_t;

static uint64_t hash_bytes(const uint8_t *p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

static uint32_t vocab_get(const orc_t *o, const uint8_t *p, size_t n) {
    uint64_t h = hash_bytes(p, n);
    uint32_t i = (uint32_t)h & (o->cap - 1);
    for (;;) {
        uint32_t r = o->slot_rank[i];
        if (r == ORC_INF) return ORC_INF;
        if (o->slot_hash[i] == h) {
            uint64_t a = o->off[r], b = o->off[r + 1];
            if (b - a == n && memcmp(o->bytes + a, p, n) == 0) return r;
        }
        i = (i + 1) & (o->cap - 1);
    }
}

orc_t *orc_new(const uint8_t *bytes, const uint64_t *off, uint32_t n_vocab, uint32_t num_special,
               const uint8_t *sp_bytes, const uint64_t *sp_off, int bos_id, int eos_id) {
    init_classes();
    orc_t *o = (orc_t *)calloc(1, sizeof *o);
    o->n_vocab = n_vocab;
    o->num_special = num_special;
    uint8_t *b = (uint8_t *)malloc(off[n_vocab] ? off[n_vocab] : 1);
    memcpy(b, bytes, off[n_vocab]);
    uint64_t *f = (uint64_t *)malloc((n_vocab + 1) * sizeof *f);
    memcpy(f, off, (n_vocab + 1) * sizeof *f);
    o->bytes = b; o->off = f;
    uint8_t *sb = (uint8_t *)malloc(sp_off[num_special] ? sp_off[num_special] : 1);
    memcpy(sb, sp_bytes, sp_off[num_special]);
    uint64_t *sf = (uint64_t *)malloc((num_special + 1) * sizeof *sf);
    memcpy(sf, sp_off, (num_special + 1) * sizeof *sf);
    o->sp_bytes = sb; o->sp_off = sf;
    o->bos_id = bos_id; o->eos_id = eos_id;
    uint32_t cap = 1024;
    while (cap < 4u * n_vocab) cap <<= 1;
    o->cap = cap;
    o->slot_rank = (uint32_t *)malloc(cap * sizeof(uint32_t));
    o->slot_hash = (uint64_t *)malloc(cap * sizeof(uint64_t));
    memset(o->slot_rank, 0xFF, cap * sizeof(uint32_t));
    /* FxHashMap::insert semantics (src/tekkenizer.rs:801): a later duplicate byte string
       overwrites the earlier rank. */
    for (uint32_t r = 0; r < n_vocab; r++) {
        const uint8_t *p = b + f[r];
        size_t n = (size_t)(f[r + 1] - f[r]);
        uint64_t h = hash_bytes(p, n);
        uint32_t i = (uint32_t)h & (cap - 1);
        for (;;) {
            uint32_t q = o->slot_rank[i];
            if (q == ORC_INF) { o->slot_rank[i] = r; o->slot_hash[i] = h; break; }
            if (o->slot_hash[i] == h && f[q + 1] - f[q] == n && memcmp(b + f[q], p, n) == 0) {
                o->slot_rank[i] = r; break;
            }
            i = (i + 1) & (cap - 1);
        }
    }
    return o;
}

void orc_free(orc_t *o) {
    if (!o) return;
    free((void *)o->bytes); free((void *)o->off);
    free((void *)o->sp_bytes); free((void *)o->sp_off);
    free(o->slot_rank); free(o->slot_hash); free(o);
}

/* ------------------------------------------------------------------ UTF-8 */

/* Strict UTF-8 decode of one scalar value at p (end = one past the text).  Returns the
   sequence length, or 0 if the bytes at p are not a valid encoding (Rust `str` invariant /
   String::from_utf8 rules: no overlongs, no surrogates, <= U+10FFFF). */
static int utf8_decode(const uint8_t *p, const uint8_t *end, uint32_t *cp) {
    uint8_t b0 = p[0];
    if (b0 < 0x80) { *cp = b0; return 1; }
    if (b0 < 0xC2) return 0;
    if (b0 < 0xE0) {
        if (end - p < 2 || (p[1] & 0xC0) != 0x80) return 0;
        *cp = ((uint32_t)(b0 & 0x1F) << 6) | (p[1] & 0x3F);
        return 2;
    }
    if (b0 < 0xF0) {
        if (end - p < 3 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80) return 0;
        uint32_t c = ((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
        if (c < 0x800 || (c >= 0xD800 && c <= 0xDFFF)) return 0;
        *cp = c;
        return 3;
    }
    if (b0 < 0xF5) {
        if (end - p < 4 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80)
            return 0;
        uint32_t c = ((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) |
                     ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
        if (c < 0x10000 || c > 0x10FFFF) return 0;
        *cp = c;
        return 4;
    }
    return 0;
}

/* Returns 1 if [p, p+n) is valid UTF-8. */
int orc_utf8_valid(const uint8_t *p, uint64_t n) {
    const uint8_t *end = p + n;
    while (p < end) {
        uint32_t cp;
        int l = utf8_decode(p, end, &cp);
        if (!l) return 0;
        p += l;
    }
    return 1;
}

/* ------------------------------------------------------------------ regex split */

typedef struct { const uint8_t *p; const uint8_t *end; } cur_t;

/* class and length of the char at q (q < end, text is valid UTF-8) */
static inline int char_at(const uint8_t *q, const uint8_t *end, uint32_t *cp, int *cls) {
    int l = utf8_decode(q, end, cp);
    *cls = g_cls[*cp];
    return l;
}

static inline int is_ws(int c) { return c == CL_R || c == CL_W; }

/* Unicode simple case folding restricted to the contraction letters: the only non-ASCII
   scalar value that folds onto one of s,t,r,e,v,m,l,d is U+017F (long s) -> 's'
   (measured from the engine: oracle/unicode_tables.json "fold"). */
static inline uint32_t fold_letter(uint32_t cp) {
    if (cp >= 'A' && cp <= 'Z') return cp + 32;
    if (cp == 0x17F) return 's';
    return cp;
}

/* Length in bytes of the leftmost-first match of the pattern at q (q < end).  Every position
   of a valid text is matched by some alternative, so the result is >= 1. */
static size_t match_at(const uint8_t *q, const uint8_t *end) {
    uint32_t c0, c1 = 0, c2 = 0;
    int k0, k1 = -1, k2 = -1;
    int l0 = char_at(q, end, &c0, &k0), l1 = 0, l2 = 0;
    if (q + l0 < end) {
        l1 = char_at(q + l0, end, &c1, &k1);
        if (q + l0 + l1 < end) l2 = char_at(q + l0 + l1, end, &c2, &k2);
    }
    /* A1: (?i:'s|'t|'re|'ve|'m|'ll|'d) */
    if (c0 == '\'' && l1) {
        uint32_t f1 = fold_letter(c1), f2 = l2 ? fold_letter(c2) : 0;
        if (f1 == 's' || f1 == 't' || f1 == 'm' || f1 == 'd') return (size_t)(l0 + l1);
        if (l2 && ((f1 == 'r' && f2 == 'e') || (f1 == 'v' && f2 == 'e') || (f1 == 'l' && f2 == 'l')))
            return (size_t)(l0 + l1 + l2);
    }
    /* A2: [^\r\n\p{L}\p{N}]?\p{L}+   (greedy optional prefix first, then without it) */
    {
        const uint8_t *s = 0;
        if (k0 != CL_R && k0 != CL_L && k0 != CL_N && k1 == CL_L) s = q + l0;
        else if (k0 == CL_L) s = q;
        if (s) {
            while (s < end) {
                uint32_t c; int k;
                int l = char_at(s, end, &c, &k);
                if (k != CL_L) break;
                s += l;
            }
            return (size_t)(s - q);
        }
    }
    /* A3: \p{N}{1,3} */
    if (k0 == CL_N) {
        size_t n = (size_t)l0;
        if (k1 == CL_N) { n += (size_t)l1; if (k2 == CL_N) n += (size_t)l2; }
        return n;
    }
    /* A4:  ?[^\s\p{L}\p{N}]+[\r\n]* */
    {
        const uint8_t *s = 0;
        if (c0 == ' ' && k1 == CL_O) s = q + l0;
        else if (k0 == CL_O) s = q;
        if (s) {
            while (s < end) {
                uint32_t c; int k;
                int l = char_at(s, end, &c, &k);
                if (k != CL_O) break;
                s += l;
            }
            while (s < end && (*s == '\r' || *s == '\n')) s++;
            return (size_t)(s - q);
        }
    }
    /* here the char at q is whitespace: find the maximal \s run [q, e) */
    {
        const uint8_t *e = q, *last_r_end = 0, *last_char = q;
        while (e < end) {
            uint32_t c; int k;
            int l = char_at(e, end, &c, &k);
            if (!is_ws(k)) break;
            last_char = e;
            e += l;
            if (k == CL_R) last_r_end = e;
        }
        /* A5: \s*[\r\n]+  -- greedy \s* backs off to the last CR/LF of the run */
        if (last_r_end) return (size_t)(last_r_end - q);
        /* A6: \s+(?!\S)  -- whole run at end of text, else all but its last char */
        if (e == end) return (size_t)(e - q);
        if (last_char > q) return (size_t)(last_char - q);
        /* A7: \s+ */
        return (size_t)(e - q);
    }
}

/* Piece end offsets of one text.  ends[] needs room for n entries at most.  Returns the
   number of pieces, or -1 if the text is not valid UTF-8 (the reference takes &str). */
int64_t orc_split(const uint8_t *text, uint64_t n, uint64_t *ends) {
    init_classes();
    if (!orc_utf8_valid(text, n)) return -1;
    const uint8_t *q = text, *end = text + n;
    int64_t k = 0;
    while (q < end) {
        q += match_at(q, end);
        ends[k++] = (uint64_t)(q - text);
    }
    return k;
}

/* ------------------------------------------------------------------ the pattern stored in tekken.json
 *
 * GROUNDWORK for SURVEY section 8(f) rank 1 ("config-driven pattern").  The reference IGNORES config.pattern
 * (src/tekkenizer.rs:74,123); Mistral's own tokenizer (mistral_common) compiles it:
 *   [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|
 *   [^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|
 *   \p{N}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+
 * Restated here as a literal matcher with the backtracking outcomes written out, pinned against the engine in
 * tests/test_oracle.py.  U = {Lu,Lt,Lm,Lo,M} ("upper or caseless"), L = {Ll,Lm,Lo,M} ("lower or caseless");
 * Lm, Lo and M are in both sets, and M is neither \p{L} nor \p{N} (it can be a prefix or punctuation char).
 */
static inline int in_U(int sub) { return sub == SUB_UPPER || sub == SUB_BOTH || sub == SUB_MARK; }
static inline int in_L(int sub) { return sub == SUB_LOWER || sub == SUB_BOTH || sub == SUB_MARK; }

/* U*L+ at s: end of the match, or 0.  Greedy U* takes the whole U run; L+ then takes the L run that follows.  If
   nothing in L follows, U* gives characters back until the one it gives back is in L (Lm, Lo or M): the match
   ends after the LAST character of the U run that is in both sets (what follows it in the run is Lu/Lt). */
static const uint8_t *match_UL(const uint8_t *s, const uint8_t *end) {
    const uint8_t *q = s, *last_both_end = 0;
    while (q < end) {
        uint32_t c;
        int l = utf8_decode(q, end, &c);
        int sub = g_sub[c];
        if (!in_U(sub)) break;
        q += l;
        if (in_L(sub)) last_both_end = q;
    }
    const uint8_t *r = q;
    while (r < end) {
        uint32_t c;
        int l = utf8_decode(r, end, &c);
        if (!in_L(g_sub[c])) break;
        r += l;
    }
    if (r > q) return r;
    return last_both_end;
}

/* U+L* at s: end of the match, or 0 */
static const uint8_t *match_UpL(const uint8_t *s, const uint8_t *end) {
    const uint8_t *q = s;
    while (q < end) {
        uint32_t c;
        int l = utf8_decode(q, end, &c);
        if (!in_U(g_sub[c])) break;
        q += l;
    }
    if (q == s) return 0;
    while (q < end) {
        uint32_t c;
        int l = utf8_decode(q, end, &c);
        if (!in_L(g_sub[c])) break;
        q += l;
    }
    return q;
}

static size_t match_at_config(const uint8_t *q, const uint8_t *end) {
    uint32_t c0, c1 = 0;
    int k0, k1 = -1;
    int l0 = char_at(q, end, &c0, &k0), l1 = 0;
    if (q + l0 < end) l1 = char_at(q + l0, end, &c1, &k1);
    const int prefix_ok = k0 != CL_R && k0 != CL_L && k0 != CL_N;          /* [^\r\n\p{L}\p{N}] */
    const uint8_t *e;
    /* B1: P?U*L+  (greedy optional prefix first, then without it) */
    if (prefix_ok && l1 && (e = match_UL(q + l0, end)) != 0) return (size_t)(e - q);
    if ((e = match_UL(q, end)) != 0) return (size_t)(e - q);
    /* B2: P?U+L* */
    if (prefix_ok && l1 && (e = match_UpL(q + l0, end)) != 0) return (size_t)(e - q);
    if ((e = match_UpL(q, end)) != 0) return (size_t)(e - q);
    /* B3: \p{N} */
    if (k0 == CL_N) return (size_t)l0;
    /* B4:  ?[^\s\p{L}\p{N}]+[\r\n/]* */
    {
        const uint8_t *s = 0;
        if (c0 == ' ' && k1 == CL_O) s = q + l0;
        else if (k0 == CL_O) s = q;
        if (s) {
            while (s < end) {
                uint32_t c; int k;
                int l = char_at(s, end, &c, &k);
                if (k != CL_O) break;
                s += l;
            }
            while (s < end && (*s == '\r' || *s == '\n' || *s == '/')) s++;
            return (size_t)(s - q);
        }
    }
    /* whitespace: B5 \s*[\r\n]+, B6 \s+(?!\S), B7 \s+ -- as A5..A7 above */
    {
        const uint8_t *e2 = q, *last_r_end = 0, *last_char = q;
        while (e2 < end) {
            uint32_t c; int k;
            int l = char_at(e2, end, &c, &k);
            if (!is_ws(k)) break;
            last_char = e2;
            e2 += l;
            if (k == CL_R) last_r_end = e2;
        }
        if (last_r_end) return (size_t)(last_r_end - q);
        if (e2 == end) return (size_t)(e2 - q);
        if (last_char > q) return (size_t)(last_char - q);
        return (size_t)(e2 - q);
    }
}

/* orc_split / orc_encode for the stored pattern (same contracts) */
int64_t orc_split_config(const uint8_t *text, uint64_t n, uint64_t *ends) {
    init_classes();
    if (!orc_utf8_valid(text, n)) return -1;
    const uint8_t *q = text, *end = text + n;
    int64_t k = 0;
    while (q < end) {
        q += match_at_config(q, end);
        ends[k++] = (uint64_t)(q - text);
    }
    return k;
}

/* ------------------------------------------------------------------ byte pair merge */

typedef struct { uint32_t start; uint32_t rank; } part_t;

static inline uint32_t get_rank3(const orc_t *o, const uint8_t *piece, const part_t *parts,
                                 size_t nparts, size_t i) {
    if (i + 3 < nparts)
        return vocab_get(o, piece + parts[i].start, parts[i + 3].start - parts[i].start);
    return ORC_INF;
}

/* Literal restatement of upstream tiktoken `_byte_pair_merge` + `byte_pair_encode`:
   quadratic (linear min scan + array removal per merge).  out needs room for n ranks. */
static size_t bpe_piece_literal(const orc_t *o, const uint8_t *piece, size_t n, uint32_t *out,
                                part_t *parts) {
    if (n == 1) { out[0] = vocab_get(o, piece, 1); return 1; }
    size_t np = 0;
    uint32_t min_rank = ORC_INF; size_t min_i = 0;
    for (size_t i = 0; i + 1 < n; i++) {
        uint32_t r = vocab_get(o, piece + i, 2);
        if (r < min_rank) { min_rank = r; min_i = i; }
        parts[np].start = (uint32_t)i; parts[np].rank = r; np++;
    }
    parts[np].start = (uint32_t)(n - 1); parts[np].rank = ORC_INF; np++;
    parts[np].start = (uint32_t)n; parts[np].rank = ORC_INF; np++;
    while (min_rank != ORC_INF) {
        size_t i = min_i;
        if (i > 0) parts[i - 1].rank = get_rank3(o, piece, parts, np, i - 1);
        parts[i].rank = get_rank3(o, piece, parts, np, i);
        memmove(parts + i + 1, parts + i + 2, (np - i - 2) * sizeof(part_t));
        np--;
        min_rank = ORC_INF;
        for (size_t j = 0; j + 1 < np; j++)
            if (parts[j].rank < min_rank) { min_rank = parts[j].rank; min_i = j; }
    }
    for (size_t j = 0; j + 1 < np; j++)
        out[j] = vocab_get(o, piece + parts[j].start, parts[j + 1].start - parts[j].start);
    return np - 1;
}

/* Same definition (lowest rank first, leftmost on ties), evaluated with a binary heap over
   (rank, start) and a doubly linked part list: O(n log n).  Used for long pieces where the
   literal form is too slow; cross-checked against bpe_piece_literal in tests/. */
typedef struct { uint32_t rank; uint32_t pos; uint32_t right_end; } hent_t;

static inline int hless(hent_t a, hent_t b) {
    return a.rank < b.rank || (a.rank == b.rank && a.pos < b.pos);
}

static void hpush(hent_t *h, size_t *n, hent_t e) {
    size_t i = (*n)++;
    while (i > 0) {
        size_t p = (i - 1) >> 1;
        if (!hless(e, h[p])) break;
        h[i] = h[p]; i = p;
    }
    h[i] = e;
}

static hent_t hpop(hent_t *h, size_t *n) {
    hent_t top = h[0], e = h[--(*n)];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && hless(h[c + 1], h[c])) c++;
        if (!hless(h[c], e)) break;
        h[i] = h[c]; i = c;
    }
    if (*n) h[i] = e;
    return top;
}

static size_t bpe_piece_heap(const orc_t *o, const uint8_t *piece, size_t n, uint32_t *out) {
    if (n == 1) { out[0] = vocab_get(o, piece, 1); return 1; }
    /* part i starts at byte i while alive; nxt[i] = start of the following part (n = none);
       prv[i] = start of the previous part (UINT32_MAX = none). */
    uint32_t *nxt = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint32_t *prv = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    uint8_t *alive = (uint8_t *)malloc(n + 1);
    hent_t *heap = (hent_t *)malloc((3 * n + 4) * sizeof(hent_t));
    size_t hn = 0;
    for (size_t i = 0; i < n; i++) { nxt[i] = (uint32_t)(i + 1); prv[i] = i ? (uint32_t)(i - 1) : 0xFFFFFFFFu; alive[i] = 1; }
    for (size_t i = 0; i + 1 < n; i++) {
        uint32_t r = vocab_get(o, piece + i, 2);
        if (r != ORC_INF) { hent_t e = { r, (uint32_t)i, (uint32_t)(i + 2) }; hpush(heap, &hn, e); }
    }
    while (hn) {
        hent_t e = hpop(heap, &hn);
        uint32_t a = e.pos;
        if (!alive[a]) continue;
        uint32_t b = nxt[a];
        if (b >= n) continue;
        if (nxt[b] != e.right_end) continue;         /* stale: right part changed */
        /* also stale if the left part grew since the entry was pushed: entries are pushed
           whenever a part changes, keyed by its (start, end of pair); a changed left part has
           the same start, so compare the recorded pair extent only -- the left part's own end
           is b, which is implied by nxt[a] at pop time, and the rank was computed for
           [a, right_end), which is still exactly the bytes of (part a, part b). */
        uint32_t c = nxt[b];
        alive[b] = 0;
        nxt[a] = c;
        if (c < n) prv[c] = a;
        if (c < n) {
            uint32_t r = vocab_get(o, piece + a, nxt[c] - a);
            if (r != ORC_INF) { hent_t x = { r, a, nxt[c] }; hpush(heap, &hn, x); }
        }
        uint32_t p = prv[a];
        if (p != 0xFFFFFFFFu) {
            uint32_t r = vocab_get(o, piece + p, c - p);
            if (r != ORC_INF) { hent_t x = { r, p, c }; hpush(heap, &hn, x); }
        }
    }
    size_t k = 0;
    for (uint32_t i = 0; i < n; i = nxt[i]) out[k++] = vocab_get(o, piece + i, nxt[i] - i);
    free(nxt); free(prv); free(alive); free(heap);
    return k;
}

#define ORC_LITERAL_MAX 512

/* Encode one piece (CoreBPE: whole-piece lookup, else merge).  mode 0 = auto (literal up to
   ORC_LITERAL_MAX bytes, heap above), 1 = always literal, 2 = always heap. */
uint64_t orc_encode_piece(const orc_t *o, const uint8_t *piece, uint64_t n, uint32_t *out, int mode) {
    uint32_t r = vocab_get(o, piece, n);
    if (r != ORC_INF) { out[0] = r; return 1; }
    if (mode == 1 || (mode == 0 && n <= ORC_LITERAL_MAX)) {
        part_t *parts = (part_t *)malloc((n + 2) * sizeof(part_t));
        size_t k = bpe_piece_literal(o, piece, n, out, parts);
        free(parts);
        return k;
    }
    return bpe_piece_heap(o, piece, n, out);
}

/* ------------------------------------------------------------------ encode */

/* Tekkenizer::encode (src/tekkenizer.rs:378-405) over one text.  out needs room for
   n + 2 ids.  Returns the id count, -1 for invalid UTF-8, -2 when BOS/EOS was requested but
   the control token is absent (TokenNotFound, src/tekkenizer.rs:335-340). */
int64_t orc_encode(const orc_t *o, const uint8_t *text, uint64_t n, int add_bos, int add_eos,
                   uint32_t *out, int mode) {
    /* mode bit 2 (value 4): split with the pattern STORED in tekken.json (match_at_config) instead of the
       reference's hard-coded one; the low two bits choose the merge loop as in orc_encode_piece */
    const int cfg = mode & 4;
    mode &= 3;
    if (cfg) init_classes();
    if (!orc_utf8_valid(text, n)) return -1;
    int64_t k = 0;
    if (add_bos) { if (o->bos_id < 0) return -2; out[k++] = (uint32_t)o->bos_id; }
    const uint8_t *q = text, *end = text + n;
    while (q < end) {
        size_t len = cfg ? match_at_config(q, end) : match_at(q, end);
        uint64_t c = orc_encode_piece(o, q, len, out + k, mode);
        for (uint64_t j = 0; j < c; j++) out[k + j] += o->num_special;
        k += (int64_t)c;
        q += len;
    }
    if (add_eos) { if (o->eos_id < 0) return -2; out[k++] = (uint32_t)o->eos_id; }
    return k;
}

/* orc_encode with the pattern stored in tekken.json (see match_at_config) */
int64_t orc_encode_config(const orc_t *o, const uint8_t *text, uint64_t n, int add_bos, int add_eos,
                          uint32_t *out, int mode) {
    init_classes();
    if (!orc_utf8_valid(text, n)) return -1;
    int64_t k = 0;
    if (add_bos) { if (o->bos_id < 0) return -2; out[k++] = (uint32_t)o->bos_id; }
    const uint8_t *q = text, *end = text + n;
    while (q < end) {
        size_t len = match_at_config(q, end);
        uint64_t c = orc_encode_piece(o, q, len, out + k, mode);
        for (uint64_t j = 0; j < c; j++) out[k + j] += o->num_special;
        k += (int64_t)c;
        q += len;
    }
    if (add_eos) { if (o->eos_id < 0) return -2; out[k++] = (uint32_t)o->eos_id; }
    return k;
}

/* Batch form used by the parity tests and the CPU baseline: docs are data[doc_off[d] ..
   doc_off[d+1]); ids are written back to back, tok_off gets n_docs+1 offsets.  Returns 0,
   or -(d+1)*4-1 / -(d+1)*4-2 style codes folded into: -1 invalid UTF-8, -2 token not found
   (first failing doc index in *bad_doc). */
int orc_encode_batch(const orc_t *o, const uint8_t *data, const uint64_t *doc_off, uint64_t n_docs,
                     int add_bos, int add_eos, uint32_t *out, uint64_t *tok_off, int mode,
                     uint64_t *bad_doc) {
    uint64_t k = 0;
    for (uint64_t d = 0; d < n_docs; d++) {
        tok_off[d] = k;
        int64_t c = orc_encode(o, data + doc_off[d], doc_off[d + 1] - doc_off[d], add_bos, add_eos,
                               out + k, mode);
        if (c < 0) { if (bad_doc) *bad_doc = d; return (int)c; }
        k += (uint64_t)c;
    }
    tok_off[n_docs] = k;
    return 0;
}

/* ------------------------------------------------------------------ decode */

enum { POL_IGNORE = 0, POL_KEEP = 1, POL_RAISE = 2 };

/* Tekkenizer::decode (src/tekkenizer.rs:436-560) for one id sequence.  out needs room for
   orc_decode_bound() bytes.  part_end (may be NULL) receives the end offset of every element
   decode_all would return (one per ordinary run, one per kept special id); *n_parts their
   count.  Returns the byte count, or
     -3  SpecialTokenPolicy error (Raise and a special id present)        :531-535
     -4  Tokenizers error: unknown rank, or an ordinary run whose bytes are not UTF-8  :552-555 */
int64_t orc_decode(const orc_t *o, const uint32_t *ids, uint64_t n, int policy, uint8_t *out,
                   uint64_t *part_end, uint64_t *n_parts) {
    uint64_t k = 0, np = 0, i = 0;
    while (i < n) {
        int special = ids[i] < o->num_special;
        uint64_t j = i;
        while (j < n && (ids[j] < o->num_special) == special) j++;
        if (special) {
            if (policy == POL_RAISE) return -3;
            if (policy == POL_KEEP) {
                for (uint64_t t = i; t < j; t++) {
                    uint64_t a = o->sp_off[ids[t]], b = o->sp_off[ids[t] + 1];
                    memcpy(out + k, o->sp_bytes + a, b - a);
                    k += b - a;
                    if (part_end) part_end[np] = k;
                    np++;
                }
            }
        } else {
            uint64_t k0 = k;
            for (uint64_t t = i; t < j; t++) {
                uint32_t r = ids[t] - o->num_special;
                if (r >= o->n_vocab) return -4;
                uint64_t a = o->off[r], b = o->off[r + 1];
                memcpy(out + k, o->bytes + a, b - a);
                k += b - a;
            }
            if (!orc_utf8_valid(out + k0, k - k0)) return -4;
            if (part_end) part_end[np] = k;
            np++;
        }
        i = j;
    }
    if (n_parts) *n_parts = np;
    return (int64_t)k;
}

/* Upper bound on decode output bytes for an id sequence. */
uint64_t orc_decode_bound(const orc_t *o, const uint32_t *ids, uint64_t n) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (ids[i] < o->num_special) k += o->sp_off[ids[i] + 1] - o->sp_off[ids[i]];
        else if (ids[i] - o->num_special < o->n_vocab) {
            uint32_t r = ids[i] - o->num_special;
            k += o->off[r + 1] - o->off[r];
        }
    }
    return k;
}

/* ------------------------------------------------------------------ threaded batch (CPU baseline) */

#include <pthread.h>

typedef struct {
    const orc_t *o; const uint8_t *data; const uint64_t *doc_off; uint64_t d0, d1;
    int add_bos, add_eos, mode; uint64_t n_tok; int rc;
} job_t;

static void *count_worker(void *arg) {
    job_t *j = (job_t *)arg;
    uint64_t maxlen = 0;
    for (uint64_t d = j->d0; d < j->d1; d++) {
        uint64_t l = j->doc_off[d + 1] - j->doc_off[d];
        if (l > maxlen) maxlen = l;
    }
    uint32_t *buf = (uint32_t *)malloc((maxlen + 2) * sizeof(uint32_t));
    uint64_t total = 0;
    for (uint64_t d = j->d0; d < j->d1; d++) {
        int64_t c = orc_encode(j->o, j->data + j->doc_off[d], j->doc_off[d + 1] - j->doc_off[d],
                               j->add_bos, j->add_eos, buf, j->mode);
        if (c < 0) { j->rc = (int)c; break; }
        total += (uint64_t)c;
    }
    free(buf);
    j->n_tok = total;
    return 0;
}

/* Encode all docs on n_threads host threads (contiguous doc ranges balanced by bytes) and
   return only the total id count: the timing loop of the CPU baseline.  Returns 0 or the
   first negative per-doc code. */
int orc_encode_count_mt(const orc_t *o, const uint8_t *data, const uint64_t *doc_off, uint64_t n_docs,
                        int add_bos, int add_eos, int mode, int n_threads, uint64_t *n_tok) {
    if (n_threads < 1) n_threads = 1;
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t total = doc_off[n_docs] - doc_off[0], d = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t target = doc_off[0] + total * (uint64_t)(t + 1) / (uint64_t)n_threads;
        uint64_t d1 = d;
        while (d1 < n_docs && (t == n_threads - 1 || doc_off[d1 + 1] <= target)) d1++;
        jobs[t].o = o; jobs[t].data = data; jobs[t].doc_off = doc_off; jobs[t].d0 = d; jobs[t].d1 = d1;
        jobs[t].add_bos = add_bos; jobs[t].add_eos = add_eos; jobs[t].mode = mode;
        d = d1;
        pthread_create(&th[t], 0, count_worker, &jobs[t]);
    }
    uint64_t sum = 0; int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], 0);
        sum += jobs[t].n_tok;
        if (jobs[t].rc && !rc) rc = jobs[t].rc;
    }
    free(jobs); free(th);
    if (n_tok) *n_tok = sum;
    return rc;
}

/* Multi-threaded batch encode that returns the ids (parity checks at full config sizes): every
   thread encodes a contiguous byte-balanced document range into its own buffer; the buffers are
   then concatenated in document order. */
typedef struct {
    const orc_t *o; const uint8_t *data; const uint64_t *doc_off; uint64_t d0, d1;
    int add_bos, add_eos, mode; uint32_t *buf; uint64_t *cnt; uint64_t n_tok; int rc; uint64_t bad;
} ejob_t;

static void *encode_worker(void *arg) {
    ejob_t *j = (ejob_t *)arg;
    uint64_t k = 0;
    for (uint64_t d = j->d0; d < j->d1; d++) {
        int64_t c = orc_encode(j->o, j->data + j->doc_off[d], j->doc_off[d + 1] - j->doc_off[d],
                               j->add_bos, j->add_eos, j->buf + k, j->mode);
        if (c < 0) { j->rc = (int)c; j->bad = d; break; }
        j->cnt[d] = (uint64_t)c;
        k += (uint64_t)c;
    }
    j->n_tok = k;
    return 0;
}

int orc_encode_batch_mt(const orc_t *o, const uint8_t *data, const uint64_t *doc_off, uint64_t n_docs,
                        int add_bos, int add_eos, uint32_t *out, uint64_t *tok_off, int mode,
                        int n_threads, uint64_t *bad_doc) {
    if (n_threads < 1) n_threads = 1;
    ejob_t *jobs = (ejob_t *)calloc((size_t)n_threads, sizeof(ejob_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t *cnt = (uint64_t *)calloc(n_docs + 1, sizeof(uint64_t));
    uint64_t total = doc_off[n_docs] - doc_off[0], d = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t target = doc_off[0] + total * (uint64_t)(t + 1) / (uint64_t)n_threads;
        uint64_t d1 = d;
        while (d1 < n_docs && (t == n_threads - 1 || doc_off[d1 + 1] <= target)) d1++;
        jobs[t].o = o; jobs[t].data = data; jobs[t].doc_off = doc_off; jobs[t].d0 = d; jobs[t].d1 = d1;
        jobs[t].add_bos = add_bos; jobs[t].add_eos = add_eos; jobs[t].mode = mode; jobs[t].cnt = cnt;
        jobs[t].buf = (uint32_t *)malloc(((doc_off[d1] - doc_off[d]) + 2 * (d1 - d) + 1) * sizeof(uint32_t));
        d = d1;
        pthread_create(&th[t], 0, encode_worker, &jobs[t]);
    }
    int rc = 0;
    uint64_t k = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], 0);
        if (jobs[t].rc && !rc) { rc = jobs[t].rc; if (bad_doc) *bad_doc = jobs[t].bad; }
        if (!rc) { memcpy(out + k, jobs[t].buf, jobs[t].n_tok * sizeof(uint32_t)); k += jobs[t].n_tok; }
        free(jobs[t].buf);
    }
    if (!rc) {
        uint64_t acc = 0;
        for (uint64_t i = 0; i < n_docs; i++) { tok_off[i] = acc; acc += cnt[i]; }
        tok_off[n_docs] = acc;
    }
    free(cnt); free(jobs); free(th);
    return rc;
}
