// tk_pretok.h -- the pre-tokeniser (regex split) as position-wise bit logic over 32-byte windows.
//
// Replaces the fancy-regex `find_iter` loop inside tiktoken-rs `CoreBPE::encode` (called from
// src/tekkenizer.rs:384-386) for the pattern hard-coded at src/tekkenizer.rs:123:
//
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ?[^\s\p{L}\p{N}]+[\r\n]*
//   | \s*[\r\n]+ | \s+(?!\S) | \s+
//
// Instead of walking matches left to right, every byte decides independently whether a piece
// (regex match) starts there, from: the class of its char and of a few neighbouring chars, and
// three run properties (offset inside a digit run mod 3; whether a CR/LF run was swallowed by
// the preceding punctuation piece; whether a CR/LF occurs later in the same whitespace run).
// One thread owns one 32-byte window and works on 32-bit masks (bit i = byte i of the window),
// so the per-byte cost is a handful of bitwise instructions.  The run properties are carried
// between windows as tiny monoid elements (see TkRunSummary) that the kernel combines with
// warp/block scans and, across thread blocks, with a scan kernel + a rare fix-up pass.
//
// Everything here is __host__ __device__: tests/native/ runs exactly this code on the CPU
// against the oracle, the kernels in tk_kernels.cu run it on the GPU.
#pragma once
#include "tk_common.h"

#if defined(__CUDA_ARCH__)
#define TK_CLZ(x) __clz((int)(x))
#define TK_FFS(x) __ffs((int)(x))
#define TK_POPC(x) __popc(x)
#define TK_FFSLL(x) __ffsll((long long)(x))
#else
static inline int TK_CLZ(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int TK_FFS(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int TK_POPC(uint32_t x) { return __builtin_popcount(x); }
static inline int TK_FFSLL(uint64_t x) { return x ? __builtin_ctzll(x) + 1 : 0; }
#endif

// Class masks of one 32-byte window.  Class bits are "spread": every byte of a multi-byte char
// carries the class of the char, so `mask << 1` answers "class of the previous char" at a lead.
struct TkWin {
    uint32_t lead;  // byte starts a char (is not 10xxxxxx)
    uint32_t mL, mN, mR, mW;  // \p{L}, \p{N}, CR/LF, other \s ; O = none of them
    uint32_t sp, ap;          // U+0020, U+0027
    uint32_t ds;              // a document starts at this byte (or: end-of-data sentinel)
    uint32_t bad;             // invalid UTF-8 detected at this byte
};

TK_HD uint32_t tk_mO(const TkWin& w) { return ~(w.mL | w.mN | w.mR | w.mW); }

// ---- SWAR helpers: four 7-bit bytes per 32-bit word, result in the 0x80 bit of each byte ----
#define TK_H 0x80808080u
TK_HD uint32_t tk_swar_ge(uint32_t w7, uint32_t c) { return ((w7 | TK_H) - c * 0x01010101u) & TK_H; }
TK_HD uint32_t tk_swar_le(uint32_t w7, uint32_t c) { return ((c | 0x80u) * 0x01010101u - w7) & TK_H; }
TK_HD uint32_t tk_swar_eq(uint32_t w7, uint32_t c) { return ~((w7 ^ (c * 0x01010101u)) + 0x7F7F7F7Fu) & TK_H; }
// gather the four 0x80 flag bits of a word into bits 0..3
TK_HD uint32_t tk_swar_nib(uint32_t flags) { return (((flags >> 7) * 0x00204081u) >> 21) & 0xFu; }

// Where the bytes around a window come from.  TkBytesChecked reads the text itself (bytes outside
// [0, n) read as 0).  TkBytesTile reads a shared-memory copy of a tile and its halo in which the
// bytes outside the text are already zero, so no bounds are checked; a character cut off by the
// end of the text then fails on its missing continuation bytes.
struct TkBytesChecked {
    const uint8_t* data;
    uint64_t n;
    TK_HD uint32_t at(int64_t pos) const { return (pos >= 0 && (uint64_t)pos < n) ? (uint32_t)data[pos] : 0u; }
    TK_HD bool past_end(int64_t end) const { return (uint64_t)end > n; }
};
struct TkBytesTile {
    const uint8_t* base;   // base[pos - origin] is the byte at text position pos
    int64_t origin;
    TK_HD uint32_t at(int64_t pos) const { return (uint32_t)base[pos - origin]; }
    TK_HD bool past_end(int64_t) const { return false; }
};

TK_HD uint32_t tk_byte_at(const uint8_t* data, uint64_t n, int64_t pos) {
    return (pos >= 0 && (uint64_t)pos < n) ? (uint32_t)data[pos] : 0u;
}

// Strict decode of the scalar value whose lead byte is at pos.  Returns its length (1..4) and
// class in *cls, or 0 if the bytes are not valid UTF-8.
template <class B>
TK_HD int tk_decode_at(const B& src, int64_t pos, const TkDeviceTables& T, uint32_t* cls) {
    uint32_t b0 = src.at(pos);
    uint32_t cp;
    int len;
    if (b0 < 0x80u) { cp = b0; len = 1; }
    else if (b0 < 0xC2u) return 0;
    else if (b0 < 0xE0u) {
        uint32_t b1 = src.at(pos + 1);
        if ((b1 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
        len = 2;
    } else if (b0 < 0xF0u) {
        uint32_t b1 = src.at(pos + 1), b2 = src.at(pos + 2);
        if ((b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
        if (cp < 0x800u || (cp >= 0xD800u && cp <= 0xDFFFu)) return 0;
        len = 3;
    } else if (b0 < 0xF5u) {
        uint32_t b1 = src.at(pos + 1), b2 = src.at(pos + 2), b3 = src.at(pos + 3);
        if ((b1 & 0xC0u) != 0x80u || (b2 & 0xC0u) != 0x80u || (b3 & 0xC0u) != 0x80u) return 0;
        cp = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
        if (cp < 0x10000u || cp > 0x10FFFFu) return 0;
        len = 4;
    } else return 0;
    if (src.past_end(pos + len)) return 0;
    if (cp == 0x0Au || cp == 0x0Du) *cls = TK_CL_R;
    else *cls = tk_class_lookup(T.uni_stage1, T.uni_stage2, cp);
    return len;
}

// The ASCII half of the classification of a 32-byte window: class masks of its ASCII bytes, lead / continuation
// bytes, *hi = bytes >= 0x80 (their classes are still missing: tk_classify_window adds them char by char, the
// pre-tokeniser kernel spreads that work over the block).  bad = 0.
TK_HD TkWin tk_classify_ascii(const uint32_t* w, uint32_t ds_word, uint32_t* hi_out) {
    TkWin r;
    uint32_t mL = 0, mN = 0, mR = 0, mW = 0, sp = 0, ap = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t x = w[j];
        uint32_t h = x & TK_H;
        uint32_t w7 = x & 0x7F7F7F7Fu;
        uint32_t ok = ~h;
        uint32_t f = w7 | 0x20202020u;
        uint32_t alpha = tk_swar_ge(f, 0x61u) & tk_swar_le(f, 0x7Au) & ok;
        uint32_t digit = tk_swar_ge(w7, 0x30u) & tk_swar_le(w7, 0x39u) & ok;
        uint32_t c9_13 = tk_swar_ge(w7, 0x09u) & tk_swar_le(w7, 0x0Du) & ok;
        uint32_t rr = (tk_swar_eq(w7, 0x0Au) | tk_swar_eq(w7, 0x0Du)) & ok;
        uint32_t e20 = tk_swar_eq(w7, 0x20u) & ok;
        uint32_t e27 = tk_swar_eq(w7, 0x27u) & ok;
        int s = 4 * j;
        mL |= tk_swar_nib(alpha) << s;
        mN |= tk_swar_nib(digit) << s;
        mR |= tk_swar_nib(rr) << s;
        mW |= tk_swar_nib((c9_13 & ~rr) | e20) << s;
        sp |= tk_swar_nib(e20) << s;
        ap |= tk_swar_nib(e27) << s;
        hi |= tk_swar_nib(h) << s;
    }
    uint32_t lead = 0xFFFFFFFFu;
    if (hi) {
        // continuation bytes: 10xxxxxx.  (hi bytes with bit 6 clear)
        uint32_t cont = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) cont |= tk_swar_nib((w[j] & TK_H) & ~((w[j] << 1) & TK_H)) << (4 * j);
        lead = ~cont;
    }
    r.lead = lead; r.mL = mL; r.mN = mN; r.mR = mR; r.mW = mW; r.sp = sp; r.ap = ap; r.ds = ds_word; r.bad = 0;
    *hi_out = hi;
    return r;
}

// Leading continuation bytes of the window at pos belong to a char that starts in the window before: its class for the
// bytes that fall into this window (*cov = the bytes it covers; continuation bytes beyond them stay uncovered).
template <class B>
TK_HD void tk_cover_leading(const B& src, uint64_t pos, uint32_t cont, const TkDeviceTables& T, uint32_t* cov, uint32_t* mL,
                            uint32_t* mN, uint32_t* mW) {
    uint32_t lc = (uint32_t)(TK_FFS(~cont) - 1);  // number of leading continuation bytes (cont != ~0 -> <32)
    if (cont == 0xFFFFFFFFu) lc = 32;
    if (lc == 0) return;
    int back = 0;
    int64_t q = (int64_t)pos - 1;
    while (back < 3 && q >= 0 && (src.at(q) & 0xC0u) == 0x80u) { --q; ++back; }
    uint32_t cls = TK_CL_O;
    const int len = (q >= 0) ? tk_decode_at(src, q, T, &cls) : 0;
    const int64_t over = len ? q + len - (int64_t)pos : 0;
    if (over > 0) {
        const uint32_t m = (1u << (over < (int64_t)lc ? (uint32_t)over : lc)) - 1u;
        *cov |= m;
        if (cls == TK_CL_L) *mL |= m; else if (cls == TK_CL_N) *mN |= m; else if (cls == TK_CL_W) *mW |= m;
    }
}

// Classify the 32-byte window at byte offset `pos` (a multiple of 32).  w[0..7] are its bytes as
// little-endian words, zero padded beyond n.  ds_word = document-start bits of the window.
template <class B>
TK_HD TkWin tk_classify_window(const B& src, uint64_t pos, const uint32_t* w, uint32_t ds_word, const TkDeviceTables& T) {
    uint32_t hi;
    TkWin r = tk_classify_ascii(w, ds_word, &hi);
    uint32_t bad = 0;
    if (hi) {
        const uint32_t cont = ~r.lead;
        uint32_t covered = ~hi;  // ASCII bytes are done
        tk_cover_leading(src, pos, cont, T, &covered, &r.mL, &r.mN, &r.mW);
        // chars whose lead byte is in this window
        uint32_t todo = hi & r.lead;
        while (todo) {
            int i = TK_FFS(todo) - 1;
            todo &= todo - 1;
            uint32_t cls = TK_CL_O;
            int len = tk_decode_at(src, (int64_t)pos + i, T, &cls);
            if (len == 0) { bad |= 1u << i; continue; }
            uint32_t m = (len >= 32 - i) ? (0xFFFFFFFFu << i) : (((1u << len) - 1u) << i);
            covered |= m;
            if (cls == TK_CL_L) r.mL |= m; else if (cls == TK_CL_N) r.mN |= m; else if (cls == TK_CL_W) r.mW |= m;
        }
        bad |= ~covered;               // stray continuation bytes
    }
    bad |= ds_word & ~r.lead;          // a document may not start inside a char
    r.bad = bad;
    return r;
}

TK_HD TkWin tk_classify_window(const uint8_t* data, uint64_t n, uint64_t pos, const uint32_t* w,
                               uint32_t ds_word, const TkDeviceTables& T) {
    return tk_classify_window(TkBytesChecked{data, n}, pos, w, ds_word, T);
}

// ---- contraction alternative (?i:'s|'t|'re|'ve|'m|'ll|'d) ------------------------------------
// Unicode simple case folding adds exactly one non-ASCII member: U+017F (long s, C5 BF) ~ 's'.
// Returns 0 (no match), 2 ('x), 3 ('xy) or 4 (' + long s: 3 bytes).  `stop` = number of bytes
// after the apostrophe that still belong to the same document.
template <class B>
TK_HD int tk_contraction_len(const B& src, int64_t apos, int stop) {
    if (stop < 1) return 0;
    uint32_t c1 = src.at(apos + 1) | 0x20u;
    if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') return 2;
    if (stop < 2) return 0;
    uint32_t b1 = src.at(apos + 1), b2 = src.at(apos + 2);
    if (b1 == 0xC5u && b2 == 0xBFu) return 4;
    uint32_t c2 = b2 | 0x20u;
    if ((c1 == 'r' || c1 == 'v') && c2 == 'e') return 3;
    if (c1 == 'l' && c2 == 'l') return 3;
    return 0;
}
// (c | 0x20 maps only 'S'/'s' to 's' among bytes < 0x80 that matter: 0x53|0x20 = 0x73; the other
// byte with that image, 0x13|0x20 = 0x33, is '3', not a letter -- so the fold is exact.)

// Derived masks of a window that later windows look back at.
struct TkDerived {
    uint32_t sO;     // spread: char is an O-class char that starts a piece (opens ` ?[^\s\p{L}\p{N}]+`)
    uint32_t f2, f3, f4;  // apostrophes that fire a contraction of 2 / 3 bytes / ' + long s
};

TK_HD uint32_t tk_shl(uint32_t c, uint32_t p, int k) { return (c << k) | (p >> (32 - k)); }
TK_HD uint32_t tk_shr(uint32_t c, uint32_t nx, int k) { return (c >> k) | (nx << (32 - k)); }

// p = previous window (may be all-zero when unknown; then only bits >= 4 of the result are exact).
template <class B>
TK_HD TkDerived tk_derive(const B& src, uint64_t pos, const TkWin& p, const TkWin& c, const TkWin& nx, uint32_t sO_prev) {
    TkDerived d;
    uint32_t P_O = tk_shl(tk_mO(c), tk_mO(p), 1), P_SP = tk_shl(c.sp, p.sp, 1);
    uint32_t startok = c.ds | ~(P_O | P_SP);
    uint32_t startO = c.lead & tk_mO(c) & startok;
    uint32_t cont = ~c.lead;
    uint32_t s = startO, carry = sO_prev >> 31;
    s |= ((s << 1) | carry) & cont;
    s |= ((s << 1) | carry) & cont;
    s |= ((s << 1) | carry) & cont;
    d.sO = s;
    d.f2 = d.f3 = d.f4 = 0;
    uint32_t a = c.ap & startok;
    while (a) {
        int i = TK_FFS(a) - 1;
        a &= a - 1;
        // bytes after the apostrophe up to the next document start
        uint64_t dsn = ((uint64_t)nx.ds << 32 | c.ds) >> (i + 1);
        int stop = dsn ? TK_FFSLL(dsn) - 1 : 8;
        int len = tk_contraction_len(src, (int64_t)pos + i, stop);
        if (len == 2) d.f2 |= 1u << i; else if (len == 3) d.f3 |= 1u << i; else if (len == 4) d.f4 |= 1u << i;
    }
    return d;
}

TK_HD TkDerived tk_derive(const uint8_t* data, uint64_t n, uint64_t pos, const TkWin& p, const TkWin& c,
                          const TkWin& nx, uint32_t sO_prev) {
    return tk_derive(TkBytesChecked{data, n}, pos, p, c, nx, sO_prev);
}

// ---- run summaries carried between windows ---------------------------------------------------
// n_all/n_val : digit-run offset.  If n_all, the window is one unbroken run of N bytes with no
//               document start in it and the count entering the next window is
//               (count entering this one + n_val) mod 3; otherwise it is n_val.
// r_mode      : "CR/LF run swallowed by the punctuation piece before it".  0/1 = the last byte of
//               the window is / is not in such a run, known; 2 = same as entering the window
//               (the window is all CR/LF with no document start).
// head        : first event of the window for a whitespace run entering it: 0 = none (all W, no
//               CR/LF, no document start), 1 = a CR/LF comes first, 2 = the run ends first.
struct TkRunSummary {
    uint32_t n_all, n_val, r_mode, head;
};

TK_HD TkRunSummary tk_compose(const TkRunSummary& a, const TkRunSummary& b) {  // a then b
    TkRunSummary r;
    r.n_all = a.n_all & b.n_all;
    r.n_val = b.n_all ? (a.n_val + b.n_val) % 3u : b.n_val;
    r.r_mode = (b.r_mode == 2u) ? a.r_mode : b.r_mode;
    r.head = a.head ? a.head : b.head;
    return r;
}

TK_HD TkRunSummary tk_summarize(const TkWin& c) {
    TkRunSummary s;
    // digit run reaching the end of the window
    uint32_t nstop = ~c.mN | c.ds;
    if (nstop == 0) { s.n_all = 1; s.n_val = (uint32_t)TK_POPC(c.lead) % 3u; }
    else {
        s.n_all = 0;
        if (c.mN >> 31) {
            int q = 31 - TK_CLZ(nstop);   // highest non-N byte / document start
            int st = ((c.mN >> q) & 1u) ? q : q + 1;   // an N byte can only stop a run as a ds
            s.n_val = (uint32_t)TK_POPC(c.lead & (0xFFFFFFFFu << st)) % 3u;
        } else s.n_val = 0;
    }
    // "a CR/LF at the next byte would be swallowed by the punctuation piece before it"
    if (!(c.mR >> 31)) s.r_mode = (tk_mO(c) >> 31) & 1u;     // right after an O char: yes
    else {
        uint32_t rstop = ~c.mR | c.ds;
        if (rstop == 0) s.r_mode = 2;                         // all CR/LF: same as on entry
        else {
            int q = 31 - TK_CLZ(rstop);
            s.r_mode = ((c.mR >> q) & 1u) ? 0u                // run starts at a document start
                                          : ((tk_mO(c) >> q) & 1u);
        }
    }
    // first event for a whitespace run entering the window
    uint32_t endm = ~(c.mR | c.mW) | c.ds;
    int e = endm ? TK_FFS(endm) - 1 : 32;
    int r = c.mR ? TK_FFS(c.mR) - 1 : 32;
    s.head = (e == 32 && r == 32) ? 0u : (r < e ? 1u : 2u);
    return s;
}

// Result of evaluating one window.
struct TkEval {
    uint32_t start;     // piece-start bits decided here
    int pend;           // bit index of a W-after-CR/LF candidate whose whitespace run continues past
                        // the next window with no CR/LF so far (-1 if none): start iff the run ends
                        // before another CR/LF
};

// Evaluate window c.  p/nx = neighbours, dp = derived masks of p (bits >= 28 exact suffices),
// dc = derived masks of c.  n_in = number of digits (mod 3) of the digit run entering the window,
// abs_in = the CR/LF run entering the window was swallowed by a punctuation piece.
TK_HD TkEval tk_eval_window(const TkWin& p, const TkWin& c, const TkWin& nx, const TkDerived& dp,
                            const TkDerived& dc, uint32_t n_in, uint32_t abs_in) {
    TkEval out;
    const uint32_t cO = tk_mO(c), pO = tk_mO(p);
    const uint32_t ds = c.ds, nds = ~ds;
    const uint32_t P_L = tk_shl(c.mL, p.mL, 1), P_N = tk_shl(c.mN, p.mN, 1), P_R = tk_shl(c.mR, p.mR, 1),
                   P_W = tk_shl(c.mW, p.mW, 1), P_O = tk_shl(cO, pO, 1);
    // rule 4: an O char starts a piece unless it continues an O run or follows a space
    // (dc.sO already holds exactly that, spread; take the lead bits)
    uint32_t start = ds | (dc.sO & c.lead & cO);
    // rule 1+2: letters
    uint32_t interior = tk_shl(dc.f2 | dc.f3 | dc.f4, dp.f2 | dp.f3 | dp.f4, 1) | tk_shl(dc.f3, dp.f3, 2);
    uint32_t after = tk_shl(dc.f2, dp.f2, 2) | tk_shl(dc.f3 | dc.f4, dp.f3 | dp.f4, 3);
    uint32_t P_sO = tk_shl(dc.sO, dp.sO, 1);
    start |= c.lead & c.mL & ~interior & (after | (nds & (P_N | P_R | (P_O & ~P_sO))));
    // CR/LF: starts a piece only right after a letter or digit
    start |= c.lead & c.mR & nds & (P_L | P_N);
    // other whitespace
    const uint32_t ws = c.mR | c.mW;
    {
        uint64_t ws64 = (uint64_t)(nx.mR | nx.mW) << 32 | ws;
        uint64_t ds64 = (uint64_t)nx.ds << 32 | ds;
        uint64_t lead64 = (uint64_t)nx.lead << 32 | c.lead;
        uint64_t nxt = (~ws64 >> 1) & ~(ds64 >> 1) & 0x7FFFFFFFFFFFFFFFull;   // next byte is a non-ws char of the same doc
        uint64_t t = nxt & ((uint64_t)nx.mW << 32 | c.mW);
        uint64_t cont64 = ~lead64;
        t |= (t & cont64) >> 1;
        t |= (t & cont64) >> 1;
        uint32_t nxt_lead = (uint32_t)t;
        start |= c.lead & c.mW & nds & (P_L | P_N | P_O | (P_W & nxt_lead));
    }
    // digits: every third char of a run, counted from the run start
    if ((c.mN & ~c.lead) == 0u) {
        // every digit of the window is one byte (no continuation byte is N): bit arithmetic instead of a loop over
        // the digits.  Seeds: the first digit of every run that starts here, and for the run that enters the window
        // with n_in digits counted the first position where the count is a multiple of three again; then every third
        // position of the same run (doubling: 3, 6, 12, 24).
        const uint32_t D = c.mN;
        uint32_t S = D & (ds | ~P_N);
        if ((D & 1u) && !(ds & 1u) && (P_N & 1u)) {
            const uint32_t j0 = (3u - n_in % 3u) % 3u;
            const uint32_t need = (2u << j0) - 1u;                       // bits 0 .. j0
            if ((D & need) == need && !(ds & need)) S |= 1u << j0;
        }
        const uint32_t dsx = ds | (ds << 1) | (ds << 2);
        const uint32_t ok3 = D & (D << 1) & (D << 2) & ~dsx;             // i, i-1, i-2 digits of one run: i-3 is the same run's
        S |= (S << 3) & ok3;
        const uint32_t ok6 = ok3 & (ok3 << 3);
        S |= (S << 6) & ok6;
        const uint32_t ok12 = ok6 & (ok6 << 6);
        S |= (S << 12) & ok12;
        const uint32_t ok24 = ok12 & (ok12 << 12);
        S |= (S << 24) & ok24;
        start |= S;
    } else {
        uint32_t m = c.lead & c.mN;
        uint32_t stop = ~c.mN | ds;
        while (m) {
            int i = TK_FFS(m) - 1;
            m &= m - 1;
            uint32_t below = stop & ((2u << i) - 1u);   // stops at or below i (a ds AT i counts)
            // a non-N byte can't be at i itself (i is N), so "at i" can only be a ds
            uint32_t cnt;
            if (below) {
                int q = 31 - TK_CLZ(below);
                int st = ((c.mN >> q) & 1u) ? q : q + 1;   // an N byte can only stop a run as a ds
                cnt = (uint32_t)TK_POPC(c.lead & c.mN & ((1u << i) - 1u) & (0xFFFFFFFFu << st));
            } else {
                cnt = (uint32_t)TK_POPC(c.lead & c.mN & ((1u << i) - 1u)) + n_in;
            }
            if (cnt % 3u == 0) start |= 1u << i;
        }
    }
    // W right after CR/LF: starts a piece iff the CR/LF run was swallowed by punctuation, or no
    // CR/LF follows in the same whitespace run.
    out.pend = -1;
    {
        uint32_t cand = c.lead & c.mW & P_R & nds;
        uint32_t rstop = ~c.mR | ds;
        uint64_t ws64 = (uint64_t)(nx.mR | nx.mW) << 32 | ws;
        uint64_t ds64 = (uint64_t)nx.ds << 32 | ds;
        uint64_t r64 = (uint64_t)nx.mR << 32 | c.mR;
        while (cand) {
            int i = TK_FFS(cand) - 1;
            cand &= cand - 1;
            uint32_t absorbed;
            uint32_t below = rstop & ((1u << i) - 1u);
            if (below) {
                int q = 31 - TK_CLZ(below);
                absorbed = ((c.mR >> q) & 1u) ? 0u : ((cO >> q) & 1u);
            } else absorbed = abs_in;
            if (absorbed) { start |= 1u << i; continue; }
            uint64_t above = ~((2ull << i) - 1ull);
            uint64_t endm = (~ws64 | ds64) & above;
            if (endm) {
                int e = TK_FFSLL(endm) - 1;
                uint64_t rl = r64 & above & ((1ull << e) - 1ull);
                if (!rl) start |= 1u << i;
            } else if (!(r64 & above)) {
                out.pend = i;
            }
        }
    }
    out.start = start;
    return out;
}
