// cfgsplit_host.cpp -- runs the groundwork split for the pattern stored in tekken.json
// (tekken_rs_b200/csrc/tk_pretok_cfg.h: safe starts + one sequential matcher walk per segment) on the CPU, so the
// CPU test-suite can compare its piece boundaries with the oracle.  Test infrastructure.
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../tekken_rs_b200/csrc/tk_host.h"         // build_cfg_unicode_tables (the library's own table builder)
#include "../../tekken_rs_b200/csrc/tk_pretok.h"       // TkBytesChecked
#include "../../tekken_rs_b200/csrc/tk_pretok_cfg.h"

static std::vector<uint16_t> g_s1;
static std::vector<uint8_t> g_s2;

// the tables the library itself would upload; returns the number of scalar values whose class differs from cls_flat
// (one class per scalar value, measured from the engine by the oracle tools) -- must be 0
extern "C" int64_t cfgsplit_use_library_tables(const uint8_t* cls_flat) {
    tk::build_cfg_unicode_tables(g_s1, g_s2);
    const TkCfgTables T{g_s1.data(), g_s2.data()};
    int64_t diff = 0;
    for (uint32_t c = 0; c < 0x110000u; ++c)
        if (c != 0x0A && c != 0x0D && tk_cfg_class(T, c) != cls_flat[c]) ++diff;
    // the table-free ASCII classes must be the table's
    for (uint32_t c = 0; c < 0x80u; ++c) {
        if (c == 0x0A || c == 0x0D) { if (tk_cfg_class_ascii(c) != TK_CC_R) ++diff; continue; }
        const uint32_t blk = T.stage1[c >> 7];
        const uint32_t b = T.stage2[blk * 64u + ((c & 127u) >> 1)];
        if (((b >> ((c & 1u) * 4u)) & 15u) != tk_cfg_class_ascii(c)) ++diff;
    }
    return diff;
}

// cls_flat: one class (TK_CC_*, CR/LF excluded) per scalar value 0 .. 0x10FFFF
extern "C" void cfgsplit_set_classes(const uint8_t* cls_flat) {
    g_s1.assign(0x110000 >> 7, 0);
    g_s2.clear();
    std::map<std::string, uint16_t> seen;
    for (uint32_t b = 0; b < (0x110000u >> 7); ++b) {
        std::string blk(64, '\0');
        for (uint32_t i = 0; i < 128; ++i) {
            const uint8_t c = cls_flat[b * 128 + i] & 15u;
            blk[i >> 1] = (char)((uint8_t)blk[i >> 1] | (c << ((i & 1u) * 4u)));
        }
        auto it = seen.find(blk);
        if (it == seen.end()) {
            it = seen.emplace(blk, (uint16_t)(g_s2.size() / 64)).first;
            g_s2.insert(g_s2.end(), blk.begin(), blk.end());
        }
        g_s1[b] = it->second;
    }
}

// start_mask: (n/32 + 1) words, bit = a piece starts at that byte.  Returns the number of safe starts, or
// -1 - position when a matcher walk from a safe start does not land on the next safe start (the scheme would be wrong).
extern "C" int64_t cfgsplit_host(const uint8_t* data, uint64_t n, const uint64_t* doc_off, uint64_t n_docs, uint32_t* start_mask) {
    const TkCfgTables T{g_s1.data(), g_s2.data()};
    const TkBytesChecked src{data, n};
    memset(start_mask, 0, (n / 32 + 1) * 4);
    int64_t n_safe = 0;
    std::vector<uint32_t> ds_bits(n / 32 + 2, 0);
    for (uint64_t d = 0; d <= n_docs; ++d) if (doc_off[d] < n) ds_bits[doc_off[d] >> 5] |= 1u << (doc_off[d] & 31);
    for (uint64_t d = 0; d < n_docs; ++d) {
        const int64_t a = (int64_t)doc_off[d], e = (int64_t)doc_off[d + 1];
        // pass 1: safe starts by the scalar rule (pass 1b below: the same from window bit logic)
        std::vector<int64_t> safe;
        uint32_t prev = 0xFFu;
        for (int64_t p = a; p < e;) {
            uint32_t c, cp;
            const int l = tk_cfg_char(src, p, T, &c, &cp);
            if (tk_cfg_safe_start(p == a, prev, c)) safe.push_back(p);
            prev = c;
            p += l;
        }
        safe.push_back(e);
        n_safe += (int64_t)safe.size() - 1;
        // pass 2 (one lane per segment): sequential matcher from a safe start to the next one
        for (size_t i = 0; i + 1 < safe.size(); ++i) {
            int64_t q = safe[i];
            // the document's end is given the way a kernel has it: the document-start bitmask
            const TkCfgEndMask stop{ds_bits.data(), (int64_t)n, a};
            while (q < safe[i + 1]) {
                start_mask[q >> 5] |= 1u << (q & 31);
                const int64_t e1 = tk_cfg_match_end(src, q, stop, T);
                if (e1 != tk_cfg_match_end(src, q, TkCfgEndAt{e}, T)) return -3000000000LL - q;
                q = e1;
            }
            if (q != safe[i + 1]) return -1 - safe[i + 1];
        }
    }
    // pass 1b: what K1 would run -- one "thread" per 32-byte window, class masks + bit logic; must give the same safe
    // starts (documents back to back: a document start is always a safe start)
    {
        const uint64_t nw = n / 32 + 1;
        std::vector<uint32_t> ds(nw + 1, 0);
        for (uint64_t d = 0; d < n_docs; ++d) if (doc_off[d] < n) ds[doc_off[d] >> 5] |= 1u << (doc_off[d] & 31);
        TkCfgWin prev{};
        std::vector<uint32_t> want(nw, 0);
        for (uint64_t d = 0; d < n_docs; ++d) {
            uint32_t pc = 0xFFu;
            for (int64_t p = (int64_t)doc_off[d]; p < (int64_t)doc_off[d + 1];) {
                uint32_t c, cp;
                const int l = tk_cfg_char(src, p, T, &c, &cp);
                if (tk_cfg_safe_start(p == (int64_t)doc_off[d], pc, c)) want[p >> 5] |= 1u << (p & 31);
                pc = c;
                p += l;
            }
        }
        for (uint64_t w = 0; w < nw; ++w) {
            uint32_t words[8] = {0};
            const uint64_t pos = w * 32;
            const uint64_t m = pos < n ? (n - pos < 32 ? n - pos : 32) : 0;
            memcpy(words, data + pos, m);
            const TkCfgWin c = tk_cfg_classify_window(src, pos, words, ds[w], T);
            const uint64_t valid = (pos + 32 <= n) ? 0xFFFFFFFFull : ((1ull << (n - pos)) - 1ull);
            if (c.bad & (uint32_t)valid) return -2000000000LL - (int64_t)pos;
            const uint32_t got = tk_cfg_safe_mask(prev, c) & (uint32_t)valid;
            if (got != want[w]) return -1000000000LL - (int64_t)pos;
            prev = c;
        }
        // pass 2b: the walk exactly as a kernel would run it -- one "lane" per safe start, in no particular order,
        // knowing only the safe-start mask and the document-start mask; must reproduce the piece-start mask of pass 2
        std::vector<uint32_t> starts(nw, 0);
        for (uint64_t wi = nw; wi-- > 0;) {
            uint32_t m = want[wi];
            starts[wi] |= m;
            while (m) {
                const int b = __builtin_ctz(m);
                m &= m - 1;
                tk_cfg_walk(src, (int64_t)(wi * 32 + b), want.data(), ds_bits.data(), (int64_t)n, T,
                            [&](int64_t pos) { starts[pos >> 5] |= 1u << (pos & 31); });
            }
        }
        for (uint64_t w = 0; w < nw; ++w)
            if (starts[w] != start_mask[w]) return -4000000000LL - (int64_t)(w * 32);
    }
    return n_safe;
}
