// pretok_host.cpp -- runs the product's pre-tokeniser logic (tekken_rs_b200/csrc/tk_pretok.h, the
// exact code the CUDA kernel executes per 32-byte window) on the CPU, so the CPU test-suite can
// compare its piece boundaries with the oracle without a GPU.  Test infrastructure.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../tekken_rs_b200/csrc/tk_host.h"
#include "../../tekken_rs_b200/csrc/tk_pretok.h"

static std::vector<uint16_t> g_s1;
static std::vector<uint8_t> g_s2;

// start_mask: (n/32 + 1) words; doc_off: n_docs+1 offsets.  Returns -1 - position of the first
// invalid UTF-8 byte, or the number of windows.
extern "C" int64_t pretok_host(const uint8_t* data, uint64_t n, const uint64_t* doc_off, uint64_t n_docs,
                               uint32_t* start_mask) {
    if (g_s1.empty()) tk::build_unicode_tables(g_s1, g_s2);
    TkDeviceTables T{};
    T.uni_stage1 = g_s1.data();
    T.uni_stage2 = g_s2.data();
    uint64_t nw = n / 32 + 1;
    std::vector<uint32_t> ds(nw + 2, 0);
    for (uint64_t d = 0; d <= n_docs; ++d) ds[doc_off[d] >> 5] |= 1u << (doc_off[d] & 31);
    std::vector<TkWin> win(nw + 2);
    TkWin zero{};
    zero.lead = 0xFFFFFFFFu;
    win[0] = zero;          // window -1
    win[nw + 1] = zero;     // window nw
    for (uint64_t w = 0; w < nw; ++w) {
        uint32_t words[8] = {0};
        uint64_t pos = w * 32;
        uint64_t m = pos < n ? (n - pos < 32 ? n - pos : 32) : 0;
        memcpy(words, data + pos, m);
        win[w + 1] = tk_classify_window(data, n, pos, words, ds[w], T);
    }
    int64_t bad = -1;
    // serial "scan" of the run summaries
    TkRunSummary run{0, 0, 0, 0};
    std::vector<uint32_t> head(nw + 1, 2);
    std::vector<int> pend(nw, -1);
    TkDerived dprev{0, 0, 0, 0};
    for (uint64_t w = 0; w < nw; ++w) {
        const TkWin &p = win[w], &c = win[w + 1], &nx = win[w + 2];
        uint64_t valid = (w * 32 + 32 <= n) ? 0xFFFFFFFFull : ((1ull << (n - w * 32)) - 1ull);
        if ((c.bad & (uint32_t)valid) && bad < 0) bad = (int64_t)(w * 32) + TK_FFS(c.bad & (uint32_t)valid) - 1;
        // derived masks of the previous window recomputed from its class masks alone (as a GPU
        // thread does), to prove that suffices
        TkDerived dp = w ? tk_derive(data, n, (w - 1) * 32, zero, p, c, 0) : TkDerived{0, 0, 0, 0};
        (void)dprev;
        TkDerived dc = tk_derive(data, n, w * 32, p, c, nx, dp.sO);
        uint32_t n_in = run.n_val, abs_in = run.r_mode == 2 ? 0 : run.r_mode;
        TkEval ev = tk_eval_window(p, c, nx, dp, dc, n_in, abs_in);
        TkRunSummary s = tk_summarize(c);
        head[w] = s.head;
        pend[w] = ev.pend;
        // positions <= n only (n itself is the sentinel)
        uint64_t keep = (w * 32 + 32 <= n) ? 0xFFFFFFFFull : ((2ull << (n - w * 32)) - 1ull);
        start_mask[w] = ev.start & (uint32_t)keep;
        TkRunSummary st{0, run.n_val, run.r_mode == 2 ? 0u : run.r_mode, 0};
        run = tk_compose(st, s);
        dprev = dc;
    }
    for (uint64_t w = 0; w < nw; ++w) {
        if (pend[w] < 0) continue;
        uint32_t verdict = 2;  // end of data = the run ends
        for (uint64_t v = w + 2; v < nw; ++v)
            if (head[v]) { verdict = head[v]; break; }
        if (verdict == 2) start_mask[w] |= 1u << pend[w];
    }
    return bad >= 0 ? -1 - bad : (int64_t)nw;
}
