// tk_decode.cu -- batched Tekkenizer::decode (src/tekkenizer.rs:436-560) on the device.
//
// The reference walks the ids once, splitting them into maximal runs of special (< num_special)
// and ordinary ids; special runs are dropped / kept as their strings / rejected by policy
// (:529-544), ordinary runs are concatenated from the rank -> bytes table and must be valid UTF-8
// as a run (CoreBPE::decode -> String::from_utf8, :552-555).  Here:
//   D0 tokmark    sequence-start bitmask over id positions (validates tok_off)
//   D1 gather     per 2048-id tile: table cells (bytes + length), block scan, text assembled in shared memory,
//                 decoupled look-back prefix over tiles, sequence byte offsets, run-boundary bits,
//                 unknown-id / Raise errors, word copy to the output
//   D2 validate   one thread per 32 output bytes: strict UTF-8 with run boundaries, as flag arithmetic on words
//   D3 status     per sequence: the error the reference would have returned first
//   S'            one id list of at most 2,048 ids in a single block over mapped pinned memory (tk_decode's latency path)
#include "tk_kernels.h"

#include "../../include/tekken_b200.h"
#include "tk_device.cuh"
#include "tk_pretok.h"

namespace tkk {

#define DC_T 256

struct DocErr {
    unsigned long long unk_tok, sp_tok, sp_byte, utf_byte;
};

// Bounds-checked debug build (-DTK_DEBUG_BOUNDS, see tk_kernels.cu): every store of the decode kernels into the
// staging buffers, the output and the per-sequence arrays first compares its index with the array's size; a violation
// is skipped and recorded (line + 1,000,000 to tell it from an encode line, index, limit).  Compiled away otherwise.
#ifdef TK_DEBUG_BOUNDS
__device__ unsigned long long g_dc_hit[4];
__device__ __forceinline__ bool dc_in(unsigned long long i, unsigned long long n, int line) {
    if (i < n) return true;
    if (atomicAdd(&g_dc_hit[0], 1ull) == 0ull) { g_dc_hit[1] = 1000000ull + (unsigned long long)line; g_dc_hit[2] = i; g_dc_hit[3] = n; }
    return false;
}
#define DC_DBG(index, limit) dc_in((unsigned long long)(index), (unsigned long long)(limit), __LINE__)
#else
#define DC_DBG(index, limit) true
#endif

__global__ void tokmark_kernel(const uint64_t* __restrict__ tok_off, uint64_t off_base, uint64_t n_docs, uint64_t total,
                               uint32_t* __restrict__ tds, uint32_t* __restrict__ seq_first, uint32_t* __restrict__ flags) {
    uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    uint64_t o = tok_off[d] - off_base;      // offsets may be a slice of a larger batch (off_base = its first entry)
    bool ok = tok_off[d] >= off_base && o <= total;
    if (d == 0 && o != 0) ok = false;
    if (d == n_docs && o != total) ok = false;
    if (d < n_docs && tok_off[d + 1] < o) ok = false;
    if (!ok) { atomicOr(flags, TKK_FLAG_BAD_OFFSETS); return; }
    atomicOr(tds + (o >> 5), 1u << (o & 31));
    atomicMin(seq_first + (o >> 5), (uint32_t)d);     // first sequence that starts in this group of 32 ids
}

// index of the sequence that contains position i (last d < n_docs with off[d] <= i)
__device__ __forceinline__ uint64_t seq_of(const uint64_t* __restrict__ off, uint64_t off_base, uint64_t n_docs, uint64_t i) {
    uint64_t lo = 0, hi = n_docs;   // find first d with off[d] > i
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] - off_base <= i) lo = mid + 1; else hi = mid;
    }
    return lo ? lo - 1 : 0;
}

__device__ __forceinline__ void mark_boundary(uint32_t* __restrict__ bmask, uint64_t o, uint64_t cap) {
    if (o <= cap) atomicOr(bmask + (o >> 5), 1u << (o & 31));
}

// ---- D1: gather ------------------------------------------------------------------------------------------------------
// A tile is DC_T threads x PER consecutive ids.  Order of work inside a block:
//   1. ids -> the first 8 bytes of one 16-byte table cell per id (length + 7 bytes): the lengths for the scan and the
//      bytes for the copy come from the same load, issued for all of a thread's ids at once;
//   2. block scan of the lengths; the tile's total is published for the tiles behind it;
//   3. the text of the tile is assembled in (zeroed) shared memory at tile-relative positions.  A thread's tokens are
//      adjacent in the text: it shifts them into an accumulator word and stores every word it COMPLETES -- its own
//      bytes, zeros elsewhere -- with a plain predicated store; no branch depends on a token's length (round-2 ncu: the
//      branchy version spent 45 % of the kernel's instructions here at 5-18 active lanes);
//   4. after a barrier the bytes that are not in a completed word are OR-ed in: every thread's last partial word and
//      the tokens that do not come from a cell (longer than 15 bytes, special strings under Keep).  Every word has
//      exactly one plain store or none, and all ORs come after all stores;
//   5. only now warp 0 collects the prefix of the tiles before this one (decoupled look-back) -- by then it is usually
//      there, and the other warps were not waiting for it while they assembled;
//   6. the rare work that needs absolute positions (sequence starts, special and unknown ids) and the copy to the output
//      in aligned words, shifted by the output's phase.
// A tile whose text is larger than the staging buffer (more than 7 bytes per id) is written to the output byte by byte.

__device__ __forceinline__ uint32_t dc_slow_len(const TkDeviceTables& T, int policy, uint32_t v) {
    if (v < T.num_special) return policy == TK_POLICY_KEEP ? T.special_off[v + 1] - T.special_off[v] : 0u;
    const uint32_t r = v - T.num_special;
    return r < T.n_vocab ? T.vocab_off[r + 1] - T.vocab_off[r] : 0u;     // <= 65,535 (checked when the file is loaded)
}

// bytes of a token that is not in a table cell (longer than 15 bytes, or a special string under Keep); null if none
__device__ __forceinline__ const uint8_t* dc_slow_bytes(const TkDeviceTables& T, int policy, uint32_t v) {
    if (v < T.num_special) return policy == TK_POLICY_KEEP ? T.special_bytes + T.special_off[v] : nullptr;
    const uint32_t r = v - T.num_special;
    return r < T.n_vocab ? T.vocab_bytes + T.vocab_off[r] : nullptr;
}

struct DcStream {
    uint32_t a0;               // the word being filled: my bytes of it, zeros elsewhere
    uint32_t wi;               // its tile-relative index
    uint32_t fill;             // bytes of it that are decided (0..3 between calls)
};
// l <= 7 bytes in (v1:v0), zero beyond l; branch-free: up to two completed words are stored under predicates
#define DC_BUFW_OF(PER) (DC_T * (PER) * 7u / 4u + 4u)       // words of a tile's staging buffer
__device__ __forceinline__ void dc_append(uint32_t* __restrict__ bufw, DcStream& s, uint32_t v0, uint32_t v1, uint32_t l, uint32_t bufw_n) {
    const uint32_t sh = 8u * s.fill;
    s.a0 |= v0 << sh;
    const uint32_t a1 = __funnelshift_l(v0, v1, sh);        // bits 32..63 of (v1:v0) << sh
    const uint32_t a2 = __funnelshift_l(v1, 0u, sh);        // bits 64..95
    const uint32_t end = s.fill + l, c = end >> 2;          // words completed: 0, 1 or 2
    (void)bufw_n;
    if (c >= 1u && DC_DBG(s.wi, bufw_n)) bufw[s.wi] = s.a0;
    if (c >= 2u && DC_DBG(s.wi + 1u, bufw_n)) bufw[s.wi + 1u] = a1;
    s.a0 = c == 0u ? s.a0 : (c == 1u ? a1 : a2);
    s.wi += c;
    s.fill = end & 3u;
}
// l bytes that somebody else writes (zeros here)
__device__ __forceinline__ void dc_gap(uint32_t* __restrict__ bufw, DcStream& s, uint32_t l, uint32_t bufw_n) {
    const uint32_t end = s.fill + l;
    (void)bufw_n;
    if (end >= 4u) {
        if (DC_DBG(s.wi, bufw_n)) bufw[s.wi] = s.a0;         // my bytes of this word; the rest of it is the gap's
        s.a0 = 0u;
    }
    s.wi += end >> 2;
    s.fill = end & 3u;
}

template <int PER>
__device__ __forceinline__ void dc_locate(const uint32_t (&len16)[PER / 2], uint32_t k, uint32_t& rel, uint32_t& l) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {                          // bytes of my ids before position k; its length
        const uint32_t lj = (len16[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
        rel += (uint32_t)j < k ? lj : 0u;
        l = (uint32_t)j == k ? lj : l;
    }
}

#ifndef DC_PER
#define DC_PER 8
#endif
#ifndef DC_MINB
#define DC_MINB 5     // 48 registers: measured 3.35 ms for the bench batch (4 blocks, 62 registers: 3.56; 6 blocks, 40 registers and spills: 3.30)
#endif
template <int PER, int MINB>
__global__ void __launch_bounds__(DC_T, MINB) decode_gather_kernel(const uint32_t* __restrict__ ids, uint64_t n_ids,
                                                             const uint64_t* __restrict__ tok_off, uint64_t off_base, uint64_t n_docs,
                                                             const uint32_t* __restrict__ tds, const uint32_t* __restrict__ seq_first,
                                                             int policy, TkDeviceTables T,
                                                             uint8_t* __restrict__ out, uint64_t out_cap,
                                                             uint64_t* __restrict__ byte_off, uint32_t* __restrict__ bmask,
                                                             DocErr* __restrict__ docerr, unsigned long long* __restrict__ tile_state,
                                                             uint32_t* __restrict__ ticket, unsigned long long* __restrict__ total_out,
                                                             uint32_t* __restrict__ flags) {
    constexpr uint32_t TILE = DC_T * PER, BUF = TILE * 7u, BUFW = BUF / 4u + 4u;
    static_assert(BUFW % 4u == 0u && 32 % PER == 0 && PER % 4 == 0, "tile shape");
    __shared__ __align__(16) uint32_t bufw[BUFW];
    __shared__ uint32_t wsum[DC_T / 32];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_tile;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
    for (uint32_t c = t; c < BUFW / 4u; c += DC_T) reinterpret_cast<uint4*>(bufw)[c] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t i0 = (uint64_t)tile * TILE + (uint64_t)t * PER;
    const uint32_t tw = (__ldg(tds + (i0 >> 5)) >> (i0 & 31)) & ((1u << PER) - 1u);   // sequence starts among my ids (PER divides 32)
    // 1. the first half of one table cell per id: the length and 7 bytes.  Ids without a cell: special and unknown ids
    //    (`spec`: no bytes unless Keep, looked at again in step 6), positions past the end.
    uint2 e[PER];
    uint32_t spec = 0;
    {
        uint32_t raw[PER];
        if (i0 + PER <= n_ids && ((uintptr_t)ids & 15u) == 0) {
#pragma unroll
            for (int q = 0; q < PER / 4; ++q) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + i0) + q);
                raw[4 * q] = a.x; raw[4 * q + 1] = a.y; raw[4 * q + 2] = a.z; raw[4 * q + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) raw[k] = i0 + k < n_ids ? __ldg(ids + i0 + k) : 0u;     // 0 is a special id: no cell
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const uint32_t r = raw[k] - T.num_special;
            const bool cell = raw[k] >= T.num_special && r < T.n_vocab;
            e[k] = make_uint2(0u, 0u);
            if (cell) e[k] = __ldg(reinterpret_cast<const uint2*>(T.vocab_e16 + r));
            spec |= cell ? 0u : 1u << k;
        }
    }
    // 2. lengths (16 bits each), scan
    uint32_t len16[PER / 2];
#pragma unroll
    for (int k = 0; k < PER / 2; ++k) len16[k] = 0u;
    uint32_t slow = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const uint32_t l = e[k].y >> 24;
        slow |= l == 0xFFu ? 1u << k : 0u;
        len16[k >> 1] |= (l == 0xFFu ? 0u : l) << (16 * (k & 1));
    }
    if (policy == TK_POLICY_KEEP) slow |= spec;
    if (slow) {                                             // rare: lengths that are not in a cell
        uint32_t todo = slow;
        slow = 0;
#pragma unroll 1
        while (todo) {
            const uint32_t k = (uint32_t)(__ffs((int)todo) - 1);
            todo &= todo - 1;
            const uint32_t l = i0 + k < n_ids ? dc_slow_len(T, policy, __ldg(ids + i0 + k)) : 0u;
            if (!l) continue;
            slow |= 1u << k;
#pragma unroll
            for (int j = 0; j < PER; ++j) len16[j >> 1] |= (uint32_t)j == k ? l << (16 * (j & 1)) : 0u;
        }
    }
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < PER / 2; ++k) sum += (len16[k] & 0xFFFFu) + (len16[k] >> 16);
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < DC_T / 32; ++w) { if (w < (int)warp) before += wsum[w]; tile_total += wsum[w]; }
    if (t == 0) tk_lookback_publish(tile_state, tile, tile_total);
    const bool fits = tile_total <= BUF;                   // block-uniform
    const uint32_t p = before + inc - sum;                 // my first byte, tile-relative
    // 3. assemble: plain stores of completed words
    DcStream s;
    s.a0 = 0u; s.wi = p >> 2; s.fill = p & 3u;
    if (fits) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const uint32_t l = (len16[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
            if ((slow >> k) & 1u) {                        // its bytes come in step 4
                dc_gap(bufw, s, l, BUFW);
            } else {
                dc_append(bufw, s, e[k].x, e[k].y & 0x00FFFFFFu, l < 7u ? l : 7u, BUFW);
                if (l > 7u) {                               // the second half of the cell (a fraction of a percent of the ids)
                    const uint2 c = __ldg(reinterpret_cast<const uint2*>(T.vocab_e16 + (__ldg(ids + i0 + k) - T.num_special)) + 1);
                    dc_append(bufw, s, c.x, 0u, l < 11u ? l - 7u : 4u, BUFW);
                    dc_append(bufw, s, c.y, 0u, l < 11u ? 0u : l - 11u, BUFW);
                }
            }
        }
    }
    __syncthreads();
    // 4. OR in what is not part of a completed word
    if (fits) {
        if (s.fill && DC_DBG(s.wi, BUFW)) atomicOr(bufw + s.wi, s.a0);
        uint32_t todo = slow;
#pragma unroll 1
        while (todo) {                                      // tokens outside the cells: byte by byte
            const uint32_t k = (uint32_t)(__ffs((int)todo) - 1);
            todo &= todo - 1;
            uint32_t rel = p, l = 0;
            dc_locate<PER>(len16, k, rel, l);
            const uint8_t* src = dc_slow_bytes(T, policy, __ldg(ids + i0 + k));
            if (src)
                for (uint32_t j = 0; j < l; ++j)
                    if (DC_DBG((rel + j) >> 2, BUFW)) atomicOr(bufw + ((rel + j) >> 2), (uint32_t)__ldg(src + j) << (8u * ((rel + j) & 3u)));
        }
    }
    // 5. where the tile starts in the output
    if (warp == 0) {
        const unsigned long long excl = tk_lookback_collect(tile_state, tile, tile_total);
        if (lane == 0) {
            s_base = excl;
            if (tile == gridDim.x - 1) {
                *total_out = excl + tile_total;
                if (excl + tile_total > out_cap) atomicOr(flags, TKK_FLAG_OUT_FULL);
            }
        }
    }
    __syncthreads();
    const uint64_t base = s_base;
    // 6a. positions that need more than a copy: sequence starts (the end sentinel n_ids included), special and unknown ids;
    //     in an oversized tile every token (it is copied here)
    {
        uint64_t seq = tw ? seq_first[i0 >> 5] : 0;         // first sequence of my group of 32 ids; advanced below
        uint32_t todo = fits ? (tw | spec) : ((1u << PER) - 1u);
        // sequence starts among my positions and the one after them (bit PER)
        const uint32_t starts = todo ? tw | (((__ldg(tds + ((i0 + PER) >> 5)) >> ((i0 + PER) & 31)) & 1u) << PER) : 0u;
#pragma unroll 1
        while (todo) {
            const uint32_t k = (uint32_t)(__ffs((int)todo) - 1);
            todo &= todo - 1;
            const uint64_t i = i0 + k;
            if (i > n_ids) break;
            // the usual special id is the </s> that ends a sequence: dropped (Ignore), it sits on the byte where the next
            // sequence starts, and that start marks the run boundary -- nothing to do here (half of this loop's trips)
            if (fits && policy == TK_POLICY_IGNORE && ((starts >> k) & 3u) == 2u && i < n_ids && __ldg(ids + i) < T.num_special) continue;
            uint32_t rel = p, l = 0;
            dc_locate<PER>(len16, k, rel, l);
            const uint64_t ok = base + rel;
            if ((tw >> k) & 1u) {
                while (tok_off[seq] - off_base < i) ++seq;   // sequences that start earlier in the group
                for (; seq <= n_docs && tok_off[seq] - off_base == i; ++seq) if (DC_DBG(seq, n_docs + 1)) byte_off[seq] = ok;
                mark_boundary(bmask, ok, out_cap);
            }
            if (i == n_ids) break;
            const uint32_t v = __ldg(ids + i);
            if (v < T.num_special) {
                // a special id ends the ordinary run before it and starts a new one after it
                if (!((tw >> k) & 1u)) mark_boundary(bmask, ok, out_cap);
                if (policy == TK_POLICY_RAISE) {
                    const uint64_t d = seq_of(tok_off, off_base, n_docs, i);
                    atomicMin(&docerr[d].sp_tok, (unsigned long long)i);
                    atomicMin(&docerr[d].sp_byte, (unsigned long long)ok);
                } else if (policy == TK_POLICY_KEEP) {
                    mark_boundary(bmask, ok + l, out_cap);
                }
            } else if (v - T.num_special >= T.n_vocab) {
                const uint64_t d = seq_of(tok_off, off_base, n_docs, i);
                atomicMin(&docerr[d].unk_tok, (unsigned long long)i);
            }
            if (!fits && l && ok + l <= out_cap) {
                const uint8_t* src = dc_slow_bytes(T, policy, v);
                if (src)
                    for (uint32_t j = 0; j < l; ++j) if (DC_DBG(ok + j, out_cap)) out[ok + j] = __ldg(src + j);
            }
        }
    }
    if (!fits) return;
    // 6b. staging buffer -> output: aligned words, the tile's text shifted by the phase of its first byte
    {
        const uint64_t lim = base + tile_total < out_cap ? base + tile_total : out_cap;   // never write past the caller's buffer
        if (lim <= base) return;
        const uint32_t sh = (uint32_t)(((uintptr_t)out + base) & 3u);
        uint8_t* const o0 = out + base - sh;                    // word aligned; output word j = tile bytes 4j - sh .. 4j - sh + 3
        const uint32_t n_bytes = (uint32_t)(lim - base) + sh;   // bytes from o0 to the end of the tile's text
        const uint32_t n_full = n_bytes >> 2;                   // words 1 .. n_full - 1 are whole words of this tile
        const uint32_t fs = 8u * (4u - sh);
        uint32_t* const ow = reinterpret_cast<uint32_t*>(o0);
        // (debug build: the word's last byte must be inside the caller's buffer, and the staging index inside its array)
        if (sh) {
            for (uint32_t j = 1u + t; j < n_full; j += DC_T)
                if (DC_DBG(base - sh + 4ull * j + 3u, out_cap) && DC_DBG(j, BUFW)) ow[j] = __funnelshift_r(bufw[j - 1], bufw[j], fs);
        } else {
            for (uint32_t j = 1u + t; j < n_full; j += DC_T)
                if (DC_DBG(base + 4ull * j + 3u, out_cap) && DC_DBG(j, BUFW)) ow[j] = bufw[j];
        }
        // the two ragged ends, byte by byte: word 0 (bytes sh..3) and the bytes after the last whole word
        if (t < 8u) {
            const uint32_t q = t < 4u ? t : 4u * n_full + (t - 4u);      // byte index from o0
            const bool mine = t < 4u ? (q >= sh && q < n_bytes) : (n_full >= 1u && q < n_bytes);
            if (mine && DC_DBG(base - sh + q, out_cap) && DC_DBG((q - sh) >> 2, BUFW)) o0[q] = reinterpret_cast<const uint8_t*>(bufw)[q - sh];
        }
    }
}

// ---- D2: strict UTF-8 per ordinary run ----------------------------------------------------------------------------------
// One thread per 32 output bytes, nine words (the window and the four bytes before it), flags kept where the bytes are
// (bit 7 of every byte of a word; "the byte before" is a funnel shift by 8 across two words) -- no per-character loop
// and no table: a window of CJK text costs what a window of Latin text costs.
//   lead2/3/4 bytes start a chain of continuation bytes; a continuation byte ON a run boundary does not continue
//   anything (contp);
//   errA (charged to the run that holds the byte): a continuation byte no chain reaches, C0 C1 F5..FF, a second byte
//     outside the range its lead allows (E0: A0..BF, ED: 80..9F, F0: 90..BF, F4: 80..8F);
//   errB (charged to the run that holds the byte BEFORE): a chain expects a continuation byte here and there is none --
//     the text, the run or the character just ends.
// Which sequence an error belongs to is looked up only when there is one.
struct DvPrev {
    uint32_t l234, l34, l4, c1_34, c1_4, c2_4, e0, ed, f0, f4;
};

__device__ __forceinline__ uint64_t dv_seq_of_byte(const uint64_t* __restrict__ byte_off, uint64_t n_docs, uint64_t q) {
    uint64_t lo = 0, hi = n_docs;                    // last d with byte_off[d] <= q
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (byte_off[mid] <= q) lo = mid + 1; else hi = mid;
    }
    return lo ? lo - 1 : 0;
}

// The nine words of one window.  REPORT = false: is anything wrong (the hot path); true: charge every error to its
// sequence.
template <bool REPORT>
__device__ __forceinline__ uint32_t dv_window(const uint32_t (&w)[9], uint32_t bm, uint32_t bm0, uint64_t pos, uint64_t n,
                                              const uint64_t* __restrict__ byte_off, uint64_t n_docs, DocErr* __restrict__ docerr) {
    DvPrev pv = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    uint32_t any_err = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint32_t x = w[j];
        if (!((x & TK_H) | ((pv.l234 | pv.c1_34 | pv.c2_4) >> 24))) {      // ASCII, and nothing reaches into it
            pv = DvPrev{0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            continue;
        }
        const uint32_t bn = j ? (bm >> (4 * (j - 1))) & 0xFu : bm0;
        const uint32_t bnd = ((bn * 0x00204081u) & 0x01010101u) << 7;       // boundary bits -> bit 7 of their bytes
        const uint32_t c6 = x << 1, c5 = x << 2, c4 = x << 3;
        const uint32_t hi = x & TK_H;
        const uint32_t cont = hi & ~c6, contp = cont & ~bnd;
        const uint32_t a = x & c6;
        DvPrev c;
        c.l234 = a & TK_H;                       // >= C0
        c.l34 = a & c5 & TK_H;                   // >= E0
        c.l4 = c.l34 & c4;                       // >= F0
        const uint32_t e1 = __funnelshift_l(pv.l234, c.l234, 8);
        c.c1_34 = __funnelshift_l(pv.l34, c.l34, 8) & contp;
        c.c1_4 = __funnelshift_l(pv.l4, c.l4, 8) & contp;
        const uint32_t t1 = __funnelshift_l(pv.c1_34, c.c1_34, 8);
        c.c2_4 = __funnelshift_l(pv.c1_4, c.c1_4, 8) & contp;
        const uint32_t t2 = __funnelshift_l(pv.c2_4, c.c2_4, 8);
        const uint32_t expect = e1 | t1 | t2;
        uint32_t A = cont & ~(expect & contp);                              // stray continuation bytes
        const uint32_t B = expect & ~contp;                                 // a continuation byte is missing here
        // C0, C1: two-byte leads with bits 4..1 clear
        A |= c.l234 & ~c5 & ~(((x & 0x1E1E1E1Eu) + 0x7F7F7F7Fu) & TK_H);
        c.e0 = c.ed = c.f0 = c.f4 = 0u;
        if (c.l34 | pv.e0 | pv.ed | pv.f0 | pv.f4) {
            const uint32_t w7 = x & 0x7F7F7F7Fu;
            c.e0 = tk_swar_eq(w7, 0x60u) & hi;
            c.ed = tk_swar_eq(w7, 0x6Du) & hi;
            if (c.l4) {
                c.f0 = tk_swar_eq(w7, 0x70u) & hi;
                c.f4 = tk_swar_eq(w7, 0x74u) & hi;
                A |= tk_swar_ge(w7, 0x75u) & hi;                            // F5..FF
            }
            A |= __funnelshift_l(pv.e0, c.e0, 8) & contp & ~c5;             // E0 80..9F: overlong
            A |= __funnelshift_l(pv.ed, c.ed, 8) & contp & c5;              // ED A0..BF: surrogates
            A |= __funnelshift_l(pv.f0, c.f0, 8) & contp & ~c5 & ~c4;       // F0 80..8F: overlong
            A |= __funnelshift_l(pv.f4, c.f4, 8) & contp & (c5 | c4);       // F4 90..BF: above U+10FFFF
        }
        pv = c;
        if (j) {
            any_err |= (A | B) & TK_H;
            if (REPORT) {
                uint32_t na = tk_swar_nib(A & TK_H), nb = tk_swar_nib(B & TK_H);
                while (na) {
                    const uint64_t q = pos + 4u * (uint32_t)(j - 1) + (uint32_t)(__ffs((int)na) - 1);
                    na &= na - 1;
                    if (q < n) atomicMin(&docerr[dv_seq_of_byte(byte_off, n_docs, q)].utf_byte, (unsigned long long)q);
                }
                while (nb) {
                    const uint64_t q = pos + 4u * (uint32_t)(j - 1) + (uint32_t)(__ffs((int)nb) - 1);
                    nb &= nb - 1;
                    if (q >= 1 && q <= n) atomicMin(&docerr[dv_seq_of_byte(byte_off, n_docs, q - 1)].utf_byte, (unsigned long long)(q - 1));
                }
            }
        }
    }
    return any_err;
}

__global__ void __launch_bounds__(256, 4) decode_validate_kernel(const uint8_t* __restrict__ out, const unsigned long long* __restrict__ total_out,
                                                              uint64_t out_cap, const uint32_t* __restrict__ bmask,
                                                              const uint64_t* __restrict__ byte_off, uint64_t n_docs,
                                                              DocErr* __restrict__ docerr) {
    uint64_t n = *total_out;
    if (n > out_cap) n = out_cap;
    const uint64_t n_windows = n / 32 + 1;           // position n itself is looked at: a character cut off by the end of the text
    const bool aligned = ((uintptr_t)out & 15u) == 0;
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_windows; wi += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t pos = wi * 32u;
        uint32_t w[9];
        if (pos + 32 <= n && aligned) {
            const uint4 a = __ldg((const uint4*)(out + pos));
            const uint4 b = __ldg((const uint4*)(out + pos) + 1);
            w[1] = a.x; w[2] = a.y; w[3] = a.z; w[4] = a.w; w[5] = b.x; w[6] = b.y; w[7] = b.z; w[8] = b.w;
            w[0] = pos ? __ldg((const uint32_t*)(out + pos) - 1) : 0u;
        } else {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t q = pos + 4 * j + k;         // byte q - 4
                    if (q >= 4 && q - 4 < n) v |= (uint32_t)out[q - 4] << (8 * k);
                }
                w[j] = v;
            }
        }
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) any |= w[j];
        if (!(any & TK_H)) continue;                             // ASCII: nothing can be wrong
        const uint32_t bm = bmask[wi], bm0 = wi ? bmask[wi - 1] >> 28 : 0u;
        if (!dv_window<false>(w, bm, bm0, pos, n, byte_off, n_docs, docerr)) continue;
        dv_window<true>(w, bm, bm0, pos, n, byte_off, n_docs, docerr);         // rare: say where, and whose
    }
}

__global__ void decode_status_kernel(const DocErr* __restrict__ docerr, uint64_t n_docs, int32_t* __restrict__ status,
                                     unsigned long long* __restrict__ first_bad) {
    const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    const DocErr e = docerr[d];
    int32_t s = TK_OK;
    const bool ordinary_fail = e.unk_tok != ~0ull || e.utf_byte != ~0ull;
    if (e.sp_tok != ~0ull) {
        // Raise: the special run fails unless an ordinary run before it already failed
        const bool earlier = (e.unk_tok != ~0ull && e.unk_tok < e.sp_tok) || (e.utf_byte != ~0ull && e.utf_byte < e.sp_byte);
        s = earlier ? TK_ERR_TOKENIZERS : TK_ERR_SPECIAL_TOKEN_POLICY;
    } else if (ordinary_fail) {
        s = TK_ERR_TOKENIZERS;
    }
    if (status) status[d] = s;
    if (s != TK_OK) atomicMin(first_bad, (unsigned long long)d);
}

// ---- the latency path: one sequence of at most kSmallDecodeIds ids in ONE single-block kernel ---------------------------
// (the reference's own call shape: Tekkenizer::decode of one id list).  The ids come from and the text goes to mapped
// pinned memory that the calling thread polls: no copies, no memsets, no stream synchronisation.  Same arithmetic as
// D0-D3 -- lengths from the table cells, block scan, run boundaries, strict UTF-8 per ordinary run (dv_window), the
// status rule of decode_status_kernel -- on one tile in shared memory.
#define DS_T 256
#define DS_PER (kSmallDecodeIds / DS_T)
struct DsSmem {
    uint32_t bufw[kSmallDecodeBytes / 4 + 8];
    uint32_t bmask[kSmallDecodeBytes / 32 + 4];
    uint32_t wsum[DS_T / 32];
    uint64_t boff[2];
    DocErr err;
};
struct DsResult {
    uint32_t n_bytes, status, flags, pad;
    uint32_t done;                          // == the call's sequence number when everything above and the text are visible
    uint32_t pad2[3];
};

__global__ void __launch_bounds__(DS_T) decode_small_kernel(const uint32_t* __restrict__ ids, uint32_t n, int policy, TkDeviceTables T,
                                                            uint32_t* __restrict__ out, uint32_t seq) {
    __shared__ __align__(16) DsSmem S;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    for (uint32_t i = t; i < sizeof(S.bufw) / 4; i += DS_T) S.bufw[i] = 0u;
    for (uint32_t i = t; i < sizeof(S.bmask) / 4; i += DS_T) S.bmask[i] = 0u;
    if (t == 0) { S.err.unk_tok = S.err.sp_tok = S.err.sp_byte = S.err.utf_byte = ~0ull; S.boff[0] = 0; }
    uint32_t v[DS_PER], len[DS_PER];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < DS_PER; ++k) {
        const uint32_t i = t * DS_PER + k;
        v[k] = i < n ? ids[i] : 0u;
        uint32_t l = 0;
        if (i < n) {
            if (v[k] < T.num_special) l = policy == TK_POLICY_KEEP ? T.special_off[v[k] + 1] - T.special_off[v[k]] : 0u;
            else {
                const uint32_t r = v[k] - T.num_special;
                if (r < T.n_vocab) {
                    l = __ldg(reinterpret_cast<const uint32_t*>(T.vocab_e16 + r) + 1) >> 24;
                    if (l == 0xFFu) l = T.vocab_off[r + 1] - T.vocab_off[r];
                }
            }
        }
        len[k] = l;
        sum += l;
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    if (lane == 31) S.wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < DS_T / 32; ++w) { if (w < (int)warp) before += S.wsum[w]; total += S.wsum[w]; }
    DsResult* res = reinterpret_cast<DsResult*>(out);
    if (total > kSmallDecodeBytes) {                       // block-uniform: more text than the tile holds -> the batch path
        if (t == 0) {
            res->n_bytes = 0; res->status = 0; res->flags = kSmallNeedBatch;
            __threadfence_system();
            *(volatile uint32_t*)&res->done = seq;
        }
        return;
    }
    uint8_t* buf = reinterpret_cast<uint8_t*>(S.bufw);
    uint32_t o = before + inc - sum;
    if (t == 0) { atomicOr(&S.bmask[0], 1u); atomicOr(&S.bmask[total >> 5], 1u << (total & 31u)); S.boff[1] = total; }
#pragma unroll 1
    for (int k = 0; k < DS_PER; ++k) {
        const uint32_t i = t * DS_PER + k;
        if (i >= n) break;
        const uint32_t l = len[k];
        const uint8_t* src = nullptr;
        if (v[k] < T.num_special) {
            // a special id ends the ordinary run before it and starts a new one after it
            atomicOr(&S.bmask[o >> 5], 1u << (o & 31u));
            if (policy == TK_POLICY_RAISE) { atomicMin(&S.err.sp_tok, (unsigned long long)i); atomicMin(&S.err.sp_byte, (unsigned long long)o); }
            else if (policy == TK_POLICY_KEEP) { atomicOr(&S.bmask[(o + l) >> 5], 1u << ((o + l) & 31u)); src = T.special_bytes + T.special_off[v[k]]; }
        } else {
            const uint32_t r = v[k] - T.num_special;
            if (r >= T.n_vocab) atomicMin(&S.err.unk_tok, (unsigned long long)i);
            else src = T.vocab_bytes + T.vocab_off[r];
        }
        if (src)
            for (uint32_t j = 0; j < l; ++j) if (DC_DBG(o + j, kSmallDecodeBytes)) buf[o + j] = __ldg(src + j);
        o += l;
    }
    __syncthreads();
    // strict UTF-8 per ordinary run
    for (uint32_t wi = t; wi < total / 32u + 1u; wi += DS_T) {
        uint32_t w[9];
        w[0] = wi ? S.bufw[wi * 8u - 1u] : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j + 1] = S.bufw[wi * 8u + j];          // zero beyond the text
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) any |= w[j];
        if (!(any & TK_H)) continue;
        const uint32_t bm = S.bmask[wi], bm0 = wi ? S.bmask[wi - 1u] >> 28 : 0u;
        if (dv_window<false>(w, bm, bm0, (uint64_t)wi * 32u, (uint64_t)total, S.boff, 1, &S.err))
            dv_window<true>(w, bm, bm0, (uint64_t)wi * 32u, (uint64_t)total, S.boff, 1, &S.err);
    }
    __syncthreads();
    for (uint32_t j = t; j < (total + 3u) / 4u; j += DS_T) if (DC_DBG(j, kSmallDecodeBytes / 4u)) out[8 + j] = S.bufw[j];
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        const DocErr e = S.err;
        int32_t st = TK_OK;
        if (e.sp_tok != ~0ull) {
            const bool earlier = (e.unk_tok != ~0ull && e.unk_tok < e.sp_tok) || (e.utf_byte != ~0ull && e.utf_byte < e.sp_byte);
            st = earlier ? TK_ERR_TOKENIZERS : TK_ERR_SPECIAL_TOKEN_POLICY;
        } else if (e.unk_tok != ~0ull || e.utf_byte != ~0ull) st = TK_ERR_TOKENIZERS;
        res->n_bytes = total; res->status = (uint32_t)st; res->flags = 0;
        __threadfence_system();
        *(volatile uint32_t*)&res->done = seq;
    }
}

cudaError_t decode_small(const TkDeviceTables& T, const uint32_t* d_ids, uint32_t n, int policy, uint32_t* d_out, uint32_t seq, cudaStream_t st) {
    decode_small_kernel<<<1, DS_T, 0, st>>>(d_ids, n, policy, T, d_out, seq);
    count_launch();
    return cudaGetLastError();
}

long long decode_debug_bounds_violations(unsigned long long* out4) {
#ifdef TK_DEBUG_BOUNDS
    unsigned long long h[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(h, g_dc_hit, sizeof h) != cudaSuccess) return -2;
    cudaMemcpyToSymbol(g_dc_hit, z, sizeof z);
    if (out4) for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return (long long)h[0];
#else
    (void)out4;
    return -1;
#endif
}

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

size_t decode_workspace_bytes(uint64_t n_ids, uint64_t n_docs, uint64_t out_cap, DecodeLayout* L) {
    DecodeLayout l{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    l.n_tiles = ceil_div(n_ids + 1, DC_T * DC_PER);
    l.mask_words_tok = l.n_tiles * (DC_T * DC_PER / 32) + 8;
    l.mask_words_out = out_cap / 32 + 8;
    l.off_small = take(256);
    l.off_tds = take(l.mask_words_tok * 4);
    l.off_seqfirst = take(l.mask_words_tok * 4);
    l.off_bmask = take(l.mask_words_out * 4);
    l.off_tilestate = take(l.n_tiles * 8);
    l.off_docerr = take((n_docs + 1) * sizeof(DocErr));
    l.total = off;
    if (L) *L = l;
    return off;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t decode_device(const TkDeviceTables& T, const uint32_t* d_ids, const uint64_t* d_tok_off, uint64_t off_base, uint64_t n_docs,
                          uint64_t n_ids, int policy, uint8_t* d_out, uint64_t out_cap, uint64_t* d_byte_off,
                          int32_t* d_doc_status, void* d_ws, const DecodeLayout& L, cudaStream_t st) {
    unsigned char* ws = (unsigned char*)d_ws;
    uint32_t* small = (uint32_t*)(ws + L.off_small);
    uint32_t* tds = (uint32_t*)(ws + L.off_tds);
    uint32_t* seq_first = (uint32_t*)(ws + L.off_seqfirst);
    uint32_t* bmask = (uint32_t*)(ws + L.off_bmask);
    unsigned long long* tilestate = (unsigned long long*)(ws + L.off_tilestate);
    DocErr* docerr = (DocErr*)(ws + L.off_docerr);
    uint32_t* flags = small + TKK_S_FLAGS;
    uint32_t* ticket = small + TKK_S_TICKET;
    unsigned long long* total_out = (unsigned long long*)(small + TKK_S_TOTAL);
    unsigned long long* first_bad = (unsigned long long*)(small + TKK_S_BADDOC);
    CK(cudaMemsetAsync(small, 0, 256, st));
    CK(cudaMemsetAsync(first_bad, 0xFF, 8, st));
    CK(cudaMemsetAsync(tds, 0, L.mask_words_tok * 4, st));
    CK(cudaMemsetAsync(seq_first, 0xFF, L.mask_words_tok * 4, st));
    CK(cudaMemsetAsync(bmask, 0, L.mask_words_out * 4, st));
    CK(cudaMemsetAsync(tilestate, 0, L.n_tiles * 8, st));
    CK(cudaMemsetAsync(docerr, 0xFF, (n_docs + 1) * sizeof(DocErr), st));
    tokmark_kernel<<<(unsigned)ceil_div(n_docs + 1, 256), 256, 0, st>>>(d_tok_off, off_base, n_docs, n_ids, tds, seq_first, flags);
    count_launch();
    decode_gather_kernel<DC_PER, DC_MINB><<<(unsigned)L.n_tiles, DC_T, 0, st>>>(d_ids, n_ids, d_tok_off, off_base, n_docs, tds, seq_first, policy, T, d_out, out_cap,
                                                             d_byte_off, bmask, docerr, tilestate, ticket, total_out, flags);
    count_launch();
    {
        uint64_t blocks = ceil_div(out_cap / 32 + 1, 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        decode_validate_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, total_out, out_cap, bmask, d_byte_off, n_docs, docerr);
        count_launch();
    }
    if (n_docs) {
        decode_status_kernel<<<(unsigned)ceil_div(n_docs, 256), 256, 0, st>>>(docerr, n_docs, d_doc_status, first_bad);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace tkk
