// tk_decode.cu -- batched Tekkenizer::decode (src/tekkenizer.rs:436-560) on the device.
//
// The reference walks the ids once, splitting them into maximal runs of special (< num_special)
// and ordinary ids; special runs are dropped / kept as their strings / rejected by policy
// (:529-544), ordinary runs are concatenated from the rank -> bytes table and must be valid UTF-8
// as a run (CoreBPE::decode -> String::from_utf8, :552-555).  Here:
//   D0 tokmark    sequence-start bitmask over id positions (validates tok_off)
//   D1 gather     per 2048-id tile: lengths, block scan, decoupled look-back prefix over tiles,
//                 byte copy, sequence byte offsets, run-boundary bits, unknown-id / Raise errors
//   D2 validate   one thread per 32 output bytes: strict UTF-8 with run boundaries (same window
//                 classifier the encoder uses)
//   D3 status     per sequence: the error the reference would have returned first
#include "tk_kernels.h"

#include "../../include/tekken_b200.h"
#include "tk_device.cuh"
#include "tk_pretok.h"

namespace tkk {

#define DC_T 256
#define DC_PER 8
#define DC_TILE (DC_T * DC_PER)

struct DocErr {
    unsigned long long unk_tok, sp_tok, sp_byte, utf_byte;
};

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

#define DC_BUF 12288      // bytes of a tile's text staged in shared memory (larger tiles write straight to global memory)

// Copy l bytes of a token into the tile's staging buffer or, for an oversized tile, to the output.
__device__ __forceinline__ void dc_put(uint8_t* __restrict__ buf, bool fits, uint32_t p, uint8_t* __restrict__ out, uint64_t o,
                                       uint64_t out_cap, const uint8_t* __restrict__ src, uint32_t l) {
    if (fits) {
        for (uint32_t j = 0; j < l; ++j) buf[p + j] = __ldg(src + j);
    } else if (o + l <= out_cap) {
        for (uint32_t j = 0; j < l; ++j) out[o + j] = __ldg(src + j);
    }
}

#ifndef DC_MINB
#define DC_MINB 8      // 32 registers (measured: 5.4 ms; 5.8 ms at 40 registers, 7.8 ms uncapped at 71)
#endif
__global__ void __launch_bounds__(DC_T, DC_MINB) decode_gather_kernel(const uint32_t* __restrict__ ids, uint64_t n_ids,
                                                             const uint64_t* __restrict__ tok_off, uint64_t off_base, uint64_t n_docs,
                                                             const uint32_t* __restrict__ tds, const uint32_t* __restrict__ seq_first,
                                                             int policy, TkDeviceTables T,
                                                             uint8_t* __restrict__ out, uint64_t out_cap,
                                                             uint64_t* __restrict__ byte_off, uint32_t* __restrict__ bmask,
                                                             DocErr* __restrict__ docerr, unsigned long long* __restrict__ tile_state,
                                                             uint32_t* __restrict__ ticket, unsigned long long* __restrict__ total_out,
                                                             uint32_t* __restrict__ flags) {
    __shared__ __align__(16) uint8_t buf[DC_BUF + 32];
    __shared__ uint32_t wsum[DC_T / 32];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_tile;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t i0 = (uint64_t)tile * DC_TILE + (uint64_t)t * DC_PER;
    uint32_t id[DC_PER], len[DC_PER];
    uint32_t sum = 0;
    {
        // 8 consecutive ids per thread: two 16-byte loads when they are all there
        uint32_t raw[DC_PER];
        if (i0 + DC_PER <= n_ids && ((uintptr_t)ids & 15u) == 0) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + i0)), b = __ldg(reinterpret_cast<const uint4*>(ids + i0) + 1);
            raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w; raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < DC_PER; ++k) raw[k] = i0 + k < n_ids ? __ldg(ids + i0 + k) : 0u;
        }
#pragma unroll
        for (int k = 0; k < DC_PER; ++k) {
            const uint64_t i = i0 + k;
            uint32_t l = 0;
            const uint32_t v = raw[k];
            if (i < n_ids) {
                if (v < T.num_special) l = policy == TK_POLICY_KEEP ? T.special_off[v + 1] - T.special_off[v] : 0u;
                else {
                    const uint32_t r = v - T.num_special;
                    l = r < T.n_vocab ? (uint32_t)__ldg(T.vocab_len + r) : 0u;
                    if (l == 255u) l = T.vocab_off[r + 1] - T.vocab_off[r];
                }
            }
            id[k] = v; len[k] = l; sum += l;
        }
    }
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
    if (warp == 0) {
        const unsigned long long excl = tk_lookback(tile_state, tile, tile_total);
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
    const bool fits = tile_total <= DC_BUF;              // block-uniform
    const uint32_t shift = (uint32_t)(base & 15u);         // staging keeps the output's 16-byte phase
    uint32_t p = shift + before + inc - sum;               // my first byte in the staging buffer
    uint64_t o = base + before + inc - sum;
    // sequence starts among my positions (the sentinel position n_ids included)
    const uint32_t tw = tds[i0 >> 5] >> (i0 & 31);   // DC_PER divides 32 -> my 8 bits are in one word
    uint64_t seq = (tw & 0xFFu) ? seq_first[i0 >> 5] : 0;   // first sequence of my group of 32 ids; advanced below
    // Usual case, unrolled: an ordinary token of at most 16 bytes, no sequence start -> one aligned 16-byte
    // load from the padded table.  Everything else (sequence starts, special and unknown ids, longer
    // tokens, oversized tiles, the sentinel position) is noted in `slow` and handled by one compact loop
    // below, so the kernel stays small enough for the instruction cache.
    uint32_t slow = 0;
    {
        uint32_t pk = p;
#pragma unroll
        for (int k = 0; k < DC_PER; ++k) {
            const uint64_t i = i0 + k;
            const uint32_t v = id[k], l = len[k];
            const uint32_t r = v - T.num_special;
            const bool fast = fits && i < n_ids && !((tw >> k) & 1u) && v >= T.num_special && r < T.n_vocab && l <= 16u;
            if (fast) {
                const uint4 q = __ldg(T.vocab_pad16 + r);
                uint32_t cur = q.x;
                for (uint32_t j = 0; j < l; ++j) {
                    if ((j & 3u) == 0u && j) cur = j == 4u ? q.y : j == 8u ? q.z : q.w;
                    buf[pk + j] = (uint8_t)cur;
                    cur >>= 8;
                }
            } else if (i <= n_ids) slow |= 1u << k;
            pk += l;
        }
    }
#pragma unroll 1
    while (slow) {
        const uint32_t k = (uint32_t)(__ffs((int)slow) - 1);
        slow &= slow - 1;
        const uint64_t i = i0 + k;
        uint32_t pre = 0, v = 0, l = 0;                      // bytes of my ids before position k; its id and length
#pragma unroll
        for (int j = 0; j < DC_PER; ++j) {
            pre += (uint32_t)j < k ? len[j] : 0u;
            if ((uint32_t)j == k) { v = id[j]; l = len[j]; }
        }
        const uint64_t ok = o + pre;
        const uint32_t pk = p + pre;
        if ((tw >> k) & 1u) {
            while (tok_off[seq] - off_base < i) ++seq;   // sequences that start earlier in the group
            for (; seq <= n_docs && tok_off[seq] - off_base == i; ++seq) byte_off[seq] = ok;
            mark_boundary(bmask, ok, out_cap);
        }
        if (i == n_ids) break;
        if (v < T.num_special) {
            // a special id ends the ordinary run before it and starts a new one after it
            mark_boundary(bmask, ok, out_cap);
            if (policy == TK_POLICY_RAISE) {
                const uint64_t d = seq_of(tok_off, off_base, n_docs, i);
                atomicMin(&docerr[d].sp_tok, (unsigned long long)i);
                atomicMin(&docerr[d].sp_byte, (unsigned long long)ok);
            } else if (policy == TK_POLICY_KEEP) {
                dc_put(buf, fits, pk, out, ok, out_cap, T.special_bytes + T.special_off[v], l);
                mark_boundary(bmask, ok + l, out_cap);
            }
        } else {
            const uint32_t r = v - T.num_special;
            if (r >= T.n_vocab) {
                const uint64_t d = seq_of(tok_off, off_base, n_docs, i);
                atomicMin(&docerr[d].unk_tok, (unsigned long long)i);
            } else {
                dc_put(buf, fits, pk, out, ok, out_cap, T.vocab_bytes + T.vocab_off[r], l);
            }
        }
    }
    if (!fits) return;
    __syncthreads();
    // staging buffer -> output: aligned 16-byte stores in the middle, single bytes at the ragged ends
    {
        const uint64_t lim = base + tile_total < out_cap ? base + tile_total : out_cap;   // never write past the caller's buffer
        if (lim <= base) return;
        const uint64_t a0 = base - shift;                       // 16-byte aligned global address of buf[0]
        const uint32_t n_chunks = (uint32_t)((lim - a0 + 15u) / 16u);
        const bool aligned = ((uintptr_t)out & 15u) == 0;
        for (uint32_t c = t; c < n_chunks; c += DC_T) {
            const uint64_t g = a0 + 16ull * c;
            if (aligned && g >= base && g + 16u <= lim) {
                *reinterpret_cast<uint4*>(out + g) = *reinterpret_cast<const uint4*>(buf + 16u * c);
            } else {
                for (uint32_t j = 0; j < 16u; ++j)
                    if (g + j >= base && g + j < lim) out[g + j] = buf[16u * c + j];
            }
        }
    }
}

__global__ void __launch_bounds__(256) decode_validate_kernel(const uint8_t* __restrict__ out, const unsigned long long* __restrict__ total_out,
                                                              uint64_t out_cap, const uint32_t* __restrict__ bmask,
                                                              const uint64_t* __restrict__ byte_off, uint64_t n_docs, TkDeviceTables T,
                                                              DocErr* __restrict__ docerr) {
    uint64_t n = *total_out;
    if (n > out_cap) n = out_cap;
    const uint64_t n_windows = (n + 31) / 32;
    for (uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_windows; wi += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t pos = wi * 32u;
        uint32_t w[8];
        if (pos + 32 <= n && ((uintptr_t)out & 15u) == 0) {
            const uint4 a = __ldg((const uint4*)(out + pos));
            const uint4 b = __ldg((const uint4*)(out + pos) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint64_t q = pos + 4 * j + k;
                    if (q < n) v |= (uint32_t)out[q] << (8 * k);
                }
                w[j] = v;
            }
        }
        // a char must not continue across a run boundary: the classifier flags a boundary on a
        // continuation byte and, via the byte after the window, a truncated char
        const TkWin c = tk_classify_window(out, n, pos, w, bmask[wi], T);
        const uint32_t valid = (pos + 32 <= n) ? 0xFFFFFFFFu : (uint32_t)((1ull << (n - pos)) - 1ull);
        uint32_t bad = c.bad & valid;
        // a boundary on a continuation byte: the run after it is flagged above (it starts inside a
        // char).  The run BEFORE it is invalid too iff its last char is cut short by the boundary.
        uint32_t split = c.ds & ~c.lead & valid;
        while (split) {
            const uint64_t q = pos + (uint32_t)(__ffs((int)split) - 1);
            split &= split - 1;
            if (q == 0) continue;
            // lead byte of the char that contains byte q-1, without crossing an earlier boundary
            uint64_t k = q - 1;
            int back = 0;
            bool crossed = false;
            while (back < 3 && k > 0 && (out[k] & 0xC0u) == 0x80u) {
                if ((bmask[k >> 5] >> (k & 31)) & 1u) { crossed = true; break; }
                --k;
                ++back;
            }
            if (crossed) continue;                                  // that run starts with a continuation byte: flagged already
            const uint32_t b0 = out[k];
            const uint32_t need = b0 < 0x80u ? 1u : b0 >= 0xF0u ? 4u : b0 >= 0xE0u ? 3u : b0 >= 0xC0u ? 2u : 0u;
            if (need == 0u || k + need <= q) continue;              // complete (or stray bytes, flagged on their own)
            uint64_t lo = 0, hi = n_docs;
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if (byte_off[mid] <= q - 1) lo = mid + 1; else hi = mid;
            }
            atomicMin(&docerr[lo ? lo - 1 : 0].utf_byte, (unsigned long long)(q - 1));
        }
        while (bad) {
            const uint64_t q = pos + (uint32_t)(__ffs((int)bad) - 1);
            bad &= bad - 1;
            // sequence containing byte q: last d with byte_off[d] <= q
            uint64_t lo = 0, hi = n_docs;
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if (byte_off[mid] <= q) lo = mid + 1; else hi = mid;
            }
            const uint64_t d = lo ? lo - 1 : 0;
            atomicMin(&docerr[d].utf_byte, (unsigned long long)q);
        }
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

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

size_t decode_workspace_bytes(uint64_t n_ids, uint64_t n_docs, uint64_t out_cap, DecodeLayout* L) {
    DecodeLayout l{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    l.n_tiles = ceil_div(n_ids + 1, DC_TILE);
    l.mask_words_tok = l.n_tiles * (DC_TILE / 32) + 8;
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
    decode_gather_kernel<<<(unsigned)L.n_tiles, DC_T, 0, st>>>(d_ids, n_ids, d_tok_off, off_base, n_docs, tds, seq_first, policy, T, d_out, out_cap,
                                                             d_byte_off, bmask, docerr, tilestate, ticket, total_out, flags);
    count_launch();
    {
        uint64_t blocks = ceil_div(ceil_div(out_cap, 32), 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        decode_validate_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, total_out, out_cap, bmask, d_byte_off, n_docs, T, docerr);
        count_launch();
    }
    if (n_docs) {
        decode_status_kernel<<<(unsigned)ceil_div(n_docs, 256), 256, 0, st>>>(docerr, n_docs, d_doc_status, first_bad);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace tkk
