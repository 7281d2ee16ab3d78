// tk_device.cuh -- device-side table lookups and the exact BPE merge loops.
//
// Replaces tiktoken-rs `CoreBPE`'s `ranks.get(piece)` + `byte_pair_merge` (the engine behind
// Tekkenizer::encode, src/tekkenizer.rs:384-386).  Definition kept bit-exact: parts start as
// single bytes; repeatedly merge the adjacent pair whose concatenation has the lowest rank,
// leftmost on ties; stop when no adjacent pair is a vocabulary entry.  Because every part is
// always a vocabulary entry, "rank of the concatenated bytes" is looked up by (left id, right
// id) in a pair table built at load time (tk_host.cpp) instead of by bytes.
#pragma once
#include <cuda_runtime.h>

#include "tk_common.h"

__device__ __forceinline__ uint64_t tk_ldg64(const uint64_t* p) { return __ldg((const unsigned long long*)p); }

// ---- pair table, one-slot layout (linear probing) -----------------------------------------------
// rank of bytes(l)+bytes(r), TK_INF if it is not a vocabulary entry
__device__ __forceinline__ uint32_t tk_pair_rank(const TkDeviceTables& T, uint32_t l, uint32_t r) {
    uint32_t i = tk_pair_hash(l, r) & T.pair_mask;
    const uint64_t key = tk_pair_key(l, r);
    for (;;) {
        uint64_t s = tk_ldg64(T.pair_slots + i);
        if (s == 0) return TK_INF;
        if (((s >> TK_ID_BITS) & ((1ull << (2 * TK_ID_BITS)) - 1ull)) == key) return (uint32_t)s & TK_ID_MASK;
        i = (i + 1) & T.pair_mask;
    }
}

// Two independent pair lookups with their first probes in flight together (a merge creates two
// new adjacent pairs; their ranks do not depend on each other).  l == TK_INF skips a lookup.
__device__ __forceinline__ void tk_pair_rank2_slots(const TkDeviceTables& T, uint32_t l0, uint32_t r0, uint32_t l1, uint32_t r1,
                                                    uint32_t* out0, uint32_t* out1) {
    const bool h0 = l0 != TK_INF && r0 != TK_INF, h1 = l1 != TK_INF && r1 != TK_INF;
    uint32_t i0 = tk_pair_hash(l0, r0) & T.pair_mask, i1 = tk_pair_hash(l1, r1) & T.pair_mask;
    const uint64_t k0 = tk_pair_key(l0, r0), k1 = tk_pair_key(l1, r1);
    uint64_t s0 = h0 ? tk_ldg64(T.pair_slots + i0) : 0ull;
    uint64_t s1 = h1 ? tk_ldg64(T.pair_slots + i1) : 0ull;
    const uint64_t km = (1ull << (2 * TK_ID_BITS)) - 1ull;
    uint32_t a = TK_INF, b = TK_INF;
    while (s0 != 0) {
        if (((s0 >> TK_ID_BITS) & km) == k0) { a = (uint32_t)s0 & TK_ID_MASK; break; }
        i0 = (i0 + 1) & T.pair_mask;
        s0 = tk_ldg64(T.pair_slots + i0);
    }
    while (s1 != 0) {
        if (((s1 >> TK_ID_BITS) & km) == k1) { b = (uint32_t)s1 & TK_ID_MASK; break; }
        i1 = (i1 + 1) & T.pair_mask;
        s1 = tk_ldg64(T.pair_slots + i1);
    }
    *out0 = a;
    *out1 = b;
}

#if TK_PAIR_BUCKETED
// ---- pair table, bucket layout: one bucket = four slots = one 32-byte sector, read with two 16-byte loads ----
struct TkPairBucket {
    uint4 a, b;
    __device__ __forceinline__ uint64_t slot(int k) const {
        return k == 0 ? ((uint64_t)a.y << 32 | a.x) : k == 1 ? ((uint64_t)a.w << 32 | a.z) : k == 2 ? ((uint64_t)b.y << 32 | b.x) : ((uint64_t)b.w << 32 | b.z);
    }
};
__device__ __forceinline__ TkPairBucket tk_pair_bucket_load(const TkDeviceTables& T, uint32_t bucket) {
    const uint4* p = reinterpret_cast<const uint4*>(T.pair_buckets) + 2u * bucket;
    TkPairBucket B;
    B.a = __ldg(p);
    B.b = __ldg(p + 1);
    return B;
}
// the rank stored for key (slot >> TK_ID_BITS == key | occupied bit) in bucket B, TK_INF if it is not there;
// *full = the bucket has no free slot (the key may then have spilled into the next bucket)
__device__ __forceinline__ uint32_t tk_pair_bucket_find(const TkPairBucket& B, uint64_t tag, bool* full) {
    uint32_t r = TK_INF;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t s = B.slot(k);
        if ((s >> TK_ID_BITS) == tag) r = (uint32_t)s & TK_ID_MASK;
    }
    *full = B.slot(3) != 0;
    return r;
}
__device__ __forceinline__ void tk_pair_rank2_buckets(const TkDeviceTables& T, uint32_t l0, uint32_t r0, uint32_t l1, uint32_t r1,
                                                      uint32_t* out0, uint32_t* out1) {
    const bool h0 = l0 != TK_INF && r0 != TK_INF, h1 = l1 != TK_INF && r1 != TK_INF;
    uint32_t b0 = tk_pair_hash(l0, r0) & T.bucket_mask, b1 = tk_pair_hash(l1, r1) & T.bucket_mask;
    const uint64_t t0 = (1ull << (63u - TK_ID_BITS)) | tk_pair_key(l0, r0), t1 = (1ull << (63u - TK_ID_BITS)) | tk_pair_key(l1, r1);
    TkPairBucket B0, B1;
    B0.a = B0.b = B1.a = B1.b = make_uint4(0u, 0u, 0u, 0u);
    if (h0) B0 = tk_pair_bucket_load(T, b0);
    if (h1) B1 = tk_pair_bucket_load(T, b1);
    bool f0, f1;
    uint32_t a = tk_pair_bucket_find(B0, t0, &f0), b = tk_pair_bucket_find(B1, t1, &f1);
    while (a == TK_INF && f0) {                    // rare: the first bucket is full
        b0 = (b0 + 1) & T.bucket_mask;
        a = tk_pair_bucket_find(tk_pair_bucket_load(T, b0), t0, &f0);
    }
    while (b == TK_INF && f1) {
        b1 = (b1 + 1) & T.bucket_mask;
        b = tk_pair_bucket_find(tk_pair_bucket_load(T, b1), t1, &f1);
    }
    *out0 = a;
    *out1 = b;
}
#endif

template <bool BUCKETS>
__device__ __forceinline__ void tk_pair_rank2(const TkDeviceTables& T, uint32_t l0, uint32_t r0, uint32_t l1, uint32_t r1,
                                              uint32_t* out0, uint32_t* out1) {
#if TK_PAIR_BUCKETED
    if (BUCKETS) { tk_pair_rank2_buckets(T, l0, r0, l1, r1, out0, out1); return; }
#endif
    tk_pair_rank2_slots(T, l0, r0, l1, r1, out0, out1);
}

// ---- decoupled look-back over tiles (single-pass prefix sum) -------------------------------------
// One 64-bit word per tile: flag << 62 | value; flag 0 = not ready, 1 = value is the tile's own
// total, 2 = value is the inclusive prefix.  Called by all 32 lanes of ONE warp of the block;
// publishes this tile's total, inspects 32 predecessors per step, publishes the inclusive prefix
// and returns the exclusive prefix (same value on every lane).
__device__ __forceinline__ unsigned long long tk_ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tk_st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long tk_lookback(unsigned long long* __restrict__ state, uint32_t tile,
                                                          unsigned long long total) {
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long vmask = (1ull << 62) - 1ull;
    if (tile == 0) {
        if (lane == 0) tk_st_relaxed_u64(state, (2ull << 62) | total);
        return 0ull;
    }
    if (lane == 0) tk_st_relaxed_u64(state + tile, (1ull << 62) | total);
    unsigned long long excl = 0;
    long long base = (long long)tile - 1;
    for (;;) {
        const long long j = base - (long long)lane;
        const unsigned long long v = j >= 0 ? tk_ld_relaxed_u64(state + j) : (2ull << 62);
        const uint32_t f = (uint32_t)(v >> 62);
        const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, f == 2u), m0 = __ballot_sync(0xFFFFFFFFu, f == 0u);
        const uint32_t upto = m2 ? (0xFFFFFFFFu >> (32 - __ffs((int)m2))) : 0xFFFFFFFFu;   // lanes up to the first prefix
        if (m0 & upto) {
            __nanosleep(40);
            continue;
        }
        unsigned long long c = ((upto >> lane) & 1u) ? (v & vmask) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        excl += c;
        if (m2) break;
        base -= 32;
    }
    if (lane == 0) tk_st_relaxed_u64(state + tile, (2ull << 62) | (excl + total));
    return excl;
}

// The two halves of tk_lookback, for a kernel that has other work to do between knowing its tile's total and needing
// the prefix: ONE thread publishes, later ONE WARP collects.
__device__ __forceinline__ void tk_lookback_publish(unsigned long long* __restrict__ state, uint32_t tile, unsigned long long total) {
    tk_st_relaxed_u64(state + tile, ((tile == 0 ? 2ull : 1ull) << 62) | total);
}
__device__ __forceinline__ unsigned long long tk_lookback_collect(unsigned long long* __restrict__ state, uint32_t tile,
                                                                  unsigned long long total) {
    if (tile == 0) return 0ull;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long vmask = (1ull << 62) - 1ull;
    unsigned long long excl = 0;
    long long base = (long long)tile - 1;
    for (;;) {
        const long long j = base - (long long)lane;
        const unsigned long long v = j >= 0 ? tk_ld_relaxed_u64(state + j) : (2ull << 62);
        const uint32_t f = (uint32_t)(v >> 62);
        const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, f == 2u), m0 = __ballot_sync(0xFFFFFFFFu, f == 0u);
        const uint32_t upto = m2 ? (0xFFFFFFFFu >> (32 - __ffs((int)m2))) : 0xFFFFFFFFu;
        if (m0 & upto) {
            __nanosleep(40);
            continue;
        }
        unsigned long long c = ((upto >> lane) & 1u) ? (v & vmask) : 0ull;
#pragma unroll
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        excl += c;
        if (m2) break;
        base -= 32;
    }
    if (lane == 0) tk_st_relaxed_u64(state + tile, (2ull << 62) | (excl + total));
    return excl;
}

// Whole-piece lookup (CoreBPE's `encoder.get(piece)` shortcut).  p may point to shared or global
// memory; len >= 1.
__device__ __forceinline__ uint32_t tk_vocab_lookup(const TkDeviceTables& T, const uint8_t* p, uint32_t len) {
    if (len > T.max_token_len) return TK_INF;
    TkPieceHasher h;
    h.init(len);
    uint64_t key8 = 0;
    for (uint32_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        uint32_t m = len - i < 8 ? len - i : 8;
        for (uint32_t j = 0; j < m; ++j) w |= (uint64_t)p[i + j] << (8 * j);
        if (i == 0) key8 = w;
        h.add(w);
    }
    const uint64_t hv = h.finish();
    const uint64_t key = len <= 8 ? key8 : hv;
    uint32_t i = (uint32_t)hv & T.vocab_mask;
    for (;;) {
        const uint4 raw = __ldg((const uint4*)(T.vocab_slots + i));
        const uint32_t slen = raw.w;
        if (slen == 0) return TK_INF;
        const uint64_t skey = (uint64_t)raw.y << 32 | raw.x;
        if (slen == len && skey == key) {
            if (len <= 8) return raw.z;
            const uint8_t* v = T.vocab_bytes + T.vocab_off[raw.z];
            bool same = true;
            for (uint32_t j = 0; j < len; ++j)
                if (__ldg(v + j) != p[j]) { same = false; break; }
            if (same) return raw.z;
        }
        i = (i + 1) & T.vocab_mask;
    }
}

// The same lookup for a piece held in a 4-byte aligned shared-memory array: bytes [s, s+len) of
// `base`.  Reads aligned 32-bit words and funnel-shifts them into place (up to 11 bytes past the
// piece are read; the caller's array has that slack).
__device__ __forceinline__ uint32_t tk_vocab_lookup_w32(const TkDeviceTables& T, const uint8_t* base, uint32_t s, uint32_t len) {
    if (len > T.max_token_len) return TK_INF;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(base) + (s >> 2);
    const uint32_t sh = (s & 3u) * 8u;
    TkPieceHasher h;
    h.init(len);
    uint64_t key8 = 0, key16 = 0;                          // bytes 0..7 and 8..15 of the piece, zero padded
    uint32_t w0 = wp[0];
    for (uint32_t i = 0, k = 0; i < len; i += 8, k += 2) {
        const uint32_t w1 = wp[k + 1], w2 = wp[k + 2];
        uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        w0 = w2;
        const uint32_t rem = len - i;
        if (rem < 8u) {
            if (rem <= 4u) { hi = 0u; if (rem < 4u) lo &= (1u << (8u * rem)) - 1u; }
            else hi &= (1u << (8u * (rem - 4u))) - 1u;
        }
        const uint64_t w = (uint64_t)hi << 32 | lo;
        if (i == 0) key8 = w;
        if (i == 8) key16 = w;
        h.add(w);
    }
    const uint64_t hv = h.finish();
    const uint64_t key = len <= 8 ? key8 : hv;
    uint32_t i = (uint32_t)hv & T.vocab_mask;
    for (;;) {
        const uint4 raw = __ldg((const uint4*)(T.vocab_slots + i));
        const uint32_t slen = raw.w;
        if (slen == 0) return TK_INF;
        const uint64_t skey = (uint64_t)raw.y << 32 | raw.x;
        if (slen == len && skey == key) {
            if (len <= 8) return raw.z;
            if (len <= 16) {
                // the token's bytes, zero padded to 16, in one load
                const uint4 c = __ldg(T.vocab_pad16 + raw.z);
                if (((uint64_t)c.y << 32 | c.x) == key8 && ((uint64_t)c.w << 32 | c.z) == key16) return raw.z;
            } else {
                const uint8_t* v = T.vocab_bytes + T.vocab_off[raw.z];
                const uint8_t* q = base + s;
                bool same = true;
                for (uint32_t j = 0; j < len; ++j)
                    if (__ldg(v + j) != q[j]) { same = false; break; }
                if (same) return raw.z;
            }
        }
        i = (i + 1) & T.vocab_mask;
    }
}

// ---- one lane, one piece of at most TK_LANE_MAX bytes ---------------------------------------------
// Pieces by length (bytes):  <= TK_LANE_MAX  one lane per piece from a global queue of its length
// class (lanemerge_kernel);  <= TK_MED_MAX  one warp per
// piece;  beyond: one block per piece.  All run the same sequential definition.
#define TK_LANE_MAX 96
#define TK_LANE_DEAD 0xFFFFFFFEu

// Exact byte_pair_merge of one piece by one lane.  id[j] = id of the part that starts at byte
// offset j (initially the byte itself), key[j] = rank << 7 | j of the pair (part at j, next live
// part), TK_INF if that pair is not a vocabulary entry or j is the last part.  Parts never move: a
// merge writes the new id at the left part's offset, clears the right part's bit in the live mask
// (the set of live offsets, 32, 64 or 128 bits in registers), so neighbours come from bit
// operations.  The minimum key is the lowest rank, leftmost on ties.  Returns the live mask.
// bit helpers for the live mask: 32-bit when the piece class fits, else 64-bit
__device__ __forceinline__ uint32_t tk_ffs_m(uint32_t m) { return (uint32_t)__ffs((int)m); }
__device__ __forceinline__ uint32_t tk_ffs_m(unsigned long long m) { return (uint32_t)__ffsll((long long)m); }
__device__ __forceinline__ uint32_t tk_top_m(uint32_t m) { return 31u - (uint32_t)__clz((int)m); }
__device__ __forceinline__ uint32_t tk_top_m(unsigned long long m) { return 63u - (uint32_t)__clzll((long long)m); }
__device__ __forceinline__ uint32_t tk_popc_m(uint32_t m) { return (uint32_t)__popc(m); }
__device__ __forceinline__ uint32_t tk_popc_m(unsigned long long m) { return (uint32_t)__popcll(m); }
typedef unsigned __int128 tk_u128;
__device__ __forceinline__ uint32_t tk_ffs_m(tk_u128 m) {
    const unsigned long long lo = (unsigned long long)m, hi = (unsigned long long)(m >> 64);
    return lo ? (uint32_t)__ffsll((long long)lo) : (hi ? 64u + (uint32_t)__ffsll((long long)hi) : 0u);
}
__device__ __forceinline__ uint32_t tk_top_m(tk_u128 m) {
    const unsigned long long lo = (unsigned long long)m, hi = (unsigned long long)(m >> 64);
    return hi ? 127u - (uint32_t)__clzll((long long)hi) : 63u - (uint32_t)__clzll((long long)lo);
}
__device__ __forceinline__ uint32_t tk_popc_m(tk_u128 m) {
    return (uint32_t)__popcll((unsigned long long)m) + (uint32_t)__popcll((unsigned long long)(m >> 64));
}
#define TK_KEY_SHIFT 7u                                  // key = rank << TK_KEY_SHIFT | offset; offsets < 128

#ifndef TK_BUCKET_ABOVE
#define TK_BUCKET_ABOVE 16      // lane-merge classes of pieces longer than this use the bucket layout of the pair table
#endif
template <class M, int MAXLEN>
__device__ __forceinline__ M tk_bpe_merge_loop(const TkDeviceTables& T, uint32_t len, uint32_t* id, uint32_t* key, uint32_t& lookups) {
    constexpr uint32_t kBits = sizeof(M) * 8;
    static_assert(MAXLEN % 4 == 0 && MAXLEN <= (int)kBits, "scan is unrolled by four; one live bit per offset");
    const M one = 1;
    M live = len >= kBits ? ~(M)0 : (M)((one << len) - one);
    for (;;) {
        // all MAXLEN slots are scanned (the caller set the ones past the piece to TK_INF): a fixed trip
        // count keeps the loop free of remainder branches, and a class holds lengths near MAXLEN anyway
        uint32_t best = TK_INF;
#pragma unroll
        for (int j = 0; j < MAXLEN; j += 4) {
            const uint32_t a0 = key[j], a1 = key[j + 1], a2 = key[j + 2], a3 = key[j + 3];
            best = min(min(best, a0), min(a1, min(a2, a3)));
        }
        if (best == TK_INF) break;
        const uint32_t bp = best & ((1u << TK_KEY_SHIFT) - 1u), rank = best >> TK_KEY_SHIFT;
        const M above = live & ~(M)(((one << bp) << 1) - one);     // live offsets > bp (bp is never the top bit: it has a right neighbour)
        const uint32_t q = tk_ffs_m(above) - 1u;                   // exists: the pair has a rank
        const M above_q = above & (above - one);                    // live offsets > q
        const M below = live & (M)((one << bp) - one);              // live offsets < bp
        live &= ~(M)(one << q);
        const uint32_t nn = above_q ? tk_ffs_m(above_q) - 1u : 0xFFFFFFFFu;
        const uint32_t pv = below ? tk_top_m(below) : 0xFFFFFFFFu;
        id[bp] = rank;
        key[q] = TK_INF;
        const uint32_t lft = pv != 0xFFFFFFFFu ? id[pv] : TK_INF;
        const uint32_t rgt = nn != 0xFFFFFFFFu ? id[nn] : TK_INF;
        uint32_t r0, r1;
        tk_pair_rank2<(MAXLEN > TK_BUCKET_ABOVE)>(T, lft, rank, rank, rgt, &r0, &r1);
        lookups += (lft != TK_INF ? 1u : 0u) + (rgt != TK_INF ? 1u : 0u);
        if (pv != 0xFFFFFFFFu) key[pv] = r0 == TK_INF ? TK_INF : ((r0 << TK_KEY_SHIFT) | pv);
        key[bp] = r1 == TK_INF ? TK_INF : ((r1 << TK_KEY_SHIFT) | bp);
    }
    return live;
}

// ---- pieces of 2..4 bytes that are not vocabulary entries: the merge loop written out -------------------
// With at most four parts the loop has at most two steps, and its LAST possible pair is always the whole piece,
// which is known not to be a token (the caller's whole-piece lookup missed): 0, 2 or 3 byte-pair lookups and at most
// two pair-table lookups, no scratch, no loop.  The lookup kernel runs this inline for such pieces instead of queueing
// them (a third of all queued pieces on mixed text; every 3-digit group of a long number).  out: up to 4 ranks.
__device__ __forceinline__ uint32_t tk_bpe_tiny(const TkDeviceTables& T, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t len,
                                                uint32_t* out) {
    if (len == 2) { out[0] = b0; out[1] = b1; return 2; }                 // its only pair is the whole piece
    const uint32_t r0 = __ldg(T.byte_pair + ((b0 << 8) | b1)), r1 = __ldg(T.byte_pair + ((b1 << 8) | b2));
    if (len == 3) {
        if (r0 == TK_INF && r1 == TK_INF) { out[0] = b0; out[1] = b1; out[2] = b2; return 3; }
        if (r0 <= r1) { out[0] = r0; out[1] = b2; }                       // leftmost on ties
        else { out[0] = b0; out[1] = r1; }
        return 2;
    }
    const uint32_t r2 = __ldg(T.byte_pair + ((b2 << 8) | b3));
    const uint32_t m = min(r0, min(r1, r2));
    if (m == TK_INF) { out[0] = b0; out[1] = b1; out[2] = b2; out[3] = b3; return 4; }
    uint32_t p0, p1;
    if (r0 == m) {                                                        // (a b) c d
        p0 = tk_pair_rank(T, r0, b2); p1 = r2;
        if (p0 == TK_INF && p1 == TK_INF) { out[0] = r0; out[1] = b2; out[2] = b3; return 3; }
        if (p0 <= p1) { out[0] = p0; out[1] = b3; } else { out[0] = r0; out[1] = p1; }
    } else if (r1 == m) {                                                 // a (b c) d
        tk_pair_rank2<false>(T, b0, r1, r1, b3, &p0, &p1);
        if (p0 == TK_INF && p1 == TK_INF) { out[0] = b0; out[1] = r1; out[2] = b3; return 3; }
        if (p0 <= p1) { out[0] = p0; out[1] = b3; } else { out[0] = b0; out[1] = p1; }
    } else {                                                              // a b (c d)
        p0 = r0; p1 = tk_pair_rank(T, b1, r2);
        if (p0 == TK_INF && p1 == TK_INF) { out[0] = b0; out[1] = b1; out[2] = r2; return 3; }
        if (p0 <= p1) { out[0] = p0; out[1] = r2; } else { out[0] = b0; out[1] = p1; }
    }
    return 2;
}

// ---- one warp, one medium piece ------------------------------------------------------------------
// Same loop, parts in shared memory (TK_MED_MAX entries per warp), each lane caching the minimum
// of its own slice so a merge step costs one warp-wide min + a rescan by the lanes it touched.
#define TK_MED_MAX 512
#define TK_DEAD 0xFFFFFFFEu

struct TkWarpBpeSmem {
    uint32_t id[TK_MED_MAX];
    uint32_t rk[TK_MED_MAX];
    uint16_t nx[TK_MED_MAX];
    uint16_t pv[TK_MED_MAX];
};

// key = rank << 9 | position : the warp-wide minimum is the lowest rank, leftmost on ties
__device__ __forceinline__ uint32_t tk_slice_min(const uint32_t* rk, uint32_t lo, uint32_t hi) {
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t r = rk[i];
        if (r != TK_INF) { uint32_t k = (r << 9) | i; best = k < best ? k : best; }
    }
    return best;
}

// src: the piece bytes (global or shared).  out: buffer for up to n ranks (global or shared).  Returns the count
// (same value on all lanes).  Must be called by all 32 lanes.
__device__ inline uint32_t tk_bpe_warp(const TkDeviceTables& T, TkWarpBpeSmem& S, const uint8_t* src, uint32_t n,
                                       uint32_t* out) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = lane; i < n; i += 32) {
        S.id[i] = src[i];                // plain load: src is global text in the batch path, shared memory in the single-block kernel
        S.nx[i] = (uint16_t)(i + 1);
        S.pv[i] = (uint16_t)(i - 1);   // 0xFFFF for i == 0
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) S.rk[i] = (i + 1 < n) ? tk_pair_rank(T, S.id[i], S.id[i + 1]) : TK_INF;
    __syncwarp();
    const uint32_t k = (n + 31u) / 32u;
    const uint32_t lo = lane * k < n ? lane * k : n, hi = (lane + 1) * k < n ? (lane + 1) * k : n;
    uint32_t mine = tk_slice_min(S.rk, lo, hi);
    for (;;) {
        const uint32_t g = __reduce_min_sync(0xFFFFFFFFu, mine);
        if (g == 0xFFFFFFFFu) break;
        const uint32_t pos = g & 511u, r = g >> 9;
        const uint32_t j = S.nx[pos], p = S.pv[pos];
        const uint32_t nn = S.nx[j];
        __syncwarp();
        if (lane == 0) {
            S.id[pos] = r;
            S.id[j] = TK_DEAD;
            S.rk[j] = TK_INF;
            S.nx[pos] = (uint16_t)nn;
            if (nn < n) S.pv[nn] = (uint16_t)pos;
            S.rk[pos] = nn < n ? tk_pair_rank(T, r, S.id[nn]) : TK_INF;
        } else if (lane == 1) {
            if (p != 0xFFFFu) S.rk[p] = tk_pair_rank(T, S.id[p], r);
        }
        __syncwarp();
        const bool touched = (pos >= lo && pos < hi) || (j >= lo && j < hi) || (p != 0xFFFFu && p >= lo && p < hi);
        if (touched) mine = tk_slice_min(S.rk, lo, hi);
    }
    // compact the surviving parts
    uint32_t count = 0;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < n ? S.id[i] : TK_DEAD;
        const uint32_t alive = __ballot_sync(0xFFFFFFFFu, v != TK_DEAD);
        if (v != TK_DEAD) out[count + __popc(alive & ((1u << lane) - 1u))] = v;
        count += __popc(alive);
    }
    __syncwarp();
    return count;
}
