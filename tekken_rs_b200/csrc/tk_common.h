// tk_common.h -- types and hash functions shared by the host table builders and the kernels.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <vector_types.h>
#else
struct uint4 { uint32_t x, y, z, w; };
#endif

#if defined(__CUDACC__)
#define TK_HD __host__ __device__ __forceinline__
#else
#define TK_HD inline
#endif

#define TK_INF 0xFFFFFFFFu

// ---- character classes of the split pattern (src/tekkenizer.rs:123) ----------------------
// L = \p{L}, N = \p{N}, R = {CR, LF}, W = \s minus R, O = everything else.
enum { TK_CL_O = 0, TK_CL_L = 1, TK_CL_N = 2, TK_CL_W = 3, TK_CL_R = 4 };

// Two-stage Unicode class table: stage1[cp >> 7] -> block index; stage2 holds 128 two-bit
// entries per block (32 bytes), values TK_CL_O/L/N/W.  R (CR/LF) is ASCII and handled inline.
// A block of one class is TK_UNI_UNIFORM | class in stage 1 and has no stage-2 entry.
#define TK_UNI_STAGE1_N (0x110000 >> 7)
#define TK_UNI_UNIFORM 0x8000u

TK_HD uint32_t tk_class_lookup(const uint16_t* stage1, const uint8_t* stage2, uint32_t cp) {
    uint32_t blk = stage1[cp >> 7];
    if (blk & TK_UNI_UNIFORM) return blk & 3u;
    uint32_t b = stage2[blk * 32u + ((cp & 127u) >> 2)];
    return (b >> ((cp & 3u) * 2u)) & 3u;
}

// ---- vocabulary hash table: byte string -> rank ----------------------------------------------
// Slot is 16 bytes.  len == 0 marks an empty slot.  For len <= 8 `key` is the bytes themselves
// (little-endian packed, zero padded) so a key+len match is exact; for len > 8 `key` is the
// 64-bit hash and the caller verifies the bytes against the vocabulary byte table.
struct TkVocabSlot {
    uint64_t key;
    uint32_t rank;
    uint32_t len;
};

TK_HD uint64_t tk_mix64(uint64_t x) {
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    return x;
}

// Hash of a piece given as little-endian 8-byte words (last one zero padded).  32-bit arithmetic
// only (two multiply lanes + a murmur-style finaliser): the GPU has no 64-bit multiplier, and the
// whole-piece lookup runs once per pre-token.  finish(): low 32 bits index the table, all 64 bits
// are the stored key of entries longer than 8 bytes.
struct TkPieceHasher {
    uint32_t a, b;
    TK_HD void init(uint32_t len) { a = 0x9E3779B9u * (len + 1u); b = 0x85EBCA6Bu ^ (len * 0x27D4EB2Fu); }
    TK_HD void add(uint64_t w) {
        a = (a ^ (uint32_t)w) * 0xCC9E2D51u;
        a = (a << 15) | (a >> 17);
        b = (b ^ (uint32_t)(w >> 32)) * 0x1B873593u;
        b = ((b << 13) | (b >> 19)) + a;
    }
    TK_HD uint64_t finish() const {
        uint32_t h = a ^ ((b << 16) | (b >> 16));
        h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
        uint32_t g = (b ^ (a >> 7)) * 0x9E3779B1u;
        g ^= g >> 15;
        return ((uint64_t)g << 32) | h;
    }
};

// ---- pair table: (left id, right id) -> rank of the concatenation ------------------------------
// One u64 per slot: bit 63 = occupied, bits 42..62 = left id, 21..41 = right id, 0..20 = rank.
// Ids are < 2^21 (checked at load).
// Layout (TK_PAIR_BUCKETED, the default): BUCKETS of four slots = one 32-byte sector, fetched with two
// 16-byte loads; a key lives in the bucket its hash names, slots of a bucket fill front to back, and only the
// entries that find their bucket full spill into the next one -- so a lookup (hit or miss) ends in its first
// bucket unless that bucket is full (well under 1 % of the buckets at the load factor the builder keeps).
// The merge kernels run one lane per piece: a warp-wide round of lookups takes as many dependent round trips
// to L2 as its unluckiest lane, and with one slot per probe (the round-1 layout, kept under TK_PAIR_BUCKETED=0)
// that was two to four per round; with buckets it is one.
// Both layouts are resident (TK_PAIR_BUCKETED=1, the default; 8 MB + 16 MB for the Tekken vocabulary): measured on
// the bench corpus, the classes of 17..96-byte pieces -- low occupancy, bound by the latency of the dependent
// round trips -- run 1.1-1.35x faster on buckets, while the classes up to 16 bytes -- bound by the L1TEX pipe,
// where a fully divergent load costs one cycle per lane and two 16-byte loads cost twice one 8-byte load -- are
// 1.1-1.3x faster on the one-slot table (r02d A/B in profiles/).  Kernels pick by a template flag.
#ifndef TK_PAIR_BUCKETED
#define TK_PAIR_BUCKETED 1
#endif
#define TK_PAIR_BUCKET_SLOTS 4u
#define TK_ID_BITS 21u
#define TK_ID_MASK ((1u << TK_ID_BITS) - 1u)

TK_HD uint64_t tk_pair_key(uint32_t l, uint32_t r) { return ((uint64_t)l << TK_ID_BITS) | (uint64_t)r; }
TK_HD uint64_t tk_pair_slot(uint32_t l, uint32_t r, uint32_t rank) {
    return (1ull << 63) | (tk_pair_key(l, r) << TK_ID_BITS) | (uint64_t)rank;
}
TK_HD uint32_t tk_pair_hash(uint32_t l, uint32_t r) {
    // 32-bit multiplies only (the lookup runs twice per merge step)
    uint32_t h = (l * 0x9E3779B1u) ^ (r * 0x85EBCA77u);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    return h ^ (h >> 13);
}

// Device-resident tables of one tokenizer (plain pointers; filled by the loader).
struct TkDeviceTables {
    const uint16_t* uni_stage1;    // TK_UNI_STAGE1_N entries
    const uint8_t* uni_stage2;     // n_blocks * 32 bytes
    const TkVocabSlot* vocab_slots;
    uint32_t vocab_mask;           // capacity - 1
    const uint64_t* pair_slots;    // one-slot open addressing (linear probing)
    uint32_t pair_mask;            // slot count - 1
    const uint64_t* pair_buckets;  // the same entries in four-slot buckets (null with TK_PAIR_BUCKETED=0)
    uint32_t bucket_mask;          // bucket count - 1
    const uint32_t* byte_pair;     // [b0 << 8 | b1] -> rank of the two-byte token, TK_INF if none
    const uint4* vocab_pad16;      // token bytes zero-padded to 16 (tokens longer than that: first 16 bytes)
    const uint4* vocab_e16;        // token bytes 0..6 | length (0xFF = longer than 15, see vocab_off) | token bytes 7..14; zero padded
    const uint8_t* vocab_bytes;    // concatenated token bytes, rank order
    const uint32_t* vocab_off;     // n_vocab + 1
    const uint8_t* special_bytes;  // concatenated special strings, positional order
    const uint32_t* special_off;   // num_special + 1
    uint32_t n_vocab;              // inner vocabulary size
    uint32_t num_special;
    uint32_t max_token_len;
    uint32_t bos_id, eos_id;       // TK_INF when absent
};
