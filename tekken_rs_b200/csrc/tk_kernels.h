// tk_kernels.h -- host-visible interface of the kernel launch sequences (tk_kernels.cu, tk_decode.cu).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include "tk_common.h"

namespace tkk {

// flags word
enum : uint32_t {
    TKK_FLAG_BAD_OFFSETS = 1u,    // doc_off / tok_off not monotone or not covering the data
    TKK_FLAG_SCRATCH_FULL = 2u,   // huge-piece scratch exhausted (host grows it and retries)
    TKK_FLAG_OUT_FULL = 4u,       // caller's output buffer too small
};

// layout of the 256-byte "small" block (uint32 indices)
enum {
    TKK_S_FLAGS = 0,
    TKK_S_WORKCOUNT = 1,
    TKK_S_NLONG = 2,
    TKK_S_NHUGE = 3,
    TKK_S_WC_LONG = 4,
    TKK_S_WC_HUGE = 5,
    TKK_S_TICKET = 6,    // (decode)
    TKK_S_NMID = 6,      // (encode) block-level pieces of at most HG_SPLIT bytes
    TKK_S_WC_MID = 7,
    TKK_S_ERRPOS = 8,    // u64
    TKK_S_POOLCUR = 10,  // u64
    TKK_S_SCRCUR = 12,   // u64
    TKK_S_TOTAL = 14,    // u64
    TKK_S_BADDOC = 16,   // u64 (decode)
    TKK_S_QN = 20,       // TKK_N_CLASSES counters: queued pieces per length class
    TKK_S_QW = 30,       // TKK_N_CLASSES work counters
    TKK_S_PAIRLOOK = 40, // u64: pair-table lookups issued by the lane-merge kernels (two per merge, minus piece edges)
    TKK_S_BPLOOK = 42,   // u64: byte-pair table lookups (first round of every queued piece)
    TKK_S_ROUNDS1 = 44,  // block-level long-piece kernel: single-rank rounds
    TKK_S_ROUNDSM = 45,  // ... multi-rank rounds
    TKK_S_ROUNDSCUT = 46, // ... multi-rank rounds that were cut (undecided chain or hazard)
    TKK_S_APPLIEDM = 47, // ... merges applied by multi-rank rounds
};

#define TKK_N_CLASSES 9
struct TkkQueueLayout {
    uint64_t off[TKK_N_CLASSES];   // first entry of every class in the queue array
};

struct TkkTileSummary {
    uint32_t packed;     // composite TkRunSummary of the tile
    uint32_t assumed;    // entry state used in the first pass: n | abs<<2 | n_prov<<3 | r_prov<<4
    long long pend_pos;  // byte position of an unresolved whitespace candidate, or -1
};

struct TkkLongRec {
    uint64_t start;     // byte position of the piece
    uint64_t len;       // bytes
    uint64_t tok_base;  // index into the token pool
    uint32_t count;     // ranks produced
    uint32_t pad;
};

struct EncodeLayout {
    uint64_t n_windows, n_tiles, n_ltiles, mask_words, max_long, queue_words;
    TkkQueueLayout queues;
    size_t off_small, off_ds, off_start, off_longword, off_docfirst, off_doccnt, off_summ, off_carry, off_worklist, off_seg, off_tilecount, off_tilebase, off_bsum, off_recs,
        off_huge, off_pool, off_stream, off_queues, total;
};

struct DecodeLayout {
    uint64_t n_tiles, mask_words_tok, mask_words_out;
    size_t off_small, off_tds, off_seqfirst, off_bmask, off_tilestate, off_docerr, total;
};

// optional per-stage CUDA-event timing
struct StageTimer {
    std::vector<std::string> names;
    std::vector<cudaEvent_t> events;
    void mark(cudaStream_t st, const char* name);
    void collect(std::vector<std::string>& out_names, std::vector<float>& out_ms);
    void reset();
    ~StageTimer();
};

// the allocation that holds the merge kernels' tables (pair table, then byte-pair table)
struct HotTables {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool enabled = false;     // the device accepted a persisting-L2 carve-out that covers it
};

// class tables of the split for the pattern stored in tekken.json (tk_pretok_cfg.h); null = the reference's pattern
struct CfgSplitTables {
    const uint16_t* stage1 = nullptr;
    const uint8_t* stage2 = nullptr;
};

size_t encode_workspace_bytes(uint64_t n, uint64_t n_docs, EncodeLayout* L);
cudaError_t encode_device(const TkDeviceTables& T, const uint8_t* d_data, const uint64_t* d_doc_off, uint64_t off_base,
                          uint64_t n_docs, uint64_t n, int add_bos, int add_eos, uint32_t* d_out, uint64_t out_cap, uint64_t* d_tok_off,
                          void* d_ws, const EncodeLayout& L, uint32_t* d_scratch, uint64_t scratch_cap, int sm_count,
                          cudaStream_t st, StageTimer* timer, const HotTables* hot = nullptr, const CfgSplitTables* cfg = nullptr);

// The latency path (tk_small.cuh): one single-block kernel encodes one text of at most kSmallMaxBytes.  d_text (16-byte
// aligned) and d_out may be mapped pinned host memory.  d_out: 8 header words {n_ids, flags, err_pos, -, done, ...},
// ids from word 8; `done` becomes `seq` when everything is visible to the host.
constexpr uint32_t kSmallMaxBytes = 256 * 32 - 64;
constexpr uint32_t kSmallOutWords = 8 + kSmallMaxBytes + 2 + 6;
constexpr uint32_t kSmallNeedBatch = 1u, kSmallBadUtf8 = 2u;
// ... and one id list of at most kSmallDecodeIds ids that decodes to at most kSmallDecodeBytes bytes (tk_decode.cu).  d_out:
// 8 header words {n_bytes, status, flags, -, done = seq, ...}, then the text; both buffers are mapped pinned memory.
constexpr uint32_t kSmallDecodeIds = 2048, kSmallDecodeBytes = 24576;
cudaError_t decode_small(const TkDeviceTables& T, const uint32_t* d_ids, uint32_t n, int policy, uint32_t* d_out, uint32_t seq, cudaStream_t st);
cudaError_t encode_small(const TkDeviceTables& T, const uint8_t* d_text, uint32_t n, int add_bos, int add_eos, uint32_t* d_out,
                         uint32_t seq, cudaStream_t st);

cudaError_t publish_counters(const void* d_ws, const EncodeLayout& L, uint32_t* mapped_dev, cudaStream_t st);
cudaError_t publish_small(const void* d_small, uint32_t* mapped_dev, cudaStream_t st);     // any 256-byte counter block

size_t decode_workspace_bytes(uint64_t n_ids, uint64_t n_docs, uint64_t out_cap, DecodeLayout* L);
cudaError_t decode_device(const TkDeviceTables& T, const uint32_t* d_ids, const uint64_t* d_tok_off, uint64_t off_base, uint64_t n_docs,
                          uint64_t n_ids, int policy, uint8_t* d_out, uint64_t out_cap, uint64_t* d_byte_off,
                          int32_t* d_doc_status, void* d_ws, const DecodeLayout& L, cudaStream_t st);

long long debug_bounds_violations(unsigned long long* out4);   // -1: not a -DTK_DEBUG_BOUNDS build
uint64_t launch_count();
void count_launch();
long long decode_debug_bounds_violations(unsigned long long* out4);     // tk_decode.cu's share of tk_debug_bounds_violations

// ids -> a little-endian bit stream of `bits` (18 or 24) bits per id, 16 ids per group of bits / 2 words; the host-buffer
// engine sends this over PCIe instead of the 32-bit ids and widens it on the host (tk_api.cu).  d_out holds
// packed_id_bytes(n, bits) bytes.
inline size_t packed_id_bytes(uint64_t n, int bits) { return (size_t)((n + 15) / 16) * (size_t)bits * 2; }
cudaError_t pack_ids(const uint32_t* d_ids, uint64_t n, int bits, void* d_out, cudaStream_t st);

}  // namespace tkk
