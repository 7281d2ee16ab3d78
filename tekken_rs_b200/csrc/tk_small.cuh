// tk_small.cuh -- the latency path: ONE kernel, ONE block for a single short text (included by tk_kernels.cu).
//
// Tekkenizer::encode (src/tekkenizer.rs:378-405) is called with one string at a time; the reference answers a
// short prompt in microseconds.  The batch pipeline (two dozen launches over global workspaces) cannot: its floor
// is launch overhead.  For a text of at most FS_MAX_BYTES this kernel runs every stage inside one thread block with
// all intermediates in shared memory -- stage the text (read straight from mapped pinned host memory), classify +
// split (the same pretok_tile the batch path runs), list the pieces, whole-piece lookup (one lane per piece), exact
// byte_pair_merge of the misses (one warp per piece, tk_bpe_warp), compaction (+num_special, BOS/EOS) -- and
// writes the ids and a completion word to mapped pinned host memory, which the host polls: no copies, no stream
// synchronisation, no memsets.  Pieces longer than TK_MED_MAX bytes make it report FS_NEED_BATCH and the host
// sends the text through the batch path instead (they need the block-level kernel).
#pragma once

#define FS_T PT_T                          // 256 threads = 256 windows of 32 bytes = one pre-tokeniser tile
#define FS_MAX_BYTES (PT_T * 32 - 64)      // 8128: the end sentinel and a padded last window stay inside the tile
#define FS_WARPS (FS_T / 32)
#define FS_NEED_BATCH 1u                   // result flag: a piece needs the batch path's long-piece kernels
#define FS_BAD_UTF8 2u

struct FsResult {                          // header of the mapped output buffer (ids follow at word 8)
    uint32_t n_ids, flags, err_pos, pad;
    uint32_t done;                         // == the call's sequence number when everything above and the ids are visible
    uint32_t pad2[3];
};

struct FsSmem {
    PtSmem pt;                             // text with halo + class masks (pretok_tile)
    uint32_t ds[FS_T + 4];                 // document-start bits: bit 0 and the end sentinel
    uint32_t start[FS_T + 4];              // piece-start bits
    TkkTileSummary summ;
    unsigned long long err;
    alignas(16) uint32_t stream[FS_T * 32 + 32];   // rank stream, one word per byte position (filled by 16-byte stores)
    uint16_t list[FS_T * 32 + 8];          // piece starts, in order, the sentinel last
    uint32_t missq[FS_T * 16 + 8];         // pieces to merge: start | len << 13   (later: staging of the compacted ids)
    uint32_t wsum[FS_WARPS];
    uint32_t n_pieces, n_miss, next_miss, flags;
    TkWarpBpeSmem bpe[FS_WARPS];
};

__device__ __forceinline__ uint32_t fs_block_excl(uint32_t v, uint32_t* wsum, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < FS_WARPS; ++w) { if (w < (int)warp) before += wsum[w]; all += wsum[w]; }
    *total = all;
    return before + inc - v;
}

__global__ void __launch_bounds__(FS_T, 1) encode_small_kernel(const uint8_t* __restrict__ text, uint32_t n, TkDeviceTables T,
                                                               uint32_t add_bos, uint32_t add_eos, uint32_t* __restrict__ out,
                                                               uint32_t seq) {
    extern __shared__ __align__(16) unsigned char fs_raw[];
    FsSmem& S = *reinterpret_cast<FsSmem*>(fs_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint64_t n_windows = n / 32 + 1;
    // ---- document-start bits (one document), cleared outputs ----
    {
        uint32_t d = t == 0 ? 1u : 0u;
        if ((n >> 5) == t) d |= 1u << (n & 31);
        S.ds[t] = d;
        S.start[t] = 0;
        if (t < 4) { S.ds[FS_T + t] = 0; S.start[FS_T + t] = 0; }
        if (t == 0) { S.err = ~0ull; S.n_miss = 0; S.next_miss = 0; S.flags = 0; }
        uint4* s4 = reinterpret_cast<uint4*>(S.stream);
        const uint4 inv = make_uint4(EN_INVALID, EN_INVALID, EN_INVALID, EN_INVALID);
        for (uint32_t i = t; i < (FS_T * 32 + 32) / 4; i += FS_T) s4[i] = inv;
    }
    __syncthreads();
    // ---- split: the batch path's tile routine on tile 0 (entry state of a text start; the summary says whether
    // a whitespace candidate is still pending at the end of the tile: past the end of the text the run has ended,
    // so it is a piece start) ----
    pretok_tile<false>(S.pt, 0u, text, (uint64_t)n, S.ds, S.start, n_windows, T, &S.summ, nullptr, &S.err);
    __syncthreads();
    if (t == 0 && S.summ.pend_pos >= 0) S.start[S.summ.pend_pos >> 5] |= 1u << (S.summ.pend_pos & 31);
    __syncthreads();
    if (S.err != ~0ull) {
        if (t == 0) {
            FsResult* r = reinterpret_cast<FsResult*>(out);
            r->n_ids = 0; r->flags = FS_BAD_UTF8; r->err_pos = (uint32_t)S.err;
            __threadfence_system();
            *(volatile uint32_t*)&r->done = seq;
        }
        return;
    }
    // ---- piece list (the sentinel at n is its last entry) ----
    uint32_t np;
    {
        uint32_t m = S.start[t];
        uint32_t o = fs_block_excl((uint32_t)__popc(m), S.wsum, &np);
        while (m) {
            S.list[o++] = (uint16_t)(t * 32u + (uint32_t)(__ffs((int)m) - 1));
            m &= m - 1;
        }
    }
    __syncthreads();
    np -= 1;                                            // pieces = starts minus the sentinel
    const uint8_t* bytes = S.pt.bytes + PT_HALO;
    // ---- whole-piece lookup, one lane per piece ----
    for (uint32_t k = t; k < np; k += FS_T) {
        const uint32_t s = S.list[k], len = (uint32_t)S.list[k + 1] - s;
        uint32_t r = tk_vocab_lookup_w32(T, bytes, s, len);
        if (r == TK_INF && len == 1) r = bytes[s];
        if (r != TK_INF) S.stream[s] = r;
        else if (len <= TK_MED_MAX) S.missq[atomicAdd(&S.n_miss, 1u)] = s | (len << 13);
        else S.flags = FS_NEED_BATCH;
    }
    __syncthreads();
    if (S.flags & FS_NEED_BATCH) {
        if (t == 0) {
            FsResult* r = reinterpret_cast<FsResult*>(out);
            r->n_ids = 0; r->flags = FS_NEED_BATCH; r->err_pos = 0;
            __threadfence_system();
            *(volatile uint32_t*)&r->done = seq;
        }
        return;
    }
    // ---- exact byte_pair_merge of the misses, one warp per piece; ranks go to the piece's own stream positions ----
    {
        const uint32_t nm = S.n_miss;
        for (;;) {
            uint32_t i = 0;
            if (lane == 0) i = atomicAdd(&S.next_miss, 1u);
            i = __shfl_sync(0xFFFFFFFFu, i, 0);
            if (i >= nm) break;
            const uint32_t e = S.missq[i];
            tk_bpe_warp(T, S.bpe[warp], bytes + (e & 8191u), e >> 13, S.stream + (e & 8191u));
        }
    }
    __syncthreads();
    // ---- compaction: ids (+num_special), BOS first, EOS last; staged in shared memory, then coalesced stores ----
    uint32_t* stage = S.missq;                          // FS_T * 16 words are not enough for 8 Ki ids: spill over into `list` is not
                                                        // possible (uint16), so the ids go out in two rounds of half a tile when needed
    {
        uint32_t w[32];
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) { w[j] = S.stream[t * 32 + j]; cnt += w[j] < EN_LONGREF ? 1u : 0u; }
        uint32_t total;
        const uint32_t before = fs_block_excl(cnt, S.wsum, &total);
        const uint32_t n_ids = total + (add_bos ? 1u : 0u) + (add_eos ? 1u : 0u);
        uint32_t* ids = out + 8;
        if (t == 0 && add_bos) ids[0] = T.bos_id;
        if (t == 0 && add_eos) ids[n_ids - 1] = T.eos_id;
        const uint32_t base = add_bos ? 1u : 0u;
        // rounds over the compacted sequence, FS_T * 16 ids at a time through the staging array
        for (uint32_t r0 = 0; r0 < total; r0 += FS_T * 16) {
            __syncthreads();
            uint32_t o = before;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (w[j] < EN_LONGREF) {
                    if (o >= r0 && o < r0 + FS_T * 16) stage[o - r0] = w[j] + T.num_special;
                    ++o;
                }
            }
            __syncthreads();
            const uint32_t m = total - r0 < FS_T * 16 ? total - r0 : FS_T * 16;
            for (uint32_t i = t; i < m; i += FS_T) ids[base + r0 + i] = stage[i];
        }
        __threadfence_system();
        __syncthreads();
        if (t == 0) {
            FsResult* r = reinterpret_cast<FsResult*>(out);
            r->n_ids = n_ids; r->flags = 0; r->err_pos = 0;
            __threadfence_system();
            *(volatile uint32_t*)&r->done = seq;
        }
    }
}

static cudaError_t encode_small_launch(const TkDeviceTables& T, const uint8_t* d_text, uint32_t n, int add_bos, int add_eos,
                                       uint32_t* d_out, uint32_t seq, cudaStream_t st) {
    static std::atomic<uint64_t> attr_set{0};   // bit per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!((attr_set.load() >> (dev & 63)) & 1ull)) {
        e = cudaFuncSetAttribute(encode_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FsSmem));
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(1ull << (dev & 63));
    }
    encode_small_kernel<<<1, FS_T, sizeof(FsSmem), st>>>(d_text, n, T, add_bos ? 1u : 0u, add_eos ? 1u : 0u, d_out, seq);
    count_launch();
    return cudaGetLastError();
}
