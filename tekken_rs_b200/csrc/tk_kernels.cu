// tk_kernels.cu -- sm_100a kernels of the encode path and their launch sequence.
//
// Encode (Tekkenizer::encode, src/tekkenizer.rs:378-405, batched) = the launch sequence in
// encode_device():
//   K0  docmark      document-start bitmask, first document / document count per 32-byte window,
//                    BOS/EOS counted per 4 KiB tile
//   K1  pretok       regex split -> piece-start bitmask (tk_pretok.h), one thread per 32 bytes
//   K1s/K1f          carry digit / CR-LF / whitespace run state across tiles (three small scan kernels
//                    + a fix-up pass that is rarely non-trivial)
//   K2a longmark     find pieces longer than TK_LANE_MAX bytes
//   K3  longmerge    exact BPE for those: warp per piece in shared memory, block per huge piece in rounds
//   K2  lookup       one lane per piece: whole-piece vocabulary lookup -> rank stream (one word per byte
//                    position); other pieces -> global queues by length class
//   K2m lanemerge    x9 (one launch per length class, 2..96 bytes): one lane per queued piece: exact byte_pair_merge
//   K3s tilesum/scan/apply   prefix of the per-tile token counts -> first output position of every tile
//   K4  emit         compaction of the rank stream: ids (+num_special), BOS/EOS, per-document offsets
// Decode lives in tk_decode.cu.
#include "tk_kernels.h"

#include <array>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "tk_device.cuh"
#include "tk_pretok.h"
#include "tk_pretok_cfg.h"

namespace tkk {

#define TKK_COUNT_TILE 4096u      // tokens are counted per 4 KiB of text (= lookup / emit tile)

// ---- bounds-checked debug build (-DTK_DEBUG_BOUNDS) ------------------------------------------------
// compute-sanitizer is not available on the GPU pool this library is developed on, so the kernels carry their own
// checks: in a debug build every store into a workspace / output array first compares its index with the array's
// size (the host publishes the sizes of the call in g_dbg before the launches); a violation is skipped and recorded
// (source line, index) in g_dbg_hit, which tk_debug_bounds_violations() reads.  tests/test_gpu_debug_bounds.py
// builds this variant and runs the parity corpus through it.  A regular build compiles the checks away.
struct DbgLimits {
    unsigned long long mask_words, stream_words, queue_words, pool_words, scratch_words, count_tiles, out_cap, n_docs, max_long;
};
#ifdef TK_DEBUG_BOUNDS
__device__ DbgLimits g_dbg;
__device__ unsigned long long g_dbg_hit[4];      // violations, first line, its index, its limit
__device__ __forceinline__ bool dbg_in(unsigned long long i, unsigned long long n, int line) {
    if (i < n) return true;
    if (atomicAdd(&g_dbg_hit[0], 1ull) == 0ull) { g_dbg_hit[1] = (unsigned long long)line; g_dbg_hit[2] = i; g_dbg_hit[3] = n; }
    return false;
}
#define TK_DBG(index, field) dbg_in((unsigned long long)(index), g_dbg.field, __LINE__)
#else
#define TK_DBG(index, field) true
#endif

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
#define TK_LAUNCHED() count_launch()

void StageTimer::mark(cudaStream_t st, const char* name) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    names.emplace_back(name);
    events.push_back(e);
}
void StageTimer::collect(std::vector<std::string>& out_names, std::vector<float>& out_ms) {
    out_names.clear();
    out_ms.clear();
    for (size_t i = 0; i + 1 < events.size(); ++i) {
        float ms = 0.f;
        cudaEventSynchronize(events[i + 1]);
        cudaEventElapsedTime(&ms, events[i], events[i + 1]);
        out_names.push_back(names[i]);
        out_ms.push_back(ms);
    }
}
void StageTimer::reset() {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    events.clear();
    names.clear();
}
StageTimer::~StageTimer() { reset(); }

// ---- programmatic dependent launch ---------------------------------------------------------------------
// The encode path is a chain of two dozen dependent launches on one stream.  They are launched with the
// programmatic-stream-serialization attribute: a kernel's blocks may be set up on the SMs while the kernel before it
// drains, and every kernel begins with pdl_wait() (griddepcontrol.wait), which returns when the preceding kernel's
// memory is visible -- so nothing is read early, and the launch latency of each link overlaps the tail of the one
// before.  TEKKEN_B200_PDL=0 launches the plain way (A/B measurements).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("TEKKEN_B200_PDL"); return !e || atoi(e) != 0; }();
    return on;
}
template <class... KArgs, class... Args>
static cudaError_t launch_chain(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- bulk copy of a tile's text into shared memory (TMA, 1-D: cp.async.bulk + mbarrier) ---------------------
// One thread issues the copy of the whole tile; the block's threads do other work (mask words, piece lists) and meet
// at the mbarrier before they read the bytes.  Source and destination 16-byte aligned, size a multiple of 16.
__device__ __forceinline__ void bulk_tile_begin(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* mbar) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar), dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_tile_wait(unsigned long long* mbar) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar) : "memory");
    }
}

// K0a: everything the later kernels expect cleared, in one launch (it was seven memsets): the counter block, the
// document masks / tables, the padding of the piece-start mask, the per-tile token counts.
struct SetupArgs {
    uint32_t* small;                 // 64 words: 0, except err_pos = ~0
    uint4* zero0; uint64_t n0;       // ds mask           (16-byte units)
    uint4* ones;  uint64_t n1;       // first-document table: 0xFF
    uint4* zero1; uint64_t n2;       // document counts
    uint32_t* tail; uint64_t n3;     // piece-start mask padding (words)
    unsigned long long* counts; uint64_t n4;
};
__global__ void __launch_bounds__(256) setup_kernel(SetupArgs a) {
    pdl_wait();
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (uint64_t)gridDim.x * blockDim.x;
    if (i0 < 64) a.small[i0] = (i0 == TKK_S_ERRPOS || i0 == TKK_S_ERRPOS + 1) ? 0xFFFFFFFFu : 0u;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u), f = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (uint64_t i = i0; i < a.n0; i += step) a.zero0[i] = z;
    for (uint64_t i = i0; i < a.n1; i += step) a.ones[i] = f;
    for (uint64_t i = i0; i < a.n2; i += step) a.zero1[i] = z;
    for (uint64_t i = i0; i < a.n3; i += step) a.tail[i] = 0u;
    for (uint64_t i = i0; i < a.n4; i += step) a.counts[i] = 0ull;
}

// =====================================================================================================
// K0: document-start bitmask.  Bit p is set iff some document starts at byte p; bit `total` is
// the end-of-data sentinel.  Also validates the offsets.
// =====================================================================================================
__global__ void docmark_kernel(const uint64_t* __restrict__ doc_off, uint64_t off_base, uint64_t n_docs, uint64_t total,
                               uint32_t add_bos, uint32_t add_eos, uint32_t* __restrict__ ds_mask, uint32_t* __restrict__ doc_first,
                               uint32_t* __restrict__ doc_cnt, unsigned long long* __restrict__ tile_count,
                               uint32_t* __restrict__ flags) {
    pdl_wait();
    uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_docs) return;
    // the end-of-data sentinel does not depend on the offsets being valid: every later kernel that walks the
    // piece-start mask to "the next set bit" relies on it, also when the call is about to fail with bad offsets
    if (d == n_docs && TK_DBG(total >> 5, mask_words)) atomicOr(ds_mask + (total >> 5), 1u << (total & 31));
    uint64_t o = doc_off[d] - off_base;      // offsets may be a slice of a larger batch (off_base = its first entry)
    bool ok = doc_off[d] >= off_base && o <= total;
    if (d == 0 && o != 0) ok = false;
    if (d == n_docs && o != total) ok = false;
    if (d < n_docs && doc_off[d + 1] < doc_off[d]) ok = false;
    if (!ok) { atomicOr(flags, TKK_FLAG_BAD_OFFSETS); return; }
    if (!TK_DBG(o >> 5, mask_words)) return;
    atomicOr(ds_mask + (o >> 5), 1u << (o & 31));
    atomicMin(doc_first + (o >> 5), (uint32_t)d);     // first document that starts in this 32-byte window
    atomicAdd(doc_cnt + (o >> 5), 1u);                // ... and how many do
    // the EOS of the document before and the BOS of this one count towards the tile this document starts in
    const unsigned long long sp = (d > 0 ? add_eos : 0u) + (d < n_docs ? add_bos : 0u);
    if (sp && TK_DBG(o / TKK_COUNT_TILE, count_tiles)) atomicAdd(tile_count + o / TKK_COUNT_TILE, sp);
}

// =====================================================================================================
// K1: pre-tokeniser.  256 threads = 256 windows of 32 bytes = one 8 KiB tile.
// =====================================================================================================
#define PT_T 256

__device__ __forceinline__ uint32_t rs_pack(const TkRunSummary& s) {
    return s.n_all | (s.n_val << 1) | (s.r_mode << 3) | (s.head << 5);
}
__device__ __forceinline__ TkRunSummary rs_unpack(uint32_t p) {
    TkRunSummary s;
    s.n_all = p & 1u; s.n_val = (p >> 1) & 3u; s.r_mode = (p >> 3) & 3u; s.head = (p >> 5) & 3u;
    return s;
}
#define RS_IDENTITY (1u | (2u << 3))

#define PT_HALO 64
struct PtSmem {
    uint8_t bytes[PT_HALO + PT_T * 32 + PT_HALO];   // the tile's text with a halo; zero outside the text
    uint32_t lead[PT_T + 2], mL[PT_T + 2], mN[PT_T + 2], mR[PT_T + 2], mW[PT_T + 2], sp[PT_T + 2], ap[PT_T + 2],
        ds[PT_T + 2];
    uint32_t head[PT_T + 1];
    uint32_t wtot[PT_T / 32];
    long long pend;
    unsigned long long mbar;           // mbarrier of the tile's bulk copy
};

__device__ __forceinline__ void pt_store(PtSmem& S, int i, const TkWin& w) {
    S.lead[i] = w.lead; S.mL[i] = w.mL; S.mN[i] = w.mN; S.mR[i] = w.mR; S.mW[i] = w.mW; S.sp[i] = w.sp; S.ap[i] = w.ap;
    S.ds[i] = w.ds;
}
__device__ __forceinline__ TkWin pt_load(const PtSmem& S, int i) {
    TkWin w;
    w.lead = S.lead[i]; w.mL = S.mL[i]; w.mN = S.mN[i]; w.mR = S.mR[i]; w.mW = S.mW[i]; w.sp = S.sp[i]; w.ap = S.ap[i];
    w.ds = S.ds[i]; w.bad = 0;
    return w;
}

// stage [tile - halo, tile + 8 KiB + halo) of the text in shared memory, zero outside the text (PT_T threads).
// With `mbar`, a tile that lies inside the text goes by ONE bulk copy issued by thread 0 (returns true: the caller
// waits on the mbarrier after its next __syncthreads); tiles at the edges of the text take the loop.
__device__ __forceinline__ bool pt_stage_tile(uint8_t* bytes, uint32_t b, const uint8_t* __restrict__ data, uint64_t n,
                                              unsigned long long* mbar = nullptr) {
    const int t = threadIdx.x;
    const long long g0 = (long long)b * (PT_T * 32) - PT_HALO;
    constexpr uint32_t kBytes = PT_HALO + PT_T * 32 + PT_HALO;
    if (mbar && g0 >= 0 && (uint64_t)g0 + kBytes <= n && (((uintptr_t)data) & 15u) == 0) {
        if (t == 0) bulk_tile_begin(bytes, data + g0, kBytes, mbar);
        return true;
    }
    uint4* dst = reinterpret_cast<uint4*>(bytes);
    for (int i = t; i < (int)((PT_HALO + PT_T * 32 + PT_HALO) / 16); i += PT_T) {
        const long long g = g0 + 16ll * i;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (g >= 0 && (uint64_t)g + 16u <= n) v = __ldg(reinterpret_cast<const uint4*>(data + g));
        else if (g + 16 > 0 && (uint64_t)(g < 0 ? 0 : g) < n) {
            uint32_t x[4] = {0u, 0u, 0u, 0u};
            for (int k = 0; k < 16; ++k) {
                const long long q = g + k;
                if (q >= 0 && (uint64_t)q < n) x[k >> 2] |= (uint32_t)data[q] << (8 * (k & 3));
            }
            v = make_uint4(x[0], x[1], x[2], x[3]);
        }
        dst[i] = v;
    }
    return false;
}

// classify window wi of the text from the staged tile (wl = its index inside the tile, -1 .. PT_T)
__device__ __forceinline__ TkWin pt_classify(const PtSmem& S, const TkBytesTile& src, const uint32_t* __restrict__ ds_mask,
                                             uint64_t n_windows, long long wi, int wl, const TkDeviceTables& T) {
    TkWin z;
    z.lead = 0xFFFFFFFFu; z.mL = z.mN = z.mR = z.mW = z.sp = z.ap = z.ds = z.bad = 0;
    if (wi < 0 || (uint64_t)wi >= n_windows) return z;
    const uint4* q = reinterpret_cast<const uint4*>(S.bytes + PT_HALO + wl * 32);
    const uint4 a = q[0], b = q[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    return tk_classify_window(src, (uint64_t)wi * 32u, w, ds_mask[wi], T);
}

// Process tile b.  FIX=false: first pass (entry state guessed from the halo window, summary
// written).  FIX=true: re-run with the exact entry state / pending verdict from the scan kernel.
template <bool FIX>
__device__ __forceinline__ void pretok_tile(PtSmem& S, uint32_t b, const uint8_t* __restrict__ data, uint64_t n,
                                            const uint32_t* __restrict__ ds_mask, uint32_t* __restrict__ start_mask,
                                            uint64_t n_windows, const TkDeviceTables& T, TkkTileSummary* __restrict__ summ,
                                            const uint32_t* __restrict__ carry, unsigned long long* __restrict__ err_pos) {
    const int t = threadIdx.x;
    const long long wi = (long long)b * PT_T + t;
    const uint64_t pos = (uint64_t)wi * 32u;
    const bool bulk = pt_stage_tile(S.bytes, b, data, n, FIX ? nullptr : &S.mbar);    // (the fix pass re-uses its block for many tiles)
    __syncthreads();
    if (bulk) bulk_tile_wait(&S.mbar);
    const TkBytesTile src{S.bytes + PT_HALO, (int64_t)b * (PT_T * 32)};
    TkWin c = pt_classify(S, src, ds_mask, n_windows, wi, t, T);
    pt_store(S, t + 1, c);
    if (t < 2) {
        const long long hw = t == 0 ? (long long)b * PT_T - 1 : (long long)b * PT_T + PT_T;
        TkWin h = pt_classify(S, src, ds_mask, n_windows, hw, t == 0 ? -1 : PT_T, T);
        pt_store(S, t == 0 ? 0 : PT_T + 1, h);
    }
    if (t == 0) S.pend = -1;
    __syncthreads();
    const TkWin p = pt_load(S, t), nx = pt_load(S, t + 2);
    TkWin zero;
    zero.lead = 0xFFFFFFFFu; zero.mL = zero.mN = zero.mR = zero.mW = zero.sp = zero.ap = zero.ds = zero.bad = 0;
    const TkDerived dp = tk_derive(src, pos - 32, zero, p, c, 0);
    const TkDerived dc = tk_derive(src, pos, p, c, nx, dp.sO);
    const TkRunSummary mine = tk_summarize(c);
    S.head[t] = mine.head;
    if (t == 0) S.head[PT_T] = tk_summarize(pt_load(S, PT_T + 1)).head;

    // exclusive scan of the run summaries over the tile
    const int lane = t & 31, warp = t >> 5;
    uint32_t inc = rs_pack(mine);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc = rs_pack(tk_compose(rs_unpack(o), rs_unpack(inc)));
    }
    if (lane == 31) S.wtot[warp] = inc;
    uint32_t exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = RS_IDENTITY;
    __syncthreads();
    uint32_t pre = RS_IDENTITY;
    for (int w = 0; w < warp; ++w) pre = rs_pack(tk_compose(rs_unpack(pre), rs_unpack(S.wtot[w])));
    const TkRunSummary E = tk_compose(rs_unpack(pre), rs_unpack(exc));

    // state entering the tile
    uint32_t n0, a0, n_prov = 0, r_prov = 0, confirm = 0;
    if (FIX) {
        const uint32_t cw = carry[b];
        n0 = cw & 3u; a0 = (cw >> 2) & 1u; confirm = (cw >> 3) & 1u;
    } else if (b == 0) {
        n0 = 0; a0 = 0;
    } else {
        const TkRunSummary hs = tk_summarize(pt_load(S, 0));
        n_prov = hs.n_all; n0 = hs.n_val;
        r_prov = hs.r_mode == 2u; a0 = r_prov ? 0u : hs.r_mode;
    }
    const uint32_t n_in = E.n_all ? (n0 + E.n_val) % 3u : E.n_val;
    const uint32_t abs_in = E.r_mode == 2u ? a0 : E.r_mode;

    TkEval ev = tk_eval_window(p, c, nx, dp, dc, n_in, abs_in);
    uint32_t start = ev.start;
    if (ev.pend >= 0) {
        uint32_t verdict = 0;
        for (int v = t + 2; v <= PT_T; ++v) {
            verdict = S.head[v];
            if (verdict) break;
        }
        if (verdict == 2u) start |= 1u << ev.pend;
        else if (verdict == 0u) {
            if (FIX) { if (confirm) start |= 1u << ev.pend; }
            else S.pend = (long long)(pos + (uint64_t)ev.pend);
        }
    }
    if ((uint64_t)wi < n_windows) {
        const uint32_t keep = (pos + 32 <= n) ? 0xFFFFFFFFu : (uint32_t)((2ull << (n - pos)) - 1ull);
        if (TK_DBG(wi, mask_words)) start_mask[wi] = start & keep;
        const uint32_t valid = (pos + 32 <= n) ? 0xFFFFFFFFu : (uint32_t)((1ull << (n - pos)) - 1ull);
        if (c.bad & valid) atomicMin(err_pos, (unsigned long long)(pos + (uint64_t)(__ffs((int)(c.bad & valid)) - 1)));
    }
    if (!FIX) {
        __syncthreads();
        if (t == PT_T - 1) {
            TkkTileSummary s;
            s.packed = rs_pack(tk_compose(E, mine));
            s.assumed = n0 | (a0 << 2) | (n_prov << 3) | (r_prov << 4);
            s.pend_pos = S.pend;
            summ[b] = s;
        }
    }
}

#ifndef PT_MINB
#define PT_MINB 6      // resident blocks per SM the compiler plans registers for (measured: 6 beats 4, 5 and 8)
#endif
__global__ void __launch_bounds__(PT_T, PT_MINB) pretok_kernel(const uint8_t* __restrict__ data, uint64_t n,
                                                      const uint32_t* __restrict__ ds_mask, uint32_t* __restrict__ start_mask,
                                                      uint64_t n_windows, TkDeviceTables T, TkkTileSummary* __restrict__ summ,
                                                      unsigned long long* __restrict__ err_pos) {
    pdl_wait();
    __shared__ __align__(16) PtSmem S;
    pretok_tile<false>(S, blockIdx.x, data, n, ds_mask, start_mask, n_windows, T, summ, nullptr, err_pos);
}

__global__ void __launch_bounds__(PT_T) pretok_fix_kernel(const uint8_t* __restrict__ data, uint64_t n,
                                                          const uint32_t* __restrict__ ds_mask, uint32_t* __restrict__ start_mask,
                                                          uint64_t n_windows, TkDeviceTables T, const uint32_t* __restrict__ carry,
                                                          const uint32_t* __restrict__ worklist, const uint32_t* __restrict__ work_count,
                                                          unsigned long long* __restrict__ err_pos) {
    pdl_wait();
    __shared__ __align__(16) PtSmem S;
    const uint32_t cnt = *work_count;
    for (uint32_t w = blockIdx.x; w < cnt; w += gridDim.x) {
        pretok_tile<true>(S, worklist[w], data, n, ds_mask, start_mask, n_windows, T, nullptr, carry, err_pos);
        __syncthreads();
    }
}

// K1s: scan of the tile summaries in three small kernels: (1) composite of every segment of
// SG_TILES tiles, (2) one block scans the segment composites (state entering every segment, first
// whitespace event after it), (3) every segment again: exact entry state of every tile, verdict
// for every pending whitespace candidate, and the list of tiles whose guess was wrong.
#define SG_T 256
#define SG_PER 8
#define SG_TILES (SG_T * SG_PER)
#define SC_T 1024

__device__ __forceinline__ void rs_step(uint32_t& nst, uint32_t& ast, const TkRunSummary& g) {
    nst = g.n_all ? (nst + g.n_val) % 3u : g.n_val;
    ast = g.r_mode == 2u ? ast : g.r_mode;
}

__global__ void __launch_bounds__(SG_T) pretok_seg_kernel(const TkkTileSummary* __restrict__ summ, uint32_t n_tiles,
                                                          uint32_t* __restrict__ seg_packed) {
    pdl_wait();
    __shared__ uint32_t f_chunk[SG_T];
    const uint32_t t = threadIdx.x;
    const uint64_t lo0 = (uint64_t)blockIdx.x * SG_TILES + (uint64_t)t * SG_PER;
    const uint32_t lo = lo0 < n_tiles ? (uint32_t)lo0 : n_tiles, hi = lo + SG_PER < n_tiles ? lo + SG_PER : n_tiles;
    TkRunSummary f = rs_unpack(RS_IDENTITY);
    for (uint32_t b = lo; b < hi; ++b) f = tk_compose(f, rs_unpack(summ[b].packed));
    f_chunk[t] = rs_pack(f);
    __syncthreads();
    // ordered reduction (the composition is not commutative): warp 0, 8 chunks per lane, then a shuffle chain
    if (t < 32) {
        TkRunSummary g = rs_unpack(RS_IDENTITY);
        for (int j = 0; j < SG_T / 32; ++j) g = tk_compose(g, rs_unpack(f_chunk[t * (SG_T / 32) + j]));
        uint32_t v = rs_pack(g);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if ((int)t >= d) v = rs_pack(tk_compose(rs_unpack(o), rs_unpack(v)));
        }
        if (t == 31) seg_packed[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(SC_T) pretok_segscan_kernel(const uint32_t* __restrict__ seg_packed, uint32_t n_seg,
                                                              uint32_t* __restrict__ seg_in, uint32_t* __restrict__ seg_after) {
    pdl_wait();
    __shared__ uint32_t f_chunk[SC_T], in_state[SC_T], h_chunk[SC_T], h_after[SC_T];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (n_seg + SC_T - 1) / SC_T;
    const uint32_t lo = t * per < n_seg ? t * per : n_seg, hi = (t + 1) * per < n_seg ? (t + 1) * per : n_seg;
    TkRunSummary f = rs_unpack(RS_IDENTITY);
    for (uint32_t b = lo; b < hi; ++b) f = tk_compose(f, rs_unpack(seg_packed[b]));
    f_chunk[t] = rs_pack(f);
    h_chunk[t] = f.head;
    __syncthreads();
    // threads past `act` hold no segment; the two serial walks (one lane each, in different warps) skip them
    const uint32_t act = per ? (n_seg + per - 1) / per : 0u;
    if (t == 0) {
        uint32_t nst = 0, ast = 0;
        for (uint32_t j = 0; j < act; ++j) {
            in_state[j] = nst | (ast << 2);
            rs_step(nst, ast, rs_unpack(f_chunk[j]));
        }
    } else if (t == 32) {
        uint32_t nh = 2;  // past the end of the data the run has ended
        for (int j = (int)act - 1; j >= 0; --j) {
            h_after[j] = nh;
            if (h_chunk[j]) nh = h_chunk[j];
        }
    }
    __syncthreads();
    if (t >= act) return;
    uint32_t nst = in_state[t] & 3u, ast = in_state[t] >> 2;
    for (uint32_t b = lo; b < hi; ++b) {
        seg_in[b] = nst | (ast << 2);
        rs_step(nst, ast, rs_unpack(seg_packed[b]));
    }
    uint32_t nh = h_after[t];
    for (uint32_t b = hi; b-- > lo;) {
        seg_after[b] = nh;
        const uint32_t h = rs_unpack(seg_packed[b]).head;
        if (h) nh = h;
    }
}

__global__ void __launch_bounds__(SG_T) pretok_apply_kernel(const TkkTileSummary* __restrict__ summ, uint32_t n_tiles,
                                                            const uint32_t* __restrict__ seg_in, const uint32_t* __restrict__ seg_after,
                                                            uint32_t* __restrict__ carry, uint32_t* __restrict__ worklist,
                                                            uint32_t* __restrict__ work_count, uint32_t* __restrict__ start_mask) {
    pdl_wait();
    __shared__ uint32_t f_chunk[SG_T];     // composite of each thread's tiles
    __shared__ uint32_t in_state[SG_T];    // n | abs<<2 entering each thread's tiles
    __shared__ uint32_t h_after[SG_T];     // first head event after each thread's tiles
    const uint32_t t = threadIdx.x;
    const uint64_t lo0 = (uint64_t)blockIdx.x * SG_TILES + (uint64_t)t * SG_PER;
    const uint32_t lo = lo0 < n_tiles ? (uint32_t)lo0 : n_tiles, hi = lo + SG_PER < n_tiles ? lo + SG_PER : n_tiles;
    TkkTileSummary mine[SG_PER];
    TkRunSummary f = rs_unpack(RS_IDENTITY);
#pragma unroll
    for (int k = 0; k < SG_PER; ++k)
        if (lo + k < hi) { mine[k] = summ[lo + k]; f = tk_compose(f, rs_unpack(mine[k].packed)); }
    f_chunk[t] = rs_pack(f);
    __syncthreads();
    // threads of the last block past `act` hold no tile
    const uint64_t left = (uint64_t)n_tiles - (uint64_t)blockIdx.x * SG_TILES;
    const uint32_t act = left >= SG_TILES ? SG_T : (uint32_t)((left + SG_PER - 1) / SG_PER);
    if (t == 0) {
        uint32_t nst = seg_in[blockIdx.x] & 3u, ast = seg_in[blockIdx.x] >> 2;
        for (uint32_t j = 0; j < act; ++j) {
            in_state[j] = nst | (ast << 2);
            rs_step(nst, ast, rs_unpack(f_chunk[j]));
        }
    } else if (t == 32) {
        uint32_t nh = seg_after[blockIdx.x];
        for (int j = (int)act - 1; j >= 0; --j) {
            h_after[j] = nh;
            const uint32_t h = rs_unpack(f_chunk[j]).head;
            if (h) nh = h;
        }
    }
    __syncthreads();
    // forward: entry states; backward: verdicts
    uint32_t cw[SG_PER];
    uint32_t nst = in_state[t] & 3u, ast = in_state[t] >> 2;
#pragma unroll
    for (int k = 0; k < SG_PER; ++k)
        if (lo + k < hi) { cw[k] = nst | (ast << 2); rs_step(nst, ast, rs_unpack(mine[k].packed)); }
    uint32_t nh = h_after[t];
#pragma unroll
    for (int k = SG_PER - 1; k >= 0; --k) {
        if (lo + k < hi) {
            const TkkTileSummary s = mine[k];
            const uint32_t confirm = nh == 2u;
            const uint32_t n0 = s.assumed & 3u, a0 = (s.assumed >> 2) & 1u, n_prov = (s.assumed >> 3) & 1u, r_prov = (s.assumed >> 4) & 1u;
            const bool wrong = (n_prov && (cw[k] & 3u) != n0) || (r_prov && ((cw[k] >> 2) & 1u) != a0);
            carry[lo + k] = cw[k] | (confirm << 3);
            if (wrong) worklist[atomicAdd(work_count, 1u)] = lo + k;
            else if (s.pend_pos >= 0 && confirm) atomicOr(start_mask + (s.pend_pos >> 5), 1u << (s.pend_pos & 31));
            const uint32_t h = rs_unpack(s.packed).head;
            if (h) nh = h;
        }
    }
}

// =====================================================================================================
// K1c: the split for the pattern STORED in tekken.json (tk_pretok_cfg.h), used by handles
// created with TK_SPLIT_CONFIG.  Two kernels replace K1/K1s/K1f; every later stage is pattern-independent.
//   cfg_mask_kernel  one thread per 32-byte window: 4-bit classes (SWAR for ASCII, two-stage table otherwise,
//                    strict UTF-8), then the SAFE starts as bit logic -- positions that four purely local rules
//                    prove to be piece starts (document start, a digit, after a digit, whitespace after
//                    non-whitespace, punctuation after a letter).  They seed the piece-start mask.
//   cfg_walk_kernel  one lane per safe start: the sequential leftmost-first matcher from piece to piece until the
//                    next safe start / document / end of text, marking the piece starts in between (atomicOr).
// A walk is as long as the distance to the next safe start: a word, a camelCase identifier, a whitespace run.
// =====================================================================================================
struct CfgSmem {
    uint8_t bytes[PT_HALO + PT_T * 32 + PT_HALO];
    TkCfgWin win[PT_T + 1];                      // [0] = the window before the tile
};

__global__ void __launch_bounds__(PT_T) cfg_mask_kernel(const uint8_t* __restrict__ data, uint64_t n, const uint32_t* __restrict__ ds_mask,
                                                        uint32_t* __restrict__ safe_mask, uint32_t* __restrict__ start_mask,
                                                        uint64_t n_windows, TkCfgTables T, unsigned long long* __restrict__ err_pos) {
    pdl_wait();
    __shared__ __align__(16) CfgSmem S;
    const int t = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const long long wi = (long long)b * PT_T + t;
    const uint64_t pos = (uint64_t)wi * 32u;
    pt_stage_tile(S.bytes, b, data, n);
    __syncthreads();
    const TkBytesTile src{S.bytes + PT_HALO, (int64_t)b * (PT_T * 32)};
    auto classify = [&](long long w, int wl) -> TkCfgWin {
        TkCfgWin z;
        z.lead = 0xFFFFFFFFu; z.mU = z.mLO = z.mC = z.mM = z.mN = z.mW = z.mR = z.ds = z.bad = 0;
        if (w < 0 || (uint64_t)w >= n_windows) return z;
        const uint4* q = reinterpret_cast<const uint4*>(S.bytes + PT_HALO + wl * 32);
        const uint4 a = q[0], c = q[1];
        const uint32_t words[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        return tk_cfg_classify_window(src, (uint64_t)w * 32u, words, ds_mask[w], T);
    };
    const TkCfgWin c = classify(wi, t);
    S.win[t + 1] = c;
    if (t == 0) S.win[0] = classify((long long)b * PT_T - 1, -1);
    __syncthreads();
    if ((uint64_t)wi < n_windows) {
        const uint32_t keep = (pos + 32 <= n) ? 0xFFFFFFFFu : (uint32_t)((2ull << (n - pos)) - 1ull);   // bit n: the end sentinel
        const uint32_t safe = tk_cfg_safe_mask(S.win[t], c) & keep;
        safe_mask[wi] = safe;
        if (TK_DBG(wi, mask_words)) start_mask[wi] = safe;
        const uint32_t valid = (pos + 32 <= n) ? 0xFFFFFFFFu : (uint32_t)((1ull << (n - pos)) - 1ull);
        if (c.bad & valid) atomicMin(err_pos, (unsigned long long)(pos + (uint64_t)(__ffs((int)(c.bad & valid)) - 1)));
    }
}

// (Measured, r02f: a variant that walks over a shared-memory copy of the tile and of the two bitmasks was 1.4x SLOWER
// than reading the text through L1 -- the walk is a chain of dependent ALU work per character, not memory-bound -- so
// the text and the masks are read from global memory.)
#define CW_T 256
__global__ void __launch_bounds__(CW_T) cfg_walk_kernel(const uint8_t* __restrict__ data, uint64_t n, const uint32_t* __restrict__ ds_mask,
                                                        const uint32_t* __restrict__ safe_mask, uint32_t* __restrict__ start_mask,
                                                        uint64_t n_windows, TkCfgTables T, const unsigned long long* __restrict__ err_pos) {
    pdl_wait();
    __shared__ uint16_t list[CW_T * 32];
    __shared__ uint32_t wsum[CW_T / 32];
    __shared__ uint32_t n_list;
    if (*err_pos != ~0ull) return;                    // invalid UTF-8: the call fails, and the matcher assumes valid text
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint64_t w = (uint64_t)blockIdx.x * CW_T + t;
    uint32_t m = w < n_windows ? safe_mask[w] : 0u;
    // the tile's safe starts, in order, one list entry each
    uint32_t inc = (uint32_t)__popc(m);
    const uint32_t mine = inc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t before = inc - mine, all = 0;
    for (uint32_t x = 0; x < CW_T / 32; ++x) { if (x < warp) before += wsum[x]; all += wsum[x]; }
    if (t == 0) n_list = all;
    while (m) {
        list[before++] = (uint16_t)(t * 32u + (uint32_t)(__ffs((int)m) - 1));
        m &= m - 1;
    }
    __syncthreads();
    const uint32_t cnt = n_list;
    const TkBytesChecked src{data, n};
    const uint64_t tile_pos = (uint64_t)blockIdx.x * CW_T * 32u;
    for (uint32_t k = t; k < cnt; k += CW_T) {
        const int64_t q0 = (int64_t)(tile_pos + list[k]);
        if ((uint64_t)q0 >= n) continue;              // the end sentinel is not a piece
        tk_cfg_walk(src, q0, safe_mask, ds_mask, (int64_t)n, T,
                    [&](int64_t p) { atomicOr(start_mask + (p >> 5), 1u << (p & 31)); });
    }
}

// =====================================================================================================
// K2a: pieces longer than TK_LANE_MAX bytes.  Such a piece must start at the highest set bit of its
// mask word, so one thread per word finds them all.
// =====================================================================================================
// one atomicAdd per warp: lanes with `pred` get consecutive slots starting at the returned base
__device__ __forceinline__ uint32_t warp_claim(uint32_t* counter, bool pred) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, pred);
    if (!m) return 0;
    const int leader = __ffs((int)m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

__global__ void longmark_kernel(const uint32_t* __restrict__ start_mask, uint64_t n_windows, uint64_t n,
                                uint32_t* __restrict__ long_of_word, TkkLongRec* __restrict__ recs,
                                uint32_t* __restrict__ n_long) {
    pdl_wait();
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // blockDim.x is a multiple of 32
    const uint32_t m = w < n_windows ? start_mask[w] : 0u;
    bool need = false;
    uint64_t pos = 0;
    if (m) {
        const uint32_t hb = 31u - (uint32_t)__clz((int)m);
        pos = w * 32u + hb;
        if (pos < n) {
            // next start within TK_LANE_MAX bytes?  (words beyond n_windows are zero-padded)
            static_assert(TK_LANE_MAX <= 97, "a piece of TK_LANE_MAX bytes ends within three mask words");
            const uint32_t m1 = start_mask[w + 1], m2 = start_mask[w + 2], m3 = start_mask[w + 3];
            uint64_t next;
            if (m1) next = (w + 1) * 32u + (uint32_t)(__ffs((int)m1) - 1);
            else if (m2) next = (w + 2) * 32u + (uint32_t)(__ffs((int)m2) - 1);
            else if (m3) next = (w + 3) * 32u + (uint32_t)(__ffs((int)m3) - 1);
            else next = ~0ull;
            need = next == ~0ull || next - pos > TK_LANE_MAX;
        }
    }
    const uint32_t slot = warp_claim(n_long, need);
    if (need) {
        TkkLongRec r;
        r.start = pos; r.len = 0; r.count = 0; r.tok_base = 0; r.pad = 0;
        if (TK_DBG(slot, max_long)) recs[slot] = r;
    }
    if (w < n_windows) long_of_word[w] = need ? slot + 1u : 0u;
}

// =====================================================================================================
// K3: long pieces.  Warp per record: measure the piece, claim output space, then either merge it
// in shared memory (<= TK_MED_MAX bytes) or queue it for the block-level kernel.
// =====================================================================================================
#define LM_WARPS 4
#define HG_SPLIT 8192u            // block-level pieces up to this many bytes get 8 warps, longer ones 32 (see longmerge_block_kernel)

__device__ __forceinline__ uint64_t piece_end(const uint32_t* __restrict__ start_mask, uint64_t pos, uint64_t n_windows, uint64_t n) {
    // position of the next set bit after pos (the sentinel at n guarantees there is one; the walk is bounded by
    // the mask's length all the same, so that a damaged mask cannot send it into the neighbouring arrays)
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t w = pos >> 5;
    uint32_t first = start_mask[w] & ~((2u << (pos & 31)) - 1u);
    if ((pos & 31) == 31) first = 0;
    if (first) return w * 32u + (uint32_t)(__ffs((int)first) - 1);
    for (w += 1;; w += 32) {
        if (w >= n_windows) return n;
        const uint32_t m = w + lane < n_windows ? start_mask[w + lane] : 0u;
        const uint32_t any = __ballot_sync(0xFFFFFFFFu, m != 0);
        if (any) {
            const int l = __ffs((int)any) - 1;
            const uint32_t mm = __shfl_sync(0xFFFFFFFFu, m, l);
            return (w + (uint64_t)l) * 32u + (uint32_t)(__ffs((int)mm) - 1);
        }
    }
}

__global__ void __launch_bounds__(LM_WARPS * 32) longmerge_warp_kernel(const uint8_t* __restrict__ data,
                                                                       const uint32_t* __restrict__ start_mask,
                                                                       TkDeviceTables T, TkkLongRec* __restrict__ recs,
                                                                       const uint32_t* __restrict__ n_long, uint32_t* __restrict__ pool,
                                                                       unsigned long long* __restrict__ pool_cursor,
                                                                       uint32_t* __restrict__ huge_list, uint32_t* __restrict__ n_huge,
                                                                       uint32_t* __restrict__ mid_list, uint32_t* __restrict__ n_mid,
                                                                       uint32_t* __restrict__ work_counter,
                                                                       unsigned long long* __restrict__ tile_count,
                                                                       uint64_t n_windows, uint64_t n, const uint32_t* __restrict__ flags) {
    pdl_wait();
    __shared__ TkWarpBpeSmem S[LM_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (*flags & TKK_FLAG_BAD_OFFSETS) return;          // the call fails; the masks describe no valid batch
    const uint32_t total = *n_long;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(work_counter, 1u);
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r >= total) break;
        const uint64_t pos = recs[r].start;
        const uint64_t end = piece_end(start_mask, pos, n_windows, n);
        const uint64_t len = end - pos;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(pool_cursor, (unsigned long long)len);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        uint32_t count = 0;
        if (len <= TK_MED_MAX) {
            // whole-piece shortcut first (CoreBPE checks the encoder map before merging)
            uint32_t whole = TK_INF;
            if (len <= T.max_token_len) {
                if (lane == 0) whole = tk_vocab_lookup(T, data + pos, (uint32_t)len);
                whole = __shfl_sync(0xFFFFFFFFu, whole, 0);
            }
            if (whole != TK_INF) {
                if (lane == 0 && TK_DBG(base, pool_words)) pool[base] = whole;
                count = 1;
            } else {
                count = tk_bpe_warp(T, S[warp], data + pos, (uint32_t)len, pool + base);
            }
        } else if (lane == 0) {
            {   // block-level pieces: two lists by size (the kernel that takes them runs with 8 or 32 warps per piece)
                const bool mid = len <= HG_SPLIT;
                const uint32_t hslot = atomicAdd(mid ? n_mid : n_huge, 1u);
                if (TK_DBG(hslot, max_long)) (mid ? mid_list : huge_list)[hslot] = r;
            }
        }
        if (lane == 0) {
            recs[r].len = len;
            recs[r].count = count;
            if (count) atomicAdd(tile_count + pos / TKK_COUNT_TILE, (unsigned long long)count);
            recs[r].tok_base = base;
        }
        __syncwarp();
    }
}

// Block per huge piece (> TK_MED_MAX bytes): the same result as the sequential merge loop, computed in
// ROUNDS ("rank-rounds with hazard cut").  A round takes the lowest pair rank m of the
// piece, selects the rank-m pairs left to right without overlaps, and computes for every selected pair
// the ranks of the two pairs its merge creates, exactly as the sequential order would see them (left
// neighbour already merged if it was selected, right neighbour not yet).  If one of those new ranks is
// <= m the sequential loop would turn to that pair next: the round applies the selected merges up to and
// including the leftmost such pair and drops the rest.  Parts and pair ranks live in compact arrays in
// global scratch (L2-resident for pieces of tens of KiB) and are rebuilt each round with a block scan.
// A run of thousands of spaces takes a handful of rounds instead of thousands of dependent steps.  Text without
// repetition (random letters, CJK: every rank occurs a handful of times) would need a round per rank; there MULTI-RANK
// rounds take over: all pairs of rank <= a threshold are candidates, the ones the sequential loop would merge are
// selected by key order ((rank, position): a candidate merges unless an overlapping neighbour with a smaller key did),
// each computes the pairs its merge creates at its time, and the round is cut at the first created pair that could
// overtake a later candidate.  64 KiB of random letters: 15-30 rounds instead of 40,000 dependent merges
// (verified against the literal loop in oracle/research/multirank_rounds.py, incl. shuffled-rank vocabularies).
// Threads per piece: a round is a handful of latency-bound sweeps over the parts, so a piece of tens of KiB gets a
// whole SM (32 warps); pieces up to HG_SPLIT bytes -- whitespace runs, long identifiers: there can be thousands in a
// batch -- get 8 warps each, eight pieces per SM.
#define HG_T_BIG 1024
#define HG_T_MID 256
#define HG_ARRAYS 7                 // id / rank, double-buffered, + the two new-rank arrays of a round + the selection state
#define HG_UNSEL 0xFFFFFFFEu        // rL marker: pair not selected this round

template <int HG_T>
__device__ __forceinline__ uint32_t hg_block_min(uint32_t v, uint32_t* s_tmp) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    v = __reduce_min_sync(0xFFFFFFFFu, v);
    __syncthreads();
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    uint32_t r = s_tmp[lane < HG_T / 32 ? lane : 0];
    r = __reduce_min_sync(0xFFFFFFFFu, r);
    return r;
}

// block-wide minimum of a 64-bit key (rank << 32 | position)
template <int HG_T>
__device__ __forceinline__ unsigned long long hg_block_min64(unsigned long long v, unsigned long long* s_tmp64) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        v = o < v ? o : v;
    }
    __syncthreads();
    if (lane == 0) s_tmp64[warp] = v;
    __syncthreads();
    unsigned long long r = s_tmp64[lane < HG_T / 32 ? lane : 0];
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, r, d);
        r = o < r ? o : r;
    }
    return r;
}
// selection state of a pair in a multi-rank round
#define HG_ST_NONE 0u
#define HG_ST_UND 1u
#define HG_ST_SEL 2u
#define HG_ST_NOT 3u
#define HG_ST_APP 4u                   // selected and applied this round (before the cut, up to the first hazard)
#define HG_PASSES 4

template <int HG_T>
__global__ void __launch_bounds__(HG_T) longmerge_block_kernel(const uint8_t* __restrict__ data, TkDeviceTables T,
                                                               TkkLongRec* __restrict__ recs, const uint32_t* __restrict__ huge_list,
                                                               const uint32_t* __restrict__ n_huge, uint32_t* __restrict__ pool,
                                                               uint32_t* __restrict__ scratch, unsigned long long scratch_cap,
                                                               unsigned long long* __restrict__ scratch_cursor,
                                                               uint32_t* __restrict__ work_counter, uint32_t* __restrict__ flags,
                                                               unsigned long long* __restrict__ tile_count) {
    pdl_wait();
    __shared__ uint32_t s_tmp[HG_T / 32];
    __shared__ uint32_t s_par[HG_T / 32];      // run-parity summaries of the warps
    __shared__ uint32_t s_cnt[HG_T / 32];
    __shared__ unsigned long long s_key[HG_T / 32];
    __shared__ uint32_t s_rec;
    __shared__ unsigned long long s_base;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t total = *n_huge;
    for (;;) {
        __syncthreads();
        if (t == 0) s_rec = atomicAdd(work_counter, 1u);
        __syncthreads();
        if (s_rec >= total) break;
        const uint32_t r = huge_list[s_rec];
        const uint64_t pos = recs[r].start;
        const uint32_t n = (uint32_t)recs[r].len;   // a call holds < 4 GiB of text (tk_api.cu encode_issue), so does a piece
        if (t == 0) s_base = atomicAdd(scratch_cursor, (unsigned long long)HG_ARRAYS * n);   // the cursor also tells the host how much is needed
        __syncthreads();
        if (s_base + (unsigned long long)HG_ARRAYS * n > scratch_cap) {
            if (t == 0) atomicOr(flags, TKK_FLAG_SCRATCH_FULL);
            continue;
        }
        uint32_t* out = pool + recs[r].tok_base;
        if (n <= T.max_token_len) {                    // whole-piece shortcut (only with giant vocab entries)
            __shared__ uint32_t s_whole;
            if (t == 0) s_whole = tk_vocab_lookup(T, data + pos, n);
            __syncthreads();
            if (s_whole != TK_INF) {
                if (t == 0) { out[0] = s_whole; recs[r].count = 1; atomicAdd(tile_count + pos / TKK_COUNT_TILE, 1ull); }
                continue;
            }
        }
        uint32_t* id = scratch + s_base;
        uint32_t* rk = id + n;
        uint32_t* id2 = rk + n;
        uint32_t* rk2 = id2 + n;
        uint32_t* rL = rk2 + n;
        uint32_t* rR = rL + n;
        uint32_t* sel = rR + n;
        for (uint32_t i = t; i < n; i += HG_T) {
            const uint32_t b0 = data[pos + i];
            id[i] = b0;
            rk[i] = i + 1 < n ? __ldg(T.byte_pair + ((b0 << 8) | data[pos + i + 1])) : TK_INF;
        }
        __syncthreads();
        uint32_t m = n;
        // Width of the window of ranks a round may merge on top of the lowest rank present.  It starts at 0 (one rank per
        // round: runs of equal pairs -- repeated characters, thousands of spaces -- halve every round), opens while rounds go
        // through with few candidates (text without repetition: every rank occurs a handful of times) and narrows again
        // after a round that had to be cut.
        uint32_t delta = 0;
        constexpr uint32_t NW = HG_T / 32;
        for (;;) {
            // Layout: every warp owns a contiguous region of the parts and walks it in rows of 32 (lane = position in the
            // row): all loads are coalesced and a pair's neighbours sit in the adjacent lanes.  (With one contiguous chunk
            // per THREAD, as in round 1, a round over 64 K parts took 0.7 ms.)
            const uint32_t wsz = ((m + NW * 32u - 1u) / (NW * 32u)) * 32u;
            const uint32_t wlo = warp * wsz < m ? warp * wsz : m, whi = wlo + wsz < m ? wlo + wsz : m;
            // 1. lowest pair rank of the piece
            uint32_t mine = TK_INF;
            for (uint32_t i = wlo + lane; i < whi; i += 128) {          // four rows per trip: four loads in flight
                const uint32_t a = rk[i], b = i + 32 < whi ? rk[i + 32] : TK_INF, c = i + 64 < whi ? rk[i + 64] : TK_INF,
                               d = i + 96 < whi ? rk[i + 96] : TK_INF;
                mine = min(min(mine, a), min(b, min(c, d)));
            }
            const uint32_t mn = hg_block_min<HG_T>(mine, s_tmp);
            if (mn == TK_INF) break;
            // 2. candidates: pairs of rank <= thr.  The sequential loop would take them in the order of their keys
            //    (rank, position).
            const uint32_t thr = mn + delta < TK_ID_MASK ? mn + delta : TK_ID_MASK;
            for (uint32_t i = wlo + lane; i < whi; i += 128) {
                const uint32_t a = rk[i], b = i + 32 < whi ? rk[i + 32] : TK_INF, c = i + 64 < whi ? rk[i + 64] : TK_INF,
                               d = i + 96 < whi ? rk[i + 96] : TK_INF;
                sel[i] = a <= thr ? HG_ST_UND : HG_ST_NONE;
                if (i + 32 < whi) sel[i + 32] = b <= thr ? HG_ST_UND : HG_ST_NONE;
                if (i + 64 < whi) sel[i + 64] = c <= thr ? HG_ST_UND : HG_ST_NONE;
                if (i + 96 < whi) sel[i + 96] = d <= thr ? HG_ST_UND : HG_ST_NONE;
            }
            __syncthreads();
            // 3. selection = what the sequential loop merges if no merge creates a pair of rank <= thr: by key order, a
            //    candidate merges unless a neighbouring candidate (they share a part) with a smaller key merged.  A local
            //    rule iterated to its fixed point inside every row (registers + shuffles), rows of a region in order, a
            //    few passes over the regions: chains of dependent candidates -- a run of equal pairs is one, left to
            //    right -- resolve across a whole region per pass.  A state only moves from undecided to decided, so
            //    reading a neighbour's state early or late changes how soon a pair is decided, never what is decided.
            if (delta == 0) {
                // One rank: the keys are the positions, so in every maximal run of adjacent candidates the 1st, 3rd, ... merge.
                // Bit arithmetic on the row's candidate mask (adding a 1 at the start of a run clears the run: runs that
                // start on an even / odd bit keep their even / odd bits), the parity of the run that enters a row carried
                // from row to row, the one that enters a region from a scan of (whole region in one run, parity of its
                // trailing run) over the warps.  A run of 65,536 equal pairs is decided in this one sweep.
                uint32_t all = 1, par = 0;
                for (uint32_t i0 = wlo; i0 < whi; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const uint32_t valid = __ballot_sync(0xFFFFFFFFu, i < whi);
                    const uint32_t B = __ballot_sync(0xFFFFFFFFu, i < whi && sel[i] != HG_ST_NONE);
                    if (B == valid) par ^= (uint32_t)__popc(B) & 1u;
                    else { all = 0; par = (uint32_t)(__clz((int)(valid & ~B)) - __clz((int)valid)) & 1u; }     // ones above the highest zero
                }
                if (lane == 0) s_par[warp] = all | (par << 1);
                __syncthreads();
                uint32_t c = 0;                                      // parity of the run that reaches my region
                for (uint32_t w = 0; w < warp; ++w) { const uint32_t x = s_par[w]; c = (x & 1u) ? (c ^ (x >> 1)) : (x >> 1); }
                for (uint32_t i0 = wlo; i0 < whi; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const bool in = i < whi;
                    const uint32_t valid = __ballot_sync(0xFFFFFFFFu, in);
                    const uint32_t B = __ballot_sync(0xFFFFFFFFu, in && sel[i] != HG_ST_NONE);
                    uint32_t even_starts = B & ~(B << 1) & 0x55555555u;
                    if (c && (B & 1u)) even_starts &= ~1u;           // the run continues from the row before with an odd length so far
                    const uint32_t r_even = B & ~(B + even_starts), r_odd = B & ~r_even;
                    const uint32_t pick = (r_even & 0x55555555u) | (r_odd & 0xAAAAAAAAu);
                    if (in && ((B >> lane) & 1u)) sel[i] = ((pick >> lane) & 1u) ? HG_ST_SEL : HG_ST_NOT;
                    // parity of the run that leaves the row
                    const uint32_t top = 31u - (uint32_t)__clz((int)valid);
                    if (!((B >> top) & 1u)) c = 0;
                    else if (B == valid) c ^= (uint32_t)__popc(B) & 1u;
                    else c = (uint32_t)(__clz((int)(valid & ~B)) - __clz((int)valid)) & 1u;
                }
                __syncthreads();
            } else
#pragma unroll 1
            for (int pass = 0; pass < HG_PASSES; ++pass) {
                // the row being settled lives in registers; the next row is loaded before this one is settled, and the two
                // neighbours outside a row come from registers as well (the settled lane 31 of the row before, the freshly
                // loaded lane 0 of the row after), so a row never waits for memory
                uint32_t cur_i = wlo + lane;
                uint32_t r = cur_i < whi ? rk[cur_i] : TK_INF, st = cur_i < whi ? sel[cur_i] : HG_ST_NONE;
                uint32_t eL_st = HG_ST_NONE, eL_r = 0;
                if (wlo > 0 && wlo < whi) { eL_st = sel[wlo - 1]; eL_r = rk[wlo - 1]; }
                for (uint32_t i0 = wlo; i0 < whi; i0 += 32) {
                    const uint32_t nxt = i0 + 32u + lane;
                    const bool take = nxt < whi || (lane == 0 && nxt < m);      // lane 0 also looks one part past the region
                    const uint32_t r_n = take ? rk[nxt] : TK_INF, st_n = take ? sel[nxt] : HG_ST_NONE;
                    const uint32_t eR_st = __shfl_sync(0xFFFFFFFFu, st_n, 0), eR_r = __shfl_sync(0xFFFFFFFFu, r_n, 0);
                    const uint32_t st_in = st;
                    for (;;) {
                        uint32_t sl = __shfl_up_sync(0xFFFFFFFFu, st, 1), rl = __shfl_up_sync(0xFFFFFFFFu, r, 1);
                        uint32_t sr = __shfl_down_sync(0xFFFFFFFFu, st, 1), rr = __shfl_down_sync(0xFFFFFFFFu, r, 1);
                        if (lane == 0) { sl = eL_st; rl = eL_r; }
                        if (lane == 31) { sr = eR_st; rr = eR_r; }
                        uint32_t nst = st;
                        if (st == HG_ST_UND) {
                            const bool lowL = sl != HG_ST_NONE && rl <= r, lowR = sr != HG_ST_NONE && rr < r;   // smaller key: left wins ties
                            if ((lowL && sl == HG_ST_SEL) || (lowR && sr == HG_ST_SEL)) nst = HG_ST_NOT;
                            else if (!((lowL && sl == HG_ST_UND) || (lowR && sr == HG_ST_UND))) nst = HG_ST_SEL;
                        }
                        const bool changed = nst != st;
                        st = nst;
                        if (!__any_sync(0xFFFFFFFFu, changed)) break;
                    }
                    if (st != st_in) sel[i0 + lane] = st;
                    eL_st = __shfl_sync(0xFFFFFFFFu, st, 31);
                    eL_r = __shfl_sync(0xFFFFFFFFu, r, 31);
                    r = nxt < whi ? r_n : TK_INF;
                    st = nxt < whi ? st_n : HG_ST_NONE;
                }
                __syncthreads();
            }
            // candidates still undecided cut the round at their key
            unsigned long long kcut = ~0ull;
            if (delta) {                                                  // (one-rank rounds decide everything)
                for (uint32_t i = wlo + lane; i < whi; i += 128) {
                    uint32_t st4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) st4[u] = i + 32u * u < whi ? sel[i + 32u * u] : HG_ST_NONE;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (st4[u] == HG_ST_UND) {
                            const unsigned long long k = (unsigned long long)rk[i + 32u * u] << 32 | (i + 32u * u);
                            kcut = k < kcut ? k : kcut;
                        }
                }
            }
            kcut = hg_block_min64<HG_T>(kcut, s_key);
            // 4. the two pairs every selected merge creates AT ITS TIME: a neighbour two positions away is already merged
            //    iff it is selected with a smaller key.  A created pair of rank <= thr is a hazard: the sequential loop might
            //    take it before a later candidate.
            unsigned long long khaz = ~0ull;
            for (uint32_t i = wlo + lane; i < whi; i += 32) {
                if (sel[i] != HG_ST_SEL) continue;
                const uint32_t r = rk[i];
                const unsigned long long k = (unsigned long long)r << 32 | i;
                if (k >= kcut) continue;
                uint32_t lf = TK_INF, rt = TK_INF;
                if (i >= 1) lf = (i >= 2 && sel[i - 2] == HG_ST_SEL && rk[i - 2] <= r) ? rk[i - 2] : id[i - 1];
                if (i + 2 < m) rt = (sel[i + 2] == HG_ST_SEL && rk[i + 2] < r) ? rk[i + 2] : id[i + 2];
                uint32_t x, y;
                tk_pair_rank2<true>(T, lf, r, r, rt, &x, &y);
                rL[i] = x;
                rR[i] = y;
                if (x <= thr || y <= thr) khaz = k < khaz ? k : khaz;
            }
            khaz = hg_block_min64<HG_T>(khaz, s_key);
            // 5. mark the merges that are applied: selected, before the cut, up to and including the first hazard
            for (uint32_t i = wlo + lane; i < whi; i += 128) {
                uint32_t st4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) st4[u] = i + 32u * u < whi ? sel[i + 32u * u] : HG_ST_NONE;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (st4[u] == HG_ST_SEL) {
                        const unsigned long long k = (unsigned long long)rk[i + 32u * u] << 32 | (i + 32u * u);
                        if (k < kcut && k <= khaz) sel[i + 32u * u] = HG_ST_APP;
                    }
            }
            __syncthreads();
            // 6. rebuild the compact arrays: the parts that survive (everything but the right part of an applied merge) are
            //    counted per warp ...
            uint32_t kept_w = 0, napp_w = 0, nsel_w = 0;
            for (uint32_t i0 = wlo; i0 < whi; i0 += 32) {
                const uint32_t i = i0 + lane;
                const uint32_t st = i < whi ? sel[i] : HG_ST_NONE;
                uint32_t stl = __shfl_up_sync(0xFFFFFFFFu, st, 1);
                if (lane == 0) stl = i0 > 0 ? sel[i0 - 1] : HG_ST_NONE;
                kept_w += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, i < whi && stl != HG_ST_APP));
                napp_w += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, st == HG_ST_APP));
                nsel_w += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, st != HG_ST_NONE));
            }
            if (lane == 0) { s_cnt[warp] = kept_w; s_par[warp] = nsel_w; s_tmp[warp] = napp_w; }
            __syncthreads();
            uint32_t run = 0, all_app = 0, all_cand = 0, all_kept = 0;
            for (uint32_t w = 0; w < NW; ++w) { if (w < warp) run += s_cnt[w]; all_kept += s_cnt[w]; all_cand += s_par[w]; all_app += s_tmp[w]; }
            // ... and written to their new places
            for (uint32_t i0 = wlo; i0 < whi; i0 += 32) {
                const uint32_t j = i0 + lane;
                const bool in = j < whi;
                const uint32_t st = in ? sel[j] : HG_ST_NONE;
                uint32_t stl = __shfl_up_sync(0xFFFFFFFFu, st, 1), st1 = __shfl_down_sync(0xFFFFFFFFu, st, 1), st2 = __shfl_down_sync(0xFFFFFFFFu, st, 2);
                if (lane == 0) stl = i0 > 0 ? sel[i0 - 1] : HG_ST_NONE;
                if (lane == 31) st1 = j + 1 < m ? sel[j + 1] : HG_ST_NONE;
                if (lane >= 30) st2 = j + 2 < m ? sel[j + 2] : HG_ST_NONE;
                const bool keep = in && stl != HG_ST_APP;
                const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    const uint32_t q = run + (uint32_t)__popc(km & ((1u << lane) - 1u));
                    uint32_t nr;
                    if (st == HG_ST_APP) {
                        const uint32_t r = rk[j];
                        id2[q] = r;
                        if (j + 2 >= m) nr = TK_INF;
                        else if (st2 == HG_ST_APP) nr = rk[j + 2] >= r ? rL[j + 2] : rR[j];   // the later of the two merges saw the other's result
                        else nr = rR[j];
                    } else {
                        id2[q] = id[j];
                        if (j + 1 >= m) nr = TK_INF;
                        else nr = st1 == HG_ST_APP ? rL[j + 1] : rk[j];
                    }
                    rk2[q] = nr;
                }
                run += (uint32_t)__popc(km);
            }
            const bool was_cut = kcut != ~0ull || khaz != ~0ull;
            if (t == 0) {
                atomicAdd(flags + (delta ? TKK_S_ROUNDSM - TKK_S_FLAGS : TKK_S_ROUNDS1 - TKK_S_FLAGS), 1u);
                if (delta) atomicAdd(flags + (TKK_S_APPLIEDM - TKK_S_FLAGS), all_app);
                if (was_cut) atomicAdd(flags + (TKK_S_ROUNDSCUT - TKK_S_FLAGS), 1u);
            }
            m = all_kept;
            { uint32_t* z = id; id = id2; id2 = z; z = rk; rk = rk2; rk2 = z; }
            __syncthreads();
            // the window: narrow after a cut; widen when the round went through and candidates are sparse (every merge of a
            // dense round is already a good round: a run of equal pairs halves)
            if (was_cut) delta >>= 2;
            else if (all_cand * 16u < m) delta = delta ? (delta < (1u << 19) ? delta * 2u : delta) : 64u;
        }
        const uint32_t outn = m;
        for (uint32_t i = t; i < m; i += HG_T) if (TK_DBG(recs[r].tok_base + i, pool_words)) out[i] = id[i];
        if (t == 0) { recs[r].count = outn; atomicAdd(tile_count + pos / TKK_COUNT_TILE, (unsigned long long)outn); }
    }
}

// =====================================================================================================
// K2: lookup tiles.  One block per LK_TILE bytes of text, one lane per piece: the whole-piece
// vocabulary lookup (CoreBPE's shortcut).  Every piece of at most TK_LANE_MAX bytes gets room in the
// tile's slot of the rank stream: 1 slot if it is a vocabulary entry (written here, last-of-piece
// flag in bit 31), `len` slots otherwise -- those pieces are queued, by length class, for K2m, which
// writes their ranks at the front of the reserved slots.  (A piece that needs merging always yields
// >= 2 ranks, so "the first rank carries the flag" tells the reader which kind a piece is.)
// K2m: one lane per queued piece: exact byte_pair_merge (tk_bpe_merge_loop).  Persistent warps take
// 32 queue entries at a time; a queue holds one length class, so the lanes of a warp run merge
// chains of similar length, every lane always has a piece, and nothing waits for a straggler.
// K4 (emit) turns the stream into the final ids.
// =====================================================================================================
#define LK_T 256
#define LK_WINS 128
#define LK_TILE (LK_WINS * 32)
#define LK_CAP (LK_TILE + TK_LANE_MAX)      // slots a tile can need (pieces that start in it)
#define LK_PCAP (LK_TILE + 4)               // pieces that can start in a tile (+ the end sentinel)
#define EN_INVALID 0xFFFFFFFFu               // stream word: no rank at this byte position
#define EN_LONGREF 0xFFFFFFFEu               // stream word: a long piece starts here (its ranks are in K3's pool)
#define QE_START_BITS 40
#define QE_LEN_BITS 7

// length class of a piece that goes to K2m: 2..4, 5..8, 9..12, 13..16, 17..24, 25..32, 33..48, 49..64, 65..96 bytes
__host__ __device__ __forceinline__ uint32_t lane_class(uint32_t len) {
    if (len <= 8u) return len <= 4u ? 0u : 1u;
    if (len > 64u) return 8u;
    const uint32_t p = 31u - (uint32_t)TK_CLZ(len - 1u);             // 3, 4, 5 for 9..16, 17..32, 33..64
    return 2u * (p - 2u) + (((len - 1u) >> (p - 1u)) & 1u);
}

// LK_COMPOSE=1: the tile's rank-stream words are composed in shared memory and stored once (16 KB more per block:
// five resident blocks instead of eight).  0: the words are filled with EN_INVALID in global memory and the ranks
// stored over them -- the L2 writes fully dirty lines back early, so the stream goes to DRAM twice (ncu: 8.1 GB
// written by this kernel for a 3.9 GB stream).
#ifndef LK_COMPOSE
#define LK_COMPOSE 0
#endif
struct LkSmem {
    uint8_t bytes[LK_TILE + TK_LANE_MAX + 16];   // 16-byte aligned
    uint32_t mask[LK_WINS + 4];
    uint16_t list[LK_PCAP];            // tile-relative starts of all pieces, in order
    uint32_t missq[LK_TILE / 2 + 4];   // pieces to merge: start | len << 12
    uint32_t wsum[LK_T / 32];
    uint32_t cls_n[TKK_N_CLASSES], cls_base[TKK_N_CLASSES], cls_pos[TKK_N_CLASSES];
    uint32_t n_pieces, n_miss, n_hit;
    unsigned long long mbar;           // mbarrier of the tile's bulk copy
#if LK_COMPOSE
    alignas(16) uint32_t out[LK_TILE]; // the tile's stream words, composed here and written once
#endif
};

// tile-relative end of the piece that starts at tile-relative byte s (the next set bit of the start
// mask), or 0xFFFFFFFF if it is more than three mask words away (a long piece)
__device__ __forceinline__ uint32_t en_piece_end(const uint32_t* mask, uint32_t s) {
    const uint32_t w = s >> 5, b = s & 31u;
    const uint32_t m = b == 31u ? 0u : (mask[w] >> (b + 1u));
    if (m) return s + (uint32_t)__ffs((int)m);
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const uint32_t mm = mask[w + k];
        if (mm) return (w + k) * 32u + (uint32_t)(__ffs((int)mm) - 1);
    }
    return 0xFFFFFFFFu;
}

// block-wide exclusive prefix of one value per thread (LK_T threads); *total = sum over the block
__device__ __forceinline__ uint32_t lk_block_excl(uint32_t v, uint32_t* wsum, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    __syncthreads();          // wsum may still be read from an earlier use
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < LK_T / 32; ++w) { if (w < (int)warp) before += wsum[w]; all += wsum[w]; }
    *total = all;
    return before + inc - v;
}

#ifndef LK_MINB
#define LK_MINB 8      // 32 registers, eight resident blocks (measured: 3.12 ms against 3.42 ms at 34 registers)
#endif
__global__ void __launch_bounds__(LK_T, LK_MINB) lookup_kernel(const uint8_t* __restrict__ data, uint64_t n,
                                                      const uint32_t* __restrict__ start_mask, TkDeviceTables T,
                                                      uint32_t* __restrict__ stream, unsigned long long* __restrict__ queues,
                                                      TkkQueueLayout Q, uint32_t* __restrict__ q_n,
                                                      unsigned long long* __restrict__ tile_count) {
    pdl_wait();
    static_assert(LK_TILE == TKK_COUNT_TILE, "token counts are kept per lookup tile");
    __shared__ __align__(16) LkSmem S;
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t tile = blockIdx.x;
    const uint64_t tile_pos = (uint64_t)tile * LK_TILE;
    const uint64_t win0 = (uint64_t)tile * LK_WINS;
    if (t < TKK_N_CLASSES) { S.cls_n[t] = 0; S.cls_pos[t] = 0; }
    if (t == 0) { S.n_miss = 0; S.n_hit = 0; }
    // ---- A: stage bytes and mask words; list the piece starts ----
    bool bulk;
    {
        const uint64_t avail = n > tile_pos ? n - tile_pos : 0;
        constexpr uint32_t want = LK_TILE + TK_LANE_MAX + 16;
        static_assert(want % 16 == 0, "bulk copies move multiples of 16 bytes");
        bulk = avail >= want;                              // a tile inside the text: one bulk copy (TMA), issued by thread 0,
        if (bulk) {                                        // in flight while the block lists the piece starts
            if (t == 0) bulk_tile_begin(S.bytes, data + tile_pos, want, &S.mbar);
        } else {
            const uint32_t full16 = (uint32_t)(avail / 16);
            uint4* dst = reinterpret_cast<uint4*>(S.bytes);
            const uint4* src = reinterpret_cast<const uint4*>(data + tile_pos);
            for (uint32_t i = t; i < full16; i += LK_T) dst[i] = __ldg(src + i);
            for (uint32_t i = full16 * 16 + t; i < want; i += LK_T) S.bytes[i] = (tile_pos + i < n) ? data[tile_pos + i] : 0;
        }
        uint32_t m = 0;
        if (t < LK_WINS + 4) { m = start_mask[win0 + t]; S.mask[t] = m; }
        if (t >= LK_WINS) m = 0;
        uint32_t np;
        uint32_t o = lk_block_excl((uint32_t)__popc(m), S.wsum, &np);
        if (t == 0) S.n_pieces = np;
        while (m) {
            S.list[o++] = (uint16_t)(t * 32u + (uint32_t)(__ffs((int)m) - 1));
            m &= m - 1;
        }
    }
    // the tile's stream words start out EN_INVALID (coalesced 16-byte stores); ranks and long-piece marks are
    // stored over them after the barrier (same block: ordered), K2m fills in the merged pieces later
#if LK_COMPOSE
    uint32_t* dst = S.out;
    {
        const uint4 inv = make_uint4(EN_INVALID, EN_INVALID, EN_INVALID, EN_INVALID);
        uint4* d4 = reinterpret_cast<uint4*>(S.out);
        for (uint32_t i = t; i < LK_TILE / 4; i += LK_T) d4[i] = inv;
    }
#else
    uint32_t* dst = stream + tile_pos;
    {
        const uint4 inv = make_uint4(EN_INVALID, EN_INVALID, EN_INVALID, EN_INVALID);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t i = t; i < LK_TILE / 4; i += LK_T) if (TK_DBG(tile_pos + 4u * i + 3u, stream_words)) d4[i] = inv;
    }
#endif
    __threadfence_block();
    __syncthreads();
    if (bulk) bulk_tile_wait(&S.mbar);
    const uint32_t np = S.n_pieces;

    // ---- B: one lane per piece: whole-piece vocabulary lookup; entries written, misses collected ----
    uint32_t hits = 0;                                     // ranks this thread stored, summed once after the loop
    for (uint32_t k0 = 0; k0 < np; k0 += LK_T) {
        const uint32_t k = k0 + t;
        uint32_t s = 0, len = 0, cls = 0xFFFFFFFFu;
        uint32_t hit = 0;                                  // ranks this lane stored (1 for a vocabulary entry, 2..4 for a tiny piece)
        if (k < np) {
            s = S.list[k];
            if (tile_pos + s < n) {                        // the end-of-data sentinel is not a piece
                const uint32_t e = en_piece_end(S.mask, s);
                if (e != 0xFFFFFFFFu && e - s <= TK_LANE_MAX) {
                    len = e - s;
                    const uint32_t whole = tk_vocab_lookup_w32(T, S.bytes, s, len);
                    if (!TK_DBG(tile_pos + s, stream_words)) { }
                    else if (whole != TK_INF) { dst[s] = whole; hit = 1; }
                    else if (len == 1) { dst[s] = (uint32_t)S.bytes[s]; hit = 1; }
                    else if (len <= 4u && s + len <= LK_TILE) {
                        // 2..4 bytes: the merge loop written out (tk_bpe_tiny), ranks straight into the piece's own positions
                        // (a piece that reaches into the next tile's words is queued: that tile's block clears them)
                        uint32_t r[4];
                        hit = tk_bpe_tiny(T, S.bytes[s], S.bytes[s + 1], S.bytes[s + 2], S.bytes[s + 3], len, r);
                        dst[s] = r[0]; dst[s + 1] = r[1];
                        if (hit > 2u) dst[s + 2] = r[2];
                        if (hit > 3u) dst[s + 3] = r[3];
                    }
                    else cls = lane_class(len);
                } else if (TK_DBG(tile_pos + s, stream_words)) dst[s] = EN_LONGREF;      // longer pieces: K3
            }
        }
        hits += hit;
        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, cls != 0xFFFFFFFFu);
        if (mm) {
            uint32_t base = 0;
            const int leader = __ffs((int)mm) - 1;
            if ((int)lane == leader) base = atomicAdd(&S.n_miss, (uint32_t)__popc(mm));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (cls != 0xFFFFFFFFu) S.missq[base + (uint32_t)__popc(mm & ((1u << lane) - 1u))] = s | (len << 12);
            // misses per length class (one shared-memory atomic per class per warp)
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
            if (cls != 0xFFFFFFFFu && (uint32_t)(__ffs((int)peers) - 1) == lane) atomicAdd(&S.cls_n[cls], (uint32_t)__popc(peers));
        }
    }
    {
        const uint32_t hm = __reduce_add_sync(0xFFFFFFFFu, hits);
        if (lane == 0 && hm) atomicAdd(&S.n_hit, hm);
    }
    __syncthreads();
#if LK_COMPOSE
    {
        uint4* g4 = reinterpret_cast<uint4*>(stream + tile_pos);
        const uint4* s4 = reinterpret_cast<const uint4*>(S.out);
        for (uint32_t i = t; i < LK_TILE / 4; i += LK_T) if (TK_DBG(tile_pos + 4u * i + 3u, stream_words)) g4[i] = s4[i];
    }
#endif
    if (t < TKK_N_CLASSES && S.cls_n[t]) S.cls_base[t] = atomicAdd(q_n + t, S.cls_n[t]);   // this tile's range of every queue
    if (t == 32 && S.n_hit) atomicAdd(tile_count + tile, (unsigned long long)S.n_hit);
    __syncthreads();
    // ---- C: misses into the queue of their length class ----
    const uint32_t nm = S.n_miss;
    for (uint32_t i0 = 0; i0 < nm; i0 += LK_T) {
        const uint32_t i = i0 + t;
        uint32_t e = 0, cls = 0xFFFFFFFFu;
        if (i < nm) { e = S.missq[i]; cls = lane_class(e >> 12); }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
        if (cls != 0xFFFFFFFFu) {
            const int leader = __ffs((int)peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&S.cls_pos[cls], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            const uint32_t pos = S.cls_base[cls] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (TK_DBG(Q.off[cls] + pos, queue_words)) queues[Q.off[cls] + pos] = (tile_pos + (e & 4095u)) | ((unsigned long long)(e >> 12) << QE_START_BITS);
        }
    }
}

// the aligned 32-bit words that hold bytes [start, start + len) of the text (NW = MAXLEN / 4 + 1 of them cover any
// piece of the class); words that do not overlap the piece are 0
template <int NW>
__device__ __forceinline__ void lm_load_words(const uint8_t* __restrict__ data, uint64_t n, uint64_t start, uint32_t len,
                                              uint32_t (&w)[NW]) {
    const uint64_t base = start & ~3ull, end = start + len;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint64_t a = base + 4u * k;
        uint32_t v = 0;
        if (len && a < end) {
            if (a + 4u <= n) v = __ldg(reinterpret_cast<const uint32_t*>(data + a));
            else for (uint32_t c = 0; c < 4u; ++c) if (a + c < n) v |= (uint32_t)__ldg(data + a + c) << (8u * c);   // last word of the text
        }
        w[k] = v;
    }
}

// register budgets of the three short classes (resident blocks the compiler plans for), from A/B builds
#ifndef LM_MINB_12
#define LM_MINB_12 6
#endif
#ifndef LM_MINB_8
#define LM_MINB_8 8
#endif
#ifndef LM_MINB_4
#define LM_MINB_4 6
#endif
template <int MAXLEN, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) lanemerge_kernel(const uint8_t* __restrict__ data, uint64_t n, TkDeviceTables T,
                                                            const unsigned long long* __restrict__ queue,
                                                            const uint32_t* __restrict__ q_n, uint32_t* __restrict__ q_w,
                                                            uint32_t* __restrict__ stream, unsigned long long* __restrict__ tile_count,
                                                            unsigned long long* __restrict__ stats) {
    pdl_wait();
    constexpr int STRIDE = MAXLEN + 1;      // odd: lane i's arrays start at bank i (no conflicts when lanes sweep together)
    extern __shared__ __align__(16) uint32_t lm_raw[];
    uint32_t* id = lm_raw + threadIdx.x * STRIDE;
    uint32_t* key = lm_raw + THREADS * STRIDE + threadIdx.x * STRIDE;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t total = *q_n;
    // short pieces cost about the same: warps stride over the queue.  Long ones vary more: warps take
    // the next 32 entries from a work counter.
    constexpr bool kDynamic = MAXLEN >= 12;
    // Software pipeline over a warp's batches of 32 entries: the queue entries are read two batches ahead and, for
    // the short classes (where the two dependent round trips entry -> text bytes are most of a piece's time), the
    // aligned words of the text that hold the piece one batch ahead, in registers.
    constexpr bool kWordsAhead = MAXLEN <= 16;
    constexpr int NW = MAXLEN / 4 + 1;
    const uint32_t warps = gridDim.x * (THREADS / 32);
    auto next_batch = [&](uint32_t prev) -> uint32_t {
        if (prev >= total) return prev;                  // past the end: stay there
        uint32_t k = prev + warps * 32u;
        if (kDynamic) {
            if (lane == 0) k = atomicAdd(q_w, 32u);
            k = __shfl_sync(0xFFFFFFFFu, k, 0);
        } else if (k < prev) k = 0xFFFFFFFFu;            // wrapped
        return k;
    };
    auto entry_of = [&](uint32_t k0) -> unsigned long long {    // 0 = no piece (len field 0)
        return (k0 < total && k0 + lane < total) ? queue[k0 + lane] : 0ull;
    };
    uint32_t k1;                                          // batch whose entry (and words) are loaded
    if (kDynamic) {
        k1 = 0;
        if (lane == 0) k1 = atomicAdd(q_w, 32u);
        k1 = __shfl_sync(0xFFFFFFFFu, k1, 0);
    } else k1 = (blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 32u;
    unsigned long long e1 = entry_of(k1);
    uint32_t w1[kWordsAhead ? NW : 1];
    if (kWordsAhead)
        lm_load_words<kWordsAhead ? NW : 1>(data, n, e1 & ((1ull << QE_START_BITS) - 1ull),
                                            (uint32_t)(e1 >> QE_START_BITS) & ((1u << QE_LEN_BITS) - 1u), w1);
    uint32_t k2 = next_batch(k1);
    unsigned long long e2 = entry_of(k2);
    uint32_t n_pair = 0, n_bp = 0;                          // table lookups issued by this lane (reported by bench.py)
    while (k1 < total) {
        const unsigned long long e = e1;
        uint32_t w[kWordsAhead ? NW : 1];
        if (kWordsAhead) {
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = w1[k];
        }
        // advance the pipeline before working on this batch
        k1 = k2; e1 = e2;
        if (kWordsAhead)
            lm_load_words<kWordsAhead ? NW : 1>(data, n, e1 & ((1ull << QE_START_BITS) - 1ull),
                                                (uint32_t)(e1 >> QE_START_BITS) & ((1u << QE_LEN_BITS) - 1u), w1);
        k2 = next_batch(k2);
        e2 = entry_of(k2);

        const uint64_t start = e & ((1ull << QE_START_BITS) - 1ull);
        const uint32_t len = (uint32_t)(e >> QE_START_BITS) & ((1u << QE_LEN_BITS) - 1u);
        if (len) {
            if (kWordsAhead) {
                // realign (funnel shift by the start's offset in its word), then one byte per part
                const uint32_t sh = ((uint32_t)start & 3u) * 8u;
#pragma unroll
                for (int k = 0; k < MAXLEN / 4; ++k) {
                    const uint32_t r = __funnelshift_r(w[k], w[k + 1], sh);
                    id[4 * k] = r & 0xFFu; id[4 * k + 1] = (r >> 8) & 0xFFu; id[4 * k + 2] = (r >> 16) & 0xFFu; id[4 * k + 3] = r >> 24;
                }
            } else {
                const uint8_t* b = data + start;
                for (uint32_t i = 0; i < len; i += 4) {
                    uint32_t v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = i + c < len ? (uint32_t)__ldg(b + i + c) : 0u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (i + c < len) id[i + c] = v[c];
                }
            }
            // parts = single bytes; rank of every adjacent byte pair from the direct table.  Slots past the
            // piece get TK_INF: the merge loop scans all MAXLEN of them.
#pragma unroll(MAXLEN <= 12 ? MAXLEN : 8)
            for (int j = 0; j < MAXLEN; ++j) {
                uint32_t r = TK_INF;
                if ((uint32_t)j + 1u < len) r = __ldg(T.byte_pair + ((id[j] << 8) | id[j + 1]));
                key[j] = r == TK_INF ? TK_INF : ((r << TK_KEY_SHIFT) | (uint32_t)j);
            }
            using Mask = typename std::conditional<(MAXLEN <= 32), uint32_t,
                                                   typename std::conditional<(MAXLEN <= 64), unsigned long long, tk_u128>::type>::type;
            n_bp += len - 1u;
            Mask live = tk_bpe_merge_loop<Mask, MAXLEN>(T, len, id, key, n_pair);
            {
                // the ranks go to consecutive byte positions from `start`: count them for the tile each one lands in
                const uint32_t cnt = tk_popc_m(live);
                const uint64_t tile0 = start / TKK_COUNT_TILE;
                const uint32_t room = (uint32_t)((tile0 + 1) * TKK_COUNT_TILE - start);
                atomicAdd(tile_count + tile0, (unsigned long long)(cnt < room ? cnt : room));
                if (cnt > room) atomicAdd(tile_count + tile0 + 1, (unsigned long long)(cnt - room));
            }
            // ranks to stream[start ...] (the piece's remaining positions stay EN_INVALID)
            uint32_t* dst = stream + start;
            while (live) {
                const uint32_t j = tk_ffs_m(live) - 1u;
                live &= live - 1;
                if (TK_DBG(dst - stream, stream_words)) *dst = id[j];
                ++dst;
            }
        }
        __syncwarp();
    }
    n_pair = __reduce_add_sync(0xFFFFFFFFu, n_pair);
    n_bp = __reduce_add_sync(0xFFFFFFFFu, n_bp);
    if (lane == 0 && (n_pair | n_bp)) { atomicAdd(stats, (unsigned long long)n_pair); atomicAdd(stats + 1, (unsigned long long)n_bp); }
}

// =====================================================================================================
// K3s: exclusive prefix of the per-tile token counts (lookup, lane-merge, long-piece and document
// kernels add to them) -> first output position of every tile.  Three tiny kernels: block sums,
// scan of the block sums, apply.
// =====================================================================================================
#define TS_T 256
#define TS_PER 8
#define TS_BLOCK (TS_T * TS_PER)

__device__ __forceinline__ unsigned long long ts_block_excl(unsigned long long v, unsigned long long* wsum, unsigned long long* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < TS_T / 32; ++w) { if (w < (int)warp) before += wsum[w]; all += wsum[w]; }
    *total = all;
    return before + inc - v;
}

__global__ void __launch_bounds__(TS_T) tilesum_kernel(const unsigned long long* __restrict__ tile_count, uint32_t n_tiles,
                                                       unsigned long long* __restrict__ bsum) {
    pdl_wait();
    __shared__ unsigned long long wsum[TS_T / 32];
    unsigned long long v = 0;
#pragma unroll
    for (int j = 0; j < TS_PER; ++j) {
        const uint64_t i = (uint64_t)blockIdx.x * TS_BLOCK + (uint64_t)j * TS_T + threadIdx.x;
        if (i < n_tiles) v += tile_count[i];
    }
    unsigned long long total;
    ts_block_excl(v, wsum, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(TS_T) tilescan_kernel(unsigned long long* __restrict__ bsum, uint32_t n_blocks, uint64_t out_cap,
                                                        unsigned long long* __restrict__ total_out, uint32_t* __restrict__ flags) {
    pdl_wait();
    __shared__ unsigned long long wsum[TS_T / 32];
    unsigned long long run = 0;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += TS_T) {
        const uint32_t i = b0 + threadIdx.x;
        const unsigned long long v = i < n_blocks ? bsum[i] : 0ull;
        unsigned long long total;
        const unsigned long long e = ts_block_excl(v, wsum, &total);
        if (i < n_blocks) bsum[i] = run + e;          // in place: sum -> exclusive prefix
        run += total;
    }
    if (threadIdx.x == 0) {
        *total_out = run;
        if (run > out_cap) atomicOr(flags, TKK_FLAG_OUT_FULL);
    }
}

__global__ void __launch_bounds__(TS_T) tileapply_kernel(const unsigned long long* __restrict__ tile_count, uint32_t n_tiles,
                                                         const unsigned long long* __restrict__ bbase,
                                                         unsigned long long* __restrict__ tile_base) {
    pdl_wait();
    __shared__ unsigned long long wsum[TS_T / 32];
    unsigned long long run = bbase[blockIdx.x];
#pragma unroll
    for (int j = 0; j < TS_PER; ++j) {
        const uint64_t i = (uint64_t)blockIdx.x * TS_BLOCK + (uint64_t)j * TS_T + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_count[i] : 0ull;
        unsigned long long total;
        const unsigned long long e = ts_block_excl(v, wsum, &total);
        if (i < n_tiles) tile_base[i] = run + e;
        run += total;
    }
}

// =====================================================================================================
// K4: emit.  The rank stream holds, for every byte position of the text, the rank of the token that
// the reference emits there, EN_INVALID, or EN_LONGREF; emit is a compaction of it.  One block per
// 4 KiB of text, one warp per 512 positions, one 32-byte window per warp iteration (lane = position in
// the window): count (ranks + the long piece's tokens + BOS/EOS of the documents that start here) per
// window, prefix over windows / warps, ids (+num_special) written from the tile's first output position
// (known from K3s: no tile waits for another); per-document token offsets.  Long pieces are copied from
// K3's pool by the whole block.
// =====================================================================================================
#define E3_T 256
#define E3_PER 16
#define E3_LONGCAP (LK_TILE / 64 + 2)

struct E3Long {
    unsigned long long src, dst;
    uint32_t count, pad;
};

// BOS/EOS tokens emitted at the start position of documents d .. d+k-1 (the EOS of the document before
// each, the BOS of each; document n_docs is the virtual end document)
__device__ __forceinline__ uint32_t e3_specials(uint64_t d, uint32_t k, uint64_t n_docs, uint32_t add_bos, uint32_t add_eos) {
    const uint32_t n_eos = k - (d == 0 ? 1u : 0u), n_bos = k - (d + k - 1 == n_docs ? 1u : 0u);
    return (add_eos ? n_eos : 0u) + (add_bos ? n_bos : 0u);
}

// number of documents that start at batch byte position s, given the index of the first of them
__device__ __forceinline__ uint32_t docs_from(const uint64_t* __restrict__ doc_off, uint64_t n_docs, uint64_t s, uint64_t first) {
    uint64_t e = first + 1;
    while (e <= n_docs && doc_off[e] == s) ++e;     // doc_off has n_docs+1 entries; the last one is the virtual end doc
    return (uint32_t)(e - first);
}

#ifndef E3_MINB
#define E3_MINB 5
#endif
// state of one lane's position in a window that needs the slow path (a document starts in it or a long piece)
struct E3Slow {
    uint64_t d;          // first document that starts at my position
    uint32_t nd;         // documents that start at my position
    uint32_t extra;      // BOS/EOS ids emitted before my position's own ids
    uint32_t own;        // my position's own ids: 1 (a rank), the long piece's count, or 0
};

__device__ __forceinline__ E3Slow e3_slow_lane(uint32_t word, uint32_t wds, uint64_t w_first, uint32_t w_cnt, uint64_t win_pos,
                                               uint32_t lane, const uint64_t* __restrict__ doc_off, uint64_t off_base,
                                               uint64_t n_docs, uint32_t add_bos, uint32_t add_eos, uint32_t long_count) {
    E3Slow r;
    r.d = 0; r.nd = 0; r.extra = 0;
    r.own = word < EN_LONGREF ? 1u : (word == EN_LONGREF ? long_count : 0u);
    if ((wds >> lane) & 1u) {
        // Usual case: one start position in the 32-byte window; then the window's document count (K0) says how
        // many documents start there and no offset has to be read.  Several start positions in one window
        // (documents shorter than 32 bytes): read offsets.
        uint64_t d = w_first;
        if ((wds & (wds - 1u)) == 0u) r.nd = w_cnt;
        else {
            uint32_t earlier = wds & ((1u << lane) - 1u);
            while (earlier) {
                const uint32_t b = (uint32_t)(__ffs((int)earlier) - 1);
                earlier &= earlier - 1;
                d += docs_from(doc_off, n_docs, win_pos + b + off_base, d);
            }
            r.nd = docs_from(doc_off, n_docs, win_pos + lane + off_base, d);
        }
        r.d = d;
        r.extra = e3_specials(d, r.nd, n_docs, add_bos, add_eos);
    }
    return r;
}

__device__ __forceinline__ unsigned long long e3_warp_incl(unsigned long long v) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= (uint32_t)d) v += o;
    }
    return v;
}

// One block per 4 KiB tile, one warp per 512 positions, INTERLEAVED: in iteration k lane l holds position
// 32 k + l of the warp's range (one 32-byte window per iteration).  The ids of a window are then consecutive
// in the output, so a store instruction of the warp writes one contiguous run (two or three 32-byte sectors)
// instead of one sector per lane, and the offset of a lane's id is a ballot + popcount.  Windows in which a
// document starts or a long piece sits (warp-uniform test) take a slower path with per-lane counts.
__global__ void __launch_bounds__(E3_T, E3_MINB) emit_kernel(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ ds_mask,
                                                    const uint32_t* __restrict__ doc_first, const uint32_t* __restrict__ doc_cnt,
                                                    const uint32_t* __restrict__ long_of_word,
                                                    const TkkLongRec* __restrict__ recs, const uint32_t* __restrict__ pool,
                                                    const uint64_t* __restrict__ doc_off, uint64_t off_base, uint64_t n_docs,
                                                    uint32_t add_bos, uint32_t add_eos, uint32_t nsp, uint32_t bos_id, uint32_t eos_id,
                                                    uint32_t* __restrict__ out, uint64_t out_cap, uint64_t* __restrict__ tok_off,
                                                    const unsigned long long* __restrict__ tile_base) {
    pdl_wait();
    static_assert(E3_T * E3_PER == LK_TILE && E3_PER == 16, "one emit block per lookup tile, 16 windows per warp");
    __shared__ unsigned long long wsum[E3_T / 32];
    __shared__ E3Long longs[E3_LONGCAP];
    __shared__ uint32_t n_longs;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t tile = blockIdx.x;
    if (t == 0) n_longs = 0;
    const uint64_t wp0 = (uint64_t)tile * LK_TILE + (uint64_t)warp * (32u * E3_PER);   // first position of my warp
    const uint64_t word0 = wp0 >> 5;                                                    // its first mask word
    // ---- loads: 16 coalesced rows of the stream; lanes 0..15 hold the window tables of window `lane` ----
    uint32_t w[E3_PER];
#pragma unroll
    for (int k = 0; k < E3_PER; ++k) w[k] = __ldg(stream + wp0 + 32u * k + lane);
    uint32_t wds_l = 0, wfirst_l = 0, wcnt_l = 0, wlong_l = 0;
    if (lane < E3_PER) {
        wds_l = ds_mask[word0 + lane];
        wfirst_l = doc_first[word0 + lane];
        wcnt_l = doc_cnt[word0 + lane];
        wlong_l = long_of_word[word0 + lane];
    }
    // ---- pass 1: ids per window (lane k keeps the count of window k) ----
    // Windows by kind (bit k = window k): `single` = documents start at exactly one position and no long piece:
    // the BOS/EOS count is a property of the window (lane k computes it from its own table entries) and shifts
    // the lanes from the start position on; `general` = several start positions or a long piece (the long-piece
    // table is only read for windows whose stream row marks one: it is not written past the end of the text).
    uint32_t long_windows = 0;
    {
        bool mine = false;
#pragma unroll
        for (int k = 0; k < E3_PER; ++k) mine |= w[k] == EN_LONGREF;
        if (__any_sync(0xFFFFFFFFu, mine)) {                     // rare: a long piece starts in my warp's range
#pragma unroll
            for (int k = 0; k < E3_PER; ++k)
                if (__ballot_sync(0xFFFFFFFFu, w[k] == EN_LONGREF)) long_windows |= 1u << k;
        }
    }
    const uint32_t single_windows = __ballot_sync(0xFFFFFFFFu, wds_l != 0u && (wds_l & (wds_l - 1u)) == 0u) & ~long_windows;
    const uint32_t general_windows = (__ballot_sync(0xFFFFFFFFu, wds_l != 0u) | long_windows) & ~single_windows;
    const uint32_t extra_l = ((single_windows >> lane) & 1u) ? e3_specials(wfirst_l, wcnt_l, n_docs, add_bos, add_eos) : 0u;
    unsigned long long cnt_l = 0;
#pragma unroll
    for (int k = 0; k < E3_PER; ++k) {
        const uint32_t c = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, w[k] < EN_LONGREF));
        if (lane == (uint32_t)k) cnt_l = c;
    }
    cnt_l += extra_l;
    // general windows: one compact loop (not unrolled: the kernel has to stay small for the instruction cache), the
    // window's stream row is read again (an L1 hit)
#pragma unroll 1
    for (uint32_t sw = general_windows; sw; sw &= sw - 1) {
        const uint32_t k = (uint32_t)(__ffs((int)sw) - 1);
        const uint32_t wk = __ldg(stream + wp0 + 32u * k + lane);
        const uint32_t wds = __shfl_sync(0xFFFFFFFFu, wds_l, k);
        const uint32_t wl = ((long_windows >> k) & 1u) ? __shfl_sync(0xFFFFFFFFu, wlong_l, k) : 0u;
        const uint32_t lc = wl ? recs[wl - 1].count : 0u;
        const E3Slow sl = e3_slow_lane(wk, wds, __shfl_sync(0xFFFFFFFFu, wfirst_l, k), __shfl_sync(0xFFFFFFFFu, wcnt_l, k),
                                       wp0 + 32u * k, lane, doc_off, off_base, n_docs, add_bos, add_eos, lc);
        unsigned long long mine = (unsigned long long)sl.extra + sl.own;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
        if (lane == k) cnt_l = mine;
    }
    // ---- exclusive prefix over the windows of the warp, then over the warps of the block ----
    const unsigned long long inc = e3_warp_incl(cnt_l);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long warp_out = tile_base[tile];
    for (uint32_t x = 0; x < warp; ++x) warp_out += wsum[x];
    const unsigned long long excl_l = warp_out + inc - cnt_l;                            // lane k: first output index of window k
    // ---- pass 2: write ----
    const uint32_t le_mask = (2u << lane) - 1u;                                          // lanes 0..lane
    if ((single_windows | general_windows) == 0u) {
        // no document start and no long piece in my warp's 512 positions (warp-uniform, the usual case): a window adds
        // at most 32 ids, so offsets inside the warp's range fit 32 bits and need one shuffle per window
        const uint32_t rel_l = (uint32_t)(inc - cnt_l);
#pragma unroll
        for (int k = 0; k < E3_PER; ++k) {
            const uint32_t rel = __shfl_sync(0xFFFFFFFFu, rel_l, k);
            const bool valid = w[k] < EN_LONGREF;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid) {
                const unsigned long long o = warp_out + (rel + (uint32_t)__popc(m & (le_mask >> 1)));
                if (o < out_cap) out[o] = w[k] + nsp;
            }
        }
    } else {
#pragma unroll
    for (int k = 0; k < E3_PER; ++k) {
        const unsigned long long base = __shfl_sync(0xFFFFFFFFu, excl_l, k);
        const uint32_t ex = __shfl_sync(0xFFFFFFFFu, extra_l, k), wds = __shfl_sync(0xFFFFFFFFu, wds_l, k);
        const bool valid = w[k] < EN_LONGREF;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid && !((general_windows >> k) & 1u)) {
            // in a single window the BOS/EOS ids come before the id of the start position itself
            const unsigned long long o = base + (unsigned long long)__popc(m & (le_mask >> 1)) + ((wds & le_mask) ? ex : 0u);
            if (o < out_cap) out[o] = w[k] + nsp;
        }
    }
    // single windows: the lane of the start position writes the BOS/EOS ids and the documents' offsets
#pragma unroll 1
    for (uint32_t sw = single_windows; sw; sw &= sw - 1) {
        const uint32_t k = (uint32_t)(__ffs((int)sw) - 1);
        const unsigned long long base = __shfl_sync(0xFFFFFFFFu, excl_l, k);
        const uint32_t wds = __shfl_sync(0xFFFFFFFFu, wds_l, k);
        const uint32_t nd = __shfl_sync(0xFFFFFFFFu, wcnt_l, k);
        const uint64_t d0 = __shfl_sync(0xFFFFFFFFu, wfirst_l, k);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, __ldg(stream + wp0 + 32u * k + lane) < EN_LONGREF);
        if ((wds >> lane) & 1u) {
            unsigned long long o = base + (unsigned long long)__popc(m & (le_mask >> 1));
            for (uint64_t x = d0; x < d0 + nd; ++x) {
                if (x > 0 && add_eos) { if (o < out_cap) out[o] = eos_id; ++o; }
                if (TK_DBG(x, n_docs)) tok_off[x] = o;
                if (x < n_docs && add_bos) { if (o < out_cap) out[o] = bos_id; ++o; }
            }
        }
    }
#pragma unroll 1
    for (uint32_t sw = general_windows; sw; sw &= sw - 1) {
        const uint32_t k = (uint32_t)(__ffs((int)sw) - 1);
        const unsigned long long base = __shfl_sync(0xFFFFFFFFu, excl_l, k);
        const uint32_t wk = __ldg(stream + wp0 + 32u * k + lane);
        const uint32_t wds = __shfl_sync(0xFFFFFFFFu, wds_l, k);
        const uint32_t wl = ((long_windows >> k) & 1u) ? __shfl_sync(0xFFFFFFFFu, wlong_l, k) : 0u;
        TkkLongRec lr;
        lr.count = 0; lr.tok_base = 0;
        if (wl) lr = recs[wl - 1];
        const E3Slow sl = e3_slow_lane(wk, wds, __shfl_sync(0xFFFFFFFFu, wfirst_l, k), __shfl_sync(0xFFFFFFFFu, wcnt_l, k),
                                       wp0 + 32u * k, lane, doc_off, off_base, n_docs, add_bos, add_eos, lr.count);
        const unsigned long long mine = (unsigned long long)sl.extra + sl.own;
        unsigned long long o = base + e3_warp_incl(mine) - mine;
        for (uint64_t x = sl.d; x < sl.d + sl.nd; ++x) {
            if (x > 0 && add_eos) { if (o < out_cap) out[o] = eos_id; ++o; }
            tok_off[x] = o;
            if (x < n_docs && add_bos) { if (o < out_cap) out[o] = bos_id; ++o; }
        }
        if (wk < EN_LONGREF) { if (o < out_cap) out[o] = wk + nsp; }
        else if (wk == EN_LONGREF) {
            E3Long L;
            L.src = lr.tok_base; L.dst = o; L.count = lr.count; L.pad = 0;
            longs[atomicAdd(&n_longs, 1u)] = L;
        }
    }
    }   // warps with document starts or long pieces
    __syncthreads();
    const uint32_t nl = n_longs;
    for (uint32_t l = 0; l < nl; ++l) {
        const E3Long L = longs[l];
        for (uint32_t i = t; i < L.count; i += E3_T)
            if (L.dst + i < out_cap) out[L.dst + i] = pool[L.src + i] + nsp;
    }
}

#include "tk_small.cuh"

// =====================================================================================================
// launch sequence
// =====================================================================================================
static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

size_t encode_workspace_bytes(uint64_t n, uint64_t n_docs, EncodeLayout* L) {
    EncodeLayout l{};
    const uint64_t n_windows = n / 32 + 1;
    const uint64_t n_tiles = ceil_div(n_windows, PT_T);
    const uint64_t words = n_tiles * PT_T + 64;     // padded so tile kernels may read a little past the end
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    l.n_windows = n_windows; l.n_tiles = n_tiles; l.mask_words = words;
    l.off_small = take(256);                            // counters + flags (zeroed every call)
    l.off_ds = take(words * 4);
    l.off_start = take(words * 4);
    l.off_longword = take(words * 4);
    l.off_docfirst = take(words * 4);
    l.off_doccnt = take(words * 4);
    l.off_summ = take(n_tiles * sizeof(TkkTileSummary));
    l.off_carry = take(n_tiles * 4);
    l.off_worklist = take(n_tiles * 4);
    l.off_seg = take((ceil_div(n_tiles, SG_TILES) + 1) * 3 * 4);
    l.n_ltiles = n_tiles * (PT_T / LK_WINS);            // lookup tiles cover exactly the emit tiles
    l.off_tilecount = take((l.n_ltiles + 1) * 8);
    l.off_tilebase = take((l.n_ltiles + 1) * 8);
    l.off_bsum = take((ceil_div(l.n_ltiles, TS_BLOCK) + 1) * 8);
    l.off_stream = take((l.n_ltiles * (size_t)LK_TILE + TK_LANE_MAX + 64) * 4);
    {
        // a queue per length class; a class whose shortest piece has m bytes holds at most n/m + 1 pieces
        const uint32_t shortest[TKK_N_CLASSES] = {2, 5, 9, 13, 17, 25, 33, 49, 65};
        uint64_t e = 0;
        for (int c = 0; c < TKK_N_CLASSES; ++c) { l.queues.off[c] = e; e += n / shortest[c] + 32; }
        l.queue_words = e;
        l.off_queues = take(e * 8);
    }
    l.max_long = n / (TK_LANE_MAX + 1) + 2;
    l.off_recs = take(l.max_long * sizeof(TkkLongRec));
    l.off_huge = take(l.max_long * 8);                  // two lists: pieces of more than HG_SPLIT bytes, and the others
    l.off_pool = take((n + 16) * 4);
    (void)n_docs;
    l.total = off;
    if (L) *L = l;
    return off;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// The 256-byte counter block goes to mapped pinned host memory with a kernel, not a copy: a copy
// would queue behind the large id transfers of other chunks on the device-to-host copy engine.
__global__ void publish_kernel(const uint32_t* __restrict__ small, uint32_t* __restrict__ mapped) {
    pdl_wait();
    mapped[threadIdx.x] = small[threadIdx.x];
    __threadfence_system();
}
cudaError_t publish_small(const void* d_small, uint32_t* mapped_dev, cudaStream_t st) {
    return launch_chain(publish_kernel, 1, 64, 0, st, (const uint32_t*)d_small, mapped_dev);
}
cudaError_t publish_counters(const void* d_ws, const EncodeLayout& L, uint32_t* mapped_dev, cudaStream_t st) {
    return launch_chain(publish_kernel, 1, 64, 0, st, (const uint32_t*)((const unsigned char*)d_ws + L.off_small), mapped_dev);
}

template <int MAXLEN, int THREADS, int MINB>
static cudaError_t launch_lanemerge(int blocks_per_sm, int sm_count, const uint8_t* d_data, uint64_t n, const TkDeviceTables& T,
                                    const unsigned long long* queue, const uint32_t* q_n, uint32_t* q_w, uint32_t* stream,
                                    unsigned long long* tile_count, unsigned long long* stats, const HotTables* hot, cudaStream_t st) {
    const size_t smem = (size_t)2 * THREADS * (MAXLEN + 1) * sizeof(uint32_t);
    static std::atomic<uint64_t> attr_set{0};   // bit per device ordinal
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (!((attr_set.load() >> (dev & 63)) & 1ull)) {
        CK(cudaFuncSetAttribute(lanemerge_kernel<MAXLEN, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.fetch_or(1ull << (dev & 63));
    }
    // Launch attribute: an L2 access-policy window over the pair + byte-pair tables (persisting), everything else this
    // kernel touches (queue entries, piece bytes, stream words: each used once) streaming.  Per launch, so the
    // caller's stream keeps its own attributes.  TEKKEN_B200_L2PIN=0 switches it off (A/B measurements).
    static const bool pin = [] { const char* e = getenv("TEKKEN_B200_L2PIN"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sm_count * blocks_per_sm));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (pin && hot && hot->enabled) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = hot->ptr;
        attr[0].val.accessPolicyWindow.num_bytes = hot->bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 1;
    }
    if (pdl_enabled()) {
        attr[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
        ++cfg.numAttrs;
    }
    CK(cudaLaunchKernelEx(&cfg, lanemerge_kernel<MAXLEN, THREADS, MINB>, d_data, n, T, queue, q_n, (uint32_t*)q_w, stream, tile_count, stats));
    count_launch();
    return cudaSuccess;
}

cudaError_t encode_device(const TkDeviceTables& T, const uint8_t* d_data, const uint64_t* d_doc_off, uint64_t off_base,
                          uint64_t n_docs, uint64_t n, int add_bos, int add_eos, uint32_t* d_out, uint64_t out_cap, uint64_t* d_tok_off,
                          void* d_ws, const EncodeLayout& L, uint32_t* d_scratch, uint64_t scratch_cap, int sm_count,
                          cudaStream_t st, StageTimer* timer, const HotTables* hot, const CfgSplitTables* cfg) {
    unsigned char* ws = (unsigned char*)d_ws;
    uint32_t* small = (uint32_t*)(ws + L.off_small);
    uint32_t* ds = (uint32_t*)(ws + L.off_ds);
    uint32_t* start = (uint32_t*)(ws + L.off_start);
    uint32_t* longword = (uint32_t*)(ws + L.off_longword);
    uint32_t* docfirst = (uint32_t*)(ws + L.off_docfirst);
    uint32_t* doccnt = (uint32_t*)(ws + L.off_doccnt);
    TkkTileSummary* summ = (TkkTileSummary*)(ws + L.off_summ);
    uint32_t* carry = (uint32_t*)(ws + L.off_carry);
    uint32_t* worklist = (uint32_t*)(ws + L.off_worklist);
    unsigned long long* tile_count = (unsigned long long*)(ws + L.off_tilecount);
    unsigned long long* tile_base = (unsigned long long*)(ws + L.off_tilebase);
    unsigned long long* bsum = (unsigned long long*)(ws + L.off_bsum);
    TkkLongRec* recs = (TkkLongRec*)(ws + L.off_recs);
    uint32_t* huge = (uint32_t*)(ws + L.off_huge);
    uint32_t* pool = (uint32_t*)(ws + L.off_pool);
    uint32_t* stream = (uint32_t*)(ws + L.off_stream);
    unsigned long long* queues = (unsigned long long*)(ws + L.off_queues);
    // small block layout
    uint32_t* flags = small + TKK_S_FLAGS;
    unsigned long long* err_pos = (unsigned long long*)(small + TKK_S_ERRPOS);
    uint32_t* work_count = small + TKK_S_WORKCOUNT;
    uint32_t* n_long = small + TKK_S_NLONG;
    uint32_t* n_huge = small + TKK_S_NHUGE;
    uint32_t* wc_long = small + TKK_S_WC_LONG;
    uint32_t* wc_huge = small + TKK_S_WC_HUGE;
    uint32_t* n_mid = small + TKK_S_NMID;
    uint32_t* wc_mid = small + TKK_S_WC_MID;
    uint32_t* q_n = small + TKK_S_QN;
    uint32_t* q_w = small + TKK_S_QW;
    unsigned long long* pool_cursor = (unsigned long long*)(small + TKK_S_POOLCUR);
    unsigned long long* scratch_cursor = (unsigned long long*)(small + TKK_S_SCRCUR);
    unsigned long long* total_out = (unsigned long long*)(small + TKK_S_TOTAL);

#ifdef TK_DEBUG_BOUNDS
    {
        DbgLimits lim{};
        lim.mask_words = L.mask_words; lim.stream_words = L.n_ltiles * (unsigned long long)LK_TILE + TK_LANE_MAX + 64;
        lim.queue_words = L.queue_words; lim.pool_words = n + 16; lim.scratch_words = scratch_cap; lim.count_tiles = L.n_ltiles + 1;
        lim.out_cap = out_cap; lim.n_docs = n_docs + 1; lim.max_long = L.max_long;
        CK(cudaMemcpyToSymbolAsync(g_dbg, &lim, sizeof lim, 0, cudaMemcpyHostToDevice, st));
    }
#endif
    if (timer) timer->mark(st, "setup");
    {
        SetupArgs sa;
        sa.small = small;
        sa.zero0 = (uint4*)ds; sa.n0 = L.mask_words / 4;                   // mask_words is a multiple of 64
        sa.ones = (uint4*)docfirst; sa.n1 = L.mask_words / 4;
        sa.zero1 = (uint4*)doccnt; sa.n2 = L.mask_words / 4;
        sa.tail = start + L.n_windows; sa.n3 = L.mask_words - L.n_windows;
        sa.counts = tile_count; sa.n4 = L.n_ltiles + 1;
        const uint64_t work = L.mask_words / 4 + 64;
        const unsigned grid = (unsigned)std::min<uint64_t>(ceil_div(work, 256), (uint64_t)sm_count * 8);
        CK(launch_chain(setup_kernel, grid, 256, 0, st, sa));
    }
    CK(launch_chain(docmark_kernel, (unsigned)ceil_div(n_docs + 1, 256), 256, 0, st, d_doc_off, off_base, n_docs, n, add_bos ? 1u : 0u, add_eos ? 1u : 0u, ds, docfirst, doccnt, tile_count, flags));
    if (timer) timer->mark(st, "pretok");
    if (cfg) {
        // the pattern stored in tekken.json: safe starts by bit logic, then one matcher walk per safe start.  The safe
        // mask lives in the long-piece index array, which is not written before K2a.
        const TkCfgTables CT{cfg->stage1, cfg->stage2};
        CK(launch_chain(cfg_mask_kernel, (unsigned)L.n_tiles, PT_T, 0, st, d_data, n, ds, longword, start, L.n_windows, CT, err_pos));
        if (timer) timer->mark(st, "pretok_walk");
        CK(launch_chain(cfg_walk_kernel, (unsigned)ceil_div(L.n_windows, CW_T), CW_T, 0, st, d_data, n, ds, longword, start, L.n_windows, CT, err_pos));
    } else {
    CK(launch_chain(pretok_kernel, (unsigned)L.n_tiles, PT_T, 0, st, d_data, n, ds, start, L.n_windows, T, summ, err_pos));
    if (timer) timer->mark(st, "pretok_carry");
    {
        const uint32_t n_seg = (uint32_t)ceil_div(L.n_tiles, SG_TILES);
        uint32_t* seg_packed = (uint32_t*)(ws + L.off_seg);
        uint32_t* seg_in = seg_packed + n_seg;
        uint32_t* seg_after = seg_in + n_seg;
        CK(launch_chain(pretok_seg_kernel, n_seg, SG_T, 0, st, summ, (uint32_t)L.n_tiles, seg_packed));
        CK(launch_chain(pretok_segscan_kernel, 1, SC_T, 0, st, seg_packed, n_seg, seg_in, seg_after));
        CK(launch_chain(pretok_apply_kernel, n_seg, SG_T, 0, st, summ, (uint32_t)L.n_tiles, seg_in, seg_after, carry, worklist, work_count, start));
    }
    CK(launch_chain(pretok_fix_kernel, (unsigned)(L.n_tiles < (uint64_t)(2 * sm_count) ? L.n_tiles : (uint64_t)(2 * sm_count)), PT_T, 0, st, d_data, n, ds, start, L.n_windows, T, carry, worklist, work_count, err_pos));
    }
    if (timer) timer->mark(st, "longmark");
    CK(launch_chain(longmark_kernel, (unsigned)ceil_div(L.n_windows, 256), 256, 0, st, start, L.n_windows, n, longword, recs, n_long));
    if (timer) timer->mark(st, "longmerge");
    {
        uint64_t blocks = ceil_div(L.max_long, LM_WARPS);
        const uint64_t cap = (uint64_t)sm_count * 12;
        if (blocks > cap) blocks = cap;
        uint32_t* mid = huge + L.max_long;
        CK(launch_chain(longmerge_warp_kernel, (unsigned)blocks, LM_WARPS * 32, 0, st, d_data, start, T, recs, n_long, pool, pool_cursor, huge,
                                                                        n_huge, mid, n_mid, wc_long, tile_count, L.n_windows, n, flags));
        uint64_t hb = L.max_long < (uint64_t)(2 * sm_count) ? L.max_long : (uint64_t)(2 * sm_count);
        CK(launch_chain(longmerge_block_kernel<HG_T_MID>, (unsigned)std::min<uint64_t>(L.max_long, (uint64_t)sm_count * 8), HG_T_MID, 0, st, d_data, T, recs,
                        mid, n_mid, pool, d_scratch, scratch_cap, scratch_cursor, wc_mid, flags, tile_count));
        CK(launch_chain(longmerge_block_kernel<HG_T_BIG>, (unsigned)hb, HG_T_BIG, 0, st, d_data, T, recs, huge, n_huge, pool, d_scratch, scratch_cap,
                                                             scratch_cursor, wc_huge, flags, tile_count));
    }
    if (timer) timer->mark(st, "lookup");
    {
        CK(launch_chain(lookup_kernel, (unsigned)L.n_ltiles, LK_T, 0, st, d_data, n, start, T, stream, queues, L.queues, q_n, tile_count));
    }
    // resident blocks per SM of the lane-merge launches, longest class first (what shared memory allows;
    // tuning knob: TEKKEN_B200_LM_BPS="a,b,...")
    static const std::array<int, TKK_N_CLASSES> bps = [] {
        std::array<int, TKK_N_CLASSES> v{4, 3, 4, 3, 4, 4, 6, 8, 6};
        if (const char* e = getenv("TEKKEN_B200_LM_BPS")) {
            for (int i = 0; i < TKK_N_CLASSES && *e; ++i) {
                char* end = nullptr;
                const long x = strtol(e, &end, 10);
                if (end == e) break;
                if (x > 0 && x <= 32) v[i] = (int)x;
                e = *end == ',' ? end + 1 : end;
            }
        }
        return v;
    }();
#define TK_LANEMERGE(MAXLEN, THREADS, MINB, CLS, NAME)                                                                            \
    if (timer) timer->mark(st, NAME);                                                                                             \
    CK((launch_lanemerge<MAXLEN, THREADS, MINB>(bps[TKK_N_CLASSES - 1 - CLS], sm_count, d_data, n, T, queues + L.queues.off[CLS],     \
                                                q_n + CLS, q_w + CLS, stream, tile_count, (unsigned long long*)(small + TKK_S_PAIRLOOK), hot, st)));
    // MINB (the resident blocks the compiler plans registers for) is set from measurements: capping the 12- and
    // 16-byte classes at 32 / 40 registers for more resident warps made them 1.4x slower
    TK_LANEMERGE(96, 64, 1, 8, "lanemerge96")
    TK_LANEMERGE(64, 128, 1, 7, "lanemerge64")
    TK_LANEMERGE(48, 128, 1, 6, "lanemerge48")
    TK_LANEMERGE(32, 256, 1, 5, "lanemerge32")
    TK_LANEMERGE(24, 256, 1, 4, "lanemerge24")
    TK_LANEMERGE(16, 256, 1, 3, "lanemerge16")
    TK_LANEMERGE(12, 256, LM_MINB_12, 2, "lanemerge12")
    TK_LANEMERGE(8, 256, LM_MINB_8, 1, "lanemerge8")
    TK_LANEMERGE(4, 256, LM_MINB_4, 0, "lanemerge4")
#undef TK_LANEMERGE
    if (timer) timer->mark(st, "emit");
    {
        const uint32_t nb = (uint32_t)ceil_div(L.n_ltiles, TS_BLOCK);
        CK(launch_chain(tilesum_kernel, nb, TS_T, 0, st, tile_count, (uint32_t)L.n_ltiles, bsum));
        CK(launch_chain(tilescan_kernel, 1, TS_T, 0, st, bsum, nb, out_cap, total_out, flags));
        CK(launch_chain(tileapply_kernel, nb, TS_T, 0, st, tile_count, (uint32_t)L.n_ltiles, bsum, tile_base));
    }
    CK(launch_chain(emit_kernel, (unsigned)L.n_ltiles, E3_T, 0, st, stream, ds, docfirst, doccnt, longword, recs, pool, d_doc_off, off_base, n_docs,
                                                      add_bos ? 1u : 0u, add_eos ? 1u : 0u, T.num_special, T.bos_id, T.eos_id, d_out,
                                                      out_cap, d_tok_off, tile_base));
    if (timer) timer->mark(st, "end");
    return cudaGetLastError();
}

// (debug build) violations recorded by the kernels of this device since the last call; out: {count, line, index, limit}
long long debug_bounds_violations(unsigned long long* out4) {
#ifdef TK_DEBUG_BOUNDS
    unsigned long long h[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(h, g_dbg_hit, sizeof h) != cudaSuccess) return -2;
    cudaMemcpyToSymbol(g_dbg_hit, z, sizeof z);
    if (out4) for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return (long long)h[0];
#else
    (void)out4;
    return -1;
#endif
}

cudaError_t encode_small(const TkDeviceTables& T, const uint8_t* d_text, uint32_t n, int add_bos, int add_eos, uint32_t* d_out,
                         uint32_t seq, cudaStream_t st) {
    return encode_small_launch(T, d_text, n, add_bos, add_eos, d_out, seq, st);
}

// =====================================================================================================
// Id packing for the trip over PCIe (host-buffer calls): ids are < 2^18 for Tekken (131,072 + specials), so 18 of the
// 32 bits carry everything.  One thread packs 16 ids into bits / 2 words; all shifts are compile-time constants.
// =====================================================================================================
template <int BITS>
__global__ void __launch_bounds__(256) pack_ids_kernel(const uint32_t* __restrict__ ids, uint64_t n, uint32_t* __restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;       // group of 16 ids
    const uint64_t i0 = g * 16u;
    if (i0 >= n) return;
    uint32_t v[16];
    if (i0 + 16u <= n && ((uintptr_t)ids & 15u) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + i0) + q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = i0 + k < n ? __ldg(ids + i0 + k) : 0u;
    }
    constexpr int W = BITS / 2;
    constexpr uint32_t MASK = (1u << BITS) - 1u;
    uint32_t* o = out + g * W;
    unsigned long long acc = 0;
    int fill = 0, wi = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        acc |= (unsigned long long)(v[k] & MASK) << fill;
        fill += BITS;
        if (fill >= 32) { o[wi++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
    }
}

cudaError_t pack_ids(const uint32_t* d_ids, uint64_t n, int bits, void* d_out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned blocks = (unsigned)(((n + 15) / 16 + 255) / 256);
    if (bits == 18) pack_ids_kernel<18><<<blocks, 256, 0, st>>>(d_ids, n, (uint32_t*)d_out);
    else if (bits == 24) pack_ids_kernel<24><<<blocks, 256, 0, st>>>(d_ids, n, (uint32_t*)d_out);
    else return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

}  // namespace tkk
