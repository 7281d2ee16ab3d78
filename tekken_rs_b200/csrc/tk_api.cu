// tk_api.cu -- the extern "C" boundary declared in include/tekken_b200.h.
//
// Host-side glue only: argument checks, device memory for the vocabulary tables and per-call
// workspaces, host<->device copies for the host-buffer entry points, and the mapping from
// device-reported conditions to the reference's TokenizerError variants (src/errors.rs:23-59).
// All arithmetic of the path runs in the kernels (tk_kernels.cu, tk_decode.cu); there is no CPU
// implementation of encode or decode in this library.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tekken_b200.h"
#include "tk_host.h"
#include "tk_kernels.h"

// ------------------------------------------------------------------------------------------ errors

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

extern "C" const char* tk_last_error(void) { return g_last_error.c_str(); }

extern "C" const char* tk_status_name(int s) {
    switch (s) {
        case TK_OK: return "Ok";
        case TK_ERR_IO: return "Io";
        case TK_ERR_JSON: return "Json";
        case TK_ERR_BASE64: return "Base64";
        case TK_ERR_TOKENIZERS: return "Tokenizers";
        case TK_ERR_AUDIO: return "Audio";
        case TK_ERR_INVALID_CONFIG: return "InvalidConfig";
        case TK_ERR_TOKEN_NOT_FOUND: return "TokenNotFound";
        case TK_ERR_SPECIAL_TOKEN_POLICY: return "SpecialTokenPolicy";
        case TK_ERR_UNSUPPORTED_FORMAT: return "UnsupportedFormat";
        case TK_ERR_INVALID_UTF8: return "InvalidUtf8";
        case TK_ERR_CUDA: return "Cuda";
        case TK_ERR_BUFFER_TOO_SMALL: return "BufferTooSmall";
        case TK_ERR_INVALID_ARGUMENT: return "InvalidArgument";
    }
    return "Unknown";
}

#define CUDA_OR_FAIL(x)                                                                           \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) return fail(TK_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_));    \
    } while (0)

// ------------------------------------------------------------------------------------------ buffers

// Output buffers handed to the caller are pinned host allocations; freed ones are kept in a
// small cache so a steady stream of calls does not pay cudaHostAlloc every time.
namespace {
struct HostPool {
    std::mutex mu;
    std::map<void*, size_t> live;                 // ptr -> capacity
    std::multimap<size_t, void*> free_list;       // capacity -> ptr
    size_t cached = 0;
    static constexpr size_t kMaxCached = 6ull << 30;   // enough for the id buffers of two 1 GB batches in flight

    // small results (accessor strings, the ids of one short text) are plain heap memory: pinning 4 KiB for a
    // five-byte answer costs more than it saves, and a host-only handle must not create a CUDA context
    static constexpr size_t kSmall = 16384;
    void* get(size_t bytes, bool pinned = true) {
        if (bytes == 0) bytes = 1;
        std::lock_guard<std::mutex> g(mu);
        if (!pinned || bytes <= kSmall) {
            void* p = malloc(bytes);
            if (p) live[p] = bytes | 1;
            return p;
        }
        auto it = free_list.lower_bound(bytes);
        if (it != free_list.end() && it->first <= bytes * 2 + 4096) {
            void* p = it->second;
            live[p] = it->first;
            cached -= it->first;
            free_list.erase(it);
            return p;
        }
        void* p = nullptr;
        // capacities come in steps of 1/8 of the size's power of two, so that the slightly different result sizes of
        // successive batches of one corpus reuse each other's buffers instead of pinning a new 2 GB every call
        // (plus a tenth of headroom when a new buffer has to be pinned: estimates of one corpus scatter by a few per cent)
        size_t step = 4096;
        while (step * 16 <= bytes) step <<= 1;
        const size_t want = bytes >= ((size_t)64 << 20) ? bytes + bytes / 10 : bytes;
        size_t cap = (want + step - 1) / step * step;
        if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            // no CUDA context (host-only handle): plain memory
            p = malloc(cap);
            if (!p) return nullptr;
            live[p] = cap | 1;   // low bit: malloc'ed
            return p;
        }
        live[p] = cap;
        return p;
    }
    // drop the cached (not the live) pinned buffers: called when the last tokenizer handle goes away
    void trim() {
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : free_list) cudaFreeHost(kv.second);
        free_list.clear();
        cached = 0;
    }
    void put(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        auto it = live.find(p);
        if (it == live.end()) return;
        size_t cap = it->second;
        live.erase(it);
        if (cap & 1) { free(p); return; }
        if (cached + cap > kMaxCached) { cudaFreeHost(p); return; }
        cached += cap;
        free_list.emplace(cap, p);
    }
};
HostPool g_pool;
std::atomic<int> g_live_handles{0};
}  // namespace

extern "C" void tk_buffer_free(void* p) { g_pool.put(p); }

// ------------------------------------------------------------------------------------------ handle

// Is a CUDA context current on this thread?  (cudaGetDevice answers 0 also when none is, and restoring "device 0"
// would then create a primary context -- a few hundred MB -- on a GPU this process never meant to touch.)
static bool thread_has_context() {
    typedef int (*ctx_fn)(void**);
    static const ctx_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (ctx_fn) nullptr;
        }
        return (ctx_fn)p;
    }();
    if (!fn) return true;
    void* ctx = nullptr;
    return fn(&ctx) == 0 && ctx != nullptr;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (thread_has_context()) cudaGetDevice(&prev);
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);   // retry without slack
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct tk_tokenizer {
    tk::HostModel host;
    int device = -1;
    int sm_count = 148;
    TkDeviceTables tables{};
    std::vector<void*> table_allocs;
    std::mutex mu;                 // serialises device work issued through this handle
    cudaStream_t stream = nullptr; // used by the host-buffer entry points
    DevBuf ws, in_data, in_off, out_a, out_b, status;     // decode + single-shot paths
    struct EncSlot {
        cudaStream_t st = nullptr;
        DevBuf ws, scratch, in_data, in_off, out_tok, out_off;
        uint32_t* h_small = nullptr;   // mapped pinned copy of the workspace's counter block
        uint32_t* d_small_map = nullptr;   // its device address
        cudaEvent_t done = nullptr;    // recorded after the counters have been published
        cudaEvent_t ev_in = nullptr, ev_out = nullptr;   // pipeline: text has arrived / ids have left
        uint8_t* h_stage = nullptr;    // pinned staging of pageable caller memory (offsets, then text)
        size_t stage_cap = 0;
        tkk::EncodeLayout L;
    };
    static constexpr int kSlots = 4;
    EncSlot slot[kSlots];          // host-buffer encode: chunks of a batch pipeline through these
    // ids of a chunk wait here for their download.  There are more of these than of the slots above: the download
    // engine is the slowest stage of a large call (4 bytes come back per id), and with as many id buffers as text
    // buffers the kernels -- and behind them the uploads -- were paced by it for the whole call, so uploads and
    // downloads shared the bus from start to end.  With the kernels free to run ahead the text is up after half the
    // call and the ids have the bus to themselves for the rest.
    struct OutSlot {
        DevBuf tok, off, pack;             // pack: the ids as an 18/24-bit stream (what crosses PCIe in large calls)
        cudaEvent_t ev_out = nullptr;      // the slot's ids have left
    };
    // pinned landing buffers of the packed ids; a host thread widens them into the result (encode_worker)
    struct PackStage {
        uint8_t* h = nullptr;
        size_t cap = 0;
        cudaEvent_t ev = nullptr;          // the packed ids have arrived
        cudaEvent_t ev_packed = nullptr;   // the pack kernel has run
        std::atomic<int> busy{0};
    };
    static constexpr int kPackStages = 4;
    PackStage pstage[kPackStages];
    size_t pack_stage_want = 0;
    static constexpr int kOutSlots = 12;
    OutSlot oslot[kOutSlots];
    struct DecSlot {
        DevBuf ids, off, out, boff, status, ws;
        uint32_t* h_small = nullptr;
        uint32_t* d_small_map = nullptr;
        cudaEvent_t done = nullptr, ev_in = nullptr, ev_out = nullptr;
        uint8_t* h_stage = nullptr;
        size_t stage_cap = 0;
        tkk::DecodeLayout L;
    };
    DecSlot dslot[kSlots];         // host-buffer decode
    cudaStream_t pipe_st[4] = {nullptr, nullptr, nullptr, nullptr};   // ... on an upload, a kernel, a download and a pack stream
    EncSlot dev_slot;              // device-pointer encode (caller's stream)
    // the latency path (tk_encode of a short text): mapped pinned in/out buffers + a stream per slot
    struct FastSlot {
        std::mutex mu;
        cudaStream_t st = nullptr;
        uint8_t* h_in = nullptr; uint8_t* d_in = nullptr;
        uint32_t* h_out = nullptr; uint32_t* d_out = nullptr;
        uint32_t seq = 0;
    };
    static constexpr int kFastSlots = 4;
    FastSlot fast[kFastSlots];
    std::atomic<float> ratio_hint{0.f};   // ids per text byte of recent large host-buffer calls (sizes the next result buffer)
    bool timing = false;
    bool counted = false;          // included in g_live_handles
    tkk::HotTables hot;            // pair + byte-pair table allocation (L2 persistence window of the merge kernels)
    int split_mode = TK_SPLIT_REFERENCE;
    tkk::CfgSplitTables cfg;       // device class tables of the TK_SPLIT_CONFIG split (null pointers otherwise)
    const tkk::CfgSplitTables* cfg_ptr() const { return split_mode == TK_SPLIT_CONFIG ? &cfg : nullptr; }
    tkk::StageTimer timer;
    std::vector<std::string> stage_names;
    std::vector<float> stage_ms;
};

template <class V>
static cudaError_t upload(tk_tokenizer* t, const V& v, const void** out) {
    void* p = nullptr;
    size_t bytes = v.size() * sizeof(v[0]);
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    t->table_allocs.push_back(p);
    if (bytes) e = cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice);
    *out = p;
    return e;
}

static int finish_handle(tk_tokenizer* t, int device, int split_mode, tk_tokenizer** out) {
    t->device = device;
    if (split_mode != TK_SPLIT_REFERENCE && split_mode != TK_SPLIT_CONFIG) {
        delete t;
        return fail(TK_ERR_INVALID_ARGUMENT, "unknown split mode %d", split_mode);
    }
    if (split_mode == TK_SPLIT_CONFIG && t->host.pattern != tk::tekken_config_pattern()) {
        delete t;
        return fail(TK_ERR_INVALID_CONFIG, "TK_SPLIT_CONFIG: the pattern stored in the tokenizer file is not the Tekken pattern this library implements");
    }
    t->split_mode = split_mode;
    if (device >= 0) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || device >= count) {
            cudaGetLastError();
            delete t;
            return fail(TK_ERR_CUDA, "CUDA device %d is not available (%s); this library has no CPU fallback", device,
                        e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range");
        }
        DeviceGuard dg(device);
        cudaError_t err = dg.ok ? cudaSuccess : cudaErrorInvalidDevice;
        const tk::HostModel& h = t->host;
        TkDeviceTables& T = t->tables;
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
        if (err == cudaSuccess) err = upload(t, h.uni_stage1, (const void**)&T.uni_stage1);
        if (err == cudaSuccess) err = upload(t, h.uni_stage2, (const void**)&T.uni_stage2);
        if (err == cudaSuccess) err = upload(t, h.vocab_slots, (const void**)&T.vocab_slots);
        if (err == cudaSuccess) {
            // the merge kernels' tables (pair table + byte-pair table, 8.25 MB for the Tekken vocabulary) live in ONE
            // allocation so that one L2 access-policy window can keep them resident while text / queue / stream
            // traffic passes through (tk_kernels.cu launch_lanemerge)
            const size_t pb = h.pair_slots.size() * sizeof(uint64_t), bb = h.byte_pair.size() * sizeof(uint32_t),
                         kb = h.pair_buckets.size() * sizeof(uint64_t);
            void* hot = nullptr;
            err = cudaMalloc(&hot, pb + bb + kb + 32);
            if (err == cudaSuccess) {
                t->table_allocs.push_back(hot);
                err = cudaMemcpy(hot, h.pair_slots.data(), pb, cudaMemcpyHostToDevice);
                if (err == cudaSuccess) err = cudaMemcpy((char*)hot + pb, h.byte_pair.data(), bb, cudaMemcpyHostToDevice);
                if (err == cudaSuccess && kb) err = cudaMemcpy((char*)hot + pb + bb, h.pair_buckets.data(), kb, cudaMemcpyHostToDevice);
                T.pair_slots = (const uint64_t*)hot;
                T.byte_pair = (const uint32_t*)((const char*)hot + pb);
                T.pair_buckets = kb ? (const uint64_t*)((const char*)hot + pb + bb) : nullptr;     // pb, bb: multiples of 32 bytes
                T.bucket_mask = h.bucket_mask;
                t->hot.ptr = hot;
                t->hot.bytes = pb + bb + kb;
                int max_persist = 0, max_window = 0;
                cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
                cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
                const size_t want = std::min<size_t>((size_t)max_persist, std::max<size_t>(t->hot.bytes * 2, (size_t)32 << 20));
                if (max_persist > 0 && (size_t)max_window >= t->hot.bytes && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
                    t->hot.enabled = true;
                else cudaGetLastError();
            }
        }
        if (err == cudaSuccess) err = upload(t, h.vocab_pad16, (const void**)&T.vocab_pad16);
        if (err == cudaSuccess) err = upload(t, h.vocab_e16, (const void**)&T.vocab_e16);
        if (err == cudaSuccess) err = upload(t, h.vocab_bytes, (const void**)&T.vocab_bytes);
        if (err == cudaSuccess) err = upload(t, h.vocab_off, (const void**)&T.vocab_off);
        if (err == cudaSuccess) err = upload(t, h.special_bytes, (const void**)&T.special_bytes);
        if (err == cudaSuccess) err = upload(t, h.special_off, (const void**)&T.special_off);
        if (err == cudaSuccess && split_mode == TK_SPLIT_CONFIG) {
            std::vector<uint16_t> s1;
            std::vector<uint8_t> s2;
            tk::build_cfg_unicode_tables(s1, s2);
            err = upload(t, s1, (const void**)&t->cfg.stage1);
            if (err == cudaSuccess) err = upload(t, s2, (const void**)&t->cfg.stage2);
        }
        if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
        T.vocab_mask = (uint32_t)h.vocab_slots.size() - 1;
        T.pair_mask = h.pair_mask;
        T.n_vocab = (uint32_t)h.n_vocab();
        T.num_special = (uint32_t)h.num_special;
        T.max_token_len = h.max_token_len;
        T.bos_id = h.has_control_token("<s>") ? h.control_token("<s>") : TK_INF;
        T.eos_id = h.has_control_token("</s>") ? h.control_token("</s>") : TK_INF;
        if (err != cudaSuccess) {
            int rc = fail(TK_ERR_CUDA, "uploading vocabulary tables: %s", cudaGetErrorString(err));
            tk_free(t);
            return rc;
        }
    }
    *out = t;
    t->counted = true;
    g_live_handles.fetch_add(1);
    return TK_OK;
}

extern "C" int tk_load_file(const char* path, int device, tk_tokenizer** out) {
    return tk_load_file_ex(path, device, TK_SPLIT_REFERENCE, out);
}

extern "C" int tk_load_file_ex(const char* path, int device, int split_mode, tk_tokenizer** out) {
    if (!path || !out) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    tk_tokenizer* t = new tk_tokenizer();
    try {
        t->host = tk::HostModel::from_file(path);
    } catch (const tk::Error& e) {
        delete t;
        return fail(e.code, "%s", e.what());
    } catch (const std::exception& e) {
        delete t;
        return fail(TK_ERR_IO, "%s", e.what());
    }
    return finish_handle(t, device, split_mode, out);
}

extern "C" int tk_new(const tk_vocab_entry* vocab, size_t n_vocab, const tk_special_entry* special, size_t n_special,
                      const char* pattern, size_t vocab_size, size_t num_special_tokens, int version, int device,
                      tk_tokenizer** out) {
    return tk_new_ex(vocab, n_vocab, special, n_special, pattern, vocab_size, num_special_tokens, version, device, TK_SPLIT_REFERENCE, out);
}

extern "C" int tk_new_ex(const tk_vocab_entry* vocab, size_t n_vocab, const tk_special_entry* special, size_t n_special,
                         const char* pattern, size_t vocab_size, size_t num_special_tokens, int version, int device,
                         int split_mode, tk_tokenizer** out) {
    if (!out || (!vocab && n_vocab) || (!special && n_special)) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (version != TK_V3 && version != TK_V7 && version != TK_V11 && version != TK_V13)
        return fail(TK_ERR_INVALID_CONFIG, "Unknown version: %d", version);
    std::vector<tk::VocabEntry> v(n_vocab);
    for (size_t i = 0; i < n_vocab; ++i) {
        if (!vocab[i].token_bytes_b64) return fail(TK_ERR_INVALID_ARGUMENT, "vocab[%zu].token_bytes_b64 is null", i);
        v[i].rank = vocab[i].rank;
        v[i].token_bytes_b64 = vocab[i].token_bytes_b64;
    }
    std::vector<tk::SpecialEntry> s(n_special);
    for (size_t i = 0; i < n_special; ++i) {
        if (!special[i].token_str) return fail(TK_ERR_INVALID_ARGUMENT, "special[%zu].token_str is null", i);
        s[i].rank = special[i].rank;
        s[i].token_str = special[i].token_str;
        s[i].is_control = special[i].is_control != 0;
    }
    tk_tokenizer* t = new tk_tokenizer();
    try {
        t->host = tk::HostModel::build(v, s, pattern ? pattern : "", vocab_size, num_special_tokens, version);
    } catch (const tk::Error& e) {
        delete t;
        return fail(e.code, "%s", e.what());
    }
    return finish_handle(t, device, split_mode, out);
}

extern "C" size_t tk_deprecated_special_tokens(const tk_special_entry** out) {
    static std::vector<tk_special_entry> v;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const auto& e : tk::deprecated_special_tokens()) v.push_back({e.rank, e.token_str.c_str(), e.is_control ? 1 : 0});
    });
    if (out) *out = v.data();
    return v.size();
}

extern "C" void tk_free(tk_tokenizer* t) {
    if (!t) return;
    if (t->device >= 0) {
        DeviceGuard dg(t->device);
        t->timer.reset();
        for (void* p : t->table_allocs) cudaFree(p);
        t->ws.release(); t->in_data.release(); t->in_off.release();
        t->out_a.release(); t->out_b.release(); t->status.release();
        auto drop = [](tk_tokenizer::EncSlot& s) {
            if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
            s.ws.release(); s.scratch.release(); s.in_data.release(); s.in_off.release(); s.out_tok.release(); s.out_off.release();
            if (s.h_small) cudaFreeHost(s.h_small);
            if (s.h_stage) cudaFreeHost(s.h_stage);
            if (s.done) cudaEventDestroy(s.done);
            if (s.ev_in) cudaEventDestroy(s.ev_in);
            if (s.ev_out) cudaEventDestroy(s.ev_out);
        };
        for (auto& f : t->fast) {
            if (f.st) { cudaStreamSynchronize(f.st); cudaStreamDestroy(f.st); }
            if (f.h_in) cudaFreeHost(f.h_in);
            if (f.h_out) cudaFreeHost(f.h_out);
        }
        for (auto& ps : t->pipe_st) if (ps) { cudaStreamSynchronize(ps); cudaStreamDestroy(ps); }
        for (auto& sl : t->slot) drop(sl);
        for (auto& ps : t->pstage) {
            if (ps.h) cudaFreeHost(ps.h);
            if (ps.ev) cudaEventDestroy(ps.ev);
            if (ps.ev_packed) cudaEventDestroy(ps.ev_packed);
        }
        for (auto& o : t->oslot) {
            o.tok.release(); o.off.release(); o.pack.release();
            if (o.ev_out) cudaEventDestroy(o.ev_out);
        }
        for (auto& d : t->dslot) {
            d.ids.release(); d.off.release(); d.out.release(); d.boff.release(); d.status.release(); d.ws.release();
            if (d.h_small) cudaFreeHost(d.h_small);
            if (d.h_stage) cudaFreeHost(d.h_stage);
            if (d.done) cudaEventDestroy(d.done);
            if (d.ev_in) cudaEventDestroy(d.ev_in);
            if (d.ev_out) cudaEventDestroy(d.ev_out);
        }
        drop(t->dev_slot);
        if (t->stream) cudaStreamDestroy(t->stream);
    }
    const bool last = t->counted && g_live_handles.fetch_sub(1) == 1;
    delete t;
    if (last) g_pool.trim();
}

// ------------------------------------------------------------------------------------------ accessors

extern "C" size_t tk_vocab_size(const tk_tokenizer* t) { return t ? t->host.vocab_size : 0; }
extern "C" size_t tk_num_special_tokens(const tk_tokenizer* t) { return t ? t->host.num_special : 0; }
extern "C" int tk_version_of(const tk_tokenizer* t) { return t ? t->host.version : 0; }
extern "C" int tk_device_of(const tk_tokenizer* t) { return t ? t->device : -1; }
extern "C" int tk_split_mode_of(const tk_tokenizer* t) { return t ? t->split_mode : TK_SPLIT_REFERENCE; }

extern "C" int tk_get_control_token(const tk_tokenizer* t, const char* s, uint32_t* id) {
    if (!t || !s || !id) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    try {
        *id = t->host.control_token(s);
    } catch (const tk::Error& e) {
        return fail(e.code, "%s", e.what());
    }
    return TK_OK;
}
extern "C" int tk_bos_id(const tk_tokenizer* t, uint32_t* id) { return tk_get_control_token(t, "<s>", id); }
extern "C" int tk_eos_id(const tk_tokenizer* t, uint32_t* id) { return tk_get_control_token(t, "</s>", id); }
extern "C" int tk_pad_id(const tk_tokenizer* t, uint32_t* id) { return tk_get_control_token(t, "<pad>", id); }
extern "C" int tk_unk_id(const tk_tokenizer* t, uint32_t* id) { return tk_get_control_token(t, "<unk>", id); }
extern "C" int tk_is_special_token(const tk_tokenizer* t, uint32_t id) { return t && id < t->host.num_special; }
extern "C" int tk_is_byte(const tk_tokenizer* t, uint32_t id) {
    return t && id >= t->host.num_special && id - t->host.num_special < 256;
}
extern "C" int tk_vocab_piece(const tk_tokenizer* t, uint32_t id, const char** str, size_t* len) {
    if (!t || !str || !len) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    if (id >= t->host.vocab_strings.size())
        return fail(TK_ERR_INVALID_CONFIG, "Token ID %u is out of vocabulary range (0-%zu)", id, t->host.vocab_size - 1);
    *str = t->host.vocab_strings[id].data();
    *len = t->host.vocab_strings[id].size();
    return TK_OK;
}

static int give_bytes(const uint8_t* p, size_t n, uint8_t** out, size_t* n_out) {
    uint8_t* b = (uint8_t*)g_pool.get(n + 1);
    if (!b) return fail(TK_ERR_CUDA, "out of host memory");
    if (n) memcpy(b, p, n);
    b[n] = 0;
    *out = b;
    *n_out = n;
    return TK_OK;
}

// id_to_piece (:617-628) = bounds check + decode(&[id], Keep).  One id needs no kernel: a special
// id yields its string; an ordinary id yields its bytes if they are valid UTF-8 on their own,
// else the Tokenizers error the reference's CoreBPE::decode returns.
extern "C" int tk_id_to_piece(const tk_tokenizer* t, uint32_t id, uint8_t** out, size_t* n) {
    if (!t || !out || !n) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    const tk::HostModel& h = t->host;
    if (id >= h.vocab_size)
        return fail(TK_ERR_INVALID_CONFIG, "Token ID %u is out of vocabulary range (0-%zu)", id, h.vocab_size - 1);
    if (id < h.num_special) {
        const std::string& s = h.special_tokens[id].token_str;
        return give_bytes((const uint8_t*)s.data(), s.size(), out, n);
    }
    size_t r = id - h.num_special;
    if (r >= h.n_vocab()) return fail(TK_ERR_TOKENIZERS, "DecodeKeyError: Invalid token for decoding: %zu", r);
    const uint8_t* p = h.vocab_bytes.data() + h.vocab_off[r];
    size_t len = h.vocab_off[r + 1] - h.vocab_off[r];
    if (!tk::utf8_valid(p, len)) return fail(TK_ERR_TOKENIZERS, "DecodeError: token %u is not valid UTF-8 on its own", id);
    return give_bytes(p, len, out, n);
}

extern "C" int tk_id_to_byte_piece(const tk_tokenizer* t, uint32_t id, int policy, uint8_t** out, size_t* n) {
    if (!t || !out || !n) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    const tk::HostModel& h = t->host;
    if (id >= h.vocab_size)
        return fail(TK_ERR_INVALID_CONFIG, "Token ID %u is out of vocabulary range (0-%zu)", id, h.vocab_size - 1);
    if (id < h.num_special) {
        const std::string& s = h.special_tokens[id].token_str;
        if (policy == TK_POLICY_KEEP) return give_bytes((const uint8_t*)s.data(), s.size(), out, n);
        if (policy == TK_POLICY_RAISE)
            return fail(TK_ERR_SPECIAL_TOKEN_POLICY,
                        "Token ID %u is a special token (%s), cannot convert to byte piece with Raise policy", id, s.c_str());
        return give_bytes(nullptr, 0, out, n);
    }
    size_t r = id - h.num_special;
    if (r < h.n_vocab()) {
        const uint8_t* p = h.vocab_bytes.data() + h.vocab_off[r];
        size_t len = h.vocab_off[r + 1] - h.vocab_off[r];
        if (tk::utf8_valid(p, len)) return give_bytes(p, len, out, n);
    }
    // :683-687 -- on decode failure the reference returns the bytes of the LOSSY vocab string
    const std::string& s = h.vocab_strings[id];
    return give_bytes((const uint8_t*)s.data(), s.size(), out, n);
}

// ------------------------------------------------------------------------------------------ encode

static int check_encode_args(const tk_tokenizer* t, int add_bos, int add_eos) {
    if (!t) return fail(TK_ERR_INVALID_ARGUMENT, "null tokenizer");
    if (t->device < 0)
        return fail(TK_ERR_CUDA, "this tokenizer handle is host-only (device -1); encode/decode need a CUDA device and there is no CPU fallback");
    try {
        if (add_bos) t->host.control_token("<s>");   // bos_id()? (:395)
        if (add_eos) t->host.control_token("</s>");  // eos_id()? (:400)
    } catch (const tk::Error& e) {
        return fail(e.code, "%s", e.what());
    }
    return TK_OK;
}

// Queue the encode kernels for one batch (or one chunk of a batch) on `st`, plus the copy of the
// workspace's counter block to pinned host memory.  Does not wait.  The caller holds t->mu and has
// set the device.
static int encode_issue(tk_tokenizer* t, tk_tokenizer::EncSlot& s, const uint8_t* d_data, const uint64_t* d_doc_off, uint64_t off_base,
                        size_t n_docs, uint64_t total, int add_bos, int add_eos, uint32_t* d_tokens, uint64_t cap,
                        uint64_t* d_tok_off, cudaStream_t st, bool timing) {
    if (((uintptr_t)d_data & 15u) != 0 && total) return fail(TK_ERR_INVALID_ARGUMENT, "device text pointer must be 16-byte aligned");
    // piece lengths, per-class piece counters and queue indices are 32-bit: one device call takes < 4 GiB of text
    // (the host-buffer entry points stream larger batches through in chunks; a single document is limited to this)
    if (total >= (1ull << 32)) return fail(TK_ERR_INVALID_ARGUMENT, "batch too large; shard it (limit 4 GiB of text per device call)");
    if ((uint64_t)n_docs >= 0xFFFFFFFEull) return fail(TK_ERR_INVALID_ARGUMENT, "too many documents in one call; shard the batch");
    size_t ws_bytes = tkk::encode_workspace_bytes(total, n_docs, &s.L);
    CUDA_OR_FAIL(s.ws.ensure(ws_bytes));
    if (s.scratch.cap == 0) CUDA_OR_FAIL(s.scratch.ensure(1 << 20));
    if (!s.h_small) {
        CUDA_OR_FAIL(cudaHostAlloc((void**)&s.h_small, 256, cudaHostAllocMapped));
        CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s.d_small_map, s.h_small, 0));
        CUDA_OR_FAIL(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    if (timing) t->timer.reset();
    cudaError_t e = tkk::encode_device(t->tables, d_data, d_doc_off, off_base, n_docs, total, add_bos, add_eos, d_tokens, cap, d_tok_off,
                                       s.ws.p, s.L, (uint32_t*)s.scratch.p, s.scratch.cap / 4, t->sm_count, st,
                                       timing ? &t->timer : nullptr, &t->hot, t->cfg_ptr());
    if (e != cudaSuccess) return fail(TK_ERR_CUDA, "encode launch: %s", cudaGetErrorString(e));
    CUDA_OR_FAIL(tkk::publish_counters(s.ws.p, s.L, s.d_small_map, st));
    CUDA_OR_FAIL(cudaEventRecord(s.done, st));
    return TK_OK;
}

// Wait for an issued encode and interpret what the kernels reported.  *retry is set when the
// huge-piece scratch was too small: it has been grown and the same encode must be issued again.
static int encode_finish(tk_tokenizer* t, tk_tokenizer::EncSlot& s, cudaStream_t st, uint64_t total, uint64_t cap, uint64_t byte_base,
                         uint64_t* n_tokens, bool* retry, bool timing) {
    *retry = false;
    (void)st;
    CUDA_OR_FAIL(cudaEventSynchronize(s.done));
    if (timing) t->timer.collect(t->stage_names, t->stage_ms);
    const uint32_t* small = s.h_small;
    const uint32_t flags = small[tkk::TKK_S_FLAGS];
    uint64_t err_pos, total_out;
    memcpy(&err_pos, small + tkk::TKK_S_ERRPOS, 8);
    memcpy(&total_out, small + tkk::TKK_S_TOTAL, 8);
    if (flags & tkk::TKK_FLAG_BAD_OFFSETS)
        return fail(TK_ERR_INVALID_ARGUMENT, "document offsets must start at 0, be non-decreasing and end at the text length");
    if (err_pos != ~0ull)
        return fail(TK_ERR_INVALID_UTF8, "input is not valid UTF-8 at byte %llu", (unsigned long long)(err_pos + byte_base));
    if (flags & tkk::TKK_FLAG_SCRATCH_FULL) {
        // pieces longer than TK_MED_MAX bytes are merged in global scratch; the kernel's cursor says how many
        // 32-bit words this batch needs: grow and rerun
        uint64_t need_words;
        memcpy(&need_words, small + tkk::TKK_S_SCRCUR, 8);
        (void)total;
        CUDA_OR_FAIL(s.scratch.ensure((size_t)need_words * 4 + 4096));
        *retry = true;
        return TK_OK;
    }
    *n_tokens = total_out;
    if (flags & tkk::TKK_FLAG_OUT_FULL)
        return fail(TK_ERR_BUFFER_TOO_SMALL, "token buffer holds %llu ids, %llu needed", (unsigned long long)cap,
                    (unsigned long long)total_out);
    return TK_OK;
}

extern "C" int tk_encode_batch_device(const tk_tokenizer* tc, const uint8_t* d_data, const uint64_t* d_doc_off, size_t n_docs,
                                      uint64_t total_bytes, int add_bos, int add_eos, uint32_t* d_tokens,
                                      uint64_t tokens_capacity, uint64_t* d_tok_off, uint64_t* n_tokens, void* stream) {
    int rc = check_encode_args(tc, add_bos, add_eos);
    if (rc) return rc;
    if (!d_doc_off || !d_tok_off || !n_tokens || (!d_data && total_bytes) || (!d_tokens && tokens_capacity))
        return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    tk_tokenizer* t = const_cast<tk_tokenizer*>(tc);
    std::lock_guard<std::mutex> g(t->mu);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device);
    cudaStream_t st = (cudaStream_t)stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = encode_issue(t, t->dev_slot, d_data, d_doc_off, 0, n_docs, total_bytes, add_bos, add_eos, d_tokens, tokens_capacity,
                          d_tok_off, st, t->timing);
        if (rc) return rc;
        bool retry = false;
        rc = encode_finish(t, t->dev_slot, st, total_bytes, tokens_capacity, 0, n_tokens, &retry, t->timing);
        if (rc || !retry) return rc;
    }
    return fail(TK_ERR_CUDA, "huge-piece scratch could not be grown");
}

// ------------------------------------------------------------------------------------------ host-buffer encode engine
//
// Host buffers in, pinned host buffers out, on one GPU or on all GPUs of the box with ONE call.
//
//   plan      the batch is cut at document boundaries into chunks of about kChunkBytes.  A document larger than 1.5
//             chunks is cut INSIDE, at a context-free piece boundary: before an ASCII space that follows an ASCII
//             letter / digit and precedes an ASCII letter -- both patterns start a piece there whatever comes before
//             (the previous piece cannot contain the space, and " x..." matches the word alternative), and no lookahead
//             of an earlier piece reaches past the letter / digit.  The slices are encoded as separate texts; BOS / EOS
//             go to the first / last slice.  So a 1 GiB document pipelines like a batch and needs a chunk's workspace.
//   deal      chunk c goes to device c mod N.  Every device runs the stream pipeline (upload, kernels,
//             download; kSlots buffer slots) over its chunks on its own host thread.
//   stitch    the id count of a chunk is known when its kernels finish; its place in the output when every earlier
//             chunk has reported its count.  Ids are then copied straight to that place in ONE pinned result buffer
//             (sized from the first chunk's ids-per-byte ratio; if the estimate is too small the call is repeated
//             once with the worst-case size), token offsets likewise: no per-shard buffers, no stitching pass.
//   pageable  caller memory that is not page-locked is staged through pinned per-slot buffers by a small pool of
//             copy threads (a direct cudaMemcpyAsync from pageable memory is synchronous and runs at a third of the
//             speed).
// TEKKEN_B200_TRACE=1 prints the device timeline of every chunk on stderr.

namespace {

class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();      // leaked on purpose: worker threads may outlive static destructors
        return *p;
    }
    // dst/src do not overlap.  The caller copies too; returns when all n bytes are in place.
    void parallel_memcpy(void* dst, const void* src, size_t n) {
        constexpr size_t kPiece = 1 << 20;
        if (n <= 2 * kPiece || threads_.empty()) { copy_piece((char*)dst, (const char*)src, n); return; }
        auto job = std::make_shared<Job>();
        job->dst = (char*)dst; job->src = (const char*)src; job->n = n;
        job->pieces = (n + kPiece - 1) / kPiece;
        const size_t helpers = std::min(threads_.size(), job->pieces - 1);
        {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t i = 0; i < helpers; ++i) q_.push_back(job);
        }
        cv_.notify_all();
        run(*job);
        while (job->done.load(std::memory_order_acquire) < job->pieces) std::this_thread::yield();
    }

    // dst[0 .. n) = the ids of a `bits`-bit stream (tkk::pack_ids); src is readable 32 bytes past its end
    void parallel_unpack(uint32_t* dst, const uint8_t* src, size_t n, int bits) {
        if (!n) return;
        auto job = std::make_shared<Job>();
        job->dst = (char*)dst; job->src = (const char*)src; job->n = n; job->bits = bits;
        job->pieces = (n + kIdPiece - 1) / kIdPiece;
        const size_t helpers = std::min(threads_.size(), job->pieces - 1);
        if (helpers) {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t i = 0; i < helpers; ++i) q_.push_back(job);
        }
        if (helpers) cv_.notify_all();
        run(*job);
        while (job->done.load(std::memory_order_acquire) < job->pieces) std::this_thread::yield();
    }

  private:
    static constexpr size_t kIdPiece = 1 << 18;      // ids per piece of an unpack job (a multiple of 16)
    struct Job {
        char* dst; const char* src; size_t n, pieces;
        int bits = 0;                                 // 0: copy n bytes; 18 / 24: widen n ids
        std::atomic<size_t> next{0}, done{0};
    };
    static inline uint32_t unpack_one(const uint8_t* src, size_t i, int bits) {
        const uint64_t bit = (uint64_t)i * (uint64_t)bits;
        uint32_t w;
        memcpy(&w, src + (bit >> 3), 4);
        return (w >> (bit & 7)) & ((1u << bits) - 1u);
    }
#if defined(__x86_64__)
    // ids i0 .. i0 + n of the stream -> dst[0 .. n).  Eight ids per step: two 16-byte loads (4 ids = bits / 2 bytes
    // apart), a byte shuffle that puts the three bytes of every id into its lane, a per-lane shift, a mask, and a
    // non-temporal store (the result is not read again by this thread).
    __attribute__((target("avx2"))) static void unpack_avx2(uint32_t* dst, const uint8_t* src, size_t i0, size_t n, int bits) {
        size_t i = 0;
        while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = unpack_one(src, i0 + i, bits); ++i; }
        const uint64_t bit0 = (uint64_t)(i0 + i) * (uint64_t)bits;
        const uint32_t s0 = (uint32_t)(bit0 & 7);    // bit phase of the first vector id; constant from step to step
        alignas(32) int8_t sh8[32];
        alignas(32) uint32_t sv[8];
        for (int j = 0; j < 4; ++j) {
            const uint32_t bo = s0 + (uint32_t)bits * (uint32_t)j;
            for (int b = 0; b < 3; ++b) sh8[4 * j + b] = sh8[16 + 4 * j + b] = (int8_t)((bo >> 3) + b);
            sh8[4 * j + 3] = sh8[16 + 4 * j + 3] = (int8_t)-1;
            sv[j] = sv[4 + j] = bo & 7;
        }
        const __m256i shuf = _mm256_load_si256((const __m256i*)sh8), shv = _mm256_load_si256((const __m256i*)sv);
        const __m256i mask = _mm256_set1_epi32((int)((1u << bits) - 1u));
        const size_t lane = (size_t)bits / 2;          // bytes per 4 ids
        const uint8_t* p = src + (bit0 >> 3);
        for (; i + 8 <= n; i += 8, p += 2 * lane) {
            const __m128i lo = _mm_loadu_si128((const __m128i*)p), hi = _mm_loadu_si128((const __m128i*)(p + lane));
            __m256i v = _mm256_set_m128i(hi, lo);
            v = _mm256_and_si256(_mm256_srlv_epi32(_mm256_shuffle_epi8(v, shuf), shv), mask);
            _mm256_stream_si256((__m256i*)(dst + i), v);
        }
        _mm_sfence();
        for (; i < n; ++i) dst[i] = unpack_one(src, i0 + i, bits);
    }
#endif
    static void unpack_piece(uint32_t* dst, const uint8_t* src, size_t i0, size_t n, int bits) {
#if defined(__x86_64__)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (avx2) { unpack_avx2(dst, src, i0, n, bits); return; }
#endif
        for (size_t i = 0; i < n; ++i) dst[i] = unpack_one(src, i0 + i, bits);
    }
    // The destination is a pinned staging buffer that the DMA engine reads next and the CPU never reads back: write it
    // with non-temporal stores (no read-for-ownership of the destination lines, no cache pollution).  glibc's memcpy
    // only switches to them far above the 1 MB pieces used here.
#if defined(__x86_64__)
    __attribute__((target("avx2"))) static void copy_nt_avx2(char* dst, const char* src, size_t n) {
        size_t head = (64 - ((uintptr_t)dst & 63)) & 63;
        if (head > n) head = n;
        memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        size_t i = 0;
        for (; i + 128 <= n; i += 128) {
            const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
            const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
            _mm256_stream_si256((__m256i*)(dst + i), a);
            _mm256_stream_si256((__m256i*)(dst + i + 32), b);
            _mm256_stream_si256((__m256i*)(dst + i + 64), c);
            _mm256_stream_si256((__m256i*)(dst + i + 96), d);
        }
        _mm_sfence();
        memcpy(dst + i, src + i, n - i);
    }
#endif
    static void copy_piece(char* dst, const char* src, size_t n) {
#if defined(__x86_64__)
        static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("TEKKEN_B200_NO_NT_COPY");
        if (avx2) { copy_nt_avx2(dst, src, n); return; }
#endif
        memcpy(dst, src, n);
    }
    static void run(Job& j) {
        constexpr size_t kPiece = 1 << 20;
        for (;;) {
            const size_t p = j.next.fetch_add(1);
            if (p >= j.pieces) return;
            if (j.bits) {
                const size_t o = p * kIdPiece, len = std::min(kIdPiece, j.n - o);
                unpack_piece((uint32_t*)j.dst + o, (const uint8_t*)j.src, o, len, j.bits);
            } else {
                const size_t o = p * kPiece, len = std::min(kPiece, j.n - o);
                copy_piece(j.dst + o, j.src + o, len);
            }
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    CopyPool() {
        // half the CPUs, but on a small host (16 CPUs: the 1-GPU boxes of the pool) all but four: the copies are
        // bound by how many cores issue stores, not by the memory system
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned n = std::max(hw / 2, hw > 4 ? hw - 4 : 1u);
        if (const char* e = getenv("TEKKEN_B200_COPY_THREADS")) n = (unsigned)atoi(e);
        n = std::max(1u, std::min(n, 12u));
        for (unsigned i = 0; i + 1 < n; ++i) threads_.emplace_back([this] { loop(); });
        for (auto& t : threads_) t.detach();
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [this] { return !q_.empty(); });
                j = q_.front();
                q_.pop_front();
            }
            run(*j);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> q_;
};

struct Chunk {
    size_t doc_begin, n_docs;        // documents of the chunk; a slice of a big document: that document, n_docs = 1
    uint64_t byte_begin, n_bytes;
    bool partial, first, last;       // slice of one document; its first / last slice
};

inline bool ascii_alpha(uint8_t c) { return (uint8_t)((c | 0x20u) - 'a') < 26u; }
inline bool ascii_alnum(uint8_t c) { return ascii_alpha(c) || (uint8_t)(c - '0') < 10u; }
// a piece starts at byte p whatever the context (see above); p - 1 and p + 1 are inside the document
inline bool context_free_cut(const uint8_t* d, uint64_t p) { return d[p] == ' ' && ascii_alpha(d[p + 1]) && ascii_alnum(d[p - 1]); }

constexpr uint64_t kMaxDeviceCall = (1ull << 32) - (1ull << 20);     // one device call takes < 4 GiB of text

int plan_chunks(const uint8_t* data, const uint64_t* doc_off, size_t n_docs, uint64_t full_chunk, size_t n_devices, std::vector<Chunk>& out) {
    const uint64_t big = full_chunk + full_chunk / 2;
    // The first chunks are small and double up to the full size: the download engine -- the resource that limits a
    // large call (4 bytes come back for every 2.4 that go up) -- starts after a fraction of a millisecond instead of
    // after a full chunk's upload and kernels.
    const uint64_t total = doc_off[n_docs];
    uint64_t chunk_bytes = total > 2 * full_chunk * n_devices ? std::max<uint64_t>(full_chunk / 16, 1u << 20) : full_chunk;
    size_t d = 0;
    while (d < n_docs) {
        if (!out.empty() && out.size() % n_devices == 0) chunk_bytes = std::min(full_chunk, chunk_bytes * 2);
        const uint64_t len_d = doc_off[d + 1] - doc_off[d];
        if (len_d > big) {
            uint64_t pos = doc_off[d];
            const uint64_t end = doc_off[d + 1];
            bool first = true;
            while (pos < end) {
                if (!first && out.size() % n_devices == 0) chunk_bytes = std::min(full_chunk, chunk_bytes * 2);   // slices grow like chunks do
                uint64_t cut = end;
                if (end - pos > big) {
                    const uint64_t target = pos + chunk_bytes, lo = pos + chunk_bytes / 2;
                    cut = 0;
                    for (uint64_t p = target; p > lo; --p) if (context_free_cut(data, p)) { cut = p; break; }
                    if (!cut) {
                        const uint64_t hi = std::min(end - 1, pos + kMaxDeviceCall);
                        for (uint64_t p = target + 1; p < hi; ++p) if (context_free_cut(data, p)) { cut = p; break; }
                    }
                    if (!cut) cut = end;      // no such boundary (one endless run): the rest goes as one piece of work
                }
                if (cut - pos >= kMaxDeviceCall)
                    return fail(TK_ERR_INVALID_ARGUMENT, "document %zu has no cut point within 4 GiB (a single run of that length); it cannot be encoded in one device call", d);
                out.push_back({d, 1, pos, cut - pos, true, first, cut == end});
                first = false;
                pos = cut;
            }
            ++d;
            continue;
        }
        const size_t b = d;
        const uint64_t begin = doc_off[b];
        size_t e = std::upper_bound(doc_off + b + 1, doc_off + n_docs + 1, begin + chunk_bytes) - doc_off;   // first doc END beyond the target
        if (e > n_docs) e = n_docs;
        if (e <= b) e = b + 1;
        // a big document is never part of an ordinary chunk (it is sliced above)
        for (size_t k = (e - b > 4 ? e - 2 : b); k < e; ++k)
            if (doc_off[k + 1] - doc_off[k] > big) { e = k > b ? k : b + 1; break; }
        out.push_back({b, e - b, begin, doc_off[e] - begin, false, true, true});
        d = e;
    }
    if (out.empty()) out.push_back({0, 0, 0, 0, false, true, true});
    return TK_OK;
}

struct EncodeJob {
    const uint8_t* data = nullptr;
    const uint64_t* doc_off = nullptr;
    size_t n_docs = 0;
    uint64_t total = 0;
    int add_bos = 0, add_eos = 0;
    bool pageable = false;
    std::vector<Chunk> chunks;
    uint64_t forced_cap = 0;          // second attempt: worst-case output size
    float ratio_hint = 0.f;           // ids per byte recent calls on this handle needed
    int pack_bits = 0;                // 18 / 24: the ids cross PCIe as a bit stream and are widened on the host; 0: as they are
    // progress shared by the device workers
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int64_t> ntok;        // id count per chunk, -1 = not known yet
    std::vector<uint64_t> prefix;     // prefix[c] = ids before chunk c, valid for c <= known
    size_t known = 0;
    uint32_t* h_tok = nullptr;
    uint64_t h_cap = 0;
    uint64_t* h_off = nullptr;
    bool alloc_failed = false;
    std::atomic<bool> abort{false};
    bool overflow = false;            // the estimate was too small: repeat with the worst-case size
    int rc = TK_OK;
    std::string err;

    void fail_with(int code, const std::string& msg) {
        std::lock_guard<std::mutex> g(mu);
        if (rc == TK_OK) { rc = code; err = msg; }
        abort.store(true);
        cv.notify_all();
    }
};

const uint64_t kDefaultChunkBytes = [] {
    // tuning knob; default 128 MB: a chunk costs about 0.8 ms of launches, ramps and tails on top of its bytes, and the
    // kernels of a chunk must outrun the download of the previous chunk's ids (measured: 48 MB chunks run at 31 GB/s
    // of text, the ids leave at the equivalent of 27-33 GB/s; 128 MB chunks run at 45 GB/s)
    const char* e = getenv("TEKKEN_B200_CHUNK_MB");
    const long mb = e ? atol(e) : 0;
    return (uint64_t)(mb > 0 ? mb : 128) << 20;
}();
std::atomic<uint64_t> g_chunk_bytes{0};                     // tk_set_chunk_bytes; 0 = the default
std::atomic<int> g_pack_mode{[] { const char* e = getenv("TEKKEN_B200_PACK_IDS"); return e ? atoi(e) : -1; }()};   // tk_set_pack_ids

}  // namespace

// One device's share of an encode job: chunks g, g + stride, ...  Runs on its own host thread (the caller's for
// device 0).  Returns TK_OK or the status it also stored in the job.
static int encode_worker(tk_tokenizer* t, EncodeJob& J, size_t g, size_t stride) {
    std::lock_guard<std::mutex> lock(t->mu);
    DeviceGuard dg(t->device);
    auto bail = [&](int code) {
        J.fail_with(code, g_last_error);
        for (auto& ps : t->pipe_st) if (ps) cudaStreamSynchronize(ps);
        return code;
    };
    if (!dg.ok) return bail(fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device));
#define W_CUDA(x)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) return bail(fail(TK_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)));         \
    } while (0)
    for (auto& ps : t->pipe_st) if (!ps) W_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    cudaStream_t st_up = t->pipe_st[0], st_k = t->pipe_st[1], st_down = t->pipe_st[2], st_pack = t->pipe_st[3];
    std::vector<size_t> mine;
    for (size_t c = g; c < J.chunks.size(); c += stride) mine.push_back(c);
    const size_t n_mine = mine.size();

    static const bool kTrace = getenv("TEKKEN_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<double> thost;
    const auto host_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    if (kTrace) {
        tev.resize(n_mine * 6 + 1);
        thost.resize(n_mine * 3 + 1);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], st_up);
        thost[0] = host_now();
    }
    struct TraceGuard {
        std::vector<cudaEvent_t>& ev;
        ~TraceGuard() { for (auto& e : ev) cudaEventDestroy(e); }
    } trace_guard{tev};
    auto mark = [&](size_t i, int k, cudaStream_t st) { if (kTrace) cudaEventRecord(tev[1 + i * 6 + k], st); };

    auto caps = [&](const Chunk& c) { return c.n_bytes + 2 * (uint64_t)c.n_docs + 2; };
    auto bos_of = [&](const Chunk& c) { return J.add_bos && c.first ? 1 : 0; };
    auto eos_of = [&](const Chunk& c) { return J.add_eos && c.last ? 1 : 0; };
    auto launch = [&](size_t i) -> int {      // the kernels of my i-th chunk (its text is on the device or on its way)
        tk_tokenizer::EncSlot& s = t->slot[i % tk_tokenizer::kSlots];
        tk_tokenizer::OutSlot& o = t->oslot[i % tk_tokenizer::kOutSlots];
        const Chunk& c = J.chunks[mine[i]];
        return encode_issue(t, s, (const uint8_t*)s.in_data.p, (const uint64_t*)s.in_off.p, c.partial ? 0 : c.byte_begin, c.n_docs, c.n_bytes,
                            bos_of(c), eos_of(c), (uint32_t*)o.tok.p, caps(c), (uint64_t*)o.off.p, st_k, false);
    };
    // upload(i): the text and offsets of my i-th chunk on their way to the device (staged through pinned memory when
    // the caller's memory is pageable, 16 MB at a time so that the copy of one piece overlaps the staging of the next).
    // kick(i): its kernels queued behind the upload.  With pageable input a stager thread runs the uploads ahead of
    // this thread, which then only queues kernels, waits for results and starts downloads.
    std::vector<std::atomic<int>> up_state(n_mine), kicked(n_mine);
    for (size_t i = 0; i < n_mine; ++i) { up_state[i].store(0); kicked[i].store(0); }
    auto upload = [&](size_t i) -> int {
        tk_tokenizer::EncSlot& s = t->slot[i % tk_tokenizer::kSlots];
        const Chunk& c = J.chunks[mine[i]];
        if (!s.ev_in) CUDA_OR_FAIL(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        if (kTrace) thost[1 + i * 3] = host_now();
        const size_t off_bytes = (c.n_docs + 1) * 8;
        CUDA_OR_FAIL(s.in_data.ensure(c.n_bytes + 64));
        CUDA_OR_FAIL(s.in_off.ensure(off_bytes));
        const uint8_t* src = J.data + c.byte_begin;
        const uint64_t* src_off = J.doc_off + c.doc_begin;
        uint8_t* sd = nullptr;
        if (J.pageable || c.partial) {
            // pinned staging: [offsets | text]; the previous upload from it has long finished (kSlots chunks ago), but say so
            const size_t need = ((off_bytes + 63) & ~(size_t)63) + (J.pageable ? c.n_bytes : 0) + 64;
            if (s.stage_cap < need) {
                if (s.h_stage) { cudaStreamSynchronize(st_up); cudaFreeHost(s.h_stage); }
                s.h_stage = nullptr; s.stage_cap = 0;
                CUDA_OR_FAIL(cudaHostAlloc((void**)&s.h_stage, need + need / 8, cudaHostAllocDefault));
                s.stage_cap = need + need / 8;
            } else if (i >= (size_t)tk_tokenizer::kSlots) CUDA_OR_FAIL(cudaEventSynchronize(s.ev_in));
            uint64_t* so = (uint64_t*)s.h_stage;
            if (c.partial) { so[0] = 0; so[1] = c.n_bytes; }
            else memcpy(so, src_off, off_bytes);
            src_off = so;
            if (J.pageable) sd = s.h_stage + ((off_bytes + 63) & ~(size_t)63);
        }
        if (i >= (size_t)tk_tokenizer::kSlots) CUDA_OR_FAIL(cudaStreamWaitEvent(st_up, s.done, 0));
        mark(i, 0, st_up);
        if (sd) {
            constexpr uint64_t kPiece = 16u << 20;
            for (uint64_t o = 0; o < c.n_bytes; o += kPiece) {
                const uint64_t len = std::min<uint64_t>(kPiece, c.n_bytes - o);
                CopyPool::get().parallel_memcpy(sd + o, src + o, len);
                CUDA_OR_FAIL(cudaMemcpyAsync((uint8_t*)s.in_data.p + o, sd + o, len, cudaMemcpyHostToDevice, st_up));
            }
        } else if (c.n_bytes) CUDA_OR_FAIL(cudaMemcpyAsync(s.in_data.p, src, c.n_bytes, cudaMemcpyHostToDevice, st_up));
        CUDA_OR_FAIL(cudaMemcpyAsync(s.in_off.p, src_off, off_bytes, cudaMemcpyHostToDevice, st_up));
        mark(i, 1, st_up);
        CUDA_OR_FAIL(cudaEventRecord(s.ev_in, st_up));
        return TK_OK;
    };
    auto kick = [&](size_t i) -> int {
        tk_tokenizer::EncSlot& s = t->slot[i % tk_tokenizer::kSlots];
        tk_tokenizer::OutSlot& o = t->oslot[i % tk_tokenizer::kOutSlots];
        const Chunk& c = J.chunks[mine[i]];
        const size_t off_bytes = (c.n_docs + 1) * 8;
        if (!o.ev_out) CUDA_OR_FAIL(cudaEventCreateWithFlags(&o.ev_out, cudaEventDisableTiming));
        if (i >= (size_t)tk_tokenizer::kOutSlots && (o.tok.cap < caps(c) * 4 || o.off.cap < off_bytes))
            CUDA_OR_FAIL(cudaEventSynchronize(o.ev_out));      // about to reallocate a buffer whose ids may still be leaving
        CUDA_OR_FAIL(o.tok.ensure(caps(c) * 4));
        CUDA_OR_FAIL(o.off.ensure(off_bytes));
        CUDA_OR_FAIL(cudaStreamWaitEvent(st_k, s.ev_in, 0));
        if (i >= (size_t)tk_tokenizer::kOutSlots) CUDA_OR_FAIL(cudaStreamWaitEvent(st_k, o.ev_out, 0));
        mark(i, 2, st_k);
        const int r = launch(i);
        mark(i, 3, st_k);
        if (kTrace) thost[2 + i * 3] = host_now();
        kicked[i].store(1, std::memory_order_release);
        return r;
    };
    // the stager (pageable input only): uploads in order, at most kSlots chunks ahead of the kernels
    std::thread stager;
    struct StagerJoin {
        std::thread& th; EncodeJob& J; bool failed_flag = false;
        ~StagerJoin() { if (th.joinable()) { if (std::uncaught_exceptions() || failed_flag) J.abort.store(true); th.join(); } }
    } stager_join{stager, J};
    const bool use_stager = J.pageable && n_mine > 1;
    if (use_stager) {
        stager = std::thread([&] {
            DeviceGuard g2(t->device);
            for (size_t i = 0; i < n_mine; ++i) {
                if (i >= (size_t)tk_tokenizer::kSlots)      // the slot's previous chunk must be through its kernels (a re-run after
                                                            // growing the long-piece scratch reads the slot's text again)
                    while (kicked[i - tk_tokenizer::kSlots].load(std::memory_order_acquire) < 2) {
                        if (J.abort.load()) { up_state[i].store(-1); return; }
                        std::this_thread::yield();
                    }
                if (J.abort.load()) { up_state[i].store(-1); return; }
                const int r = upload(i);
                if (r) { J.fail_with(r, g_last_error); for (size_t k = i; k < n_mine; ++k) up_state[k].store(-1); return; }
                up_state[i].store(1, std::memory_order_release);
            }
        });
    }
    auto issue = [&](size_t i) -> int {
        if (use_stager) {
            int st;
            while ((st = up_state[i].load(std::memory_order_acquire)) == 0) {
                if (J.abort.load()) return J.rc ? J.rc : TK_ERR_CUDA;
                std::this_thread::yield();
            }
            if (st < 0) return fail(J.rc ? J.rc : TK_ERR_CUDA, "%s", J.err.c_str());
        } else {
            const int r = upload(i);
            if (r) return r;
        }
        return kick(i);
    };

    // Packed ids (large calls): the download lands in a pinned stage and this thread widens it into the result while
    // the next chunk's ids are on the bus.  1.64 GB of ids per GB of text become 0.92 GB (18 bits) on the link that
    // bounds the call.
    struct UnpackTask { int stage; uint32_t* dst; uint64_t n; };
    std::deque<UnpackTask> uq;
    std::mutex umu;
    std::condition_variable ucv;
    bool u_done = false;
    std::thread unpacker;
    struct UnpackJoin {
        std::thread& th; std::mutex& mu; std::condition_variable& cv; bool& done;
        void finish() {
            if (!th.joinable()) return;
            { std::lock_guard<std::mutex> g(mu); done = true; }
            cv.notify_all();
            th.join();
        }
        ~UnpackJoin() { finish(); }
    } unpack_join{unpacker, umu, ucv, u_done};
    const int pack_bits = J.pack_bits;
    size_t n_packed = 0;
    if (pack_bits) {
        unpacker = std::thread([&] {
            DeviceGuard g2(t->device);
            for (;;) {
                UnpackTask task;
                {
                    std::unique_lock<std::mutex> lk(umu);
                    ucv.wait(lk, [&] { return u_done || !uq.empty(); });
                    if (uq.empty()) return;
                    task = uq.front();
                    uq.pop_front();
                }
                tk_tokenizer::PackStage& ps = t->pstage[task.stage];
                const cudaError_t e = cudaEventSynchronize(ps.ev);
                if (e != cudaSuccess) J.fail_with(TK_ERR_CUDA, std::string("copying ids back: ") + cudaGetErrorString(e));
                else if (!J.abort.load()) CopyPool::get().parallel_unpack(task.dst, ps.h, task.n, pack_bits);
                ps.busy.store(0, std::memory_order_release);
            }
        });
    }

    // kAhead chunks are always queued ahead, so this thread sits in the wait for chunk i when it completes and its
    // ids start their way back at once
    constexpr size_t kAhead = tk_tokenizer::kSlots - 1;
    int rc = TK_OK;
    for (size_t i = 0; i < kAhead && i < n_mine; ++i) { rc = issue(i); if (rc) return bail(rc); }
    for (size_t i = 0; i < n_mine; ++i) {
        if (J.abort.load()) return bail(J.rc ? J.rc : TK_ERR_CUDA);
        tk_tokenizer::EncSlot& s = t->slot[i % tk_tokenizer::kSlots];
        const size_t ci = mine[i];
        const Chunk& c = J.chunks[ci];
        uint64_t n_tok = 0;
        for (int attempt = 0;; ++attempt) {
            bool retry = false;
            rc = encode_finish(t, s, st_k, c.n_bytes, caps(c), c.byte_begin, &n_tok, &retry, false);
            if (rc) return bail(rc);
            if (!retry) break;
            if (attempt) return bail(fail(TK_ERR_CUDA, "huge-piece scratch could not be grown"));
            rc = launch(i);
            if (rc) return bail(rc);
        }
        kicked[i].store(2, std::memory_order_release);       // the slot's text buffers may be refilled
        uint64_t my_prefix = 0;
        {
            std::unique_lock<std::mutex> lk(J.mu);
            J.ntok[ci] = (int64_t)n_tok;
            while (J.known < J.chunks.size() && J.ntok[J.known] >= 0) { J.prefix[J.known + 1] = J.prefix[J.known] + (uint64_t)J.ntok[J.known]; ++J.known; }
            if (ci == 0 && !J.h_tok) {
                // the result buffer, sized from the first chunk's ids per byte (the worst case on the second attempt)
                const uint64_t worst = J.total + 2 * (uint64_t)J.n_docs + 2;
                uint64_t cap = worst;
                if (!J.forced_cap && J.chunks.size() > 1) {
                    // the first chunk's ids per byte, or what recent calls on this handle needed if that was more (a
                    // document whose first megabytes are not typical of the rest would otherwise overflow every time)
                    const double ratio = std::max((double)n_tok / (double)std::max<uint64_t>(1, c.n_bytes), (double)J.ratio_hint);
                    cap = std::min(worst, (uint64_t)(ratio * 1.25 * (double)J.total) + 2 * (uint64_t)J.n_docs + (1u << 16));
                }
                if (J.chunks.size() == 1) cap = n_tok + 2;
                lk.unlock();
                uint32_t* p = (uint32_t*)g_pool.get(std::max<uint64_t>(cap, 1) * 4);
                lk.lock();
                if (!p) J.alloc_failed = true;
                J.h_tok = p;
                J.h_cap = cap;
            }
            J.cv.notify_all();
            J.cv.wait(lk, [&] { return J.abort.load() || ((J.known >= ci) && (J.h_tok || J.alloc_failed)); });
            if (J.alloc_failed) { lk.unlock(); return bail(fail(TK_ERR_CUDA, "out of pinned host memory")); }
            if (J.abort.load()) { lk.unlock(); return bail(J.rc ? J.rc : TK_ERR_CUDA); }
            my_prefix = J.prefix[ci];
            if (my_prefix + n_tok > J.h_cap) {
                J.overflow = true;
                lk.unlock();
                J.fail_with(TK_ERR_BUFFER_TOO_SMALL, "result estimate too small");
                for (auto& ps : t->pipe_st) if (ps) cudaStreamSynchronize(ps);
                return TK_ERR_BUFFER_TOO_SMALL;
            }
        }
        cudaError_t e = cudaSuccess;
        if (kTrace) thost[3 + i * 3] = host_now();
        mark(i, 4, st_down);
        tk_tokenizer::OutSlot& o = t->oslot[i % tk_tokenizer::kOutSlots];
        // a landing buffer that has been widened already; if the host threads are behind (all four still busy: a small
        // or crowded host) this chunk's ids travel as they are -- the link never waits for the CPU
        // ... and the last chunk always does: nothing is left on the bus to hide its widening behind, so the packed
        // trip (fewer bytes, then a host pass) would end later than the plain one
        int stage = -1;
        if (n_tok && pack_bits && (i + 1 < n_mine || n_mine == 1)) {
            int n_busy = 0;
            for (int k = 0; k < tk_tokenizer::kPackStages; ++k) {
                if (t->pstage[k].busy.load(std::memory_order_acquire)) ++n_busy;
                else if (stage < 0) stage = k;
            }
            if (n_busy >= 2) stage = -1;                     // two chunks behind already
        }
        if (stage >= 0) {
            tk_tokenizer::PackStage& ps = t->pstage[stage];
            const size_t pb = tkk::packed_id_bytes(n_tok, pack_bits);
            // all landing buffers get the size of the largest chunk seen (page-locked allocation is slow: ~70 ms for
            // 200 MB; sizes must settle after the first call, whichever buffer meets whichever chunk)
            t->pack_stage_want = std::max(t->pack_stage_want, pb + pb / 4 + 4096);
            if (ps.cap < pb + 64) {
                if (ps.h) cudaFreeHost(ps.h);
                ps.h = nullptr; ps.cap = 0;
                W_CUDA(cudaHostAlloc((void**)&ps.h, t->pack_stage_want, cudaHostAllocPortable));
                ps.cap = t->pack_stage_want;
            }
            if (!ps.ev) { W_CUDA(cudaEventCreateWithFlags(&ps.ev, cudaEventDisableTiming)); W_CUDA(cudaEventCreateWithFlags(&ps.ev_packed, cudaEventDisableTiming)); }
            // the slot's pack buffer was last read by the download kOutSlots chunks ago (o.ev_out); packing on its own
            // stream overlaps the previous chunk's download
            if (i >= (size_t)tk_tokenizer::kOutSlots) W_CUDA(cudaStreamWaitEvent(st_pack, o.ev_out, 0));
            W_CUDA(o.pack.ensure(pb));
            W_CUDA(tkk::pack_ids((const uint32_t*)o.tok.p, n_tok, pack_bits, o.pack.p, st_pack));
            W_CUDA(cudaEventRecord(ps.ev_packed, st_pack));
            W_CUDA(cudaStreamWaitEvent(st_down, ps.ev_packed, 0));
            e = cudaMemcpyAsync(ps.h, o.pack.p, pb, cudaMemcpyDeviceToHost, st_down);
            if (e == cudaSuccess) e = cudaEventRecord(ps.ev, st_down);
            if (e == cudaSuccess) {
                ps.busy.store(1, std::memory_order_release);
                { std::lock_guard<std::mutex> g2(umu); uq.push_back({stage, J.h_tok + my_prefix, n_tok}); }
                ++n_packed;
                ucv.notify_all();
            }
        } else if (n_tok) e = cudaMemcpyAsync(J.h_tok + my_prefix, o.tok.p, n_tok * 4, cudaMemcpyDeviceToHost, st_down);
        if (e == cudaSuccess && !c.partial && c.n_docs)
            e = cudaMemcpyAsync(J.h_off + c.doc_begin, o.off.p, c.n_docs * 8, cudaMemcpyDeviceToHost, st_down);
        mark(i, 5, st_down);
        if (e == cudaSuccess) e = cudaEventRecord(o.ev_out, st_down);
        if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "copying ids back: %s", cudaGetErrorString(e)));
        if (i + kAhead < n_mine) { rc = issue(i + kAhead); if (rc) return bail(rc); }
    }
    {
        cudaError_t e = cudaStreamSynchronize(st_down);
        if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "copying ids back: %s", cudaGetErrorString(e)));
        unpack_join.finish();                                  // the last chunks' ids are in the result
        if (J.abort.load()) return bail(J.rc ? J.rc : TK_ERR_CUDA);
    }
    // chunk-local token offsets -> batch offsets (every prefix is known by now: my last chunk waited for them)
    for (size_t i = 0; i < n_mine; ++i) {
        const Chunk& c = J.chunks[mine[i]];
        const uint64_t add = J.prefix[mine[i]];
        if (c.partial) { if (c.first) J.h_off[c.doc_begin] = add; }
        else if (add) for (size_t d = c.doc_begin; d < c.doc_begin + c.n_docs; ++d) J.h_off[d] += add;
    }
    if (kTrace) {
        fprintf(stderr, "[tekken_b200 trace] device %d chunk: host issue..issued | h2d begin..end | kernels begin..end | host saw done | d2h begin..end  (ms)\n", t->device);
        for (size_t i = 0; i < n_mine; ++i) {
            float v[6];
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&v[k], tev[0], tev[1 + i * 6 + k]);
            fprintf(stderr, "[tekken_b200 trace] %d/%3zu: %7.2f..%7.2f | %7.2f..%7.2f | %7.2f..%7.2f | %7.2f | %7.2f..%7.2f\n", t->device, mine[i],
                    thost[1 + i * 3] - thost[0], thost[2 + i * 3] - thost[0], v[0], v[1], v[2], v[3], thost[3 + i * 3] - thost[0], v[4], v[5]);
        }
        fprintf(stderr, "[tekken_b200 trace] device %d: all copies done at host %.2f ms; %zu of %zu chunks came back as %d-bit ids\n", t->device,
                host_now() - thost[0], n_packed, n_mine, pack_bits);
    }
#undef W_CUDA
    return TK_OK;
}

static int encode_batch_engine(tk_tokenizer* const* handles, size_t n_handles, const uint8_t* data, const uint64_t* doc_off, size_t n_docs,
                               int add_bos, int add_eos, uint32_t** tokens, uint64_t** tok_off) {
    if (!doc_off || !tokens || !tok_off) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    *tokens = nullptr;
    *tok_off = nullptr;
    {
        // the chunk plan slices the caller's buffers by these offsets: check them before anything is copied
        uint64_t bad = doc_off[0] != 0;
        for (size_t d = 0; d < n_docs; ++d) bad |= (uint64_t)(doc_off[d + 1] < doc_off[d]);
        if (bad) return fail(TK_ERR_INVALID_ARGUMENT, "document offsets must start at 0, be non-decreasing and end at the text length");
    }
    const uint64_t total = doc_off[n_docs];
    if (!data && total) return fail(TK_ERR_INVALID_ARGUMENT, "null text");
    bool pageable = false;
    if (total) {
        DeviceGuard dg(handles[0]->device);
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, data) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = pa.type == cudaMemoryTypeUnregistered;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        EncodeJob J;
        J.data = data; J.doc_off = doc_off; J.n_docs = n_docs; J.total = total; J.add_bos = add_bos; J.add_eos = add_eos;
        J.pageable = pageable && total >= (1u << 16);       // small inputs: the driver's own staging is as good
        J.forced_cap = attempt ? total + 2 * (uint64_t)n_docs + 2 : 0;
        J.ratio_hint = handles[0]->ratio_hint.load();
        {
            // tk_set_pack_ids / TEKKEN_B200_PACK_IDS: 0 = never, 18 / 24 = that width whenever the ids fit, -1 (default) =
            // the narrowest width that fits, for calls of 32 MB and more (below that the extra kernel and host pass cost
            // more than the bytes they save)
            const int mode = g_pack_mode.load();
            const uint64_t n_ids_max = (uint64_t)handles[0]->tables.n_vocab + handles[0]->tables.num_special;
            int bits = n_ids_max <= (1u << 18) ? 18 : n_ids_max <= (1u << 24) ? 24 : 0;
            if (mode == 24 && bits) bits = 24;
            if (mode == 0 || (mode < 0 && total < (32u << 20))) bits = 0;
            if (mode < 0 && bits) {
                // the widening needs CPUs: with one process per GPU (torchrun sets LOCAL_WORLD_SIZE) and fewer than twelve
                // CPUs for each (the copy pool's size), or with several devices driven by this call, leave the ids as
                // they are
                const char* lws = getenv("LOCAL_WORLD_SIZE");
                const unsigned procs = lws && atoi(lws) > 0 ? (unsigned)atoi(lws) : 1u;
                if (std::thread::hardware_concurrency() / (procs * (unsigned)n_handles) < 12u) bits = 0;
            }
            J.pack_bits = bits;
        }
        // chunks small enough that every device gets several, large enough to keep the launch overhead low
        uint64_t chunk = g_chunk_bytes.load() ? g_chunk_bytes.load() : kDefaultChunkBytes;
        if (n_handles > 1) chunk = std::max<uint64_t>(4u << 20, std::min<uint64_t>(chunk, total / (n_handles * 4) + 1));
        int rc = plan_chunks(data, doc_off, n_docs, chunk, n_handles, J.chunks);
        if (rc) return rc;
        // ... and only a pipeline of several chunks per device hides the widening behind later transfers
        if (g_pack_mode.load() < 0 && J.chunks.size() < 4 * n_handles) J.pack_bits = 0;
        J.ntok.assign(J.chunks.size(), -1);
        J.prefix.assign(J.chunks.size() + 1, 0);
        J.h_off = (uint64_t*)g_pool.get((n_docs + 1) * 8);
        if (!J.h_off) return fail(TK_ERR_CUDA, "out of pinned host memory");
        const size_t n_dev = std::min(n_handles, J.chunks.size());
        std::vector<std::thread> th;
        for (size_t g = 1; g < n_dev; ++g) th.emplace_back([&, g] { encode_worker(handles[g], J, g, n_dev); });
        encode_worker(handles[0], J, 0, n_dev);
        for (auto& x : th) x.join();
        if (J.rc == TK_OK) {
            if (total >= (64u << 20)) {
                const float r = (float)((double)J.prefix[J.chunks.size()] / (double)total), old = handles[0]->ratio_hint.load();
                handles[0]->ratio_hint.store(std::max(r, old * 0.9f));
            }
            J.h_off[n_docs] = J.prefix[J.chunks.size()];
            if (!J.h_tok) J.h_tok = (uint32_t*)g_pool.get(4);      // (not reached: chunk 0 always allocates)
            *tokens = J.h_tok;
            *tok_off = J.h_off;
            return TK_OK;
        }
        g_pool.put(J.h_tok);
        g_pool.put(J.h_off);
        if (!(J.overflow && attempt == 0)) return fail(J.rc, "%s", J.err.c_str());
    }
    return fail(TK_ERR_CUDA, "result buffer could not be sized");
}

extern "C" int tk_encode_batch(const tk_tokenizer* tc, const uint8_t* data, const uint64_t* doc_off, size_t n_docs, int add_bos,
                               int add_eos, uint32_t** tokens, uint64_t** tok_off) {
    int rc = check_encode_args(tc, add_bos, add_eos);
    if (rc) return rc;
    tk_tokenizer* t = const_cast<tk_tokenizer*>(tc);
    return encode_batch_engine(&t, 1, data, doc_off, n_docs, add_bos, add_eos, tokens, tok_off);
}

extern "C" int tk_encode_batch_multi(tk_tokenizer* const* handles, size_t n_handles, const uint8_t* data, const uint64_t* doc_off,
                                     size_t n_docs, int add_bos, int add_eos, uint32_t** tokens, uint64_t** tok_off) {
    if (!handles || n_handles == 0) return fail(TK_ERR_INVALID_ARGUMENT, "no tokenizer handles");
    for (size_t g = 0; g < n_handles; ++g) {
        int rc = check_encode_args(handles[g], add_bos, add_eos);
        if (rc) return rc;
        const tk_tokenizer* a = handles[0];
        const tk_tokenizer* b = handles[g];
        if (b->host.vocab_size != a->host.vocab_size || b->host.num_special != a->host.num_special || b->split_mode != a->split_mode ||
            b->host.vocab_bytes != a->host.vocab_bytes)
            return fail(TK_ERR_INVALID_ARGUMENT, "handle %zu was not created from the same tokenizer as handle 0", g);
        for (size_t k = 0; k < g; ++k)      // (two handles on one device work -- separate streams and workspaces -- but gain nothing)
            if (handles[k] == handles[g]) return fail(TK_ERR_INVALID_ARGUMENT, "handle %zu is passed twice; pass one handle per GPU", g);
    }
    return encode_batch_engine(handles, n_handles, data, doc_off, n_docs, add_bos, add_eos, tokens, tok_off);
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

// The latency path: a text of at most kSmallMaxBytes goes through ONE single-block kernel that reads it from mapped
// pinned memory and writes the ids + a completion word back to mapped pinned memory, which this thread polls.
// *fallback is set when the text has a piece only the batch path's long-piece kernels handle.
static int encode_small_text(tk_tokenizer* t, const uint8_t* utf8, size_t len, int add_bos, int add_eos, uint32_t** out, size_t* n_out,
                             bool* fallback) {
    *fallback = false;
    tk_tokenizer::FastSlot* s = nullptr;
    for (auto& f : t->fast) if (f.mu.try_lock()) { s = &f; break; }
    if (!s) { s = &t->fast[0]; s->mu.lock(); }
    std::lock_guard<std::mutex> g(s->mu, std::adopt_lock);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device);
    if (!s->st) {
        CUDA_OR_FAIL(cudaHostAlloc((void**)&s->h_in, tkk::kSmallMaxBytes + 128, cudaHostAllocMapped));
        CUDA_OR_FAIL(cudaHostAlloc((void**)&s->h_out, tkk::kSmallOutWords * 4, cudaHostAllocMapped));
        CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s->d_in, s->h_in, 0));
        CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s->d_out, s->h_out, 0));
        memset(s->h_out, 0, 64);
        CUDA_OR_FAIL(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    }
    if (len) memcpy(s->h_in, utf8, len);
    const uint32_t seq = ++s->seq ? s->seq : ++s->seq;            // never 0: the cleared header reads as "not done"
    cudaError_t e = tkk::encode_small(t->tables, s->d_in, (uint32_t)len, add_bos, add_eos, s->d_out, seq, s->st);
    if (e != cudaSuccess) return fail(TK_ERR_CUDA, "encode launch: %s", cudaGetErrorString(e));
    volatile uint32_t* hdr = s->h_out;
    for (uint32_t spin = 1;; ++spin) {
        if (hdr[4] == seq) break;
        if ((spin & 4095u) == 0u) {
            // the kernel cannot finish without setting the word: a finished (or failed) stream without it is an error
            const cudaError_t q = cudaStreamQuery(s->st);
            if (q == cudaSuccess) { if (hdr[4] == seq) break; return fail(TK_ERR_CUDA, "single-block encode kernel ended without a result"); }
            if (q != cudaErrorNotReady) return fail(TK_ERR_CUDA, "single-block encode kernel: %s", cudaGetErrorString(q));
        }
        cpu_relax();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const uint32_t n_ids = hdr[0], flags = hdr[1], err_pos = hdr[2];
    if (flags & tkk::kSmallBadUtf8) return fail(TK_ERR_INVALID_UTF8, "input is not valid UTF-8 at byte %u", err_pos);
    if (flags & tkk::kSmallNeedBatch) { *fallback = true; return TK_OK; }
    uint32_t* ids = (uint32_t*)g_pool.get((size_t)n_ids * 4 + 4, false);
    if (!ids) return fail(TK_ERR_CUDA, "out of host memory");
    if (n_ids) memcpy(ids, s->h_out + 8, (size_t)n_ids * 4);
    *out = ids;
    *n_out = n_ids;
    return TK_OK;
}

extern "C" int tk_encode(const tk_tokenizer* t, const uint8_t* utf8, size_t len, int add_bos, int add_eos, uint32_t** out,
                         size_t* n_out) {
    if (!out || !n_out) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    static const bool kNoFast = getenv("TEKKEN_B200_NO_FAST") != nullptr;      // measurements: force the batch path
    if (t && t->device >= 0 && len <= tkk::kSmallMaxBytes && t->split_mode == TK_SPLIT_REFERENCE && !kNoFast) {
        int rc = check_encode_args(t, add_bos, add_eos);
        if (rc) return rc;
        if (!utf8 && len) return fail(TK_ERR_INVALID_ARGUMENT, "null text");
        bool fallback = false;
        rc = encode_small_text(const_cast<tk_tokenizer*>(t), utf8, len, add_bos, add_eos, out, n_out, &fallback);
        if (rc || !fallback) return rc;
    }
    uint64_t off[2] = {0, len};
    uint64_t* tok_off = nullptr;
    int rc = tk_encode_batch(t, utf8, off, 1, add_bos, add_eos, out, &tok_off);
    if (rc) return rc;
    *n_out = (size_t)tok_off[1];
    tk_buffer_free(tok_off);
    return TK_OK;
}

// ------------------------------------------------------------------------------------------ decode

static int run_decode(tk_tokenizer* t, const uint32_t* d_ids, const uint64_t* d_tok_off, size_t n_docs, uint64_t n_ids, int policy,
                      uint8_t* d_out, uint64_t cap, uint64_t* d_byte_off, int32_t* d_status, uint64_t* n_bytes, uint64_t* bad_doc,
                      cudaStream_t st) {
    if (policy != TK_POLICY_IGNORE && policy != TK_POLICY_KEEP && policy != TK_POLICY_RAISE)
        return fail(TK_ERR_INVALID_ARGUMENT, "unknown special token policy %d", policy);
    if (((uintptr_t)d_ids & 3u) != 0) return fail(TK_ERR_INVALID_ARGUMENT, "id pointer must be 4-byte aligned");
    if ((uint64_t)n_docs >= 0xFFFFFFFEull) return fail(TK_ERR_INVALID_ARGUMENT, "too many sequences in one call; shard the batch");
    tkk::DecodeLayout L;
    size_t ws_bytes = tkk::decode_workspace_bytes(n_ids, n_docs, cap, &L);
    CUDA_OR_FAIL(t->ws.ensure(ws_bytes));
    cudaError_t e = tkk::decode_device(t->tables, d_ids, d_tok_off, 0, n_docs, n_ids, policy, d_out, cap, d_byte_off, d_status, t->ws.p,
                                       L, st);
    if (e != cudaSuccess) return fail(TK_ERR_CUDA, "decode launch: %s", cudaGetErrorString(e));
    uint32_t small[64];
    CUDA_OR_FAIL(cudaMemcpyAsync(small, (unsigned char*)t->ws.p + L.off_small, sizeof small, cudaMemcpyDeviceToHost, st));
    CUDA_OR_FAIL(cudaStreamSynchronize(st));
    const uint32_t flags = small[tkk::TKK_S_FLAGS];
    uint64_t total_out, first_bad;
    memcpy(&total_out, small + tkk::TKK_S_TOTAL, 8);
    memcpy(&first_bad, small + tkk::TKK_S_BADDOC, 8);
    if (flags & tkk::TKK_FLAG_BAD_OFFSETS)
        return fail(TK_ERR_INVALID_ARGUMENT, "id offsets must start at 0, be non-decreasing and end at the id count");
    *n_bytes = total_out;
    if (flags & tkk::TKK_FLAG_OUT_FULL)
        return fail(TK_ERR_BUFFER_TOO_SMALL, "byte buffer holds %llu bytes, %llu needed", (unsigned long long)cap,
                    (unsigned long long)total_out);
    if (first_bad != ~0ull) {
        if (bad_doc) *bad_doc = first_bad;
        int32_t s = 0;
        // the status of the first failing sequence decides the error (Result<Vec<_>> semantics)
        if (d_status) {
            CUDA_OR_FAIL(cudaMemcpyAsync(&s, d_status + first_bad, 4, cudaMemcpyDeviceToHost, st));
            CUDA_OR_FAIL(cudaStreamSynchronize(st));
        }
        if (s == TK_ERR_SPECIAL_TOKEN_POLICY)
            return fail(TK_ERR_SPECIAL_TOKEN_POLICY, "Decoding tokens that contain special tokens is not allowed (sequence %llu)",
                        (unsigned long long)first_bad);
        return fail(TK_ERR_TOKENIZERS, "decode failed for sequence %llu: unknown token id or the bytes of an ordinary run are not valid UTF-8",
                    (unsigned long long)first_bad);
    }
    return TK_OK;
}

extern "C" int tk_decode_batch_device(const tk_tokenizer* tc, const uint32_t* d_ids, const uint64_t* d_tok_off, size_t n_docs,
                                      uint64_t total_ids, int policy, uint8_t* d_out, uint64_t out_capacity, uint64_t* d_byte_off,
                                      int32_t* d_doc_status, uint64_t* n_bytes, uint64_t* bad_doc, void* stream) {
    int rc = check_encode_args(tc, 0, 0);
    if (rc) return rc;
    if (!d_tok_off || !d_byte_off || !n_bytes || (!d_ids && total_ids) || (!d_out && out_capacity))
        return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    tk_tokenizer* t = const_cast<tk_tokenizer*>(tc);
    std::lock_guard<std::mutex> g(t->mu);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device);
    int32_t* st_buf = d_doc_status;
    if (!st_buf) {
        CUDA_OR_FAIL(t->status.ensure((n_docs + 1) * 4));
        st_buf = (int32_t*)t->status.p;
    }
    return run_decode(t, d_ids, d_tok_off, n_docs, total_ids, policy, d_out, out_capacity, d_byte_off, st_buf, n_bytes, bad_doc,
                      (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ host-buffer decode engine
//
// tk_decode_batch: ids in host memory (pageable or pinned), text out in pinned host memory.  The batch is cut at
// sequence boundaries into chunks of about kChunkBytes / 4 ids that flow through kSlots buffer slots on the
// upload / kernel / download streams, like the encode engine.  A sequence larger than 1.5 chunks is cut inside,
// before an id whose bytes start with an ASCII byte (or before a special id): the bytes before such a cut must end a
// character for the run to be valid UTF-8, so validating the slices separately is validating the run.  The size of
// the text is not known in advance: the result buffer is sized from the first chunk's bytes per id, and if that
// turns out too small the call is repeated with the exact size (a host walk over the ids).  Nothing else walks the
// ids on the host.

namespace {

struct DChunk {
    size_t seq_begin, n_seqs;
    uint64_t id_begin, n_ids;
    bool partial, first, last;
};

// may the id sequence be cut BEFORE position p (see above)?
inline bool decode_cut_ok(const tk::HostModel& h, uint32_t id) {
    if (id < h.num_special) return true;
    const size_t r = id - h.num_special;
    if (r >= h.n_vocab()) return true;                         // unknown id: the sequence fails anyway
    const uint32_t a = h.vocab_off[r], b = h.vocab_off[r + 1];
    return b > a && h.vocab_bytes[a] < 0x80u;
}

int plan_decode_chunks(const tk::HostModel& h, const uint32_t* ids, const uint64_t* tok_off, size_t n_seqs, uint64_t chunk_ids,
                       std::vector<DChunk>& out) {
    const uint64_t big = chunk_ids + chunk_ids / 2, kMaxIds = (1ull << 31);
    size_t d = 0;
    while (d < n_seqs) {
        const uint64_t len_d = tok_off[d + 1] - tok_off[d];
        if (len_d > big) {
            uint64_t pos = tok_off[d];
            const uint64_t end = tok_off[d + 1];
            bool first = true;
            while (pos < end) {
                uint64_t cut = end;
                if (end - pos > big) {
                    const uint64_t target = pos + chunk_ids, lo = pos + chunk_ids / 2;
                    cut = 0;
                    for (uint64_t p = target; p > lo; --p) if (decode_cut_ok(h, ids[p])) { cut = p; break; }
                    if (!cut) {
                        const uint64_t hi = std::min(end, pos + kMaxIds);
                        for (uint64_t p = target + 1; p < hi; ++p) if (decode_cut_ok(h, ids[p])) { cut = p; break; }
                    }
                    if (!cut) cut = end;
                }
                if (cut - pos >= kMaxIds) return fail(TK_ERR_INVALID_ARGUMENT, "sequence %zu has no cut point within 2^31 ids", d);
                out.push_back({d, 1, pos, cut - pos, true, first, cut == end});
                first = false;
                pos = cut;
            }
            ++d;
            continue;
        }
        const size_t b = d;
        const uint64_t begin = tok_off[b];
        size_t e = std::upper_bound(tok_off + b + 1, tok_off + n_seqs + 1, begin + chunk_ids) - tok_off;
        if (e > n_seqs) e = n_seqs;
        if (e <= b) e = b + 1;
        for (size_t k = (e - b > 4 ? e - 2 : b); k < e; ++k)
            if (tok_off[k + 1] - tok_off[k] > big) { e = k > b ? k : b + 1; break; }
        out.push_back({b, e - b, begin, tok_off[e] - begin, false, true, true});
        d = e;
    }
    if (out.empty()) out.push_back({0, 0, 0, 0, false, true, true});
    return TK_OK;
}

}  // namespace

static int decode_batch_engine(tk_tokenizer* t, const uint32_t* ids, const uint64_t* tok_off, size_t n_seqs, int policy, uint8_t** out,
                               uint64_t** byte_off, uint64_t* bad_doc, uint64_t forced_cap) {
    const uint64_t n_ids = tok_off[n_seqs];
    std::lock_guard<std::mutex> lock(t->mu);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device);
    bool pageable = false;
    if (n_ids) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, ids) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = pa.type == cudaMemoryTypeUnregistered;
    }
    if (n_ids * 4 < (1u << 16)) pageable = false;
    std::vector<DChunk> chunks;
    const uint64_t chunk_ids = std::max<uint64_t>(1024, (g_chunk_bytes.load() ? g_chunk_bytes.load() : kDefaultChunkBytes) / 4);
    int rc = plan_decode_chunks(t->host, ids, tok_off, n_seqs, chunk_ids, chunks);
    if (rc) return rc;
    const size_t n_chunks = chunks.size();
    for (auto& ps : t->pipe_st) if (!ps) CUDA_OR_FAIL(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    cudaStream_t st_up = t->pipe_st[0], st_k = t->pipe_st[1], st_down = t->pipe_st[2];

    uint8_t* h_out = nullptr;
    uint64_t h_cap = 0;
    uint64_t* h_boff = (uint64_t*)g_pool.get((n_seqs + 1) * 8);
    if (!h_boff) return fail(TK_ERR_CUDA, "out of pinned host memory");
    auto bail = [&](int code) {
        for (auto& ps : t->pipe_st) if (ps) cudaStreamSynchronize(ps);
        g_pool.put(h_out); g_pool.put(h_boff);
        return code;
    };
    auto launch = [&](size_t i) -> int {
        tk_tokenizer::DecSlot& s = t->dslot[i % tk_tokenizer::kSlots];
        const DChunk& c = chunks[i];
        size_t ws_bytes = tkk::decode_workspace_bytes(c.n_ids, c.n_seqs, s.out.cap, &s.L);
        CUDA_OR_FAIL(s.ws.ensure(ws_bytes));
        cudaError_t e = tkk::decode_device(t->tables, (const uint32_t*)s.ids.p, (const uint64_t*)s.off.p, c.partial ? 0 : c.id_begin, c.n_seqs, c.n_ids,
                                           policy, (uint8_t*)s.out.p, s.out.cap, (uint64_t*)s.boff.p, (int32_t*)s.status.p, s.ws.p, s.L, st_k);
        if (e != cudaSuccess) return fail(TK_ERR_CUDA, "decode launch: %s", cudaGetErrorString(e));
        CUDA_OR_FAIL(tkk::publish_small((const unsigned char*)s.ws.p + s.L.off_small, s.d_small_map, st_k));
        CUDA_OR_FAIL(cudaEventRecord(s.done, st_k));
        return TK_OK;
    };
    auto issue = [&](size_t i) -> int {
        tk_tokenizer::DecSlot& s = t->dslot[i % tk_tokenizer::kSlots];
        const DChunk& c = chunks[i];
        if (!s.h_small) {
            CUDA_OR_FAIL(cudaHostAlloc((void**)&s.h_small, 256, cudaHostAllocMapped));
            CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s.d_small_map, s.h_small, 0));
            CUDA_OR_FAIL(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            CUDA_OR_FAIL(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
            CUDA_OR_FAIL(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
        }
        const size_t off_bytes = (c.n_seqs + 1) * 8;
        CUDA_OR_FAIL(s.ids.ensure(c.n_ids * 4 + 64));
        CUDA_OR_FAIL(s.off.ensure(off_bytes));
        CUDA_OR_FAIL(s.boff.ensure(off_bytes));
        CUDA_OR_FAIL(s.status.ensure((c.n_seqs + 1) * 4));
        CUDA_OR_FAIL(s.out.ensure(std::max<uint64_t>(s.out.cap, c.n_ids * 6 + (1u << 16))));
        const uint32_t* src = ids + c.id_begin;
        const uint64_t* src_off = tok_off + c.seq_begin;
        if (pageable || c.partial) {
            const size_t need = ((off_bytes + 63) & ~(size_t)63) + (pageable ? c.n_ids * 4 : 0) + 64;
            if (s.stage_cap < need) {
                if (s.h_stage) { cudaStreamSynchronize(st_up); cudaFreeHost(s.h_stage); }
                s.h_stage = nullptr; s.stage_cap = 0;
                CUDA_OR_FAIL(cudaHostAlloc((void**)&s.h_stage, need + need / 8, cudaHostAllocDefault));
                s.stage_cap = need + need / 8;
            } else if (i >= (size_t)tk_tokenizer::kSlots) CUDA_OR_FAIL(cudaEventSynchronize(s.ev_in));
            uint64_t* so = (uint64_t*)s.h_stage;
            if (c.partial) { so[0] = 0; so[1] = c.n_ids; }
            else memcpy(so, src_off, off_bytes);
            src_off = so;
            if (pageable) {
                uint8_t* sd = s.h_stage + ((off_bytes + 63) & ~(size_t)63);
                CopyPool::get().parallel_memcpy(sd, src, c.n_ids * 4);
                src = (const uint32_t*)sd;
            }
        }
        if (i >= (size_t)tk_tokenizer::kSlots) CUDA_OR_FAIL(cudaStreamWaitEvent(st_up, s.done, 0));
        if (c.n_ids) CUDA_OR_FAIL(cudaMemcpyAsync(s.ids.p, src, c.n_ids * 4, cudaMemcpyHostToDevice, st_up));
        CUDA_OR_FAIL(cudaMemcpyAsync(s.off.p, src_off, off_bytes, cudaMemcpyHostToDevice, st_up));
        CUDA_OR_FAIL(cudaEventRecord(s.ev_in, st_up));
        CUDA_OR_FAIL(cudaStreamWaitEvent(st_k, s.ev_in, 0));
        if (i >= (size_t)tk_tokenizer::kSlots) CUDA_OR_FAIL(cudaStreamWaitEvent(st_k, s.ev_out, 0));
        return launch(i);
    };
    std::vector<uint64_t> prefix(n_chunks + 1, 0);
    constexpr size_t kAhead = tk_tokenizer::kSlots - 1;
    for (size_t i = 0; i < kAhead && i < n_chunks; ++i) { rc = issue(i); if (rc) return bail(rc); }
    for (size_t i = 0; i < n_chunks; ++i) {
        tk_tokenizer::DecSlot& s = t->dslot[i % tk_tokenizer::kSlots];
        const DChunk& c = chunks[i];
        uint64_t n_bytes = 0;
        for (int attempt = 0;; ++attempt) {
            cudaError_t e = cudaEventSynchronize(s.done);
            if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "decode: %s", cudaGetErrorString(e)));
            const uint32_t flags = s.h_small[tkk::TKK_S_FLAGS];
            uint64_t first_bad;
            memcpy(&n_bytes, s.h_small + tkk::TKK_S_TOTAL, 8);
            memcpy(&first_bad, s.h_small + tkk::TKK_S_BADDOC, 8);
            if (flags & tkk::TKK_FLAG_BAD_OFFSETS)
                return bail(fail(TK_ERR_INVALID_ARGUMENT, "id offsets must start at 0, be non-decreasing and end at the id count"));
            if (flags & tkk::TKK_FLAG_OUT_FULL) {
                // this chunk's text is longer than the slot's buffer: the kernel reported the size, grow and run it again
                if (attempt) return bail(fail(TK_ERR_CUDA, "decode buffer could not be grown"));
                cudaError_t g = s.out.ensure(n_bytes + 4096);
                if (g != cudaSuccess) return bail(fail(TK_ERR_CUDA, "decode buffer: %s", cudaGetErrorString(g)));
                rc = launch(i);
                if (rc) return bail(rc);
                continue;
            }
            if (first_bad != ~0ull) {
                int32_t st = 0;
                e = cudaMemcpy(&st, (int32_t*)s.status.p + first_bad, 4, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "decode: %s", cudaGetErrorString(e)));
                const uint64_t bad = c.seq_begin + first_bad;
                if (bad_doc) *bad_doc = bad;
                if (st == TK_ERR_SPECIAL_TOKEN_POLICY)
                    return bail(fail(TK_ERR_SPECIAL_TOKEN_POLICY, "Decoding tokens that contain special tokens is not allowed (sequence %llu)",
                                     (unsigned long long)bad));
                return bail(fail(TK_ERR_TOKENIZERS, "decode failed for sequence %llu: unknown token id or the bytes of an ordinary run are not valid UTF-8",
                                 (unsigned long long)bad));
            }
            break;
        }
        if (i == 0) {
            uint64_t cap = forced_cap;
            if (!cap) cap = n_chunks == 1 ? n_bytes : (uint64_t)((double)n_bytes / (double)std::max<uint64_t>(1, c.n_ids) * 1.25 * (double)n_ids) + (1u << 16);
            h_out = (uint8_t*)g_pool.get(cap + 1);
            if (!h_out) return bail(fail(TK_ERR_CUDA, "out of pinned host memory"));
            h_cap = cap;
        }
        if (prefix[i] + n_bytes > h_cap) { bail(TK_OK); return TK_ERR_BUFFER_TOO_SMALL; }     // caller repeats with the exact size
        cudaError_t e = cudaSuccess;
        if (n_bytes) e = cudaMemcpyAsync(h_out + prefix[i], s.out.p, n_bytes, cudaMemcpyDeviceToHost, st_down);
        if (e == cudaSuccess && !c.partial && c.n_seqs) e = cudaMemcpyAsync(h_boff + c.seq_begin, s.boff.p, c.n_seqs * 8, cudaMemcpyDeviceToHost, st_down);
        if (e == cudaSuccess) e = cudaEventRecord(s.ev_out, st_down);
        if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "copying text back: %s", cudaGetErrorString(e)));
        prefix[i + 1] = prefix[i] + n_bytes;
        if (i + kAhead < n_chunks) { rc = issue(i + kAhead); if (rc) return bail(rc); }
    }
    {
        cudaError_t e = cudaStreamSynchronize(st_down);
        if (e != cudaSuccess) return bail(fail(TK_ERR_CUDA, "copying text back: %s", cudaGetErrorString(e)));
    }
    for (size_t i = 0; i < n_chunks; ++i) {
        const DChunk& c = chunks[i];
        const uint64_t add = prefix[i];
        if (c.partial) { if (c.first) h_boff[c.seq_begin] = add; }
        else if (add) for (size_t d = c.seq_begin; d < c.seq_begin + c.n_seqs; ++d) h_boff[d] += add;
    }
    h_boff[n_seqs] = prefix[n_chunks];
    if (!h_out) { h_out = (uint8_t*)g_pool.get(1); if (!h_out) return bail(fail(TK_ERR_CUDA, "out of host memory")); }
    h_out[prefix[n_chunks]] = 0;
    *out = h_out;
    *byte_off = h_boff;
    return TK_OK;
}

extern "C" int tk_decode_batch(const tk_tokenizer* tc, const uint32_t* ids, const uint64_t* tok_off, size_t n_docs, int policy,
                               uint8_t** out, uint64_t** byte_off, uint64_t* bad_doc) {
    int rc = check_encode_args(tc, 0, 0);
    if (rc) return rc;
    if (!tok_off || !out || !byte_off) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    *byte_off = nullptr;
    if (policy != TK_POLICY_IGNORE && policy != TK_POLICY_KEEP && policy != TK_POLICY_RAISE)
        return fail(TK_ERR_INVALID_ARGUMENT, "unknown special token policy %d", policy);
    {
        uint64_t bad = tok_off[0] != 0;
        for (size_t d = 0; d < n_docs; ++d) bad |= (uint64_t)(tok_off[d + 1] < tok_off[d]);
        if (bad) return fail(TK_ERR_INVALID_ARGUMENT, "id offsets must start at 0, be non-decreasing and end at the id count");
    }
    const uint64_t n_ids = tok_off[n_docs];
    if (!ids && n_ids) return fail(TK_ERR_INVALID_ARGUMENT, "null ids");
    if (((uintptr_t)ids & 3u) != 0) return fail(TK_ERR_INVALID_ARGUMENT, "id pointer must be 4-byte aligned");
    tk_tokenizer* t = const_cast<tk_tokenizer*>(tc);
    rc = decode_batch_engine(t, ids, tok_off, n_docs, policy, out, byte_off, bad_doc, 0);
    if (rc != TK_ERR_BUFFER_TOO_SMALL) return rc;
    // the estimate of the text size was too small: the exact size from the host tables, once
    const tk::HostModel& h = t->host;
    uint64_t cap = 0;
    for (uint64_t i = 0; i < n_ids; ++i) {
        const uint32_t v = ids[i];
        if (v < h.num_special) { if (policy == TK_POLICY_KEEP) cap += h.special_off[v + 1] - h.special_off[v]; }
        else if (v - h.num_special < h.n_vocab()) cap += h.vocab_off[v - h.num_special + 1] - h.vocab_off[v - h.num_special];
    }
    return decode_batch_engine(t, ids, tok_off, n_docs, policy, out, byte_off, bad_doc, cap + 1);
}

// The latency path of tk_decode: one id list of at most kSmallDecodeIds ids in one single-block kernel that reads the ids
// from and writes the text to mapped pinned memory (tk_decode.cu, decode_small_kernel); the slots are the ones the
// short-text encode uses.  *fallback: the text is longer than the kernel's tile, take the batch path.
static int decode_small_ids(tk_tokenizer* t, const uint32_t* ids, size_t n, int policy, uint8_t** out, size_t* n_out, bool* fallback) {
    *fallback = false;
    tk_tokenizer::FastSlot* s = nullptr;
    for (auto& f : t->fast) if (f.mu.try_lock()) { s = &f; break; }
    if (!s) { s = &t->fast[0]; s->mu.lock(); }
    std::lock_guard<std::mutex> g(s->mu, std::adopt_lock);
    DeviceGuard dg(t->device);
    if (!dg.ok) return fail(TK_ERR_CUDA, "cudaSetDevice(%d) failed", t->device);
    static_assert(tkk::kSmallDecodeIds * 4 <= tkk::kSmallMaxBytes + 128 && (8 + tkk::kSmallDecodeBytes / 4) <= tkk::kSmallOutWords,
                  "the decode tile fits the latency slots");
    if (!s->st) {
        CUDA_OR_FAIL(cudaHostAlloc((void**)&s->h_in, tkk::kSmallMaxBytes + 128, cudaHostAllocMapped));
        CUDA_OR_FAIL(cudaHostAlloc((void**)&s->h_out, tkk::kSmallOutWords * 4, cudaHostAllocMapped));
        CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s->d_in, s->h_in, 0));
        CUDA_OR_FAIL(cudaHostGetDevicePointer((void**)&s->d_out, s->h_out, 0));
        memset(s->h_out, 0, 64);
        CUDA_OR_FAIL(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    }
    if (n) memcpy(s->h_in, ids, n * 4);
    const uint32_t seq = ++s->seq ? s->seq : ++s->seq;
    cudaError_t e = tkk::decode_small(t->tables, (const uint32_t*)s->d_in, (uint32_t)n, policy, s->d_out, seq, s->st);
    if (e != cudaSuccess) return fail(TK_ERR_CUDA, "decode launch: %s", cudaGetErrorString(e));
    volatile uint32_t* hdr = s->h_out;
    for (uint32_t spin = 1;; ++spin) {
        if (hdr[4] == seq) break;
        if ((spin & 4095u) == 0u) {
            const cudaError_t q = cudaStreamQuery(s->st);
            if (q == cudaSuccess) { if (hdr[4] == seq) break; return fail(TK_ERR_CUDA, "single-block decode kernel ended without a result"); }
            if (q != cudaErrorNotReady) return fail(TK_ERR_CUDA, "single-block decode kernel: %s", cudaGetErrorString(q));
        }
        cpu_relax();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const uint32_t n_bytes = hdr[0], flags = hdr[2];
    const int status = (int)(int32_t)hdr[1];
    if (flags & tkk::kSmallNeedBatch) { *fallback = true; return TK_OK; }
    if (status == TK_ERR_SPECIAL_TOKEN_POLICY)
        return fail(TK_ERR_SPECIAL_TOKEN_POLICY, "Decoding tokens that contain special tokens is not allowed (sequence 0)");
    if (status != TK_OK)
        return fail(TK_ERR_TOKENIZERS, "decode failed for sequence 0: unknown token id or the bytes of an ordinary run are not valid UTF-8");
    uint8_t* b = (uint8_t*)g_pool.get((size_t)n_bytes + 1, false);
    if (!b) return fail(TK_ERR_CUDA, "out of host memory");
    if (n_bytes) memcpy(b, s->h_out + 8, n_bytes);
    b[n_bytes] = 0;
    *out = b;
    *n_out = n_bytes;
    return TK_OK;
}

extern "C" int tk_decode(const tk_tokenizer* t, const uint32_t* ids, size_t n, int policy, uint8_t** out, size_t* n_out) {
    if (!out || !n_out) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    static const bool kNoFast = getenv("TEKKEN_B200_NO_FAST") != nullptr;      // measurements: force the batch path
    if (t && t->device >= 0 && n <= tkk::kSmallDecodeIds && (ids || !n) && !kNoFast &&
        (policy == TK_POLICY_IGNORE || policy == TK_POLICY_KEEP || policy == TK_POLICY_RAISE)) {
        bool fallback = false;
        const int rc = decode_small_ids(const_cast<tk_tokenizer*>(t), ids, n, policy, out, n_out, &fallback);
        if (rc || !fallback) return rc;
    }
    uint64_t off[2] = {0, n};
    uint64_t* byte_off = nullptr;
    int rc = tk_decode_batch(t, ids, off, 1, policy, out, &byte_off, nullptr);
    if (rc) return rc;
    *n_out = (size_t)byte_off[1];
    tk_buffer_free(byte_off);
    return TK_OK;
}

// decode_all (:463-511): same bytes as decode; the element boundaries are a host-side walk over
// the ids (one element per ordinary run, one per kept special id).
extern "C" int tk_decode_all(const tk_tokenizer* t, const uint32_t* ids, size_t n, int policy, uint8_t** out, uint64_t** part_end,
                             size_t* n_parts) {
    if (!out || !part_end || !n_parts) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    size_t n_out = 0;
    int rc = tk_decode(t, ids, n, policy, out, &n_out);
    if (rc) return rc;
    const tk::HostModel& h = t->host;
    std::vector<uint64_t> ends;
    uint64_t o = 0;
    size_t i = 0;
    while (i < n) {
        const bool special = ids[i] < h.num_special;
        size_t j = i;
        while (j < n && (ids[j] < h.num_special) == special) ++j;
        if (special) {
            if (policy == TK_POLICY_KEEP)
                for (size_t k = i; k < j; ++k) { o += h.special_off[ids[k] + 1] - h.special_off[ids[k]]; ends.push_back(o); }
        } else {
            for (size_t k = i; k < j; ++k) { uint32_t r = ids[k] - (uint32_t)h.num_special; o += h.vocab_off[r + 1] - h.vocab_off[r]; }
            ends.push_back(o);
        }
        i = j;
    }
    uint64_t* pe = (uint64_t*)g_pool.get(ends.size() * 8 + 8);
    if (!pe) { tk_buffer_free(*out); *out = nullptr; return fail(TK_ERR_CUDA, "out of host memory"); }
    if (!ends.empty()) memcpy(pe, ends.data(), ends.size() * 8);
    *part_end = pe;
    *n_parts = ends.size();
    return TK_OK;
}

// ------------------------------------------------------------------------------------------ streaming: text file in, id shards out
//
// the step either side of the path -- corpus bytes in, token shards out.  The text file is
// memory-mapped (pageable memory: the engine stages it through pinned buffers with its copy threads), cut into
// documents at a delimiter byte, and encoded window by window (about 1 GiB of text per window, whole documents)
// with the host-buffer engine on all given devices; the ids of window i are written to the shard file by a writer
// thread while window i + 1 is being encoded.  Memory stays bounded by two windows' results whatever the file size.

namespace {

struct OutFile {
    int fd = -1;
    bool npy = false;
    char dtype = '4';
    uint64_t count = 0;
    static constexpr size_t kHeader = 128;
    int open_for(const char* path, bool as_npy, char dt) {
        fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) return fail(TK_ERR_IO, "cannot create %s: %s", path, strerror(errno));
        npy = as_npy; dtype = dt;
        if (npy) return write_header();           // placeholder: the element count is patched in by close()
        return TK_OK;
    }
    int write_header() {
        char h[kHeader];
        memset(h, ' ', sizeof h);
        memcpy(h, "\x93NUMPY\x01\x00", 8);
        const uint16_t hl = kHeader - 10;
        h[8] = (char)(hl & 0xFF); h[9] = (char)(hl >> 8);
        char d[96];
        const int n = snprintf(d, sizeof d, "{'descr': '<u%c', 'fortran_order': False, 'shape': (%llu,), }", dtype, (unsigned long long)count);
        memcpy(h + 10, d, (size_t)n);
        h[kHeader - 1] = '\n';
        if (::pwrite(fd, h, kHeader, 0) != (ssize_t)kHeader) return fail(TK_ERR_IO, "write failed: %s", strerror(errno));
        if (::lseek(fd, 0, SEEK_END) < (off_t)kHeader) ::lseek(fd, kHeader, SEEK_SET);
        return TK_OK;
    }
    int append(const void* p, size_t bytes, size_t elems) {
        const char* c = (const char*)p;
        while (bytes) {
            const ssize_t w = ::write(fd, c, std::min<size_t>(bytes, (size_t)1 << 30));
            if (w < 0) { if (errno == EINTR) continue; return fail(TK_ERR_IO, "write failed: %s", strerror(errno)); }
            c += w; bytes -= (size_t)w;
        }
        count += elems;
        return TK_OK;
    }
    int close_file() {
        int rc = TK_OK;
        if (fd >= 0) {
            if (npy) rc = write_header();
            if (::close(fd) != 0 && rc == TK_OK) rc = fail(TK_ERR_IO, "close failed: %s", strerror(errno));
            fd = -1;
        }
        return rc;
    }
    ~OutFile() { if (fd >= 0) ::close(fd); }
};

}  // namespace

extern "C" int tk_encode_file(tk_tokenizer* const* handles, size_t n_handles, const char* text_path, int delimiter, int add_bos, int add_eos,
                              const char* tokens_path, const char* offsets_path, int format, tk_file_stats* stats) {
    if (!handles || n_handles == 0 || !text_path || !tokens_path) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    if (format != TK_SHARDS_RAW_U32 && format != TK_SHARDS_NPY) return fail(TK_ERR_INVALID_ARGUMENT, "unknown shard format %d", format);
    if (delimiter < -1 || delimiter > 255) return fail(TK_ERR_INVALID_ARGUMENT, "delimiter must be a byte value or -1");
    for (size_t g = 0; g < n_handles; ++g) {
        int rc = check_encode_args(handles[g], add_bos, add_eos);
        if (rc) return rc;
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int fd = ::open(text_path, O_RDONLY);
    if (fd < 0) return fail(TK_ERR_IO, "cannot open %s: %s", text_path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0) { ::close(fd); return fail(TK_ERR_IO, "cannot stat %s: %s", text_path, strerror(errno)); }
    const uint64_t n = (uint64_t)sb.st_size;
    const uint8_t* data = nullptr;
    if (n) {
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); return fail(TK_ERR_IO, "cannot map %s: %s", text_path, strerror(errno)); }
        madvise(m, n, MADV_SEQUENTIAL);
        data = (const uint8_t*)m;
    }
    ::close(fd);
    struct Unmap { const uint8_t* p; uint64_t n; ~Unmap() { if (p) munmap((void*)p, n); } } unmap{data, n};
    // documents: every document ends with (and includes) its delimiter; a last one without delimiter is kept
    std::vector<uint64_t> off{0};
    if (delimiter < 0) { if (n) off.push_back(n); }
    else {
        const uint8_t* p = data;
        const uint8_t* end = data + n;
        while (p < end) {
            const uint8_t* q = (const uint8_t*)memchr(p, delimiter, (size_t)(end - p));
            p = q ? q + 1 : end;
            off.push_back((uint64_t)(p - data));
        }
    }
    const size_t n_docs = off.size() - 1;
    OutFile ftok, foff;
    int rc = ftok.open_for(tokens_path, format == TK_SHARDS_NPY, '4');
    if (rc) return rc;
    if (offsets_path) { rc = foff.open_for(offsets_path, format == TK_SHARDS_NPY, '8'); if (rc) return rc; }
    static const uint64_t kWindow = [] { const char* e = getenv("TEKKEN_B200_FILE_WINDOW_MB"); const long mb = e ? atol(e) : 0; return (uint64_t)(mb > 0 ? mb : 1024) << 20; }();
    // writer thread state: the result of the previous window
    std::thread writer;
    int writer_rc = TK_OK;
    std::string writer_err;
    uint64_t tok_base = 0;
    auto join_writer = [&]() -> int {
        if (writer.joinable()) writer.join();
        if (writer_rc) return fail(writer_rc, "%s", writer_err.c_str());
        return TK_OK;
    };
    size_t a = 0;
    std::vector<uint64_t> local;
    while (a < n_docs || (n_docs == 0 && a == 0)) {
        size_t b = a;
        while (b < n_docs && (b == a || off[b + 1] - off[a] <= kWindow)) ++b;
        local.resize(b - a + 1);
        for (size_t d = a; d <= b; ++d) local[d - a] = off[d] - off[a];
        uint32_t* ids = nullptr;
        uint64_t* toff = nullptr;
        rc = encode_batch_engine(handles, n_handles, data ? data + off[a] : nullptr, local.data(), b - a, add_bos, add_eos, &ids, &toff);
        if (rc) { join_writer(); return rc; }
        rc = join_writer();                      // the previous window is on disk (its buffers went back to the pool)
        if (rc) { g_pool.put(ids); g_pool.put(toff); return rc; }
        const uint64_t n_tok = toff[b - a], base = tok_base;
        const size_t nd = b - a;
        const bool last = b >= n_docs;
        writer = std::thread([&, ids, toff, n_tok, base, nd, last] {
            int r = ftok.append(ids, (size_t)n_tok * 4, (size_t)n_tok);
            if (r == TK_OK && foff.fd >= 0) {
                for (size_t d = 0; d <= nd; ++d) toff[d] += base;
                r = foff.append(toff, (nd + (last ? 1 : 0)) * 8, nd + (last ? 1 : 0));      // the closing offset once, at the very end
            }
            if (r) { writer_rc = r; writer_err = g_last_error; }
            g_pool.put(ids);
            g_pool.put(toff);
        });
        tok_base += n_tok;
        a = b;
        if (n_docs == 0) break;
    }
    rc = join_writer();
    if (rc) return rc;
    rc = ftok.close_file();
    if (rc == TK_OK) rc = foff.close_file();
    if (rc) return rc;
    if (stats) {
        stats->n_docs = n_docs; stats->n_bytes = n; stats->n_tokens = tok_base;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return TK_OK;
}

// ------------------------------------------------------------------------------------------ audio token counting

// the text-side half of Tekkenizer::encode_audio (src/tekkenizer.rs:728-735 -> AudioEncoder::encode,
// src/audio.rs:555-591): how many [AUDIO] ids a clip turns into.  O(1) integer / f64 arithmetic per clip on the host --
// there is nothing for a GPU to do; resampling and the mel spectrogram are outside this library.
extern "C" int tk_has_audio_support(const tk_tokenizer* t) { return t && t->host.audio.present; }

extern "C" int tk_audio_config_of(const tk_tokenizer* t, tk_audio_config* out) {
    if (!t || !out) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    const tk::AudioConfigData& a = t->host.audio;
    if (!a.present) return fail(TK_ERR_AUDIO, "Audio encoder not configured");
    out->sampling_rate = a.sampling_rate; out->frame_rate = a.frame_rate; out->num_mel_bins = a.num_mel_bins;
    out->hop_length = a.hop_length; out->window_size = a.window_size; out->chunk_length_s = a.chunk_length_s;
    return TK_OK;
}

extern "C" int tk_audio_token_count(const tk_audio_config* cfg, uint64_t n_samples, uint64_t* padded_samples, uint64_t* n_audio_tokens) {
    if (!cfg || !n_audio_tokens) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    tk::AudioConfigData a;
    a.present = true; a.sampling_rate = cfg->sampling_rate; a.frame_rate = cfg->frame_rate; a.num_mel_bins = cfg->num_mel_bins;
    a.hop_length = cfg->hop_length; a.window_size = cfg->window_size; a.chunk_length_s = cfg->chunk_length_s;
    uint64_t padded = 0;
    try {
        tk::audio_token_count(a, n_samples, &padded, n_audio_tokens);
    } catch (const tk::Error& e) {
        return fail(e.code, "%s", e.what());
    }
    if (padded_samples) *padded_samples = padded;
    return TK_OK;
}

extern "C" int tk_encode_audio_tokens(const tk_tokenizer* t, uint64_t n_samples, uint32_t** out, size_t* n_out) {
    if (!t || !out || !n_out) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    if (!t->host.audio.present) return fail(TK_ERR_AUDIO, "Audio encoder not configured");        // src/tekkenizer.rs:731-734
    uint64_t padded = 0, n = 0;
    try {
        tk::audio_token_count(t->host.audio, n_samples, &padded, &n);
    } catch (const tk::Error& e) {
        return fail(e.code, "%s", e.what());
    }
    uint32_t* ids = (uint32_t*)g_pool.get((size_t)(n + 1) * 4, false);
    if (!ids) return fail(TK_ERR_CUDA, "out of host memory");
    ids[0] = t->host.begin_audio_token_id;                                                         // src/audio.rs:586-587
    for (uint64_t i = 0; i < n; ++i) ids[1 + i] = t->host.audio_token_id;
    *out = ids;
    *n_out = (size_t)(n + 1);
    return TK_OK;
}

// ------------------------------------------------------------------------------------------ sharding, misc

extern "C" int tk_shard_plan(const uint64_t* doc_off, size_t n_docs, size_t n_shards, uint64_t* shard_begin) {
    if (!doc_off || !shard_begin || n_shards == 0) return fail(TK_ERR_INVALID_ARGUMENT, "null argument");
    const uint64_t base = doc_off[0], total = doc_off[n_docs] - base;
    shard_begin[0] = 0;
    size_t d = 0;
    for (size_t s = 1; s < n_shards; ++s) {
        // first document whose start is at or past the s-th byte quantile
        const uint64_t target = base + (uint64_t)((unsigned __int128)total * s / n_shards);
        size_t lo = d, hi = n_docs;
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (doc_off[mid] < target) lo = mid + 1; else hi = mid;
        }
        d = lo;
        shard_begin[s] = d;
    }
    shard_begin[n_shards] = n_docs;
    return TK_OK;
}

extern "C" void tk_set_chunk_bytes(uint64_t bytes) { g_chunk_bytes.store(bytes ? std::max<uint64_t>(bytes, 4096) : 0); }

extern "C" void tk_set_pack_ids(int mode) { g_pack_mode.store(mode == 0 || mode == 18 || mode == 24 ? mode : -1); }

extern "C" int tk_debug_unpack_ids(const uint8_t* src, size_t n, int bits, uint32_t* dst) {
    if ((bits != 18 && bits != 24) || (n && (!src || !dst))) return fail(TK_ERR_INVALID_ARGUMENT, "bits must be 18 or 24");
    CopyPool::get().parallel_unpack(dst, src, n, bits);
    return TK_OK;
}

extern "C" long long tk_debug_bounds_violations(const tk_tokenizer* t, uint64_t* detail4) {
    if (!t || t->device < 0) return -1;
    DeviceGuard dg(t->device);
    cudaDeviceSynchronize();
    unsigned long long d[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
    const long long n = tkk::debug_bounds_violations(d), n2 = tkk::decode_debug_bounds_violations(d2);
    if (n < 0 || n2 < 0) return n < 0 ? n : n2;
    if (detail4) for (int i = 0; i < 4; ++i) detail4[i] = n ? d[i] : d2[i];     // (decode lines are reported + 1,000,000)
    if (detail4) detail4[0] = (unsigned long long)(n + n2);
    return n + n2;
}

extern "C" uint64_t tk_kernel_launch_count(void) { return tkk::launch_count(); }

extern "C" void tk_set_stage_timing(tk_tokenizer* t, int enabled) {
    if (t) t->timing = enabled != 0;
}

extern "C" size_t tk_last_encode_counters(const tk_tokenizer* t, uint64_t* out, size_t cap) {
    if (!t || !out || !t->dev_slot.h_small) return 0;
    const uint32_t* small = t->dev_slot.h_small;
    uint64_t v[TKK_N_CLASSES + 8];
    for (int c = 0; c < TKK_N_CLASSES; ++c) v[c] = small[tkk::TKK_S_QN + c];
    v[TKK_N_CLASSES] = small[tkk::TKK_S_NLONG];
    v[TKK_N_CLASSES + 1] = small[tkk::TKK_S_NHUGE] + small[tkk::TKK_S_NMID];
    memcpy(&v[TKK_N_CLASSES + 2], small + tkk::TKK_S_PAIRLOOK, 8);
    memcpy(&v[TKK_N_CLASSES + 3], small + tkk::TKK_S_BPLOOK, 8);
    v[TKK_N_CLASSES + 4] = small[tkk::TKK_S_ROUNDS1]; v[TKK_N_CLASSES + 5] = small[tkk::TKK_S_ROUNDSM];
    v[TKK_N_CLASSES + 6] = small[tkk::TKK_S_ROUNDSCUT]; v[TKK_N_CLASSES + 7] = small[tkk::TKK_S_APPLIEDM];
    const size_t n = std::min(cap, (size_t)(TKK_N_CLASSES + 8));
    for (size_t i = 0; i < n; ++i) out[i] = v[i];
    return n;
}

extern "C" size_t tk_last_stage_times(const tk_tokenizer* t, const char** names, float* ms, size_t cap) {
    if (!t) return 0;
    size_t n = std::min(cap, t->stage_names.size());
    for (size_t i = 0; i < n; ++i) {
        if (names) names[i] = t->stage_names[i].c_str();
        if (ms) ms[i] = t->stage_ms[i];
    }
    return n;
}
