/*
 * tekken_b200.h -- C ABI of the B200-native Tekkenizer encode/decode path.
 *
 * This is the drop-in boundary: the entry points a `Tekkenizer` shim in the reference's host
 * language (Rust, over `extern "C"`) binds in place of
 * `tiktoken_rs::CoreBPE` and the glue around it.  Every function cites the reference
 * interface it replaces (paths are into jorge-menjivar/tekken-rs).
 *
 * Conventions
 *   - Return value: TK_OK (0) or a negative tk_status.  One code per variant of the
 *     reference's `TokenizerError` (src/errors.rs:23-59) plus a few boundary-only codes.
 *     The message of the last failure on the calling thread is at tk_last_error().
 *   - Ownership: output buffers returned through `T** out` are allocated by the library
 *     (pinned host memory) and released with tk_buffer_free().  The reference returns owned
 *     `Vec<u32>` / `String` (src/tekkenizer.rs:383, 440).
 *   - Text is UTF-8.  The reference takes `&str`, which is valid by construction; this ABI
 *     takes bytes and rejects invalid UTF-8 with TK_ERR_INVALID_UTF8 instead of guessing.
 *   - A tokenizer handle is immutable after construction and may be used from many host
 *     threads at once (the reference's methods all take `&self`).
 *   - There is no CPU fallback: encode/decode run on the CUDA device the handle was created
 *     on and fail with TK_ERR_CUDA if that is impossible.
 */
#ifndef TEKKEN_B200_H
#define TEKKEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tk_tokenizer tk_tokenizer;

/* src/errors.rs:23-59 */
typedef enum tk_status {
    TK_OK = 0,
    TK_ERR_IO = -1,                   /* TokenizerError::Io                  :25-26 */
    TK_ERR_JSON = -2,                 /* TokenizerError::Json                :29-30 */
    TK_ERR_BASE64 = -3,               /* TokenizerError::Base64              :33-34 */
    TK_ERR_TOKENIZERS = -4,           /* TokenizerError::Tokenizers          :37-38 */
    TK_ERR_AUDIO = -5,                /* TokenizerError::Audio (unused here) :41-42 */
    TK_ERR_INVALID_CONFIG = -6,       /* TokenizerError::InvalidConfig       :45-46 */
    TK_ERR_TOKEN_NOT_FOUND = -7,      /* TokenizerError::TokenNotFound       :49-50 */
    TK_ERR_SPECIAL_TOKEN_POLICY = -8, /* TokenizerError::SpecialTokenPolicy  :53-54 */
    TK_ERR_UNSUPPORTED_FORMAT = -9,   /* TokenizerError::UnsupportedFormat   :57-58 */
    /* boundary-only codes (no reference variant) */
    TK_ERR_INVALID_UTF8 = -20,        /* encode input is not valid UTF-8 (`&str` in the reference) */
    TK_ERR_CUDA = -21,                /* CUDA runtime failure / no usable device                  */
    TK_ERR_BUFFER_TOO_SMALL = -22,    /* caller-provided device buffer cannot hold the result     */
    TK_ERR_INVALID_ARGUMENT = -23
} tk_status;

/* src/special_tokens.rs:129-136 */
typedef enum tk_policy { TK_POLICY_IGNORE = 0, TK_POLICY_KEEP = 1, TK_POLICY_RAISE = 2 } tk_policy;

/* src/config.rs:97-103 */
typedef enum tk_version { TK_V3 = 3, TK_V7 = 7, TK_V11 = 11, TK_V13 = 13 } tk_version;

/* Which split pattern a handle uses.  TK_SPLIT_REFERENCE: the literal hard-coded at src/tekkenizer.rs:123 -- what the
   reference computes (it drops `config.pattern`, `_pattern` at :74).  TK_SPLIT_CONFIG: the pattern STORED in
   tekken.json (Mistral's own Tekken regex, the one tests/test_small_vocab.rs:62 and examples/basic_usage.rs:141
   write and `mistral_common` compiles) -- SURVEY 8(f) rank 1. */
typedef enum tk_split_mode { TK_SPLIT_REFERENCE = 0, TK_SPLIT_CONFIG = 1 } tk_split_mode;

/* src/config.rs:16-23 `TokenInfo` (token_str is display-only and not needed) */
typedef struct tk_vocab_entry {
    uint64_t rank;
    const char *token_bytes_b64; /* base64 (STANDARD alphabet, padded), NUL-terminated */
} tk_vocab_entry;

/* src/special_tokens.rs:161-168 `SpecialTokenInfo` */
typedef struct tk_special_entry {
    uint64_t rank;
    const char *token_str; /* UTF-8, NUL-terminated */
    int is_control;
} tk_special_entry;

/* ---- construction ------------------------------------------------------------------ */

/* Tekkenizer::from_file (src/tekkenizer.rs:222-248).  `device` is the CUDA ordinal that will
   hold the vocabulary tables and run the kernels; -1 builds a host-only handle (accessors
   work, encode/decode return TK_ERR_CUDA) for use on machines without a GPU. */
int tk_load_file(const char *path, int device, tk_tokenizer **out);

/* Tekkenizer::new (src/tekkenizer.rs:71-191).  `pattern` is accepted and ignored exactly as
   the reference ignores it (`_pattern`, :74): the split pattern is the literal of :123.
   `special` may be NULL with n_special == 0. */
int tk_new(const tk_vocab_entry *vocab, size_t n_vocab, const tk_special_entry *special,
           size_t n_special, const char *pattern, size_t vocab_size, size_t num_special_tokens,
           int version, int device, tk_tokenizer **out);

/* The same constructors with an explicit tk_split_mode (tk_load_file / tk_new use TK_SPLIT_REFERENCE).  With
   TK_SPLIT_CONFIG the file's / argument's pattern must be the Tekken pattern (TK_ERR_INVALID_CONFIG otherwise). */
int tk_load_file_ex(const char *path, int device, int split_mode, tk_tokenizer **out);
int tk_new_ex(const tk_vocab_entry *vocab, size_t n_vocab, const tk_special_entry *special,
              size_t n_special, const char *pattern, size_t vocab_size, size_t num_special_tokens,
              int version, int device, int split_mode, tk_tokenizer **out);
int tk_split_mode_of(const tk_tokenizer *t); /* a tk_split_mode */

/* The 20 built-in special tokens used when the file has no `special_tokens`
   (get_deprecated_special_tokens, src/tekkenizer.rs:827-930).  Returns their count. */
size_t tk_deprecated_special_tokens(const tk_special_entry **out);

/* Drop (the reference frees on scope exit). */
void tk_free(tk_tokenizer *t);

/* ---- accessors (host-side) ----------------------------------------------------------- */

size_t tk_vocab_size(const tk_tokenizer *t);                 /* src/tekkenizer.rs:261-263 */
size_t tk_num_special_tokens(const tk_tokenizer *t);         /* :269-271 */
int tk_version_of(const tk_tokenizer *t);                    /* :277-279, a tk_version */
int tk_device_of(const tk_tokenizer *t);
int tk_get_control_token(const tk_tokenizer *t, const char *token_str, uint32_t *id); /* :331-341 */
int tk_bos_id(const tk_tokenizer *t, uint32_t *id);          /* :286-288 */
int tk_eos_id(const tk_tokenizer *t, uint32_t *id);          /* :295-297 */
int tk_pad_id(const tk_tokenizer *t, uint32_t *id);          /* :304-306 */
int tk_unk_id(const tk_tokenizer *t, uint32_t *id);          /* :313-315 */
int tk_is_special_token(const tk_tokenizer *t, uint32_t id); /* :574-576 */
int tk_is_byte(const tk_tokenizer *t, uint32_t id);          /* :591-600 */
/* vocab()[id] (:348-350): special string, or the lossy UTF-8 rendering of the token bytes.
   The pointer is owned by the handle. */
int tk_vocab_piece(const tk_tokenizer *t, uint32_t id, const char **str, size_t *len);
/* id_to_piece (:617-628) and id_to_byte_piece (:648-695, including the lossy fallback of
   :685-686).  Output is library-allocated; free with tk_buffer_free. */
int tk_id_to_piece(const tk_tokenizer *t, uint32_t id, uint8_t **out, size_t *n);
int tk_id_to_byte_piece(const tk_tokenizer *t, uint32_t id, int policy, uint8_t **out, size_t *n);

/* ---- encode: Tekkenizer::encode (src/tekkenizer.rs:378-405) --------------------------- */

/* One text.  ids = CoreBPE ranks + num_special_tokens, optional BOS first / EOS last.  A text of at most 8,128 bytes is
   encoded by ONE single-block kernel that reads it from and writes the ids to mapped pinned memory (about 23 us host to
   host); longer texts go through the chunk pipeline of tk_encode_batch.  *out is library-allocated (tk_buffer_free). */
int tk_encode(const tk_tokenizer *t, const uint8_t *utf8, size_t len, int add_bos, int add_eos,
              uint32_t **out, size_t *n_out);

/* New: encode_batch.  Documents are data[doc_off[d] .. doc_off[d+1]) (n_docs+1 offsets,
   doc_off[0] == 0).  ids of all documents are returned back to back in *tokens; *tok_off
   gets n_docs+1 offsets into it.  Host buffers in, pinned host buffers out.  The batch may be
   of any size: it is streamed through the device in chunks cut at document boundaries, and a document larger than a
   chunk is cut at context-free piece boundaries (an ASCII space between ASCII letters / digits), so its size is not
   limited either unless it contains no such boundary for 4 GiB.  `data` may be pageable or page-locked memory. */
int tk_encode_batch(const tk_tokenizer *t, const uint8_t *data, const uint64_t *doc_off,
                    size_t n_docs, int add_bos, int add_eos, uint32_t **tokens, uint64_t **tok_off);

/* One call, all GPUs: the same as tk_encode_batch over `n_handles` handles of the SAME tokenizer on different
   devices (tk_load_file(path, g, ..) for g = 0 .. n-1).  The batch's chunks are dealt to the devices, each device runs
   its pipeline on its own host thread, and ids / offsets land in one result buffer in document order (no stitching
   pass; no collective: documents are independent).  SURVEY 8(b) "batch calls fan out to all GPUs internally". */
int tk_encode_batch_multi(tk_tokenizer *const *handles, size_t n_handles, const uint8_t *data,
                          const uint64_t *doc_off, size_t n_docs, int add_bos, int add_eos,
                          uint32_t **tokens, uint64_t **tok_off);

/* Zero-copy form: every pointer is a device pointer on the handle's device, `stream` is a
   cudaStream_t (NULL = default stream).  d_tokens must hold tokens_capacity ids
   (total_bytes + 2*n_docs always suffices); d_tok_off holds n_docs+1 offsets.  The call
   synchronises the stream before returning; *n_tokens is the total id count.  On
   TK_ERR_BUFFER_TOO_SMALL *n_tokens is the capacity that would have been enough.  d_data must
   be 16-byte aligned; total_bytes < 4 GiB and n_docs < 2^32 - 2 per call (TK_ERR_INVALID_ARGUMENT
   otherwise: shard the batch, tk_shard_plan). */
int tk_encode_batch_device(const tk_tokenizer *t, const uint8_t *d_data, const uint64_t *d_doc_off,
                           size_t n_docs, uint64_t total_bytes, int add_bos, int add_eos,
                           uint32_t *d_tokens, uint64_t tokens_capacity, uint64_t *d_tok_off,
                           uint64_t *n_tokens, void *stream);

/* ---- decode: Tekkenizer::decode / decode_all (src/tekkenizer.rs:436-560) -------------- */

/* decode (:436-443): the bytes of the concatenated elements (valid UTF-8 on success).  policy: tk_policy.  Errors as
   the reference: TK_ERR_SPECIAL_TOKEN_POLICY (Raise and a special id, :531-535), TK_ERR_TOKENIZERS (unknown id, or an
   ordinary run that is not valid UTF-8, :555).  At most 2,048 ids that decode to at most 24 KiB take a single-block
   kernel over mapped pinned memory (about 20 us host to host); longer lists the pipeline of tk_decode_batch. */
int tk_decode(const tk_tokenizer *t, const uint32_t *ids, size_t n, int policy, uint8_t **out,
              size_t *n_out);

/* decode_all (:463-511): the concatenated bytes plus the end offset of every element the
   reference would return (one per ordinary run, one per kept special id). */
int tk_decode_all(const tk_tokenizer *t, const uint32_t *ids, size_t n, int policy, uint8_t **out,
                  uint64_t **part_end, size_t *n_parts);

/* New: decode_batch.  Sequences are ids[tok_off[d] .. tok_off[d+1]).  On success *byte_off
   gets n_docs+1 offsets into *out.  If any sequence fails the call returns the error of the
   first failing sequence (Rust `collect::<Result<Vec<_>>>` semantics) and *bad_doc (may be
   NULL) its index. */
int tk_decode_batch(const tk_tokenizer *t, const uint32_t *ids, const uint64_t *tok_off,
                    size_t n_docs, int policy, uint8_t **out, uint64_t **byte_off, uint64_t *bad_doc);

/* Device form.  d_out must hold out_capacity bytes; d_byte_off n_docs+1 offsets;
   d_doc_status (may be NULL) n_docs int32 per-sequence status codes. */
int tk_decode_batch_device(const tk_tokenizer *t, const uint32_t *d_ids, const uint64_t *d_tok_off,
                           size_t n_docs, uint64_t total_ids, int policy, uint8_t *d_out,
                           uint64_t out_capacity, uint64_t *d_byte_off, int32_t *d_doc_status,
                           uint64_t *n_bytes, uint64_t *bad_doc, void *stream);

/* ---- audio token counting: AudioEncoder::encode, the token part (src/audio.rs:555-591) ---- */

/* AudioConfig + AudioSpectrogramConfig (src/audio.rs:18-22, 86-91).  chunk_length_s <= 0 means None. */
typedef struct tk_audio_config {
    uint64_t sampling_rate;
    double frame_rate;
    uint64_t num_mel_bins, hop_length, window_size;
    double chunk_length_s;
} tk_audio_config;

int tk_has_audio_support(const tk_tokenizer *t);                      /* src/tekkenizer.rs:746-748 */
int tk_audio_config_of(const tk_tokenizer *t, tk_audio_config *out);  /* :757-759; TK_ERR_AUDIO when the file has no `audio` */
/* Samples after Audio::pad (:439-463) and the number of [AUDIO] ids (:563-584) for a clip of n_samples samples at
   cfg->sampling_rate (i.e. after the resampling of :557).  Host arithmetic, O(1). */
int tk_audio_token_count(const tk_audio_config *cfg, uint64_t n_samples, uint64_t *padded_samples,
                         uint64_t *n_audio_tokens);
/* AudioEncoding::tokens of encode_audio (src/tekkenizer.rs:728-735): [BEGIN_AUDIO] then one [AUDIO] per frame group,
   ready to be spliced between text ids (examples/audio_tokenization_test.rs:52-62).  TK_ERR_AUDIO without audio config. */
int tk_encode_audio_tokens(const tk_tokenizer *t, uint64_t n_samples, uint32_t **out, size_t *n_out);

/* ---- multi-GPU sharding (documents are independent: no collective) -------------------- */

/* Byte-balanced contiguous document ranges: shard s gets documents
   [shard_begin[s], shard_begin[s+1]).  shard_begin has n_shards+1 entries. */
int tk_shard_plan(const uint64_t *doc_off, size_t n_docs, size_t n_shards, uint64_t *shard_begin);

/* ---- streaming: text file in, id shards out ------------------------- */

typedef enum tk_shard_format {
    TK_SHARDS_RAW_U32 = 0, /* ids: little-endian u32, back to back (the usual training `.bin`); offsets: u64 */
    TK_SHARDS_NPY = 1      /* the same data behind a NumPy .npy header (1-D '<u4' / '<u8') */
} tk_shard_format;

typedef struct tk_file_stats {
    uint64_t n_docs, n_bytes, n_tokens;
    double seconds;
} tk_file_stats;

/* Encode a UTF-8 text file into a shard of ids.  The file is memory-mapped and cut into documents at `delimiter` (a
   byte value such as '\n' or 0; every document includes its delimiter; -1: the whole file is one document), encoded
   window by window on all `n_handles` devices (handles of one tokenizer, as for tk_encode_batch_multi) while the
   previous window's ids are written out, so memory use does not grow with the file.  `offsets_path` (may be NULL)
   receives n_docs + 1 token offsets. */
int tk_encode_file(tk_tokenizer *const *handles, size_t n_handles, const char *text_path, int delimiter,
                   int add_bos, int add_eos, const char *tokens_path, const char *offsets_path,
                   int format, tk_file_stats *stats);

/* ---- misc ---------------------------------------------------------------------------- */

void tk_buffer_free(void *p);
const char *tk_last_error(void);
const char *tk_status_name(int status);
/* Tuning knob of the host-buffer calls: size of the chunks a batch is streamed through the device in (process-wide;
   0 restores the default of 128 MB, also settable with TEKKEN_B200_CHUNK_MB).  The tests use it to exercise the
   chunk-boundary and document-slicing logic on small inputs. */
void tk_set_chunk_bytes(uint64_t bytes);
/* How the ids of a host-buffer encode cross PCIe (process-wide; also TEKKEN_B200_PACK_IDS): -1 = as an 18- or 24-bit
   stream (the narrowest width the vocabulary fits) for calls of 32 MB of text and more, widened to uint32 on the host
   while the next chunk is on the bus -- the default; 0 = always as uint32; 18 / 24 = that width for every call whose
   ids fit.  The result is the same array either way. */
void tk_set_pack_ids(int mode);
/* Bounds-checked debug build (compile the library with -DTK_DEBUG_BOUNDS; compute-sanitizer substitute): number of
   out-of-range stores the encode and decode kernels of the handle's device refused since the last call of this function,
   with detail4 = {count, source line (tk_kernels.cu; + 1,000,000: tk_decode.cu), index, limit} of the first one.  -1 in a regular build (the checks compile away). */
long long tk_debug_bounds_violations(const tk_tokenizer *t, uint64_t *detail4);
/* Test hook for the host half of the packed-id download (large host-buffer encodes send their ids over PCIe as an
   18- or 24-bit little-endian bit stream, 16 ids per group, and widen them on the host): dst[0..n) = the ids of the
   stream at src, using the same multi-threaded routine the engine uses.  src must be readable 32 bytes past the
   stream's end. */
int tk_debug_unpack_ids(const uint8_t *src, size_t n, int bits, uint32_t *dst);
/* Kernel launches issued by this process so far (for benchmark accounting). */
uint64_t tk_kernel_launch_count(void);
/* Per-stage device time of the most recent tk_encode_batch_device call on this handle, in
   milliseconds, measured with CUDA events on the call's stream when profiling is enabled
   with tk_set_stage_timing(t, 1).  names/ms arrays hold up to `cap` entries; returns count. */
void tk_set_stage_timing(tk_tokenizer *t, int enabled);
size_t tk_last_stage_times(const tk_tokenizer *t, const char **names, float *ms, size_t cap);

/* Work counters of the most recent tk_encode_batch_device call on this handle (benchmark accounting: the
   north star asks for hash probes/s of the merge stage).  out[0..8] = pieces sent to the merge kernels per
   length class (2-4, 5-8, 9-12, 13-16, 17-24, 25-32, 33-48, 49-64, 65-96 bytes); out[9] = pieces longer than
   96 bytes; out[10] = of those, longer than 512 bytes; out[11] = pair-table lookups; out[12] = byte-pair table
   lookups; out[13..16] = rounds of the block-level long-piece kernel: single-rank rounds, multi-rank rounds, multi-rank
   rounds that were cut, merges applied by multi-rank rounds.  Returns the number of entries written (at most `cap`). */
size_t tk_last_encode_counters(const tk_tokenizer *t, uint64_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TEKKEN_B200_H */
