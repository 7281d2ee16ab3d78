//! `extern "C"` declarations: one item per entry of `include/tekken_b200.h`, in the header's order.
//! Status codes map 1:1 onto `TokenizerError` (reference `src/errors.rs:23-59`), see [`check`].
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct tk_tokenizer {
    _private: [u8; 0],
}

#[repr(C)]
pub struct tk_vocab_entry {
    pub rank: u64,
    pub token_bytes_b64: *const c_char,
}

#[repr(C)]
pub struct tk_special_entry {
    pub rank: u64,
    pub token_str: *const c_char,
    pub is_control: c_int,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct tk_audio_config {
    pub sampling_rate: u64,
    pub frame_rate: f64,
    pub num_mel_bins: u64,
    pub hop_length: u64,
    pub window_size: u64,
    /// <= 0: not set
    pub chunk_length_s: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct tk_file_stats {
    pub n_docs: u64,
    pub n_bytes: u64,
    pub n_tokens: u64,
    pub seconds: f64,
}

pub const TK_OK: c_int = 0;
pub const TK_ERR_IO: c_int = -1;
pub const TK_ERR_JSON: c_int = -2;
pub const TK_ERR_BASE64: c_int = -3;
pub const TK_ERR_TOKENIZERS: c_int = -4;
pub const TK_ERR_AUDIO: c_int = -5;
pub const TK_ERR_INVALID_CONFIG: c_int = -6;
pub const TK_ERR_TOKEN_NOT_FOUND: c_int = -7;
pub const TK_ERR_SPECIAL_TOKEN_POLICY: c_int = -8;
pub const TK_ERR_UNSUPPORTED_FORMAT: c_int = -9;
pub const TK_ERR_INVALID_UTF8: c_int = -20;
pub const TK_ERR_CUDA: c_int = -21;
pub const TK_ERR_BUFFER_TOO_SMALL: c_int = -22;
pub const TK_ERR_INVALID_ARGUMENT: c_int = -23;

pub const TK_SPLIT_REFERENCE: c_int = 0;
pub const TK_SPLIT_CONFIG: c_int = 1;

pub const TK_SHARDS_RAW_U32: c_int = 0;
pub const TK_SHARDS_NPY: c_int = 1;

extern "C" {
    // ---- construction
    pub fn tk_load_file(path: *const c_char, device: c_int, out: *mut *mut tk_tokenizer) -> c_int;
    pub fn tk_new(
        vocab: *const tk_vocab_entry, n_vocab: usize, special: *const tk_special_entry, n_special: usize, pattern: *const c_char,
        vocab_size: usize, num_special_tokens: usize, version: c_int, device: c_int, out: *mut *mut tk_tokenizer,
    ) -> c_int;
    pub fn tk_load_file_ex(path: *const c_char, device: c_int, split_mode: c_int, out: *mut *mut tk_tokenizer) -> c_int;
    pub fn tk_new_ex(
        vocab: *const tk_vocab_entry, n_vocab: usize, special: *const tk_special_entry, n_special: usize, pattern: *const c_char,
        vocab_size: usize, num_special_tokens: usize, version: c_int, device: c_int, split_mode: c_int, out: *mut *mut tk_tokenizer,
    ) -> c_int;
    pub fn tk_split_mode_of(t: *const tk_tokenizer) -> c_int;
    pub fn tk_deprecated_special_tokens(out: *mut *const tk_special_entry) -> usize;
    pub fn tk_free(t: *mut tk_tokenizer);

    // ---- accessors
    pub fn tk_vocab_size(t: *const tk_tokenizer) -> usize;
    pub fn tk_num_special_tokens(t: *const tk_tokenizer) -> usize;
    pub fn tk_version_of(t: *const tk_tokenizer) -> c_int;
    pub fn tk_device_of(t: *const tk_tokenizer) -> c_int;
    pub fn tk_get_control_token(t: *const tk_tokenizer, token_str: *const c_char, id: *mut u32) -> c_int;
    pub fn tk_bos_id(t: *const tk_tokenizer, id: *mut u32) -> c_int;
    pub fn tk_eos_id(t: *const tk_tokenizer, id: *mut u32) -> c_int;
    pub fn tk_pad_id(t: *const tk_tokenizer, id: *mut u32) -> c_int;
    pub fn tk_unk_id(t: *const tk_tokenizer, id: *mut u32) -> c_int;
    pub fn tk_is_special_token(t: *const tk_tokenizer, id: u32) -> c_int;
    pub fn tk_is_byte(t: *const tk_tokenizer, id: u32) -> c_int;
    pub fn tk_vocab_piece(t: *const tk_tokenizer, id: u32, s: *mut *const c_char, len: *mut usize) -> c_int;
    pub fn tk_id_to_piece(t: *const tk_tokenizer, id: u32, out: *mut *mut u8, n: *mut usize) -> c_int;
    pub fn tk_id_to_byte_piece(t: *const tk_tokenizer, id: u32, policy: c_int, out: *mut *mut u8, n: *mut usize) -> c_int;

    // ---- encode
    pub fn tk_encode(t: *const tk_tokenizer, utf8: *const u8, len: usize, add_bos: c_int, add_eos: c_int, out: *mut *mut u32, n_out: *mut usize) -> c_int;
    pub fn tk_encode_batch(
        t: *const tk_tokenizer, data: *const u8, doc_off: *const u64, n_docs: usize, add_bos: c_int, add_eos: c_int,
        tokens: *mut *mut u32, tok_off: *mut *mut u64,
    ) -> c_int;
    pub fn tk_encode_batch_multi(
        handles: *const *mut tk_tokenizer, n_handles: usize, data: *const u8, doc_off: *const u64, n_docs: usize, add_bos: c_int,
        add_eos: c_int, tokens: *mut *mut u32, tok_off: *mut *mut u64,
    ) -> c_int;
    pub fn tk_encode_batch_device(
        t: *const tk_tokenizer, d_data: *const u8, d_doc_off: *const u64, n_docs: usize, total_bytes: u64, add_bos: c_int, add_eos: c_int,
        d_tokens: *mut u32, tokens_capacity: u64, d_tok_off: *mut u64, n_tokens: *mut u64, stream: *mut c_void,
    ) -> c_int;

    // ---- decode
    pub fn tk_decode(t: *const tk_tokenizer, ids: *const u32, n: usize, policy: c_int, out: *mut *mut u8, n_out: *mut usize) -> c_int;
    pub fn tk_decode_all(
        t: *const tk_tokenizer, ids: *const u32, n: usize, policy: c_int, out: *mut *mut u8, part_end: *mut *mut u64, n_parts: *mut usize,
    ) -> c_int;
    pub fn tk_decode_batch(
        t: *const tk_tokenizer, ids: *const u32, tok_off: *const u64, n_docs: usize, policy: c_int, out: *mut *mut u8,
        byte_off: *mut *mut u64, bad_doc: *mut u64,
    ) -> c_int;
    pub fn tk_decode_batch_device(
        t: *const tk_tokenizer, d_ids: *const u32, d_tok_off: *const u64, n_docs: usize, total_ids: u64, policy: c_int, d_out: *mut u8,
        out_capacity: u64, d_byte_off: *mut u64, d_doc_status: *mut i32, n_bytes: *mut u64, bad_doc: *mut u64, stream: *mut c_void,
    ) -> c_int;

    // ---- multi-GPU sharding, streaming
    pub fn tk_shard_plan(doc_off: *const u64, n_docs: usize, n_shards: usize, shard_begin: *mut u64) -> c_int;
    pub fn tk_encode_file(
        handles: *const *mut tk_tokenizer, n_handles: usize, text_path: *const c_char, delimiter: c_int, add_bos: c_int, add_eos: c_int,
        tokens_path: *const c_char, offsets_path: *const c_char, format: c_int, stats: *mut tk_file_stats,
    ) -> c_int;

    // ---- audio token counting (src/audio.rs:555-591)
    pub fn tk_has_audio_support(t: *const tk_tokenizer) -> c_int;
    pub fn tk_audio_config_of(t: *const tk_tokenizer, out: *mut tk_audio_config) -> c_int;
    pub fn tk_audio_token_count(cfg: *const tk_audio_config, n_samples: u64, padded_samples: *mut u64, n_audio_tokens: *mut u64) -> c_int;
    pub fn tk_encode_audio_tokens(t: *const tk_tokenizer, n_samples: u64, out: *mut *mut u32, n_out: *mut usize) -> c_int;

    // ---- misc
    pub fn tk_buffer_free(p: *mut c_void);
    pub fn tk_last_error() -> *const c_char;
    pub fn tk_status_name(status: c_int) -> *const c_char;
    pub fn tk_set_chunk_bytes(bytes: u64);
    pub fn tk_debug_bounds_violations(t: *const tk_tokenizer, detail4: *mut u64) -> i64;
    pub fn tk_set_pack_ids(mode: c_int);
    pub fn tk_debug_unpack_ids(src: *const u8, n: usize, bits: c_int, dst: *mut u32) -> c_int;
    pub fn tk_kernel_launch_count() -> u64;
    pub fn tk_set_stage_timing(t: *mut tk_tokenizer, enabled: c_int);
    pub fn tk_last_stage_times(t: *const tk_tokenizer, names: *mut *const c_char, ms: *mut f32, cap: usize) -> usize;
    pub fn tk_last_encode_counters(t: *const tk_tokenizer, out: *mut u64, cap: usize) -> usize;
}

/// Status code -> `TokenizerError` (the message of the failing call is thread-local in the library).
pub fn check(rc: c_int) -> crate::Result<()> {
    use crate::TokenizerError as E;
    if rc == TK_OK {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(tk_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        TK_ERR_IO => E::Io(std::io::Error::new(std::io::ErrorKind::Other, msg)),
        TK_ERR_JSON => E::Json(msg),
        TK_ERR_BASE64 => E::Base64(msg),
        TK_ERR_TOKENIZERS => E::Tokenizers(msg),
        TK_ERR_AUDIO => E::Audio(msg),
        TK_ERR_INVALID_CONFIG => E::InvalidConfig(msg),
        TK_ERR_TOKEN_NOT_FOUND => E::TokenNotFound(msg),
        TK_ERR_SPECIAL_TOKEN_POLICY => E::SpecialTokenPolicy(msg),
        TK_ERR_UNSUPPORTED_FORMAT => E::UnsupportedFormat(msg),
        // boundary-only codes (no reference variant): invalid UTF-8 cannot reach here through `&str`
        TK_ERR_INVALID_UTF8 | TK_ERR_INVALID_ARGUMENT | TK_ERR_BUFFER_TOO_SMALL => E::Tokenizers(msg),
        TK_ERR_CUDA => E::Device(msg),
        _ => E::Tokenizers(msg),
    })
}
