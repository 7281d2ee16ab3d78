//! `tekken_b200` -- the text path of `tekken::Tekkenizer` (jorge-menjivar/tekken-rs, `src/tekkenizer.rs`) running on
//! NVIDIA B200 GPUs: same method names, argument meaning and error variants as the reference's struct, plus the
//! batched calls the accelerated path adds.  Every encode / decode goes to the CUDA library behind
//! `include/tekken_b200.h`; there is no CPU implementation here.
//!
//! ```no_run
//! use tekken_b200::{Tekkenizer, SpecialTokenPolicy};
//! let tk = Tekkenizer::from_file("tekken.json")?;                       // device 0, the reference's split pattern
//! let ids = tk.encode("Hello, world!", true, true)?;                    // [1, 22177, 1044, 4304, 1033, 2]
//! let text = tk.decode(&ids, SpecialTokenPolicy::Keep)?;                // "<s>Hello, world!</s>"
//! let many = tk.encode_batch(&["a", "b c"], false, false)?;             // one call, pipelined through the GPU
//! # Ok::<(), tekken_b200::TokenizerError>(())
//! ```
//!
//! This crate could not be compiled in the repository's build image (no Rust toolchain); the ABI it binds is
//! exercised by the Python mirror (`tekken_rs_b200/tekkenizer.py`) in every GPU test.
pub mod ffi;

use std::ffi::{CStr, CString};
use std::os::raw::c_int;
use std::path::Path;

/// Reference `src/errors.rs:23-59`, plus `Device` for CUDA failures (no reference counterpart).
#[derive(thiserror::Error, Debug)]
pub enum TokenizerError {
    #[error("IO error: {0}")]
    Io(#[from] std::io::Error),
    #[error("JSON error: {0}")]
    Json(String),
    #[error("Base64 decode error: {0}")]
    Base64(String),
    #[error("Tokenizers error: {0}")]
    Tokenizers(String),
    #[error("Audio processing error: {0}")]
    Audio(String),
    #[error("Invalid configuration: {0}")]
    InvalidConfig(String),
    #[error("Token not found: {0}")]
    TokenNotFound(String),
    #[error("Special token policy error: {0}")]
    SpecialTokenPolicy(String),
    #[error("Unsupported audio format: {0}")]
    UnsupportedFormat(String),
    #[error("CUDA device error: {0}")]
    Device(String),
}
pub type Result<T> = std::result::Result<T, TokenizerError>;

/// Reference `src/special_tokens.rs:129-136`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum SpecialTokenPolicy {
    Ignore = 0,
    Keep = 1,
    Raise = 2,
}

/// Reference `src/config.rs:97-103`.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum TokenizerVersion {
    V3 = 3,
    V7 = 7,
    V11 = 11,
    V13 = 13,
}
impl TokenizerVersion {
    pub fn from_string(s: &str) -> Option<Self> {
        match s { "v3" => Some(Self::V3), "v7" => Some(Self::V7), "v11" => Some(Self::V11), "v13" => Some(Self::V13), _ => None }
    }
    pub fn as_str(&self) -> &'static str {
        match self { Self::V3 => "v3", Self::V7 => "v7", Self::V11 => "v11", Self::V13 => "v13" }
    }
    fn from_code(c: c_int) -> Self {
        match c { 3 => Self::V3, 7 => Self::V7, 11 => Self::V11, _ => Self::V13 }
    }
}

/// Which split pattern the handle uses: the literal the reference hard-codes (`src/tekkenizer.rs:123`, what
/// `tekken-rs` computes) or the pattern stored in `tekken.json` (what `mistral_common` computes).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum SplitMode {
    Reference = 0,
    Config = 1,
}

/// Reference `src/config.rs:16-23` (`token_str` is display-only and not needed by the path).
pub struct TokenInfo {
    pub rank: usize,
    /// base64, STANDARD alphabet
    pub token_bytes: String,
}

/// Reference `src/special_tokens.rs:161-168`.
pub struct SpecialTokenInfo {
    pub rank: usize,
    pub token_str: String,
    pub is_control: bool,
}

/// Reference `src/audio.rs:86-91` + `:18-22`, as far as token counting needs it.
pub type AudioConfig = ffi::tk_audio_config;

/// Drop-in for `tekken::tekkenizer::Tekkenizer` (text path) on one B200.
pub struct Tekkenizer {
    handle: *mut ffi::tk_tokenizer,
    vocab: Vec<String>,
    version: TokenizerVersion,
}
// the handle is immutable after construction and the library serialises device work per handle
unsafe impl Send for Tekkenizer {}
unsafe impl Sync for Tekkenizer {}

impl Drop for Tekkenizer {
    fn drop(&mut self) {
        unsafe { ffi::tk_free(self.handle) }
    }
}

/// Copy a library-owned buffer into a `Vec` and release it.
unsafe fn take<T: Copy>(p: *mut T, n: usize) -> Vec<T> {
    let v = if n == 0 { Vec::new() } else { std::slice::from_raw_parts(p, n).to_vec() };
    ffi::tk_buffer_free(p as *mut _);
    v
}

fn pack(texts: &[&str]) -> (Vec<u8>, Vec<u64>) {
    let mut data = Vec::with_capacity(texts.iter().map(|t| t.len()).sum());
    let mut off = Vec::with_capacity(texts.len() + 1);
    off.push(0u64);
    for t in texts {
        data.extend_from_slice(t.as_bytes());
        off.push(data.len() as u64);
    }
    (data, off)
}

impl Tekkenizer {
    fn wrap(handle: *mut ffi::tk_tokenizer) -> Result<Self> {
        let n = unsafe { ffi::tk_vocab_size(handle) };
        let mut vocab = Vec::with_capacity(n);
        for id in 0..n as u32 {
            let (mut p, mut len) = (std::ptr::null(), 0usize);
            ffi::check(unsafe { ffi::tk_vocab_piece(handle, id, &mut p, &mut len) })?;
            let bytes = unsafe { std::slice::from_raw_parts(p as *const u8, len) };
            vocab.push(String::from_utf8_lossy(bytes).into_owned());
        }
        let version = TokenizerVersion::from_code(unsafe { ffi::tk_version_of(handle) });
        Ok(Self { handle, vocab, version })
    }

    /// `Tekkenizer::from_file` (`src/tekkenizer.rs:222-248`) on CUDA device 0.
    pub fn from_file<P: AsRef<Path>>(path: P) -> Result<Self> {
        Self::from_file_on(path, 0, SplitMode::Reference)
    }

    /// The same on a chosen device / split pattern.  `device = -1` builds a host-only handle (accessors only).
    pub fn from_file_on<P: AsRef<Path>>(path: P, device: i32, split: SplitMode) -> Result<Self> {
        let c = CString::new(path.as_ref().to_string_lossy().as_bytes()).map_err(|e| TokenizerError::InvalidConfig(e.to_string()))?;
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::tk_load_file_ex(c.as_ptr(), device, split as c_int, &mut h) })?;
        Self::wrap(h)
    }

    /// `Tekkenizer::new` (`src/tekkenizer.rs:71-191`).  `_pattern` is ignored exactly as the reference ignores it
    /// unless `split` is `SplitMode::Config`; audio configuration travels in the file (use `from_file`).
    #[allow(clippy::too_many_arguments)]
    pub fn new(
        vocab: Vec<TokenInfo>, special_tokens: &[SpecialTokenInfo], pattern: String, vocab_size: usize, num_special_tokens: usize,
        version: TokenizerVersion, device: i32, split: SplitMode,
    ) -> Result<Self> {
        let b64: Vec<CString> = vocab.iter().map(|v| CString::new(v.token_bytes.as_str()).unwrap_or_default()).collect();
        let ve: Vec<ffi::tk_vocab_entry> = vocab.iter().zip(&b64).map(|(v, s)| ffi::tk_vocab_entry { rank: v.rank as u64, token_bytes_b64: s.as_ptr() }).collect();
        let strs: Vec<CString> = special_tokens.iter().map(|s| CString::new(s.token_str.as_str()).unwrap_or_default()).collect();
        let se: Vec<ffi::tk_special_entry> = special_tokens
            .iter()
            .zip(&strs)
            .map(|(s, c)| ffi::tk_special_entry { rank: s.rank as u64, token_str: c.as_ptr(), is_control: s.is_control as c_int })
            .collect();
        let pat = CString::new(pattern).unwrap_or_default();
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe {
            ffi::tk_new_ex(ve.as_ptr(), ve.len(), se.as_ptr(), se.len(), pat.as_ptr(), vocab_size, num_special_tokens, version as c_int, device,
                           split as c_int, &mut h)
        })?;
        Self::wrap(h)
    }

    // ---- accessors (`src/tekkenizer.rs:261-350`, `:574-600`)
    pub fn vocab_size(&self) -> usize { unsafe { ffi::tk_vocab_size(self.handle) } }
    pub fn num_special_tokens(&self) -> usize { unsafe { ffi::tk_num_special_tokens(self.handle) } }
    pub fn version(&self) -> &TokenizerVersion { &self.version }
    pub fn device(&self) -> i32 { unsafe { ffi::tk_device_of(self.handle) } }
    pub fn split_mode(&self) -> SplitMode { if unsafe { ffi::tk_split_mode_of(self.handle) } == 1 { SplitMode::Config } else { SplitMode::Reference } }
    pub fn vocab(&self) -> &[String] { &self.vocab }
    pub fn get_control_token(&self, token_str: &str) -> Result<u32> {
        let c = CString::new(token_str).map_err(|e| TokenizerError::TokenNotFound(e.to_string()))?;
        let mut id = 0u32;
        ffi::check(unsafe { ffi::tk_get_control_token(self.handle, c.as_ptr(), &mut id) })?;
        Ok(id)
    }
    pub fn bos_id(&self) -> Result<u32> { self.get_control_token("<s>") }
    pub fn eos_id(&self) -> Result<u32> { self.get_control_token("</s>") }
    pub fn pad_id(&self) -> Result<u32> { self.get_control_token("<pad>") }
    pub fn unk_id(&self) -> Result<u32> { self.get_control_token("<unk>") }
    pub fn is_special_token(&self, token_id: u32) -> bool { unsafe { ffi::tk_is_special_token(self.handle, token_id) != 0 } }
    pub fn is_byte(&self, token_id: u32) -> bool { unsafe { ffi::tk_is_byte(self.handle, token_id) != 0 } }
    pub fn id_to_piece(&self, token_id: u32) -> Result<String> {
        let (mut p, mut n) = (std::ptr::null_mut::<u8>(), 0usize);
        ffi::check(unsafe { ffi::tk_id_to_piece(self.handle, token_id, &mut p, &mut n) })?;
        Ok(unsafe { String::from_utf8_unchecked(take(p, n)) })
    }
    pub fn id_to_byte_piece(&self, token_id: u32, policy: SpecialTokenPolicy) -> Result<Vec<u8>> {
        let (mut p, mut n) = (std::ptr::null_mut::<u8>(), 0usize);
        ffi::check(unsafe { ffi::tk_id_to_byte_piece(self.handle, token_id, policy as c_int, &mut p, &mut n) })?;
        Ok(unsafe { take(p, n) })
    }

    // ---- encode (`src/tekkenizer.rs:378-405`)
    /// One text.  Texts up to 8128 bytes take the single-block latency path (about 20 us host to host).
    pub fn encode(&self, text: &str, add_bos: bool, add_eos: bool) -> Result<Vec<u32>> {
        let (mut p, mut n) = (std::ptr::null_mut::<u32>(), 0usize);
        ffi::check(unsafe { ffi::tk_encode(self.handle, text.as_ptr(), text.len(), add_bos as c_int, add_eos as c_int, &mut p, &mut n) })?;
        Ok(unsafe { take(p, n) })
    }

    /// New: one call for many texts (documents are independent; the library streams them through the GPU in chunks).
    pub fn encode_batch(&self, texts: &[&str], add_bos: bool, add_eos: bool) -> Result<Vec<Vec<u32>>> {
        let (data, off) = pack(texts);
        let (ids, toff) = self.encode_batch_flat(&data, &off, add_bos, add_eos)?;
        Ok((0..texts.len()).map(|d| ids[toff[d] as usize..toff[d + 1] as usize].to_vec()).collect())
    }

    /// Flat form: documents are `data[doc_off[d]..doc_off[d+1]]`; returns (ids back to back, `n_docs + 1` offsets).
    pub fn encode_batch_flat(&self, data: &[u8], doc_off: &[u64], add_bos: bool, add_eos: bool) -> Result<(Vec<u32>, Vec<u64>)> {
        assert!(!doc_off.is_empty());
        let n_docs = doc_off.len() - 1;
        let (mut tok, mut toff) = (std::ptr::null_mut::<u32>(), std::ptr::null_mut::<u64>());
        ffi::check(unsafe {
            ffi::tk_encode_batch(self.handle, data.as_ptr(), doc_off.as_ptr(), n_docs, add_bos as c_int, add_eos as c_int, &mut tok, &mut toff)
        })?;
        let toff_v = unsafe { take(toff, n_docs + 1) };
        let ids = unsafe { take(tok, toff_v[n_docs] as usize) };
        Ok((ids, toff_v))
    }

    /// One call, all GPUs: `self` and `others` are handles of the same tokenizer on different devices.
    pub fn encode_batch_flat_multi(&self, others: &[&Tekkenizer], data: &[u8], doc_off: &[u64], add_bos: bool, add_eos: bool) -> Result<(Vec<u32>, Vec<u64>)> {
        let n_docs = doc_off.len() - 1;
        let mut hs: Vec<*mut ffi::tk_tokenizer> = vec![self.handle];
        hs.extend(others.iter().map(|t| t.handle));
        let (mut tok, mut toff) = (std::ptr::null_mut::<u32>(), std::ptr::null_mut::<u64>());
        ffi::check(unsafe {
            ffi::tk_encode_batch_multi(hs.as_ptr(), hs.len(), data.as_ptr(), doc_off.as_ptr(), n_docs, add_bos as c_int, add_eos as c_int, &mut tok, &mut toff)
        })?;
        let toff_v = unsafe { take(toff, n_docs + 1) };
        let ids = unsafe { take(tok, toff_v[n_docs] as usize) };
        Ok((ids, toff_v))
    }

    /// Zero-copy form on raw device pointers (`cudaStream_t` as `*mut c_void`); returns the id count.
    ///
    /// # Safety
    /// All pointers must be valid device pointers on this handle's device with the documented capacities.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn encode_batch_device(
        &self, d_data: *const u8, d_doc_off: *const u64, n_docs: usize, total_bytes: u64, add_bos: bool, add_eos: bool, d_tokens: *mut u32,
        capacity: u64, d_tok_off: *mut u64, stream: *mut std::os::raw::c_void,
    ) -> Result<u64> {
        let mut n = 0u64;
        ffi::check(ffi::tk_encode_batch_device(self.handle, d_data, d_doc_off, n_docs, total_bytes, add_bos as c_int, add_eos as c_int, d_tokens,
                                               capacity, d_tok_off, &mut n, stream))?;
        Ok(n)
    }

    // ---- decode (`src/tekkenizer.rs:436-560`)
    pub fn decode(&self, tokens: &[u32], special_token_policy: SpecialTokenPolicy) -> Result<String> {
        let (mut p, mut n) = (std::ptr::null_mut::<u8>(), 0usize);
        ffi::check(unsafe { ffi::tk_decode(self.handle, tokens.as_ptr(), tokens.len(), special_token_policy as c_int, &mut p, &mut n) })?;
        // the library validated every ordinary run as UTF-8 (`:555`); special strings are UTF-8 by construction
        Ok(unsafe { String::from_utf8_unchecked(take(p, n)) })
    }

    pub fn decode_all(&self, tokens: &[u32], special_token_policy: SpecialTokenPolicy) -> Result<Vec<String>> {
        let (mut p, mut pe, mut np) = (std::ptr::null_mut::<u8>(), std::ptr::null_mut::<u64>(), 0usize);
        ffi::check(unsafe { ffi::tk_decode_all(self.handle, tokens.as_ptr(), tokens.len(), special_token_policy as c_int, &mut p, &mut pe, &mut np) })?;
        let ends = unsafe { take(pe, np) };
        let total = ends.last().copied().unwrap_or(0) as usize;
        let raw = unsafe { take(p, total) };
        let mut out = Vec::with_capacity(np);
        let mut s = 0usize;
        for e in ends {
            out.push(unsafe { String::from_utf8_unchecked(raw[s..e as usize].to_vec()) });
            s = e as usize;
        }
        Ok(out)
    }

    /// New: decode many id sequences with one call (`Result<Vec<_>>` semantics: the first failing sequence's error).
    pub fn decode_batch(&self, sequences: &[&[u32]], special_token_policy: SpecialTokenPolicy) -> Result<Vec<String>> {
        let mut ids = Vec::with_capacity(sequences.iter().map(|s| s.len()).sum());
        let mut off = Vec::with_capacity(sequences.len() + 1);
        off.push(0u64);
        for s in sequences {
            ids.extend_from_slice(s);
            off.push(ids.len() as u64);
        }
        let (mut p, mut bo, mut bad) = (std::ptr::null_mut::<u8>(), std::ptr::null_mut::<u64>(), 0u64);
        ffi::check(unsafe {
            ffi::tk_decode_batch(self.handle, ids.as_ptr(), off.as_ptr(), sequences.len(), special_token_policy as c_int, &mut p, &mut bo, &mut bad)
        })?;
        let boff = unsafe { take(bo, sequences.len() + 1) };
        let raw = unsafe { take(p, boff[sequences.len()] as usize) };
        Ok((0..sequences.len()).map(|d| unsafe { String::from_utf8_unchecked(raw[boff[d] as usize..boff[d + 1] as usize].to_vec()) }).collect())
    }

    // ---- audio token counting (`src/audio.rs:555-591`, `src/tekkenizer.rs:728-760`)
    pub fn has_audio_support(&self) -> bool { unsafe { ffi::tk_has_audio_support(self.handle) != 0 } }
    pub fn audio_config(&self) -> Option<AudioConfig> {
        let mut c = AudioConfig::default();
        if unsafe { ffi::tk_audio_config_of(self.handle, &mut c) } == 0 { Some(c) } else { None }
    }
    /// The token sequence `encode_audio` returns for a clip of `n_samples` samples at the configured sampling
    /// rate: `[BEGIN_AUDIO]` followed by one `[AUDIO]` per audio frame.
    pub fn encode_audio_tokens(&self, n_samples: u64) -> Result<Vec<u32>> {
        let (mut p, mut n) = (std::ptr::null_mut::<u32>(), 0usize);
        ffi::check(unsafe { ffi::tk_encode_audio_tokens(self.handle, n_samples, &mut p, &mut n) })?;
        Ok(unsafe { take(p, n) })
    }

    // ---- streaming: text file in, id shards out
    /// Encode a text file (documents end at `delimiter`, e.g. `b'\n'`; `None` = the whole file is one document)
    /// into a `.bin` / `.npy` file of u32 ids plus an optional file of u64 token offsets.
    pub fn encode_file(&self, text_path: &str, delimiter: Option<u8>, add_bos: bool, add_eos: bool, tokens_path: &str, offsets_path: Option<&str>,
                       npy: bool) -> Result<ffi::tk_file_stats> {
        let (tp, op) = (CString::new(text_path).unwrap_or_default(), CString::new(tokens_path).unwrap_or_default());
        let off = offsets_path.map(|s| CString::new(s).unwrap_or_default());
        let mut st = ffi::tk_file_stats::default();
        let hs = [self.handle];
        ffi::check(unsafe {
            ffi::tk_encode_file(hs.as_ptr(), 1, tp.as_ptr(), delimiter.map(|d| d as c_int).unwrap_or(-1), add_bos as c_int, add_eos as c_int, op.as_ptr(),
                                off.as_ref().map(|s| s.as_ptr()).unwrap_or(std::ptr::null()), if npy { ffi::TK_SHARDS_NPY } else { ffi::TK_SHARDS_RAW_U32 }, &mut st)
        })?;
        Ok(st)
    }
}

/// Byte-balanced contiguous document ranges for `n_shards` GPUs / processes (no collective needed).
pub fn shard_plan(doc_off: &[u64], n_shards: usize) -> Result<Vec<u64>> {
    let mut out = vec![0u64; n_shards + 1];
    ffi::check(unsafe { ffi::tk_shard_plan(doc_off.as_ptr(), doc_off.len() - 1, n_shards, out.as_mut_ptr()) })?;
    Ok(out)
}

/// Name of a status code (`"Ok"`, `"InvalidConfig"`, ...).
pub fn status_name(status: i32) -> String {
    unsafe { CStr::from_ptr(ffi::tk_status_name(status)) }.to_string_lossy().into_owned()
}

/// Tuning knob of the host-buffer calls: chunk size in bytes (0 = the default of 128 MB).
pub fn set_chunk_bytes(bytes: u64) {
    unsafe { ffi::tk_set_chunk_bytes(bytes) }
}

/// How ids cross PCIe in host-buffer encodes: -1 = as an 18/24-bit stream for large pipelined calls (default),
/// 0 = always as `u32`, 18 / 24 = that width whenever the ids fit.  The result is the same either way.
pub fn set_pack_ids(mode: i32) {
    unsafe { ffi::tk_set_pack_ids(mode) }
}
