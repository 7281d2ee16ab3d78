// tk_host.h -- host-side model of a loaded Tekkenizer: the tekken.json parser, the reference's
// construction-time validations, and the builders of the tables the kernels read.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "tk_common.h"

namespace tk {

// Mirrors TokenizerError (src/errors.rs:23-59); `code` is a tk_status.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct VocabEntry {
    uint64_t rank;
    std::string token_bytes_b64;
};

struct SpecialEntry {
    uint64_t rank;
    std::string token_str;
    bool is_control;
};

// AudioConfig + AudioSpectrogramConfig (src/audio.rs:18-22, 86-91), as far as token counting needs them
struct AudioConfigData {
    bool present = false;
    uint64_t sampling_rate = 0;
    double frame_rate = 0.0;
    uint64_t num_mel_bins = 0, hop_length = 0, window_size = 0;
    double chunk_length_s = -1.0;      // <= 0: None
};

// What Tekkenizer::from_file reads from the file (src/config.rs:73-82).
struct ModelData {
    std::vector<VocabEntry> vocab;
    bool has_special_tokens = false;
    std::vector<SpecialEntry> special_tokens;
    std::string pattern;
    uint64_t num_vocab_tokens = 0;
    uint64_t default_vocab_size = 0;
    uint64_t default_num_special_tokens = 0;
    std::string version;
    AudioConfigData audio;
};

ModelData parse_tekken_json(const std::string& text);           // throws Error(TK_ERR_JSON)
std::vector<uint8_t> base64_decode_standard(const std::string&); // throws Error(TK_ERR_BASE64)
const std::vector<SpecialEntry>& deprecated_special_tokens();     // src/tekkenizer.rs:827-930
int parse_version(const std::string& s);                          // src/config.rs:124-131; 0 if unknown

// The immutable host state of a tokenizer (src/tekkenizer.rs:34-44 minus audio).
struct HostModel {
    size_t vocab_size = 0;
    size_t num_special = 0;
    int version = 0;
    std::vector<SpecialEntry> special_tokens;                 // incl. <SPECIAL_i> fillers, positional
    std::unordered_map<std::string, uint64_t> special_map;    // token_str -> rank
    std::vector<uint8_t> vocab_bytes;                         // rank order
    std::vector<uint32_t> vocab_off;                          // n_vocab + 1
    std::vector<std::string> vocab_strings;                   // vocab() (:141-155), lossy UTF-8
    uint32_t max_token_len = 0;
    std::string pattern;                                      // config.pattern as given (ignored by the reference, :74)
    AudioConfigData audio;                                    // :42 audio_config
    uint32_t audio_token_id = 0, begin_audio_token_id = 0;    // :158-175 (valid when audio.present)
    void set_audio(const AudioConfigData& a);                 // throws TokenNotFound like Tekkenizer::new(.., Some(cfg))

    // device table images
    std::vector<uint16_t> uni_stage1;
    std::vector<uint8_t> uni_stage2;
    std::vector<TkVocabSlot> vocab_slots;
    std::vector<uint64_t> pair_slots;
    uint32_t pair_mask = 0;                                   // slot count - 1
    std::vector<uint64_t> pair_buckets;                       // the same entries in four-slot buckets (see tk_common.h)
    uint32_t bucket_mask = 0;
    std::vector<uint32_t> byte_pair;                          // 65536 entries, direct-indexed
    std::vector<uint8_t> vocab_pad16;                         // 16 bytes per rank (the encoder's whole-piece check)
    std::vector<uint8_t> vocab_e16;                           // 16 bytes per rank: bytes 0..6, length (0xFF = longer than 15), bytes 7..14; decode's gather source
    std::vector<uint8_t> special_bytes;
    std::vector<uint32_t> special_off;
    size_t n_pairs = 0;

    size_t n_vocab() const { return vocab_off.size() - 1; }
    uint32_t control_token(const std::string& s) const;        // :331-341, throws TokenNotFound
    bool has_control_token(const std::string& s) const { return special_map.count(s) != 0; }

    // Tekkenizer::new (src/tekkenizer.rs:71-191)
    static HostModel build(const std::vector<VocabEntry>& vocab, const std::vector<SpecialEntry>& special,
                           const std::string& pattern_ignored, size_t vocab_size, size_t num_special,
                           int version);
    // Tekkenizer::from_file (src/tekkenizer.rs:222-248)
    static HostModel from_file(const std::string& path);
};

void audio_token_count(const AudioConfigData& c, uint64_t n_samples, uint64_t* padded, uint64_t* n_tokens);   // src/audio.rs:555-584
void build_unicode_tables(std::vector<uint16_t>& stage1, std::vector<uint8_t>& stage2);
void build_cfg_unicode_tables(std::vector<uint16_t>& stage1, std::vector<uint8_t>& stage2);   // TK_SPLIT_CONFIG, see tk_pretok_cfg.h
const char* tekken_config_pattern();                          // the stored pattern the TK_SPLIT_CONFIG kernels implement
std::string utf8_lossy(const uint8_t* p, size_t n);  // String::from_utf8_lossy
bool utf8_valid(const uint8_t* p, size_t n);         // String::from_utf8(..).is_ok()

// Hash of a piece exactly as the kernels compute it (for table construction and host tests).
uint64_t piece_hash(const uint8_t* p, uint32_t len, uint64_t* key8);

}  // namespace tk
