// tk_host.cpp -- tekken.json parser, construction-time validation and device-table builders.
//
// Follows (re-designed, not translated) the load path of the reference:
//   Tekkenizer::from_file     src/tekkenizer.rs:222-248
//   Tekkenizer::new           src/tekkenizer.rs:71-191
//   reload_mergeable_ranks    src/tekkenizer.rs:776-816
//   serde model               src/config.rs:16-82, src/special_tokens.rs:161-168
// The output is not a HashMap-backed CoreBPE but flat table images for the GPU: an
// open-addressing byte-string table, an id-pair merge table, a byte table for decode, and a
// two-stage Unicode class table.
#include "tk_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_set>

#include "../../include/tekken_b200.h"
#include "unicode_ranges.inc"
#include "unicode_subclasses.inc"

namespace tk {

// ------------------------------------------------------------------------------------------ JSON

namespace {

struct JsonReader {
    const char* p;
    const char* end;
    const char* begin;

    [[noreturn]] void fail(const std::string& what) const {
        size_t line = 1, col = 1;
        for (const char* q = begin; q < p && q < end; ++q) {
            if (*q == '\n') { ++line; col = 1; } else ++col;
        }
        std::ostringstream os;
        os << what << " at line " << line << " column " << col;
        throw Error(TK_ERR_JSON, os.str());
    }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    char peek() {
        ws();
        if (p >= end) fail("EOF while parsing a value");
        return *p;
    }
    void expect(char c) {
        if (peek() != c) fail(std::string("expected `") + c + "`");
        ++p;
    }
    bool consume(char c) {
        if (peek() == c) { ++p; return true; }
        return false;
    }
    static void put_utf8(std::string& s, uint32_t cp) {
        if (cp < 0x80) s.push_back((char)cp);
        else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            s.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            s.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); s.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t hex4() {
        if (end - p < 4) fail("EOF while parsing a string");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("invalid escape");
        }
        return v;
    }
    std::string string() {
        expect('"');
        std::string s;
        for (;;) {
            if (p >= end) fail("EOF while parsing a string");
            unsigned char c = (unsigned char)*p++;
            if (c == '"') break;
            if (c < 0x20) fail("control character (\\u0000-\\u001F) found while parsing a string");
            if (c != '\\') { s.push_back((char)c); continue; }
            if (p >= end) fail("EOF while parsing a string");
            char e = *p++;
            switch (e) {
                case '"': s.push_back('"'); break;
                case '\\': s.push_back('\\'); break;
                case '/': s.push_back('/'); break;
                case 'b': s.push_back('\b'); break;
                case 'f': s.push_back('\f'); break;
                case 'n': s.push_back('\n'); break;
                case 'r': s.push_back('\r'); break;
                case 't': s.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF) {
                        if (end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            uint32_t lo = hex4();
                            if (lo < 0xDC00 || lo > 0xDFFF) fail("lone leading surrogate in hex escape");
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        } else fail("unexpected end of hex escape");
                    } else if (cp >= 0xDC00 && cp <= 0xDFFF) fail("lone trailing surrogate in hex escape");
                    put_utf8(s, cp);
                    break;
                }
                default: fail("invalid escape");
            }
        }
        if (!utf8_valid((const uint8_t*)s.data(), s.size())) fail("invalid unicode code point");
        return s;
    }
    // serde `usize`: a non-negative integer literal
    uint64_t usize() {
        ws();
        const char* s = p;
        if (p < end && *p == '-') fail("invalid value: expected usize");
        if (p >= end || *p < '0' || *p > '9') fail("invalid type: expected usize");
        uint64_t v = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            uint64_t d = (uint64_t)(*p - '0');
            if (v > (UINT64_MAX - d) / 10) fail("number out of range");
            v = v * 10 + d;
            ++p;
        }
        if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) { p = s; fail("invalid type: floating point, expected usize"); }
        if (p - s > 1 && *s == '0') { p = s; fail("invalid number"); }
        return v;
    }
    // serde `f64`: any JSON number
    double f64() {
        ws();
        const char* s = p;
        if (p < end && *p == '-') ++p;
        if (p >= end || *p < '0' || *p > '9') { p = s; fail("invalid type: expected f64"); }
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) ++p;
        return strtod(std::string(s, p).c_str(), nullptr);
    }
    bool boolean() {
        ws();
        if (end - p >= 4 && !memcmp(p, "true", 4)) { p += 4; return true; }
        if (end - p >= 5 && !memcmp(p, "false", 5)) { p += 5; return false; }
        fail("invalid type: expected a boolean");
    }
    bool null() {
        ws();
        if (end - p >= 4 && !memcmp(p, "null", 4)) { p += 4; return true; }
        return false;
    }
    void skip_value(int depth = 0) {
        if (depth > 128) fail("recursion limit exceeded");
        char c = peek();
        if (c == '{') {
            ++p;
            if (consume('}')) return;
            for (;;) {
                string();
                expect(':');
                skip_value(depth + 1);
                if (consume(',')) continue;
                expect('}');
                return;
            }
        } else if (c == '[') {
            ++p;
            if (consume(']')) return;
            for (;;) {
                skip_value(depth + 1);
                if (consume(',')) continue;
                expect(']');
                return;
            }
        } else if (c == '"') {
            string();
        } else if (c == 't' || c == 'f') {
            boolean();
        } else if (c == 'n') {
            if (!null()) fail("expected value");
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            if (*p == '-') ++p;
            if (p >= end || *p < '0' || *p > '9') fail("invalid number");
            while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) ++p;
        } else {
            fail("expected value");
        }
    }
    // iterate the keys of an object: f(key) must consume the value
    template <class F>
    void object(F&& f) {
        if (peek() != '{') fail("invalid type: expected a map");
        ++p;
        if (consume('}')) return;
        for (;;) {
            std::string k = string();
            expect(':');
            f(k);
            if (consume(',')) continue;
            expect('}');
            return;
        }
    }
    template <class F>
    void array(F&& f) {
        if (peek() != '[') fail("invalid type: expected a sequence");
        ++p;
        if (consume(']')) return;
        for (;;) {
            f();
            if (consume(',')) continue;
            expect(']');
            return;
        }
    }
};

}  // namespace

ModelData parse_tekken_json(const std::string& text) {
    JsonReader r{text.data(), text.data() + text.size(), text.data()};
    ModelData md;
    bool have_vocab = false, have_config = false;
    bool c_pattern = false, c_nvt = false, c_dvs = false, c_dnst = false, c_version = false;
    r.object([&](const std::string& key) {
        if (key == "vocab") {
            have_vocab = true;
            md.vocab.clear();
            r.array([&] {
                VocabEntry e{};
                bool hr = false, hb = false;
                r.object([&](const std::string& k) {
                    if (k == "rank") { e.rank = r.usize(); hr = true; }
                    else if (k == "token_bytes") { e.token_bytes_b64 = r.string(); hb = true; }
                    else if (k == "token_str") { if (!r.null()) r.string(); }
                    else r.skip_value();
                });
                if (!hr) r.fail("missing field `rank`");
                if (!hb) r.fail("missing field `token_bytes`");
                md.vocab.push_back(std::move(e));
            });
        } else if (key == "special_tokens") {
            if (r.null()) { md.has_special_tokens = false; return; }
            md.has_special_tokens = true;
            md.special_tokens.clear();
            r.array([&] {
                SpecialEntry e{};
                bool hr = false, hs = false, hc = false;
                r.object([&](const std::string& k) {
                    if (k == "rank") { e.rank = r.usize(); hr = true; }
                    else if (k == "token_str") { e.token_str = r.string(); hs = true; }
                    else if (k == "is_control") { e.is_control = r.boolean(); hc = true; }
                    else r.skip_value();
                });
                if (!hr) r.fail("missing field `rank`");
                if (!hs) r.fail("missing field `token_str`");
                if (!hc) r.fail("missing field `is_control`");
                md.special_tokens.push_back(std::move(e));
            });
        } else if (key == "config") {
            have_config = true;
            r.object([&](const std::string& k) {
                if (k == "pattern") { md.pattern = r.string(); c_pattern = true; }
                else if (k == "num_vocab_tokens") { md.num_vocab_tokens = r.usize(); c_nvt = true; }
                else if (k == "default_vocab_size") { md.default_vocab_size = r.usize(); c_dvs = true; }
                else if (k == "default_num_special_tokens") { md.default_num_special_tokens = r.usize(); c_dnst = true; }
                else if (k == "version") { md.version = r.string(); c_version = true; }
                else r.skip_value();
            });
            if (!c_pattern) r.fail("missing field `pattern`");
            if (!c_nvt) r.fail("missing field `num_vocab_tokens`");
            if (!c_dvs) r.fail("missing field `default_vocab_size`");
            if (!c_dnst) r.fail("missing field `default_num_special_tokens`");
            if (!c_version) r.fail("missing field `version`");
        } else if (key == "audio") {
            // ModelData::audio: Option<AudioConfig> (src/config.rs:81, src/audio.rs:86-91).  Only the token COUNT of
            // an audio clip is on this library's path; the waveform processing is not.
            if (r.null()) { md.audio.present = false; return; }
            md.audio = AudioConfigData{};
            md.audio.present = true;
            bool a_sr = false, a_fr = false, a_enc = false;
            r.object([&](const std::string& k) {
                if (k == "sampling_rate") { md.audio.sampling_rate = r.usize(); a_sr = true; }
                else if (k == "frame_rate") { md.audio.frame_rate = r.f64(); a_fr = true; }
                else if (k == "chunk_length_s") { if (r.null()) md.audio.chunk_length_s = -1.0; else md.audio.chunk_length_s = r.f64(); }
                else if (k == "audio_encoding_config") {
                    a_enc = true;
                    bool e_m = false, e_h = false, e_w = false;
                    r.object([&](const std::string& k2) {
                        if (k2 == "num_mel_bins") { md.audio.num_mel_bins = r.usize(); e_m = true; }
                        else if (k2 == "hop_length") { md.audio.hop_length = r.usize(); e_h = true; }
                        else if (k2 == "window_size") { md.audio.window_size = r.usize(); e_w = true; }
                        else r.skip_value();
                    });
                    if (!e_m) r.fail("missing field `num_mel_bins`");
                    if (!e_h) r.fail("missing field `hop_length`");
                    if (!e_w) r.fail("missing field `window_size`");
                }
                else r.skip_value();
            });
            if (!a_sr) r.fail("missing field `sampling_rate`");
            if (!a_fr) r.fail("missing field `frame_rate`");
            if (!a_enc) r.fail("missing field `audio_encoding_config`");
        } else {
            // `image` / other keys are ignored by the reference's serde model too
            r.skip_value();
        }
    });
    r.ws();
    if (r.p != r.end) r.fail("trailing characters");
    if (!have_vocab) r.fail("missing field `vocab`");
    if (!have_config) r.fail("missing field `config`");
    return md;
}

// base64 0.22 STANDARD engine: alphabet A-Za-z0-9+/, canonical padding required.
std::vector<uint8_t> base64_decode_standard(const std::string& s) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof T);
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) T[(unsigned char)a[i]] = (int8_t)i;
        init = true;
    }
    size_t n = s.size();
    if (n % 4 != 0) throw Error(TK_ERR_BASE64, "Invalid padding");
    std::vector<uint8_t> out;
    out.reserve(n / 4 * 3);
    for (size_t i = 0; i < n; i += 4) {
        int v[4];
        int pad = 0;
        for (int j = 0; j < 4; ++j) {
            unsigned char c = (unsigned char)s[i + j];
            if (c == '=') {
                if (i + 4 != n || j < 2) throw Error(TK_ERR_BASE64, "Invalid padding");
                v[j] = 0;
                ++pad;
            } else {
                if (pad) throw Error(TK_ERR_BASE64, "Invalid padding");
                if (T[c] < 0) {
                    std::ostringstream os;
                    os << "Invalid symbol " << (int)c << ", offset " << (i + j) << ".";
                    throw Error(TK_ERR_BASE64, os.str());
                }
                v[j] = T[c];
            }
        }
        uint32_t w = ((uint32_t)v[0] << 18) | ((uint32_t)v[1] << 12) | ((uint32_t)v[2] << 6) | (uint32_t)v[3];
        out.push_back((uint8_t)(w >> 16));
        if (pad < 2) out.push_back((uint8_t)(w >> 8));
        if (pad < 1) out.push_back((uint8_t)w);
        // canonical encoding: the unused low bits of the last symbol must be zero
        if (pad == 2 && (v[1] & 0xF)) throw Error(TK_ERR_BASE64, "Invalid last symbol");
        if (pad == 1 && (v[2] & 0x3)) throw Error(TK_ERR_BASE64, "Invalid last symbol");
    }
    return out;
}

const std::vector<SpecialEntry>& deprecated_special_tokens() {
    static const std::vector<SpecialEntry> v = [] {
        const char* names[] = {"<unk>", "<s>", "</s>", "[INST]", "[/INST]", "[AVAILABLE_TOOLS]",
                               "[/AVAILABLE_TOOLS]", "[TOOL_RESULTS]", "[/TOOL_RESULTS]", "[TOOL_CALLS]",
                               "[IMG]", "<pad>", "[IMG_BREAK]", "[IMG_END]", "[PREFIX]", "[MIDDLE]",
                               "[SUFFIX]", "[SYSTEM_PROMPT]", "[/SYSTEM_PROMPT]", "[TOOL_CONTENT]"};
        std::vector<SpecialEntry> out;
        for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); ++i) out.push_back({i, names[i], true});
        return out;
    }();
    return v;
}

int parse_version(const std::string& s) {
    if (s == "v3") return TK_V3;
    if (s == "v7") return TK_V7;
    if (s == "v11") return TK_V11;
    if (s == "v13") return TK_V13;
    return 0;
}

// ------------------------------------------------------------------------------------------ UTF-8

static inline bool is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }

// Length of the valid scalar value at p, or -(number of bytes a lossy decoder replaces by one
// U+FFFD) when the bytes are not valid (Rust core::str::lossy maximal-subpart rule).
static int utf8_step(const uint8_t* p, size_t n) {
    uint8_t b0 = p[0];
    if (b0 < 0x80) return 1;
    if (b0 < 0xC2 || b0 > 0xF4) return -1;
    if (b0 < 0xE0) {
        if (n < 2 || !is_cont(p[1])) return -1;
        return 2;
    }
    if (b0 < 0xF0) {
        if (n < 2) return -1;
        uint8_t b1 = p[1];
        bool ok = (b0 == 0xE0) ? (b1 >= 0xA0 && b1 <= 0xBF) : (b0 == 0xED) ? (b1 >= 0x80 && b1 <= 0x9F) : is_cont(b1);
        if (!ok) return -1;
        if (n < 3 || !is_cont(p[2])) return -2;
        return 3;
    }
    if (n < 2) return -1;
    uint8_t b1 = p[1];
    bool ok = (b0 == 0xF0) ? (b1 >= 0x90 && b1 <= 0xBF) : (b0 == 0xF4) ? (b1 >= 0x80 && b1 <= 0x8F) : is_cont(b1);
    if (!ok) return -1;
    if (n < 3 || !is_cont(p[2])) return -2;
    if (n < 4 || !is_cont(p[3])) return -3;
    return 4;
}

bool utf8_valid(const uint8_t* p, size_t n) {
    size_t i = 0;
    while (i < n) {
        int s = utf8_step(p + i, n - i);
        if (s < 0) return false;
        i += (size_t)s;
    }
    return true;
}

std::string utf8_lossy(const uint8_t* p, size_t n) {
    std::string out;
    size_t i = 0;
    while (i < n) {
        int s = utf8_step(p + i, n - i);
        if (s > 0) { out.append((const char*)p + i, (size_t)s); i += (size_t)s; }
        else { out.append("\xEF\xBF\xBD"); i += (size_t)(-s); }
    }
    return out;
}

// ------------------------------------------------------------------------------------------ tables

uint64_t piece_hash(const uint8_t* p, uint32_t len, uint64_t* key8) {
    TkPieceHasher h;
    h.init(len);
    uint64_t first = 0;
    for (uint32_t i = 0; i < len; i += 8) {
        uint64_t w = 0;
        uint32_t m = std::min<uint32_t>(8, len - i);
        for (uint32_t j = 0; j < m; ++j) w |= (uint64_t)p[i + j] << (8 * j);
        if (i == 0) first = w;
        h.add(w);
    }
    if (key8) *key8 = first;
    return h.finish();
}

void build_unicode_tables(std::vector<uint16_t>& stage1, std::vector<uint8_t>& stage2) {
    std::vector<uint8_t> cls(0x110000, TK_CL_O);
    auto fill = [&](const uint32_t (*r)[2], int n, uint8_t v) {
        for (int i = 0; i < n; ++i)
            for (uint32_t c = r[i][0]; c <= r[i][1]; ++c) cls[c] = v;
    };
    fill(UNI_L_RANGES, UNI_L_COUNT, TK_CL_L);
    fill(UNI_N_RANGES, UNI_N_COUNT, TK_CL_N);
    fill(UNI_S_RANGES, UNI_S_COUNT, TK_CL_W);
    stage1.assign(TK_UNI_STAGE1_N, 0);
    stage2.clear();
    std::unordered_map<std::string, uint16_t> seen;
    for (uint32_t b = 0; b < TK_UNI_STAGE1_N; ++b) {
        std::string blk(32, '\0');
        for (uint32_t i = 0; i < 128; ++i) {
            uint8_t c = cls[b * 128 + i];
            blk[i >> 2] = (char)((uint8_t)blk[i >> 2] | (uint8_t)(c << ((i & 3) * 2)));
        }
        // a block whose 128 code points share one class (CJK ideographs, Hangul, unassigned planes: most of the
        // non-ASCII text there is) is answered by stage 1 alone: TK_UNI_UNIFORM | class, no second dependent load
        bool uniform = true;
        for (uint32_t i = 1; i < 128; ++i) uniform &= cls[b * 128 + i] == cls[b * 128];
        if (uniform) { stage1[b] = (uint16_t)(TK_UNI_UNIFORM | cls[b * 128]); continue; }
        auto it = seen.find(blk);
        if (it == seen.end()) {
            uint16_t idx = (uint16_t)seen.size();
            seen.emplace(blk, idx);
            stage2.insert(stage2.end(), blk.begin(), blk.end());
            stage1[b] = idx;
        } else {
            stage1[b] = it->second;
        }
    }
    if (stage2.empty()) stage2.assign(32, 0);
}

// Class tables of the TK_SPLIT_CONFIG split: 4-bit classes of tk_pretok_cfg.h (TK_CC_*), two-stage:
// stage1[cp >> 7] -> block, 64 bytes (128 nibbles) per block.  CR/LF are tested inline by tk_cfg_class.
void build_cfg_unicode_tables(std::vector<uint16_t>& stage1, std::vector<uint8_t>& stage2) {
    enum { CC_O = 0, CC_U = 1, CC_LO = 2, CC_C = 3, CC_M = 4, CC_N = 5, CC_W = 6 };
    std::vector<uint8_t> cls(0x110000, CC_O);
    auto fill = [&](const uint32_t (*r)[2], int n, uint8_t v) {
        for (int i = 0; i < n; ++i)
            for (uint32_t c = r[i][0]; c <= r[i][1]; ++c) cls[c] = v;
    };
    fill(UNI_S_RANGES, UNI_S_COUNT, CC_W);
    fill(UNI_N_RANGES, UNI_N_COUNT, CC_N);
    fill(UNI_SUB_UPPER_RANGES, UNI_SUB_UPPER_COUNT, CC_U);
    fill(UNI_SUB_LOWER_RANGES, UNI_SUB_LOWER_COUNT, CC_LO);
    fill(UNI_SUB_BOTH_RANGES, UNI_SUB_BOTH_COUNT, CC_C);
    fill(UNI_SUB_MARK_RANGES, UNI_SUB_MARK_COUNT, CC_M);
    stage1.assign(TK_UNI_STAGE1_N, 0);
    stage2.clear();
    std::unordered_map<std::string, uint16_t> seen;
    for (uint32_t b = 0; b < TK_UNI_STAGE1_N; ++b) {
        std::string blk(64, '\0');
        for (uint32_t i = 0; i < 128; ++i)
            blk[i >> 1] = (char)((uint8_t)blk[i >> 1] | (uint8_t)(cls[b * 128 + i] << ((i & 1) * 4)));
        auto it = seen.find(blk);
        if (it == seen.end()) {
            it = seen.emplace(blk, (uint16_t)seen.size()).first;
            stage2.insert(stage2.end(), blk.begin(), blk.end());
        }
        stage1[b] = it->second;
    }
}

uint32_t HostModel::control_token(const std::string& s) const {
    auto it = special_map.find(s);
    if (it == special_map.end()) {
        std::ostringstream os;
        os << "Unknown control token: '" << s << "'. Available special tokens: [";
        bool first = true;
        for (const auto& kv : special_map) {
            if (!first) os << ", ";
            first = false;
            os << '"' << kv.first << '"';
        }
        os << "]";
        throw Error(TK_ERR_TOKEN_NOT_FOUND, os.str());
    }
    return (uint32_t)it->second;
}

struct BytesKey {
    const uint8_t* p;
    uint32_t n;
    bool operator==(const BytesKey& o) const { return n == o.n && memcmp(p, o.p, n) == 0; }
};
struct BytesKeyHash {
    size_t operator()(const BytesKey& k) const { return (size_t)piece_hash(k.p, k.n, nullptr); }
};

// Mistral's Tekken pattern as tekken.json stores it: the one pattern the TK_SPLIT_CONFIG kernels implement
// (tk_pretok_cfg.h).  A file with another pattern cannot be honoured and is refused in that mode.
const char* tekken_config_pattern() {
    return "[^\\r\\n\\p{L}\\p{N}]?[\\p{Lu}\\p{Lt}\\p{Lm}\\p{Lo}\\p{M}]*[\\p{Ll}\\p{Lm}\\p{Lo}\\p{M}]+"
           "|[^\\r\\n\\p{L}\\p{N}]?[\\p{Lu}\\p{Lt}\\p{Lm}\\p{Lo}\\p{M}]+[\\p{Ll}\\p{Lm}\\p{Lo}\\p{M}]*"
           "|\\p{N}| ?[^\\s\\p{L}\\p{N}]+[\\r\\n/]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+";
}

HostModel HostModel::build(const std::vector<VocabEntry>& vocab, const std::vector<SpecialEntry>& special,
                           const std::string& pattern, size_t vocab_size, size_t num_special,
                           int version) {
    HostModel m;
    m.pattern = pattern;      // the reference drops it (`_pattern`, :74); kept for handles that ask for TK_SPLIT_CONFIG
    // src/tekkenizer.rs:80-87
    if (vocab_size > vocab.size() + num_special) {
        std::ostringstream os;
        os << "vocab_size (" << vocab_size << ") must be <= vocab.len() (" << vocab.size()
           << ") + num_special_tokens (" << num_special << ")";
        throw Error(TK_ERR_INVALID_CONFIG, os.str());
    }
    // :90-98
    {
        std::unordered_set<std::string> seen;
        for (const auto& t : special)
            if (!seen.insert(t.token_str).second)
                throw Error(TK_ERR_INVALID_CONFIG, "Duplicate special token: " + t.token_str);
    }
    // :100-106
    if (special.size() > num_special) {
        std::ostringstream os;
        os << "special_tokens.len() (" << special.size() << ") must be <= num_special_tokens (" << num_special << ")";
        throw Error(TK_ERR_INVALID_CONFIG, os.str());
    }
    // :118 underflows (panics) in the reference; report it as a configuration error instead.
    if (vocab_size < num_special)
        throw Error(TK_ERR_INVALID_CONFIG, "vocab_size must be >= num_special_tokens");
    // :108-116
    m.special_tokens = special;
    for (size_t i = special.size(); i < num_special; ++i)
        m.special_tokens.push_back({i, "<SPECIAL_" + std::to_string(i) + ">", true});
    m.vocab_size = vocab_size;
    m.num_special = num_special;
    m.version = version;

    // reload_mergeable_ranks (:776-816): first `inner` entries by position; bytes -> rank map
    // where a later duplicate byte string overwrites the earlier rank; ranks must be 0..len-1.
    size_t inner = vocab_size - num_special;
    size_t take = std::min(inner, vocab.size());
    std::vector<std::vector<uint8_t>> decoded(take);
    std::unordered_map<BytesKey, uint64_t, BytesKeyHash> ranks;
    ranks.reserve(take * 2);
    for (size_t i = 0; i < take; ++i) {
        decoded[i] = base64_decode_standard(vocab[i].token_bytes_b64);
        const auto& b = decoded[i];
        uint64_t rank = vocab[i].rank;
        if (rank < 256 && !(b.size() == 1 && b[0] == (uint8_t)rank)) {
            std::ostringstream os;
            os << "Expected byte token at rank " << rank << " to be [" << rank << "], got [";
            for (size_t j = 0; j < b.size(); ++j) os << (j ? ", " : "") << (int)b[j];
            os << "]";
            throw Error(TK_ERR_INVALID_CONFIG, os.str());
        }
        ranks[BytesKey{b.data(), (uint32_t)b.size()}] = rank;
    }
    {
        size_t n = ranks.size();
        std::vector<uint8_t> hit(n, 0);
        size_t distinct = 0;
        for (const auto& kv : ranks) {
            if (kv.second >= n) throw Error(TK_ERR_INVALID_CONFIG, "Vocabulary ranks are not contiguous");
            if (!hit[kv.second]) { hit[kv.second] = 1; ++distinct; }
        }
        if (distinct != n) throw Error(TK_ERR_INVALID_CONFIG, "Vocabulary ranks are not contiguous");
    }
    size_t n_vocab = ranks.size();
    // The merge kernels start from single-byte parts, i.e. they need all 256 byte tokens.  The
    // reference builds such a tokenizer but panics inside CoreBPE on the first missing byte.
    if (n_vocab < 256)
        throw Error(TK_ERR_INVALID_CONFIG, "vocabulary must contain the 256 single-byte tokens (ranks 0..255)");
    if (n_vocab + num_special >= (1u << TK_ID_BITS))
        throw Error(TK_ERR_INVALID_CONFIG, "vocabulary too large for the 21-bit id tables");

    // rank-ordered byte table
    std::vector<const std::vector<uint8_t>*> by_rank(n_vocab, nullptr);
    for (size_t i = 0; i < take; ++i) {
        auto it = ranks.find(BytesKey{decoded[i].data(), (uint32_t)decoded[i].size()});
        if (it->second == vocab[i].rank) by_rank[vocab[i].rank] = &decoded[i];
    }
    m.vocab_off.assign(n_vocab + 1, 0);
    for (size_t r = 0; r < n_vocab; ++r) {
        if (!by_rank[r]) throw Error(TK_ERR_INVALID_CONFIG, "Vocabulary ranks are not contiguous");
        m.vocab_off[r + 1] = m.vocab_off[r] + (uint32_t)by_rank[r]->size();
        m.max_token_len = std::max<uint32_t>(m.max_token_len, (uint32_t)by_rank[r]->size());
    }
    if (m.max_token_len > 65535u)      // the decoder keeps token lengths in 16 bits (tk_decode.cu)
        throw Error(TK_ERR_INVALID_CONFIG, "Vocabulary token longer than 65,535 bytes");
    m.vocab_bytes.resize(m.vocab_off[n_vocab]);
    for (size_t r = 0; r < n_vocab; ++r)
        if (!by_rank[r]->empty()) memcpy(&m.vocab_bytes[m.vocab_off[r]], by_rank[r]->data(), by_rank[r]->size());

    // special map (:129-132): token_str -> rank, later entries win
    for (const auto& t : m.special_tokens) m.special_map[t.token_str] = t.rank;

    // vocab() strings (:141-155)
    m.vocab_strings.resize(vocab_size);
    for (size_t i = 0; i < vocab_size; ++i) {
        if (i < num_special) m.vocab_strings[i] = m.special_tokens[i].token_str;
        else {
            size_t r = i - num_special;
            if (r < n_vocab) m.vocab_strings[i] = utf8_lossy(&m.vocab_bytes[m.vocab_off[r]], m.vocab_off[r + 1] - m.vocab_off[r]);
            else m.vocab_strings[i] = "<?>";
        }
    }

    // ---- device table images ----
    build_unicode_tables(m.uni_stage1, m.uni_stage2);

    // byte string -> rank
    {
        uint32_t cap = 1024;
        while (cap < 4 * n_vocab) cap <<= 1;
        m.vocab_slots.assign(cap, TkVocabSlot{0, 0, 0});
        for (size_t r = 0; r < n_vocab; ++r) {
            uint32_t len = m.vocab_off[r + 1] - m.vocab_off[r];
            if (len == 0) continue;  // an empty token can never match a (non-empty) piece
            uint64_t key8;
            uint64_t h = piece_hash(&m.vocab_bytes[m.vocab_off[r]], len, &key8);
            uint32_t i = (uint32_t)h & (cap - 1);
            while (m.vocab_slots[i].len != 0) i = (i + 1) & (cap - 1);
            m.vocab_slots[i] = TkVocabSlot{len <= 8 ? key8 : h, (uint32_t)r, len};
        }
    }
    // (left id, right id) -> rank: every split of every token whose halves are both tokens.
    // Parts of a piece are always tokens (they start as single bytes, all present), so looking
    // up the concatenated BYTES of two adjacent parts -- what CoreBPE does -- is the same as
    // looking up this table by their ids.
    {
        std::vector<uint64_t> entries;
        for (size_t r = 0; r < n_vocab; ++r) {
            uint32_t len = m.vocab_off[r + 1] - m.vocab_off[r];
            const uint8_t* p = &m.vocab_bytes[m.vocab_off[r]];
            for (uint32_t k = 1; k < len; ++k) {
                auto a = ranks.find(BytesKey{p, k});
                if (a == ranks.end()) continue;
                auto b = ranks.find(BytesKey{p + k, len - k});
                if (b == ranks.end()) continue;
                entries.push_back(tk_pair_slot((uint32_t)a->second, (uint32_t)b->second, (uint32_t)r));
            }
        }
        m.n_pairs = entries.size();
        {
            uint32_t cap = 1024;
            while (cap < 3 * entries.size()) cap <<= 1;
            m.pair_slots.assign(cap, 0);
            for (uint64_t e : entries) {
                uint32_t l = (uint32_t)((e >> (2 * TK_ID_BITS)) & TK_ID_MASK), r = (uint32_t)((e >> TK_ID_BITS) & TK_ID_MASK);
                uint32_t i = tk_pair_hash(l, r) & (cap - 1);
                while (m.pair_slots[i] != 0) i = (i + 1) & (cap - 1);
                m.pair_slots[i] = e;
            }
            m.pair_mask = cap - 1;
        }
#if TK_PAIR_BUCKETED
        {
            // buckets of four slots, on average at most 0.55 entries per bucket: about 0.2 % of the buckets are full
            uint32_t nb = 256;
            while ((double)nb * 0.55 < (double)entries.size()) nb <<= 1;
            if (const char* e = getenv("TEKKEN_B200_PAIR_BUCKETS_LOG2")) { const int k = atoi(e); if (k >= 8 && k <= 26) nb = 1u << k; }
            while ((uint64_t)nb * TK_PAIR_BUCKET_SLOTS < 2 * entries.size()) nb <<= 1;      // never more than half full
            m.pair_buckets.assign((size_t)nb * TK_PAIR_BUCKET_SLOTS, 0);
            for (uint64_t e : entries) {
                uint32_t l = (uint32_t)((e >> (2 * TK_ID_BITS)) & TK_ID_MASK), r = (uint32_t)((e >> TK_ID_BITS) & TK_ID_MASK);
                uint32_t b = tk_pair_hash(l, r) & (nb - 1);
                for (;;) {
                    uint64_t* slot = &m.pair_buckets[(size_t)b * TK_PAIR_BUCKET_SLOTS];
                    uint32_t k = 0;
                    while (k < TK_PAIR_BUCKET_SLOTS && slot[k] != 0) ++k;
                    if (k < TK_PAIR_BUCKET_SLOTS) { slot[k] = e; break; }
                    b = (b + 1) & (nb - 1);
                }
            }
            m.bucket_mask = nb - 1;
        }
#endif
    }
    // decode's gather source: every token in an aligned 16-byte cell + its length in a byte
    {
        m.vocab_pad16.assign(16 * std::max<size_t>(n_vocab, 1), 0);
        m.vocab_e16.assign(16 * std::max<size_t>(n_vocab, 1), 0);
        for (size_t r = 0; r < n_vocab; ++r) {
            uint32_t len = m.vocab_off[r + 1] - m.vocab_off[r];
            memcpy(&m.vocab_pad16[16 * r], &m.vocab_bytes[m.vocab_off[r]], std::min<uint32_t>(len, 16));
            // decode's cell: bytes 0..6, the length, bytes 7..14 -- the first 8-byte load gives the length and 7 bytes,
            // which is the whole token for all but a fraction of a percent of the ids in text.  0xFF = longer than 15
            // bytes (see vocab_off)
            if (len <= 15) {
                uint8_t* c = &m.vocab_e16[16 * r];
                const uint8_t* b = &m.vocab_bytes[m.vocab_off[r]];
                for (uint32_t j = 0; j < len; ++j) c[j < 7 ? j : j + 1] = b[j];
                c[7] = (uint8_t)len;
            } else m.vocab_e16[16 * r + 7] = 0xFF;
        }
    }
    // first-round pair ranks (both parts are single bytes): direct-indexed, no probing
    {
        m.byte_pair.assign(65536, TK_INF);
        uint8_t two[2];
        for (uint32_t a = 0; a < 256; ++a)
            for (uint32_t b = 0; b < 256; ++b) {
                two[0] = (uint8_t)a; two[1] = (uint8_t)b;
                auto it = ranks.find(BytesKey{two, 2});
                if (it != ranks.end()) m.byte_pair[(a << 8) | b] = (uint32_t)it->second;
            }
    }
    // special strings, positional (decode(Keep) indexes special_tokens[id], :538)
    m.special_off.assign(num_special + 1, 0);
    for (size_t i = 0; i < num_special; ++i)
        m.special_off[i + 1] = m.special_off[i] + (uint32_t)m.special_tokens[i].token_str.size();
    m.special_bytes.resize(m.special_off[num_special]);
    for (size_t i = 0; i < num_special; ++i)
        if (m.special_off[i + 1] - m.special_off[i] > 65535u) throw Error(TK_ERR_INVALID_CONFIG, "Special token string longer than 65,535 bytes");
    for (size_t i = 0; i < num_special; ++i)
        memcpy(m.special_bytes.data() + m.special_off[i], m.special_tokens[i].token_str.data(),
               m.special_tokens[i].token_str.size());
    return m;
}

HostModel HostModel::from_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(TK_ERR_IO, "cannot open " + path + ": No such file or directory (os error 2)");
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (!utf8_valid((const uint8_t*)text.data(), text.size()))
        throw Error(TK_ERR_IO, "stream did not contain valid UTF-8");  // read_to_string (:223)
    ModelData md = parse_tekken_json(text);
    int version = parse_version(md.version);
    if (!version) throw Error(TK_ERR_INVALID_CONFIG, "Unknown version: " + md.version);
    const std::vector<SpecialEntry>& sp = md.has_special_tokens ? md.special_tokens : deprecated_special_tokens();
    HostModel m = build(md.vocab, sp, md.pattern, md.default_vocab_size, md.default_num_special_tokens, version);
    m.set_audio(md.audio);
    return m;
}

// Tekkenizer::new with Some(audio_config) (src/tekkenizer.rs:157-178): the audio special tokens must exist
void HostModel::set_audio(const AudioConfigData& a) {
    audio = a;
    if (!a.present) return;
    if (!has_control_token("[AUDIO]")) throw Error(TK_ERR_TOKEN_NOT_FOUND, "Audio token not found");
    if (!has_control_token("[BEGIN_AUDIO]")) throw Error(TK_ERR_TOKEN_NOT_FOUND, "BeginAudio token not found");
    audio_token_id = control_token("[AUDIO]");
    begin_audio_token_id = control_token("[BEGIN_AUDIO]");
}

// The token count of AudioEncoder::encode (src/audio.rs:555-591) for a clip of n_samples samples at the configured
// sampling rate: Audio::pad (:439-463), the spectrogram length (:563-578), ceil(length / audio_length_per_tok) (:584,
// :188-199).  Same integer / f64 arithmetic as the reference.
void audio_token_count(const AudioConfigData& c, uint64_t n_samples, uint64_t* padded, uint64_t* n_tokens) {
    if (c.sampling_rate == 0) throw Error(TK_ERR_INVALID_CONFIG, "sampling_rate must be > 0");
    if (!(c.frame_rate > 0.0)) throw Error(TK_ERR_INVALID_CONFIG, "frame_rate must be > 0");
    if (c.hop_length == 0) throw Error(TK_ERR_INVALID_CONFIG, "hop_length must be > 0");
    if (c.window_size == 0) throw Error(TK_ERR_INVALID_CONFIG, "window_size must be > 0");
    uint64_t len = n_samples;
    if (c.chunk_length_s > 0.0) {
        const uint64_t chunk_frames = (uint64_t)(c.chunk_length_s * (double)c.sampling_rate);      // :157-175
        if (chunk_frames == 0) throw Error(TK_ERR_INVALID_CONFIG, "chunk_length_s too small");
        len = (len + chunk_frames - 1) / chunk_frames * chunk_frames;                                // div_ceil * chunk_frames
    } else if (len < c.window_size) {
        len = c.window_size;
    }
    *padded = len;
    uint64_t sig;
    if (len % c.hop_length != 0) {
        const double v = std::ceil((double)len / (double)c.hop_length - 1.0);
        sig = v <= 0.0 ? 0 : (uint64_t)v;                                                            // `as usize` saturates at 0
    } else sig = len / c.hop_length;
    double f = (double)c.sampling_rate / c.frame_rate;
    f /= (double)c.hop_length;
    const uint64_t per_tok = f <= 0.0 ? 0 : (uint64_t)f;                                             // audio_length_per_tok
    if (per_tok == 0) throw Error(TK_ERR_INVALID_CONFIG, "frame_rate * hop_length exceeds sampling_rate: audio_length_per_tok is 0");
    *n_tokens = (uint64_t)std::ceil((double)sig / (double)per_tok);
}

}  // namespace tk
