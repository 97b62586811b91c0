// The UCE artifact: the safetensors file the reference writes with safetensors.torch.save_file (trainscripts/uce_sd_erase.py:85-88,
// uce_sd_debias.py) and reads with load_file (evalscripts/generate-images-sd.py:17-19) — only the edited attn2.to_k / to_v weights,
// fp32, key = module path + ".weight".  Host-only code (no CUDA calls): writer byte-identical to save_file for the same dict
// (checked in tests/test_artifact.py against the safetensors package), reader for any safetensors file with F32 tensors.
//
// Format (safetensors 0.4+):  u64 LE header length N | N bytes of JSON, padded with spaces to a multiple of 8 | raw little-endian
// tensors back to back.  JSON: {"<name>":{"dtype":"F32","shape":[r,c],"data_offsets":[begin,end]},...}, tensors ordered by
// (dtype alignment descending, name ascending) — all F32 here, so by name — with offsets relative to the end of the header.
#include "uce_common.cuh"
#include <algorithm>
#include <cerrno>
#include <cstring>
#include <string>
#include <vector>

namespace uce {
namespace {

void json_escape(const char* s, std::string& out) {
    out.push_back('"');
    for (const unsigned char* p = (const unsigned char*)s; *p; ++p) {
        switch (*p) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (*p < 0x20) { char buf[8]; snprintf(buf, sizeof(buf), "\\u%04x", *p); out += buf; }
                else out.push_back((char)*p);
        }
    }
    out.push_back('"');
}

// ---- minimal JSON reader for the header: objects, arrays, strings, non-negative integers (+ skipping of anything else) ----
struct Cur { const char* p; const char* end; bool ok = true; };
void ws(Cur& c) { while (c.p < c.end && (*c.p == ' ' || *c.p == '\n' || *c.p == '\t' || *c.p == '\r')) ++c.p; }
bool eat(Cur& c, char ch) { ws(c); if (c.p < c.end && *c.p == ch) { ++c.p; return true; } return false; }
bool parse_string(Cur& c, std::string& out) {
    ws(c);
    if (c.p >= c.end || *c.p != '"') return c.ok = false;
    ++c.p; out.clear();
    while (c.p < c.end && *c.p != '"') {
        if (*c.p == '\\') {
            if (++c.p >= c.end) return c.ok = false;
            switch (*c.p) {
                case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break; case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break; case 'f': out.push_back('\f'); break;
                case 'u': {
                    if (c.end - c.p < 5) return c.ok = false;
                    unsigned v = 0;
                    for (int i = 1; i <= 4; ++i) { const char h = c.p[i]; v = v * 16 + (h <= '9' ? h - '0' : (h | 32) - 'a' + 10); }
                    if (v < 0x80) out.push_back((char)v);
                    else if (v < 0x800) { out.push_back((char)(0xC0 | (v >> 6))); out.push_back((char)(0x80 | (v & 63))); }
                    else { out.push_back((char)(0xE0 | (v >> 12))); out.push_back((char)(0x80 | ((v >> 6) & 63))); out.push_back((char)(0x80 | (v & 63))); }
                    c.p += 4; break;
                }
                default: out.push_back(*c.p);
            }
            ++c.p;
        } else out.push_back(*c.p++);
    }
    if (c.p >= c.end) return c.ok = false;
    ++c.p;
    return true;
}
bool parse_uint(Cur& c, unsigned long long& v) {
    ws(c);
    if (c.p >= c.end || *c.p < '0' || *c.p > '9') return c.ok = false;
    v = 0;
    while (c.p < c.end && *c.p >= '0' && *c.p <= '9') v = v * 10 + (unsigned long long)(*c.p++ - '0');
    return true;
}
bool skip_value(Cur& c);
bool skip_container(Cur& c, char open, char close) {
    if (!eat(c, open)) return c.ok = false;
    if (eat(c, close)) return true;
    do {
        if (open == '{') { std::string k; if (!parse_string(c, k) || !eat(c, ':')) return c.ok = false; }
        if (!skip_value(c)) return false;
    } while (eat(c, ','));
    return eat(c, close) ? true : (c.ok = false);
}
bool skip_value(Cur& c) {
    ws(c);
    if (c.p >= c.end) return c.ok = false;
    if (*c.p == '{') return skip_container(c, '{', '}');
    if (*c.p == '[') return skip_container(c, '[', ']');
    if (*c.p == '"') { std::string s; return parse_string(c, s); }
    while (c.p < c.end && *c.p != ',' && *c.p != '}' && *c.p != ']') ++c.p;      // number / true / false / null
    return true;
}

struct Entry { std::string name, dtype; std::vector<long> shape; unsigned long long begin = 0, end = 0; };

}  // namespace
}  // namespace uce

struct uce_artifact {
    FILE* f = nullptr;
    unsigned long long data_start = 0, file_size = 0;
    std::vector<uce::Entry> entries;
};

using namespace uce;

extern "C" {

int uce_artifact_write_f32(const char* path, int n, const char* const* names, const float* const* data, const long* rows, const long* cols) {
    if (!path || n < 0 || (n > 0 && (!names || !data || !rows || !cols))) { set_error("uce_artifact_write_f32: bad argument"); return UCE_E_ARG; }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) {
        if (!names[i] || !data[i] || rows[i] <= 0 || cols[i] <= 0) { set_error("uce_artifact_write_f32: entry %d invalid", i); return UCE_E_ARG; }
        order[i] = i;
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return strcmp(names[a], names[b]) < 0; });
    for (int i = 1; i < n; ++i)
        if (strcmp(names[order[i - 1]], names[order[i]]) == 0) { set_error("uce_artifact_write_f32: duplicate key '%s'", names[order[i]]); return UCE_E_ARG; }
    std::string hdr = "{";
    unsigned long long off = 0;
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        const unsigned long long bytes = (unsigned long long)rows[i] * (unsigned long long)cols[i] * 4ull;
        if (k) hdr.push_back(',');
        json_escape(names[i], hdr);
        char buf[160];
        snprintf(buf, sizeof(buf), ":{\"dtype\":\"F32\",\"shape\":[%ld,%ld],\"data_offsets\":[%llu,%llu]}", rows[i], cols[i], off, off + bytes);
        hdr += buf;
        off += bytes;
    }
    hdr.push_back('}');
    while (hdr.size() % 8) hdr.push_back(' ');
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot create '%s': %s", path, strerror(errno)); return UCE_E_STATE; }
    const unsigned long long hn = hdr.size();
    unsigned char le[8];
    for (int b = 0; b < 8; ++b) le[b] = (unsigned char)(hn >> (8 * b));
    bool ok = fwrite(le, 1, 8, f) == 8 && fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    for (int k = 0; k < n && ok; ++k) {
        const int i = order[k];
        const size_t cnt = (size_t)rows[i] * (size_t)cols[i];
        ok = fwrite(data[i], sizeof(float), cnt, f) == cnt;          // the format is little-endian, and so is every host this library runs on
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) { set_error("write to '%s' failed: %s", path, strerror(errno)); remove(path); return UCE_E_STATE; }
    return 0;
}

int uce_artifact_open(const char* path, uce_artifact** out) {
    if (!path || !out) { set_error("uce_artifact_open: bad argument"); return UCE_E_ARG; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open '%s': %s", path, strerror(errno)); return UCE_E_STATE; }
    auto fail = [&](const char* why) { set_error("'%s' is not a valid safetensors file: %s", path, why); fclose(f); return UCE_E_STATE; };
    if (fseek(f, 0, SEEK_END) != 0) return fail("seek failed");
    const long long fsz = ftell(f);
    rewind(f);
    unsigned char le[8];
    if (fsz < 8 || fread(le, 1, 8, f) != 8) return fail("shorter than its length field");
    unsigned long long hn = 0;
    for (int b = 0; b < 8; ++b) hn |= (unsigned long long)le[b] << (8 * b);
    if (hn > (unsigned long long)fsz - 8 || hn > (100ull << 20)) return fail("header length out of range");
    std::string hdr(hn, '\0');
    if (hn && fread(&hdr[0], 1, hn, f) != hn) return fail("truncated header");
    uce_artifact* a = new uce_artifact();
    a->f = f; a->data_start = 8 + hn; a->file_size = (unsigned long long)fsz;
    Cur c{hdr.data(), hdr.data() + hdr.size()};
    auto bad = [&](const char* why) { delete a; return fail(why); };
    if (!eat(c, '{')) return bad("header is not a JSON object");
    if (!eat(c, '}')) {
        do {
            std::string key;
            if (!parse_string(c, key) || !eat(c, ':')) return bad("malformed key");
            if (key == "__metadata__") { if (!skip_value(c)) return bad("malformed metadata"); continue; }
            Entry e; e.name = key;
            bool have_off = false;
            if (!eat(c, '{')) return bad("tensor entry is not an object");
            do {
                std::string field;
                if (!parse_string(c, field) || !eat(c, ':')) return bad("malformed tensor entry");
                if (field == "dtype") { if (!parse_string(c, e.dtype)) return bad("malformed dtype"); }
                else if (field == "shape") {
                    if (!eat(c, '[')) return bad("malformed shape");
                    if (!eat(c, ']')) {
                        do { unsigned long long v; if (!parse_uint(c, v)) return bad("malformed shape"); e.shape.push_back((long)v); } while (eat(c, ','));
                        if (!eat(c, ']')) return bad("malformed shape");
                    }
                } else if (field == "data_offsets") {
                    if (!eat(c, '[') || !parse_uint(c, e.begin) || !eat(c, ',') || !parse_uint(c, e.end) || !eat(c, ']')) return bad("malformed data_offsets");
                    have_off = true;
                } else if (!skip_value(c)) return bad("malformed tensor entry");
            } while (eat(c, ','));
            if (!eat(c, '}')) return bad("malformed tensor entry");
            if (!have_off || e.dtype.empty() || e.end < e.begin || a->data_start + e.end > a->file_size) return bad("tensor data outside the file");
            a->entries.push_back(std::move(e));
        } while (eat(c, ','));
        if (!eat(c, '}')) return bad("unterminated header");
    }
    *out = a;
    return 0;
}

int uce_artifact_count(const uce_artifact* a) { return a ? (int)a->entries.size() : UCE_E_ARG; }

int uce_artifact_entry(const uce_artifact* a, int i, const char** name, const char** dtype, int* ndim, long shape[8]) {
    if (!a || i < 0 || i >= (int)a->entries.size()) { set_error("uce_artifact_entry: index out of range"); return UCE_E_ARG; }
    const Entry& e = a->entries[i];
    if (e.shape.size() > 8) { set_error("tensor '%s' has more than 8 dimensions", e.name.c_str()); return UCE_E_STATE; }
    if (name) *name = e.name.c_str();
    if (dtype) *dtype = e.dtype.c_str();
    if (ndim) *ndim = (int)e.shape.size();
    if (shape) for (size_t d = 0; d < e.shape.size(); ++d) shape[d] = e.shape[d];
    return 0;
}

int uce_artifact_read_f32(uce_artifact* a, int i, float* dst, size_t cap_elems) {
    if (!a || !dst || i < 0 || i >= (int)a->entries.size()) { set_error("uce_artifact_read_f32: bad argument"); return UCE_E_ARG; }
    const Entry& e = a->entries[i];
    if (e.dtype != "F32") { set_error("tensor '%s' is %s, the UCE artifact holds F32 (uce_sd_erase.py:117)", e.name.c_str(), e.dtype.c_str()); return UCE_E_STATE; }
    unsigned long long n = 1;
    for (long d : e.shape) n *= (unsigned long long)d;
    if (n * 4 != e.end - e.begin) { set_error("tensor '%s': shape and data_offsets disagree", e.name.c_str()); return UCE_E_STATE; }
    if (n > cap_elems) { set_error("tensor '%s' needs %llu floats, buffer holds %zu", e.name.c_str(), n, cap_elems); return UCE_E_ARG; }
    if (fseek(a->f, (long)(a->data_start + e.begin), SEEK_SET) != 0 || fread(dst, 4, (size_t)n, a->f) != (size_t)n) {
        set_error("read of tensor '%s' failed", e.name.c_str()); return UCE_E_STATE;
    }
    return 0;
}

int uce_artifact_close(uce_artifact* a) {
    if (!a) return 0;
    if (a->f) fclose(a->f);
    delete a;
    return 0;
}

}  // extern "C"
