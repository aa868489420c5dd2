// Parser for the consensus-specs trusted-setup file (presets/mainnet/trusted_setups/trusted_setup_4096.json):
//   {"g1_monomial": ["0x..96 hex..", x4096], "g1_lagrange": [...], "g2_monomial": ["0x..192 hex..", x65]}
// Replaces serde_json + hex decoding of TrustedSetupJSON (crates/trusted_setup/src/lib.rs:23-37,66-73) and the length / prefix
// rules of deserialize_g1_points / deserialize_g2_points (crates/serialization/src/trusted_setup.rs): every string starts with
// "0x" and decodes to exactly 48 (G1) or 96 (G2) bytes.  Where the reference panics, this returns an error.  Keys the reference's
// struct does not name (g1_lagrange) are skipped whatever their value is.  The points themselves are validated later: G1 on the
// device (decompression, curve equation, subgroup), G2 on the host (host_pairing.cpp).
#include <cstring>
#include "kzg_runtime.h"

namespace ekzg {
namespace {

struct Cursor {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    bool peek(char c) { ws(); return p < end && *p == c; }
};

// a JSON string without escape processing beyond skipping \x pairs (keys and hex strings never hold any)
bool read_string(Cursor& c, const char** s, size_t* n) {
    if (!c.eat('"')) return false;
    *s = c.p;
    while (c.p < c.end && *c.p != '"') {
        if (*c.p == '\\') c.p++;
        c.p++;
    }
    if (c.p >= c.end) return false;
    *n = (size_t)(c.p - *s);
    c.p++;
    return true;
}

bool skip_value(Cursor& c, int depth) {
    if (depth > 64) return false;
    c.ws();
    if (c.p >= c.end) return false;
    if (*c.p == '"') { const char* s; size_t n; return read_string(c, &s, &n); }
    if (*c.p == '[' || *c.p == '{') {
        const char open = *c.p, close = open == '[' ? ']' : '}';
        c.p++;
        if (c.eat(close)) return true;
        for (;;) {
            if (open == '{') { const char* s; size_t n; if (!read_string(c, &s, &n) || !c.eat(':')) return false; }
            if (!skip_value(c, depth + 1)) return false;
            if (c.eat(',')) continue;
            return c.eat(close);
        }
    }
    const char* start = c.p;   // number / true / false / null
    while (c.p < c.end && *c.p != ',' && *c.p != ']' && *c.p != '}' && *c.p != ' ' && *c.p != '\n' && *c.p != '\r' && *c.p != '\t') c.p++;
    return c.p > start;
}

int hex_nibble(char ch) {
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
    return -1;
}

Status read_points(Cursor& c, const char* key, size_t point_bytes, std::vector<uint8_t>* out) {
    out->clear();
    if (!c.eat('[')) return Status::Error(std::string("trusted setup JSON: ") + key + " is not an array");
    if (c.eat(']')) return Status::Ok();
    for (size_t idx = 0;; idx++) {
        const char* s; size_t n;
        if (!read_string(c, &s, &n)) return Status::Error(std::string("trusted setup JSON: ") + key + "[" + std::to_string(idx) + "] is not a string");
        if (n < 2 || s[0] != '0' || s[1] != 'x') return Status::Error(std::string("trusted setup JSON: ") + key + "[" + std::to_string(idx) + "] does not start with 0x");
        if (n - 2 != 2 * point_bytes)
            return Status::Error(std::string("trusted setup JSON: ") + key + "[" + std::to_string(idx) + "] holds " + std::to_string((n - 2) / 2) + " bytes, expected " + std::to_string(point_bytes));
        for (size_t i = 0; i < point_bytes; i++) {
            const int hi = hex_nibble(s[2 + 2 * i]), lo = hex_nibble(s[3 + 2 * i]);
            if (hi < 0 || lo < 0) return Status::Error(std::string("trusted setup JSON: ") + key + "[" + std::to_string(idx) + "] is not hexadecimal");
            out->push_back((uint8_t)(hi << 4 | lo));
        }
        if (c.eat(',')) continue;
        if (c.eat(']')) return Status::Ok();
        return Status::Error(std::string("trusted setup JSON: malformed array ") + key);
    }
}

}  // namespace

Status parse_trusted_setup_json(const char* json, size_t len, SetupBytes* out) {
    if (!json) return Status::Error("trusted setup JSON: null pointer");
    Cursor c{json, json + len};
    if (!c.eat('{')) return Status::Error("trusted setup JSON: not an object");
    bool have_g1 = false, have_g2 = false;
    if (!c.eat('}')) {
        for (;;) {
            const char* k; size_t kn;
            if (!read_string(c, &k, &kn) || !c.eat(':')) return Status::Error("trusted setup JSON: malformed object");
            if (kn == 11 && !memcmp(k, "g1_monomial", 11)) {
                Status s = read_points(c, "g1_monomial", 48, &out->g1_monomial);
                if (!s.ok) return s;
                have_g1 = true;
            } else if (kn == 11 && !memcmp(k, "g2_monomial", 11)) {
                Status s = read_points(c, "g2_monomial", 96, &out->g2_monomial);
                if (!s.ok) return s;
                have_g2 = true;
            } else if (!skip_value(c, 0)) {
                return Status::Error("trusted setup JSON: malformed value for key " + std::string(k, kn));
            }
            if (c.eat(',')) continue;
            if (c.eat('}')) break;
            return Status::Error("trusted setup JSON: malformed object");
        }
    }
    c.ws();
    if (c.p != c.end && !(c.p + 1 == c.end && *c.p == '\0')) return Status::Error("trusted setup JSON: trailing characters");
    if (!have_g1) return Status::Error("trusted setup JSON: missing field g1_monomial");
    if (!have_g2) return Status::Error("trusted setup JSON: missing field g2_monomial");
    return Status::Ok();
}

}  // namespace ekzg
