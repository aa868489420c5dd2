// see host_sha256.h
#include "host_sha256.h"
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ekzg {
namespace host {

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void blocks_portable(uint32_t* h, const uint8_t* p, size_t nblocks) {
    for (; nblocks; nblocks--, p += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
}

#if defined(__x86_64__)
// Four rounds per sha256rnds2 pair; the message schedule of rounds 4r..4r+3 (r >= 4) is
// msg2(msg1(W[r-4], W[r-3]) + (W[r-2] : W[r-1] >> 32 bits), W[r-1]).
__attribute__((target("sha,sse4.1,ssse3"))) static void blocks_sha_ni(uint32_t* h, const uint8_t* p, size_t nblocks) {
    const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);     // DCBA
    __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);     // HGFE
    tmp = _mm_shuffle_epi32(tmp, 0xB1);                       // CDAB
    st1 = _mm_shuffle_epi32(st1, 0x1B);                       // EFGH
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);               // ABEF
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);                    // CDGH
    for (; nblocks; nblocks--, p += 64) {
        const __m128i save0 = st0, save1 = st1;
        __m128i m[4];
        for (int r = 0; r < 16; r++) {
            if (r < 4) {
                m[r] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * r)), bswap);
            } else {
                __m128i t = _mm_sha256msg1_epu32(m[r & 3], m[(r + 1) & 3]);
                t = _mm_add_epi32(t, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
                m[r & 3] = _mm_sha256msg2_epu32(t, m[(r + 3) & 3]);
            }
            __m128i msg = _mm_add_epi32(m[r & 3], _mm_loadu_si128((const __m128i*)&K256[4 * r]));
            st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
            msg = _mm_shuffle_epi32(msg, 0x0E);
            st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
        }
        st0 = _mm_add_epi32(st0, save0);
        st1 = _mm_add_epi32(st1, save1);
    }
    tmp = _mm_shuffle_epi32(st0, 0x1B);                       // FEBA
    st1 = _mm_shuffle_epi32(st1, 0xB1);                       // DCHG
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);                    // DCBA
    st1 = _mm_alignr_epi8(st1, tmp, 8);                       // HGFE
    _mm_storeu_si128((__m128i*)&h[0], st0);
    _mm_storeu_si128((__m128i*)&h[4], st1);
}
#endif

bool sha256_uses_sha_ni() {
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3");
    return ok;
#else
    return false;
#endif
}

static void blocks(uint32_t* h, const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    if (sha256_uses_sha_ni()) { blocks_sha_ni(h, p, n); return; }
#endif
    blocks_portable(h, p, n);
}

Sha256Stream::Sha256Stream() : buflen(0), total(0) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof iv);
}

void Sha256Stream::update(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    total += n;
    if (buflen) {
        size_t take = 64 - buflen < n ? 64 - buflen : n;
        memcpy(buf + buflen, p, take);
        buflen += (uint32_t)take; p += take; n -= take;
        if (buflen == 64) { blocks(h, buf, 1); buflen = 0; }
    }
    if (n >= 64) { blocks(h, p, n / 64); p += (n / 64) * 64; n %= 64; }
    if (n) { memcpy(buf, p, n); buflen = (uint32_t)n; }
}

void Sha256Stream::final(uint8_t out[32]) {
    uint64_t bits = total * 8;
    buf[buflen++] = 0x80;
    if (buflen > 56) { memset(buf + buflen, 0, 64 - buflen); blocks(h, buf, 1); buflen = 0; }
    memset(buf + buflen, 0, 56 - buflen);
    for (int i = 0; i < 8; i++) buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    blocks(h, buf, 1);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

}  // namespace host
}  // namespace ekzg

// test hook for tests/test_host_sha256.py (no GPU needed)
extern "C" void eth_kzg_b200_debug_sha256(const uint8_t* data, uint64_t n, uint64_t split, int force_portable, uint8_t out[32]) {
    ekzg::host::Sha256Stream s;
    if (force_portable) {
        // portable path, fed in one piece after the split prefix
        uint32_t h[8];
        memcpy(h, s.h, sizeof h);
        uint64_t full = n / 64;
        ekzg::host::blocks_portable(h, data, full);
        memcpy(s.h, h, sizeof h);
        s.total = full * 64;
        s.update(data + full * 64, n - full * 64);
        s.total = n;
    } else {
        if (split > n) split = n;
        s.update(data, split);
        s.update(data + split, n - split);
    }
    s.final(out);
}
