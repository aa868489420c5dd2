// Host SHA-256 for the one strictly sequential piece of the verifiers: the Fiat-Shamir transcript of a batch
// (kzg_multi_open/src/fk20/verifier.rs:269-328 hashes every commitment, cell and proof -- 34.6 MB for 128 blobs x 128
// cells -- into ONE digest, so it cannot be spread over GPU threads).  Uses the x86 SHA extensions when the CPU has
// them (the reference's sha2 crate does the same), else portable C.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ekzg {
namespace host {

struct Sha256Stream {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t buflen;
    uint64_t total;
    Sha256Stream();
    void update(const void* data, size_t n);
    void final(uint8_t out[32]);
};

bool sha256_uses_sha_ni();

}  // namespace host
}  // namespace ekzg
