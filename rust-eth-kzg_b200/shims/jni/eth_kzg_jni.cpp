// JNI shim: the 12 native methods of ethereum.cryptography.LibEthKZG on top of the C ABI (include/c_eth_kzg.h).
//
// Replaces bindings/java/rust_code/src/lib.rs (a Rust cdylib, `java_eth_kzg`, that links the reference's c_eth_kzg crate as an rlib):
// same exported symbols (Java_ethereum_cryptography_LibEthKZG_*), same argument marshalling, same result objects
// (ethereum/cryptography/CellsAndProofs([[B[[B)V, ethereum/cryptography/Cells([[B)V, byte[][]{proof, y}), same failure behaviour:
// java.lang.IllegalArgumentException("function <name> has thrown an exception, with reason: <reason>") and a null / false return
// (lib.rs:509-524), wrong-length arrays reported as "<name> is not the correct size. expected: <n>\ngot: <m>" (errors.rs, lib.rs:530-541).
// The Java sources of the reference (bindings/java/java_code) work unchanged on top of libjava_eth_kzg.so built from this file.
//
// Built against the JDK's <jni.h> when there is one; this image has no JDK, so the build falls back to jni_min.h (the same
// function-table layout) and the shim is exercised through a mock JNIEnv (tests/jni/jni_mock.cpp, tests/test_jni_shim.py) -- it has never run inside a JVM.
#if defined(__has_include)
#if __has_include(<jni.h>)
#include <jni.h>
#endif
#endif
#ifndef JNI_VERSION_1_8
#include "jni_min.h"
#endif
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/c_eth_kzg.h"

#ifdef EKZG_JNI_MIN_HEADER
#define JT(env) (*(env))
#else
#define JT(env) ((env)->functions)
#endif

namespace {

constexpr size_t BLOB = 131072, CELL = 2048, G1 = 48, FR = 32, NCELLS = 128;

// a failure on its way to Java: either a JNI call failed with an exception already pending (rethrown as is), or a reason for
// throw_on_error
struct Failure {
    bool pending = false;
    std::string reason;
};

struct Shim {
    JNIEnv* env;
    const char* fn;
    Failure fail;
    bool failed = false;

    bool set(const std::string& reason) { failed = true; fail.reason = reason; return false; }
    bool jni_failed(const char* what) {
        if (JT(env)->ExceptionCheck(env)) { failed = true; fail.pending = true; return true; }
        (void)what;
        return false;
    }

    // env.convert_byte_array + slice_to_array_ref::<N>
    bool bytes(jbyteArray a, size_t expect, const char* name, std::vector<uint8_t>* out) {
        if (!a) return set(std::string("NullPtr(\"") + name + "\")");
        const jsize n = JT(env)->GetArrayLength(env, a);
        if (jni_failed("GetArrayLength")) return false;
        out->resize((size_t)n);
        if (n) JT(env)->GetByteArrayRegion(env, a, 0, n, reinterpret_cast<jbyte*>(out->data()));
        if (jni_failed("GetByteArrayRegion")) return false;
        if ((size_t)n != expect)
            return set(std::string(name) + " is not the correct size. expected: " + std::to_string(expect) + "\ngot: " + std::to_string(n));
        return true;
    }
    // jobject_array_to_2d_byte_array + slice_to_array_ref per element: contiguous storage + the pointer array the C ABI wants
    bool bytes2d(jobjectArray a, size_t expect, const char* name, std::vector<uint8_t>* flat, std::vector<const uint8_t*>* ptrs) {
        if (!a) return set(std::string("NullPtr(\"") + name + "\")");
        const jsize n = JT(env)->GetArrayLength(env, a);
        if (jni_failed("GetArrayLength")) return false;
        flat->assign((size_t)n * expect, 0);
        ptrs->clear();
        for (jsize i = 0; i < n; i++) {
            jbyteArray e = (jbyteArray)JT(env)->GetObjectArrayElement(env, a, i);
            if (jni_failed("GetObjectArrayElement")) return false;
            if (!e) return set(std::string("NullPtr(\"") + name + "\")");
            const jsize len = JT(env)->GetArrayLength(env, e);
            if ((size_t)len != expect) {
                JT(env)->DeleteLocalRef(env, e);
                return set(std::string(name) + " is not the correct size. expected: " + std::to_string(expect) + "\ngot: " + std::to_string(len));
            }
            JT(env)->GetByteArrayRegion(env, e, 0, len, reinterpret_cast<jbyte*>(flat->data() + (size_t)i * expect));
            JT(env)->DeleteLocalRef(env, e);          // 16 384-element batches must not exhaust the local reference table
            if (jni_failed("GetByteArrayRegion")) return false;
        }
        for (jsize i = 0; i < n; i++) ptrs->push_back(flat->data() + (size_t)i * expect);
        return true;
    }
    // jlongarray_to_vec_u64
    bool longs(jlongArray a, std::vector<uint64_t>* out) {
        if (!a) return set("NullPtr(\"cell indices\")");
        const jsize n = JT(env)->GetArrayLength(env, a);
        if (jni_failed("GetArrayLength")) return false;
        std::vector<jlong> tmp((size_t)n);
        if (n) JT(env)->GetLongArrayRegion(env, a, 0, n, tmp.data());
        if (jni_failed("GetLongArrayRegion")) return false;
        out->resize((size_t)n);
        for (jsize i = 0; i < n; i++) (*out)[i] = (uint64_t)tmp[i];
        return true;
    }
    // CResult -> Ok / Error::Cryptography
    bool ok(CResult r) {
        if (r.status == Ok) return true;
        set(r.error_msg ? r.error_msg : "error");
        eth_kzg_free_error_message(r.error_msg);
        return false;
    }

    jbyteArray new_bytes(const uint8_t* p, size_t n) {
        jbyteArray a = JT(env)->NewByteArray(env, (jsize)n);
        if (!a || jni_failed("NewByteArray")) { failed = true; fail.pending = true; return nullptr; }
        JT(env)->SetByteArrayRegion(env, a, 0, (jsize)n, reinterpret_cast<const jbyte*>(p));
        return a;
    }
    // byte[count][item] from contiguous storage
    jobjectArray new_bytes2d(const uint8_t* p, size_t count, size_t item) {
        jclass cls = JT(env)->FindClass(env, "[B");
        if (!cls || jni_failed("FindClass")) { failed = true; fail.pending = true; return nullptr; }
        jobjectArray arr = JT(env)->NewObjectArray(env, (jsize)count, cls, nullptr);
        if (!arr || jni_failed("NewObjectArray")) { failed = true; fail.pending = true; return nullptr; }
        for (size_t i = 0; i < count; i++) {
            jbyteArray e = new_bytes(p + i * item, item);
            if (!e) return nullptr;
            JT(env)->SetObjectArrayElement(env, arr, (jsize)i, e);
            JT(env)->DeleteLocalRef(env, e);
            if (jni_failed("SetObjectArrayElement")) return nullptr;
        }
        return arr;
    }
    // new <cls>(args...) with the constructor signature `sig`
    jobject construct(const char* cls_name, const char* sig, const jvalue* args) {
        jclass cls = JT(env)->FindClass(env, cls_name);
        if (!cls || jni_failed("FindClass")) { failed = true; fail.pending = true; return nullptr; }
        jmethodID ctor = JT(env)->GetMethodID(env, cls, "<init>", sig);
        if (!ctor || jni_failed("GetMethodID")) { failed = true; fail.pending = true; return nullptr; }
        jobject o = JT(env)->NewObjectA(env, cls, ctor, args);
        if (!o || jni_failed("NewObjectA")) { failed = true; fail.pending = true; return nullptr; }
        return o;
    }

    // throw_on_error (lib.rs:509-524)
    void raise() {
        if (fail.pending && JT(env)->ExceptionCheck(env)) return;   // the JVM's own exception travels on
        const std::string msg = std::string("function ") + fn + " has thrown an exception, with reason: " + (fail.reason.empty() ? "Jni error" : fail.reason);
        jclass cls = JT(env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (cls) JT(env)->ThrowNew(env, cls, msg.c_str());
    }
};

const DASContext* ctx_of(jlong p) { return reinterpret_cast<const DASContext*>(static_cast<intptr_t>(p)); }

jobject cells_and_proofs(Shim& s, const std::vector<uint8_t>& cells, const std::vector<uint8_t>& proofs) {
    jobjectArray c = s.new_bytes2d(cells.data(), NCELLS, CELL);
    if (!c) return nullptr;
    jobjectArray p = s.new_bytes2d(proofs.data(), NCELLS, G1);
    if (!p) return nullptr;
    jvalue args[2];
    args[0].l = c;
    args[1].l = p;
    return s.construct("ethereum/cryptography/CellsAndProofs", "([[B[[B)V", args);
}

// the C ABI writes through 128 pointers (pointer_utils.rs:53-62); here they all point into one block
void out_ptrs(std::vector<uint8_t>& block, size_t item, std::vector<uint8_t*>* ptrs) {
    block.assign(NCELLS * item, 0);
    ptrs->clear();
    for (size_t i = 0; i < NCELLS; i++) ptrs->push_back(block.data() + i * item);
}

}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_ethereum_cryptography_LibEthKZG_DASContextNew(JNIEnv*, jclass, jboolean use_precomp) {
    return (jlong) reinterpret_cast<intptr_t>(eth_kzg_das_context_new(use_precomp != 0));
}

JNIEXPORT void JNICALL Java_ethereum_cryptography_LibEthKZG_DASContextDestroy(JNIEnv*, jclass, jlong ctx_ptr) {
    eth_kzg_das_context_free(const_cast<DASContext*>(ctx_of(ctx_ptr)));
}

JNIEXPORT jobject JNICALL Java_ethereum_cryptography_LibEthKZG_computeCellsAndKZGProofs(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob) {
    Shim s{env, "computeCellsAndKZGProofs"};
    std::vector<uint8_t> b, cells, proofs;
    std::vector<uint8_t*> cp, pp;
    jobject out = nullptr;
    if (s.bytes(blob, BLOB, "blob", &b)) {
        out_ptrs(cells, CELL, &cp);
        out_ptrs(proofs, G1, &pp);
        if (s.ok(eth_kzg_compute_cells_and_kzg_proofs(ctx_of(ctx_ptr), b.data(), cp.data(), pp.data()))) out = cells_and_proofs(s, cells, proofs);
    }
    if (s.failed) { s.raise(); return nullptr; }
    return out;
}

JNIEXPORT jobject JNICALL Java_ethereum_cryptography_LibEthKZG_computeCells(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob) {
    Shim s{env, "computeCells"};
    std::vector<uint8_t> b, cells;
    std::vector<uint8_t*> cp;
    jobject out = nullptr;
    if (s.bytes(blob, BLOB, "blob", &b)) {
        out_ptrs(cells, CELL, &cp);
        if (s.ok(eth_kzg_compute_cells(ctx_of(ctx_ptr), b.data(), cp.data()))) {
            jobjectArray c = s.new_bytes2d(cells.data(), NCELLS, CELL);
            if (c) {
                jvalue args[1];
                args[0].l = c;
                out = s.construct("ethereum/cryptography/Cells", "([[B)V", args);
            }
        }
    }
    if (s.failed) { s.raise(); return nullptr; }
    return out;
}

JNIEXPORT jbyteArray JNICALL Java_ethereum_cryptography_LibEthKZG_blobToKZGCommitment(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob) {
    Shim s{env, "blobToKZGCommitment"};
    std::vector<uint8_t> b;
    uint8_t c[G1];
    jbyteArray out = nullptr;
    if (s.bytes(blob, BLOB, "blob", &b) && s.ok(eth_kzg_blob_to_kzg_commitment(ctx_of(ctx_ptr), b.data(), c))) out = s.new_bytes(c, G1);
    if (s.failed) { s.raise(); return nullptr; }
    return out;
}

JNIEXPORT jboolean JNICALL Java_ethereum_cryptography_LibEthKZG_verifyCellKZGProofBatch(JNIEnv* env, jclass, jlong ctx_ptr, jobjectArray commitments,
                                                                                        jlongArray cell_indices, jobjectArray cells, jobjectArray proofs) {
    Shim s{env, "verifyCellKZGProofBatch"};
    std::vector<uint8_t> cm, ce, pr;
    std::vector<const uint8_t*> cmp, cep, prp;
    std::vector<uint64_t> idx;
    bool verified = false;
    // (the reference converts all four arrays, then checks cell, commitment and proof sizes in that order -- lib.rs:137-156; here
    //  every array is checked as it is converted, so only the message differs when several arguments are wrong at once)
    if (s.bytes2d(commitments, G1, "commitment", &cm, &cmp) && s.longs(cell_indices, &idx) && s.bytes2d(cells, CELL, "cell", &ce, &cep) &&
        s.bytes2d(proofs, G1, "proof", &pr, &prp))
        s.ok(eth_kzg_verify_cell_kzg_proof_batch(ctx_of(ctx_ptr), cmp.size(), cmp.data(), idx.size(), idx.data(), cep.size(), cep.data(), prp.size(), prp.data(),
                                                 &verified));
    if (s.failed) { s.raise(); return JNI_FALSE; }
    return verified ? JNI_TRUE : JNI_FALSE;
}

JNIEXPORT jobject JNICALL Java_ethereum_cryptography_LibEthKZG_recoverCellsAndKZGProofs(JNIEnv* env, jclass, jlong ctx_ptr, jlongArray cell_ids,
                                                                                        jobjectArray cells) {
    Shim s{env, "recoverCellsAndKZGProofs"};
    std::vector<uint8_t> ce, oc, op;
    std::vector<const uint8_t*> cep;
    std::vector<uint8_t*> ocp, opp;
    std::vector<uint64_t> idx;
    jobject out = nullptr;
    if (s.longs(cell_ids, &idx) && s.bytes2d(cells, CELL, "cell", &ce, &cep)) {
        out_ptrs(oc, CELL, &ocp);
        out_ptrs(op, G1, &opp);
        if (s.ok(eth_kzg_recover_cells_and_proofs(ctx_of(ctx_ptr), cep.size(), cep.data(), idx.size(), idx.data(), ocp.data(), opp.data())))
            out = cells_and_proofs(s, oc, op);
    }
    if (s.failed) { s.raise(); return nullptr; }
    return out;
}

JNIEXPORT jobjectArray JNICALL Java_ethereum_cryptography_LibEthKZG_computeKzgProof(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob, jbyteArray z) {
    Shim s{env, "computeKzgProof"};
    std::vector<uint8_t> b, zz;
    uint8_t out[G1 + FR];   // proof, then y
    jobjectArray res = nullptr;
    if (s.bytes(blob, BLOB, "blob", &b) && s.bytes(z, FR, "z", &zz) && s.ok(eth_kzg_compute_kzg_proof(ctx_of(ctx_ptr), b.data(), zz.data(), out, out + G1))) {
        jclass cls = JT(env)->FindClass(env, "[B");
        if (cls) res = JT(env)->NewObjectArray(env, 2, cls, nullptr);
        jbyteArray p = res ? s.new_bytes(out, G1) : nullptr;
        jbyteArray y = p ? s.new_bytes(out + G1, FR) : nullptr;
        if (y) {
            JT(env)->SetObjectArrayElement(env, res, 0, p);
            JT(env)->SetObjectArrayElement(env, res, 1, y);
        } else {
            s.failed = true;
            s.fail.pending = true;
        }
    }
    if (s.failed) { s.raise(); return nullptr; }
    return res;
}

JNIEXPORT jbyteArray JNICALL Java_ethereum_cryptography_LibEthKZG_computeBlobKzgProof(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob, jbyteArray commitment) {
    Shim s{env, "computeBlobKzgProof"};
    std::vector<uint8_t> b, c;
    uint8_t proof[G1];
    jbyteArray out = nullptr;
    if (s.bytes(blob, BLOB, "blob", &b) && s.bytes(commitment, G1, "commitment", &c) &&
        s.ok(eth_kzg_compute_blob_kzg_proof(ctx_of(ctx_ptr), b.data(), c.data(), proof)))
        out = s.new_bytes(proof, G1);
    if (s.failed) { s.raise(); return nullptr; }
    return out;
}

JNIEXPORT jboolean JNICALL Java_ethereum_cryptography_LibEthKZG_verifyKzgProof(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray commitment, jbyteArray z,
                                                                               jbyteArray y, jbyteArray proof) {
    Shim s{env, "verifyKzgProof"};
    std::vector<uint8_t> c, zz, yy, p;
    bool verified = false;
    if (s.bytes(commitment, G1, "commitment", &c) && s.bytes(z, FR, "z", &zz) && s.bytes(y, FR, "y", &yy) && s.bytes(proof, G1, "proof", &p))
        s.ok(eth_kzg_verify_kzg_proof(ctx_of(ctx_ptr), c.data(), zz.data(), yy.data(), p.data(), &verified));
    if (s.failed) { s.raise(); return JNI_FALSE; }
    return verified ? JNI_TRUE : JNI_FALSE;
}

JNIEXPORT jboolean JNICALL Java_ethereum_cryptography_LibEthKZG_verifyBlobKzgProof(JNIEnv* env, jclass, jlong ctx_ptr, jbyteArray blob, jbyteArray commitment,
                                                                                   jbyteArray proof) {
    Shim s{env, "verifyBlobKzgProof"};
    std::vector<uint8_t> b, c, p;
    bool verified = false;
    if (s.bytes(blob, BLOB, "blob", &b) && s.bytes(commitment, G1, "commitment", &c) && s.bytes(proof, G1, "proof", &p))
        s.ok(eth_kzg_verify_blob_kzg_proof(ctx_of(ctx_ptr), b.data(), c.data(), p.data(), &verified));
    if (s.failed) { s.raise(); return JNI_FALSE; }
    return verified ? JNI_TRUE : JNI_FALSE;
}

JNIEXPORT jboolean JNICALL Java_ethereum_cryptography_LibEthKZG_verifyBlobKzgProofBatch(JNIEnv* env, jclass, jlong ctx_ptr, jobjectArray blobs,
                                                                                        jobjectArray commitments, jobjectArray proofs) {
    Shim s{env, "verifyBlobKzgProofBatch"};
    std::vector<uint8_t> b, c, p;
    std::vector<const uint8_t*> bp, cp, pp;
    bool verified = false;
    if (s.bytes2d(blobs, BLOB, "blob", &b, &bp) && s.bytes2d(commitments, G1, "commitment", &c, &cp) && s.bytes2d(proofs, G1, "proof", &p, &pp))
        s.ok(eth_kzg_verify_blob_kzg_proof_batch(ctx_of(ctx_ptr), bp.size(), bp.data(), cp.size(), cp.data(), pp.size(), pp.data(), &verified));
    if (s.failed) { s.raise(); return JNI_FALSE; }
    return verified ? JNI_TRUE : JNI_FALSE;
}

}  // extern "C"
