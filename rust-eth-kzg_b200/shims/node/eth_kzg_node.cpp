// Node-API addon: the reference's Node binding (bindings/node/src/lib.rs, a napi-rs cdylib over the Rust crate) on top of the C ABI
// (include/c_eth_kzg.h).  Same exports as bindings/node/index.d.ts: the six size constants, class CellsAndProofs {cells, proofs},
// class DasContextJs with constructor() (usePrecomp = true, lib.rs:49-60), static create({usePrecomp}) and the ten methods in their
// synchronous and async* (Promise) forms.  Argument rules of the reference: Uint8Array arguments of the exact size
// ("<name> must have size <N>, found size <len>", lib.rs:440-453), cell indices as number | bigint (lib.rs:427-436), failures as
// Error("failed to compute <function>: <reason>"), an invalid proof is `false`, not an exception (lib.rs:229-238).
//
// async* methods: the reference runs the same body on napi-rs' tokio pool; here the arguments are copied on the JS thread, the C-ABI
// call runs on the libuv pool (napi_async_work) -- where concurrent single-blob calls meet in the library's coalescing queues and
// share GPU batches -- and the result objects are built back on the JS thread.
//
// Built against node's <node_api.h> when there is one; this image has no node, so the build falls back to node_api_min.h and the addon
// is exercised through a mock Node-API (tests/napi/napi_mock.cpp, tests/test_node_shim.py) -- it has never been loaded by node.
#if defined(__has_include)
#if __has_include(<node_api.h>)
#include <node_api.h>
#endif
#endif
#ifndef NAPI_AUTO_LENGTH
#include "node_api_min.h"
#endif
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/c_eth_kzg.h"

namespace {

constexpr size_t BLOB = 131072, CELL = 2048, G1 = 48, FR = 32, NCELLS = 128;

enum Kind { Commit, CellsProofs, Cells, Recover, VerifyCellBatch, KzgProof, BlobProof, VerifyKzg, VerifyBlob, VerifyBlobBatch, N_KINDS };
const char* const FN_NAME[N_KINDS] = {"blob_to_kzg_commitment", "compute_cells_and_kzg_proofs", "compute_cells", "recover_cells_and_kzg_proofs",
                                      "verify_cell_kzg_proof_batch", "compute_kzg_proof", "compute_blob_kzg_proof", "verify_kzg_proof",
                                      "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch"};

struct Addon {             // per-environment state (napi_set_instance_data)
    napi_ref cells_and_proofs_ctor = nullptr;
    napi_ref context_ctor = nullptr;
};

struct Native {            // what a DasContextJs wraps
    DASContext* ctx = nullptr;
};

// one call: inputs copied out of the JS values, outputs as flat bytes
struct Job {
    Kind kind;
    const DASContext* ctx = nullptr;
    std::vector<uint8_t> in[4];                 // flat copies of the (arrays of) Uint8Array arguments, in argument order
    std::vector<const uint8_t*> ptrs[4];
    std::vector<uint64_t> indices;
    std::vector<uint8_t> out_a, out_b;          // cells / commitment / proof, proofs / y
    bool verified = false;
    bool ok = true;
    std::string error;
    // async only
    napi_deferred deferred = nullptr;
    napi_async_work work = nullptr;
};

bool throw_msg(napi_env env, const std::string& m) {
    napi_throw_error(env, nullptr, m.c_str());
    return false;
}

std::string size_error(const char* name, size_t want, size_t got) {   // slice_to_array_ref (lib.rs:440-453)
    return std::string(name) + " must have size " + std::to_string(want) + ", found size " + std::to_string(got) + "\n err:could not convert slice to array";
}

bool get_bytes(napi_env env, napi_value v, const char* name, size_t want, std::vector<uint8_t>* out, bool append) {
    bool is_ta = false;
    napi_typedarray_type type;
    size_t len = 0;
    void* data = nullptr;
    if (napi_is_typedarray(env, v, &is_ta) != napi_ok || !is_ta || napi_get_typedarray_info(env, v, &type, &len, &data, nullptr, nullptr) != napi_ok ||
        type != napi_uint8_array)
        return throw_msg(env, std::string(name) + " must be a Uint8Array");
    if (len != want) return throw_msg(env, size_error(name, want, len));
    const uint8_t* p = static_cast<const uint8_t*>(data);
    if (append) out->insert(out->end(), p, p + len);
    else out->assign(p, p + len);
    return true;
}

bool get_bytes_array(napi_env env, napi_value v, const char* name, size_t want, std::vector<uint8_t>* flat, std::vector<const uint8_t*>* ptrs) {
    bool is_arr = false;
    uint32_t n = 0;
    if (napi_is_array(env, v, &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, v, &n) != napi_ok)
        return throw_msg(env, std::string(name) + "s must be an array of Uint8Array");
    flat->clear();
    flat->reserve((size_t)n * want);
    for (uint32_t i = 0; i < n; i++) {
        napi_value e;
        if (napi_get_element(env, v, i, &e) != napi_ok) return throw_msg(env, std::string(name) + "s must be an array of Uint8Array");
        if (!get_bytes(env, e, name, want, flat, true)) return false;
    }
    ptrs->clear();
    for (uint32_t i = 0; i < n; i++) ptrs->push_back(flat->data() + (size_t)i * want);
    return true;
}

// Vec<Either<u32, BigInt>> -> u64 (u32_or_bigint_to_u64, lib.rs:427-436)
bool get_indices(napi_env env, napi_value v, std::vector<uint64_t>* out) {
    bool is_arr = false;
    uint32_t n = 0;
    if (napi_is_array(env, v, &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, v, &n) != napi_ok)
        return throw_msg(env, "cell indices must be an array of number | bigint");
    out->clear();
    for (uint32_t i = 0; i < n; i++) {
        napi_value e;
        napi_valuetype t;
        if (napi_get_element(env, v, i, &e) != napi_ok || napi_typeof(env, e, &t) != napi_ok) return throw_msg(env, "cell indices must be an array of number | bigint");
        if (t == napi_number) {
            uint32_t x = 0;
            if (napi_get_value_uint32(env, e, &x) != napi_ok) return throw_msg(env, "cell index is not a u32");
            out->push_back(x);
        } else if (t == napi_bigint) {
            uint64_t x = 0;
            bool lossless = true;
            if (napi_get_value_bigint_uint64(env, e, &x, &lossless) != napi_ok) return throw_msg(env, "cell index is not a bigint");
            out->push_back(x);   // the reference truncates a wider value to 64 bits as well (`value_u128 as u64`)
        } else {
            return throw_msg(env, "cell indices must be an array of number | bigint");
        }
    }
    return true;
}

// JS thread: arguments -> job.  false = an exception is pending.
bool parse(napi_env env, napi_callback_info info, Kind kind, Job* job) {
    size_t argc = 4;
    napi_value argv[4] = {nullptr, nullptr, nullptr, nullptr}, self;
    if (napi_get_cb_info(env, info, &argc, argv, &self, nullptr) != napi_ok) return throw_msg(env, "napi_get_cb_info failed");
    Native* nat = nullptr;
    if (napi_unwrap(env, self, reinterpret_cast<void**>(&nat)) != napi_ok || !nat || !nat->ctx) return throw_msg(env, "not a DasContextJs");
    job->kind = kind;
    job->ctx = nat->ctx;
    static const int want_args[N_KINDS] = {1, 1, 1, 2, 4, 2, 2, 4, 3, 3};
    if (argc < (size_t)want_args[kind]) return throw_msg(env, std::string(FN_NAME[kind]) + ": expected " + std::to_string(want_args[kind]) + " arguments");
    switch (kind) {
        case Commit: case CellsProofs: case Cells:
            return get_bytes(env, argv[0], "blob", BLOB, &job->in[0], false);
        case Recover:
            return get_indices(env, argv[0], &job->indices) && get_bytes_array(env, argv[1], "cell", CELL, &job->in[0], &job->ptrs[0]);
        case VerifyCellBatch:   // the reference converts the indices first, then checks commitments, cells, proofs (lib.rs:213-226)
            return get_indices(env, argv[1], &job->indices) && get_bytes_array(env, argv[0], "commitment", G1, &job->in[0], &job->ptrs[0]) &&
                   get_bytes_array(env, argv[2], "cell", CELL, &job->in[1], &job->ptrs[1]) && get_bytes_array(env, argv[3], "proof", G1, &job->in[2], &job->ptrs[2]);
        case KzgProof:
            return get_bytes(env, argv[0], "blob", BLOB, &job->in[0], false) && get_bytes(env, argv[1], "z", FR, &job->in[1], false);
        case BlobProof:
            return get_bytes(env, argv[0], "blob", BLOB, &job->in[0], false) && get_bytes(env, argv[1], "commitment", G1, &job->in[1], false);
        case VerifyKzg:
            return get_bytes(env, argv[0], "commitment", G1, &job->in[0], false) && get_bytes(env, argv[1], "z", FR, &job->in[1], false) &&
                   get_bytes(env, argv[2], "y", FR, &job->in[2], false) && get_bytes(env, argv[3], "proof", G1, &job->in[3], false);
        case VerifyBlob:
            return get_bytes(env, argv[0], "blob", BLOB, &job->in[0], false) && get_bytes(env, argv[1], "commitment", G1, &job->in[1], false) &&
                   get_bytes(env, argv[2], "proof", G1, &job->in[2], false);
        case VerifyBlobBatch:
            return get_bytes_array(env, argv[0], "blob", BLOB, &job->in[0], &job->ptrs[0]) && get_bytes_array(env, argv[1], "commitment", G1, &job->in[1], &job->ptrs[1]) &&
                   get_bytes_array(env, argv[2], "proof", G1, &job->in[2], &job->ptrs[2]);
        default:
            return throw_msg(env, "unknown method");
    }
}

// any thread: the C-ABI call
void run(Job* j) {
    CResult r{Ok, nullptr};
    std::vector<uint8_t*> pa, pb;
    auto outs = [&](std::vector<uint8_t>& block, size_t item, std::vector<uint8_t*>& ptrs) {
        block.assign(NCELLS * item, 0);
        for (size_t i = 0; i < NCELLS; i++) ptrs.push_back(block.data() + i * item);
    };
    switch (j->kind) {
        case Commit:
            j->out_a.assign(G1, 0);
            r = eth_kzg_blob_to_kzg_commitment(j->ctx, j->in[0].data(), j->out_a.data());
            break;
        case CellsProofs:
            outs(j->out_a, CELL, pa); outs(j->out_b, G1, pb);
            r = eth_kzg_compute_cells_and_kzg_proofs(j->ctx, j->in[0].data(), pa.data(), pb.data());
            break;
        case Cells:
            outs(j->out_a, CELL, pa);
            r = eth_kzg_compute_cells(j->ctx, j->in[0].data(), pa.data());
            break;
        case Recover:
            outs(j->out_a, CELL, pa); outs(j->out_b, G1, pb);
            r = eth_kzg_recover_cells_and_proofs(j->ctx, j->ptrs[0].size(), j->ptrs[0].data(), j->indices.size(), j->indices.data(), pa.data(), pb.data());
            break;
        case VerifyCellBatch:
            r = eth_kzg_verify_cell_kzg_proof_batch(j->ctx, j->ptrs[0].size(), j->ptrs[0].data(), j->indices.size(), j->indices.data(), j->ptrs[1].size(),
                                                    j->ptrs[1].data(), j->ptrs[2].size(), j->ptrs[2].data(), &j->verified);
            break;
        case KzgProof:
            j->out_a.assign(G1, 0); j->out_b.assign(FR, 0);
            r = eth_kzg_compute_kzg_proof(j->ctx, j->in[0].data(), j->in[1].data(), j->out_a.data(), j->out_b.data());
            break;
        case BlobProof:
            j->out_a.assign(G1, 0);
            r = eth_kzg_compute_blob_kzg_proof(j->ctx, j->in[0].data(), j->in[1].data(), j->out_a.data());
            break;
        case VerifyKzg:
            r = eth_kzg_verify_kzg_proof(j->ctx, j->in[0].data(), j->in[1].data(), j->in[2].data(), j->in[3].data(), &j->verified);
            break;
        case VerifyBlob:
            r = eth_kzg_verify_blob_kzg_proof(j->ctx, j->in[0].data(), j->in[1].data(), j->in[2].data(), &j->verified);
            break;
        case VerifyBlobBatch:
            r = eth_kzg_verify_blob_kzg_proof_batch(j->ctx, j->ptrs[0].size(), j->ptrs[0].data(), j->ptrs[1].size(), j->ptrs[1].data(), j->ptrs[2].size(),
                                                    j->ptrs[2].data(), &j->verified);
            break;
        default:
            break;
    }
    if (r.status != Ok) {
        j->ok = false;
        j->error = std::string("failed to compute ") + FN_NAME[j->kind] + ": " + (r.error_msg ? r.error_msg : "error");
        eth_kzg_free_error_message(r.error_msg);
    }
}

napi_value make_u8(napi_env env, const uint8_t* p, size_t n) {
    void* data = nullptr;
    napi_value ab, ta;
    if (napi_create_arraybuffer(env, n, &data, &ab) != napi_ok) return nullptr;
    memcpy(data, p, n);
    if (napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &ta) != napi_ok) return nullptr;
    return ta;
}
napi_value make_u8_array(napi_env env, const std::vector<uint8_t>& flat, size_t item) {
    napi_value arr;
    const size_t n = flat.size() / item;
    if (napi_create_array_with_length(env, n, &arr) != napi_ok) return nullptr;
    for (size_t i = 0; i < n; i++) {
        napi_value e = make_u8(env, flat.data() + i * item, item);
        if (!e || napi_set_element(env, arr, (uint32_t)i, e) != napi_ok) return nullptr;
    }
    return arr;
}

// JS thread: a finished, successful job -> its JS result (nullptr if Node-API itself failed)
napi_value to_js(napi_env env, const Job& j) {
    switch (j.kind) {
        case Commit: case BlobProof:
            return make_u8(env, j.out_a.data(), j.out_a.size());
        case Cells:
            return make_u8_array(env, j.out_a, CELL);
        case KzgProof: {   // [proof, y]
            napi_value arr, p = make_u8(env, j.out_a.data(), G1), y = make_u8(env, j.out_b.data(), FR);
            if (!p || !y || napi_create_array_with_length(env, 2, &arr) != napi_ok) return nullptr;
            napi_set_element(env, arr, 0, p);
            napi_set_element(env, arr, 1, y);
            return arr;
        }
        case CellsProofs: case Recover: {
            Addon* addon = nullptr;
            napi_value ctor, obj, cells = make_u8_array(env, j.out_a, CELL), proofs = make_u8_array(env, j.out_b, G1);
            if (!cells || !proofs || napi_get_instance_data(env, reinterpret_cast<void**>(&addon)) != napi_ok || !addon) return nullptr;
            if (napi_get_reference_value(env, addon->cells_and_proofs_ctor, &ctor) != napi_ok || napi_new_instance(env, ctor, 0, nullptr, &obj) != napi_ok) return nullptr;
            napi_set_named_property(env, obj, "cells", cells);
            napi_set_named_property(env, obj, "proofs", proofs);
            return obj;
        }
        default: {
            napi_value b;
            return napi_get_boolean(env, j.verified, &b) == napi_ok ? b : nullptr;
        }
    }
}

template <int KIND>
napi_value method_sync(napi_env env, napi_callback_info info) {
    Job job;
    if (!parse(env, info, (Kind)KIND, &job)) return nullptr;
    run(&job);
    if (!job.ok) { throw_msg(env, job.error); return nullptr; }
    napi_value v = to_js(env, job);
    if (!v) throw_msg(env, "could not create the result value");
    return v;
}

void async_execute(napi_env, void* data) { run(static_cast<Job*>(data)); }
void async_complete(napi_env env, napi_status status, void* data) {
    Job* job = static_cast<Job*>(data);
    napi_value v = nullptr;
    std::string err = status != napi_ok ? "async work was cancelled" : job->error;
    if (status == napi_ok && job->ok) {
        v = to_js(env, *job);
        if (!v) err = "could not create the result value";
    }
    if (v) {
        napi_resolve_deferred(env, job->deferred, v);
    } else {
        napi_value msg, e;
        napi_create_string_utf8(env, err.c_str(), err.size(), &msg);
        napi_create_error(env, nullptr, msg, &e);
        napi_reject_deferred(env, job->deferred, e);
    }
    napi_delete_async_work(env, job->work);
    delete job;
}
template <int KIND>
napi_value method_async(napi_env env, napi_callback_info info) {
    Job* job = new Job();
    if (!parse(env, info, (Kind)KIND, job)) { delete job; return nullptr; }   // argument errors throw synchronously, as napi-rs' conversions do
    napi_value promise, name;
    if (napi_create_promise(env, &job->deferred, &promise) != napi_ok) { delete job; throw_msg(env, "napi_create_promise failed"); return nullptr; }
    napi_create_string_utf8(env, FN_NAME[KIND], NAPI_AUTO_LENGTH, &name);
    if (napi_create_async_work(env, nullptr, name, async_execute, async_complete, job, &job->work) != napi_ok || napi_queue_async_work(env, job->work) != napi_ok) {
        delete job;
        throw_msg(env, "could not queue the async work");
        return nullptr;
    }
    return promise;
}

void finalize_native(napi_env, void* data, void*) {
    Native* n = static_cast<Native*>(data);
    if (n) { eth_kzg_das_context_free(n->ctx); delete n; }
}

bool attach_context(napi_env env, napi_value self, bool use_precomp) {
    DASContext* c = eth_kzg_das_context_new(use_precomp);
    if (!c) return throw_msg(env, "DASContext creation failed: no usable CUDA device (this backend has no CPU fallback)");
    Native* n = new Native();
    n->ctx = c;
    if (napi_wrap(env, self, n, finalize_native, nullptr, nullptr) != napi_ok) {
        finalize_native(env, n, nullptr);
        return throw_msg(env, "napi_wrap failed");
    }
    return true;
}

bool option_use_precomp(napi_env env, napi_value options, bool* out) {
    napi_valuetype t;
    napi_value v;
    if (napi_typeof(env, options, &t) != napi_ok || t != napi_object || napi_get_named_property(env, options, "usePrecomp", &v) != napi_ok ||
        napi_get_value_bool(env, v, out) != napi_ok)
        return throw_msg(env, "options must be an object {usePrecomp: boolean}");
    return true;
}

// new DasContextJs(): DASContextOptions::default() = {usePrecomp: true}.  create() passes its options object through (an internal route:
// the TypeScript signature of the constructor takes no arguments).
napi_value context_ctor(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1] = {nullptr}, self;
    if (napi_get_cb_info(env, info, &argc, argv, &self, nullptr) != napi_ok) { throw_msg(env, "napi_get_cb_info failed"); return nullptr; }
    bool use_precomp = true;
    if (argc >= 1 && argv[0]) {
        napi_valuetype t;
        if (napi_typeof(env, argv[0], &t) == napi_ok && t != napi_undefined && !option_use_precomp(env, argv[0], &use_precomp)) return nullptr;
    }
    if (!attach_context(env, self, use_precomp)) return nullptr;
    return self;
}

napi_value context_create(napi_env env, napi_callback_info info) {   // static create(options)
    size_t argc = 1;
    napi_value argv[1] = {nullptr};
    Addon* addon = nullptr;
    if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok || argc < 1) { throw_msg(env, "create(options): options missing"); return nullptr; }
    bool use_precomp = true;
    if (!option_use_precomp(env, argv[0], &use_precomp)) return nullptr;
    napi_value ctor, obj;
    if (napi_get_instance_data(env, reinterpret_cast<void**>(&addon)) != napi_ok || !addon || napi_get_reference_value(env, addon->context_ctor, &ctor) != napi_ok ||
        napi_new_instance(env, ctor, 1, argv, &obj) != napi_ok)
        return nullptr;   // (an exception thrown by the constructor is pending)
    return obj;
}

napi_value plain_ctor(napi_env env, napi_callback_info info) {
    napi_value self = nullptr;
    napi_get_cb_info(env, info, nullptr, nullptr, &self, nullptr);
    return self;
}

#define EKZG_METHOD(js, kind) {js, nullptr, method_sync<kind>, nullptr, nullptr, nullptr, napi_default, nullptr}
#define EKZG_ASYNC(js, kind) {js, nullptr, method_async<kind>, nullptr, nullptr, nullptr, napi_default, nullptr}

napi_value init(napi_env env, napi_value exports) {
    Addon* addon = new Addon();
    napi_set_instance_data(env, addon, [](napi_env, void* d, void*) { delete static_cast<Addon*>(d); }, nullptr);
    const struct { const char* name; uint32_t v; } consts[] = {{"BYTES_PER_COMMITMENT", 48}, {"BYTES_PER_PROOF", 48}, {"BYTES_PER_FIELD_ELEMENT", 32},
                                                                {"BYTES_PER_BLOB", 131072}, {"MAX_NUM_COLUMNS", 128}, {"BYTES_PER_CELL", 2048}};
    for (auto& c : consts) {
        napi_value v;
        napi_create_uint32(env, c.v, &v);
        napi_set_named_property(env, exports, c.name, v);
    }
    napi_value cap;
    napi_define_class(env, "CellsAndProofs", NAPI_AUTO_LENGTH, plain_ctor, nullptr, 0, nullptr, &cap);
    napi_create_reference(env, cap, 1, &addon->cells_and_proofs_ctor);
    napi_set_named_property(env, exports, "CellsAndProofs", cap);
    const napi_property_descriptor props[] = {
        {"create", nullptr, context_create, nullptr, nullptr, nullptr, napi_static, nullptr},
        EKZG_METHOD("blobToKzgCommitment", Commit), EKZG_ASYNC("asyncBlobToKzgCommitment", Commit),
        EKZG_METHOD("computeCellsAndKzgProofs", CellsProofs), EKZG_ASYNC("asyncComputeCellsAndKzgProofs", CellsProofs),
        EKZG_METHOD("computeCells", Cells), EKZG_ASYNC("asyncComputeCells", Cells),
        EKZG_METHOD("recoverCellsAndKzgProofs", Recover), EKZG_ASYNC("asyncRecoverCellsAndKzgProofs", Recover),
        EKZG_METHOD("verifyCellKzgProofBatch", VerifyCellBatch), EKZG_ASYNC("asyncVerifyCellKzgProofBatch", VerifyCellBatch),
        EKZG_METHOD("computeKzgProof", KzgProof), EKZG_ASYNC("asyncComputeKzgProof", KzgProof),
        EKZG_METHOD("computeBlobKzgProof", BlobProof), EKZG_ASYNC("asyncComputeBlobKzgProof", BlobProof),
        EKZG_METHOD("verifyKzgProof", VerifyKzg), EKZG_ASYNC("asyncVerifyKzgProof", VerifyKzg),
        EKZG_METHOD("verifyBlobKzgProof", VerifyBlob), EKZG_ASYNC("asyncVerifyBlobKzgProof", VerifyBlob),
        EKZG_METHOD("verifyBlobKzgProofBatch", VerifyBlobBatch), EKZG_ASYNC("asyncVerifyBlobKzgProofBatch", VerifyBlobBatch),
    };
    napi_value cls;
    napi_define_class(env, "DasContextJs", NAPI_AUTO_LENGTH, context_ctor, nullptr, sizeof(props) / sizeof(props[0]), props, &cls);
    napi_create_reference(env, cls, 1, &addon->context_ctor);
    napi_set_named_property(env, exports, "DasContextJs", cls);
    return exports;
}

}  // namespace

// what NAPI_MODULE_INIT() expands to: node looks this symbol up when it loads the .node file
extern "C" NAPI_MODULE_EXPORT napi_value napi_register_module_v1(napi_env env, napi_value exports) { return init(env, exports); }
