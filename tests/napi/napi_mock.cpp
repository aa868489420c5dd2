// A mock Node-API for exercising rust-eth-kzg_b200/shims/node/eth_kzg_node.cpp without node (test infrastructure).
// Implements the ~40 napi_* functions the addon calls over a small heap value model; napi_queue_async_work starts the execute
// callback on a thread of its own at once (the libuv pool), mock_napi_drain joins them and runs the complete callbacks on the calling
// ("JS") thread.  Load this library with RTLD_GLOBAL before the addon so that the addon's napi_* references resolve here.
#include "../../rust-eth-kzg_b200/shims/node/node_api_min.h"
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct napi_value__ {
    enum T { Undef, Null, Bool, Number, BigInt, String, Object, Array, U8Array, ArrayBuffer, Class, Error, Promise } t = Undef;
    bool b = false;
    double num = 0;
    uint64_t big = 0;
    std::string s;                       // String / Error message / Class name
    std::vector<uint8_t> bytes;          // ArrayBuffer
    napi_value ab = nullptr;             // U8Array
    size_t off = 0, len = 0;
    std::vector<napi_value> elems;
    std::map<std::string, napi_value> props;
    napi_callback ctor = nullptr;        // Class
    std::vector<napi_property_descriptor> methods;
    napi_value cls = nullptr;            // instance
    void* wrapped = nullptr;
    napi_finalize fin = nullptr;
    int pstate = 0;                      // Promise: 0 pending, 1 resolved, 2 rejected
    napi_value presult = nullptr;
};
struct napi_ref__ { napi_value v; };
struct napi_deferred__ { napi_value promise; };
struct napi_callback_info__ { std::vector<napi_value> args; napi_value self; void* data; };
struct napi_async_work__ { napi_async_execute_callback ex; napi_async_complete_callback done; void* data; std::thread th; bool queued = false; };
struct napi_env__ {
    std::vector<std::unique_ptr<napi_value__>> arena;
    std::vector<std::unique_ptr<napi_ref__>> refs;
    std::vector<napi_async_work> queued;
    napi_value pending = nullptr;
    void* instance_data = nullptr;
    napi_finalize instance_fin = nullptr;
    napi_value make(napi_value__::T t) { arena.emplace_back(new napi_value__()); arena.back()->t = t; return arena.back().get(); }
};

extern "C" {
napi_status napi_define_class(napi_env env, const char* name, size_t, napi_callback ctor, void*, size_t n, const napi_property_descriptor* p, napi_value* result) {
    napi_value c = env->make(napi_value__::Class);
    c->s = name; c->ctor = ctor;
    for (size_t i = 0; i < n; i++) c->methods.push_back(p[i]);
    *result = c;
    return napi_ok;
}
napi_status napi_wrap(napi_env, napi_value o, void* native, napi_finalize fin, void*, napi_ref*) { o->wrapped = native; o->fin = fin; return napi_ok; }
napi_status napi_unwrap(napi_env, napi_value o, void** result) { if (!o || !o->wrapped) return napi_invalid_arg; *result = o->wrapped; return napi_ok; }
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* self, void** data) {
    if (argc) {
        const size_t cap = *argc;
        for (size_t i = 0; i < cap && argv; i++) argv[i] = i < info->args.size() ? info->args[i] : nullptr;
        *argc = info->args.size();
    }
    if (self) *self = info->self;
    if (data) *data = info->data;
    return napi_ok;
}
napi_status napi_new_instance(napi_env env, napi_value cls, size_t argc, const napi_value* argv, napi_value* result) {
    if (!cls || cls->t != napi_value__::Class) return napi_function_expected;
    napi_value o = env->make(napi_value__::Object);
    o->cls = cls;
    napi_callback_info__ info;
    for (size_t i = 0; i < argc; i++) info.args.push_back(argv[i]);
    info.self = o; info.data = nullptr;
    napi_value r = cls->ctor(env, &info);
    if (env->pending) return napi_pending_exception;
    *result = r ? r : o;
    return napi_ok;
}
napi_status napi_create_reference(napi_env env, napi_value v, uint32_t, napi_ref* result) { env->refs.emplace_back(new napi_ref__{v}); *result = env->refs.back().get(); return napi_ok; }
napi_status napi_get_reference_value(napi_env, napi_ref r, napi_value* result) { if (!r) return napi_invalid_arg; *result = r->v; return napi_ok; }
napi_status napi_set_instance_data(napi_env env, void* d, napi_finalize f, void*) { env->instance_data = d; env->instance_fin = f; return napi_ok; }
napi_status napi_get_instance_data(napi_env env, void** d) { *d = env->instance_data; return napi_ok; }
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* r) {
    if (!v) return napi_invalid_arg;
    switch (v->t) {
        case napi_value__::Undef: *r = napi_undefined; break;
        case napi_value__::Null: *r = napi_null; break;
        case napi_value__::Bool: *r = napi_boolean; break;
        case napi_value__::Number: *r = napi_number; break;
        case napi_value__::BigInt: *r = napi_bigint; break;
        case napi_value__::String: *r = napi_string; break;
        case napi_value__::Class: *r = napi_function; break;
        default: *r = napi_object; break;
    }
    return napi_ok;
}
napi_status napi_is_array(napi_env, napi_value v, bool* r) { *r = v && v->t == napi_value__::Array; return napi_ok; }
napi_status napi_is_typedarray(napi_env, napi_value v, bool* r) { *r = v && v->t == napi_value__::U8Array; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* ab, size_t* off) {
    if (!v || v->t != napi_value__::U8Array) return napi_invalid_arg;
    if (type) *type = napi_uint8_array;
    if (length) *length = v->len;
    if (data) *data = v->ab->bytes.data() + v->off;
    if (ab) *ab = v->ab;
    if (off) *off = v->off;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t n, void** data, napi_value* result) {
    napi_value a = env->make(napi_value__::ArrayBuffer);
    a->bytes.assign(n, 0);
    if (data) *data = a->bytes.data();
    *result = a;
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t off, napi_value* result) {
    if (type != napi_uint8_array || !ab || ab->t != napi_value__::ArrayBuffer || off + length > ab->bytes.size()) return napi_invalid_arg;
    napi_value t = env->make(napi_value__::U8Array);
    t->ab = ab; t->off = off; t->len = length;
    *result = t;
    return napi_ok;
}
napi_status napi_create_array_with_length(napi_env env, size_t n, napi_value* result) { napi_value a = env->make(napi_value__::Array); a->elems.assign(n, nullptr); *result = a; return napi_ok; }
napi_status napi_get_array_length(napi_env, napi_value v, uint32_t* r) { if (!v || v->t != napi_value__::Array) return napi_array_expected; *r = (uint32_t)v->elems.size(); return napi_ok; }
napi_status napi_get_element(napi_env env, napi_value v, uint32_t i, napi_value* r) {
    if (!v || v->t != napi_value__::Array) return napi_array_expected;
    *r = i < v->elems.size() && v->elems[i] ? v->elems[i] : env->make(napi_value__::Undef);
    return napi_ok;
}
napi_status napi_set_element(napi_env, napi_value v, uint32_t i, napi_value e) {
    if (!v || v->t != napi_value__::Array) return napi_array_expected;
    if (i >= v->elems.size()) v->elems.resize(i + 1, nullptr);
    v->elems[i] = e;
    return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* r) { *r = env->make(napi_value__::Object); return napi_ok; }
napi_status napi_set_named_property(napi_env, napi_value o, const char* name, napi_value v) { if (!o) return napi_object_expected; o->props[name] = v; return napi_ok; }
napi_status napi_get_named_property(napi_env env, napi_value o, const char* name, napi_value* r) {
    if (!o) return napi_object_expected;
    auto it = o->props.find(name);
    *r = it == o->props.end() ? env->make(napi_value__::Undef) : it->second;
    return napi_ok;
}
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t* r) { if (!v || v->t != napi_value__::Number) return napi_number_expected; *r = (uint32_t)(int64_t)v->num; return napi_ok; }
napi_status napi_get_value_bigint_uint64(napi_env, napi_value v, uint64_t* r, bool* lossless) {
    if (!v || v->t != napi_value__::BigInt) return napi_bigint_expected;
    *r = v->big; *lossless = true;
    return napi_ok;
}
napi_status napi_get_value_bool(napi_env, napi_value v, bool* r) { if (!v || v->t != napi_value__::Bool) return napi_boolean_expected; *r = v->b; return napi_ok; }
napi_status napi_get_boolean(napi_env env, bool b, napi_value* r) { napi_value v = env->make(napi_value__::Bool); v->b = b; *r = v; return napi_ok; }
napi_status napi_get_undefined(napi_env env, napi_value* r) { *r = env->make(napi_value__::Undef); return napi_ok; }
napi_status napi_create_uint32(napi_env env, uint32_t x, napi_value* r) { napi_value v = env->make(napi_value__::Number); v->num = x; *r = v; return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char* s, size_t n, napi_value* r) {
    napi_value v = env->make(napi_value__::String);
    v->s = n == NAPI_AUTO_LENGTH ? std::string(s) : std::string(s, n);
    *r = v;
    return napi_ok;
}
napi_status napi_create_error(napi_env env, napi_value, napi_value msg, napi_value* r) { napi_value e = env->make(napi_value__::Error); e->s = msg ? msg->s : ""; *r = e; return napi_ok; }
napi_status napi_throw_error(napi_env env, const char*, const char* msg) { napi_value e = env->make(napi_value__::Error); e->s = msg; env->pending = e; return napi_ok; }
napi_status napi_create_promise(napi_env env, napi_deferred* d, napi_value* promise) {
    napi_value p = env->make(napi_value__::Promise);
    *d = new napi_deferred__{p};
    *promise = p;
    return napi_ok;
}
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->pstate = 1; d->promise->presult = v; delete d; return napi_ok; }
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->pstate = 2; d->promise->presult = v; delete d; return napi_ok; }
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback ex, napi_async_complete_callback done, void* data, napi_async_work* r) {
    napi_async_work w = new napi_async_work__();
    w->ex = ex; w->done = done; w->data = data;
    *r = w;
    return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) {
    w->queued = true;
    w->th = std::thread([env, w] { w->ex(env, w->data); });
    env->queued.push_back(w);
    return napi_ok;
}
napi_status napi_delete_async_work(napi_env, napi_async_work w) { delete w; return napi_ok; }

// ---- test-side API ---------------------------------------------------------------------------------------------------------
void* mock_napi_env_new() { return new napi_env__(); }
void mock_napi_env_free(void* e) {
    napi_env env = (napi_env)e;
    for (auto& v : env->arena) if (v->wrapped && v->fin) v->fin(env, v->wrapped, nullptr);   // garbage collection of the wrapped contexts
    if (env->instance_fin) env->instance_fin(env, env->instance_data, nullptr);
    delete env;
}
void* mock_napi_load(void* e, void* register_fn) {
    napi_env env = (napi_env)e;
    napi_value exports = env->make(napi_value__::Object);
    typedef napi_value (*reg_t)(napi_env, napi_value);
    return ((reg_t)register_fn)(env, exports);
}
void* mock_napi_u8(void* e, const uint8_t* p, long n) {
    napi_env env = (napi_env)e;
    napi_value ab = env->make(napi_value__::ArrayBuffer), t = env->make(napi_value__::U8Array);
    ab->bytes.assign(p, p + n);
    t->ab = ab; t->off = 0; t->len = (size_t)n;
    return t;
}
void* mock_napi_number(void* e, double x) { napi_value v = ((napi_env)e)->make(napi_value__::Number); v->num = x; return v; }
void* mock_napi_bigint(void* e, uint64_t x) { napi_value v = ((napi_env)e)->make(napi_value__::BigInt); v->big = x; return v; }
void* mock_napi_bool(void* e, int b) { napi_value v = ((napi_env)e)->make(napi_value__::Bool); v->b = b != 0; return v; }
void* mock_napi_string(void* e, const char* s) { napi_value v = ((napi_env)e)->make(napi_value__::String); v->s = s; return v; }
void* mock_napi_array(void* e, long n) { napi_value v = ((napi_env)e)->make(napi_value__::Array); v->elems.assign((size_t)n, nullptr); return v; }
void mock_napi_array_set(void* a, long i, void* v) { ((napi_value)a)->elems[(size_t)i] = (napi_value)v; }
void* mock_napi_object(void* e) { return ((napi_env)e)->make(napi_value__::Object); }
void mock_napi_set_prop(void* o, const char* name, void* v) { ((napi_value)o)->props[name] = (napi_value)v; }
void* mock_napi_get_prop(void* o, const char* name) { auto& p = ((napi_value)o)->props; auto it = p.find(name); return it == p.end() ? nullptr : it->second; }
int mock_napi_type(void* v) { return (int)((napi_value)v)->t; }
long mock_napi_len(void* v) { napi_value x = (napi_value)v; return (long)(x->t == napi_value__::U8Array ? x->len : x->elems.size()); }
void mock_napi_u8_get(void* v, uint8_t* out) { napi_value x = (napi_value)v; memcpy(out, x->ab->bytes.data() + x->off, x->len); }
void* mock_napi_elem(void* v, long i) { return ((napi_value)v)->elems[(size_t)i]; }
double mock_napi_number_value(void* v) { return ((napi_value)v)->num; }
int mock_napi_bool_value(void* v) { return ((napi_value)v)->b ? 1 : 0; }
const char* mock_napi_text(void* v) { return ((napi_value)v)->s.c_str(); }
const char* mock_napi_class_name(void* v) { napi_value x = (napi_value)v; return x->cls ? x->cls->s.c_str() : ""; }
int mock_napi_method_count(void* cls) { return (int)((napi_value)cls)->methods.size(); }
const char* mock_napi_method_name(void* cls, int i) { return ((napi_value)cls)->methods[(size_t)i].utf8name; }
void* mock_napi_new(void* e, void* cls, int argc, void** argv) {
    napi_value r = nullptr;
    std::vector<napi_value> a;
    for (int i = 0; i < argc; i++) a.push_back((napi_value)argv[i]);
    return napi_new_instance((napi_env)e, (napi_value)cls, (size_t)argc, a.data(), &r) == napi_ok ? r : nullptr;
}
// obj.method(args) for an instance, Class.method(args) for a class value (static methods)
void* mock_napi_call(void* e, void* target, const char* name, int argc, void** argv) {
    napi_env env = (napi_env)e;
    napi_value t = (napi_value)target;
    const bool is_static = t->t == napi_value__::Class;
    napi_value cls = is_static ? t : t->cls;
    if (!cls) return nullptr;
    for (auto& m : cls->methods) {
        if (strcmp(m.utf8name, name) != 0 || ((m.attributes & napi_static) != 0) != is_static) continue;
        napi_callback_info__ info;
        for (int i = 0; i < argc; i++) info.args.push_back((napi_value)argv[i]);
        info.self = t; info.data = m.data;
        napi_value r = m.method(env, &info);
        return env->pending ? nullptr : r;
    }
    napi_throw_error(env, nullptr, "no such method");
    return nullptr;
}
int mock_napi_take_exception(void* e, char* buf, long cap) {
    napi_env env = (napi_env)e;
    if (!env->pending) return 0;
    strncpy(buf, env->pending->s.c_str(), (size_t)cap - 1);
    buf[cap - 1] = 0;
    env->pending = nullptr;
    return 1;
}
int mock_napi_promise_state(void* v) { return ((napi_value)v)->pstate; }
void* mock_napi_promise_result(void* v) { return ((napi_value)v)->presult; }
// the event loop turning: wait for every queued execute callback, then run the complete callbacks on this thread
int mock_napi_drain(void* e) {
    napi_env env = (napi_env)e;
    std::vector<napi_async_work> q;
    q.swap(env->queued);
    for (napi_async_work w : q) w->th.join();
    for (napi_async_work w : q) w->done(env, napi_ok, w->data);
    return (int)q.size();
}
}
