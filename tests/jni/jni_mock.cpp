// A mock JNIEnv for exercising rust-eth-kzg_b200/shims/jni/eth_kzg_jni.cpp without a JVM (test infrastructure).
// Implements the 16 function-table slots the shim calls over plain heap objects, records the pending exception, and exposes a small
// C API so that tests/test_jni_shim.py can build byte[] / long[] / byte[][] arguments and inspect the returned objects.
#include "../../rust-eth-kzg_b200/shims/jni/jni_min.h"
#include <cstring>
#include <memory>
#include <string>
#include <vector>

struct _jobject {
    enum Kind { Bytes, Longs, Objects, Class, Instance } kind;
    std::vector<int8_t> bytes;
    std::vector<int64_t> longs;
    std::vector<_jobject*> elems;    // Objects: elements; Instance: constructor arguments
    std::string name;                // Class: its name; Instance: class name + constructor signature
};
struct _jmethodID { std::string cls, name, sig; };

namespace {
struct Mock {
    const JNINativeInterface_* table;   // must be first: JNIEnv* points here
    std::vector<std::unique_ptr<_jobject>> arena;
    std::vector<std::unique_ptr<_jmethodID>> methods;
    bool pending = false;
    std::string exc_class, exc_msg;
    long live_local_refs = 0, max_local_refs = 0;
    _jobject* make(_jobject::Kind k) {
        arena.emplace_back(new _jobject());
        arena.back()->kind = k;
        if (++live_local_refs > max_local_refs) max_local_refs = live_local_refs;
        return arena.back().get();
    }
};
Mock* M(JNIEnv* env) { return reinterpret_cast<Mock*>(env); }

jint GetVersion(JNIEnv*) { return JNI_VERSION_1_8; }
jclass FindClass(JNIEnv* env, const char* name) {
    const std::string n = name;
    if (n != "[B" && n != "ethereum/cryptography/CellsAndProofs" && n != "ethereum/cryptography/Cells" && n != "java/lang/IllegalArgumentException") {
        M(env)->pending = true; M(env)->exc_class = "java/lang/NoClassDefFoundError"; M(env)->exc_msg = n;
        return nullptr;
    }
    _jobject* c = M(env)->make(_jobject::Class);
    c->name = n;
    return c;
}
jint ThrowNew(JNIEnv* env, jclass cls, const char* msg) { M(env)->pending = true; M(env)->exc_class = cls->name; M(env)->exc_msg = msg; return 0; }
void ExceptionClear(JNIEnv* env) { M(env)->pending = false; }
jboolean ExceptionCheck(JNIEnv* env) { return M(env)->pending ? JNI_TRUE : JNI_FALSE; }
void DeleteLocalRef(JNIEnv* env, jobject) { M(env)->live_local_refs--; }
jmethodID GetMethodID(JNIEnv* env, jclass cls, const char* name, const char* sig) {
    const bool ok = !strcmp(name, "<init>") && ((cls->name == "ethereum/cryptography/CellsAndProofs" && !strcmp(sig, "([[B[[B)V")) ||
                                                (cls->name == "ethereum/cryptography/Cells" && !strcmp(sig, "([[B)V")));
    if (!ok) { M(env)->pending = true; M(env)->exc_class = "java/lang/NoSuchMethodError"; M(env)->exc_msg = cls->name + "." + name + sig; return nullptr; }
    M(env)->methods.emplace_back(new _jmethodID{cls->name, name, sig});
    return M(env)->methods.back().get();
}
jobject NewObjectA(JNIEnv* env, jclass cls, jmethodID m, const jvalue* args) {
    _jobject* o = M(env)->make(_jobject::Instance);
    o->name = cls->name + m->sig;
    const int nargs = m->sig == "([[B[[B)V" ? 2 : 1;
    for (int i = 0; i < nargs; i++) o->elems.push_back(args[i].l);
    return o;
}
jsize GetArrayLength(JNIEnv*, jarray a) {
    return (jsize)(a->kind == _jobject::Bytes ? a->bytes.size() : a->kind == _jobject::Longs ? a->longs.size() : a->elems.size());
}
jobjectArray NewObjectArray(JNIEnv* env, jsize n, jclass, jobject init) {
    _jobject* a = M(env)->make(_jobject::Objects);
    a->elems.assign((size_t)n, init);
    return a;
}
jobject GetObjectArrayElement(JNIEnv* env, jobjectArray a, jsize i) {
    if (i < 0 || (size_t)i >= a->elems.size()) { M(env)->pending = true; M(env)->exc_class = "java/lang/ArrayIndexOutOfBoundsException"; return nullptr; }
    M(env)->live_local_refs++;
    if (M(env)->live_local_refs > M(env)->max_local_refs) M(env)->max_local_refs = M(env)->live_local_refs;
    return a->elems[(size_t)i];
}
void SetObjectArrayElement(JNIEnv* env, jobjectArray a, jsize i, jobject v) {
    if (i < 0 || (size_t)i >= a->elems.size()) { M(env)->pending = true; M(env)->exc_class = "java/lang/ArrayIndexOutOfBoundsException"; return; }
    a->elems[(size_t)i] = v;
}
jbyteArray NewByteArray(JNIEnv* env, jsize n) {
    _jobject* a = M(env)->make(_jobject::Bytes);
    a->bytes.assign((size_t)n, 0);
    return a;
}
void GetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize start, jsize len, jbyte* buf) {
    if (a->kind != _jobject::Bytes || start < 0 || len < 0 || (size_t)(start + len) > a->bytes.size()) {
        M(env)->pending = true; M(env)->exc_class = "java/lang/ArrayIndexOutOfBoundsException"; return;
    }
    memcpy(buf, a->bytes.data() + start, (size_t)len);
}
void SetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize start, jsize len, const jbyte* buf) {
    if (a->kind != _jobject::Bytes || start < 0 || len < 0 || (size_t)(start + len) > a->bytes.size()) {
        M(env)->pending = true; M(env)->exc_class = "java/lang/ArrayIndexOutOfBoundsException"; return;
    }
    memcpy(a->bytes.data() + start, buf, (size_t)len);
}
void GetLongArrayRegion(JNIEnv* env, jlongArray a, jsize start, jsize len, jlong* buf) {
    if (a->kind != _jobject::Longs || start < 0 || len < 0 || (size_t)(start + len) > a->longs.size()) {
        M(env)->pending = true; M(env)->exc_class = "java/lang/ArrayIndexOutOfBoundsException"; return;
    }
    memcpy(buf, a->longs.data() + start, (size_t)len * 8);
}

JNINativeInterface_ make_table() {
    JNINativeInterface_ t;
    memset(&t, 0, sizeof t);
    t.GetVersion = GetVersion; t.FindClass = FindClass; t.ThrowNew = ThrowNew; t.ExceptionClear = ExceptionClear; t.DeleteLocalRef = DeleteLocalRef;
    t.NewObjectA = NewObjectA; t.GetMethodID = GetMethodID; t.GetArrayLength = GetArrayLength; t.NewObjectArray = NewObjectArray;
    t.GetObjectArrayElement = GetObjectArrayElement; t.SetObjectArrayElement = SetObjectArrayElement; t.NewByteArray = NewByteArray;
    t.GetByteArrayRegion = GetByteArrayRegion; t.GetLongArrayRegion = GetLongArrayRegion; t.SetByteArrayRegion = SetByteArrayRegion;
    t.ExceptionCheck = ExceptionCheck;
    return t;
}
const JNINativeInterface_ g_table = make_table();
}  // namespace

extern "C" {
void* mock_env_new() { Mock* m = new Mock(); m->table = &g_table; return m; }
void mock_env_free(void* env) { delete reinterpret_cast<Mock*>(env); }
int mock_table_slots() { return (int)(sizeof(JNINativeInterface_) / sizeof(void*)); }
void* mock_new_bytes(void* env, const uint8_t* p, long n) { _jobject* a = reinterpret_cast<Mock*>(env)->make(_jobject::Bytes); a->bytes.assign(p, p + n); return a; }
void* mock_new_longs(void* env, const int64_t* p, long n) { _jobject* a = reinterpret_cast<Mock*>(env)->make(_jobject::Longs); a->longs.assign(p, p + n); return a; }
void* mock_new_array(void* env, long n) { _jobject* a = reinterpret_cast<Mock*>(env)->make(_jobject::Objects); a->elems.assign((size_t)n, nullptr); return a; }
void mock_array_set(void* arr, long i, void* v) { reinterpret_cast<_jobject*>(arr)->elems[(size_t)i] = reinterpret_cast<_jobject*>(v); }
long mock_len(void* o) { _jobject* a = reinterpret_cast<_jobject*>(o); return (long)(a->kind == _jobject::Bytes ? a->bytes.size() : a->elems.size()); }
void mock_bytes_get(void* o, uint8_t* out) { _jobject* a = reinterpret_cast<_jobject*>(o); memcpy(out, a->bytes.data(), a->bytes.size()); }
void* mock_elem(void* o, long i) { return reinterpret_cast<_jobject*>(o)->elems[(size_t)i]; }
const char* mock_name(void* o) { return reinterpret_cast<_jobject*>(o)->name.c_str(); }
int mock_kind(void* o) { return (int)reinterpret_cast<_jobject*>(o)->kind; }
// pending exception -> "class: message" (and clears it); returns 0 if none
int mock_take_exception(void* env, char* buf, long cap) {
    Mock* m = reinterpret_cast<Mock*>(env);
    if (!m->pending) return 0;
    const std::string s = m->exc_class + ": " + m->exc_msg;
    strncpy(buf, s.c_str(), (size_t)cap - 1);
    buf[cap - 1] = 0;
    m->pending = false;
    return 1;
}
long mock_max_local_refs(void* env) { return reinterpret_cast<Mock*>(env)->max_local_refs; }
}
