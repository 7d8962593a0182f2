// napi_mock.cc — TEST INFRASTRUCTURE: a small in-process stand-in for the Node-API runtime, so that addon/rm_napi.cc can be
// EXECUTED (not just compiled) in an image without Node.js.  It implements the calls addon/napi_min.h declares over a plain
// object model and enforces the one rule of Node-API that compile checks cannot: a napi_value is only valid inside the handle
// scope it was created in.  Every entry into addon code (module init, a method call, an async completion callback) runs in a
// fresh scope; using a handle from an older scope aborts with a message — exactly the bug class of keeping a napi_value
// across napi_queue_async_work.  Async work executes on real threads, concurrently, and completes on the calling thread when
// the harness drains the queue (like libuv's thread pool + the JS main loop).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../addon/napi_min.h"
#include "napi_mock.h"

namespace mock {

uint64_t g_scope = 1;
std::vector<std::unique_ptr<Handle>> g_handles;
std::vector<std::unique_ptr<Val>> g_vals;
std::string g_exception;
int g_live_refs = 0;
std::vector<Work*> g_queue;

[[noreturn]] void die(const char* msg) {
    std::fprintf(stderr, "napi_mock: %s\n", msg);
    std::abort();
}

Val* new_val(Val::Kind k) {
    g_vals.emplace_back(new Val());
    g_vals.back()->kind = k;
    return g_vals.back().get();
}
napi_value wrap(Val* v) {
    g_handles.emplace_back(new Handle{v, g_scope});
    return reinterpret_cast<napi_value>(g_handles.back().get());
}
Val* unwrap(napi_value h) {
    if (!h) die("null napi_value");
    Handle* hd = reinterpret_cast<Handle*>(h);
    if (hd->scope != g_scope) die("napi_value used outside the handle scope it was created in (keep a napi_ref instead)");
    return hd->v;
}
void open_scope() { ++g_scope; }

Val* make_number(double d) {
    Val* v = new_val(Val::Num);
    v->num = d;
    return v;
}
Val* make_string(const std::string& s) {
    Val* v = new_val(Val::Str);
    v->str = s;
    return v;
}
Val* make_object() { return new_val(Val::Obj); }
Val* make_typedarray(int type, void* data, size_t length, size_t elem) {
    Val* ab = new_val(Val::ArrayBuffer);
    ab->data = data;
    ab->len = length * elem;
    Val* ta = new_val(Val::TypedArray);
    ta->ta_type = type;
    ta->len = length;
    ta->ab = ab;
    ta->offset = 0;
    return ta;
}

Val* call(Val* fn, const std::vector<Val*>& args) {
    if (!fn || fn->kind != Val::Func) die("call of a non-function");
    open_scope();
    CallInfo ci;
    for (Val* a : args) ci.argv.push_back(wrap(a));
    g_exception.clear();
    napi_value r = fn->cb(reinterpret_cast<napi_env>(&g_scope), reinterpret_cast<napi_callback_info>(&ci));
    Val* out = r ? unwrap(r) : nullptr;
    open_scope();  // the callback's handles die here
    return out;
}

void drain() {
    // libuv: execute callbacks ran on pool threads (started at queue time); completions run on the main thread, each in its own scope
    while (!g_queue.empty()) {
        std::vector<Work*> q;
        q.swap(g_queue);
        for (Work* w : q) {
            w->th.join();
            open_scope();
            w->complete(reinterpret_cast<napi_env>(&g_scope), napi_ok, w->data);
            open_scope();
        }
    }
}

int gc() {  // collect every external ArrayBuffer that is not reachable from a live reference: here simply all of them
    int n = 0;
    for (auto& v : g_vals)
        if (v->kind == Val::ArrayBuffer && v->fin) {
            v->fin(reinterpret_cast<napi_env>(&g_scope), v->data, v->hint);
            v->fin = nullptr;
            v->data = nullptr;
            ++n;
        }
    return n;
}

}  // namespace mock

using namespace mock;

extern "C" {

napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    CallInfo* ci = reinterpret_cast<CallInfo*>(info);
    const size_t want = argc ? *argc : 0;
    for (size_t i = 0; i < want; ++i) argv[i] = i < ci->argv.size() ? ci->argv[i] : wrap(new_val(Val::Undef));
    if (argc) *argc = ci->argv.size();
    if (this_arg) *this_arg = nullptr;
    if (data) *data = nullptr;
    return napi_ok;
}
napi_status napi_get_named_property(napi_env, napi_value object, const char* name, napi_value* result) {
    Val* o = unwrap(object);
    auto it = o->props.find(name);
    *result = wrap(it == o->props.end() ? new_val(Val::Undef) : it->second);
    return napi_ok;
}
napi_status napi_has_named_property(napi_env, napi_value object, const char* name, bool* result) {
    Val* o = unwrap(object);
    *result = o->props.count(name) != 0;
    return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value object, const char* name, napi_value value) {
    unwrap(object)->props[name] = unwrap(value);
    return napi_ok;
}
napi_status napi_get_value_double(napi_env, napi_value v, double* result) {
    Val* x = unwrap(v);
    if (x->kind != Val::Num) return (napi_status)1;
    *result = x->num;
    return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* result) {
    Val* x = unwrap(v);
    if (x->kind != Val::Num) return (napi_status)1;
    *result = (int32_t)x->num;
    return napi_ok;
}
napi_status napi_get_value_string_utf8(napi_env, napi_value v, char* buf, size_t bufsize, size_t* result) {
    Val* x = unwrap(v);
    if (x->kind != Val::Str) {
        if (result) *result = 0;
        return (napi_status)1;
    }
    const size_t n = std::min(bufsize ? bufsize - 1 : 0, x->str.size());
    if (buf && bufsize) {
        std::memcpy(buf, x->str.data(), n);
        buf[n] = 0;
    }
    if (result) *result = n;
    return napi_ok;
}
napi_status napi_create_object(napi_env, napi_value* result) {
    *result = wrap(make_object());
    return napi_ok;
}
napi_status napi_create_double(napi_env, double d, napi_value* result) {
    *result = wrap(make_number(d));
    return napi_ok;
}
napi_status napi_create_int32(napi_env, int32_t d, napi_value* result) {
    *result = wrap(make_number(d));
    return napi_ok;
}
napi_status napi_create_string_utf8(napi_env, const char* s, size_t n, napi_value* result) {
    *result = wrap(make_string(std::string(s, n)));
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result) {
    Val* ab = new_val(Val::ArrayBuffer);
    ab->owned.resize(byte_length ? byte_length : 1);
    ab->data = ab->owned.data();
    ab->len = byte_length;
    if (data) *data = ab->data;
    *result = wrap(ab);
    return napi_ok;
}
napi_status napi_create_external_arraybuffer(napi_env, void* external_data, size_t byte_length, napi_finalize finalize_cb, void* finalize_hint,
                                             napi_value* result) {
    if (std::getenv("NAPI_MOCK_NO_EXTERNAL_BUFFERS")) return (napi_status)22;  // napi_no_external_buffers_allowed
    Val* ab = new_val(Val::ArrayBuffer);
    ab->data = external_data;
    ab->len = byte_length;
    ab->fin = finalize_cb;
    ab->hint = finalize_hint;
    *result = wrap(ab);
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result) {
    Val* ab = unwrap(arraybuffer);
    if (ab->kind != Val::ArrayBuffer) die("typed array over a non-ArrayBuffer");
    const size_t elem = (type == napi_uint16_array) ? 2 : (type == napi_float32_array ? 4 : (type == napi_float64_array ? 8 : 1));
    if (byte_offset % elem != 0 || byte_offset + length * elem > ab->len) die("typed array out of range / misaligned");
    Val* ta = new_val(Val::TypedArray);
    ta->ta_type = (int)type;
    ta->len = length;
    ta->ab = ab;
    ta->offset = byte_offset;
    *result = wrap(ta);
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer,
                                     size_t* byte_offset) {
    Val* ta = unwrap(v);
    if (ta->kind != Val::TypedArray) return (napi_status)1;
    if (type) *type = (napi_typedarray_type)ta->ta_type;
    if (length) *length = ta->len;
    if (data) *data = (char*)ta->ab->data + ta->offset;
    if (arraybuffer) *arraybuffer = wrap(ta->ab);
    if (byte_offset) *byte_offset = ta->offset;
    return napi_ok;
}
napi_status napi_create_promise(napi_env, napi_deferred* deferred, napi_value* promise) {
    Val* p = new_val(Val::Promise);
    *deferred = reinterpret_cast<napi_deferred>(p);
    *promise = wrap(p);
    return napi_ok;
}
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value resolution) {
    Val* p = reinterpret_cast<Val*>(d);
    if (p->state != 0) die("promise settled twice");
    p->state = 1;
    p->result = unwrap(resolution);
    return napi_ok;
}
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value rejection) {
    Val* p = reinterpret_cast<Val*>(d);
    if (p->state != 0) die("promise settled twice");
    p->state = 2;
    p->result = unwrap(rejection);
    return napi_ok;
}
napi_status napi_create_error(napi_env, napi_value, napi_value msg, napi_value* result) {
    Val* e = new_val(Val::Error);
    e->str = unwrap(msg)->str;
    *result = wrap(e);
    return napi_ok;
}
napi_status napi_throw_error(napi_env, const char* code, const char* msg) {
    g_exception = std::string(code ? code : "") + ": " + (msg ? msg : "");
    return napi_ok;
}
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback ex, napi_async_complete_callback co, void* data,
                                   napi_async_work* result) {
    Work* w = new Work();
    w->execute = ex;
    w->complete = co;
    w->data = data;
    *result = reinterpret_cast<napi_async_work>(w);
    return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work work) {
    Work* w = reinterpret_cast<Work*>(work);
    w->th = std::thread([w, env] { w->execute(env, w->data); });  // runs concurrently with the caller, like a libuv pool thread
    g_queue.push_back(w);
    return napi_ok;
}
napi_status napi_delete_async_work(napi_env, napi_async_work work) {
    delete reinterpret_cast<Work*>(work);
    return napi_ok;
}
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref* result) {
    Ref* r = new Ref{unwrap(value), initial_refcount};
    ++g_live_refs;
    *result = reinterpret_cast<napi_ref>(r);
    return napi_ok;
}
napi_status napi_delete_reference(napi_env, napi_ref ref) {
    delete reinterpret_cast<Ref*>(ref);
    --g_live_refs;
    return napi_ok;
}
napi_status napi_get_reference_value(napi_env, napi_ref ref, napi_value* result) {
    Ref* r = reinterpret_cast<Ref*>(ref);
    *result = wrap(r->v);
    return napi_ok;
}
napi_status napi_define_properties(napi_env, napi_value object, size_t n, const napi_property_descriptor* props) {
    Val* o = unwrap(object);
    for (size_t i = 0; i < n; ++i) {
        Val* f = new_val(Val::Func);
        f->cb = props[i].method;
        o->props[props[i].utf8name] = f;
    }
    return napi_ok;
}
void napi_module_register(napi_module*) {}

}  // extern "C"
