// napi_mock.h — TEST INFRASTRUCTURE: object model of the mock Node-API runtime (tests/napi_mock.cc).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../addon/napi_min.h"

namespace mock {

struct Val {
    enum Kind { Undef, Num, Str, Obj, ArrayBuffer, TypedArray, Promise, Error, Func } kind = Undef;
    double num = 0;
    std::string str;                   // Str, Error message
    std::map<std::string, Val*> props;  // Obj
    // ArrayBuffer
    void* data = nullptr;
    size_t len = 0;  // ArrayBuffer: bytes; TypedArray: elements
    std::vector<char> owned;
    napi_finalize fin = nullptr;
    void* hint = nullptr;
    // TypedArray
    int ta_type = 0;
    Val* ab = nullptr;
    size_t offset = 0;
    // Promise: 0 pending, 1 resolved, 2 rejected
    int state = 0;
    Val* result = nullptr;
    // Func
    napi_callback cb = nullptr;
};
struct Handle {
    Val* v;
    uint64_t scope;
};
struct CallInfo {
    std::vector<napi_value> argv;
};
struct Ref {
    Val* v;
    uint32_t count;
};
struct Work {
    napi_async_execute_callback execute;
    napi_async_complete_callback complete;
    void* data;
    std::thread th;
};

extern std::string g_exception;  // message of the last napi_throw_error during a call (empty: none)
extern int g_live_refs;          // napi_create_reference minus napi_delete_reference

Val* make_number(double d);
Val* make_string(const std::string& s);
Val* make_object();
Val* make_typedarray(int type, void* data, size_t length, size_t elem);  // a view over caller memory
napi_value wrap(Val* v);
void open_scope();
Val* call(Val* fn, const std::vector<Val*>& args);  // one native method call in its own handle scope
void drain();                                        // join the async work and run the completion callbacks
int gc();                                            // finalize external ArrayBuffers; returns how many

}  // namespace mock
