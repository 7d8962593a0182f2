/* napi_min.h — the handful of Node-API (N-API v8) declarations rm_napi.cc needs.
 * node_api.h is not present in this image (no Node.js); these prototypes restate the stable C ABI of
 * Node-API so the addon can be compiled (not run) here.  On a machine with Node.js, include <node_api.h>
 * instead (-DRM_HAVE_NODE_API_H). */
#ifndef NAPI_MIN_H
#define NAPI_MIN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_ref__* napi_ref;
typedef enum { napi_ok = 0 } napi_status;
typedef enum { napi_uint8_array = 1, napi_uint8_clamped_array = 2, napi_uint16_array = 4, napi_int32_array = 5, napi_float32_array = 9, napi_float64_array = 10 } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
typedef struct {
    const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
    napi_value value; int attributes; void* data;
} napi_property_descriptor;
typedef struct { int nm_version; unsigned int nm_flags; const char* nm_filename;
    napi_value (*nm_register_func)(napi_env, napi_value); const char* nm_modname; void* nm_priv; void* reserved[4]; } napi_module;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_named_property(napi_env, napi_value object, const char* name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* name, bool* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* name, napi_value value);
napi_status napi_get_value_double(napi_env, napi_value, double* result);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t* result);
napi_status napi_get_value_string_utf8(napi_env, napi_value, char* buf, size_t bufsize, size_t* result);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_double(napi_env, double, napi_value* result);
napi_status napi_create_int32(napi_env, int32_t, napi_value* result);
napi_status napi_create_string_utf8(napi_env, const char*, size_t, napi_value* result);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
napi_status napi_create_external_arraybuffer(napi_env, void* external_data, size_t byte_length, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_promise(napi_env, napi_deferred*, napi_value* promise);
napi_status napi_resolve_deferred(napi_env, napi_deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env, napi_deferred, napi_value rejection);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name, napi_async_execute_callback, napi_async_complete_callback, void* data, napi_async_work* result);
napi_status napi_queue_async_work(napi_env, napi_async_work);
napi_status napi_delete_async_work(napi_env, napi_async_work);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
napi_status napi_get_reference_value(napi_env, napi_ref ref, napi_value* result);
napi_status napi_define_properties(napi_env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
void napi_module_register(napi_module*);
#ifdef __cplusplus
}
#endif
#endif
