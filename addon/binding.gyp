{
  "targets": [{
    "target_name": "rm_napi",
    "sources": ["rm_napi.cc"],
    "defines": ["RM_HAVE_NODE_API_H", "NAPI_VERSION=8"],
    "cflags_cc": ["-std=c++17", "-fPIC"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../cpu_raymarcher_b200", "-lrm_b200", "-Wl,-rpath,<(module_root_dir)/../cpu_raymarcher_b200"]
  }]
}
