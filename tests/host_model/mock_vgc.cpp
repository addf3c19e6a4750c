// TEST-ONLY.  LD_PRELOAD shim that lets the REAL vechat_racon_b200 binary run on a box without a GPU so that the
// host side of the binding (csrc/racon_binding/b200polisher.cpp: device ranges, batching, packing, store, stitch) is
// checked in the CPU suite.  It interposes vgc_create / vgc_polish / vgc_destroy / vgc_last_error and answers
// vgc_polish with the CHECKER (oracle/_ref/libvechat_ref.so, the unmodified reference window code, path in
// MOCK_VGC_REF_SO).  Everything else (vgc_result_bound ...) still resolves to the real libvgc.so.  It never ships:
// the product has no CPU path, and tests/test_example_binary.py asserts the un-preloaded binary fails without a GPU.
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vgc.h"

namespace {
thread_local std::string g_err;
typedef int (*ref_polish_fn)(const vgc_batch*, const vgc_params*, vgc_result*, int);
ref_polish_fn ref_polish() {
  static ref_polish_fn fn = [] {
    const char* path = std::getenv("MOCK_VGC_REF_SO");
    void* so = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
    return so ? reinterpret_cast<ref_polish_fn>(dlsym(so, "ref_polish")) : nullptr;
  }();
  return fn;
}
std::atomic<int> g_handles{0};
}  // namespace

struct vgc_engine {
  vgc_params params;
  int device;
  int id;
};

extern "C" {

const char* vgc_last_error(void) { return g_err.c_str(); }

int vgc_create(vgc_handle* out, int device, const vgc_params* params) {
  if (!ref_polish()) {
    g_err = "mock: MOCK_VGC_REF_SO does not name a loadable libvechat_ref.so";
    return VGC_ERR_NO_DEVICE;
  }
  vgc_engine* h = new vgc_engine();
  h->params = *params;
  h->device = device;
  h->id = g_handles++;
  *out = h;
  return VGC_OK;
}

int vgc_destroy(vgc_handle h) {
  delete h;
  return VGC_OK;
}

int vgc_polish(vgc_handle h, const vgc_batch* batch, vgc_result* result, vgc_stats* stats) {
  std::fprintf(stderr, "[mock_vgc] handle %d device %d: %u windows, %u layers, %llu bytes\n", h->id, h->device,
               batch->n_windows, batch->n_layers, static_cast<unsigned long long>(batch->seq_off[batch->n_layers]));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const int rc = ref_polish()(batch, &h->params, result, 4);
  if (rc != 0) g_err = "mock: ref_polish failed";
  return rc;
}

}  // extern "C"
