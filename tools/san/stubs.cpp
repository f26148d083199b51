// Development aid (tools/san_host_check.sh): the pieces of capi.cu that the host-only sources (config.cpp, pipe.cpp,
// audio.cpp) link against, without CUDA, so those sources can be built with -fsanitize=address,undefined and driven
// from tools/san/drive.py.  Not part of the product.
#include "../../glava_b200/csrc/internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace glb {
static thread_local std::string g_last_error;
static thread_local bool g_has_error = false;
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf; g_has_error = true;
    return code;
}
void clear_error() { g_has_error = false; }
bool has_error() { return g_has_error; }
}  // namespace glb
using namespace glb;

extern "C" {
const char* glava_b200_last_error(void) { return g_last_error.c_str(); }
int glava_b200_load_config_binds(glava_b200_params* out, const char* const* paths, const char* entry, const char* const* requests,
                                 const char* force_module, const char* const* binds) {
    return load_config(out, paths, entry, requests, force_module, binds);
}
void* glava_b200_host_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
void  glava_b200_host_free(void* p) { free(p); }
// a stand-in renderer for glava_b200_audio_frame / _fifo_pump / _pipe_apply: records what it was handed
struct glava_b200 { int batch; glava_b200_params p; long updates, ingests; double sum; };
glava_b200* san_renderer_new(int batch) { glava_b200* r = new glava_b200(); r->batch = batch; fill_defaults(&r->p, 0); return r; }
void san_renderer_free(glava_b200* r) { delete r; }
double san_renderer_sum(glava_b200* r) { return r->sum; }
long san_renderer_updates(glava_b200* r) { return r->updates; }
int glava_b200_batch(const glava_b200* r) { return r->batch; }
int glava_b200_sync(glava_b200*) { return 0; }
int glava_b200_update(glava_b200* r, const float* lb, const float* rb, size_t bsz, int) {
    for (size_t i = 0; i < (size_t) r->batch * bsz; ++i) r->sum += lb[i] + rb[i];      // touches every byte it may read
    ++r->updates; return 0;
}
int glava_b200_ingest_fifo(glava_b200* r, const int16_t* chunks, int frames) {
    for (size_t i = 0; i < (size_t) r->batch * frames * 2; ++i) r->sum += chunks[i];
    ++r->ingests; return 0;
}
int glava_b200_update_rings(glava_b200* r, int) { ++r->updates; return 0; }
int glava_b200_get_params(const glava_b200* r, glava_b200_params* out) { *out = r->p; return 0; }
int glava_b200_reconfigure(glava_b200* r, const glava_b200_params* p) { r->p = *p; return validate_params(p); }
}
