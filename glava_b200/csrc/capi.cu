// capi.cu — the C ABI of include/glava_b200.h: renderer handle, device memory, streams.
//
// Plays the role of rd_new / rd_update / rd_destroy (reference glava/render.c:867,1743,2456)
// for a batch of independent streams on one CUDA device.  All device state (PCM staging,
// gravity/average state per stream and channel, R16 textures, framebuffers) is allocated once
// in glava_b200_new and stays resident in HBM.
#include "internal.h"
#include "raster_core.h"
#include "chain_core.h"
#include "tables.h"

#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace glb {

// ---- error plumbing -----------------------------------------------------------------------------
static thread_local std::string g_last_error;
static thread_local bool g_has_error = false;
static void default_hook(const char* msg) { fprintf(stderr, "glava_b200: %s\n", msg); }
static void (*g_hook)(const char*) = default_hook;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf; g_has_error = true;
    if (g_hook) g_hook(buf);
    return code;
}
void clear_error() { g_has_error = false; }
bool has_error() { return g_has_error; }

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(GLAVA_B200_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace glb

using namespace glb;

struct glava_b200 {
    glava_b200_params p;        // EFFECTIVE parameters the kernels run with: n = setbufsize / setbufscale, accel_fft
                                // cleared when a transform follows "fft" (render.c:2143-2154)
    glava_b200_params p_user;   // as given (glava_b200_get_params)
    int n_in;                   // setbufsize: floats per channel the caller hands to glava_b200_update
    bool post_chain;            // transform_smooth and / or keyframe lerp: the spectrum kernel stops at the float chain
                                // result; upload + K5 run as separate kernels
    bool interp_on;             // keyframe interpolation active (render.c:1761-1763, 2161-2168)
    int  kcounter;              // frames since the last modified update (render.c:2380-2383)
    float* d_scaled[2];         // bufscale output [batch][n] x {l, r}
    float* d_key[3]; int key_start, key_end;   // keyframe buffers [batch*2][n] (start, end, the one being written)
    float* d_spec_cur;          // latest post-transform buffer (glava_b200_spectrum)
    void* d_ts_tab; int ts_asz, ts_lim;        // transform_smooth {smin, smax} table
    std::atomic<unsigned long long> sizereq;   // pending glava_b200_sizereq: (1 << 63) | w << 32 | h, 0 = none
    int tap_ku; bool fused_k5, no_texmm;       // environment knobs, read once at creation
    int batch, device, slots;
    cudaStream_t stream;        // raster kernels, read-backs (the stream glava_b200_cuda_stream returns)
    cudaStream_t spec_stream;   // spectrum kernels + FIFO ingest, lowest priority: the latency-bound spectrum
                                // kernel of update i+1 runs under the HBM-bound raster kernel of update i
    int    tex_cur;             // which half of d_tex the latest spectrum wrote / the raster reads
    cudaEvent_t ev_spec_done[2], ev_raster_done[2];
    // inputs
    float* d_pcm[2][2];         // H2D staging for glava_b200_update, double-buffered  [2][batch][n] x {l, r}
    int    stage_cur;
    cudaEvent_t ev_input;       // glava_b200_input_event
    bool   async_input; int last_in;   // glava_b200_set_async_input: update() does not wait for its H2D copy
    cudaStream_t copy_stream;   // H2D of update i+1 overlaps the kernels of update i
    // asynchronous frame read-back: D2D into a staging frame on the raster stream (microseconds), D2H from there on its
    // own stream — the PCIe copy of frame i runs under raster i+1 instead of in front of it
    cudaStream_t out_stream; uint8_t* d_stage[2]; size_t stage_bytes; int out_cur;
    cudaEvent_t ev_stage_ready[2], ev_stage_free[2];
    cudaEvent_t ev_copied[2], ev_free[2];
    float* d_ring[2][2];        // FIFO rings, ping-pong                      [2][batch][n] x {l, r}
    int    ring_cur;
    void* d_chunks; size_t chunks_cap;
    // constants
    double* d_window; float* d_twiddle; void* d_rowtab; int* d_need; int need_count;
    TapEntry* d_tap_tab; int* d_tap_cnt; float* d_tap_wsum; int tap_max; int epi_n;
    K5Table k5; void* d_k5_blk; void* d_k5_ent; void* d_k5_wsum;   // full-plane K5 tap table (null: evaluate taps in the kernel)
    // need-list K5 as its own kernel (one table per channel): the serial per-texel sums run at full occupancy on
    // (texel, plane) pairs instead of on a sixth of the threads of one spectrum CTA
    void* d_need_blk; int need_nblk, need_max_rows;    // need-list K5 out of shared memory: blocks of sampled texels (tables.h build_need_blocks)
    bool k5_split_lazy, csr_in_smem, split_epilogue; int av_t_len; float* d_av_t; int spec_oop, spec_t;   // need-list K5 as its own kernel (k5_need_smem_kernel; k5_need_kernel + transposed copy as the fall-back)
    unsigned char* d_csr; int csr_bytes, csr_idx_off, csr_off_off;   // the same taps, texel-major, for the shared-memory path
    void* d_geo; int geo_box[4];   // polar geometry cache (radial / circle), see raster_kernels.cu
    void* d_ctile; int ctile_nx, ctile_count;   // circle: per-tile texel reference ranges (launch_circle_tiles)
    float* d_coltab; bool no_coltab;   // graph / wave: per-stream column table (launch_raster fills it), GLAVA_B200_NO_COLTAB
    uint32_t* d_texmm;             // circle: per-plane {min, max} of the sampled texture, refreshed before each raster
    // state + outputs
    float* d_spec; float* d_applied; float* d_ring_f;
    uint16_t* d_gr_store; uint16_t* d_ring_u; uint16_t* d_tex;
    uint16_t* d_av;             // pre-smoothing textures [batch*2][n] (full-plane K5 runs as its own kernel)
    uint8_t* d_fb;
    unsigned long long updates;
    // per-stream `modified` masks (glava_b200_update_masked): once streams have advanced unevenly, every stream has its
    // own average-ring cursor; words {bit 31 = modified, low bits = cursor} go to the device through a small pinned ring
    bool desync;
    std::vector<uint32_t> cursor;          // [batch] modified updates of that stream so far, mod avg_frames (valid once desync)
    uint32_t* h_umask[4]; uint32_t* d_umask[4]; cudaEvent_t ev_umask[4]; int umask_cur;
    uint64_t launches;
    std::vector<void*> allocs;
    // optional per-kernel device timing (glava_b200_set_timing): events around each launch
    bool timing;
    std::vector<cudaEvent_t> ev_spec; // pairs around each spectrum launch (spec_stream)
    std::vector<cudaEvent_t> ev_ras;  // pairs around each raster launch (stream)
};

// effective parameters and mode flags from the user's parameters
static void derive(glava_b200* r) {
    r->p = r->p_user;
    r->n_in = r->p_user.n;
    r->p.n = r->p_user.n / r->p_user.bufscale;
    const bool is_fft = r->p.module != GLAVA_B200_MOD_WAVE;
    if (r->p.transform_smooth == 1 && is_fft) r->p.accel_fft = 0;           // (2 = "smooth" before "fft": applied to the PCM)
    const float fr = r->p.fr > 0.0f ? r->p.fr : r->p.ur;
    r->interp_on = r->p.interpolate && !(r->p.accel_fft && is_fft) && (r->p.ur / fr) <= 0.9f;
    r->post_chain = r->p.transform_smooth == 1 || r->interp_on;
}

static int dev_alloc(glava_b200* r, void** out, size_t bytes, bool zero) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess)
        return fail(GLAVA_B200_ECUDA, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    r->allocs.push_back(*out);
    if (zero) CU(cudaMemsetAsync(*out, 0, bytes, r->stream));
    return 0;
}

extern "C" {

void glava_b200_set_abort_hook(void (*hook)(const char*)) { g_hook = hook ? hook : default_hook; }
const char* glava_b200_last_error(void) { return g_last_error.c_str(); }
const char* glava_b200_version(void) { return "glava_b200 0.1 (sm_100a)"; }

int glava_b200_default_params(glava_b200_params* out, const char* module) {
    clear_error();
    if (!out || !module) return fail(GLAVA_B200_EINVAL, "null argument");
    int m = module_from_name(module);
    if (m < 0) return fail(GLAVA_B200_ECONFIG, "Could not find module '%s'", module);
    fill_defaults(out, m);
    return GLAVA_B200_OK;
}

int glava_b200_load_config(glava_b200_params* out, const char* const* paths, const char* entry,
                           const char* const* requests, const char* force_module) {
    if (!out) return fail(GLAVA_B200_EINVAL, "null argument");
    return load_config(out, paths, entry, requests, force_module, nullptr);
}

int glava_b200_load_config_binds(glava_b200_params* out, const char* const* paths, const char* entry,
                                 const char* const* requests, const char* force_module, const char* const* binds) {
    if (!out) return fail(GLAVA_B200_EINVAL, "null argument");
    return load_config(out, paths, entry, requests, force_module, binds);
}

// ---- NUMA placement of pinned host memory -----------------------------------------------------------------------------
// On a two-socket box (the 8-GPU B200 node: GPUs 0-3 behind socket 0, 4-7 behind socket 1) a pinned buffer on the other
// socket makes every H2D / D2H DMA cross the inter-socket link; eight ranks doing so together was what collapsed round 1's
// end-to-end scaling.  Pinned buffers are therefore page-placed on the NUMA node of the device they feed: anonymous
// mapping + mbind(MPOL_BIND) + first touch + cudaHostRegister.  No libnuma in the image: the raw syscall.
static std::mutex g_host_mu;
static std::map<void*, size_t> g_host_maps;          // mmap'd + registered blocks (everything else came from cudaHostAlloc)

static int device_numa_node(int device) {
    char bus[32] = { 0 };
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char) (*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

extern "C" {

int glava_b200_device_numa_node(int device) { return device_numa_node(device); }
int glava_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Pin the calling thread to the CPUs of the device's NUMA node (what `numactl --cpunodebind` does for a one-rank-per-GPU
// process).  Returns the node, or -1 when the topology is not exposed (nothing changed).
int glava_b200_bind_thread_to_device(int device) {
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    cpu_set_t set; CPU_ZERO(&set);
    int a = 0, b = 0, any = 0;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int ch = fgetc(f);
        if (ch == '-') { if (fscanf(f, "%d", &b) != 1) b = a; ch = fgetc(f); }
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); any = 1; }
        if (ch != ',') break;
    }
    fclose(f);
    if (!any || sched_setaffinity(0, sizeof(set), &set) != 0) return -1;
    return node;
}

void* glava_b200_host_alloc_on(size_t bytes, int device) {
    if (bytes == 0) bytes = 16;
    const int node = getenv("GLAVA_B200_NO_NUMA") ? -1 : device_numa_node(device);
    if (node >= 0 && node < 64) {
        const size_t page = (size_t) sysconf(_SC_PAGESIZE);
        const size_t len = (bytes + page - 1) / page * page;
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask = 1ul << node;
            const long rc = syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, &mask, 64ul, 0u);
            if (rc == 0) {
                for (size_t o = 0; o < len; o += page) ((volatile char*) p)[o] = 0;      // first touch: the pages now exist on `node`
                if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                    std::lock_guard<std::mutex> lk(g_host_mu);
                    g_host_maps[p] = len;
                    return p;
                }
                cudaGetLastError();
            }
            munmap(p, len);
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        fail(GLAVA_B200_ECUDA, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void* glava_b200_host_alloc(size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return glava_b200_host_alloc_on(bytes, dev);
}
void glava_b200_host_free(void* p) {
    if (!p) return;
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_maps.find(p);
        if (it != g_host_maps.end()) { len = it->second; g_host_maps.erase(it); }
    }
    if (len) { cudaHostUnregister(p); munmap(p, len); }
    else cudaFreeHost(p);
}

}  // extern "C"

static void dev_free(glava_b200* r, void* ptr) {
    if (!ptr) return;
    for (size_t i = 0; i < r->allocs.size(); ++i) if (r->allocs[i] == ptr) { r->allocs.erase(r->allocs.begin() + i); break; }
    cudaFree(ptr);
}

// Upload a K5 tap table (tables.h) for k5_table_kernel; slots = {blk, ent, wsum, out} device pointers owned by the handle.
static int upload_k5_table(glava_b200* r, const K5TableHost& t, K5Table* out, void** slots, bool with_out) {
    static_assert(sizeof(K5Blk) == sizeof(int4) && sizeof(K5Ent) == sizeof(int2), "table records are read as int4 / int2");
    int rc;
    if ((rc = dev_alloc(r, &slots[0], t.blk.size() * sizeof(K5Blk), false)) != 0) return rc;
    if ((rc = dev_alloc(r, &slots[1], (t.ent.empty() ? 1 : t.ent.size()) * sizeof(K5Ent), false)) != 0) return rc;
    if ((rc = dev_alloc(r, &slots[2], t.wsum.size() * sizeof(float), false)) != 0) return rc;
    CU(cudaMemcpyAsync(slots[0], t.blk.data(), t.blk.size() * sizeof(K5Blk), cudaMemcpyHostToDevice, r->stream));
    if (!t.ent.empty()) CU(cudaMemcpyAsync(slots[1], t.ent.data(), t.ent.size() * sizeof(K5Ent), cudaMemcpyHostToDevice, r->stream));
    CU(cudaMemcpyAsync(slots[2], t.wsum.data(), t.wsum.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    if (with_out) {
        if ((rc = dev_alloc(r, &slots[3], (t.out.empty() ? 1 : t.out.size()) * sizeof(int), false)) != 0) return rc;
        if (!t.out.empty()) CU(cudaMemcpyAsync(slots[3], t.out.data(), t.out.size() * sizeof(int), cudaMemcpyHostToDevice, r->stream));
    }
    CU(cudaStreamSynchronize(r->stream));
    out->blk = (const int4*) slots[0]; out->ent = (const int2*) slots[1]; out->wsum = (const float*) slots[2];
    out->out = with_out ? (const int*) slots[3] : nullptr;
    out->count = (int) t.out.size(); out->max_span = t.max_span;
    out->smem_bytes = K5_S_PLANES * t.max_span * (int) sizeof(float);
    return 0;
}

// Full-plane K5 (every texel wanted: lazy_smooth = 0, circle, the optional-stage path): the taps of ALL n output
// texels (tables.h build_k5_table_host), uploaded for k5_table_kernel.
static int build_k5_table(glava_b200* r) {
    K5TableHost t;
    build_k5_table_host(r->p, &t);
    if ((size_t) 2 * t.max_span * sizeof(float) > 200 * 1024 || t.ent.size() * sizeof(K5Ent) > ((size_t) 256 << 20)) return 0;   // keep the in-kernel evaluation
    void* slots[4] = { nullptr, nullptr, nullptr, nullptr };
    int rc = upload_k5_table(r, t, &r->k5, slots, false);
    r->d_k5_blk = slots[0]; r->d_k5_ent = slots[1]; r->d_k5_wsum = slots[2];
    return rc;
}

// Everything derived from the parameters that does not depend on the audio: the lazy-K5 need-list and tap
// table, the bars / graph row-colour table, the polar geometry cache.  Called at creation and again by
// glava_b200_reconfigure (the analogue of a `--pipe` uniform update, render.c:1846-2005).
static int build_tables(glava_b200* r) {
    const glava_b200_params& p = r->p;
    int rc;
    dev_free(r, r->d_need); dev_free(r, r->d_tap_tab); dev_free(r, r->d_tap_cnt); dev_free(r, r->d_tap_wsum); dev_free(r, r->d_geo);
    dev_free(r, r->d_k5_blk); dev_free(r, r->d_k5_ent); dev_free(r, r->d_k5_wsum);
    r->d_k5_blk = r->d_k5_ent = r->d_k5_wsum = nullptr; memset(&r->k5, 0, sizeof(r->k5));
    dev_free(r, r->d_csr); r->d_csr = nullptr; r->csr_bytes = r->csr_idx_off = r->csr_off_off = 0;
    dev_free(r, r->d_av_t); r->d_av_t = nullptr;
    dev_free(r, r->d_need_blk); r->d_need_blk = nullptr; r->need_nblk = r->need_max_rows = 0;
    dev_free(r, r->d_ctile); r->d_ctile = nullptr; r->ctile_nx = r->ctile_count = 0;
    r->k5_split_lazy = false; r->csr_in_smem = false; r->av_t_len = 0;
    r->d_need = nullptr; r->need_count = 0; r->d_tap_tab = nullptr; r->d_tap_cnt = nullptr; r->d_tap_wsum = nullptr;
    r->tap_max = 0; r->epi_n = 0; r->d_geo = nullptr; r->geo_box[0] = r->geo_box[1] = r->geo_box[2] = r->geo_box[3] = 0;
    if (p.transform_smooth) {
        // transform_smooth's sampling window of output t (render.c:700-707) depends on t and the parameters only:
        // evaluated once on the host (chain_core.h), with the same libm calls and types as the reference
        const int sz = p.n;
        std::vector<SmoothWin> tab;
        r->ts_asz = transform_smooth_windows(sz, p.smooth_distance, p.smooth_ratio, &tab, &r->ts_lim);
        CU(cudaMemcpyAsync(r->d_ts_tab, tab.data(), (size_t) sz * sizeof(SmoothWin), cudaMemcpyHostToDevice, r->stream));
        CU(cudaStreamSynchronize(r->stream));
    }
    if (p.lazy_smooth && !r->post_chain) {
        std::vector<int> lists[2];
        if (build_need_list(p, lists)) {
            // need-list, K5 tap table (tap-major, for the L2 path) and its texel-major blob (tables.h)
            LazyTables t;
            build_lazy_tables(p, lists, &t);
            if ((rc = dev_alloc(r, (void**) &r->d_need, t.need.size() * sizeof(int), false)) != 0) return rc;
            CU(cudaMemcpyAsync(r->d_need, t.need.data(), t.need.size() * sizeof(int), cudaMemcpyHostToDevice, r->stream));
            CU(cudaStreamSynchronize(r->stream));
            r->need_count = (int) t.cnt;
            r->epi_n = getenv("GLAVA_B200_NO_EPI_PRUNE") ? 0 : t.epi_n;
            if (t.tab.size() * sizeof(TapEntry) <= (size_t) 64 << 20 && !getenv("GLAVA_B200_NO_TAPTAB")) {
                if ((rc = dev_alloc(r, (void**) &r->d_tap_tab, t.tab.size() * sizeof(TapEntry), false)) != 0) return rc;
                if ((rc = dev_alloc(r, (void**) &r->d_tap_cnt, t.tcnt.size() * sizeof(int), false)) != 0) return rc;
                if ((rc = dev_alloc(r, (void**) &r->d_tap_wsum, t.wsum.size() * sizeof(float), false)) != 0) return rc;
                CU(cudaMemcpyAsync(r->d_tap_tab, t.tab.data(), t.tab.size() * sizeof(TapEntry), cudaMemcpyHostToDevice, r->stream));
                CU(cudaMemcpyAsync(r->d_tap_cnt, t.tcnt.data(), t.tcnt.size() * sizeof(int), cudaMemcpyHostToDevice, r->stream));
                CU(cudaMemcpyAsync(r->d_tap_wsum, t.wsum.data(), t.wsum.size() * sizeof(float), cudaMemcpyHostToDevice, r->stream));
                CU(cudaStreamSynchronize(r->stream));
                r->tap_max = (int) t.tap_max;
                // When one channel's blob is small enough for two CTAs per SM to hold it in shared memory next to the FFT
                // buffers, the kernel's serial per-texel sums read their taps from there (no L2 round trips on the chain).
                const int base = spectrum_smem_bytes(p.n);
                const bool fits = base > 0 && base + t.blob <= (size_t) 112 * 1024 && !getenv("GLAVA_B200_NO_SMEM_TAPS");
                if ((rc = dev_alloc(r, (void**) &r->d_csr, t.csr.size(), false)) != 0) return rc;
                CU(cudaMemcpyAsync(r->d_csr, t.csr.data(), t.csr.size(), cudaMemcpyHostToDevice, r->stream));
                CU(cudaStreamSynchronize(r->stream));
                r->csr_bytes = (int) t.blob; r->csr_idx_off = (int) t.idx_off; r->csr_off_off = (int) t.off_off;
                r->csr_in_smem = fits;
                // Need-list K5 as its own kernel (k5_need_smem_kernel / k5_need_kernel, lanes = streams): default whenever the taps do not fit shared
                // memory next to the FFT, and from setbufsize 4096 up anyway; GLAVA_B200_K5_SPLIT=1 / 0 forces it on / off.
                const char* ks = getenv("GLAVA_B200_K5_SPLIT");
                r->k5_split_lazy = ks ? atoi(ks) != 0 : (!fits || p.n >= 4096);     // measured: also ahead at 4096 (759 k vs 753 k frames/s at the headline)
                r->av_t_len = t.epi_n > 0 ? t.epi_n : p.n;
                if (r->k5_split_lazy) {
                    // blocks of sampled texels for the shared-memory form; GLAVA_B200_K5N_SMEM=0 keeps the L2 form (k5_need_kernel)
                    NeedBlocks nb;
                    int target = -105, tpb = 16;         // tile rows: 105 % of the tallest window (tables.h), or an absolute number / -per cent
                    if (const char* e = getenv("GLAVA_B200_K5N_ROWS")) { const int v = atoi(e); if (v >= 32 || v < 0) target = v; }
                    if (const char* e = getenv("GLAVA_B200_K5N_TPB")) { const int v = atoi(e); if (v >= 1) tpb = v; }
                    build_need_blocks(t, p.n, target, tpb, &nb);
                    const char* ke = getenv("GLAVA_B200_K5N_SMEM");
                    if (nb.nblk > 0 && (size_t) nb.max_rows * 33 * sizeof(float) <= (size_t) 200 * 1024 && !(ke && atoi(ke) == 0)) {
                        if ((rc = dev_alloc(r, &r->d_need_blk, nb.blk.size() * sizeof(int), false)) != 0) return rc;
                        CU(cudaMemcpyAsync(r->d_need_blk, nb.blk.data(), nb.blk.size() * sizeof(int), cudaMemcpyHostToDevice, r->stream));
                        CU(cudaStreamSynchronize(r->stream));
                        r->need_nblk = nb.nblk; r->need_max_rows = nb.max_rows;
                    } else if ((rc = dev_alloc(r, (void**) &r->d_av_t, (size_t) 2 * r->av_t_len * r->batch * sizeof(float), true)) != 0) return rc;
                }
            }
        }
    }
    if (p.smooth_pass && (!r->d_need || r->post_chain) && !getenv("GLAVA_B200_NO_K5_TABLE")) {
        if ((rc = build_k5_table(r)) != 0) return rc;
    }
    if (p.module == GLAVA_B200_MOD_BARS || p.module == GLAVA_B200_MOD_GRAPH) { if ((rc = launch_bars_rowtab(p, r->d_rowtab, r->stream)) != 0) return rc; ++r->launches; }
    // polar modules: cache the audio-independent per-pixel geometry (circle: only when the module samples a
    // pre-smoothed texture, i.e. smooth_audio() is a single texelFetch)
    if (p.module == GLAVA_B200_MOD_RADIAL || (p.module == GLAVA_B200_MOD_CIRCLE && p.smooth_pass)) {
        const size_t bytes = polar_geo_box(p, r->geo_box);
        if (bytes > 0 && bytes <= ((size_t) 512 << 20) && !getenv("GLAVA_B200_NO_GEO")) {
            if ((rc = dev_alloc(r, &r->d_geo, bytes, false)) != 0) return rc;
            if ((rc = launch_polar_geo(p, r->d_geo, r->geo_box, r->stream)) != 0) return rc;
            ++r->launches;
            // per-tile annulus from bucketed texture min / max: measured a net LOSS at 1080p x 1024 streams, noise and
            // synthetic-music input alike (529 k vs 548 k frames/s: the curve's halo keeps most tiles of the annulus alive and every
            // live tile pays the bucket reduction) — kept behind GLAVA_B200_CTILE=1 so that the negative result is reproducible
            if (p.module == GLAVA_B200_MOD_CIRCLE && getenv("GLAVA_B200_CTILE")) {
                int ntx = 0, nty = 0;
                const size_t tb = circle_tile_bytes(p, &ntx, &nty);
                if ((rc = dev_alloc(r, &r->d_ctile, tb, false)) != 0) return rc;
                if ((rc = launch_circle_tiles(p, r->d_geo, r->geo_box, r->d_ctile, r->stream)) != 0) return rc;
                r->ctile_nx = ntx; r->ctile_count = ntx * nty;
                r->launches += 2;
            }
        }
    }
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

static int build(glava_b200* r) {
    const glava_b200_params& p = r->p;
    const size_t n = (size_t) p.n, planes = (size_t) r->batch * 2, F = (size_t) p.avg_frames;
    CU(cudaSetDevice(r->device));
    CU(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        // Spectrum stream priority.  Measured on B200 (bars, 1024 streams/GPU, whole step): low/equal 705 k
        // frames/s with the raster kernel at 0.96 of the HBM peak while co-running; high 703 k with the
        // raster kernel at 0.89 (the spectrum CTAs then grab most of each SM's registers first); no
        // overlap at all 680 k.  Default: lowest.  GLAVA_B200_SPEC_PRIO=high|equal overrides.
        const char* pr = getenv("GLAVA_B200_SPEC_PRIO");
        int prio = lo;
        if (pr && !strcmp(pr, "high")) prio = hi; else if (pr && !strcmp(pr, "equal")) prio = 0;
        if (getenv("GLAVA_B200_NO_OVERLAP")) r->spec_stream = r->stream;   // tuning aid: serialise spectrum and raster
        else CU(cudaStreamCreateWithPriority(&r->spec_stream, cudaStreamNonBlocking, prio));
    }
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&r->ev_spec_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&r->ev_raster_done[i], cudaEventDisableTiming));
    }
    int rc;
#define ALLOC(ptr, bytes, zero) if ((rc = dev_alloc(r, (void**) &(ptr), (bytes), (zero))) != 0) return rc
    CU(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
    const size_t n_in = (size_t) r->n_in;
    for (int i = 0; i < 2; ++i) {
        ALLOC(r->d_pcm[i][0], (size_t) r->batch * n_in * 4, true); ALLOC(r->d_pcm[i][1], (size_t) r->batch * n_in * 4, true);
        CU(cudaEventCreateWithFlags(&r->ev_copied[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&r->ev_free[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) for (int c = 0; c < 2; ++c) ALLOC(r->d_ring[i][c], (size_t) r->batch * n_in * 4, true);
    if (r->p_user.bufscale > 1 || r->p_user.transform_smooth == 2) { ALLOC(r->d_scaled[0], (size_t) r->batch * n * 4, true); ALLOC(r->d_scaled[1], (size_t) r->batch * n * 4, true); }
    if (r->interp_on) {
        for (int i = 0; i < 3; ++i) ALLOC(r->d_key[i], planes * n * 4, true);      // keyframes start at 0 (render.c:1681 calloc)
        r->key_start = 0; r->key_end = 1;
    }
    if (r->p.transform_smooth) ALLOC(r->d_ts_tab, n * sizeof(int2), true);
    ALLOC(r->d_window, n * 8, false); ALLOC(r->d_twiddle, n * 4, false);
    ALLOC(r->d_spec, planes * n * 4, true);
    r->d_spec_cur = r->d_spec;
    // gravity + average state: ONE allocation, so a single L2 access-policy window can cover it (below)
    char* state = nullptr; size_t state_bytes = 0;
    if (p.accel_fft) {
        state_bytes = planes * n * 2 * (1 + F);
        ALLOC(state, state_bytes, true);
        r->d_gr_store = (uint16_t*) state; r->d_ring_u = (uint16_t*) (state + planes * n * 2);
    } else {
        state_bytes = planes * n * 4 * (1 + F);
        ALLOC(state, state_bytes, true);
        r->d_applied = (float*) state; r->d_ring_f = (float*) (state + planes * n * 4);
    }
    // EXPERIMENT, off by default (GLAVA_B200_L2_PERSIST=1): pin the spectrum kernel's read-modify-write
    // state in the 126 MB L2 with a persisting access window, so that its DRAM reads do not get mixed into
    // the raster kernel's write stream.  Measured on B200: the persisting carve-out takes L2 away from
    // the raster kernel's write-back path and HALVES its store bandwidth (1.03 -> 0.44 of the HBM peak,
    // 702 k -> 328 k frames/s).  Kept only so the negative result is reproducible.
    if (getenv("GLAVA_B200_L2_PERSIST")) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, r->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            size_t want = state_bytes < (size_t) prop.persistingL2CacheMaxSize ? state_bytes : (size_t) prop.persistingL2CacheMaxSize;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            size_t win = state_bytes < (size_t) prop.accessPolicyMaxWindowSize ? state_bytes : (size_t) prop.accessPolicyMaxWindowSize;
            attr.accessPolicyWindow.base_ptr = state;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = win > 0 ? (float) ((double) want / (double) win) : 0.0f;
            if (attr.accessPolicyWindow.hitRatio > 1.0f) attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(r->spec_stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
        }
    }
    ALLOC(r->d_tex, 2 * planes * n * 2, true);
    ALLOC(r->d_av, planes * n * 2, true);
    ALLOC(r->d_texmm, planes * GLB_TEXMM_STRIDE * sizeof(uint32_t), true);          // double-buffered: spectrum i+1 writes one half while raster i reads the other
    ALLOC(r->d_rowtab, (size_t) p.h * 8, true);
    if (p.module == GLAVA_B200_MOD_GRAPH || p.module == GLAVA_B200_MOD_WAVE)          // (the module of a handle never changes: reconfigure refuses)
        ALLOC(r->d_coltab, (size_t) r->batch * GLB_COLTAB_PLANES * p.w * sizeof(float), true);
    // framebuffers: [slots][h][w] RGBA8
    size_t frame = (size_t) p.w * p.h * 4;
    r->slots = (p.fb_slots > 0 && p.fb_slots < r->batch) ? p.fb_slots : r->batch;
    ALLOC(r->d_fb, frame * r->slots, true);
#undef ALLOC
    // window LUT: render.c:660 macro as expanded at render.c:794 — cos(TWOPI*i/N - 1), double
    std::vector<double> w(n);
    for (size_t i = 0; i < n; ++i) w[i] = 0.53836 - (0.46164 * cos(6.28318530718 * (double) i / (double) n - 1));
    CU(cudaMemcpyAsync(r->d_window, w.data(), n * 8, cudaMemcpyHostToDevice, r->stream));
    // twiddles exp(-2*pi*i*k/M), M = n/2, evaluated in double
    const size_t M = n / 2;
    std::vector<float> tw(2 * M);
    for (size_t k = 0; k < M; ++k) {
        double ang = -2.0 * M_PI * (double) k / (double) M;
        tw[2 * k] = (float) cos(ang); tw[2 * k + 1] = (float) sin(ang);
    }
    CU(cudaMemcpyAsync(r->d_twiddle, tw.data(), n * 4, cudaMemcpyHostToDevice, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    if ((rc = build_tables(r)) != 0) return rc;
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

glava_b200* glava_b200_new(const glava_b200_params* params, int batch, int device) {
    clear_error();
    if (!params || batch < 1) { fail(GLAVA_B200_EINVAL, "glava_b200_new: bad arguments"); return nullptr; }
    if (validate_params(params) != 0) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fail(GLAVA_B200_ECUDA, "no CUDA device available: the B200 path has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(GLAVA_B200_EINVAL, "device %d out of range (%d devices)", device, ndev); return nullptr; }
    glava_b200* r = new glava_b200();
    r->p_user = *params; r->batch = batch; r->device = device;
    derive(r);
    r->sizereq.store(0);
    { const char* e = getenv("GLAVA_B200_TAP_KU"); r->tap_ku = e ? atoi(e) : 8; }
    r->fused_k5 = getenv("GLAVA_B200_FUSED_K5") != nullptr;
    // Launch structure of the lazy pipeline-B update: wherever the need-list K5 runs as its own kernel (auto:
    // setbufsize >= 8192) the R16 state update does too, and at 8192 the FFT passes run out of place with 256 threads —
    // the measured optima (tools/variant_sweep.sh).  GLAVA_B200_SPLIT_EPI / _SPEC_OOP / _SPEC_T override.
    { const char* e = getenv("GLAVA_B200_SPLIT_EPI"); r->split_epilogue = e ? atoi(e) != 0 : true; }
    { const char* e = getenv("GLAVA_B200_SPEC_OOP"); r->spec_oop = e ? atoi(e) : (params->n == 8192 ? 1 : 0); }
    { const char* e = getenv("GLAVA_B200_SPEC_T"); r->spec_t = e ? atoi(e) : (params->n == 8192 ? 256 : 0); }
    r->no_texmm = getenv("GLAVA_B200_NO_TEXMM") != nullptr;
    r->no_coltab = getenv("GLAVA_B200_NO_COLTAB") != nullptr;
    r->kcounter = 0; r->d_scaled[0] = r->d_scaled[1] = nullptr; r->d_key[0] = r->d_key[1] = r->d_key[2] = nullptr;
    r->key_start = 0; r->key_end = 1; r->d_spec_cur = nullptr; r->d_ts_tab = nullptr; r->ts_asz = r->ts_lim = 0;
    r->stream = nullptr; r->spec_stream = nullptr; r->tex_cur = 0; r->ring_cur = 0;
    for (int i = 0; i < 2; ++i) { r->ev_spec_done[i] = nullptr; r->ev_raster_done[i] = nullptr; } r->d_chunks = nullptr; r->chunks_cap = 0;
    r->d_window = nullptr; r->d_twiddle = nullptr; r->d_rowtab = nullptr; r->d_need = nullptr; r->need_count = 0;
    r->d_tap_tab = nullptr; r->d_tap_cnt = nullptr; r->d_tap_wsum = nullptr; r->tap_max = 0; r->epi_n = 0;
    r->d_csr = nullptr; r->csr_bytes = r->csr_idx_off = r->csr_off_off = 0;
    r->d_k5_blk = r->d_k5_ent = r->d_k5_wsum = nullptr; memset(&r->k5, 0, sizeof(r->k5));
    r->k5_split_lazy = false; r->csr_in_smem = false; r->av_t_len = 0; r->d_av_t = nullptr;
    r->d_need_blk = nullptr; r->need_nblk = r->need_max_rows = 0;
    r->d_ctile = nullptr; r->ctile_nx = r->ctile_count = 0;
    r->d_geo = nullptr; r->geo_box[0] = r->geo_box[1] = r->geo_box[2] = r->geo_box[3] = 0;
    r->d_spec = r->d_applied = r->d_ring_f = nullptr; r->d_gr_store = r->d_ring_u = r->d_tex = nullptr; r->d_av = nullptr; r->d_texmm = nullptr; r->d_coltab = nullptr; r->d_fb = nullptr;
    for (int i = 0; i < 2; ++i) { r->d_pcm[i][0] = r->d_pcm[i][1] = nullptr; r->ev_copied[i] = r->ev_free[i] = nullptr; }
    r->stage_cur = 0; r->copy_stream = nullptr;
    r->out_stream = nullptr; r->d_stage[0] = r->d_stage[1] = nullptr; r->stage_bytes = 0; r->out_cur = 0;
    for (int i = 0; i < 2; ++i) { r->ev_stage_ready[i] = nullptr; r->ev_stage_free[i] = nullptr; }
    r->updates = 0; r->launches = 0; r->timing = false;
    r->desync = false; r->umask_cur = 0; r->async_input = false; r->last_in = -1; r->ev_input = nullptr;
    for (int i = 0; i < 4; ++i) { r->h_umask[i] = nullptr; r->d_umask[i] = nullptr; r->ev_umask[i] = nullptr; }
    if (build(r) != 0) { glava_b200_destroy(r); return nullptr; }
    return r;
}

void glava_b200_destroy(glava_b200* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->spec_stream) cudaStreamSynchronize(r->spec_stream);
    if (r->stream) cudaStreamSynchronize(r->stream);
    for (cudaEvent_t e : r->ev_spec) cudaEventDestroy(e);
    for (cudaEvent_t e : r->ev_ras) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { if (r->ev_spec_done[i]) cudaEventDestroy(r->ev_spec_done[i]); if (r->ev_raster_done[i]) cudaEventDestroy(r->ev_raster_done[i]); }
    if (r->spec_stream && r->spec_stream != r->stream) cudaStreamDestroy(r->spec_stream);
    for (void* p : r->allocs) cudaFree(p);
    if (r->d_chunks) cudaFree(r->d_chunks);
    for (int i = 0; i < 2; ++i) { if (r->ev_copied[i]) cudaEventDestroy(r->ev_copied[i]); if (r->ev_free[i]) cudaEventDestroy(r->ev_free[i]); }
    for (int i = 0; i < 4; ++i) {
        if (r->h_umask[i]) cudaFreeHost(r->h_umask[i]);
        if (r->d_umask[i]) cudaFree(r->d_umask[i]);
        if (r->ev_umask[i]) cudaEventDestroy(r->ev_umask[i]);
    }
    if (r->ev_input) cudaEventDestroy(r->ev_input);
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    if (r->out_stream) { cudaStreamSynchronize(r->out_stream); cudaStreamDestroy(r->out_stream); }
    for (int i = 0; i < 2; ++i) {
        if (r->d_stage[i]) cudaFree(r->d_stage[i]);
        if (r->ev_stage_ready[i]) cudaEventDestroy(r->ev_stage_ready[i]);
        if (r->ev_stage_free[i]) cudaEventDestroy(r->ev_stage_free[i]);
    }
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

int glava_b200_reconfigure(glava_b200* r, const glava_b200_params* params) {
    clear_error();
    if (!r || !params) return fail(GLAVA_B200_EINVAL, "glava_b200_reconfigure: null argument");
    int rc = validate_params(params);
    if (rc) return rc;
    const glava_b200_params& o = r->p_user;
    glava_b200 probe; probe.p_user = *params; derive(&probe);
    if (params->n != o.n || params->w != o.w || params->h != o.h || params->module != o.module || params->accel_fft != o.accel_fft ||
        params->avg_frames != o.avg_frames || params->fb_slots != o.fb_slots || params->bufscale != o.bufscale ||
        params->transform_smooth != o.transform_smooth || probe.interp_on != r->interp_on)
        return fail(GLAVA_B200_EINVAL, "glava_b200_reconfigure: setbufsize, setbufscale, geometry, module, setaccelfft, setavgframes, "
                                       "fb_slots, the \"smooth\" transform and whether interpolation is active size the device "
                                       "state and cannot change on a live renderer; create a new one");
    CU(cudaSetDevice(r->device));
    CU(cudaStreamSynchronize(r->spec_stream));
    CU(cudaStreamSynchronize(r->stream));
    r->p_user = *params;
    derive(r);
    return build_tables(r);
}

int glava_b200_get_params(const glava_b200* r, glava_b200_params* out) {
    if (!r || !out) return fail(GLAVA_B200_EINVAL, "null argument");
    *out = r->p_user; return 0;
}
int glava_b200_batch(const glava_b200* r) { return r ? r->batch : 0; }
const char* glava_b200_module_name(const glava_b200* r) { return r ? module_name(r->p.module) : "?"; }
int glava_b200_spectrum_size(const glava_b200* r) { return r ? r->p.n : 0; }
const void* glava_b200_framebuffer_device(const glava_b200* r) { return r ? r->d_fb : nullptr; }
void* glava_b200_cuda_stream(const glava_b200* r) { return r ? (void*) r->stream : nullptr; }
uint64_t glava_b200_launch_count(const glava_b200* r) { return r ? r->launches : 0; }

static int timing_mark(std::vector<cudaEvent_t>& v, cudaStream_t st) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    CU(cudaEventRecord(e, st));
    v.push_back(e);
    return 0;
}

static uint16_t* tex_half(glava_b200* r, int b) { return r->d_tex + (size_t) b * r->batch * 2 * r->p.n; }
static int sync_all(glava_b200* r) {
    CU(cudaStreamSynchronize(r->spec_stream));
    CU(cudaStreamSynchronize(r->stream));
    if (r->out_stream) CU(cudaStreamSynchronize(r->out_stream));
    return 0;
}

// Runtime resize (render.c:1811-1830 resizes the stage FBOs when the framebuffer size changed; offscreen consumers ask
// through glava_sizereq, glava.c:263-267).  Spectrum state is untouched; the framebuffers and every table that depends
// on the geometry (row colours, polar geometry cache, lazy-K5 need-list) are rebuilt.
static int apply_resize(glava_b200* r, int w, int h) {
    glava_b200_params q = r->p_user;
    q.w = w; q.h = h;
    int rc = validate_params(&q);
    if (rc) return rc;
    if ((rc = sync_all(r)) != 0) return rc;
    dev_free(r, r->d_fb); dev_free(r, r->d_rowtab); dev_free(r, r->d_coltab);
    r->d_fb = nullptr; r->d_rowtab = nullptr; r->d_coltab = nullptr;
    r->p_user = q;
    derive(r);
    if ((rc = dev_alloc(r, (void**) &r->d_rowtab, (size_t) q.h * 8, true)) != 0) return rc;
    if ((q.module == GLAVA_B200_MOD_GRAPH || q.module == GLAVA_B200_MOD_WAVE) &&
        (rc = dev_alloc(r, (void**) &r->d_coltab, (size_t) r->batch * GLB_COLTAB_PLANES * q.w * sizeof(float), true)) != 0) return rc;
    if ((rc = dev_alloc(r, (void**) &r->d_fb, (size_t) q.w * q.h * 4 * r->slots, true)) != 0) return rc;
    return build_tables(r);
}

// One update = spectrum kernel on spec_stream (if modified) + raster kernel on stream.
//   spectrum(i) waits for: its input (caller-provided event on spec_stream), spectrum(i-1) (same stream:
//                          gravity / average state is read-modify-write), raster(i-2) (last reader of the
//                          texture half it overwrites)
//   raster(i)   waits for: spectrum(i)
// so raster(i) and spectrum(i+1) run concurrently: one is HBM-store bound, the other latency bound.
// Per-stream mask -> device words.  Returns the device array for this update (nullptr: every stream modified in lock step).
static int stage_umask(glava_b200* r, const uint8_t* mask, const uint32_t** d_out) {
    const int F = r->p.avg_frames;
    if (!r->desync) {                                           // first uneven update: every stream starts from the shared cursor
        r->cursor.assign((size_t) r->batch, (uint32_t) (r->updates % (unsigned long long) F));
        r->desync = true;
    }
    const int k = r->umask_cur;
    if (!r->h_umask[k]) {
        CU(cudaHostAlloc((void**) &r->h_umask[k], (size_t) r->batch * 4, cudaHostAllocDefault));
        CU(cudaMalloc((void**) &r->d_umask[k], (size_t) r->batch * 4));
        CU(cudaEventCreateWithFlags(&r->ev_umask[k], cudaEventDisableTiming));
    } else {
        CU(cudaEventSynchronize(r->ev_umask[k]));               // the copy that last read this pinned buffer (4 updates ago)
    }
    for (int s = 0; s < r->batch; ++s) {
        const bool m = !mask || mask[s];
        r->h_umask[k][s] = (m ? 0x80000000u : 0u) | r->cursor[s];
        if (m) r->cursor[s] = (r->cursor[s] + 1u) % (uint32_t) F;
    }
    CU(cudaMemcpyAsync(r->d_umask[k], r->h_umask[k], (size_t) r->batch * 4, cudaMemcpyHostToDevice, r->spec_stream));
    CU(cudaEventRecord(r->ev_umask[k], r->spec_stream));
    r->umask_cur = (k + 1) & 3;
    *d_out = r->d_umask[k];
    return 0;
}

// `mask` (may be null): one byte per stream, non-zero = that stream has new audio (glava.c:528-537 per renderer).
static int run_update(glava_b200* r, const float* d_l, const float* d_r, int modified, bool raster_only = false,
                      const uint8_t* mask = nullptr) {
    int rc;
    if (mask && !raster_only) {
        int cnt = 0;
        for (int s = 0; s < r->batch; ++s) cnt += mask[s] ? 1 : 0;
        modified = cnt > 0;
        if (cnt == 0 || cnt == r->batch) mask = nullptr;        // nobody / everybody: the plain forms
        else if (r->interp_on)
            return fail(GLAVA_B200_EINVAL, "per-stream modified masks are not available with keyframe interpolation active "
                                           "(setinterpolate): the keyframe rotation is per renderer; update the streams in lock step");
    } else mask = nullptr;
    if (unsigned long long req = r->sizereq.exchange(0)) {                  // render.c:1811-1815: at the start of a frame
        if ((rc = apply_resize(r, (int) ((req >> 32) & 0x7fffffffu), (int) (req & 0xffffffffu))) != 0) return rc;
    }
    if (!r->d_fb) return fail(GLAVA_B200_ECUDA, "renderer has no framebuffer (a resize failed)");
    const glava_b200_params& p = r->p;
    const bool new_tex = !raster_only && (modified || r->interp_on);   // an interpolated frame has a new texture without new audio
    const int planes = r->batch * 2;
    const size_t total = (size_t) planes * p.n;
    if (new_tex) CU(cudaStreamWaitEvent(r->spec_stream, r->ev_raster_done[r->tex_cur ^ 1], 0));
    if (modified && r->p_user.bufscale > 1) {
        const int chans = p.module == GLAVA_B200_MOD_WAVE ? 1 : 2;
        if ((rc = launch_bufscale(d_l, d_r, r->d_scaled[0], r->d_scaled[1], r->batch, r->n_in, r->p_user.bufscale, chans, r->spec_stream)) != 0) return rc;
        ++r->launches;
        d_l = r->d_scaled[0]; d_r = r->d_scaled[1];
    }
    if (modified && p.transform_smooth == 2) {
        // `#request transform <u> "smooth"` listed BEFORE "fft" (render.c:1218-1286): handle_audio applies it on the CPU to
        // the (scaled) PCM ring, then meets "fft" and carries on as usual (render.c:2131-2156).  Here: on a copy of the
        // rings (rd_update transforms the caller's buffers in place; this ABI leaves them alone)
        const int chans = p.module == GLAVA_B200_MOD_WAVE ? 1 : 2;
        const float* src[2] = { d_l, d_r };
        for (int c = 0; c < chans; ++c) {
            if (src[c] != r->d_scaled[c]) CU(cudaMemcpyAsync(r->d_scaled[c], src[c], (size_t) r->batch * p.n * 4, cudaMemcpyDeviceToDevice, r->spec_stream));
            if ((rc = launch_transform_smooth(r->d_scaled[c], p.n, r->d_ts_tab, r->ts_asz, r->ts_lim, r->batch, r->spec_stream, nullptr, 0)) != 0) return rc;
            ++r->launches;
        }
        d_l = r->d_scaled[0]; d_r = chans == 2 ? r->d_scaled[1] : r->d_scaled[0];
    }
    float* chain_out = r->d_spec;                         // where the float chain result of this update goes
    if (r->interp_on) {
        int nxt = 0; while (nxt == r->key_start || nxt == r->key_end) ++nxt;
        chain_out = r->d_key[nxt];
    }
    if (modified) {
        const int b = r->tex_cur ^ 1;
        SpectrumArgs a;
        memset(&a, 0, sizeof(a));
        a.pcm_l = d_l; a.pcm_r = d_r; a.window = r->d_window; a.twiddle = r->d_twiddle;
        a.spec = chain_out; a.skip_tex = r->post_chain ? 1 : 0; a.applied = r->d_applied; a.ring_f = r->d_ring_f;
        a.gr_store = r->d_gr_store; a.ring_u = r->d_ring_u; a.tex = tex_half(r, b);
        const bool split_lazy = r->k5_split_lazy && p.lazy_smooth && r->d_need && !r->post_chain && p.smooth_pass;
        a.need = (p.lazy_smooth && r->d_need && !split_lazy) ? r->d_need : nullptr; a.need_count = r->need_count;
        a.tap_tab = a.need ? r->d_tap_tab : nullptr; a.tap_cnt = r->d_tap_cnt; a.tap_wsum = r->d_tap_wsum; a.tap_max = r->tap_max;
        a.epi_n = (a.need || split_lazy) ? r->epi_n : 0;
        a.tap_ku = r->tap_ku;
        a.csr = (a.need && a.tap_tab && r->csr_in_smem) ? r->d_csr : nullptr; a.csr_bytes = r->csr_bytes; a.csr_idx_off = r->csr_idx_off; a.csr_off_off = r->csr_off_off;
        a.batch = r->batch; a.update = r->updates;
        a.variant_oop = r->spec_oop; a.variant_t = r->spec_t;
        a.umask = nullptr; a.tex_prev = tex_half(r, r->tex_cur);
        if (mask || r->desync) { if ((rc = stage_umask(r, mask, &a.umask)) != 0) return rc; }
        const int F = p.avg_frames;
        for (int f = 0; f < F; ++f) {
            // pipeline A: window_frame(f, avg_frames - 1) -> cos(TWOPI*f/F - 1), double (render.c:661,766)
            a.avg_w_a[f] = 0.6 - (0.4 * cos(6.28318530718 * (double) f / (double) F - 1));
            // pipeline B: window(I, _AVG_FRAMES - 1) -> cos(TWOPI*I/F - 1), GLSL float (average_pass.frag:41)
            a.avg_w_b[f] = 0.53836f - (0.46164f * cosf(6.28318530718f * (float) f / (float) F - 1.0f));
        }
        a.avg_b_windowed = (p.avg_window && F != 2) ? 1 : 0;
        const bool is_fft = p.module != GLAVA_B200_MOD_WAVE;
        if (r->timing && (rc = timing_mark(r->ev_spec, r->spec_stream)) != 0) return rc;
        // full-plane smoothing (every texel wanted): the spectrum kernel exports the pre-smoothing texture and
        // a second kernel smooths all planes, sharing the tap weights between planes
        const bool split_k5 = p.smooth_pass && !a.need && !r->post_chain && !r->fused_k5 && !split_lazy;
        a.av_out = (split_k5 || split_lazy) ? r->d_av : nullptr;
        a.av_t_len = split_lazy ? r->av_t_len : 0;
        // three-kernel form of the lazy pipeline-B update: transform_fft only in the plane-per-CTA kernel, the R16 state
        // update as an elementwise kernel, K5 on (texel, stream) pairs.  (F = 1 has no ring: in-kernel form.)
        const int epi8 = ((r->epi_n + 7) / 8) * 8;
        const bool split_epi = split_lazy && is_fft && p.accel_fft && r->split_epilogue && epi8 > 0 && epi8 <= p.n && F >= 1;
        a.fft_only = split_epi ? 1 : 0;
        if (split_epi) a.epi_n = epi8;
        if ((rc = launch_spectrum(p, a, is_fft, r->spec_stream)) != 0) return rc;
        if (split_epi) {
            if ((rc = launch_epilogue_b(p, a, epi8, r->spec_stream)) != 0) return rc;
            ++r->launches;
        }
        if (split_lazy) {
            if (r->d_need_blk) {
                if ((rc = launch_k5_need_smem(p, r->d_av, a.tex, r->batch, is_fft ? 2 : 1, r->d_csr, r->csr_bytes, r->csr_idx_off, r->csr_off_off,
                                              r->d_need, r->d_tap_wsum, r->need_count, r->d_need_blk, r->need_nblk, r->need_max_rows, r->spec_stream)) != 0) return rc;
                r->launches += 1;
            } else {
                if ((rc = launch_k5_need(p, r->d_av, r->d_av_t, r->av_t_len, a.tex, r->batch, is_fft ? 2 : 1, r->d_csr, r->csr_bytes, r->csr_idx_off,
                                         r->csr_off_off, r->d_need, r->d_tap_wsum, r->need_count, r->spec_stream)) != 0) return rc;
                r->launches += 2;
            }
        }
        if (split_k5) {
            // wave uses plane 0 of each stream only; smoothing the (zero) odd planes too keeps the launch simple
            if ((rc = launch_smooth_only(p, r->d_av, a.tex, r->batch * 2, r->spec_stream, &r->k5)) != 0) return rc;
            ++r->launches;
        }
        if (r->post_chain && p.transform_smooth == 1) {                     // render.c:694-718, after the module's chain
            if ((rc = launch_transform_smooth(chain_out, p.n, r->d_ts_tab, r->ts_asz, r->ts_lim, planes, r->spec_stream, a.umask)) != 0) return rc;
            ++r->launches;
        }
        if (!r->post_chain) {
            if (r->timing && (rc = timing_mark(r->ev_spec, r->spec_stream)) != 0) return rc;
            CU(cudaEventRecord(r->ev_spec_done[b], r->spec_stream));
            r->tex_cur = b;
        }
        ++r->launches; ++r->updates;
        r->d_spec_cur = chain_out;
    }
    if (r->post_chain && new_tex) {
        // R16 upload of the buffer this frame shows (render.c:2185) and K5 (render.c:2276-2303).  With keyframe
        // interpolation that is the lerp of the two PREVIOUS post-transform buffers (render.c:1792-1809): the
        // update just computed becomes visible one update later (rc.glsl:129-130).
        const int b = r->tex_cur ^ 1;
        uint16_t* dst = p.smooth_pass ? r->d_av : tex_half(r, b);
        if (r->interp_on) {
            const float mod = keyframe_mod(p.ur, p.fr > 0.0f ? p.fr : p.ur, r->kcounter);   // render.c:1761,1804
            rc = launch_upload(r->d_key[r->key_start], r->d_key[r->key_end], mod, dst, total, r->spec_stream);
        } else {
            rc = launch_upload(chain_out, nullptr, 0.0f, dst, total, r->spec_stream);
        }
        if (rc) return rc;
        ++r->launches;
        if (p.smooth_pass) {
            if ((rc = launch_smooth_only(p, r->d_av, tex_half(r, b), planes, r->spec_stream, &r->k5)) != 0) return rc;
            ++r->launches;
        }
        if (modified && r->timing && (rc = timing_mark(r->ev_spec, r->spec_stream)) != 0) return rc;
        CU(cudaEventRecord(r->ev_spec_done[b], r->spec_stream));
        r->tex_cur = b;
        if (r->interp_on && modified) {                                       // render.c:2347-2353: start <- end <- this update
            int nxt = 0; while (nxt == r->key_start || nxt == r->key_end) ++nxt;
            r->key_start = r->key_end; r->key_end = nxt;
        }
    }
    if (!raster_only) r->kcounter = modified ? 0 : r->kcounter + 1;           // render.c:2380-2383
    const int b = r->tex_cur;
    CU(cudaStreamWaitEvent(r->stream, r->ev_spec_done[b], 0));
    RasterArgs ra;
    ra.tex = tex_half(r, b); ra.fb = r->d_fb;
    ra.rowtab = (p.module == GLAVA_B200_MOD_BARS || p.module == GLAVA_B200_MOD_GRAPH) ? r->d_rowtab : nullptr;
    ra.batch = r->batch; ra.slots = r->slots; ra.stream0 = 0;
    ra.geo = r->d_geo; ra.gx0 = r->geo_box[0]; ra.gy0 = r->geo_box[1]; ra.gw = r->geo_box[2]; ra.gh = r->geo_box[3];
    ra.texmm = nullptr; ra.ctile = nullptr; ra.ctile_zero = nullptr; ra.ctile_nx = 0;
    ra.coltab = r->no_coltab ? nullptr : r->d_coltab;
    // params.shader_pre_smoothed: the module's stage-1 shader believes something else about its textures than what the
    // K5 pass did.  Only the raster launch sees that belief (as its smooth_pass); everything before it follows the real one.
    glava_b200_params pr_store;
    const glava_b200_params* pr = &p;
    if (p.shader_pre_smoothed) {
        pr_store = p; pr_store.smooth_pass = p.shader_pre_smoothed == 1 ? 1 : 0; pr = &pr_store;
        if (p.module == GLAVA_B200_MOD_CIRCLE) ra.geo = nullptr;          // the circle cache is laid out for the consistent case
    }
    if (p.module == GLAVA_B200_MOD_CIRCLE && ra.geo && !r->no_texmm) {
        if ((rc = launch_texmm(p, ra.tex, r->d_texmm, r->batch * 2, r->stream)) != 0) return rc;
        ++r->launches;
        ra.texmm = r->d_texmm;
        if (r->d_ctile) { ra.ctile = (const int4*) r->d_ctile; ra.ctile_zero = (const int*) ((const int4*) r->d_ctile + r->ctile_count); ra.ctile_nx = r->ctile_nx; }
    }
    if (r->timing && (rc = timing_mark(r->ev_ras, r->stream)) != 0) return rc;
    int raster_launches = 0;
    if ((rc = launch_raster(*pr, ra, r->stream, &raster_launches)) != 0) return rc;
    if (r->timing && (rc = timing_mark(r->ev_ras, r->stream)) != 0) return rc;
    CU(cudaEventRecord(r->ev_raster_done[b], r->stream));
    r->launches += (uint64_t) raster_launches;
    return 0;
}

static int update_host(glava_b200* r, const float* lb, const float* rb, size_t bsz, int modified, const uint8_t* mask);

int glava_b200_update(glava_b200* r, const float* lb, const float* rb, size_t bsz, int modified) {
    clear_error();
    if (!r || !lb) return fail(GLAVA_B200_EINVAL, "glava_b200_update: null argument");
    return update_host(r, lb, rb, bsz, modified, nullptr);
}

static int update_host(glava_b200* r, const float* lb, const float* rb, size_t bsz, int modified, const uint8_t* mask) {
    if (bsz != (size_t) r->n_in) return fail(GLAVA_B200_EINVAL, "glava_b200_update: bsz %zu != setbufsize %d", bsz, r->n_in);
    // rd_update always gets both rings (render.h:58); only `wave` ignores rb (wave/1.frag:7 samples audio_l alone)
    if (modified && !rb && r->p.module != GLAVA_B200_MOD_WAVE)
        return fail(GLAVA_B200_EINVAL, "glava_b200_update: module '%s' reads both channels, rb is null", module_name(r->p.module));
    CU(cudaSetDevice(r->device));
    size_t bytes = (size_t) r->batch * bsz * 4;
    const int b = r->stage_cur;
    if (modified) {
        // Staging buffer b is free once the spectrum kernel that last read it has finished; the copy runs
        // on its own stream so it overlaps the previous update's kernels.  The call returns after the
        // copy has completed, so — like rd_update — the caller may reuse lb / rb immediately.
        CU(cudaStreamWaitEvent(r->copy_stream, r->ev_free[b], 0));
        CU(cudaMemcpyAsync(r->d_pcm[b][0], lb, bytes, cudaMemcpyHostToDevice, r->copy_stream));
        if (r->p.module != GLAVA_B200_MOD_WAVE)
            CU(cudaMemcpyAsync(r->d_pcm[b][1], rb, bytes, cudaMemcpyHostToDevice, r->copy_stream));
        CU(cudaEventRecord(r->ev_copied[b], r->copy_stream));
        CU(cudaStreamWaitEvent(r->spec_stream, r->ev_copied[b], 0));
    }
    int rc = run_update(r, r->d_pcm[b][0], r->d_pcm[b][1], modified, false, mask);
    if (rc) return rc;
    if (modified) {
        CU(cudaEventRecord(r->ev_free[b], r->spec_stream));
        r->stage_cur = b ^ 1;
        r->last_in = b;
        if (!r->async_input) CU(cudaEventSynchronize(r->ev_copied[b]));
    }
    return 0;
}

// Double-buffer contract for the host rings (instead of rd_update's "consumed on return"): with async input enabled
// glava_b200_update* return as soon as the copy is ENQUEUED; the caller keeps lb / rb untouched until
// glava_b200_wait_input() — in practice it alternates between two sets of rings, and the host never blocks on a copy.
int glava_b200_set_async_input(glava_b200* r, int enable) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    r->async_input = enable != 0;
    return 0;
}
int glava_b200_wait_input(glava_b200* r) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    if (r->last_in < 0) return 0;
    CU(cudaSetDevice(r->device));
    CU(cudaEventSynchronize(r->ev_copied[r->last_in]));
    return 0;
}

int glava_b200_update_masked(glava_b200* r, const float* lb, const float* rb, size_t bsz, const uint8_t* modified) {
    clear_error();
    if (!r || !lb || !modified) return fail(GLAVA_B200_EINVAL, "glava_b200_update_masked: null argument");
    int any = 0;
    for (int s = 0; s < r->batch; ++s) any |= modified[s];
    return update_host(r, lb, rb, bsz, any ? 1 : 0, modified);
}

int glava_b200_update_device_masked(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, const uint8_t* modified) {
    clear_error();
    if (!r || !d_lb || !modified) return fail(GLAVA_B200_EINVAL, "glava_b200_update_device_masked: null argument");
    if (bsz != (size_t) r->n_in) return fail(GLAVA_B200_EINVAL, "glava_b200_update_device_masked: bsz %zu != setbufsize %d", bsz, r->n_in);
    if (((uintptr_t) d_lb & 15) || ((uintptr_t) d_rb & 15)) return fail(GLAVA_B200_EINVAL, "device PCM pointers must be 16-byte aligned");
    if (!d_rb && r->p.module != GLAVA_B200_MOD_WAVE) return fail(GLAVA_B200_EINVAL, "glava_b200_update_device_masked: module '%s' reads both channels, d_rb is null", module_name(r->p.module));
    CU(cudaSetDevice(r->device));
    return run_update(r, d_lb, d_rb ? d_rb : d_lb, 1, false, modified);
}

int glava_b200_update_rings_masked(glava_b200* r, const uint8_t* modified) {
    clear_error();
    if (!r || !modified) return fail(GLAVA_B200_EINVAL, "glava_b200_update_rings_masked: null argument");
    CU(cudaSetDevice(r->device));
    return run_update(r, r->d_ring[r->ring_cur][0], r->d_ring[r->ring_cur][1], 1, false, modified);
}

// Ordering of glava_b200_update_device against the caller's own CUDA work.  The spectrum kernel reads d_lb / d_rb on an
// internal non-blocking stream: hand over the event that marks "buffers written" (recorded on the producer's stream) and the
// kernel waits for it on the device; glava_b200_input_event() is recorded after the kernel that last READ the buffers, so
// the producer can cudaStreamWaitEvent on it before overwriting them.  Both are cudaEvent_t passed as void*.
int glava_b200_update_device_after(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, int modified, void* ready_event) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    if (ready_event) CU(cudaStreamWaitEvent(r->spec_stream, (cudaEvent_t) ready_event, 0));
    return glava_b200_update_device(r, d_lb, d_rb, bsz, modified);
}
void* glava_b200_input_event(glava_b200* r) {
    if (!r) return nullptr;
    if (!r->ev_input) { if (cudaEventCreateWithFlags(&r->ev_input, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; } }
    cudaSetDevice(r->device);
    cudaEventRecord(r->ev_input, r->spec_stream);          // after everything enqueued so far on the reading stream
    return (void*) r->ev_input;
}

int glava_b200_update_device(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, int modified) {
    clear_error();
    if (!r || !d_lb) return fail(GLAVA_B200_EINVAL, "glava_b200_update_device: null argument");
    if (bsz != (size_t) r->n_in) return fail(GLAVA_B200_EINVAL, "glava_b200_update_device: bsz %zu != setbufsize %d", bsz, r->n_in);
    if (((uintptr_t) d_lb & 15) || ((uintptr_t) d_rb & 15)) return fail(GLAVA_B200_EINVAL, "device PCM pointers must be 16-byte aligned");
    if (modified && !d_rb && r->p.module != GLAVA_B200_MOD_WAVE)
        return fail(GLAVA_B200_EINVAL, "glava_b200_update_device: module '%s' reads both channels, d_rb is null", module_name(r->p.module));
    CU(cudaSetDevice(r->device));
    return run_update(r, d_lb, d_rb ? d_rb : d_lb, modified);
}

static int ingest(glava_b200* r, const void* chunks, bool float_in, int frames, const char* who) {
    clear_error();
    if (!r || !chunks) return fail(GLAVA_B200_EINVAL, "%s: null argument", who);
    if (frames < 1 || frames > r->n_in) return fail(GLAVA_B200_EINVAL, "%s: frames %d out of range", who, frames);
    CU(cudaSetDevice(r->device));
    size_t bytes = (size_t) r->batch * frames * 2 * (float_in ? sizeof(float) : sizeof(int16_t));
    if (bytes > r->chunks_cap) {
        CU(cudaStreamSynchronize(r->spec_stream));
        if (r->d_chunks) cudaFree(r->d_chunks);
        r->d_chunks = nullptr; r->chunks_cap = 0;
        CU(cudaMalloc((void**) &r->d_chunks, bytes));
        r->chunks_cap = bytes;
    }
    // ordered with the spectrum kernels (they read the rings): same stream
    CU(cudaMemcpyAsync(r->d_chunks, chunks, bytes, cudaMemcpyHostToDevice, r->spec_stream));
    int cur = r->ring_cur, nxt = cur ^ 1;
    int rc = launch_fifo_ingest(r->p_user, r->d_chunks, float_in, frames, r->d_ring[cur][0], r->d_ring[cur][1],
                                r->d_ring[nxt][0], r->d_ring[nxt][1], r->batch, r->spec_stream);
    if (rc) return rc;
    ++r->launches;
    r->ring_cur = nxt;
    return 0;
}
int glava_b200_ingest_fifo(glava_b200* r, const int16_t* chunks, int frames) { return ingest(r, chunks, false, frames, "glava_b200_ingest_fifo"); }
int glava_b200_ingest_float(glava_b200* r, const float* chunks, int frames) { return ingest(r, chunks, true, frames, "glava_b200_ingest_float"); }

int glava_b200_update_rings(glava_b200* r, int modified) {
    clear_error();
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    return run_update(r, r->d_ring[r->ring_cur][0], r->d_ring[r->ring_cur][1], modified);
}

// ---- offscreen hand-off (glava.h:16-25, glava.c:244-267) -----------------------------------------------------------
int glava_b200_sizereq(glava_b200* r, int w, int h) {
    if (!r || w < 1 || h < 1 || w > 0x7fffffff) return fail(GLAVA_B200_EINVAL, "glava_b200_sizereq: bad arguments");
    r->sizereq.store((1ull << 63) | ((unsigned long long) (unsigned) w << 32) | (unsigned) h);
    return 0;
}
int glava_b200_wait_frame(glava_b200* r) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    CU(cudaEventSynchronize(r->ev_raster_done[r->tex_cur]));
    return 0;
}
void* glava_b200_frame_event(const glava_b200* r) { return r ? (void*) r->ev_raster_done[r->tex_cur] : nullptr; }
const void* glava_b200_frame_device(const glava_b200* r, int stream) {
    if (!r || stream < 0 || stream >= r->batch || !r->d_fb) return nullptr;
    return r->d_fb + (size_t) r->p.w * r->p.h * 4 * (size_t) (stream % r->slots);
}
int glava_b200_framebuffer_ipc(glava_b200* r, void* handle, size_t handle_bytes) {
    clear_error();
    if (!r || !handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return fail(GLAVA_B200_EINVAL, "glava_b200_framebuffer_ipc: need a %zu-byte buffer", sizeof(cudaIpcMemHandle_t));
    CU(cudaSetDevice(r->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, r->d_fb));
    memcpy(handle, &h, sizeof(h));
    return 0;
}

int glava_b200_set_timing(glava_b200* r, int enable) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    int rc0 = sync_all(r); if (rc0) return rc0;
    for (cudaEvent_t e : r->ev_spec) cudaEventDestroy(e);
    for (cudaEvent_t e : r->ev_ras) cudaEventDestroy(e);
    r->ev_spec.clear(); r->ev_ras.clear();
    r->timing = enable != 0;
    return 0;
}

int glava_b200_kernel_times(glava_b200* r, double* spectrum_ms, int* spectrum_launches, double* raster_ms, int* raster_launches) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    int rc0 = sync_all(r); if (rc0) return rc0;
    double s = 0, q = 0; int ns = 0, nq = 0;
    for (size_t i = 0; i + 1 < r->ev_spec.size(); i += 2) {
        float a = 0; CU(cudaEventElapsedTime(&a, r->ev_spec[i], r->ev_spec[i + 1])); s += a; ++ns;
    }
    for (size_t i = 0; i + 1 < r->ev_ras.size(); i += 2) {
        float a = 0; CU(cudaEventElapsedTime(&a, r->ev_ras[i], r->ev_ras[i + 1])); q += a; ++nq;
    }
    if (spectrum_ms) *spectrum_ms = s;
    if (spectrum_launches) *spectrum_launches = ns;
    if (raster_ms) *raster_ms = q;
    if (raster_launches) *raster_launches = nq;
    return 0;
}

// Event timeline of the timed launches, in ms relative to the first spectrum mark: out = [start, end] pairs, spectrum
// launches first (n_spec pairs) then raster launches.  Development aid for looking at the overlap of the two streams.
int glava_b200_timeline(glava_b200* r, double* out, int cap_pairs, int* n_spec, int* n_ras) {
    if (!r || !out) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    int rc0 = sync_all(r); if (rc0) return rc0;
    const int ns = (int) r->ev_spec.size() / 2, nq = (int) r->ev_ras.size() / 2;
    if (ns + nq > cap_pairs || ns == 0) return fail(GLAVA_B200_EINVAL, "glava_b200_timeline: %d pairs, capacity %d", ns + nq, cap_pairs);
    cudaEvent_t t0 = r->ev_spec[0];
    for (int i = 0; i < 2 * ns; ++i) { float a = 0; CU(cudaEventElapsedTime(&a, t0, r->ev_spec[i])); out[i] = a; }
    for (int i = 0; i < 2 * nq; ++i) { float a = 0; CU(cudaEventElapsedTime(&a, t0, r->ev_ras[i])); out[2 * ns + i] = a; }
    if (n_spec) *n_spec = ns;
    if (n_ras) *n_ras = nq;
    return 0;
}

int glava_b200_sync(glava_b200* r) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    return sync_all(r);
}

int glava_b200_readback(glava_b200* r, int stream, uint8_t* rgba) {
    clear_error();
    if (!r || !rgba || stream < 0 || stream >= r->batch) return fail(GLAVA_B200_EINVAL, "glava_b200_readback: bad arguments");
    CU(cudaSetDevice(r->device));
    size_t frame = (size_t) r->p.w * r->p.h * 4;
    CU(cudaMemcpyAsync(rgba, r->d_fb + frame * (size_t) (stream % r->slots), frame, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}

int glava_b200_readback_async(glava_b200* r, int stream, uint8_t* rgba) {
    clear_error();
    if (!r || !rgba || stream < 0 || stream >= r->batch) return fail(GLAVA_B200_EINVAL, "glava_b200_readback_async: bad arguments");
    CU(cudaSetDevice(r->device));
    const size_t frame = (size_t) r->p.w * r->p.h * 4;
    if (!r->out_stream) {
        CU(cudaStreamCreateWithFlags(&r->out_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&r->ev_stage_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&r->ev_stage_free[i], cudaEventDisableTiming));
        }
    }
    if (r->stage_bytes != frame) {                       // first use, or the geometry changed (glava_b200_sizereq)
        CU(cudaStreamSynchronize(r->out_stream));
        for (int i = 0; i < 2; ++i) { if (r->d_stage[i]) cudaFree(r->d_stage[i]); r->d_stage[i] = nullptr; }
        r->stage_bytes = 0;
        for (int i = 0; i < 2; ++i) CU(cudaMalloc((void**) &r->d_stage[i], frame));
        r->stage_bytes = frame;
    }
    const int i = r->out_cur;
    // the frame is snapshotted on the raster stream (ordered after the raster that produced it, before the next one
    // overwrites the slot); the slow PCIe leg then runs from the snapshot on its own stream
    CU(cudaStreamWaitEvent(r->stream, r->ev_stage_free[i], 0));
    CU(cudaMemcpyAsync(r->d_stage[i], r->d_fb + frame * (size_t) (stream % r->slots), frame, cudaMemcpyDeviceToDevice, r->stream));
    CU(cudaEventRecord(r->ev_stage_ready[i], r->stream));
    CU(cudaStreamWaitEvent(r->out_stream, r->ev_stage_ready[i], 0));
    CU(cudaMemcpyAsync(rgba, r->d_stage[i], frame, cudaMemcpyDeviceToHost, r->out_stream));
    CU(cudaEventRecord(r->ev_stage_free[i], r->out_stream));
    r->out_cur = i ^ 1;
    return 0;
}

// Orders everything enqueued on the handle's stream from now on after the read-backs issued so far (their PCIe leg runs
// on a separate stream): an event recorded on glava_b200_cuda_stream() after this call covers them.
int glava_b200_readback_fence(glava_b200* r) {
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    if (!r->out_stream) return 0;
    CU(cudaSetDevice(r->device));
    for (int i = 0; i < 2; ++i) CU(cudaStreamWaitEvent(r->stream, r->ev_stage_free[i], 0));
    return 0;
}

static int planes_to_host(glava_b200* r, const void* d, size_t elem, void* out_l, void* out_r) {
    // device layout [batch][2][n] -> two host arrays [batch][n]
    const size_t row = (size_t) r->p.n * elem;
    CU(cudaStreamSynchronize(r->spec_stream));
    if (out_l) CU(cudaMemcpy2DAsync(out_l, row, d, 2 * row, row, (size_t) r->batch, cudaMemcpyDeviceToHost, r->stream));
    if (out_r) CU(cudaMemcpy2DAsync(out_r, row, (const char*) d + row, 2 * row, row, (size_t) r->batch, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    return 0;
}
int glava_b200_spectrum(glava_b200* r, float* out_l, float* out_r) {
    clear_error();
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    return planes_to_host(r, r->d_spec_cur, 4, out_l, out_r);
}
int glava_b200_textures(glava_b200* r, uint16_t* out_l, uint16_t* out_r) {
    clear_error();
    if (!r) return fail(GLAVA_B200_EINVAL, "null argument");
    CU(cudaSetDevice(r->device));
    return planes_to_host(r, tex_half(r, r->tex_cur), 2, out_l, out_r);
}

int glava_b200_smooth_pass(glava_b200* r, const uint16_t* in, uint16_t* out, int count) {
    clear_error();
    if (!r || !in || !out || count < 1) return fail(GLAVA_B200_EINVAL, "glava_b200_smooth_pass: bad arguments");
    CU(cudaSetDevice(r->device));
    size_t bytes = (size_t) count * r->p.n * 2;
    uint16_t* d_in = nullptr; uint16_t* d_out = nullptr;
    CU(cudaMalloc((void**) &d_in, bytes));
    cudaError_t e = cudaMalloc((void**) &d_out, bytes);
    if (e != cudaSuccess) { cudaFree(d_in); return fail(GLAVA_B200_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, r->stream) != cudaSuccess) { rc = fail(GLAVA_B200_ECUDA, "H2D copy failed"); break; }
        if ((rc = launch_smooth_only(r->p, d_in, d_out, count, r->stream, &r->k5)) != 0) break;
        ++r->launches;
        if (cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, r->stream) != cudaSuccess) { rc = fail(GLAVA_B200_ECUDA, "D2H copy failed"); break; }
        cudaError_t se = cudaStreamSynchronize(r->stream);
        if (se != cudaSuccess) rc = fail(GLAVA_B200_ECUDA, "smooth pass: %s", cudaGetErrorString(se));
    } while (0);
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}

int glava_b200_raster_textures(glava_b200* r, const uint16_t* tex_l, const uint16_t* tex_r) {
    clear_error();
    if (!r || !tex_l) return fail(GLAVA_B200_EINVAL, "glava_b200_raster_textures: null argument");
    CU(cudaSetDevice(r->device));
    const size_t row = (size_t) r->p.n * 2;
    CU(cudaStreamSynchronize(r->spec_stream));
    char* dst = (char*) tex_half(r, r->tex_cur);
    CU(cudaMemcpy2DAsync(dst, 2 * row, tex_l, row, row, (size_t) r->batch, cudaMemcpyHostToDevice, r->stream));
    if (tex_r) CU(cudaMemcpy2DAsync(dst + row, 2 * row, tex_r, row, row, (size_t) r->batch, cudaMemcpyHostToDevice, r->stream));
    return run_update(r, nullptr, nullptr, 0, true);
}

int glava_b200_transform_smooth(glava_b200* r, float* planes, int count) {
    clear_error();
    if (!r || !planes || count < 1) return fail(GLAVA_B200_EINVAL, "glava_b200_transform_smooth: bad arguments");
    if (!r->p.transform_smooth) return fail(GLAVA_B200_EINVAL, "glava_b200_transform_smooth: the renderer was not created with transform_smooth");
    CU(cudaSetDevice(r->device));
    const size_t bytes = (size_t) count * r->p.n * 4;
    float* d = nullptr;
    CU(cudaMalloc((void**) &d, bytes));
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d, planes, bytes, cudaMemcpyHostToDevice, r->stream) != cudaSuccess) { rc = fail(GLAVA_B200_ECUDA, "H2D copy failed"); break; }
        if ((rc = launch_transform_smooth(d, r->p.n, r->d_ts_tab, r->ts_asz, r->ts_lim, count, r->stream)) != 0) break;
        ++r->launches;
        if (cudaMemcpyAsync(planes, d, bytes, cudaMemcpyDeviceToHost, r->stream) != cudaSuccess) { rc = fail(GLAVA_B200_ECUDA, "D2H copy failed"); break; }
        cudaError_t se = cudaStreamSynchronize(r->stream);
        if (se != cudaSuccess) rc = fail(GLAVA_B200_ECUDA, "transform_smooth: %s", cudaGetErrorString(se));
    } while (0);
    cudaFree(d);
    return rc;
}

}  // extern "C"
