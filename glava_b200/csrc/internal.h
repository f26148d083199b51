// internal.h — shared declarations of libglava_b200 (host side + kernel launchers).
#ifndef GLAVA_B200_INTERNAL_H
#define GLAVA_B200_INTERNAL_H

#include "../../include/glava_b200.h"

#include <cstddef>
#include <cstdint>

namespace glb {

// error reporting: stores the message, calls the abort hook (glava.h:17 analogue), returns `code`
int  fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();
bool has_error();

// config.cpp
void fill_defaults(glava_b200_params* p, int module);
int  module_from_name(const char* name);
const char* module_name(int id);
bool parse_hex_color(const char* s, float out[4], bool literal_rounding);
bool parse_hex_components(const char* s, float* out[4]);
int  load_config(glava_b200_params* out, const char* const* paths, const char* entry,
                 const char* const* requests, const char* force_module, const char* const* binds);
int  validate_params(const glava_b200_params* p);

#define GLB_MAX_AVG_FRAMES 16

#ifndef GLB_TAPENTRY_DEFINED
#define GLB_TAPENTRY_DEFINED
struct alignas(8) TapEntry { int idx; float w; };   // one tap of the K5 smoothing sum: texel index, weight
#endif

// ---- views handed to the kernels (all pointers are DEVICE pointers) ---------------------------
// "channel plane" c = stream * 2 + ch (ch 0 = left, 1 = right); every per-channel array is
// [batch*2][...] so one CTA of the spectrum kernel owns one contiguous plane.
struct SpectrumArgs {
    const float*  pcm_l;        // [batch][n]   ring contents, oldest first (glava.c:528-537 lb)
    const float*  pcm_r;        // [batch][n]
    const double* window;       // [n]          render.c:660,794 window LUT (double, as the macro evaluates)
    const float*  twiddle;      // [n/2][2]     exp(-2*pi*i*k/(n/2)) as (re, im)
    float*    spec;             // [batch*2][n] pipeline-A result, or raw transform_fft output (accel)
    float*    applied;          // [batch*2][n] transform_gravity state (A), render.c:725
    float*    ring_f;           // [batch*2][F][n] transform_average ring (A), slot = update index % F
    uint16_t* gr_store;         // [batch*2][n] (B) render.c:2197
    uint16_t* ring_u;           // [batch*2][F][n] (B) gr->out[], render.c:2232
    uint16_t* tex;              // [batch*2][n] R16 texture the module samples
    uint16_t* av_out;           // non-null: export the pre-smoothing texture here and skip K5 (k5_planes_kernel follows)
    const int* need;            // lazy K5: texel indices to evaluate, [2][need_count] (-1 = unused), nullptr = all n
    int       need_count;
    // precomputed K5 taps for the need-list (weights and indices do not depend on the audio):
    const TapEntry* tap_tab;    // [2][tap_max][need_count], tap-major so a warp's loads coalesce; nullptr = evaluate directly
    const int*   tap_cnt;       // [2][need_count]
    const float* tap_wsum;      // [2][need_count]  sum of the weights in loop order
    int       tap_max;
    int       tap_ku;           // table entries a thread keeps in flight (8, 4 or 2)
    int       epi_n;            // lazy K5: number of leading bins whose gravity/average state can reach a sampled texel (0 = all)
    // the same taps as one blob per channel, small enough to live in shared memory (loaded by the TMA engine while the
    // FFT runs): [float w[csr_total]] [uint16 idx[csr_total]] [int off[need_count + 1]], texel-major ("CSR"); nullptr = unused
    const unsigned char* csr;   // [2][csr_bytes]
    int       csr_bytes;        // bytes per channel (multiple of 16)
    int       csr_idx_off, csr_off_off;   // byte offsets of idx[] and off[] inside a blob
    int       skip_tex;         // 1: produce `spec` only (transform_smooth / keyframe lerp / upload / K5 follow as kernels)
    int       batch;
    unsigned long long update;  // number of modified updates before this one (ring cursor)
    // per-stream `modified` (glava.c:528-537 decides per renderer, i.e. per stream): nullptr = every stream is modified and
    // shares the cursor `update % F`; else one word per stream, bit 31 = this stream has new audio, low 16 bits = its own
    // ring cursor (modified updates of THAT stream so far, mod F).  An unmodified stream's state is left alone and its
    // texture is carried from `tex_prev` (the half of the double buffer the previous raster read) into `tex`.
    int variant_oop, variant_t; // kernel variant: out-of-place passes, threads per CTA (0 = the size's default)
    int fft_only;               // 1: stop after transform_fft — `spec` (the leading epi_n bins) is all this launch produces; gravity /
                                // average run as epilogue_b_kernel, K5 as k5_need_smem_kernel / k5_need_kernel (capi.cu run_update)
    int av_t_len;               // > 0: av_out receives only the leading av_t_len bins (need-list K5 as its own kernel downstream)
    const uint32_t* umask;      // [batch]
    const uint16_t* tex_prev;   // [batch*2][n]
    double    avg_w_a[GLB_MAX_AVG_FRAMES];   // pipeline A weights, oldest first (render.c:661,766)
    float     avg_w_b[GLB_MAX_AVG_FRAMES];   // pipeline B weights, newest first (average_pass.frag:41)
    int       avg_b_windowed;                // average_pass.frag:27-29,38-42
};

struct RasterArgs {
    const uint16_t* tex;        // [batch*2][n]
    uint8_t*  fb;               // [slots][h][w][4]
    const void* rowtab;         // module row table (bars: [h] {fill, outline} RGBA8 pairs), may be null
    int       batch, slots, stream0;
    // polar modules: per-renderer cache of the audio-independent per-pixel geometry (RadialGeo /
    // CircleGeo, 16 bytes each) over the bounding box of the disc that can be non-zero; may be null
    const void* geo;
    int       gx0, gy0, gw, gh;  // box origin and size in pixels (gx0, gw multiples of 4)
    const uint32_t* texmm;       // circle: per plane of `tex` {min, max} followed by GLB_CIRCLE_NB bucket {min, max} pairs (may be null)
    // circle: per 128 x 8 tile (+ 1-pixel halo) the range of texel indices its cells reference, {lo_l, hi_l, lo_r, hi_r}
    // (lo > hi: none), and whether any reference is out of range (reads 0); audio-independent, built with the cache
    const int4* ctile; const int* ctile_zero; int ctile_nx;
    // graph / wave: per-stream column table [batch][GLB_COLTAB_PLANES][w] floats, scratch that launch_raster fills on the
    // raster stream right before the module kernel (graph: plane 0 = column height; wave: WaveCol's five fields); may be null
    float* coltab;
};
#define GLB_COLTAB_PLANES 5
#define GLB_CIRCLE_NB 128                         // texel buckets per plane (n / 128 texels each)
#define GLB_TEXMM_STRIDE (2 + 2 * GLB_CIRCLE_NB)  // u32 per plane

// ---- kernel launchers (spectrum_kernels.cu, raster_kernels.cu); `stream` is a cudaStream_t passed as void* ------------------
int launch_spectrum(const glava_b200_params& p, const SpectrumArgs& a, bool is_fft, void* stream);
// Full-plane K5 tap table (built once per parameter set, capi.cu): for every block of K5_BLOCK output texels the taps of
// all its texels, tap-major and padded to the block's longest sum with zero-weight entries, so a warp's loads coalesce
// and the tap loop is uniform.
struct K5Table {
    const int4*  blk;      // [n / K5_BLOCK] {first entry, taps per texel (padded), first input index, input span}
    const int2*  ent;      // entries {input index - first input index, weight bits}
    const float* wsum;     // [count] sum of an output's weights in loop order
    int smem_bytes;        // dynamic shared memory the kernel needs: planes per CTA * max span * sizeof(float)
    const int* out;        // [count] texel each output position writes (nullptr: position = texel, count = n)
    int count;             // outputs (need-list tables: the texels the module samples; full table: n)
    int max_span;
};
#ifndef K5_BLOCK
#define K5_BLOCK 128
#endif
#define K5_S_PLANES 8        // planes that share one tap (K5_S in spectrum_kernels.cu)
// planes handled: plane(i) = i * plane_stride + plane_offset for i in [0, count) — stride 2 / offset ch walks one channel's
// planes of the interleaved [batch][2] layout (need-list tables differ per channel)
int launch_smooth_only(const glava_b200_params& p, const uint16_t* d_in, uint16_t* d_out, int count, void* stream,
                       const K5Table* table = nullptr, int plane_stride = 1, int plane_offset = 0);
// need-list K5, lanes = streams (d_av [batch*2][n] u16 as exported by the spectrum kernel -> transposed float d_av_t): csr = the texel-major tap blobs of SpectrumArgs (per channel float w[] | u16 idx[] | int off[]),
// need / wsum = [2][need_count]; av_t as exported by the spectrum kernel; writes tex[(stream * 2 + ch) * n + need[k]]
int launch_k5_need(const glava_b200_params& p, const uint16_t* d_av, float* d_av_t, int av_t_len, uint16_t* d_tex, int batch, int channels,
                   const unsigned char* d_csr, int csr_bytes, int csr_idx_off, int csr_off_off, const int* d_need,
                   const float* d_wsum, int need_count, void* stream);
// the same need-list K5 from shared-memory tiles over blocks of sampled texels (tables.h build_need_blocks); reads the
// [plane][bin] R16 texture directly
int launch_k5_need_smem(const glava_b200_params& p, const uint16_t* d_av, uint16_t* d_tex, int batch, int channels,
                        const unsigned char* d_csr, int csr_bytes, int csr_idx_off, int csr_off_off, const int* d_need,
                        const float* d_wsum, int need_count, const void* d_blocks, int nblk, int max_rows, void* stream);
// pipeline B epilogue as its own elementwise kernel: spec (transform_fft output) -> upload quantisation, K1 - K4 on the R16
// state, pre-smoothing texels into av_out (leading `bins` of every plane; bins is a multiple of 8)
int launch_epilogue_b(const glava_b200_params& p, const SpectrumArgs& a, int bins, void* stream);
int launch_raster(const glava_b200_params& p, const RasterArgs& a, void* stream, int* launched = nullptr);   // *launched += kernels launched
int launch_bars_rowtab(const glava_b200_params& p, void* d_rowtab, void* stream);
// geometry cache of the polar modules: box = {x0, y0, w, h}; returns bytes needed when d_geo == nullptr
size_t polar_geo_box(const glava_b200_params& p, int box[4]);
int launch_texmm(const glava_b200_params& p, const uint16_t* d_tex, uint32_t* d_out, int planes, void* stream);
int launch_polar_geo(const glava_b200_params& p, void* d_geo, const int box[4], void* stream);
// circle: the per-tile texel reference ranges from the finished cache; d_tiles = int4[ntx * nty] followed by int[ntx * nty]
size_t circle_tile_bytes(const glava_b200_params& p, int* ntx, int* nty);
int launch_circle_tiles(const glava_b200_params& p, const void* d_geo, const int box[4], void* d_tiles, void* stream);
int launch_fifo_ingest(const glava_b200_params& p, const void* d_chunks, bool float_in, int frames, const float* src_l, const float* src_r,
                       float* dst_l, float* dst_r, int batch, void* stream);
int spectrum_smem_bytes(int n);
int spectrum_threads(int n);
// chain_kernels.cu: optional stages of rd_update (bufscale, transform_smooth, keyframe lerp + R16 upload)
int launch_bufscale(const float* in_l, const float* in_r, float* out_l, float* out_r, int batch, int n_in, int k,
                    int channels, void* stream);
int launch_transform_smooth(float* d_planes, int n, const void* d_tab, int asz, int lim, int count, void* stream,
                            const uint32_t* d_umask = nullptr, int mask_shift = 1);   // d_umask: SpectrumArgs::umask, planes of unmodified
                            // streams are skipped; stream of plane i = i >> mask_shift (1: interleaved [batch*2] planes, 0: one channel's [batch])
int launch_upload(const float* d_s, const float* d_e, float mod, uint16_t* d_out, size_t total, void* stream);

}  // namespace glb

#endif
