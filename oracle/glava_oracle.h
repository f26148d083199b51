/* TEST INFRASTRUCTURE — CPU restatement of the reference's PCM->spectrum->pixels path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libglava_b200.so) never links or calls it.
 *
 * Parity status.  Pinned against the reference's OWN code, compiled where it lies by oracle/Makefile into oracle/_ref/:
 *   - the transforms (render.c transform_fft / gravity / average / wrange / smooth): bit for bit, live and through the golden
 *     vectors generated from them (tests/golden/);
 *   - the FIFO ring update (fifo.c's audio thread on real named pipes);
 *   - rd_update's orchestration — transform chain, `modified`, setbufscale, keyframe interpolation, the "smooth" transform,
 *     pipeline B's upload — and rd_new's reading of configurations: both run for real on a null OpenGL driver
 *     (libglava_ref_rd.so), the whole program included (glava_entry with fifo.c on a named pipe).
 *   - the RASTER half and the GL passes K1 - K5 (round 2): the reference's own renderer — rd_new / rd_update, its pass /
 *     gravity_pass / average_pass / smooth_pass.frag and the module shaders — run on a REAL OpenGL: Mesa 18.1.9 llvmpipe
 *     (GLava's stated software floor, README.md:121), found inside the Nsight Compute bundle of this image and of the GPU
 *     box and driven without an X server (oracle/ref_shim.c's "mesa" window backend, oracle/fakex/, oracle/ref_gl.py).
 *     tests/golden/llvmpipe_golden.npz holds what it rendered for 58 configurations (every upload, every 1-D pass texture,
 *     the final frames; BASELINE geometries, every module option, 30 random configurations); tests/test_llvmpipe_golden.py
 *     replays them through this restatement pass by pass, re-generates them live wherever the harness loads, and — with
 *     -m gpu — compares the kernels with them.  Result: uploads and every blended frame bit for bit; K1 - K4 bit for bit
 *     except <= 2 texels per case by 1 LSB16; K5 <= 1 LSB16 (4 for ROUND_FORMULA circular: that llvmpipe's sqrt); native
 *     frames <= 1 LSB with <= 8 counted hard-edge pixels per case — the places where GLSL leaves sin / cos / log / atan /
 *     sqrt to the implementation.  Where this file's round-1 conventions differed from llvmpipe (unorm store rounding,
 *     blend arithmetic) they were CHANGED to llvmpipe's.
 *   - oracle/glsl_interp.py (round 1's stand-in for a GL: it EXECUTES the reference's shader sources in float32; its input
 *     is token-identical to the texts rd_new hands to glShaderSource) is kept for exact single-stage textures and the
 *     1000-seed randomised config differential (tests/golden/glsl_golden.npz, 36 module configurations; the `test`
 *     module's #55000055 known answer, shaders/glava/test_rc.glsl:27); it is itself pinned to the llvmpipe goldens now.
 * Not pinned by anything the reference can run: setbufsize 16384 on the GL side (GL_MAX_TEXTURE_SIZE of that Mesa is 8192 and
 * bind_1d_fbo aborts) — checked against this restatement, which is pinned at every smaller size, and against float64.
 */
#ifndef GLAVA_ORACLE_H
#define GLAVA_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MOD_BARS = 0, ORC_MOD_RADIAL = 1, ORC_MOD_CIRCLE = 2, ORC_MOD_GRAPH = 3,
       ORC_MOD_WAVE = 4, ORC_MOD_TEST = 5 };

/* colour that is either a constant or mix(lo, hi, clamp(X / gradient, 0, 1)) */
typedef struct {
    int   mode;          /* 0 = gradient mix, 1 = constant (lo) */
    float lo[4], hi[4];
    float gradient;
} orc_color;

typedef struct {
    /* ---- spectrum (render.c transforms, util/ *_pass.frag) ---- */
    int   n;               /* setbufsize: floats per channel                          */
    float fft_scale, fft_cutoff, gravity_step, ur;
    int   avg_frames, avg_window;
    int   accel_fft;       /* 0: pipeline A (CPU float chain), 1: pipeline B (R16 GL passes) */
    int   smooth_pass;     /* K5 pre-smoothing pass on/off                            */
    float smooth_factor, sample_range, sample_scale, hybrid_weight;
    int   sample_mode;     /* 0 average, 1 maximum, 2 hybrid                          */
    int   round_formula;   /* 0 sinusoidal, 1 linear, 2 circular                      */
    /* ---- raster, common ---- */
    int   module, w, h, channels, premultiply_alpha;
    /* ---- bars ---- */
    float bars_width, bars_gap, bars_outline_width, bars_amplify;
    orc_color bars_color;
    int   bars_outline_mode;   /* 0: vec4(COLOR.rgb*1.5, COLOR.a), 1: constant */
    float bars_outline[4];
    int   bars_direction, bars_invert, bars_flip, bars_mirror_yx;
    /* ---- radial ---- */
    float radial_radius, radial_line;
    float radial_line_half;   /* value of `(C_LINE / 2)` as GLSL evaluates it (int division for an int literal) */
    float radial_outline[4];
    int   radial_nbars; float radial_bar_width, radial_amplify;
    orc_color radial_color;
    float radial_rotate; int radial_invert;
    float radial_bar_alias, radial_c_alias, radial_off_x, radial_off_y;
    /* ---- circle ---- */
    float circle_radius, circle_line, circle_outline[4], circle_amplify, circle_rotate;
    int   circle_invert, circle_fill, circle_smooth;
    /* ---- graph ---- */
    float graph_vscale; int graph_direction; orc_color graph_color;
    int   graph_draw_outline, graph_draw_highlight; float graph_outline[4]; int graph_invert;
    /* ---- wave ---- */
    float wave_min_thickness, wave_max_thickness, wave_base_color[4], wave_amplify, wave_outline[4];
    int   graph_join_channels;       /* JOIN_CHANNELS (graph.glsl:23) */
    int   graph_anti_alias;          /* ANTI_ALIAS (graph.glsl:19): graph/3.frag */
    int   shader_pre_smoothed;       /* the stage-1 header's _PRE_SMOOTHED_AUDIO when it contradicts smooth_pass: 1 = "smoothed", 2 = "raw" (0: consistent) */
    int   radial_bar_width_int;      /* BAR_WIDTH written as an integer literal: `BAR_WIDTH / 2` divides in integers */
    float radial_bar_outline_width, radial_bar_outline[4];   /* BAR_OUTLINE_WIDTH, BAR_OUTLINE (deprecated, radial.glsl:33-36) */
    float clear_color[4];            /* setbg / setbgf (render.c:1062-1099); only visible when premultiply_alpha == 0 */
} orc_params;

void orc_default_params(orc_params* p, int module, int n, int w, int h);

/* window LUT w[i] = 0.53836 - 0.46164*cos(2*pi*i/N - 1) as the macro at render.c:660
 * expands at render.c:794 (double), i in [0, n). */
void orc_window(double* w, int n);

/* transform_fft (render.c:783-847) restated in float32 with the same float twiddle
 * recurrence: intended to be bit-identical to the compiled reference. */
void orc_fft_f32(const orc_params* p, float* buf);
/* same maths, float64 arithmetic and exact twiddles ("truth" companion). in: float PCM,
 * out: double spectrum. */
void orc_fft_f64(const orc_params* p, const float* in, double* out);

/* per (stream, channel) persistent state */
typedef struct orc_chan orc_chan;
orc_chan* orc_chan_new(const orc_params* p);
void      orc_chan_free(orc_chan* c);

/* One audio update for one channel.
 *  pcm      : n floats (ring contents, oldest first), not modified
 *  spec_f32 : n floats out — pipeline A result (after fft+gravity+average) when
 *             accel_fft == 0; the raw transform_fft output when accel_fft == 1
 *  tex_u16  : n texels out — the R16 1-D texture the module shader samples
 *             (after upload quantisation, K1-K4 when accel, and K5 when smooth_pass)
 *  is_fft   : 1 for fft modules; 0 for `wave` (transform chain window+wrange, wave/1.frag:7-10);
 *             2 = `pcm` already holds transform_fft's output (computed by oracle/_ref)
 */
void orc_chan_update(orc_chan* c, const orc_params* p, const float* pcm, int is_fft,
                     float* spec_f32, uint16_t* tex_u16);

/* ---- optional stages of rd_update that the shipped configuration leaves off ---------------- */
/* bufscale (render.c:1765-1790): box-average `k` consecutive samples, out has n_in / k floats */
void orc_bufscale(const float* in, int n_in, int k, float* out);
/* transform_smooth (render.c:694-718): in place; b[0] becomes NaN (0/0), as in the reference */
void orc_transform_smooth(float* b, int sz, float smooth_distance, float smooth_ratio);
/* keyframe interpolation (render.c:1792-1809): out = s + (e - s) * min(ur / fr * kcounter, 1) */
void orc_interp(const float* s, const float* e, int n, float ur, float fr, int kcounter, float* out);

typedef struct {
    int   bufscale;          /* setbufscale */
    int   interpolate;       /* setinterpolate */
    float fr;                /* frame rate (rd_update calls per second); <= 0: same as ur */
    int   transform_smooth;  /* "smooth" appended to the module's transform chain */
    float smooth_distance, smooth_ratio;   /* setsmooth, setsmoothratio (render.c:917-918) */
} orc_ext;

/* One stream (both channels) through a whole rd_update, including the optional stages:
 * bufscale -> transform chain (-> transform_smooth) -> keyframe lerp -> R16 upload -> K1-K5.
 * p->n is setbufsize (what the caller passes as bsz); textures / spectra have p->n / bufscale entries.
 * Call order and state follow render.c:1761-1809, 2113-2309, 2347-2353, 2380-2383. */
typedef struct orc_stream orc_stream;
orc_stream* orc_stream_new(const orc_params* p, const orc_ext* x);
void        orc_stream_free(orc_stream* s);
int         orc_stream_n(const orc_stream* s);
void orc_stream_update(orc_stream* s, const float* lb, const float* rb, int modified,
                       float* spec_l, float* spec_r, uint16_t* tex_l, uint16_t* tex_r);

/* K5 alone: smooth_pass.frag over an R16 texture */
void orc_smooth_pass(const orc_params* p, const uint16_t* in, uint16_t* out);

/* Module raster: all active stages of p->module.  tex_l / tex_r are the R16 textures
 * bound as audio_l / audio_r.  out: h rows of w RGBA8 pixels, row 0 = bottom (GL). */
void orc_raster(const orc_params* p, const uint16_t* tex_l, const uint16_t* tex_r, uint8_t* out);
/* rows [y0, y1) only (for timing on bounded samples and for threading) */
void orc_raster_rows(const orc_params* p, const uint16_t* tex_l, const uint16_t* tex_r,
                     uint8_t* out, int y0, int y1);

/* FIFO ingest (fifo.c:89-110): slide ring left by `frames`, append int16 interleaved */
void orc_fifo_ingest(float* ring_l, float* ring_r, int n, const int16_t* interleaved,
                     int frames, int channels);

const char* orc_math_kind(void);

#ifdef __cplusplus
}
#endif
#endif
