/* glava_b200.h — C ABI of the B200-native GLava hot path (PCM -> spectrum -> pixels).
 *
 * Drop-in boundary: this is what a GLava maintainer binds instead of
 *   rd_new / rd_update / rd_destroy            (glava/render.h:53-60, glava/render.c:867,1743,2456)
 * for a BATCH of independent PCM streams.  Plain pointers and sizes only; no CUDA or
 * torch types appear in any signature.  All device memory is owned by the handle.
 * See INTEGRATION.md for the glava.c-side stub.
 *
 * Error behaviour: the reference has no error codes; fatal conditions call the
 * overridable fn-ptr `glava_abort` (glava/glava.h:17, glava/glava.c:77-80).  Here every
 * entry point returns 0 on success / negative GLAVA_B200_E* on failure AND reports the
 * message through `glava_b200_set_abort_hook` (default hook: print to stderr, do NOT exit).
 */
#ifndef GLAVA_B200_H
#define GLAVA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLAVA_B200_OK        0
#define GLAVA_B200_EINVAL   -1   /* bad argument / unsupported configuration */
#define GLAVA_B200_ECONFIG  -2   /* config file / #request error (reference: glava_abort) */
#define GLAVA_B200_ECUDA    -3   /* CUDA runtime failure (no device, OOM, launch error)   */

/* module ids: directory names under shaders/glava/ selected by `#request mod <name>`
 * (render.c:1100-1110) or -m / force_module (glava.c:391) */
enum { GLAVA_B200_MOD_BARS = 0, GLAVA_B200_MOD_RADIAL = 1, GLAVA_B200_MOD_CIRCLE = 2,
       GLAVA_B200_MOD_GRAPH = 3, GLAVA_B200_MOD_WAVE = 4, GLAVA_B200_MOD_TEST = 5 };

/* A colour macro of a module config (e.g. bars.glsl:20, radial.glsl:17, graph.glsl:11):
 *   mode 1  constant (lo)
 *   mode 0  mix(lo, hi, clamp(X / gradient, 0, 1))           — the shipped form, evaluated in closed form
 *   mode 2  any other GLSL expression of the per-pixel variable X (`d`; `pos` in graph) that the config reader could
 *           compile: the module's glava_b200_color_prog at the end of glava_b200_params is evaluated instead. */
typedef struct {
    int   mode;
    float lo[4], hi[4];
    float gradient;
} glava_b200_color;

/* Compiled colour expression: a straight-line program over 8 four-lane float registers (a scalar lives splat in all
 * four lanes, so scalar-vector arithmetic is lane-wise as in GLSL).  The value of the expression is register `result`.
 * Produced by glava_b200_load_config from the macro's text; every instruction is one individually rounded float op. */
enum { GLAVA_B200_COP_SPLAT = 0,   /* dst = imm in every lane                                   */
       GLAVA_B200_COP_VAR,         /* dst = X in every lane                                     */
       GLAVA_B200_COP_LANE,        /* dst.lane[a] = imm                                         */
       GLAVA_B200_COP_SHUF,        /* dst.lane[k] = reg[a].lane[(imm >> 3k) & 7] unless that selector is 7 */
       GLAVA_B200_COP_ADD, GLAVA_B200_COP_SUB, GLAVA_B200_COP_MUL, GLAVA_B200_COP_DIV,   /* dst = a op b */
       GLAVA_B200_COP_MIN, GLAVA_B200_COP_MAX, GLAVA_B200_COP_MOD, GLAVA_B200_COP_STEP,  /* step(edge = a, x = b) */
       GLAVA_B200_COP_NEG, GLAVA_B200_COP_ABS, GLAVA_B200_COP_FLOOR, GLAVA_B200_COP_CEIL, GLAVA_B200_COP_FRACT,
       GLAVA_B200_COP_SQRT, GLAVA_B200_COP_SIN, GLAVA_B200_COP_COS, GLAVA_B200_COP_LOG, GLAVA_B200_COP_SIGN,
       GLAVA_B200_COP_TRUNC,       /* dst = f(a)                                                */
       GLAVA_B200_COP_MIX,         /* dst = a * (1 - c) + b * c,            c = reg[(int) imm]  */
       GLAVA_B200_COP_CLAMP,       /* dst = min(max(a, b), c)                                   */
       GLAVA_B200_COP_SMOOTHSTEP,  /* edges a, b, x = c                                         */
       GLAVA_B200_COP_LT, GLAVA_B200_COP_LE, GLAVA_B200_COP_EQ, GLAVA_B200_COP_NE,   /* dst = (a op b) ? 1 : 0  (> and >= swap operands) */
       GLAVA_B200_COP_AND, GLAVA_B200_COP_OR, GLAVA_B200_COP_NOT,                    /* on 0 / 1 values                                  */
       GLAVA_B200_COP_SELECT,      /* dst = c != 0 ? a : b,                  c = reg[(int) imm]  */
       GLAVA_B200_COP_EXP, GLAVA_B200_COP_EXP2, GLAVA_B200_COP_LOG2,                 /* dst = f(a)               */
       GLAVA_B200_COP_POW,         /* dst = pow(a, b)                                            */
       GLAVA_B200_COP_ATAN2,       /* dst = atan(a, b); atan(x) is compiled as atan(x, 1)        */
       GLAVA_B200_COP_TAN,         /* dst = sin(a) / cos(a)                                      */
       GLAVA_B200_COP_COUNT };
#define GLAVA_B200_COLOR_OPS  64
#define GLAVA_B200_COLOR_REGS 8
typedef struct { uint8_t op, dst, a, b; float imm; } glava_b200_color_op;
typedef struct { int n_ops, result; glava_b200_color_op ops[GLAVA_B200_COLOR_OPS]; } glava_b200_color_prog;

/* Everything rc.glsl / smooth_parameters.glsl / <module>.glsl configure on this path.
 * Field meaning follows the `#request` (render.c:1033-1314) or `#define` named in the
 * comment.  Defaults = the shipped config files. */
typedef struct {
    /* spectrum */
    int   n;                  /* setbufsize (rc.glsl:190): floats per channel, power of two 256..16384 */
    float fft_scale;          /* setfftscale      smooth_parameters.glsl:46 */
    float fft_cutoff;         /* setfftcutoff     :51 */
    float gravity_step;       /* setgravitystep   :67 */
    float ur;                 /* updates / second the gravity step is divided by (render.c:728,2224);
                                 nominal setsamplerate / (setsamplesize / 4) */
    int   avg_frames;         /* setavgframes     :56 (1..16) */
    int   avg_window;         /* setavgwindow     :61 */
    int   accel_fft;          /* setaccelfft      rc.glsl:211 — 0: pipeline A (render.c:2149-2156 float
                                 chain), 1: pipeline B (R16 passes, render.c:2188-2267) */
    int   smooth_pass;        /* setsmoothpass    :78 */
    float smooth_factor;      /* setsmoothfactor  :72 — as the shaders see it: the "%.6f" literal of the injected
                                 `#define _SMOOTH_FACTOR` (render.c:315-324); the config reader rounds accordingly */
    float sample_range;       /* #define SAMPLE_RANGE :42 */
    float sample_scale;       /* #define SAMPLE_SCALE :37 */
    float hybrid_weight;      /* #define SAMPLE_HYBRID_WEIGHT :34 */
    int   sample_mode;        /* #define SAMPLE_MODE: 0 average, 1 maximum, 2 hybrid */
    int   round_formula;      /* #define ROUND_FORMULA: 0 sinusoidal, 1 linear, 2 circular */
    /* raster, common */
    int   module;             /* GLAVA_B200_MOD_* */
    int   w, h;               /* setgeometry w h (rc.glsl:52) */
    int   channels;           /* _CHANNELS: setmirror ? 1 : 2 (render.c:290) */
    int   premultiply_alpha;  /* setopacity "native" (render.c:1036-1040); 0: per-stage GL blending over clear_color */
    /* bars.glsl */
    float bars_width, bars_gap, bars_outline_width, bars_amplify;
    glava_b200_color bars_color;
    int   bars_outline_mode;  /* 0: vec4(COLOR.rgb * 1.5, COLOR.a) (bars.glsl:22), 1: constant, 2: bars_outline_prog */
    float bars_outline[4];
    int   bars_direction, bars_invert, bars_flip, bars_mirror_yx;
    /* radial.glsl */
    float radial_radius, radial_line;
    float radial_line_half;   /* value of the GLSL expression (C_LINE / 2), integer division for an int literal */
    float radial_outline[4];
    int   radial_nbars; float radial_bar_width, radial_amplify;
    glava_b200_color radial_color;
    float radial_rotate; int radial_invert;
    float radial_bar_alias, radial_c_alias, radial_off_x, radial_off_y;
    /* circle.glsl */
    float circle_radius, circle_line, circle_outline[4], circle_amplify, circle_rotate;
    int   circle_invert, circle_fill, circle_smooth;
    /* graph.glsl */
    float graph_vscale; int graph_direction; glava_b200_color graph_color;
    int   graph_draw_outline, graph_draw_highlight; float graph_outline[4]; int graph_invert;
    /* wave.glsl */
    float wave_min_thickness, wave_max_thickness, wave_base_color[4], wave_amplify, wave_outline[4];
    /* request read-backs the caller sizes its audio with (render.h:10-24, glava.c:487-514) */
    int   rate_request;       /* setsamplerate  rc.glsl:203 */
    int   samplesize_request; /* setsamplesize  rc.glsl:181 */
    /* engine knobs (no reference equivalent) */
    int   fb_slots;           /* framebuffers resident per device; 0 = one per stream (default),
                                 K < batch = ring of K (stream s renders into slot s % K) */
    int   lazy_smooth;        /* 1: K5 evaluates only the texels the module samples, and gravity / average state is kept only
                                 for the bins those texels' taps reach (same pixels; glava_b200_textures() / _spectrum() then hold
                                 only those texels / bins, the rest is stale or zero); 0: all n texels.  After a
                                 glava_b200_reconfigure that widens the sampled set (AMPLIFY does not; BAR_WIDTH, geometry,
                                 setsmoothfactor do), the newly reached bins start from the state they had when they were last
                                 updated (zero at creation): their bars fade in over setavgframes updates, like a GLava restart */
    /* optional stages of rd_update that the shipped configuration leaves off */
    int   bufscale;           /* setbufscale (rc.glsl:236, deprecated): box-average `bufscale` PCM samples
                                 before anything else (render.c:1765-1790); textures then have n / bufscale texels */
    int   interpolate;        /* setinterpolate (rc.glsl:131): keyframe lerp between the last two post-transform
                                 buffers (render.c:1792-1809, 2347-2353).  As in the reference it is inactive when
                                 ur / fr > 0.9 (render.c:1761-1763) and for an fft module under setaccelfft
                                 (render.c:2161-2168) */
    float fr;                 /* glava_b200_update calls per second, the reference's measured gl->fr
                                 (render.c:2386); setframerate when > 0; 0 = same as ur */
    int   transform_smooth;   /* `#request transform <uniform> "smooth"` (render.c:1218-1286; transform_smooth, render.c:694-718)
                                 in the module's bind list — the module shaders are not read here, so the position is a parameter:
                                 1 = after "fft" ... "avg": forces the CPU-order chain (pipeline A, render.c:2143-2154);
                                 2 = BEFORE "fft": applied to the (scaled) PCM ring on the CPU side of handle_audio, the fft chain
                                     and setaccelfft's GL passes follow unchanged (render.c:2131-2156) */
    float smooth_distance;    /* setsmooth       (render.c:917,1201) */
    float smooth_ratio;       /* setsmoothratio  (render.c:918,1204) */
    /* compiled colour expressions (mode 2 of the colour they belong to; n_ops == 0 otherwise) */
    glava_b200_color_prog bars_color_prog, bars_outline_prog, radial_color_prog, graph_color_prog;
    /* setbg / setbgf (render.c:1062-1099; rc.glsl:56 `setbg 00000000`): the glClear colour.  Native opacity never shows
     * it (blending is off and every stage writes every pixel); with premultiply_alpha == 0 (setopacity "none" / "xroot")
     * every module stage is blended over it with SRC_ALPHA / ONE_MINUS_SRC_ALPHA (render.c:1467-1470) */
    float clear_color[4];
    /* radial.glsl details appended later (kept at the end so earlier offsets stay put) */
    int   radial_bar_width_int;      /* BAR_WIDTH was written as an integer literal: the shader's `BAR_WIDTH / 2`
                                        (radial/1.frag:62,79,88) is then an integer division */
    float radial_bar_outline_width;  /* BAR_OUTLINE_WIDTH (deprecated, radial.glsl:33-36; default 0) */
    float radial_bar_outline[4];     /* BAR_OUTLINE (default: OUTLINE) */
    int   graph_join_channels;       /* JOIN_CHANNELS (graph.glsl:23): the two halves meet at a common height in the middle */
    int   graph_anti_alias;          /* ANTI_ALIAS (graph.glsl:19): graph/3.frag fades the column steps of the line */
    int   shader_pre_smoothed;       /* what the module's stage-1 shader believes about its audio textures
                                        (`_PRE_SMOOTHED_AUDIO`): 0 = what smooth_pass says (consistent), 1 = "already smoothed"
                                        although the K5 pass is off, 2 = "raw" although K5 ran (smoothed twice).  The reference
                                        gets into 1 / 2 when smooth_parameters.glsl flips setsmoothpass: the module's first
                                        shader header is built before its includes' requests run (render.c:284-293, 312);
                                        the config reader reproduces that */
    int   mirror_input;              /* setmirror as the AUDIO side sees it (r->mirror_input, render.c:1054-1058: the backend mixes
                                        both channels into one, fifo.c:98-102).  `channels` above is what the SHADER sees
                                        (`_CHANNELS`), which bars' `DISABLE_MONO 1` turns back into 2 (bars/1.frag:32-34) while the
                                        rings stay mono.  0 with channels == 1 is treated as 1 (older callers set only channels) */
} glava_b200_params;

typedef struct glava_b200 glava_b200;    /* plays the role of struct glava_renderer (render.h:8-30) */

/* Fatal-error hook, the analogue of the `glava_abort` fn-ptr (glava.h:17). NULL restores the default. */
void glava_b200_set_abort_hook(void (*hook)(const char* message));
const char* glava_b200_last_error(void);

/* Shipped-default parameters for a module name ("bars", "radial", "circle", "graph", "wave", "test"). */
int glava_b200_default_params(glava_b200_params* out, const char* module);

/* Config surface — the part of rd_new (render.c:1322-1435) that reads `entry` (rc.glsl) from the
 * first directory in the NULL-terminated `paths` that has it, applies `#request`s, then the
 * NULL-terminated `requests` strings (CLI --request, render.c:1415-1435), then reads
 * smooth_parameters.glsl and <module>.glsl (`#define`s) with user-dir-over-system-dir precedence.
 * `force_module` = the -m option (may be NULL).  paths may be NULL: shipped defaults. */
int glava_b200_load_config(glava_b200_params* out, const char* const* paths, const char* entry,
                           const char* const* requests, const char* force_module);

/* Same, with `--pipe` binds (glava.c:421-436): NULL-terminated "name=value" strings; a config macro written
 * `@name:default` (glsl_ext.c:516-591) takes `value` (#rrggbb[aa] or vec4(...)) when `name` is bound. */
int glava_b200_load_config_binds(glava_b200_params* out, const char* const* paths, const char* entry,
                                 const char* const* requests, const char* force_module, const char* const* binds);

/* Live `--pipe` binds (glava.c:338-411, render.c:1846-2100).  pipe_args: NULL-terminated "NAME[:TYPE]" strings exactly
 * as given to `--pipe` (TYPE one of int, float, bool, vec2, vec3, vec4; default vec4; a NULL-equivalent name is "_"),
 * validated with the reference's messages.  The other arguments are those of glava_b200_load_config and are copied.
 *   _feed    bytes as they arrive on stdin; every complete line `name = value` (or just `value` for the first bind)
 *            is parsed with the reference's rules — prefix match of the name, "#rrggbb[aa]" or "r,g,b,a" for vec4,
 *            true/TRUE/True/1 for bool ... — and its messages ("Bad assignment format", "Variable name not bound",
 *            "Bad format for color string").  Returns the number of binds that took a new value.
 *   _params  the configuration re-evaluated with every `@name:default` macro of a bound name reading the bind's current
 *            value (0 / vec4(0) until its first line, like an unwritten GL uniform).
 *   _apply   _params + glava_b200_reconfigure when a bind changed since the last call: the uniform write of
 *            render.c:2071-2100.  Call once per frame after _feed. */
typedef struct glava_b200_pipe glava_b200_pipe;
glava_b200_pipe* glava_b200_pipe_new(const char* const* paths, const char* entry, const char* const* requests,
                                     const char* force_module, const char* const* pipe_args);
int  glava_b200_pipe_feed(glava_b200_pipe* p, const char* bytes, size_t len);
int  glava_b200_pipe_params(glava_b200_pipe* p, glava_b200_params* out);
int  glava_b200_pipe_bind_count(const glava_b200_pipe* p);
int  glava_b200_pipe_bind(const glava_b200_pipe* p, int index, const char** name, const char** type, float value[4]);
void glava_b200_pipe_free(glava_b200_pipe* p);

/* rd_new (render.h:53-57): build a renderer for `batch` independent streams on CUDA device
 * `device` with the given parameters (from glava_b200_load_config / _default_params). */
glava_b200* glava_b200_new(const glava_b200_params* params, int batch, int device);
/* rd_destroy (render.h:60) */
void glava_b200_destroy(glava_b200* r);

/* Live parameter update — the analogue of a `--pipe` uniform write (render.c:1846-2005): colours, AMPLIFY,
 * gradient, smoothing parameters ... take effect from the next update.  Everything that sizes device state
 * (setbufsize, geometry, module, setaccelfft, setavgframes, fb_slots) must be unchanged. */
int  glava_b200_reconfigure(glava_b200* r, const glava_b200_params* params);
int  glava_b200_get_params(const glava_b200* r, glava_b200_params* out);
int  glava_b200_pipe_apply(glava_b200_pipe* p, glava_b200* r);   /* see glava_b200_pipe_new */
int  glava_b200_batch(const glava_b200* r);
const char* glava_b200_module_name(const glava_b200* r);

/* Pinned host memory for the PCM rings handed to glava_b200_update (so the H2D copy is a
 * true async DMA).  Plain malloc'd memory is accepted too, just slower. */
void* glava_b200_host_alloc(size_t bytes);                 /* placed on the NUMA node of the CURRENT CUDA device */
void* glava_b200_host_alloc_on(size_t bytes, int device);  /* ... of `device` (anonymous mapping + mbind + cudaHostRegister;
                                                              plain cudaHostAlloc when the topology is not exposed or
                                                              GLAVA_B200_NO_NUMA is set) */
void  glava_b200_host_free(void* p);
/* NUMA helpers for one-process-per-GPU callers: the node a device hangs off (-1: unknown), and pinning the calling
 * thread to that node's CPUs (what `numactl --cpunodebind` would do; returns the node or -1, nothing changed). */
int   glava_b200_device_numa_node(int device);
int   glava_b200_bind_thread_to_device(int device);

/* rd_update (render.h:58-59; render.c:1743-2417), batched.
 *   lb, rb : HOST, [batch][bsz] float32, ring contents oldest-first, exactly what glava.c:528-537
 *            memcpy's into lb/rb for one stream.  NOT modified (the reference transforms them in
 *            place, render.c:2140-2180).  rb is ignored by `wave` (audio_l only, wave/1.frag:7); for every
 *            other module a NULL rb with modified != 0 is GLAVA_B200_EINVAL.
 *   bsz    : must equal params.n (setbufsize; with setbufscale k the spectrum has n / k entries)
 *   modified: as rd_update's flag; 0 re-rasters the last spectrum (render.c:2268-2272), or, with keyframe
 *            interpolation active, the next interpolated one
 * Copies H2D on a dedicated copy stream (double-buffered staging: the copy of update i+1 overlaps the
 * kernels of update i), runs the fused spectrum kernel and the module raster kernel on the handle's
 * stream.  Returns once the host buffers have been consumed (they may be reused immediately) and the
 * kernels are enqueued; call glava_b200_sync to wait for the frame. */
int glava_b200_update(glava_b200* r, const float* lb, const float* rb, size_t bsz, int modified);
/* Double-buffer contract for callers that must not block: after glava_b200_set_async_input(r, 1) the update calls return
 * once the H2D copy is ENQUEUED; lb / rb must stay untouched until glava_b200_wait_input(r) (or the next glava_b200_sync).
 * A caller alternating between two sets of rings never waits on a copy. */
int glava_b200_set_async_input(glava_b200* r, int enable);
int glava_b200_wait_input(glava_b200* r);
/* same with DEVICE pointers (inputs already resident in HBM).  The kernels read d_lb / d_rb (16-byte aligned) on an INTERNAL
 * non-blocking stream: either the buffers are complete before the call and stay untouched until glava_b200_sync, or the
 * caller orders its own stream with the two event hooks below. */
int glava_b200_update_device(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, int modified);
/* ready_event: a cudaEvent_t (as void*) recorded on the producer's stream after d_lb / d_rb were written; the update waits
 * for it on the device.  glava_b200_input_event(): a cudaEvent_t recorded after the kernels that READ the buffers of the
 * updates issued so far — cudaStreamWaitEvent(producer, ev) before overwriting them. */
int   glava_b200_update_device_after(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, int modified, void* ready_event);
void* glava_b200_input_event(glava_b200* r);

/* Per-stream `modified` (glava.c:528-537 tests the flag of ITS audio thread; in a batch every stream has its own):
 * modified[s] != 0 = stream s has new audio and runs the whole chain; a stream with 0 keeps its gravity / average state
 * untouched and is re-rastered from its previous texture, exactly what rd_update(modified = false) does for it
 * (render.c:2122, 2268-2272).  All zero / all non-zero are the plain calls above.  After the first uneven update every
 * stream carries its own average-ring cursor; plain modified = 1 calls keep working.  lb / rb rows of unmodified
 * streams are not read by the kernels.  Not available (GLAVA_B200_EINVAL) for an uneven mask while keyframe
 * interpolation is active (setinterpolate with ur / fr <= 0.9: the keyframe rotation belongs to the renderer). */
int glava_b200_update_masked(glava_b200* r, const float* lb, const float* rb, size_t bsz, const uint8_t* modified /* [batch] */);
int glava_b200_update_device_masked(glava_b200* r, const float* d_lb, const float* d_rb, size_t bsz, const uint8_t* modified);
int glava_b200_update_rings_masked(glava_b200* r, const uint8_t* modified);

/* FIFO-compatible ingest (fifo.c:89-110): `frames` new interleaved int16 L,R frames per stream,
 * HOST [batch][frames*2]; slides every stream's device-resident ring and converts s16/65535.f
 * (mono: integer mean first, fifo.c:98-102, when params.mirror_input — or, for callers that only set it, channels == 1).  Then
 * glava_b200_update_rings() runs the update on the resident rings. */
int glava_b200_ingest_fifo(glava_b200* r, const int16_t* chunks, int frames);
/* The PulseAudio backend's ring update (pulse_input.c:146-174): the samples are ALREADY float (PA_SAMPLE_FLOAT32NE, no
 * / 65535), HOST [batch][frames*2] interleaved L,R; params.channels == 1 mixes (l + r) / 2 in float into both rings. */
int glava_b200_ingest_float(glava_b200* r, const float* chunks, int frames);
int glava_b200_update_rings(glava_b200* r, int modified);

int glava_b200_sync(glava_b200* r);

/* Outputs.  Frame layout: RGBA8, [h][w][4] bytes, row 0 = bottom row (GL window coordinates),
 * byte order R,G,B,A. */
int glava_b200_readback(glava_b200* r, int stream, uint8_t* rgba);              /* one stream's frame -> HOST */
int glava_b200_readback_async(glava_b200* r, int stream, uint8_t* rgba);        /* same, asynchronous: the frame is snapshotted on the
                                                                                   handle's stream (D2D) and copied out on a separate
                                                                                   one, under the next update's kernels; rgba (pinned)
                                                                                   is valid after glava_b200_sync */
int glava_b200_readback_fence(glava_b200* r);                                   /* orders later work on glava_b200_cuda_stream() after
                                                                                   the read-backs issued so far (for event timing) */
int glava_b200_spectrum(glava_b200* r, float* out_l, float* out_r);             /* HOST [batch][n]: pipeline-A result
                                                                                   (accel_fft 0) or raw transform_fft output (1);
                                                                                   n / bufscale entries per stream */
int glava_b200_spectrum_size(const glava_b200* r);                              /* entries per channel of spectra / textures */
int glava_b200_textures(glava_b200* r, uint16_t* out_l, uint16_t* out_r);       /* HOST [batch][n] R16 texels the module samples */
const void* glava_b200_framebuffer_device(const glava_b200* r);                 /* DEVICE [fb_slots][h][w] RGBA8 */
void* glava_b200_cuda_stream(const glava_b200* r);                              /* cudaStream_t, for event timing */

/* Offscreen hand-off — the analogues of glava_sizereq / glava_wait / glava_tex (glava.h:22-24, glava.c:244-267) for a
 * consumer (compositor, encoder) that takes frames straight from HBM:
 *   glava_b200_sizereq         thread-safe resize request, applied at the start of the next update (render.c:1811-1830);
 *                              spectrum state is kept, framebuffers and geometry tables are rebuilt
 *   glava_b200_wait_frame      host-blocks until the latest frame is complete (glava_wait)
 *   glava_b200_frame_event     cudaEvent_t recorded after the latest raster: cudaStreamWaitEvent(consumer, ev) orders a
 *                              consumer stream after the frame without a host round trip (glava_tex hands out the GL
 *                              texture id; here the "texture" is glava_b200_frame_device / _framebuffer_device)
 *   glava_b200_framebuffer_ipc cudaIpcMemHandle_t (64 bytes) of the framebuffer array, for a consumer in another
 *                              process (the OBS-plugin situation, glava-obs/entry.c:141-214) */
int   glava_b200_sizereq(glava_b200* r, int w, int h);
int   glava_b200_wait_frame(glava_b200* r);
void* glava_b200_frame_event(const glava_b200* r);
const void* glava_b200_frame_device(const glava_b200* r, int stream);
int   glava_b200_framebuffer_ipc(glava_b200* r, void* handle, size_t handle_bytes);

/* ---- one handle for a batch spread over several GPUs of a node (csrc/sharded.cpp) -------------------------------------
 * The streams of a batch are independent (one GLava process each, in the reference), so the batch is cut into contiguous
 * blocks, one per device — glava_b200_shard_range: counts differ by at most one — and nothing crosses between devices (no
 * collective, NCCL unused).  One worker thread per device, pinned to the device's NUMA node; every call below is issued
 * on all shards concurrently.  lb / rb / chunks / modified / textures are indexed by GLOBAL stream, [batch][...].
 *   device_mask: bit d = CUDA device d takes a shard (0 = every visible device).  _devices: explicit ordinals (an
 *   ordinal may repeat: several shards on one device).  Error behaviour as everywhere: 0 / negative code + abort hook;
 *   the message names the failing device and its stream block. */
typedef struct glava_b200_sharded glava_b200_sharded;
int  glava_b200_device_count(void);
int  glava_b200_shard_range(int batch, int shards, int k, int* first_stream, int* count);
glava_b200_sharded* glava_b200_new_sharded(const glava_b200_params* params, int batch, uint64_t device_mask);
glava_b200_sharded* glava_b200_new_sharded_devices(const glava_b200_params* params, int batch, const int* devices, int n);
void glava_b200_sharded_destroy(glava_b200_sharded* s);
int  glava_b200_sharded_shards(const glava_b200_sharded* s);
int  glava_b200_sharded_batch(const glava_b200_sharded* s);
glava_b200* glava_b200_sharded_shard(const glava_b200_sharded* s, int k, int* device, int* first_stream, int* count);
int  glava_b200_sharded_update(glava_b200_sharded* s, const float* lb, const float* rb, size_t bsz, const uint8_t* modified /* NULL: all */);
int  glava_b200_sharded_rerender(glava_b200_sharded* s);                                  /* rd_update(modified = false) */
int  glava_b200_sharded_ingest_fifo(glava_b200_sharded* s, const int16_t* chunks, int frames);   /* ingest + update on the rings */
int  glava_b200_sharded_sync(glava_b200_sharded* s);
int  glava_b200_sharded_readback(glava_b200_sharded* s, int stream, uint8_t* rgba);
const void* glava_b200_sharded_frame_device(glava_b200_sharded* s, int stream, int* device);
int  glava_b200_sharded_textures(glava_b200_sharded* s, uint16_t* out_l, uint16_t* out_r);

/* Stage-wise entry points (used by the parity tests; same kernels as the fused path). */
int glava_b200_smooth_pass(glava_b200* r, const uint16_t* in, uint16_t* out, int count);      /* K5 on HOST [count][n] */
int glava_b200_raster_textures(glava_b200* r, const uint16_t* tex_l, const uint16_t* tex_r);  /* HOST [batch][n] -> raster */
int glava_b200_transform_smooth(glava_b200* r, float* planes, int count);                     /* transform_smooth in place on HOST [count][n] */

/* Kernel launch counters since creation (bench.py `gpu_launches`). */
uint64_t glava_b200_launch_count(const glava_b200* r);

/* Per-kernel device timing: when enabled, CUDA events are recorded on the handle's stream around
 * the spectrum kernel and the raster kernel of every update; glava_b200_kernel_times() synchronises
 * and returns the summed durations (ms) and launch counts since timing was (re-)enabled. */
int glava_b200_set_timing(glava_b200* r, int enable);
int glava_b200_kernel_times(glava_b200* r, double* spectrum_ms, int* spectrum_launches,
                            double* raster_ms, int* raster_launches);
/* development aid: [start, end] ms of every timed launch relative to the first one (spectrum pairs, then raster pairs) */
int glava_b200_timeline(glava_b200* r, double* out, int cap_pairs, int* n_spec, int* n_ras);

/* ---- audio plug-in ABI kept verbatim from the reference (fifo.h:9-26) so a GLava audio
 * backend can feed this renderer: see INTEGRATION.md. ---- */

const char* glava_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GLAVA_B200_H */
