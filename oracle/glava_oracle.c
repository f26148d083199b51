/* TEST INFRASTRUCTURE — see glava_oracle.h.  CPU restatement, compiled by oracle/Makefile
 * with -ffp-contract=off so that every float operation is an individually rounded IEEE
 * binary32 operation (what GLSL `float` is on llvmpipe, and what render.c's float code
 * is on x86-64 without FMA contraction).
 *
 * GLSL semantics fixed here (all documented in DESIGN.md §"GLSL semantics"):
 *   - float literals and uniforms are binary32; int/int is integer division
 *   - mod(x, y)  = x - y * floor(x / y)
 *   - round(x)   = nearest, ties to even
 *   - mix(a,b,t) = a * (1 - t) + b * t
 *   - float -> unorm8/unorm16 store = lrintf(clamp(c, 0, 1) * MAX): ONE rounding, ties to even (round 1 used
 *     (int)(c * MAX + 0.5f), whose second float rounding is off by one in a band around every half), NaN -> 0
 *   - unorm -> float fetch          = (float) u / MAX
 *   - texelFetch outside [0, size)  = 0 (llvmpipe behaviour; undefined in GL)
 *   - texture() on the 1-D audio textures: NEAREST + REPEAT (render.c:510-518)
 *   - uninitialised R16 render targets (gr_store, ring slots; render.c:1717) read 0
 *   - #rrggbb colour literals are the "%.6f" decimal strings glsl_ext.c:505 prints
 */
#include "glava_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_PRODUCT_MATH
#include "gl_math.h"
#define G_SIN(x)      glm_sin(x)
#define G_ATAN2(y, x) glm_atan2(y, x)
#define G_LOG(x)      glm_log(x)
#define G_POW(x, k)   ((k) == 3 ? ((x) * (x)) * (x) : (x) * (x))   /* gl_math has no pow: exact products for the two integer powers used */
const char* orc_math_kind(void) { return "product-gl_math"; }
#else
#define G_SIN(x)      sinf(x)
static float orc_atan2f(float y, float x) { return (x == 0.0f && y == 0.0f) ? 0.0f : atan2f(y, x); }
#define G_ATAN2(y, x) orc_atan2f(y, x)
#define G_LOG(x)      logf(x)
#define G_POW(x, k)   powf((x), (float) (k))
const char* orc_math_kind(void) { return "libm"; }
#endif

#define G_TWOPI 6.28318530718f
#define G_PI    3.14159265359f

static inline float g_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float g_min(float a, float b) { return b < a ? b : a; }
static inline float g_max(float a, float b) { return a < b ? b : a; }
static inline float g_mod(float x, float y) { return x - y * floorf(x / y); }
static inline float g_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float g_round(float x) { return rintf(x); }
/* clamp, then the float32 product rounded to nearest-even (Mesa's _mesa_float_to_unorm; pinned by the llvmpipe goldens:
 * the reference's uploads are reproduced bit for bit); a NaN input (0/0 weight in K5) stores 0 */
static inline uint32_t unorm8(float c)  { return c > 0.0f ? (c < 1.0f ? (uint32_t) lrintf(c * 255.0f) : 255u) : 0u; }
static inline uint16_t unorm16(float c) { return c > 0.0f ? (c < 1.0f ? (uint16_t) lrintf(c * 65535.0f) : 65535u) : 0u; }
static inline float from8(uint32_t u)  { return (float) u / 255.0f; }
static inline float from16(uint16_t u) { return (float) u / 65535.0f; }

typedef struct { float r, g, b, a; } vec4;
static inline vec4 v4(float r, float g, float b, float a) { vec4 v = { r, g, b, a }; return v; }
static inline vec4 v4a(const float* c) { return v4(c[0], c[1], c[2], c[3]); }
static inline uint32_t pack8(vec4 c) {
    return unorm8(c.r) | (unorm8(c.g) << 8) | (unorm8(c.b) << 16) | (unorm8(c.a) << 24);
}
static inline vec4 unpack8(uint32_t u) {
    return v4(from8(u & 255u), from8((u >> 8) & 255u), from8((u >> 16) & 255u), from8(u >> 24));
}
static inline vec4 g_mix(vec4 a, vec4 b, float t) {
    float s = 1.0f - t;
    return v4(a.r * s + b.r * t, a.g * s + b.g * t, a.b * s + b.b * t, a.a * s + b.a * t);
}
static inline vec4 eval_color(const orc_color* c, float x) {
    if (c->mode == 1) return v4a(c->lo);
    return g_mix(v4a(c->lo), v4a(c->hi), g_clamp(x / c->gradient, 0.0f, 1.0f));
}

/* ------------------------------------------------------------------------------------ */
/* defaults: shipped rc.glsl / smooth_parameters.glsl / <module>.glsl */

/* "%.6f" of hex/255 as glsl_ext.c:505 emits, parsed back as a GLSL float literal */
static float hexc(int v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "%.6f", (double) ((float) v / (float) 255));   /* glsl_ext.c:113,505 */
    return strtof(buf, NULL);
}
static void hex3(float* out, int r, int g, int b) { out[0] = hexc(r); out[1] = hexc(g); out[2] = hexc(b); out[3] = 1.0f; }

void orc_default_params(orc_params* p, int module, int n, int w, int h) {
    memset(p, 0, sizeof(*p));
    p->n = n; p->fft_scale = 10.2f; p->fft_cutoff = 0.3f; p->gravity_step = 4.2f;
    p->ur = 22050.0f / 256.0f;
    p->avg_frames = 5; p->avg_window = 1; p->accel_fft = 1; p->smooth_pass = 1;
    p->smooth_factor = 0.025f; p->sample_range = 0.9f; p->sample_scale = 8.0f;
    p->hybrid_weight = 0.65f; p->sample_mode = 0; p->round_formula = 0;
    p->module = module; p->w = w; p->h = h; p->channels = 2; p->premultiply_alpha = 1;
    /* bars.glsl */
    p->bars_width = 5; p->bars_gap = 1; p->bars_outline_width = 1; p->bars_amplify = 300;
    p->bars_color.mode = 0; hex3(p->bars_color.lo, 0x33, 0x66, 0xb2); hex3(p->bars_color.hi, 0xa0, 0xa0, 0xb2);
    p->bars_color.gradient = 80; p->bars_outline_mode = 0;
    /* radial.glsl */
    p->radial_radius = 128; p->radial_line = 2; p->radial_line_half = 1; hex3(p->radial_outline, 0x33, 0x33, 0x33);
    hex3(p->radial_bar_outline, 0x33, 0x33, 0x33); p->radial_bar_outline_width = 0; p->radial_bar_width_int = 0;
    p->radial_nbars = 160; p->radial_bar_width = 4.5f; p->radial_amplify = 300;
    p->radial_color.mode = 0; hex3(p->radial_color.lo, 0xcc, 0x33, 0x33); hex3(p->radial_color.hi, 0xcc, 0xa0, 0xa0);
    p->radial_color.gradient = 95; p->radial_rotate = G_PI / 2; p->radial_bar_alias = 1.2f; p->radial_c_alias = 1.8f;
    /* circle.glsl */
    p->circle_radius = 128; p->circle_line = 1.5f; hex3(p->circle_outline, 0x33, 0x33, 0x33);
    p->circle_amplify = 150; p->circle_rotate = G_PI / 2; p->circle_smooth = 1;
    /* graph.glsl */
    p->graph_vscale = 300; p->graph_direction = 1;
    p->graph_color.mode = 0; hex3(p->graph_color.lo, 0x80, 0x2a, 0x2a); hex3(p->graph_color.hi, 0x4f, 0x4f, 0x92);
    p->graph_color.gradient = 75; p->graph_draw_highlight = 1; hex3(p->graph_outline, 0x26, 0x26, 0x26);
    /* wave.glsl */
    p->wave_min_thickness = 1; p->wave_max_thickness = 6; p->wave_amplify = 500;
    p->wave_base_color[0] = 0.7f; p->wave_base_color[1] = 0.2f; p->wave_base_color[2] = 0.45f; p->wave_base_color[3] = 1;
    p->wave_outline[0] = p->wave_outline[1] = p->wave_outline[2] = 0.15f; p->wave_outline[3] = 1;
}

/* ------------------------------------------------------------------------------------ */
/* transform_fft — render.c:783-847                                                      */

void orc_window(double* w, int n) {
    /* render.c:660 `window(t, sz)` called as window(i, s->sz - 1) at render.c:794: the
     * unparenthesised macro argument makes the phase TWOPI*i/sz - 1 (period sz).  TWOPI
     * is render.c:64's 6.28318530718 literal. */
    for (int i = 0; i < n; ++i)
        w[i] = 0.53836 - (0.46164 * cos(6.28318530718 * (double) i / (double) n - 1));
}

static unsigned bitrev(unsigned v, int bits) {
    unsigned r = 0;
    for (int b = 0; b < bits; ++b) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

void orc_fft_f32(const orc_params* p, float* buf) {
    const int n = p->n, half = n / 2;               /* half complex points z[k] = buf[2k] + i buf[2k+1] */
    int bits = 0; while ((1 << bits) < half) ++bits;
    for (int i = 0; i < n; ++i)                     /* render.c:793-795: float *= double -> float */
        buf[i] = (float) ((double) buf[i] * (0.53836 - (0.46164 * cos(6.28318530718 * (double) i / (double) n - 1))));
    for (int k = 0; k < half; ++k) {                /* render.c:797-812 bit-reversal permutation */
        int r = (int) bitrev((unsigned) k, bits);
        if (r > k) {
            float t0 = buf[2 * k], t1 = buf[2 * k + 1];
            buf[2 * k] = buf[2 * r]; buf[2 * k + 1] = buf[2 * r + 1];
            buf[2 * r] = t0; buf[2 * r + 1] = t1;
        }
    }
    /* render.c:814-840 Danielson-Lanczos; span = butterflies' half-distance in complex points.
     * Same float recurrence for the twiddle (wr, wi) and the same operation order. */
    for (int span = 1; span < half; span <<= 1) {
        float theta = (float) -(2 * M_PI / (double) (2 * span));
        float wtemp = (float) sin(0.5 * (double) theta);
        float wpr = (float) (-2.0 * (double) wtemp * (double) wtemp);
        float wpi = (float) sin((double) theta);
        float wr = 1.0f, wi = 0.0f;
        for (int m = 0; m < span; ++m) {
            for (int k = m; k < half; k += 2 * span) {
                int j = k + span;
                float tr = wr * buf[2 * j] - wi * buf[2 * j + 1];
                float ti = wr * buf[2 * j + 1] + wi * buf[2 * j];
                buf[2 * j]     = buf[2 * k] - tr;
                buf[2 * j + 1] = buf[2 * k + 1] - ti;
                buf[2 * k]     += tr;
                buf[2 * k + 1] += ti;
            }
            wtemp = wr;
            wr += wr * wpr - wi * wpi;
            wi += wi * wpr + wtemp * wpi;
        }
    }
    for (int i = 0; i < n; ++i) {                   /* render.c:842-846 */
        float v = buf[i];
        if (v < 0.0f) v = -v;
        v = (float) (log((double) (v + 1)) / 3);
        float ramp = (((float) i / (float) n) * p->fft_scale) + (1.0f - p->fft_cutoff);
        v *= ramp > 1.0f ? ramp : 1.0f;
        buf[i] = v;
    }
}

void orc_fft_f64(const orc_params* p, const float* in, double* out) {
    const int n = p->n, half = n / 2;
    double* re = malloc(sizeof(double) * half), * im = malloc(sizeof(double) * half);
    double* wn = malloc(sizeof(double) * n);
    orc_window(wn, n);
    int bits = 0; while ((1 << bits) < half) ++bits;
    for (int k = 0; k < half; ++k) {
        int r = (int) bitrev((unsigned) k, bits);
        /* the reference rounds the windowed sample to float before the FFT (render.c:794) */
        re[r] = (double) (float) ((double) in[2 * k] * wn[2 * k]);
        im[r] = (double) (float) ((double) in[2 * k + 1] * wn[2 * k + 1]);
    }
    for (int span = 1; span < half; span <<= 1) {
        for (int m = 0; m < span; ++m) {
            double ang = -M_PI * (double) m / (double) span;
            double wr = cos(ang), wi = sin(ang);
            for (int k = m; k < half; k += 2 * span) {
                int j = k + span;
                double tr = wr * re[j] - wi * im[j], ti = wr * im[j] + wi * re[j];
                re[j] = re[k] - tr; im[j] = im[k] - ti;
                re[k] += tr; im[k] += ti;
            }
        }
    }
    for (int k = 0; k < half; ++k) { out[2 * k] = fabs(re[k]); out[2 * k + 1] = fabs(im[k]); }
    for (int i = 0; i < n; ++i) {
        double v = log(out[i] + 1) / 3;
        double ramp = ((double) i / (double) n) * (double) p->fft_scale + (1.0 - (double) p->fft_cutoff);
        out[i] = v * (ramp > 1.0 ? ramp : 1.0);
    }
    free(re); free(im); free(wn);
}

/* ------------------------------------------------------------------------------------ */
/* K5: util/smooth_pass.frag + util/smooth.glsl:23-64                                    */

static inline float scale_audio(const orc_params* p, float idx) {      /* smooth.glsl:13-15 */
    return -G_LOG((-(p->sample_range) * idx) + 1.0f) / p->sample_scale;
}
static inline float round_formula(const orc_params* p, float x) {      /* common.glsl:17-21 */
    switch (p->round_formula) {
        case 1:  return x;
        case 2:  return sqrtf(1.0f - ((x - 1.0f) * (x - 1.0f)));
        default: return (0.5f * G_SIN((G_PI * x) - (G_PI / 2.0f))) + 0.5f;
    }
}
static inline float fetch16(const uint16_t* tex, int n, int i) {
    return (i < 0 || i >= n) ? 0.0f : from16(tex[i]);
}
static float smooth_audio_raw(const orc_params* p, const uint16_t* tex, int n, float idx) {
    float fn = (float) n;
    float smin = scale_audio(p, g_clamp(idx - p->smooth_factor, 0.0f, 1.0f)) * fn;
    float smax = scale_audio(p, g_clamp(idx + p->smooth_factor, 0.0f, 1.0f)) * fn;
    float m = ((smax - smin) / 2.0f), s, w;
    float rm = smin + m;
    if (p->sample_mode == 0) {
        float avg = 0.0f, weight = 0.0f;
        for (s = smin; s <= smax; s += 1.0f) {
            w = round_formula(p, g_clamp((m - fabsf(rm - s)) / m, 0.0f, 1.0f));
            weight += w;
            avg += fetch16(tex, n, (int) g_round(s)) * w;
        }
        avg /= weight;
        return avg;
    } else if (p->sample_mode == 2) {
        float vmax = 0.0f, avg = 0.0f, weight = 0.0f, v;
        for (s = smin; s < smax; s += 1.0f) {
            w = round_formula(p, g_clamp((m - fabsf(rm - s)) / m, 0.0f, 1.0f));
            weight += w;
            v = fetch16(tex, n, (int) g_round(s)) * w;
            avg += v;
            if (vmax < v) vmax = v;
        }
        return (vmax * (1.0f - p->hybrid_weight)) + ((avg / weight) * p->hybrid_weight);
    } else {
        float vmax = 0.0f;
        for (s = smin; s < smax; s += 1.0f) {
            w = fetch16(tex, n, (int) g_round(s)) * round_formula(p, g_clamp((m - fabsf(rm - s)) / m, 0.0f, 1.0f));
            if (vmax < w) vmax = w;
        }
        return vmax;
    }
}

void orc_smooth_pass(const orc_params* p, const uint16_t* in, uint16_t* out) {
    /* smooth_pass.frag:14-16: viewport n x 1, gl_FragCoord.x = x + 0.5, uniform w = n */
    const int n = p->n;
    for (int x = 0; x < n; ++x) {
        float v = smooth_audio_raw(p, in, n, ((float) x + 0.5f) / (float) n);
        out[x] = unorm16(v);
    }
}

/* the sampler the module shaders use: smooth_audio() of smooth.glsl with
 * _PRE_SMOOTHED_AUDIO = smooth_pass (render.c:292) — as of the moment the stage-1 header was built: shaderload forms the
 * EBIND list before the shader's own includes run their requests (render.c:284-293 vs :312), so a setsmoothpass inside
 * smooth_parameters.glsl changes what the K5 pass does but not what the module's first shader believes */
static inline float smooth_audio(const orc_params* p, const uint16_t* tex, float idx) {
    const int believed = p->shader_pre_smoothed ? (p->shader_pre_smoothed == 1) : p->smooth_pass;
    if (believed) return fetch16(tex, p->n, (int) g_round(idx * (float) p->n));
    return smooth_audio_raw(p, tex, p->n, idx);
}
static inline float smooth_audio_adj(const orc_params* p, const uint16_t* tex, float idx, float pixel) {
    float al = smooth_audio(p, tex, g_max(idx - pixel, 0.0f)),
          am = smooth_audio(p, tex, idx),
          ar = smooth_audio(p, tex, g_min(idx + pixel, 1.0f));
    return (al + am + ar) / 3.0f;
}

/* ------------------------------------------------------------------------------------ */
/* per-channel update: render.c:2113-2309 (handle_audio)                                  */

struct orc_chan {
    int n, frames;
    /* pipeline A */
    float* applied;        /* transform_gravity state, render.c:725-726 */
    float* ring_f;         /* transform_average ring, oldest slot first, render.c:747-751 */
    /* pipeline B */
    uint16_t* gr_store;    /* render.c:2197 */
    uint16_t* ring_u;      /* gr->out[], render.c:2232-2242 */
    int out_idx;
    float* tmp; uint16_t* tmp_u;
};

orc_chan* orc_chan_new(const orc_params* p) {
    orc_chan* c = calloc(1, sizeof(*c));
    c->n = p->n; c->frames = p->avg_frames;
    c->applied  = calloc((size_t) p->n, sizeof(float));
    c->ring_f   = calloc((size_t) p->n * (size_t) p->avg_frames, sizeof(float));
    c->gr_store = calloc((size_t) p->n, sizeof(uint16_t));
    c->ring_u   = calloc((size_t) p->n * (size_t) p->avg_frames, sizeof(uint16_t));
    c->tmp      = calloc((size_t) p->n, sizeof(float));
    c->tmp_u    = calloc((size_t) p->n, sizeof(uint16_t));
    return c;
}
void orc_chan_free(orc_chan* c) {
    free(c->applied); free(c->ring_f); free(c->gr_store); free(c->ring_u); free(c->tmp); free(c->tmp_u); free(c);
}

static void gravity_a(orc_chan* c, const orc_params* p, float* b) {      /* render.c:720-736 */
    float g = p->gravity_step * (1.0f / p->ur);
    for (int t = 0; t < c->n; ++t) {
        if (b[t] >= c->applied[t]) c->applied[t] = b[t] - g;
        else c->applied[t] -= g;
        b[t] = c->applied[t];
    }
}
static void average_a(orc_chan* c, const orc_params* p, float* b) {      /* render.c:738-771 */
    const int n = c->n, F = c->frames;
    memmove(c->ring_f, c->ring_f + n, sizeof(float) * (size_t) n * (size_t) (F - 1));
    memcpy(c->ring_f + (size_t) n * (size_t) (F - 1), b, sizeof(float) * (size_t) n);
    for (int t = 0; t < n; ++t) {
        float v = 0.0f;
        for (int f = 0; f < F; ++f) {
            if (p->avg_window) {
                /* window_frame(f, d->avg_frames - 1) -> cos(TWOPI*f/F - 1), double (render.c:661,766) */
                double w = 0.6 - (0.4 * cos(6.28318530718 * (double) f / (double) F - 1));
                v = (float) ((double) v + w * (double) c->ring_f[(size_t) f * n + t]);
            } else v += c->ring_f[(size_t) f * n + t];
        }
        b[t] = v / (float) F;
    }
}

void orc_chan_update(orc_chan* c, const orc_params* p, const float* pcm, int is_fft,
                     float* spec_f32, uint16_t* tex_u16) {
    const int n = c->n, F = c->frames;
    float* b = c->tmp;
    memcpy(b, pcm, sizeof(float) * (size_t) n);
    uint16_t* tex = c->tmp_u;
    if (!is_fft) {
        /* wave: transforms "window" (no-op, render.c:850) + "wrange" (render.c:773-781) */
        for (int t = 0; t < n; ++t) { b[t] += 1.0f; b[t] /= 2.0f; }
        for (int t = 0; t < n; ++t) tex[t] = unorm16(b[t]);                  /* render.c:521-524 */
    } else if (!p->accel_fft) {
        if (is_fft != 2) orc_fft_f32(p, b);
        gravity_a(c, p, b); average_a(c, p, b);                               /* render.c:2149-2156 */
        for (int t = 0; t < n; ++t) tex[t] = unorm16(b[t]);
    } else {
        if (is_fft != 2) orc_fft_f32(p, b);                                   /* render.c:2177-2180 */
        float diff = p->gravity_step * (1.0f / p->ur);                        /* render.c:2224 */
        for (int t = 0; t < n; ++t) {
            uint16_t u = unorm16(b[t]);                                       /* upload, render.c:2185 */
            uint16_t g = c->gr_store[t] > u ? c->gr_store[t] : u;             /* K1 GL_MAX, render.c:2199-2211 */
            g = unorm16(from16(g) - diff);                                    /* K2 gravity_pass.frag:8 */
            c->gr_store[t] = g;
            tex[t] = g;
        }
        if (F > 1) {
            memcpy(c->ring_u + (size_t) c->out_idx * n, c->gr_store, sizeof(uint16_t) * (size_t) n);  /* K3 */
            int windowed = p->avg_window && F != 2;                           /* average_pass.frag:27-29 */
            for (int t = 0; t < n; ++t) {                                     /* K4 average_pass.frag:24-46 */
                float r = 0.0f;
                for (int i = 0; i < F; ++i) {
                    int fr = c->out_idx - i; if (fr < 0) fr += F;             /* render.c:2250-2255 */
                    float tx = from16(c->ring_u[(size_t) fr * n + t]);
                    if (windowed) {
                        /* window(I, _AVG_FRAMES - 1): cos(TWOPI * I / F - 1), GLSL float */
                        float w = 0.53836f - (0.46164f * cosf(G_TWOPI * (float) i / (float) F - 1.0f));
                        r += w * tx;
                    } else r += tx;
                }
                tex[t] = unorm16(r / (float) F);
            }
            if (++c->out_idx >= F) c->out_idx = 0;
        }
    }
    if (spec_f32) memcpy(spec_f32, b, sizeof(float) * (size_t) n);
    if (p->smooth_pass) orc_smooth_pass(p, tex, tex_u16);                      /* render.c:2276-2303 */
    else memcpy(tex_u16, tex, sizeof(uint16_t) * (size_t) n);
}

/* ------------------------------------------------------------------------------------ */
/* optional stages of rd_update (off in the shipped configuration)                        */

void orc_bufscale(const float* in, int n_in, int k, float* out) {           /* render.c:1765-1790 */
    int nsz = n_in / k;
    for (int t = 0; t < nsz; ++t) {
        float accum = 0.0F;
        for (int a = 0; a < k; ++a) accum += in[t * k + a];
        accum /= (float) k;
        out[t] = accum;
    }
}

void orc_transform_smooth(float* b, int sz, float smooth_distance, float smooth_ratio) {   /* render.c:694-718 */
    const double E = 2.7182818284590452353;                                  /* render.c:692 */
    size_t asz = (size_t) ceil(sz / smooth_ratio);
    for (int t = 0; t < (int) asz; ++t) {
        float db = log(t), avg = 0;                                          /* log(0) = -inf at t = 0 */
        float lo = db - smooth_distance; if (!(lo > 0)) lo = 0;              /* max(db - distance, 0) */
        int smin = (int) floor(powf(E, lo));
        int smax = (int) ceil(powf(E, db + smooth_distance));
        if (smax > sz - 1) smax = sz - 1;
        int count = 0;
        for (int s = smin; s <= smax; ++s)
            if (b[s]) { avg += b[s]; count++; }
        avg /= count;                                                        /* 0/0 = NaN when nothing was summed */
        b[t] = avg;
    }
}

void orc_interp(const float* s, const float* e, int n, float ur, float fr, int kcounter, float* out) {
    float uratio = ur / fr;                                                  /* render.c:1761 */
    for (int t = 0; t < n; ++t) {
        float mod = uratio * kcounter;                                       /* render.c:1804 */
        if (mod > 1.0F) mod = 1.0F;
        out[t] = s[t] + ((e[t] - s[t]) * mod);
    }
}

struct orc_stream {
    orc_params p;            /* effective: n = setbufsize / bufscale, accel_fft as the bind ends up using it */
    orc_ext x;
    int n_in, is_fft, post_chain, interp_on, kcounter;
    orc_chan* ch[2];
    float* key[2][2];        /* [channel][start, end] keyframes, render.c:1683-1689 */
    float* last[2];          /* post-transform buffer of the last modified update (what lb / rb hold) */
    uint16_t* tex[2];        /* texture of the previous frame (kept on unmodified frames) */
    float* scaled; float* work;
};

orc_stream* orc_stream_new(const orc_params* p, const orc_ext* x) {
    orc_stream* s = calloc(1, sizeof(*s));
    s->p = *p; s->x = *x;
    if (s->x.bufscale < 1) s->x.bufscale = 1;
    s->n_in = p->n; s->p.n = p->n / s->x.bufscale;
    s->is_fft = p->module != ORC_MOD_WAVE;
    /* a transform after "fft" makes the bind fall back to the CPU chain (render.c:2143-2154) */
    if (s->x.transform_smooth == 1 && s->is_fft) s->p.accel_fft = 0;        /* (2 = "smooth" BEFORE "fft": applied to the PCM, the fft chain stays where it is) */
    /* interpolation is forced off when the fft chain is pushed to the GPU passes (render.c:2161-2168)
       and when the update rate is close to the frame rate (render.c:1761-1763) */
    float fr = s->x.fr > 0 ? s->x.fr : p->ur;
    s->x.fr = fr;
    s->interp_on = s->x.interpolate && !(s->p.accel_fft && s->is_fft) && (p->ur / fr) <= 0.9F;
    s->post_chain = s->x.transform_smooth == 1 || s->interp_on;
    size_t n = (size_t) s->p.n;
    for (int c = 0; c < 2; ++c) {
        s->ch[c] = orc_chan_new(&s->p);
        s->key[c][0] = calloc(n, sizeof(float)); s->key[c][1] = calloc(n, sizeof(float));
        s->last[c] = calloc(n, sizeof(float)); s->tex[c] = calloc(n, sizeof(uint16_t));
    }
    s->scaled = calloc((size_t) s->n_in, sizeof(float)); s->work = calloc(n, sizeof(float));
    return s;
}
void orc_stream_free(orc_stream* s) {
    for (int c = 0; c < 2; ++c) {
        orc_chan_free(s->ch[c]); free(s->key[c][0]); free(s->key[c][1]); free(s->last[c]); free(s->tex[c]);
    }
    free(s->scaled); free(s->work); free(s);
}
int orc_stream_n(const orc_stream* s) { return s->p.n; }

void orc_stream_update(orc_stream* s, const float* lb, const float* rb, int modified,
                       float* spec_l, float* spec_r, uint16_t* tex_l, uint16_t* tex_r) {
    const int n = s->p.n;
    const float* in[2] = { lb, rb };
    float* spec[2] = { spec_l, spec_r };
    uint16_t* texo[2] = { tex_l, tex_r };
    uint16_t* pre = malloc(sizeof(uint16_t) * (size_t) n);
    for (int c = 0; c < 2; ++c) {
        /* keyframe lerp first, from the keyframes of the PREVIOUS updates (render.c:1792-1809) */
        if (s->interp_on)
            orc_interp(s->key[c][0], s->key[c][1], n, s->p.ur, s->x.fr, s->kcounter, s->work);
        if (modified) {
            const float* src = in[c];
            if (s->x.bufscale > 1) { orc_bufscale(in[c], s->n_in, s->x.bufscale, s->scaled); src = s->scaled; }
            if (s->x.transform_smooth == 2) {
                /* `#request transform <u> "smooth"` listed BEFORE "fft" (render.c:1218-1286): handle_audio applies it on the
                 * CPU to the (scaled) PCM, then meets "fft" and carries on as usual — GPU passes under setaccelfft
                 * (render.c:2131-2156) */
                if (src != s->scaled) { memcpy(s->scaled, src, sizeof(float) * (size_t) n); src = s->scaled; }
                orc_transform_smooth(s->scaled, n, s->x.smooth_distance, s->x.smooth_ratio);
            }
            if (!s->post_chain) {
                orc_chan_update(s->ch[c], &s->p, src, s->is_fft, s->last[c], s->tex[c]);
            } else {
                orc_params q = s->p; q.smooth_pass = 0;                     /* chain only; upload + K5 below */
                orc_chan_update(s->ch[c], &q, src, s->is_fft, s->last[c], pre);
                if (s->x.transform_smooth == 1)
                    orc_transform_smooth(s->last[c], n, s->x.smooth_distance, s->x.smooth_ratio);
            }
        }
        if (s->post_chain && (modified || s->interp_on)) {
            const float* up = s->interp_on ? s->work : s->last[c];          /* render.c:2185 */
            for (int t = 0; t < n; ++t) pre[t] = unorm16(up[t]);
            if (s->p.smooth_pass) orc_smooth_pass(&s->p, pre, s->tex[c]);
            else memcpy(s->tex[c], pre, sizeof(uint16_t) * (size_t) n);
        }
        if (s->interp_on && modified) {                                     /* render.c:2347-2353 */
            memcpy(s->key[c][0], s->key[c][1], sizeof(float) * (size_t) n);
            memcpy(s->key[c][1], s->last[c], sizeof(float) * (size_t) n);
        }
        if (spec[c]) memcpy(spec[c], s->last[c], sizeof(float) * (size_t) n);
        if (texo[c]) memcpy(texo[c], s->tex[c], sizeof(uint16_t) * (size_t) n);
    }
    free(pre);
    s->kcounter = modified ? 0 : s->kcounter + 1;                            /* render.c:2380-2383 */
}

/* ------------------------------------------------------------------------------------ */
/* FIFO ingest — fifo.c:89-110                                                           */
void orc_fifo_ingest(float* rl, float* rr, int n, const int16_t* in, int frames, int channels) {
    memmove(rl, rl + frames, sizeof(float) * (size_t) (n - frames));
    memmove(rr, rr + frames, sizeof(float) * (size_t) (n - frames));
    for (int q = 0; q < frames; ++q) {
        int idx = n - frames + q;
        if (channels == 1) {
            float s = (float) ((in[2 * q] + in[2 * q + 1]) / 2) / (float) 65535;
            rl[idx] = s; rr[idx] = s;
        } else {
            rl[idx] = (float) in[2 * q] / (float) 65535;
            rr[idx] = (float) in[2 * q + 1] / (float) 65535;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* module stage 1 shaders                                                                */

typedef struct { const orc_params* p; const uint16_t* l; const uint16_t* r; } rctx;

static vec4 bars_px(const rctx* c, int x, int y) {                       /* bars/1.frag:36-135 */
    const orc_params* p = c->p;
    float fx = (float) x + 0.5f, fy = (float) y + 0.5f;
    int   aw = p->bars_mirror_yx ? p->h : p->w, ah = p->bars_mirror_yx ? p->w : p->h;
    float ax = p->bars_mirror_yx ? fy : fx,  ay = p->bars_mirror_yx ? fx : fy;
    float dx;
    if (p->channels == 2) dx = ax - (float) (aw / 2);
    else dx = p->bars_invert ? (float) aw - ax : ax;
    float d = p->bars_flip ? (float) ah - ay : ay;
    float section = p->bars_width + p->bars_gap;
    float center = section / 2.0f;
    float m = fabsf(g_mod(dx, section));
    float md = m - center;
    float nbars = floorf(((float) aw * 0.5f) / section) * 2.0f;
    float hi = ceilf(p->bars_width / 2.0f), lo = -floorf(p->bars_width / 2.0f);
    if (md < hi && md >= lo) {
        float s = dx / section;
        float pp = (g_sign(s) == 1.0f ? ceilf(s) : floorf(s));
        if (p->channels == 2) pp /= (nbars / 2.0f); else pp /= nbars;
        pp += g_sign(pp) * ((0.5f + center) / (float) aw);
        if (pp > 1.0f || pp < -1.0f) return v4(0, 0, 0, 0);
        float v;
        const uint16_t* tex;
        if (pp > 0.0f) {
            if (p->bars_direction == 1) pp = 1.0f - pp;
            tex = (p->channels == 1 || p->bars_invert > 0) ? c->l : c->r;
        } else {
            pp = fabsf(pp);
            if (p->bars_direction == 1) pp = 1.0f - pp;
            tex = (p->channels == 1) ? c->l : (p->bars_invert > 0 ? c->r : c->l);
        }
        v = smooth_audio(p, tex, pp);
        v *= p->bars_amplify;
        vec4 col = eval_color(&p->bars_color, d);
        vec4 outl = p->bars_outline_mode == 0 ? v4(col.r * 1.5f, col.g * 1.5f, col.b * 1.5f, col.a) : v4a(p->bars_outline);
        if (d < v - p->bars_outline_width) {
            if (p->bars_outline_width > 0) {
                if (md < hi - p->bars_outline_width && md >= lo + p->bars_outline_width) return col;
                return outl;
            }
            return col;
        }
        if (p->bars_outline_width > 0 && d <= v) return outl;
    }
    return v4(0, 0, 0, 0);
}

static inline vec4 apply_frag(vec4 f, vec4 c) {                         /* radial/1.frag:35 */
    float k = 1.0f - g_clamp(f.a, 0.0f, 1.0f);
    return v4(f.r * f.a + c.r * k, f.g * f.a + c.g * k, f.b * f.a + c.b * k, g_max(c.a, f.a));
}

static vec4 radial_px(const rctx* c, int x, int y) {                     /* radial/1.frag:32-116, _USE_ALPHA > 0 */
    const orc_params* p = c->p;
    vec4 frag = v4(0, 0, 0, 0);
    float dx = ((float) x + 0.5f) - (float) (p->w / 2) + p->radial_off_x,
          dy = ((float) y + 0.5f) - (float) (p->h / 2) + p->radial_off_y;
    float theta = G_ATAN2(dy, dx);
    float d = sqrtf((dx * dx) + (dy * dy));
    float R = p->radial_radius, hl = p->radial_line / 2.0f;
    if (d > R - hl && d < R + hl) {
        frag = apply_frag(frag, v4a(p->radial_outline));
        frag.a *= g_clamp((p->radial_line_half - fabsf(R - d)) * p->radial_c_alias, 0.0f, 1.0f);
    }
    if (d > R) {
        const float section = (G_TWOPI / (float) p->radial_nbars);
        const float center = ((G_TWOPI / (float) p->radial_nbars) / 2.0f);
        float m = g_mod(theta, section);
        float ym = d * G_SIN(center - m);
        /* `BAR_WIDTH / 2` (radial/1.frag:62,79,88) is an INTEGER division when the macro is an integer literal */
        const float bw2 = p->radial_bar_width_int ? (float) ((int) p->radial_bar_width / 2) : p->radial_bar_width / 2.0f;
        const float ow = p->radial_bar_outline_width;                    /* BAR_OUTLINE_WIDTH (deprecated, radial.glsl:33-36) */
        if (fabsf(ym) < bw2) {
            float idx = theta + p->radial_rotate;
            float dir = g_mod(fabsf(idx), G_TWOPI);
            if (dir > G_PI) idx = -g_sign(idx) * (G_TWOPI - dir);
            if (p->radial_invert == 0) idx = -idx;
            float pos = (float) (int) (fabsf(idx) / section) / (float) (p->radial_nbars / 2);
            float v = smooth_audio(p, idx > 0.0f ? c->l : c->r, pos);
            v *= p->radial_amplify;
            d -= R;
            if (d <= v - ow) {                                           /* radial/1.frag:85-99 */
                vec4 r;
                if (!(ow > 0.0f) || fabsf(ym) < bw2 - ow) r = eval_color(&p->radial_color, d);
                else r = v4a(p->radial_bar_outline);
                r.a *= ((bw2 - fabsf(ym)) * p->radial_bar_alias);
                return apply_frag(frag, r);
            }
            if (ow > 0.0f && d <= v) {                                   /* radial/1.frag:100-110 */
                vec4 r = v4a(p->radial_bar_outline);
                r.a *= ((bw2 - fabsf(ym)) * p->radial_bar_alias);
                return apply_frag(frag, r);
            }
        }
    }
    return apply_frag(frag, v4(0, 0, 0, 0));
}

static float circle_apply_smooth(const rctx* c, float theta) {          /* circle/1.frag:34-49 */
    const orc_params* p = c->p;
    float idx = theta + p->circle_rotate;
    float dir = g_mod(fabsf(idx), G_TWOPI);
    if (dir > G_PI) idx = -g_sign(idx) * (G_TWOPI - dir);
    if (p->circle_invert > 0) idx = -idx;
    float pos = fabsf(idx) / (G_PI + 0.001f);
    float v = smooth_audio(p, idx > 0.0f ? c->l : c->r, pos);
    v *= p->circle_amplify;
    return v;
}
static vec4 circle_px(const rctx* c, int x, int y) {                     /* circle/1.frag:51-84, pixel_center_integer */
    const orc_params* p = c->p;
    float dx = (float) x - (float) (p->w / 2), dy = (float) y - (float) (p->h / 2);
    float theta = G_ATAN2(dy, dx);
    float d = sqrtf((dx * dx) + (dy * dy));
    float adv = (1.0f / d) * (p->circle_line * 0.5f);
    float adj0 = theta + adv, adj1 = theta - adv;
    d -= p->circle_radius;
    float hl = p->circle_line / 2.0f;
    if (d >= -hl) {
        float v = circle_apply_smooth(c, theta);
        adj0 = circle_apply_smooth(c, adj0) - v;
        adj1 = circle_apply_smooth(c, adj1) - v;
        float dmax = g_max(adj0, adj1), dmin = g_min(adj0, adj1);
        d -= v;
        int in = p->circle_fill ? (d < hl) : ((d > -hl && d < hl) || (d <= dmax && d >= dmin));
        if (in) return v4a(p->circle_outline);
    }
    return v4(0, 0, 0, 0);
}

static float graph_height(const rctx* c, int x) {                        /* graph/1.frag:87-105,124-132 */
    const orc_params* p = c->p;
    float fx = (float) x, W = (float) p->w;
    float half_w = (float) (p->w / 2);
    float pixel = 1.0f / W;
    const uint16_t* tex; float idx;
    if (fx < half_w) { tex = c->l; idx = p->graph_direction < 0 ? fx : (half_w - fx); }
    else             { tex = c->r; idx = p->graph_direction < 0 ? (-fx + W) : (fx - half_w); }
    float s = smooth_audio_adj(p, tex, idx / half_w, pixel);
    s *= p->graph_vscale;
    float fact = g_clamp((fabsf((float) (p->w / 2) - fx) / W) * 48.0f, 0.0f, 1.0f);
    if (p->graph_join_channels) {                                        /* graph/1.frag:93-96,126 */
        float middle = (p->graph_vscale * (smooth_audio_adj(p, c->l, 1.0f, pixel) + smooth_audio_adj(p, c->r, 0.0f, pixel))) / 2.0f;
        fact = (-2.0f * G_POW(fact, 3)) + (3.0f * G_POW(fact, 2));
        s = (fact * s) + ((1.0f - fact) * middle);
    } else s *= fact;
    s *= g_clamp((g_min(fx, W - fx) / W) * 48.0f, 0.0f, 1.0f);
    return s;
}
static vec4 graph_px(const rctx* c, int x, int y) {                      /* graph/1.frag:107-122 */
    const orc_params* p = c->p;
    float s = graph_height(c, x);
    float d = p->graph_invert > 0 ? (float) p->h - (float) y : (float) y;
    if (d + 1.5f <= s) return eval_color(&p->graph_color, d);
    return v4(0, 0, 0, 0);
}

static inline float wave_tex(const rctx* c, float coord) {               /* texture(): NEAREST, REPEAT */
    const int n = c->p->n;
    float u = coord * (float) n;
    int i = (int) floorf(u);
    i %= n; if (i < 0) i += n;
    return from16(c->l[i]);
}
static vec4 wave_px(const rctx* c, int x, int y) {                       /* wave/1.frag:17-39, pixel_center_integer */
    const orc_params* p = c->p;
    float fx = (float) x, fy = (float) y, W = (float) p->w, H = (float) p->h;
#define WAVE_INDEX(off) (((wave_tex(c, (fx + (off)) / W) - 0.5f) * p->wave_amplify) + 0.5f)
    float os = WAVE_INDEX(0.0f), adj0 = WAVE_INDEX(-1.0f), adj1 = WAVE_INDEX(1.0f);
#undef WAVE_INDEX
    float s0 = adj0 - os, s1 = adj1 - os;
    float dmax = g_max(s0, s1), dmin = g_min(s0, s1);
    float s = (os + (H * 0.5f) - 0.5f);
    float diff = fy - s;
    if (fabsf(diff) < g_clamp(fabsf(s - (H * 0.5f)) * 6.0f, p->wave_min_thickness, p->wave_max_thickness)
        || (diff <= dmax && diff >= dmin)) {
        float k = (fabsf((H * 0.5f) - s) * 0.02f);
        return v4(p->wave_base_color[0] + k, p->wave_base_color[1] + k, p->wave_base_color[2] + k, p->wave_base_color[3] + k);
    }
    return v4(0, 0, 0, 0);
}

/* What a stage's `fragment` becomes in its RGBA8 target.  setopacity "native": blending is off (render.c:1467-1470), plain
 * unorm8 conversion.  Any other opacity: GL_BLEND with glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) over the target
 * glClear'd to the `setbg` colour (render.c:1700, 2028) — fixed-function blending in unorm8 fixed point (below). */
static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static uint32_t store8(const orc_params* p, vec4 s) {
    if (p->premultiply_alpha) return pack8(s);
    s = v4(clamp01(s.r), clamp01(s.g), clamp01(s.b), clamp01(s.a));
    /* Blending happens in the target's own 8-bit normalised fixed point (llvmpipe's unorm8 blend; pinned by the llvmpipe
     * goldens, where every non-native-opacity frame is reproduced bit for bit): the fragment is converted to unorm8 FIRST,
     * then  C = mul_norm(Cs, As) + mul_norm(Cd, 255 - As)  per channel (alpha included), saturating;
     * mul_norm(a, b) = (t + (t >> 8)) >> 8 with t = a * b + 128, i.e. a * b / 255 rounded. */
    {
        const uint32_t S = pack8(s), D = pack8(v4a(p->clear_color));
        const uint32_t a = S >> 24;
        uint32_t out = 0;
        for (int c = 0; c < 4; ++c) {
            const uint32_t v1 = (S >> (8 * c)) & 255u, v0 = (D >> (8 * c)) & 255u;
            uint32_t t1 = v1 * a + 128u;         t1 = (t1 + (t1 >> 8)) >> 8;
            uint32_t t0 = v0 * (255u - a) + 128u; t0 = (t0 + (t0 >> 8)) >> 8;
            uint32_t r = t1 + t0; if (r > 255u) r = 255u;
            out |= r << (8 * c);
        }
        return out;
    }
}

static uint32_t stage1(const rctx* c, int x, int y) {
    switch (c->p->module) {
        case ORC_MOD_BARS:   return store8(c->p, bars_px(c, x, y));
        case ORC_MOD_RADIAL: return store8(c->p, radial_px(c, x, y));
        case ORC_MOD_CIRCLE: return store8(c->p, circle_px(c, x, y));
        case ORC_MOD_GRAPH:  return store8(c->p, graph_px(c, x, y));
        case ORC_MOD_WAVE:   return store8(c->p, wave_px(c, x, y));
        default:             return store8(c->p, v4(1.0f, 0.0f, 0.0f, (float) 1 / (float) 3));   /* test/1.frag:32 */
    }
}

/* ------------------------------------------------------------------------------------ */
/* post stages on RGBA8 surfaces                                                         */

typedef struct { const uint32_t* px; int w, h, y0, rows; } surf;   /* rows [y0, y0+rows) stored */
static inline vec4 sfetch(const surf* s, int x, int y) {
    if (x < 0 || x >= s->w || y < 0 || y >= s->h) return v4(0, 0, 0, 0);
    return unpack8(s->px[(size_t) (y - s->y0) * s->w + x]);
}
/* the 8-tap neighbour mean written out in circle/2.frag:18-26, graph/2.frag:21-29,
 * wave/2.frag:18-26 — taps a3 and a7 repeat a0 and a4 in the source.
 * The taps are addressed as ivec2(gl_FragCoord.x - 1, gl_FragCoord.y - 1).  wave/2.frag declares
 * layout(pixel_center_integer): x - 1 is exactly -1 at the left edge, outside the surface.  circle/2.frag and
 * graph/2.frag use the default half-integer gl_FragCoord: x + 0.5 - 1 = -0.5 at x = 0, and float -> int conversion
 * drops the fraction (GLSL 3.30 5.4.1), i.e. 0 — the "x - 1" / "y - 1" taps of column 0 / row 0 read column 0 / row 0.
 * (Found by running the reference's shader text through oracle/glsl_interp.py.) */
static inline vec4 neigh_avg(const surf* s, int x, int y, int half_integer_coords) {
    int xm = x - 1, ym = y - 1;
    if (half_integer_coords) { if (xm < 0) xm = 0; if (ym < 0) ym = 0; }
    vec4 a0 = sfetch(s, x + 1, y), a1 = sfetch(s, x + 1, y + 1), a2 = sfetch(s, x, y + 1), a3 = sfetch(s, x + 1, y),
         a4 = sfetch(s, xm, y), a5 = sfetch(s, xm, ym), a6 = sfetch(s, x, ym), a7 = sfetch(s, xm, y);
    vec4 r;
    r.r = (a0.r + a1.r + a2.r + a3.r + a4.r + a5.r + a6.r + a7.r) / 8.0f;
    r.g = (a0.g + a1.g + a2.g + a3.g + a4.g + a5.g + a6.g + a7.g) / 8.0f;
    r.b = (a0.b + a1.b + a2.b + a3.b + a4.b + a5.b + a6.b + a7.b) / 8.0f;
    r.a = (a0.a + a1.a + a2.a + a3.a + a4.a + a5.a + a6.a + a7.a) / 8.0f;
    return r;
}
static inline vec4 premultiply(vec4 f) { return v4(f.r * f.a, f.g * f.a, f.b * f.a, f.a); }  /* premultiply.frag:12-15 */

/* graph/3.frag:19-105 (ANTI_ALIAS 1), default half-integer gl_FragCoord.  S = the previous stage's surface.  graph/4.frag
 * (premultiply) tests `#if ANTI_ALIAS == 0` WITHOUT including graph.glsl: the macro is undefined there, evaluates to 0 and
 * the stage is always disabled — stage 3 is the final one, in native mode too. */
static float aa_up(const orc_params* p, const surf* S, float x, float oy) {      /* get_col_height_up, :21-44 */
    float y = oy;
    if (p->graph_invert > 0) {
        while (y >= 0.0f) { if (sfetch(S, (int) x, (int) y).a <= 0.0f) { y += 1.0f; break; } y -= 1.0f; }
    } else {
        while (y < (float) S->h) { if (sfetch(S, (int) x, (int) y).a <= 0.0f) { y -= 1.0f; break; } y += 1.0f; }
    }
    return y;
}
static float aa_down(const orc_params* p, const surf* S, float x, float oy) {    /* get_col_height_down, :48-69 */
    float y = oy;
    if (p->graph_invert > 0) {
        while (y < (float) S->h) { if (sfetch(S, (int) x, (int) y).a > 0.0f) break; y += 1.0f; }
    } else {
        while (y >= 0.0f) { if (sfetch(S, (int) x, (int) y).a > 0.0f) break; y -= 1.0f; }
    }
    return y;
}
static void graph_anti_alias(const orc_params* p, const surf* S, uint32_t* dst, int y0, int y1) {
    for (int yi = y0; yi < y1; ++yi) {
        for (int xi = 0; xi < S->w; ++xi) {
            const float X = (float) xi + 0.5f, Y = (float) yi + 0.5f;
            vec4 f = sfetch(S, (int) X, (int) Y);
            if (f.a <= 0.0f) {
                int left_done = 0;
                float h2 = 0.0f, a_fact = 0.0f;
                if (sfetch(S, (int) (X - 1.0f), (int) Y).a > 0.0f) {
                    float h1 = aa_up(p, S, X - 1.0f, Y);
                    h2 = aa_down(p, S, X, Y);
                    f = sfetch(S, (int) X, (int) h2);
                    a_fact = g_clamp(fabsf((h1 - Y) / (h2 - h1)), 0.0f, 1.0f);
                    left_done = 1;
                }
                if (sfetch(S, (int) (X + 1.0f), (int) Y).a > 0.0f) {
                    if (!left_done) { h2 = aa_down(p, S, X, Y); f = sfetch(S, (int) X, (int) h2); }
                    float h3 = aa_up(p, S, X + 1.0f, Y);
                    a_fact = g_max(a_fact, g_clamp(fabsf((h3 - Y) / (h2 - h3)), 0.0f, 1.0f));
                }
                f.a *= a_fact;
            }
            dst[(size_t) yi * S->w + xi] = store8(p, f);
        }
    }
}

void orc_raster_rows(const orc_params* p, const uint16_t* tl, const uint16_t* tr, uint8_t* out, int y0, int y1) {
    rctx c = { p, tl, tr };
    const int w = p->w, h = p->h;
    uint32_t* dst = (uint32_t*) out;
    /* which modules have a neighbourhood stage */
    int stencil = (p->module == ORC_MOD_CIRCLE && p->circle_smooth) ||
                  (p->module == ORC_MOD_GRAPH && (p->graph_draw_outline || p->graph_draw_highlight)) ||
                  (p->module == ORC_MOD_WAVE);
    /* graph ANTI_ALIAS 1 (graph/3.frag) walks whole columns of the previous stage: render the full surface, then the walk */
    const int aa = p->module == ORC_MOD_GRAPH && p->graph_anti_alias;
    const int out_y0 = y0, out_y1 = y1;
    uint32_t* full = NULL;
    if (aa) { full = malloc(sizeof(uint32_t) * (size_t) w * (size_t) h); dst = full; y0 = 0; y1 = h; }
    int sy0 = stencil ? (y0 > 0 ? y0 - 1 : 0) : y0, sy1 = stencil ? (y1 < h ? y1 + 1 : h) : y1;
    uint32_t* s1 = malloc(sizeof(uint32_t) * (size_t) w * (size_t) (sy1 - sy0));
    for (int y = sy0; y < sy1; ++y)
        for (int x = 0; x < w; ++x) s1[(size_t) (y - sy0) * w + x] = stage1(&c, x, y);
    surf S = { s1, w, h, sy0, sy1 - sy0 };
    for (int y = y0; y < y1; ++y) {
        for (int x = 0; x < w; ++x) {
            uint32_t px = s1[(size_t) (y - sy0) * w + x];
            switch (p->module) {
                case ORC_MOD_BARS: break;                                   /* bars/2.frag disabled (USE_ALPHA 0) */
                case ORC_MOD_RADIAL:
                    if (p->premultiply_alpha) px = pack8(premultiply(unpack8(px)));   /* radial/2.frag */
                    break;
                case ORC_MOD_CIRCLE: {
                    vec4 f = unpack8(px);
                    if (p->circle_smooth) {                                 /* circle/2.frag:14-32 */
                        vec4 avg = neigh_avg(&S, x, y, 1);
                        if (f.a == 0.0f) f = avg;
                    }
                    /* with C_SMOOTH 0 the stage is NOT disabled: it copies its input (circle/2.frag:12) — the identity
                     * natively, one more blend over the clear colour otherwise */
                    px = store8(p, f); f = unpack8(px);
                    if (p->premultiply_alpha) px = pack8(premultiply(f));   /* circle/3.frag */
                    break;
                }
                case ORC_MOD_GRAPH: {
                    if (p->graph_draw_outline || p->graph_draw_highlight) { /* graph/2.frag:19-44 */
                        vec4 f = unpack8(px);
                        vec4 avg = neigh_avg(&S, x, y, 1);
                        if (avg.a > 0.0f) {
                            if (f.a <= 0.0f) { if (p->graph_draw_outline) f = v4a(p->graph_outline); }
                            else if (avg.a < 1.0f) {
                                if (p->graph_draw_highlight) { float k = avg.a * 2.0f; f.r *= k; f.g *= k; f.b *= k; }
                            }
                        }
                        px = store8(p, f);
                    }
                    break;                                                  /* graph/3,4.frag disabled (ANTI_ALIAS 0) */
                }
                case ORC_MOD_WAVE: {                                        /* wave/2.frag:14-33 */
                    vec4 f = unpack8(px);
                    vec4 avg = neigh_avg(&S, x, y, 0);
                    if (avg.a > 0.0f) {
                        if (f.a <= 0.0f || x == 0 || x == w - 1) f = v4a(p->wave_outline);
                    }
                    px = store8(p, f);
                    break;
                }
                default: {                                                  /* test/2.frag, test/3.frag */
                    px = store8(p, unpack8(px));
                    if (p->premultiply_alpha) px = pack8(premultiply(unpack8(px)));
                    break;
                }
            }
            dst[(size_t) y * w + x] = px;
        }
    }
    free(s1);
    if (aa) {
        surf S2 = { full, w, h, 0, h };
        graph_anti_alias(p, &S2, (uint32_t*) out, out_y0, out_y1);
        free(full);
    }
}

void orc_raster(const orc_params* p, const uint16_t* tl, const uint16_t* tr, uint8_t* out) {
    orc_raster_rows(p, tl, tr, out, 0, p->h);
}
