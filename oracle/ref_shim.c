/* TEST INFRASTRUCTURE — not product code.
 *
 * Shim that compiles the UNMODIFIED reference renderer translation unit where it
 * lies (/root/reference/glava/render.c, passed as -DGLAVA_REF_RENDER_C=...) and
 * re-exports its CPU transforms through a flat C ABI, so tests and the reference
 * arm of bench.py can call the reference's own code:
 *
 *   transform_fft      render.c:783-847
 *   transform_gravity  render.c:720-736
 *   transform_average  render.c:738-771
 *   transform_wrange   render.c:773-781
 *   transform_smooth   render.c:694-718
 *
 * The shim needs render.c's private `struct gl_data` / `struct gl_sampler_data`
 * (render.c:116-119,166-207), hence the #include of the .c file.  No reference
 * source is copied into this repository; the output (.so) goes to oracle/_ref/,
 * which is git-ignored.
 */
#include GLAVA_REF_RENDER_C

#include <stdint.h>

/* One per (stream, channel): persistent per-transform state the reference keeps in
 * gl->t_data[] (render.c:662-666 ALLOC_ONCE, render.c:2149-2156 call order). */
struct ref_chan {
    struct gl_data d;
    void* udata_gravity;
    void* udata_average;
};

void* ref_chan_new(float fft_scale, float fft_cutoff, float gravity_step, float ur,
                   int avg_frames, int avg_window) {
    struct ref_chan* c = calloc(1, sizeof(*c));
    c->d.fft_scale    = fft_scale;
    c->d.fft_cutoff   = fft_cutoff;
    c->d.gravity_step = gravity_step;
    c->d.ur           = ur;
    c->d.avg_frames   = (size_t) avg_frames;
    c->d.avg_window   = avg_window != 0;
    return c;
}

void ref_chan_free(void* p) {
    struct ref_chan* c = p;
    free(c->udata_gravity);
    free(c->udata_average);
    free(c);
}

void ref_chan_set_ur(void* p, float ur) { ((struct ref_chan*) p)->d.ur = ur; }

void ref_fft(void* p, float* buf, size_t sz) {
    struct ref_chan* c = p;
    struct gl_sampler_data s = { .buf = buf, .sz = sz };
    transform_fft(&c->d, NULL, &s);
}

void ref_gravity(void* p, float* buf, size_t sz) {
    struct ref_chan* c = p;
    struct gl_sampler_data s = { .buf = buf, .sz = sz };
    transform_gravity(&c->d, &c->udata_gravity, &s);
}

void ref_average(void* p, float* buf, size_t sz) {
    struct ref_chan* c = p;
    struct gl_sampler_data s = { .buf = buf, .sz = sz };
    transform_average(&c->d, &c->udata_average, &s);
}

void ref_wrange(float* buf, size_t sz) {
    struct gl_sampler_data s = { .buf = buf, .sz = sz };
    transform_wrange(NULL, NULL, &s);
}

/* transform_smooth (render.c:694-718): registered as "smooth", requested by no shipped module. */
void ref_smooth(float* buf, size_t sz, float smooth_distance, float smooth_ratio) {
    struct gl_data d;
    memset(&d, 0, sizeof(d));
    d.smooth_distance = smooth_distance;
    d.smooth_ratio    = smooth_ratio;
    struct gl_sampler_data s = { .buf = buf, .sz = sz };
    transform_smooth(&d, NULL, &s);
}

/* The CPU chain rd_update runs with `setaccelfft false` (render.c:2149-2156). */
void ref_update_a(void* p, float* buf, size_t sz) {
    ref_fft(p, buf, sz);
    ref_gravity(p, buf, sz);
    ref_average(p, buf, sz);
}

/* #rrggbb[aa] parsing used for colour literals (glsl_ext.c:88-122). */
int ref_parse_color(const char* str, float* rgba) {
    float* res[4] = { &rgba[0], &rgba[1], &rgba[2], &rgba[3] };
    return ext_parse_color(str, 2, res) ? 1 : 0;
}
