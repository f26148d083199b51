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

/* ------------------------------------------------------------------------------------------------------------------
 * The reference's own GLSL source extension, glsl_ext.c (compiled into this library from where it lies): `#include`
 * with the ':' / '@' directory rules, `#request` parsing + typed argument conversion, `#expand`, `#rrggbb[aa]` colour
 * literals, `@name:default` pipe binds.  ref_ext_process() runs ext_process() on one file with a handler for every
 * request name of render.c:1033-1314 (same format strings) that only logs "name|arg|arg...\n" — so tests can pin
 *   - oracle/glsl_interp.py's restatement of the extension (the front end of every shader-derived golden frame) and
 *   - the product's config reader (csrc/config.cpp)
 * to what the reference really does with the same text.  parse errors call glava_abort (a fn-ptr, glava.h:17): it is
 * pointed at a longjmp for the duration of the call and reported as return value 1. */
#include <setjmp.h>
#include <stdarg.h>

static jmp_buf ref_ext_jmp;
static void ref_ext_abort(void) { longjmp(ref_ext_jmp, 1); }

static char*  ref_log_buf; static size_t ref_log_cap, ref_log_len;
static void ref_logf(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    if (ref_log_buf && ref_log_len < ref_log_cap) {
        int n = vsnprintf(ref_log_buf + ref_log_len, ref_log_cap - ref_log_len, fmt, ap);
        if (n > 0) ref_log_len += (size_t) n < ref_log_cap - ref_log_len ? (size_t) n : ref_log_cap - ref_log_len - 1;
    }
    va_end(ap);
}

static const struct { const char* name; const char* fmt; } ref_requests[] = {   /* render.c:1033-1314 */
    { "setopacity", "s" }, { "setmirror", "b" }, { "setfullscreencheck", "b" }, { "setbg", "s" }, { "settesteval", "s" },
    { "setbgf", "ffff" }, { "mod", "s" }, { "nativeonly", "b" }, { "setfloating", "b" }, { "setdecorated", "b" },
    { "setfocused", "b" }, { "setmaximized", "b" }, { "setversion", "ii" }, { "setgeometry", "iiii" },
    { "addxwinstate", "s" }, { "setsource", "s" }, { "setclickthrough", "b" }, { "setforcegeometry", "b" },
    { "setforceraised", "b" }, { "setxwintype", "s" }, { "setshaderversion", "i" }, { "setswap", "i" },
    { "setframerate", "i" }, { "setprintframes", "b" }, { "settitle", "s" }, { "setbufsize", "i" }, { "setbufscale", "i" },
    { "setsamplerate", "i" }, { "setsamplesize", "i" }, { "setaccelfft", "b" }, { "setavgframes", "i" },
    { "setavgwindow", "b" }, { "setgravitystep", "f" }, { "setsmoothpass", "b" }, { "setsmoothfactor", "f" },
    { "setsmooth", "f" }, { "setsmoothratio", "f" }, { "setinterpolate", "b" }, { "setfftscale", "f" },
    { "setfftcutoff", "f" }, { "timecycle", "f" }, { "transform", "ss" }, { "uniform", "ss" }, { NULL, NULL }
};

static void ref_request_logger(const char* name, void** args) {
    const char* fmt = "";
    for (size_t t = 0; ref_requests[t].name; ++t) if (!strcmp(ref_requests[t].name, name)) fmt = ref_requests[t].fmt;
    ref_logf("%s", name);
    for (size_t i = 0; fmt[i]; ++i) {
        switch (fmt[i]) {
            case 'i': ref_logf("|%d", *(int*) args[i]); break;
            case 'f': ref_logf("|%.9g", (double) *(float*) args[i]); break;
            case 'b': ref_logf("|%d", *(bool*) args[i] ? 1 : 0); break;
            default:  ref_logf("|%s", (const char*) args[i]); break;
        }
    }
    ref_logf("\n");
}

/* path: file to process; cd / cfd / dd: current, config (may be NULL) and defaults directories; binds: NULL-terminated
 * `--pipe` names (may be NULL); avg_frames: the `_AVG_FRAMES` #expand count.  out / reqlog: caller buffers.
 * returns 0 ok, 1 parse error (glava_abort was called), 2 cannot read the file, 3 output truncated */
int ref_ext_process(const char* path, const char* cd, const char* cfd, const char* dd, const char** binds, int avg_frames,
                    char* out, size_t out_cap, char* reqlog, size_t req_cap) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return 2;
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    char* src = malloc((size_t) sz + 1);
    if (fread(src, 1, (size_t) sz, fp) != (size_t) sz) { fclose(fp); free(src); return 2; }
    fclose(fp); src[sz] = '\0';

    struct request_handler handlers[sizeof(ref_requests) / sizeof(ref_requests[0])];
    size_t nh = 0;
    for (; ref_requests[nh].name; ++nh)
        handlers[nh] = (struct request_handler) { .name = ref_requests[nh].name, .fmt = ref_requests[nh].fmt, .handler = ref_request_logger };
    handlers[nh] = (struct request_handler) { .name = NULL };

    struct rd_bind bd[17]; size_t nb = 0;
    for (; binds && binds[nb] && nb < 16; ++nb) bd[nb] = (struct rd_bind) { .name = binds[nb], .stype = "vec4", .type = STDIN_TYPE_VEC4 };
    bd[nb] = (struct rd_bind) { .name = NULL };

    static int s_avg; s_avg = avg_frames;
    size_t avg_call(void) { return (size_t) s_avg; }
    struct glsl_ext_efunc efuncs[] = { { .name = "_AVG_FRAMES", .call = avg_call }, { .name = NULL } };

    struct glsl_ext ext = { .source = src, .source_len = (size_t) sz, .cd = cd, .cfd = cfd, .dd = dd, .handlers = handlers,
                            .processed = NULL, .p_len = 0, .binds = bd, .efuncs = efuncs };
    ref_log_buf = reqlog; ref_log_cap = req_cap; ref_log_len = 0;
    if (reqlog && req_cap) reqlog[0] = '\0';
    void (*saved)(void) = glava_abort;
    int rc = 0;
    glava_abort = ref_ext_abort;
    if (setjmp(ref_ext_jmp) == 0) {
        ext_process(&ext, path);
        if (ext.p_len + 1 > out_cap) rc = 3;
        else { memcpy(out, ext.processed, ext.p_len); out[ext.p_len] = '\0'; }
        ext_free(&ext);
    } else rc = 1;                                          /* parse_error -> glava_abort: buffers of that run are abandoned */
    glava_abort = saved;
    free(src);
    return rc;
}
