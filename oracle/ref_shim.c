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

/* ------------------------------------------------------------------------------------------------------------------
 * rd_new / rd_update FOR REAL, with a null OpenGL driver and a null window backend.
 *
 * Every gl* the renderer calls is a glad function pointer; here they point at stubs that hand out object ids, report
 * success, and record what update_1d_tex() uploads (render.c:521-524) — the float buffer rd_update hands to GL as the
 * audio texture of a bind, i.e. the result of the reference's own transform chain, buffer scaling and keyframe
 * interpolation as rd_update orchestrates them (render.c:1743-2417).  A window backend "null" is registered the way
 * glx_wcb / glfw_wcb register themselves.  rd_new then reads rc.glsl and the module through its own request handlers
 * (render.c:1033-1314), so the fields of `struct gl_data` afterwards are the reference's reading of a config.
 * What this cannot show: anything the GLSL passes compute (no shader runs; a "compiled" stage is never disabled).
 * ------------------------------------------------------------------------------------------------------------------ */
static GLuint ng_next_id = 1;
static GLuint ng_bound_1d = 0;
static void   ng_noop(void) {}
static GLenum ng_get_error(void) { return GL_NO_ERROR; }
static GLuint ng_create_shader(GLenum type) { (void) type; return ng_next_id++; }
static GLuint ng_create_program(void) { return ng_next_id++; }
static void   ng_gen(GLsizei n, GLuint* out) { for (GLsizei i = 0; i < n; ++i) out[i] = ng_next_id++; }
/* uniform locations: one id per distinct name, so that writes to the `--pipe` uniforms (`_IN_<name>`, render.c:2071-2100) can
 * be told apart and recorded */
#define NG_MAX_NAMES 128
static char* ng_names[NG_MAX_NAMES];
static int   ng_name_count = 0;
static GLint ng_uniform_location(GLuint p, const GLchar* name) {
    (void) p;
    for (int i = 0; i < ng_name_count; ++i) if (!strcmp(ng_names[i], name)) return i + 1;
    if (ng_name_count == NG_MAX_NAMES) return 0;
    ng_names[ng_name_count] = strdup(name);
    return ++ng_name_count;
}
#define NG_MAX_PIPE_WRITES 256
static struct { int loc, count, is_int; float v[4]; } ng_pipe_writes[NG_MAX_PIPE_WRITES];
static int ng_pipe_write_count = 0;
static void ng_record_uniform(GLint loc, int count, int is_int, float a, float b, float c, float d) {
    if (loc < 1 || loc > ng_name_count || strncmp(ng_names[loc - 1], "_IN_", 4) != 0 || ng_pipe_write_count == NG_MAX_PIPE_WRITES) return;
    ng_pipe_writes[ng_pipe_write_count].loc = loc; ng_pipe_writes[ng_pipe_write_count].count = count;
    ng_pipe_writes[ng_pipe_write_count].is_int = is_int;
    ng_pipe_writes[ng_pipe_write_count].v[0] = a; ng_pipe_writes[ng_pipe_write_count].v[1] = b;
    ng_pipe_writes[ng_pipe_write_count].v[2] = c; ng_pipe_writes[ng_pipe_write_count].v[3] = d;
    ++ng_pipe_write_count;
}
static void ng_uniform1i(GLint loc, GLint v) { ng_record_uniform(loc, 1, 1, (float) v, 0, 0, 0); }
static void ng_uniform1f(GLint loc, GLfloat a) { ng_record_uniform(loc, 1, 0, a, 0, 0, 0); }
static void ng_uniform2f(GLint loc, GLfloat a, GLfloat b) { ng_record_uniform(loc, 2, 0, a, b, 0, 0); }
static void ng_uniform3f(GLint loc, GLfloat a, GLfloat b, GLfloat c) { ng_record_uniform(loc, 3, 0, a, b, c, 0); }
static void ng_uniform4f(GLint loc, GLfloat a, GLfloat b, GLfloat c, GLfloat d) { ng_record_uniform(loc, 4, 0, a, b, c, d); }
int ref_rd_pipe_write_count(void) { return ng_pipe_write_count; }
/* write `idx` since the last ref_rd_update: the uniform's name (without `_IN_`), component count (-1: an int / bool), values */
const char* ref_rd_pipe_write(int idx, int* count, float* vals) {
    if (idx < 0 || idx >= ng_pipe_write_count) return NULL;
    *count = ng_pipe_writes[idx].is_int ? -1 : ng_pipe_writes[idx].count;
    memcpy(vals, ng_pipe_writes[idx].v, sizeof(float) * 4);
    return ng_names[ng_pipe_writes[idx].loc - 1] + 4;
}
static void   ng_get_objectiv(GLuint o, GLenum pname, GLint* out) { (void) o; *out = (pname == GL_INFO_LOG_LENGTH) ? 0 : GL_TRUE; }
static void   ng_get_integerv(GLenum pname, GLint* out) { (void) pname; *out = 1024; }
static GLenum ng_fb_status(GLenum target) { (void) target; return GL_FRAMEBUFFER_COMPLETE; }
static void   ng_bind_texture(GLenum target, GLuint tex) { if (target == GL_TEXTURE_1D) ng_bound_1d = tex; }

/* the text rd_new hands to the GLSL compiler (glShaderSource, render.c:335-337): injected header + glsl_ext output */
#define NG_MAX_SOURCES 64
static char* ng_sources[NG_MAX_SOURCES];
static int   ng_source_count = 0;
static void ng_shader_source(GLuint shader, GLsizei count, const GLchar* const* strings, const GLint* lengths) {
    (void) shader;
    if (ng_source_count == NG_MAX_SOURCES || count < 1) return;
    const size_t len = lengths ? (size_t) lengths[0] : strlen(strings[0]);
    ng_sources[ng_source_count] = malloc(len + 1);
    memcpy(ng_sources[ng_source_count], strings[0], len);
    ng_sources[ng_source_count][len] = '\0';
    ++ng_source_count;
}
int ref_rd_source_count(void) { return ng_source_count; }
const char* ref_rd_source(int idx) { return (idx >= 0 && idx < ng_source_count) ? ng_sources[idx] : NULL; }
void ref_rd_clear_sources(void) { for (int i = 0; i < ng_source_count; ++i) free(ng_sources[i]); ng_source_count = 0; }

#define NG_MAX_UPLOADS 64
static struct { GLuint tex; int width; float* data; } ng_uploads[NG_MAX_UPLOADS];
static int ng_upload_count = 0;
static void ng_clear_uploads(void) {
    for (int i = 0; i < ng_upload_count; ++i) free(ng_uploads[i].data);
    ng_upload_count = 0;
}
/* whole-program runs (ref_glava_entry): every float upload that differs from the previous one of the same texture is
 * appended to a file as {int32 texture id, int32 width, float[width]} */
static FILE*  ng_log_file = NULL;
static float* ng_log_last[2] = { NULL, NULL };
static GLuint ng_log_tex[2] = { 0, 0 };
static int    ng_log_width[2] = { 0, 0 };
static void ng_log_upload(GLuint tex, int width, const float* data) {
    int slot = (ng_log_tex[0] == tex || ng_log_tex[0] == 0) ? 0 : 1;
    if (ng_log_tex[slot] != 0 && ng_log_tex[slot] != tex) return;                     /* a third texture: not an audio bind */
    if (ng_log_last[slot] && ng_log_width[slot] == width && !memcmp(ng_log_last[slot], data, sizeof(float) * (size_t) width)) return;
    ng_log_tex[slot] = tex; ng_log_width[slot] = width;
    ng_log_last[slot] = realloc(ng_log_last[slot], sizeof(float) * (size_t) width);
    memcpy(ng_log_last[slot], data, sizeof(float) * (size_t) width);
    int32_t hdr[2] = { (int32_t) tex, (int32_t) width };
    fwrite(hdr, sizeof(hdr), 1, ng_log_file); fwrite(data, sizeof(float), (size_t) width, ng_log_file); fflush(ng_log_file);
}
static void ng_tex_image_1d(GLenum target, GLint level, GLint ifmt, GLsizei width, GLint border, GLenum format, GLenum type, const void* pixels) {
    (void) target; (void) level; (void) ifmt; (void) border; (void) format;
    if (pixels && type == GL_FLOAT && ng_log_file) { ng_log_upload(ng_bound_1d, (int) width, pixels); return; }
    if (!pixels || type != GL_FLOAT || ng_upload_count == NG_MAX_UPLOADS) return;     /* bind_1d_fbo allocates with NULL */
    ng_uploads[ng_upload_count].tex = ng_bound_1d;
    ng_uploads[ng_upload_count].width = (int) width;
    ng_uploads[ng_upload_count].data = malloc(sizeof(float) * (size_t) width);
    memcpy(ng_uploads[ng_upload_count].data, pixels, sizeof(float) * (size_t) width);
    ++ng_upload_count;
}

static void ng_install(void) {
#define NG_NOOP(fn) glad_##fn = (__typeof__(glad_##fn)) ng_noop
    NG_NOOP(glUseProgram); NG_NOOP(glViewport); NG_NOOP(glBindFramebuffer); NG_NOOP(glTexParameteri);
    NG_NOOP(glDisable); NG_NOOP(glActiveTexture); NG_NOOP(glEnable); NG_NOOP(glBindVertexArray);
    NG_NOOP(glBindFragDataLocation); NG_NOOP(glAttachShader); NG_NOOP(glUniform2i); NG_NOOP(glEnableVertexAttribArray);
    NG_NOOP(glDisableVertexAttribArray); NG_NOOP(glBlendEquation); NG_NOOP(glBindBuffer); NG_NOOP(glVertexAttribPointer);
    NG_NOOP(glTextureBarrierNV); NG_NOOP(glTexImage2D);
    NG_NOOP(glReadPixels); NG_NOOP(glLinkProgram); NG_NOOP(glGetShaderInfoLog); NG_NOOP(glGetProgramInfoLog);
    NG_NOOP(glFramebufferTexture2D); NG_NOOP(glFramebufferTexture1D); NG_NOOP(glDrawArrays); NG_NOOP(glCompileShader);
    NG_NOOP(glClearColor); NG_NOOP(glClear); NG_NOOP(glBufferData); NG_NOOP(glBlendFunc);
#undef NG_NOOP
    glad_glUniform1i = ng_uniform1i; glad_glUniform1f = ng_uniform1f; glad_glUniform2f = ng_uniform2f;
    glad_glUniform3f = ng_uniform3f; glad_glUniform4f = ng_uniform4f;
    glad_glGetError = ng_get_error;
    glad_glCreateShader = ng_create_shader; glad_glCreateProgram = ng_create_program;
    glad_glGenTextures = ng_gen; glad_glGenFramebuffers = ng_gen; glad_glGenVertexArrays = ng_gen; glad_glGenBuffers = ng_gen;
    glad_glGetUniformLocation = ng_uniform_location;
    glad_glGetShaderiv = ng_get_objectiv; glad_glGetProgramiv = ng_get_objectiv;
    glad_glGetIntegerv = ng_get_integerv; glad_glCheckFramebufferStatus = ng_fb_status;
    glad_glBindTexture = ng_bind_texture; glad_glTexImage1D = ng_tex_image_1d;
    glad_glShaderSource = (__typeof__(glad_glShaderSource)) ng_shader_source;
}

/* ---- window backend "null" (struct gl_wcb, render.h:66-104) ---- */
static int nw_geom[4] = { 0, 0, 800, 600 };
static bool  nw_offscreen(void) { return false; }
static void  nw_init(void) {}
static void* nw_create_and_bind(const char* name, const char* class, const char* type, const char** states, size_t states_sz,
                                int w, int h, int x, int y, int major, int minor, bool clickthrough, bool offscreen) {
    (void) name; (void) class; (void) type; (void) states; (void) states_sz; (void) major; (void) minor; (void) clickthrough; (void) offscreen;
    nw_geom[0] = x; nw_geom[1] = y; nw_geom[2] = w; nw_geom[3] = h;
    ng_install();
    glad_instantiated = true;
    return nw_geom;
}
static bool  nw_false(void* p) { (void) p; return false; }
static double nw_close_at = 0.0;                         /* whole-program runs: the window "closes" at this CLOCK_MONOTONIC time */
static double nw_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec; }
static bool  nw_should_close(void* p) { (void) p; return nw_close_at > 0.0 && nw_now() >= nw_close_at; }
static bool  nw_true(void* p) { (void) p; return true; }
static void  nw_void(void* p) { (void) p; }
static void  nw_terminate(void) {}
static void  nw_get_pos(void* p, int* x, int* y) { (void) p; *x = nw_geom[0]; *y = nw_geom[1]; }
static void  nw_get_fbsize(void* p, int* w, int* h) { (void) p; *w = nw_geom[2]; *h = nw_geom[3]; }
static void  nw_set_geometry(void* p, int x, int y, int w, int h) { (void) p; nw_geom[0] = x; nw_geom[1] = y; nw_geom[2] = w; nw_geom[3] = h; }
static void  nw_set_int(int v) { (void) v; }
static void  nw_set_bool(bool v) { (void) v; }
static double nw_get_time(void* p) { (void) p; return 0.0; }            /* frame time 0: gl->ur / gl->fr stay what the test sets */
static void  nw_set_time(void* p, double t) { (void) p; (void) t; }
static void  nw_set_visible(void* p, bool v) { (void) p; (void) v; }
static const char* nw_environment(void) { return NULL; }
static struct gl_wcb ref_null_wcb = {
    .name = "null", .offscreen = nw_offscreen, .init = nw_init, .create_and_bind = nw_create_and_bind,
    .should_close = nw_should_close, .should_render = nw_true, .bg_changed = nw_false, .swap_buffers = nw_void, .raise = nw_void,
    .destroy = nw_void, .terminate = nw_terminate, .get_pos = nw_get_pos, .get_fbsize = nw_get_fbsize,
    .set_geometry = nw_set_geometry, .set_swap = nw_set_int, .set_floating = nw_set_bool, .set_decorated = nw_set_bool,
    .set_focused = nw_set_bool, .set_maximized = nw_set_bool, .set_transparent = nw_set_bool, .get_time = nw_get_time,
    .set_time = nw_set_time, .set_visible = nw_set_visible, .get_environment = nw_environment
};

static const char* bind_types_name(int t) { return bind_types[t].n; }
static jmp_buf ref_rd_jmp;
static void ref_rd_abort(void) { longjmp(ref_rd_jmp, 1); }

/* rd_new (render.h:53-57) with backend "null", no --pipe binds.  NULL when the reference aborted. */
void* ref_rd_new(const char** paths, const char* entry, const char** requests) {
    if (wcbs_idx == 0) register_wcb(&ref_null_wcb);
    static struct rd_bind no_binds[1] = { { .name = NULL } };
    void (*saved)(void) = glava_abort;
    struct glava_renderer* r = NULL;
    glava_abort = ref_rd_abort;
    if (setjmp(ref_rd_jmp) == 0) r = rd_new(paths, entry, requests, "null", no_binds, STDIN_TYPE_NONE, false, false, false);
    glava_abort = saved;
    return r;
}

/* the same with `--pipe NAME[:TYPE]` binds (glava.c:338-411 builds exactly this array): types as STDIN_TYPE_* (render.h:32-38) */
void* ref_rd_new_binds(const char** paths, const char* entry, const char** requests, const char** bind_names, const int* bind_types) {
    if (wcbs_idx == 0) register_wcb(&ref_null_wcb);
    static struct rd_bind binds[17];
    size_t nb = 0;
    for (; bind_names && bind_names[nb] && nb < 16; ++nb)
        binds[nb] = (struct rd_bind) { .name = bind_names[nb], .stype = bind_types_name(bind_types[nb]), .type = bind_types[nb] };
    binds[nb] = (struct rd_bind) { .name = NULL };
    void (*saved)(void) = glava_abort;
    struct glava_renderer* r = NULL;
    glava_abort = ref_rd_abort;
    if (setjmp(ref_rd_jmp) == 0) r = rd_new(paths, entry, requests, "null", binds, STDIN_TYPE_NONE, false, false, false);
    glava_abort = saved;
    return r;
}

/* what the reference read from the configuration (struct glava_renderer, render.h:8-30; struct gl_data, render.c:166-207) */
void ref_rd_config(void* rp, int* ints /* [16] */, float* floats /* [12] */) {
    struct glava_renderer* r = rp; struct gl_data* gl = r->gl;
    ints[0] = (int) r->bufsize_request; ints[1] = (int) r->rate_request; ints[2] = (int) r->samplesize_request; ints[3] = r->mirror_input;
    ints[4] = (int) gl->avg_frames; ints[5] = gl->avg_window; ints[6] = gl->smooth_pass; ints[7] = gl->accel_fft;
    ints[8] = gl->interpolate; ints[9] = (int) gl->bufscale; ints[10] = gl->premultiply_alpha; ints[11] = gl->rate;
    ints[12] = gl->geometry[2]; ints[13] = gl->geometry[3]; ints[14] = (int) gl->stages_sz; ints[15] = gl->copy_desktop;
    floats[0] = gl->fft_scale; floats[1] = gl->fft_cutoff; floats[2] = gl->gravity_step; floats[3] = gl->smooth_factor;
    floats[4] = gl->smooth_distance; floats[5] = gl->smooth_ratio;
    floats[6] = gl->clear_color.r; floats[7] = gl->clear_color.g; floats[8] = gl->clear_color.b; floats[9] = gl->clear_color.a;
    floats[10] = gl->ur; floats[11] = gl->fr;
}
void ref_rd_set_rates(void* rp, float ur, float fr) { struct gl_data* gl = ((struct glava_renderer*) rp)->gl; gl->ur = ur; gl->fr = fr; }

/* one rd_update (render.h:58-59).  lb / rb are transformed IN PLACE, as in the reference.  Returns the number of audio
 * texture uploads recorded during the call (fetch them with ref_rd_upload), or -1 when the reference aborted. */
int ref_rd_update(void* rp, float* lb, float* rb, size_t bsz, int modified) {
    ng_clear_uploads();
    ng_pipe_write_count = 0;
    void (*saved)(void) = glava_abort;
    int rc = 0;
    glava_abort = ref_rd_abort;
    if (setjmp(ref_rd_jmp) == 0) { rd_time(rp); rd_update(rp, lb, rb, bsz, modified != 0); }
    else rc = -1;
    glava_abort = saved;
    return rc < 0 ? rc : ng_upload_count;
}
/* upload `idx` of the last update: which (0 = audio_l, 1 = audio_r, 2 = another texture), its width, its floats */
int ref_rd_upload(void* rp, int idx, int* which, float* out, int cap) {
    struct gl_data* gl = ((struct glava_renderer*) rp)->gl;
    if (idx < 0 || idx >= ng_upload_count) return -1;
    *which = ng_uploads[idx].tex == gl->audio_tex_l ? 0 : (ng_uploads[idx].tex == gl->audio_tex_r ? 1 : 2);
    const int w = ng_uploads[idx].width;
    memcpy(out, ng_uploads[idx].data, sizeof(float) * (size_t) (w < cap ? w : cap));
    return w;
}
void ref_rd_destroy(void* rp) { rd_destroy(rp); }


#ifdef GLAVA_REF_WITH_PROGRAM   /* only oracle/_ref/libglava_ref_rd.so links glava.c + fifo.c */
/* The whole program: glava_entry (glava/glava.c:291-577) — argument parsing, rd_new, the audio backend thread (fifo.c), the
 * frame loop with its locked ring copy (glava.c:523-552), rd_update — on the null driver, for `run_ms` milliseconds, logging the
 * audio texture uploads to `log_path`.  argv as for the `glava` binary (a "--backend=null" is what selects the null window). */
int ref_glava_entry(int argc, char** argv, long run_ms, const char* log_path) {
    if (wcbs_idx == 0) register_wcb(&ref_null_wcb);
    ng_log_file = fopen(log_path, "wb");
    if (!ng_log_file) return -1;
    nw_close_at = nw_now() + 1e-3 * (double) run_ms;
    glava_entry(argc, argv, NULL);
    fclose(ng_log_file); ng_log_file = NULL;
    nw_close_at = 0.0;
    return 0;
}
#endif


#ifdef GLAVA_REF_WITH_GL   /* only oracle/_ref/libglava_ref_gl.so */
/* ------------------------------------------------------------------------------------------------------------------
 * rd_new / rd_update WITH A REAL OpenGL: Mesa 18.1.9 llvmpipe — GLava's stated software floor (README.md:121).
 *
 * The image ships no system GL, but Nsight Compute bundles Mesa's "libgl-xlib" software GL (llvmpipe + softpipe).  It
 * wants an X display only to describe a visual; oracle/fakex/fake_x11.c supplies that without a server.  A window
 * backend "mesa" (offscreen, like the OBS plug-in's use of glava_tex) creates an OpenGL 3.3 core context on a pbuffer
 * and loads glad from glXGetProcAddress; from there on EVERYTHING is the reference's: its config reader, its CPU
 * transforms, its GL 1-D passes (pass / gravity / average / smooth fragment shaders), its module stages — compiled by
 * Mesa's GLSL compiler and rasterised by llvmpipe.  The final stage lands in gl->off_sfbo (render.c:1603-1611), which
 * ref_gl_frame reads back with glReadPixels.
 * ------------------------------------------------------------------------------------------------------------------ */
#include <dlfcn.h>

static void* mg_dpy = NULL;
static void* (*mg_gpa)(const char*) = NULL;
static void* mg_ctx = NULL;
static unsigned long mg_pbuffer = 0;
static int mg_geom[4] = { 0, 0, 800, 600 };
static char mg_error[512];

const char* ref_gl_error(void) { return mg_error; }

/* libX11 / libXext stand-ins first (RTLD_GLOBAL, so that libGL's DT_NEEDED entries resolve to them by soname), then Mesa */
int ref_gl_load(const char* fakex_dir, const char* libgl_path) {
    if (mg_gpa) return 0;
    char path[1024];
    snprintf(path, sizeof(path), "%s/libX11.so.6", fakex_dir);
    void* x11 = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!x11) { snprintf(mg_error, sizeof(mg_error), "%s", dlerror()); return -1; }
    snprintf(path, sizeof(path), "%s/libXext.so.6", fakex_dir);
    if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) { snprintf(mg_error, sizeof(mg_error), "%s", dlerror()); return -1; }
    void* gl = dlopen(libgl_path, RTLD_NOW | RTLD_GLOBAL);
    if (!gl) { snprintf(mg_error, sizeof(mg_error), "%s", dlerror()); return -1; }
    void* (*fdpy)(void) = (void* (*)(void)) dlsym(x11, "fakex_display");
    mg_gpa = (void* (*)(const char*)) dlsym(gl, "glXGetProcAddressARB");
    if (!fdpy || !mg_gpa) { snprintf(mg_error, sizeof(mg_error), "fakex_display / glXGetProcAddressARB not found"); mg_gpa = NULL; return -1; }
    mg_dpy = fdpy();
    return 0;
}

static int mg_no_barrier = 0;
static void mg_texture_barrier(void) { glFinish(); }
int ref_gl_barrier_substituted(void) { return mg_no_barrier; }
static void (*mg_real_draw)(GLenum, GLint, GLsizei) = NULL;
static void mg_trace_draw(GLenum mode, GLint first, GLsizei count) {      /* REF_GL_TRACE_DRAWS=1: state at every draw */
    GLint fbo = 0, prog = 0, vp[4] = { 0 }, blend = 0, eq = 0, unit = 0, t1d = 0, db = 0;
    glGetIntegerv(GL_DRAW_FRAMEBUFFER_BINDING, &fbo); glGetIntegerv(GL_CURRENT_PROGRAM, &prog); glGetIntegerv(GL_VIEWPORT, vp);
    glGetIntegerv(GL_BLEND, &blend); glGetIntegerv(GL_BLEND_EQUATION_RGB, &eq); glGetIntegerv(GL_ACTIVE_TEXTURE, &unit);
    glGetIntegerv(GL_TEXTURE_BINDING_1D, &t1d); glGetIntegerv(GL_DRAW_BUFFER, &db);
    const GLenum st = glCheckFramebufferStatus(GL_DRAW_FRAMEBUFFER);
    mg_real_draw(mode, first, count);
    if (vp[3] == 1) {
        unsigned short px[4] = { 0 }; static float fx[65536];
        GLint rfb = 0; glGetIntegerv(GL_READ_FRAMEBUFFER_BINDING, &rfb);
        glBindFramebuffer(GL_READ_FRAMEBUFFER, fbo);
        glReadPixels(0, 0, 4, 1, GL_RED, GL_UNSIGNED_SHORT, px);
        glBindFramebuffer(GL_READ_FRAMEBUFFER, rfb);
        glGetTexImage(GL_TEXTURE_1D, 0, GL_RED, GL_FLOAT, fx);
        fprintf(stderr, "   target now %u %u %u %u | bound 1-D texture starts %g %g\n", px[0], px[1], px[2], px[3], fx[0], fx[1]);
    }
    fprintf(stderr, "[draw] fbo %d (status %x, draw buffer %x) prog %d viewport %d %d %d %d blend %d eq %x unit %d tex1d %d err %x\n",
            fbo, st, db, prog, vp[0], vp[1], vp[2], vp[3], blend, eq, unit - GL_TEXTURE0, t1d, glGetError());
}
static bool  mw_offscreen(void) { return true; }
static void* mw_create_and_bind(const char* name, const char* class, const char* type, const char** states, size_t states_sz,
                                int w, int h, int x, int y, int major, int minor, bool clickthrough, bool offscreen) {
    (void) name; (void) class; (void) type; (void) states; (void) states_sz; (void) clickthrough; (void) offscreen;
    mg_geom[0] = x; mg_geom[1] = y; mg_geom[2] = w; mg_geom[3] = h;
    if (!mg_gpa) { fprintf(stderr, "ref_gl: ref_gl_load was not called\n"); glava_abort(); }
    {
        /* a fresh context per renderer, as every GLava process has: rd_new relies on GL's initial state (blending off,
         * ...) and e.g. `setopacity "none"` leaves GL_BLEND enabled behind */
        if (mg_ctx) {
            int (*mkcur0)(void*, unsigned long, unsigned long, void*) = (int (*)(void*, unsigned long, unsigned long, void*)) mg_gpa("glXMakeContextCurrent");
            void (*destroy)(void*, void*) = (void (*)(void*, void*)) mg_gpa("glXDestroyContext");
            mkcur0(mg_dpy, 0, 0, NULL);
            destroy(mg_dpy, mg_ctx);
            mg_ctx = NULL;
        }
        void** (*choose)(void*, int, const int*, int*) = (void** (*)(void*, int, const int*, int*)) mg_gpa("glXChooseFBConfig");
        void*  (*mkpb)(void*, void*, const int*) = (void* (*)(void*, void*, const int*)) mg_gpa("glXCreatePbuffer");
        void*  (*mkctx)(void*, void*, void*, int, const int*) = (void* (*)(void*, void*, void*, int, const int*)) mg_gpa("glXCreateContextAttribsARB");
        const int fbattr[] = { 0x8010 /* GLX_DRAWABLE_TYPE */, 0x4 | 0x1 /* PBUFFER | WINDOW */, 0x8011 /* GLX_RENDER_TYPE */, 0x1 /* RGBA */,
                               8 /* RED */, 8, 9, 8, 10, 8, 11 /* ALPHA */, 8, 12 /* DEPTH */, 0, 5 /* DOUBLEBUFFER */, 0, 0 };
        int n = 0;
        void** cfgs = choose(mg_dpy, 0, fbattr, &n);
        if (!cfgs || n < 1) { fprintf(stderr, "ref_gl: no GLX framebuffer configuration\n"); glava_abort(); }
        const int pbattr[] = { 0x8041 /* GLX_PBUFFER_WIDTH */, 16, 0x8040 /* GLX_PBUFFER_HEIGHT */, 16, 0 };
        if (!mg_pbuffer) mg_pbuffer = (unsigned long) mkpb(mg_dpy, cfgs[0], pbattr);
        const int cattr[] = { 0x2091 /* MAJOR */, major, 0x2092 /* MINOR */, minor, 0x9126 /* PROFILE_MASK */, 0x1 /* core, as glx_wcb.c asks */, 0 };
        mg_ctx = mkctx(mg_dpy, cfgs[0], NULL, 1, cattr);
        if (!mg_ctx) { fprintf(stderr, "ref_gl: glXCreateContextAttribsARB(%d.%d core) failed\n", major, minor); glava_abort(); }
    }
    int (*mkcur)(void*, unsigned long, unsigned long, void*) = (int (*)(void*, unsigned long, unsigned long, void*)) mg_gpa("glXMakeContextCurrent");
    if (!mkcur(mg_dpy, mg_pbuffer, mg_pbuffer, mg_ctx)) { fprintf(stderr, "ref_gl: glXMakeContextCurrent failed\n"); glava_abort(); }
    if (!glad_instantiated) {
        if (!gladLoadGLLoader((GLADloadproc) mg_gpa)) { fprintf(stderr, "ref_gl: glad could not load OpenGL\n"); glava_abort(); }
        glad_instantiated = true;
    }
    /* render.c:2217 calls glTextureBarrierNV unconditionally (in-place gravity pass on one texture).  This Mesa does not
     * advertise GL_NV_texture_barrier for llvmpipe, so glad leaves the pointer NULL and the reference would jump to 0.
     * llvmpipe flushes a scene that renders to a resource before a later draw samples that resource, which is all the
     * barrier asks for; glFinish() stands in for it. */
    if (!glad_glTextureBarrierNV) { mg_no_barrier = 1; glad_glTextureBarrierNV = mg_texture_barrier; }
    if (getenv("REF_GL_TRACE_DRAWS")) { mg_real_draw = glad_glDrawArrays; glad_glDrawArrays = mg_trace_draw; }
    return mg_geom;
}
static void  mw_get_pos(void* p, int* x, int* y) { (void) p; *x = mg_geom[0]; *y = mg_geom[1]; }
static void  mw_get_fbsize(void* p, int* w, int* h) { (void) p; *w = mg_geom[2]; *h = mg_geom[3]; }
static void  mw_set_geometry(void* p, int x, int y, int w, int h) { (void) p; mg_geom[0] = x; mg_geom[1] = y; mg_geom[2] = w; mg_geom[3] = h; }
static struct gl_wcb ref_mesa_wcb = {
    .name = "mesa", .offscreen = mw_offscreen, .init = nw_init, .create_and_bind = mw_create_and_bind,
    .should_close = nw_false, .should_render = nw_true, .bg_changed = nw_false, .swap_buffers = nw_void, .raise = nw_void,
    .destroy = nw_void, .terminate = nw_terminate, .get_pos = mw_get_pos, .get_fbsize = mw_get_fbsize,
    .set_geometry = mw_set_geometry, .set_swap = nw_set_int, .set_floating = nw_set_bool, .set_decorated = nw_set_bool,
    .set_focused = nw_set_bool, .set_maximized = nw_set_bool, .set_transparent = nw_set_bool, .get_time = nw_get_time,
    .set_time = nw_set_time, .set_visible = nw_set_visible, .get_environment = nw_environment
};

const char* ref_gl_strings(int which) {
    return (const char*) glGetString(which == 0 ? GL_VERSION : which == 1 ? GL_RENDERER : GL_SHADING_LANGUAGE_VERSION);
}

/* rd_new with the real GL.  NULL when the reference aborted (shader compile errors are printed by the reference). */
void* ref_gl_new(const char** paths, const char* entry, const char** requests) {
    bool have = false;
    for (size_t t = 0; t < wcbs_idx; ++t) have |= wcbs[t] == &ref_mesa_wcb;
    if (!have) register_wcb(&ref_mesa_wcb);
    static struct rd_bind no_binds[1] = { { .name = NULL } };
    void (*saved)(void) = glava_abort;
    struct glava_renderer* r = NULL;
    glava_abort = ref_rd_abort;
    if (setjmp(ref_rd_jmp) == 0) r = rd_new(paths, entry, requests, "mesa", no_binds, STDIN_TYPE_NONE, false, false, false);
    glava_abort = saved;
    return r;
}
void ref_gl_size(void* rp, int* w, int* h) { (void) rp; *w = mg_geom[2]; *h = mg_geom[3]; }

/* One rd_update (lb / rb transformed in place, as in the reference), then the frame: RGBA8, row 0 = bottom.
 * returns 0, or -1 when the reference aborted */
int ref_gl_frame(void* rp, float* lb, float* rb, size_t bsz, int modified, unsigned char* rgba) {
    struct glava_renderer* r = rp; struct gl_data* gl = r->gl;
    void (*saved)(void) = glava_abort;
    int rc = 0;
    glava_abort = ref_rd_abort;
    if (setjmp(ref_rd_jmp) == 0) { rd_time(rp); rd_update(rp, lb, rb, bsz, modified != 0); }
    else rc = -1;
    glava_abort = saved;
    if (rc == 0 && rgba) {
        glBindFramebuffer(GL_READ_FRAMEBUFFER, gl->off_sfbo.fbo);
        glPixelStorei(GL_PACK_ALIGNMENT, 1);
        glReadPixels(0, 0, mg_geom[2], mg_geom[3], GL_RGBA, GL_UNSIGNED_BYTE, rgba);
        glBindFramebuffer(GL_READ_FRAMEBUFFER, 0);
    }
    return rc;
}

/* The 1-D texture stage 1 samples for audio_l (which = 0) / audio_r (1) after the last rd_update, as R16 texels:
 * the smooth pass' output, or the average / gravity / raw upload when later passes are off.  returns the width, or -1 */
int ref_gl_texture(void* rp, int which, unsigned short* out, int cap) {
    struct glava_renderer* r = rp; struct gl_data* gl = r->gl;
    if (gl->stages_sz == 0) return -1;
    struct gl_sfbo* st = &gl->stages[0];
    for (size_t b = 0; b < st->binds_sz; ++b) {
        struct gl_bind* bd = &st->binds[b];
        if (bd->src_type != (which ? SRC_AUDIO_R : SRC_AUDIO_L)) continue;
        GLuint tex = which ? gl->audio_tex_r : gl->audio_tex_l;
        if (bd->optimize_fft) { tex = bd->gr_store.tex; if (gl->avg_frames > 1) tex = bd->av.tex; }
        if (gl->smooth_pass) tex = bd->sm.tex;
        GLint w = 0;
        glBindTexture(GL_TEXTURE_1D, tex);
        glGetTexLevelParameteriv(GL_TEXTURE_1D, 0, GL_TEXTURE_WIDTH, &w);
        if (w < 1 || w > cap) return -1;
        glPixelStorei(GL_PACK_ALIGNMENT, 1);
        glGetTexImage(GL_TEXTURE_1D, 0, GL_RED, GL_UNSIGNED_SHORT, out);
        return (int) w;
    }
    return -1;
}
/* any 1-D texture of a bind, for looking at the passes one by one: stage 0, bind of audio_l / audio_r,
 * what = 0 upload (transform_fft output), 1 gr_store (K1 / K2), 2 av (K4), 3 sm (K5), 4.. gr.out[what - 4] (K3 ring) */
int ref_gl_pass_texture(void* rp, int which, int what, unsigned short* out, int cap) {
    struct glava_renderer* r = rp; struct gl_data* gl = r->gl;
    if (gl->stages_sz == 0) return -1;
    struct gl_sfbo* st = &gl->stages[0];
    for (size_t b = 0; b < st->binds_sz; ++b) {
        struct gl_bind* bd = &st->binds[b];
        if (bd->src_type != (which ? SRC_AUDIO_R : SRC_AUDIO_L)) continue;
        GLuint tex = 0;
        switch (what) {
            case 0: tex = which ? gl->audio_tex_r : gl->audio_tex_l; break;
            case 1: tex = bd->gr_store.tex; break;
            case 2: tex = bd->av.tex; break;
            case 3: tex = bd->sm.tex; break;
            default: if (bd->gr.out && (size_t) (what - 4) < bd->gr.out_sz) tex = bd->gr.out[what - 4].tex; break;
        }
        if (!tex) return -1;
        GLint w = 0;
        glBindTexture(GL_TEXTURE_1D, tex);
        glGetTexLevelParameteriv(GL_TEXTURE_1D, 0, GL_TEXTURE_WIDTH, &w);
        if (w < 1 || w > cap) return -1;
        glPixelStorei(GL_PACK_ALIGNMENT, 1);
        glGetTexImage(GL_TEXTURE_1D, 0, GL_RED, GL_UNSIGNED_SHORT, out);
        return (int) w;
    }
    return -1;
}
/* the shader tree packed into this library by the build (oracle/pack_shaders.py) -> files under `dir`; returns the
 * number of files written, -1 on an I/O error */
#include <sys/stat.h>
extern const unsigned char ref_shader_blob[];
extern const unsigned long ref_shader_blob_size;
int ref_gl_unpack_shaders(const char* dir) {
    const unsigned char* p = ref_shader_blob;
    unsigned count; memcpy(&count, p, 4); p += 4;
    for (unsigned i = 0; i < count; ++i) {
        unsigned short plen; memcpy(&plen, p, 2); p += 2;
        char path[2048];
        int base = snprintf(path, sizeof(path), "%s/", dir);
        if (base < 0 || (size_t) base + plen + 1 > sizeof(path)) return -1;
        memcpy(path + base, p, plen); path[base + plen] = '\0'; p += plen;
        unsigned size; memcpy(&size, p, 4); p += 4;
        for (char* q = path + base; *q; ++q) if (*q == '/') { *q = '\0'; mkdir(path, 0755); *q = '/'; }
        FILE* f = fopen(path, "wb");
        if (!f) return -1;
        if (size && fwrite(p, 1, size, f) != size) { fclose(f); return -1; }
        fclose(f); p += size;
    }
    return (int) count;
}
void ref_gl_finish(void) { glFinish(); }           /* llvmpipe rasterises on worker threads: the frame is complete after this */
int ref_gl_get_error(void) { return (int) glGetError(); }
int ref_gl_stage_count(void* rp) { return (int) ((struct glava_renderer*) rp)->gl->stages_sz; }
void ref_gl_destroy(void* rp) { rd_destroy(rp); }
#endif
