// config.cpp — the `#request` / `#define` configuration surface of the hot path.
//
// Mirrors, for this path only, what rd_new does with the entry file and the module
// config (reference: glava/render.c:1033-1314 request table, :1322-1435 entry + CLI
// requests; glava/glsl_ext.c:228-300 request argument parsing, :489-514 colour literals,
// :516-591 `@name:default` pipe binds).  It is a line-oriented reader, not a GLSL
// preprocessor: module shaders are CUDA kernels here, so only the values of the
// documented `#define`s of <module>.glsl / smooth_parameters.glsl are extracted.
#include "internal.h"
#include "raster_core.h"      // eval_color_prog: constant colour expressions are evaluated at config time

#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace glb {

static const float kPI = 3.14159265359f, kTWOPI = 6.28318530718f;   // bars/1.frag:33-34 literals

// ---- colour literals: "#rrggbb[aa]" -> the "%.6f" decimals glsl_ext.c:505 emits ------------
static float hex_channel(unsigned v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "%.6f", (double) ((float) v / (float) 255));
    return strtof(buf, nullptr);
}
// ext_parse_color with two digits per component (glsl_ext.c:88-122) as `setbg` and the `--pipe` colour values use it: optional
// `0x`, up to 8 hex digits, each complete pair sets one component to e / 255 (no "%.6f" detour), components the string does
// not reach keep their value, a trailing odd digit is dropped; false on a non-hex character.
bool parse_hex_components(const char* s, float* out[4]) {
    size_t len = strlen(s);
    if (len >= 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) { s += 2; len -= 2; }
    unsigned acc = 0; int have = 0, comp = 0;
    for (size_t k = 0; k < len && k < 8; ++k) {
        const char c = s[k];
        unsigned v;
        if (c >= 'a' && c <= 'f') v = (unsigned) (c - 'a') + 10;
        else if (c >= 'A' && c <= 'F') v = (unsigned) (c - 'A') + 10;
        else if (c >= '0' && c <= '9') v = (unsigned) (c - '0');
        else return false;
        acc = (acc << 4) | v;
        if (++have == 2) { *out[comp++] = (float) acc / (float) 255; acc = 0; have = 0; }
    }
    return true;
}
bool parse_hex_color(const char* s, float out[4], bool literal_rounding) {
    if (s[0] == '#') ++s;
    if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;             // glsl_ext.c:92-95
    size_t len = strlen(s);
    if (len > 8) len = 8;
    unsigned comp[4] = { 0, 0, 0, 255 };
    size_t ncomp = 0;
    for (size_t i = 0; i + 2 <= len && ncomp < 4; i += 2) {
        unsigned v = 0;
        for (int k = 0; k < 2; ++k) {
            char c = s[i + k]; unsigned d;
            if (c >= '0' && c <= '9') d = (unsigned) (c - '0');
            else if (c >= 'a' && c <= 'f') d = (unsigned) (c - 'a') + 10;
            else if (c >= 'A' && c <= 'F') d = (unsigned) (c - 'A') + 10;
            else return false;
            v = v * 16 + d;
        }
        comp[ncomp++] = v;
    }
    for (int i = 0; i < 4; ++i) {
        if ((size_t) i < ncomp || i == 3)
            out[i] = literal_rounding ? hex_channel(comp[i]) : (float) comp[i] / (float) 255;
        else out[i] = 0.0f;
    }
    if (ncomp < 4) out[3] = 1.0f;                                           // glsl_ext.c:503
    return true;
}

// ---- defaults: the shipped rc.glsl / smooth_parameters.glsl / <module>.glsl ------------------
static void hex3(float* o, unsigned r, unsigned g, unsigned b) {
    o[0] = hex_channel(r); o[1] = hex_channel(g); o[2] = hex_channel(b); o[3] = 1.0f;
}

int module_from_name(const char* name) {
    static const char* names[] = { "bars", "radial", "circle", "graph", "wave", "test" };
    for (int i = 0; i < 6; ++i) if (!strcmp(name, names[i])) return i;
    return -1;
}
const char* module_name(int id) {
    static const char* names[] = { "bars", "radial", "circle", "graph", "wave", "test" };
    return (id >= 0 && id < 6) ? names[id] : "?";
}

void fill_defaults(glava_b200_params* p, int module) {
    memset(p, 0, sizeof(*p));
    p->n = 4096; p->fft_scale = 10.2f; p->fft_cutoff = 0.3f; p->gravity_step = 4.2f;
    p->rate_request = 22050; p->samplesize_request = 1024;
    p->ur = (float) p->rate_request / (float) (p->samplesize_request / 4);
    p->avg_frames = 5; p->avg_window = 1; p->accel_fft = 1; p->smooth_pass = 1;
    p->smooth_factor = 0.025f; p->sample_range = 0.9f; p->sample_scale = 8.0f;
    p->hybrid_weight = 0.65f; p->sample_mode = 0; p->round_formula = 0;
    p->module = module; p->w = 800; p->h = 600; p->channels = 2; p->mirror_input = 0; p->premultiply_alpha = 1;
    p->bars_width = 5; p->bars_gap = 1; p->bars_outline_width = 1; p->bars_amplify = 300;
    p->bars_color.mode = 0; hex3(p->bars_color.lo, 0x33, 0x66, 0xb2); hex3(p->bars_color.hi, 0xa0, 0xa0, 0xb2);
    p->bars_color.gradient = 80; p->bars_outline_mode = 0;
    p->radial_radius = 128; p->radial_line = 2; p->radial_line_half = 1; hex3(p->radial_outline, 0x33, 0x33, 0x33);
    hex3(p->radial_bar_outline, 0x33, 0x33, 0x33); p->radial_bar_outline_width = 0; p->radial_bar_width_int = 0;   // radial.glsl:33-36, BAR_WIDTH 4.5
    p->radial_nbars = 160; p->radial_bar_width = 4.5f; p->radial_amplify = 300;
    p->radial_color.mode = 0; hex3(p->radial_color.lo, 0xcc, 0x33, 0x33); hex3(p->radial_color.hi, 0xcc, 0xa0, 0xa0);
    p->radial_color.gradient = 95; p->radial_rotate = kPI / 2; p->radial_bar_alias = 1.2f; p->radial_c_alias = 1.8f;
    p->circle_radius = 128; p->circle_line = 1.5f; hex3(p->circle_outline, 0x33, 0x33, 0x33);
    p->circle_amplify = 150; p->circle_rotate = kPI / 2; p->circle_smooth = 1;
    p->graph_vscale = 300; p->graph_direction = 1;
    p->graph_color.mode = 0; hex3(p->graph_color.lo, 0x80, 0x2a, 0x2a); hex3(p->graph_color.hi, 0x4f, 0x4f, 0x92);
    p->graph_color.gradient = 75; p->graph_draw_highlight = 1; hex3(p->graph_outline, 0x26, 0x26, 0x26);
    p->wave_min_thickness = 1; p->wave_max_thickness = 6; p->wave_amplify = 500;
    p->wave_base_color[0] = 0.7f; p->wave_base_color[1] = 0.2f; p->wave_base_color[2] = 0.45f; p->wave_base_color[3] = 1;
    p->wave_outline[0] = p->wave_outline[1] = p->wave_outline[2] = 0.15f; p->wave_outline[3] = 1;
    p->fb_slots = 0; p->lazy_smooth = 0;
    p->bufscale = 1; p->interpolate = 0; p->fr = 0.0f;                       // render.c:908, rc.glsl:131
    p->transform_smooth = 0; p->smooth_distance = 0.01f; p->smooth_ratio = 4.0f;   // render.c:917-918
}

// ---- tiny expression evaluator for numeric #defines: + - * / ( ) literals PI TWOPI ---------
struct Num { double v; bool is_int; };
static std::string strip_bind(const std::string& v);      // `@name:default` -> bound value or default (below)
struct ExprParser {
    const char* s; const std::map<std::string, std::string>* defs; int depth; bool ok;
    void ws() { while (*s && isspace((unsigned char) *s)) ++s; }
    Num primary() {
        ws();
        if (*s == '(') { ++s; Num r = expr(); ws(); if (*s == ')') ++s; else ok = false; return r; }
        if (*s == '-') { ++s; Num r = primary(); r.v = -r.v; return r; }
        if (*s == '+') { ++s; return primary(); }
        if (isdigit((unsigned char) *s) || *s == '.') {
            char* end; double v = strtod(s, &end);
            bool is_int = true;
            for (const char* c = s; c < end; ++c) if (*c == '.' || *c == 'e' || *c == 'E') is_int = false;
            s = end;
            if (*s == 'f' || *s == 'F') { ++s; is_int = false; }
            if (!is_int) v = (double) (float) v;                            // GLSL float literal
            return { v, is_int };
        }
        if (isalpha((unsigned char) *s) || *s == '_') {
            std::string id;
            while (isalnum((unsigned char) *s) || *s == '_') id += *s++;
            if (id == "PI") return { (double) kPI, false };
            if (id == "TWOPI") return { (double) kTWOPI, false };
            if (id == "float" || id == "int") { Num r = primary(); if (id == "float") r.is_int = false; return r; }
            auto it = defs ? defs->find(id) : decltype(defs->end())();
            if (defs && it != defs->end() && depth < 8) {
                const std::string body = strip_bind(it->second);
                ExprParser sub { body.c_str(), defs, depth + 1, true };
                Num r = sub.expr(); sub.ws();
                if (!sub.ok || *sub.s) ok = false;
                return r;
            }
        }
        ok = false; return { 0, true };
    }
    Num term() {
        Num a = primary();
        for (;;) {
            ws();
            if (*s == '*') { ++s; Num b = primary(); bool i = a.is_int && b.is_int;
                a = { i ? (double) ((long) a.v * (long) b.v) : (double) ((float) a.v * (float) b.v), i }; }
            else if (*s == '/') { ++s; Num b = primary(); bool i = a.is_int && b.is_int;
                if (i) { if ((long) b.v == 0) { ok = false; return a; } a = { (double) ((long) a.v / (long) b.v), true }; }
                else a = { (double) ((float) a.v / (float) b.v), false }; }
            else return a;
        }
    }
    Num expr() {
        Num a = term();
        for (;;) {
            ws();
            if (*s == '+') { ++s; Num b = term(); bool i = a.is_int && b.is_int;
                a = { i ? a.v + b.v : (double) ((float) a.v + (float) b.v), i }; }
            else if (*s == '-') { ++s; Num b = term(); bool i = a.is_int && b.is_int;
                a = { i ? a.v - b.v : (double) ((float) a.v - (float) b.v), i }; }
            else return a;
        }
    }
};

typedef std::map<std::string, std::string> Defs;

static std::string trim_copy(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char) s[a])) ++a;
    while (b > a && isspace((unsigned char) s[b - 1])) --b;
    return s.substr(a, b - a);
}

static bool eval_num(const Defs& d, const char* name, Num* out) {
    auto it = d.find(name);
    if (it == d.end()) return false;
    const std::string body = strip_bind(it->second);
    ExprParser p { body.c_str(), &d, 0, true };
    Num r = p.expr(); p.ws();
    if (!p.ok || *p.s) {
        fail(GLAVA_B200_ECONFIG, "cannot evaluate '#define %s %s' as a number", name, it->second.c_str());
        return false;
    }
    *out = r; return true;
}
static void getf(const Defs& d, const char* name, float* dst) { Num n; if (eval_num(d, name, &n)) *dst = (float) n.v; }
static void geti(const Defs& d, const char* name, int* dst) { Num n; if (eval_num(d, name, &n)) *dst = (int) n.v; }
// A macro the module's shader tests in a preprocessor conditional (`where` = the shader line).  GLSL's preprocessor
// evaluates integer constant expressions only: with a float spelling ("#define BAR_OUTLINE_WIDTH 0.5") the reference's
// shader does not compile — Mesa: `preprocessor error: syntax error, unexpected OTHER` — and GLava aborts in shaderload
// (render.c:352-377).  Same outcome here: a configuration error naming the macro.  (Seen when the llvmpipe harness first
// compiled the fuzzer's configurations; the in-house interpreter had accepted them.)
static bool getpp(const Defs& d, const char* name, const char* where, Num* out) {
    if (!eval_num(d, name, out)) return false;
    if (!out->is_int) {
        fail(GLAVA_B200_ECONFIG, "'#define %s %s': %s tests this macro in a preprocessor #if, which takes integer constant "
                                 "expressions only — the reference's shader fails to compile with a float here", name,
             d.find(name)->second.c_str(), where);
        return false;
    }
    return true;
}
static void getppi(const Defs& d, const char* name, const char* where, int* dst) { Num n; if (getpp(d, name, where, &n)) *dst = (int) n.v; }
static void getppf(const Defs& d, const char* name, const char* where, float* dst) { Num n; if (getpp(d, name, where, &n)) *dst = (float) n.v; }

// strip `@name:` pipe-bind prefix -> its default (glsl_ext.c:571-587 when no --pipe bind exists)
// `--pipe` binds (glava.c:421-436, glsl_ext.c:516-591): "name" -> value the bound uniform currently holds.
// A macro written `@name:default` takes the bound value when `name` is bound, else `default`.
static thread_local const std::map<std::string, std::string>* g_binds = nullptr;

static std::string strip_bind(const std::string& v) {
    size_t i = 0; while (i < v.size() && isspace((unsigned char) v[i])) ++i;
    if (i < v.size() && v[i] == '@') {
        size_t c = v.find(':', i);
        if (c != std::string::npos) {
            if (g_binds) {
                auto it = g_binds->find(v.substr(i + 1, c - i - 1));
                if (it != g_binds->end()) return it->second;
            }
            return v.substr(c + 1);
        }
        if (g_binds) {                                   // `@name` without a default (glsl_ext.c:571-587)
            auto it = g_binds->find(trim_copy(v.substr(i + 1)));
            if (it != g_binds->end()) return it->second;
        }
    }
    return v.substr(i);
}
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char) s[a])) ++a;
    while (b > a && isspace((unsigned char) s[b - 1])) --b;
    return s.substr(a, b - a);
}
// split "f(a, b(c, d), e)" argument list at top-level commas
static bool split_call(const std::string& s, const char* fn, std::vector<std::string>* args) {
    std::string t = trim(s);
    size_t fl = strlen(fn);
    if (t.compare(0, fl, fn) != 0) return false;
    size_t i = fl; while (i < t.size() && isspace((unsigned char) t[i])) ++i;
    if (i >= t.size() || t[i] != '(' || t.back() != ')') return false;
    std::string inner = t.substr(i + 1, t.size() - i - 2);
    int depth = 0; std::string cur;
    for (char c : inner) {
        if (c == '(') ++depth;
        if (c == ')') --depth;
        if (c == ',' && depth == 0) { args->push_back(trim(cur)); cur.clear(); } else cur += c;
    }
    if (depth != 0) return false;
    args->push_back(trim(cur));
    return true;
}
static bool parse_const_color(const Defs& d, const std::string& s, float out[4]) {
    std::string t = trim(s);
    if (!t.empty() && t[0] == '#') return parse_hex_color(t.c_str(), out, true);
    std::vector<std::string> a;
    if (split_call(t, "vec4", &a) && a.size() == 4) {
        for (int i = 0; i < 4; ++i) {
            ExprParser p { a[i].c_str(), &d, 0, true };
            Num n = p.expr(); p.ws();
            if (!p.ok || *p.s) return false;
            out[i] = (float) n.v;
        }
        return true;
    }
    return false;
}
}  // namespace glb
#include "color_compile.h"
namespace glb {

// COLOR forms: constant colour (mode 1) | mix(<colour>, <colour>, clamp(<var> / <num>, 0, 1)) (mode 0, the shipped form)
// | any other expression color_compile.h can compile (mode 2, program in *prog).  `vars`: names of the per-pixel variable.
static bool compile_color_text(const Defs& d, const char* name, const std::string& text, std::vector<std::string> vars,
                               glava_b200_color_prog* prog, bool* uses_var) {
    ColorCompiler cc(&d, std::move(vars));
    if (!cc.compile(text, prog)) {
        fail(GLAVA_B200_ECONFIG, "unsupported colour expression in '#define %s %s': %s", name, text.c_str(), cc.err.c_str());
        return false;
    }
    if (uses_var) *uses_var = cc.uses_var;
    return true;
}
static bool const_equals(const Defs& d, const std::string& text, double want) {
    ExprParser p { text.c_str(), &d, 0, true };
    Num n = p.expr(); p.ws();
    return p.ok && !*p.s && n.v == want;
}
static bool parse_color_macro(const Defs& d, const char* name, glava_b200_color* c, glava_b200_color_prog* prog,
                              std::vector<std::string> vars) {
    auto it = d.find(name);
    if (it == d.end()) return true;
    std::string v = strip_bind(it->second);
    prog->n_ops = 0;
    float k[4];
    if (parse_const_color(d, v, k)) { c->mode = 1; memcpy(c->lo, k, sizeof(k)); memcpy(c->hi, k, sizeof(k)); return true; }
    std::vector<std::string> a, b;
    float lo[4], hi[4];
    if (split_call(v, "mix", &a) && a.size() == 3 && parse_const_color(d, a[0], lo) && parse_const_color(d, a[1], hi)
        && split_call(a[2], "clamp", &b) && b.size() == 3 && const_equals(d, b[1], 0.0) && const_equals(d, b[2], 1.0)) {
        size_t slash = b[0].find('/');
        bool var_ok = false;
        if (slash != std::string::npos) for (const std::string& x : vars) var_ok |= trim(b[0].substr(0, slash)) == x;
        if (var_ok) {
            ExprParser p { b[0].c_str() + slash + 1, &d, 0, true };
            Num n = p.expr(); p.ws();
            if (p.ok && !*p.s) {
                c->mode = 0; c->gradient = (float) n.v;
                memcpy(c->lo, lo, sizeof(lo)); memcpy(c->hi, hi, sizeof(hi));
                return true;
            }
        }
    }
    bool uses_var = false;
    if (!compile_color_text(d, name, v, vars, prog, &uses_var)) return false;
    if (!uses_var) {                                         // a constant after all: fold it
        const f4 r = eval_color_prog(*prog, 0.0f);
        c->mode = 1; c->lo[0] = c->hi[0] = r.r; c->lo[1] = c->hi[1] = r.g; c->lo[2] = c->hi[2] = r.b; c->lo[3] = c->hi[3] = r.a;
        prog->n_ops = 0;
        return true;
    }
    c->mode = 2;
    return true;
}
static bool parse_plain_color(const Defs& d, const char* name, float out[4]) {
    auto it = d.find(name);
    if (it == d.end()) return true;
    const std::string v = strip_bind(it->second);
    if (parse_const_color(d, v, out)) return true;
    glava_b200_color_prog prog; bool uses_var = false;
    if (!compile_color_text(d, name, v, {}, &prog, &uses_var)) return false;      // no variable is in scope: constants only
    const f4 r = eval_color_prog(prog, 0.0f);
    out[0] = r.r; out[1] = r.g; out[2] = r.b; out[3] = r.a;
    return true;
}

// ---- file reading ---------------------------------------------------------------------------
static bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::stringstream ss; ss << f.rdbuf(); *out = ss.str();
    return true;
}
// remove /* */ and // comments, keep newlines
static std::string strip_comments(const std::string& in) {
    std::string o; o.reserve(in.size());
    for (size_t i = 0; i < in.size();) {
        if (in[i] == '/' && i + 1 < in.size() && in[i + 1] == '*') {
            i += 2;
            while (i + 1 < in.size() && !(in[i] == '*' && in[i + 1] == '/')) { if (in[i] == '\n') o += '\n'; ++i; }
            i += 2;
        } else if (in[i] == '/' && i + 1 < in.size() && in[i + 1] == '/') {
            while (i < in.size() && in[i] != '\n') ++i;
        } else o += in[i++];
    }
    return o;
}
// tokens of a request line: whitespace separated, "quoted strings" kept whole (glsl_ext.c:228-300)
static std::vector<std::string> tokenize(const std::string& line) {
    std::vector<std::string> t; std::string cur; bool q = false, have = false;
    for (char c : line) {
        if (q) { if (c == '"') { q = false; } else cur += c; continue; }
        if (c == '"') { q = true; have = true; continue; }
        if (isspace((unsigned char) c)) { if (have) { t.push_back(cur); cur.clear(); have = false; } continue; }
        cur += c; have = true;
    }
    if (have) t.push_back(cur);
    return t;
}

static bool parse_bool(const std::string& s, bool* v) {                     // glsl_ext.c:266-288
    if (s == "true" || s == "t" || s == "1") { *v = true; return true; }
    if (s == "false" || s == "f" || s == "0") { *v = false; return true; }
    return false;
}

struct Loader {
    glava_b200_params* p;
    std::string module;        // `#request mod`
    bool module_forced;
    bool failed;
};

// returns false on error (message already reported)
static bool apply_request(Loader& L, const std::vector<std::string>& t, const char* where, int line) {
    if (t.empty()) return true;
    const std::string& name = t[0];
    auto need = [&](size_t n) -> bool {
        if (t.size() < 1 + n) {
            fail(GLAVA_B200_ECONFIG, "[%s:%d] failed to execute request '%s': expected %d argument(s)", where, line,
                 name.c_str(), (int) n);
            return false;
        }
        return true;
    };
    auto as_int = [&](size_t i) { return (int) strtol(t[i].c_str(), nullptr, 0); };
    auto as_f   = [&](size_t i) { return strtof(t[i].c_str(), nullptr); };
    auto as_b   = [&](size_t i, bool* v) -> bool {
        if (parse_bool(t[i], v)) return true;
        fail(GLAVA_B200_ECONFIG, "[%s:%d] tried to parse invalid raw string into a boolean", where, line);
        return false;
    };
    glava_b200_params* p = L.p;
    bool b;
    if (name == "mod") { if (!need(1)) return false; if (!L.module_forced) L.module = t[1]; }
    else if (name == "setmirror") { if (!need(1) || !as_b(1, &b)) return false; p->channels = b ? 1 : 2; p->mirror_input = b ? 1 : 0; }   // render.c:1054-1058: r->mirror_input AND gl->mirror_input
    else if (name == "setopacity") {
        if (!need(1)) return false;
        if (t[1] == "native") p->premultiply_alpha = 1;
        else if (t[1] == "xroot" || t[1] == "none") p->premultiply_alpha = 0;
        else { fail(GLAVA_B200_ECONFIG, "Invalid opacity option: '%s'", t[1].c_str()); return false; }   // render.c:1047-1050
    }
    else if (name == "setgeometry") { if (!need(4)) return false; p->w = as_int(3); p->h = as_int(4); }
    else if (name == "setbufsize") { if (!need(1)) return false; p->n = as_int(1); }
    else if (name == "setsamplerate") { if (!need(1)) return false; p->rate_request = as_int(1); }
    else if (name == "setsamplesize") { if (!need(1)) return false; p->samplesize_request = as_int(1); }
    else if (name == "setaccelfft") { if (!need(1) || !as_b(1, &b)) return false; p->accel_fft = b; }
    else if (name == "setavgframes") { if (!need(1)) return false; p->avg_frames = as_int(1); }
    else if (name == "setavgwindow") { if (!need(1) || !as_b(1, &b)) return false; p->avg_window = b; }
    else if (name == "setgravitystep") { if (!need(1)) return false; p->gravity_step = as_f(1); }
    else if (name == "setsmoothpass") { if (!need(1) || !as_b(1, &b)) return false; p->smooth_pass = b; }
    else if (name == "setsmoothfactor") {
        // the shaders never see the request's float: rd_new prints it into the injected header as
        // "#define _SMOOTH_FACTOR %.6f" (render.c:315-324), so the effective factor is that 6-decimal literal
        if (!need(1)) return false;
        char buf[64]; snprintf(buf, sizeof(buf), "%.6f", (double) as_f(1));
        p->smooth_factor = strtof(buf, nullptr);
    }
    else if (name == "setfftscale") { if (!need(1)) return false; p->fft_scale = as_f(1); }
    else if (name == "setfftcutoff") { if (!need(1)) return false; p->fft_cutoff = as_f(1); }
    else if (name == "setbufscale") { if (!need(1)) return false; p->bufscale = as_int(1); }          // render.c:1178
    else if (name == "setinterpolate") { if (!need(1) || !as_b(1, &b)) return false; p->interpolate = b; }   // render.c:1207
    else if (name == "setsmooth") { if (!need(1)) return false; p->smooth_distance = as_f(1); }       // render.c:1201
    else if (name == "setsmoothratio") { if (!need(1)) return false; p->smooth_ratio = as_f(1); }     // render.c:1204
    else if (name == "setframerate") { if (!need(1)) return false; if (as_int(1) > 0) p->fr = (float) as_int(1); }   // render.c:2361: pacing target
    else if (name == "transform") {
        // render.c:1218-1286.  The module fixes its own chain (fft [+gravity +avg] or window + wrange); the one
        // transform a request can add on this path is "smooth", applied after that chain.
        if (!need(2)) return false;
        static const char* chain[] = { "window", "fft", "wrange", "avg", "gravity", nullptr };
        bool known = false;
        for (int i = 0; chain[i]; ++i) if (t[2] == chain[i]) known = true;
        if (t[2] == "smooth") p->transform_smooth = 1;
        else if (!known) {
            fail(GLAVA_B200_ECONFIG, "Cannot add transformation '%s' to uniform '%s': transform function does not exist!",
                 t[2].c_str(), t[1].c_str());
            return false;
        }
    }
    else if (name == "setbg") {                                            // render.c:1062-1075
        // ext_parse_color (glsl_ext.c:88-122): two hex digits per component, optional 0x, up to 8 digits; components the
        // string does not reach keep their value (`setbg ff0000` leaves the alpha at its default 0)
        if (!need(1)) return false;
        float* comp[4] = { &p->clear_color[0], &p->clear_color[1], &p->clear_color[2], &p->clear_color[3] };
        if (!parse_hex_components(t[1].c_str(), comp)) {
            fail(GLAVA_B200_ECONFIG, "Invalid value for `setbg` request: '%s'", t[1].c_str());
            return false;
        }
    }
    else if (name == "setbgf") {                                           // render.c:1092-1099
        if (!need(4)) return false;
        for (int k = 0; k < 4; ++k) p->clear_color[k] = as_f(k + 1);
    }
    else {
        // window / desktop / pacing requests of the reference that have no meaning on this path
        static const char* ignored[] = { "setfloating", "setdecorated", "setfocused", "setmaximized", "setversion",
            "setshaderversion", "settitle", "setxwintype", "addxwinstate", "setclickthrough", "setsource", "setswap",
            "setfullscreencheck", "setprintframes", "setforcegeometry", "setforceraised",
            "setfullscreencheck", "timecycle", "settesteval", "nativeonly",
            "uniform", nullptr };
        bool known = false;
        for (int i = 0; ignored[i]; ++i) if (name == ignored[i]) known = true;
        if (!known) {
            fail(GLAVA_B200_ECONFIG, "[%s:%d] unknown request type '%s'", where, line, name.c_str());   // glsl_ext.c:299
            return false;
        }
    }
    return true;
}

// Directory context of `#include` (glsl_ext.c:161-183): plain targets are relative to `cd`; ":x" switches `cd` to
// the config dir `cfd` (only when one is set), "@x" to the defaults dir `dd` (an error when none is set).  The
// switch is sticky for the rest of the including file and inherited by the included one, as in the reference.
struct IncCtx { std::string cd, cfd, dd; bool has_cfd, has_dd; };

// ---- `#if` / `#ifdef` / `#elif` / `#else` / `#endif` / `#undef` around the `#define`s of a config file ---------------------
// GLava's own preprocessing (glsl_ext.c) does not look at these — `#request` and `#include` lines act wherever they stand —
// but the GLSL compiler that later sees the text does, so they decide which `#define` of a user's <module>.glsl counts.
// Integer expressions as the GLSL preprocessor evaluates them: literals, macros (expanded recursively, undefined -> 0),
// defined(X) / defined X, ! - + * / % < > <= >= == != && || and parentheses.
struct IfExpr {
    const char* s; const Defs* defs; int depth; bool ok;
    void ws() { while (*s && isspace((unsigned char) *s)) ++s; }
    bool eat(const char* t) { ws(); size_t n = strlen(t); if (strncmp(s, t, n) == 0) { s += n; return true; } return false; }
    long primary() {
        ws();
        if (eat("(")) { long v = lor(); if (!eat(")")) ok = false; return v; }
        if (eat("!")) return !primary();
        if (*s == '-') { ++s; return -primary(); }
        if (*s == '+') { ++s; return primary(); }
        if (isdigit((unsigned char) *s)) { char* e; long v = strtol(s, &e, 0); s = e; while (*s == 'u' || *s == 'U') ++s; return v; }
        if (isalpha((unsigned char) *s) || *s == '_') {
            std::string id; while (isalnum((unsigned char) *s) || *s == '_') id += *s++;
            if (id == "defined") {
                const bool paren = eat("("); ws();
                std::string name; while (isalnum((unsigned char) *s) || *s == '_') name += *s++;
                if (paren && !eat(")")) ok = false;
                return defs->count(name) ? 1 : 0;
            }
            auto it = defs->find(id);
            if (it == defs->end() || depth > 8) return 0;                 // an undefined identifier evaluates to 0
            const std::string body = strip_bind(it->second);
            IfExpr sub { body.c_str(), defs, depth + 1, true };
            long v = sub.lor(); sub.ws();
            if (!sub.ok || *sub.s) ok = false;                            // e.g. a float or a colour: not an #if operand
            return v;
        }
        ok = false; return 0;
    }
    long mul() { long a = primary(); for (;;) { if (eat("*")) a *= primary(); else if (eat("/")) { long b = primary(); if (!b) { ok = false; return 0; } a /= b; }
                                                 else if (eat("%")) { long b = primary(); if (!b) { ok = false; return 0; } a %= b; } else return a; } }
    long add() { long a = mul(); for (;;) { ws(); if (*s == '+') { ++s; a += mul(); } else if (*s == '-') { ++s; a -= mul(); } else return a; } }
    long rel() { long a = add(); for (;;) { if (eat("<=")) a = a <= add(); else if (eat(">=")) a = a >= add(); else if (eat("<")) a = a < add();
                                             else if (eat(">")) a = a > add(); else return a; } }
    long eq()  { long a = rel(); for (;;) { if (eat("==")) a = a == rel(); else if (eat("!=")) a = a != rel(); else return a; } }
    long land() { long a = eq(); while (eat("&&")) { long b = eq(); a = a && b; } return a; }
    long lor() { long a = land(); while (eat("||")) { long b = land(); a = a || b; } return a; }
};
struct CondFrame { bool parent, taken, active; };

// scan a config file: dispatch `#request`s, collect `#define`s (later definitions override,
// the effect of glsl_ext.c:143-159's auto-#undef), follow `#include`s
static bool scan_file(Loader& L, const std::string& path, Defs* defs, bool requests, IncCtx ctx, int depth = 0, bool optional = true) {
    std::string src;
    if (!read_file(path, &src)) {
        if (optional) return true;                // smooth_parameters.glsl / <module>.glsl need not exist in every dir
        fail(GLAVA_B200_ECONFIG, "failed to load GLSL shader source specified by #include directive '%s'", path.c_str());   // glsl_ext.c:187
        return false;
    }
    if (depth > 32) { fail(GLAVA_B200_ECONFIG, "[%s] #include nesting too deep", path.c_str()); return false; }
    src = strip_comments(src);
    std::istringstream is(src);
    std::string line; int ln = 0;
    std::vector<CondFrame> cond;                         // conditional nesting of THIS file (an #if cannot span an #include)
    auto live = [&]() { return cond.empty() || cond.back().active; };
    auto word = [](const std::string& b, const char* w) {
        const size_t n = strlen(w);
        return b.compare(0, n, w) == 0 && (b.size() == n || !(isalnum((unsigned char) b[n]) || b[n] == '_'));
    };
    auto eval_if = [&](const std::string& e, bool* out) {
        if (!defs) { *out = true; return true; }
        IfExpr x { e.c_str(), defs, 0, true };
        const long v = x.lor(); x.ws();
        if (!x.ok || *x.s) { fail(GLAVA_B200_ECONFIG, "[%s:%d] cannot evaluate '#if %s'", path.c_str(), ln, e.c_str()); return false; }
        *out = v != 0; return true;
    };
    while (std::getline(is, line)) {
        ++ln;
        std::string t = trim(line);
        if (t.empty() || t[0] != '#') continue;
        std::string body = trim(t.substr(1));
        if (word(body, "ifdef") || word(body, "ifndef") || word(body, "if")) {
            bool v = false;
            if (live()) {
                if (word(body, "if")) { if (!eval_if(trim(body.substr(2)), &v)) return false; }
                else {
                    const bool neg = word(body, "ifndef");
                    const std::string name = trim(body.substr(neg ? 6 : 5));
                    v = defs ? ((defs->count(name) != 0) != neg) : true;
                }
            }
            cond.push_back({ live(), v, live() && v });
            continue;
        }
        if (word(body, "elif") || word(body, "else")) {
            if (cond.empty()) { fail(GLAVA_B200_ECONFIG, "[%s:%d] #%s without #if", path.c_str(), ln, word(body, "else") ? "else" : "elif"); return false; }
            CondFrame& f = cond.back();
            bool v = true;
            if (word(body, "elif") && f.parent && !f.taken) { if (!eval_if(trim(body.substr(4)), &v)) return false; }
            f.active = f.parent && !f.taken && v;
            f.taken = f.taken || f.active;
            continue;
        }
        if (word(body, "endif")) {
            if (cond.empty()) { fail(GLAVA_B200_ECONFIG, "[%s:%d] #endif without #if", path.c_str(), ln); return false; }
            cond.pop_back();
            continue;
        }
        if (word(body, "undef")) { if (defs && live()) defs->erase(trim(body.substr(5))); continue; }
        if (body.compare(0, 7, "request") == 0 && requests) {
            if (!apply_request(L, tokenize(body.substr(7)), path.c_str(), ln)) return false;
        } else if (body.compare(0, 7, "include") == 0) {
            std::vector<std::string> a = tokenize(body.substr(7));
            if (a.empty()) { fail(GLAVA_B200_ECONFIG, "[%s:%d] No arguments provided to #include directive!", path.c_str(), ln); return false; }   // glsl_ext.c:163
            std::string target = a[0];
            if (!target.empty() && target[0] == ':' && ctx.has_cfd) { target = target.substr(1); ctx.cd = ctx.cfd; }
            if (!target.empty() && target[0] == '@') {
                if (!ctx.has_dd) {
                    fail(GLAVA_B200_ECONFIG, "[%s:%d] encountered '@' path specifier while no default directory is available in the current context",
                         path.c_str(), ln);                                                      // glsl_ext.c:176
                    return false;
                }
                target = target.substr(1); ctx.cd = ctx.dd;
            }
            if (!scan_file(L, ctx.cd + "/" + target, defs, requests, ctx, depth + 1, false)) return false;
        } else if (body.compare(0, 6, "define") == 0 && defs && live()) {
            std::string rest = trim(body.substr(6));
            size_t i = 0; while (i < rest.size() && (isalnum((unsigned char) rest[i]) || rest[i] == '_')) ++i;
            if (i == 0) continue;
            if (i < rest.size() && rest[i] == '(') continue;   // function-like macro: not a setting
            (*defs)[rest.substr(0, i)] = trim(rest.substr(i));
        }
    }
    return true;
}

static bool file_exists(const std::string& p) { std::ifstream f(p.c_str()); return (bool) f; }

static void apply_defines(glava_b200_params* p, const Defs& d) {
    // smooth_parameters.glsl
    auto it = d.find("ROUND_FORMULA");
    if (it != d.end()) {
        if (it->second == "sinusoidal") p->round_formula = 0; else if (it->second == "linear") p->round_formula = 1;
        else if (it->second == "circular") p->round_formula = 2;
        else fail(GLAVA_B200_ECONFIG, "unknown ROUND_FORMULA '%s'", it->second.c_str());
    }
    it = d.find("SAMPLE_MODE");
    if (it != d.end()) {
        if (it->second == "average") p->sample_mode = 0; else if (it->second == "maximum") p->sample_mode = 1;
        else if (it->second == "hybrid") p->sample_mode = 2;
        else fail(GLAVA_B200_ECONFIG, "unknown SAMPLE_MODE '%s'", it->second.c_str());
    }
    getf(d, "SAMPLE_SCALE", &p->sample_scale); getf(d, "SAMPLE_RANGE", &p->sample_range);
    getf(d, "SAMPLE_HYBRID_WEIGHT", &p->hybrid_weight);
    Num n;
    switch (p->module) {
        case GLAVA_B200_MOD_BARS:
            getf(d, "BAR_WIDTH", &p->bars_width); getf(d, "BAR_GAP", &p->bars_gap);
            getppf(d, "BAR_OUTLINE_WIDTH", "bars/1.frag:116,127", &p->bars_outline_width); getf(d, "AMPLIFY", &p->bars_amplify);
            parse_color_macro(d, "COLOR", &p->bars_color, &p->bars_color_prog, { "d" });
            if (p->bars_color.mode == 0 && eval_num(d, "GRADIENT", &n)) p->bars_color.gradient = (float) n.v;
            if ((it = d.find("BAR_OUTLINE")) != d.end()) {
                std::string v = strip_bind(it->second);
                std::string squeezed; for (char c : v) if (!isspace((unsigned char) c)) squeezed += c;
                if (squeezed == "vec4(COLOR.rgb*1.5,COLOR.a)") p->bars_outline_mode = 0;
                else if (parse_const_color(d, v, p->bars_outline)) p->bars_outline_mode = 1;
                else {                                       // any other expression of d (COLOR expands textually inside it)
                    bool uses_var = false;
                    p->bars_outline_prog.n_ops = 0;
                    if (compile_color_text(d, "BAR_OUTLINE", v, { "d" }, &p->bars_outline_prog, &uses_var)) {
                        if (uses_var) p->bars_outline_mode = 2;
                        else {
                            const f4 r = eval_color_prog(p->bars_outline_prog, 0.0f);
                            p->bars_outline[0] = r.r; p->bars_outline[1] = r.g; p->bars_outline[2] = r.b; p->bars_outline[3] = r.a;
                            p->bars_outline_mode = 1; p->bars_outline_prog.n_ops = 0;
                        }
                    }
                }
            }
            getppi(d, "DIRECTION", "bars/1.frag:89,101", &p->bars_direction); getppi(d, "INVERT", "bars/1.frag:53,94,106", &p->bars_invert);
            getppi(d, "FLIP", "bars/1.frag:59", &p->bars_flip); getppi(d, "MIRROR_YX", "bars/1.frag:38", &p->bars_mirror_yx);
            { int dm = 0; getppi(d, "DISABLE_MONO", "bars/1.frag:32", &dm); if (dm == 1) p->channels = 2; }          // bars/1.frag:32-34
            // USE_ALPHA (bars.glsl:16) changes nothing in the reference: bars/2.frag tests `#if USE_ALPHA == 0` WITHOUT including
            // bars.glsl, the undefined macro evaluates as 0 and the premultiply stage is always disabled (pinned by the llvmpipe
            // golden `bars_use_alpha`: a translucent COLOR with USE_ALPHA 1 renders un-premultiplied).  Read and ignored.
            break;
        case GLAVA_B200_MOD_RADIAL:
            getf(d, "C_RADIUS", &p->radial_radius);
            if (eval_num(d, "C_LINE", &n)) {
                p->radial_line = (float) n.v;
                p->radial_line_half = n.is_int ? (float) ((long) n.v / 2) : (float) n.v / 2.0f;   // radial/1.frag:52 (C_LINE / 2)
            }
            parse_plain_color(d, "OUTLINE", p->radial_outline);
            geti(d, "NBARS", &p->radial_nbars);
            if (eval_num(d, "BAR_WIDTH", &n)) { p->radial_bar_width = (float) n.v; p->radial_bar_width_int = n.is_int ? 1 : 0; }
            getf(d, "AMPLIFY", &p->radial_amplify);
            parse_color_macro(d, "COLOR", &p->radial_color, &p->radial_color_prog, { "d" });
            if (p->radial_color.mode == 0 && eval_num(d, "GRADIENT", &n)) p->radial_color.gradient = (float) n.v;
            getf(d, "ROTATE", &p->radial_rotate); getppi(d, "INVERT", "radial/1.frag:67", &p->radial_invert);
            getf(d, "BAR_ALIAS_FACTOR", &p->radial_bar_alias); getf(d, "C_ALIAS_FACTOR", &p->radial_c_alias);
            getf(d, "CENTER_OFFSET_X", &p->radial_off_x); getf(d, "CENTER_OFFSET_Y", &p->radial_off_y);
            getppf(d, "BAR_OUTLINE_WIDTH", "radial/1.frag:87,101", &p->radial_bar_outline_width);   // deprecated (radial.glsl:33-36)
            memcpy(p->radial_bar_outline, p->radial_outline, sizeof(p->radial_bar_outline));   // `#define BAR_OUTLINE OUTLINE`
            parse_plain_color(d, "BAR_OUTLINE", p->radial_bar_outline);
            break;
        case GLAVA_B200_MOD_CIRCLE:
            getf(d, "C_RADIUS", &p->circle_radius); getf(d, "C_LINE", &p->circle_line);
            parse_plain_color(d, "OUTLINE", p->circle_outline);
            getf(d, "AMPLIFY", &p->circle_amplify); getf(d, "ROTATE", &p->circle_rotate);
            geti(d, "INVERT", &p->circle_invert); getppi(d, "C_FILL", "circle/1.frag:75", &p->circle_fill); getppi(d, "C_SMOOTH", "circle/2.frag:14", &p->circle_smooth);
            break;
        case GLAVA_B200_MOD_GRAPH:
            getf(d, "VSCALE", &p->graph_vscale); getppi(d, "DIRECTION", "graph/1.frag:67", &p->graph_direction);
            parse_color_macro(d, "COLOR", &p->graph_color, &p->graph_color_prog, { "pos", "d" });
            if (p->graph_color.mode == 0 && eval_num(d, "GRADIENT", &n)) p->graph_color.gradient = (float) n.v;
            getppi(d, "DRAW_OUTLINE", "graph/2.frag:12,34", &p->graph_draw_outline); getppi(d, "DRAW_HIGHLIGHT", "graph/2.frag:12,39", &p->graph_draw_highlight);
            parse_plain_color(d, "OUTLINE", p->graph_outline); getppi(d, "INVERT", "graph/1.frag:110", &p->graph_invert);
            getppi(d, "ANTI_ALIAS", "graph/3.frag:15", &p->graph_anti_alias);
            geti(d, "JOIN_CHANNELS", &p->graph_join_channels);
            break;
        case GLAVA_B200_MOD_WAVE:
            getf(d, "MIN_THICKNESS", &p->wave_min_thickness); getf(d, "MAX_THICKNESS", &p->wave_max_thickness);
            parse_plain_color(d, "BASE_COLOR", p->wave_base_color); getf(d, "AMPLIFY", &p->wave_amplify);
            parse_plain_color(d, "OUTLINE", p->wave_outline);
            break;
        default: break;
    }
}

int load_config(glava_b200_params* out, const char* const* paths, const char* entry,
                const char* const* requests, const char* force_module, const char* const* binds) {
    clear_error();
    std::map<std::string, std::string> bind_map;
    if (binds) for (int i = 0; binds[i]; ++i) {           // "name=value" (value: #rrggbb[aa] or vec4(...))
        std::string b = binds[i];
        size_t eq = b.find('=');
        if (eq == std::string::npos) { fail(GLAVA_B200_ECONFIG, "bind '%s': expected name=value", binds[i]); return GLAVA_B200_ECONFIG; }
        bind_map[trim_copy(b.substr(0, eq))] = trim_copy(b.substr(eq + 1));
    }
    struct BindScope { BindScope(const std::map<std::string, std::string>* m) { g_binds = m; } ~BindScope() { g_binds = nullptr; } } scope(&bind_map);
    fill_defaults(out, GLAVA_B200_MOD_BARS);
    if (paths && paths[0]) {
        // With a configuration directory the starting point is rd_new's own initialisers (render.c:876-934), not the
        // values the shipped rc.glsl requests: a user rc.glsl that leaves a request out gets THESE (8192 samples at
        // 22000 Hz, 6 average frames, interpolation ON, a 500 x 400 window ...).  paths == NULL keeps the shipped set.
        out->n = 8192; out->rate_request = 22000; out->samplesize_request = 1024;
        out->avg_frames = 6; out->avg_window = 1; out->gravity_step = 4.2f; out->interpolate = 1;
        out->smooth_factor = 0.025f; out->smooth_distance = 0.01f; out->smooth_ratio = 4.0f;
        out->premultiply_alpha = 1; out->accel_fft = 1; out->smooth_pass = 1; out->fft_scale = 10.2f; out->fft_cutoff = 0.3f;
        out->w = 500; out->h = 400; out->bufscale = 1; out->channels = 2; out->mirror_input = 0;
        out->clear_color[0] = out->clear_color[1] = out->clear_color[2] = out->clear_color[3] = 0.0f;
    }
    Loader L { out, "bars", false, false };
    if (force_module) { L.module = force_module; L.module_forced = true; }
    std::vector<std::string> dirs;
    if (paths) for (int i = 0; paths[i]; ++i) dirs.push_back(paths[i]);
    if (!entry) entry = "rc.glsl";
    // entry: first directory that has it (render.c:1322-1413)
    std::string entry_dir;
    for (auto& d : dirs) if (file_exists(d + "/" + entry)) { entry_dir = d; break; }
    if (!dirs.empty() && entry_dir.empty()) {
        fail(GLAVA_B200_ECONFIG, "Could not find entry point '%s' in any of the configuration paths", entry);
        return GLAVA_B200_ECONFIG;
    }
    // rc.glsl: cd = the directory it was found in, no config / defaults dir (render.c:1356-1361)
    IncCtx rc_ctx { entry_dir, "", "", false, false };
    if (!entry_dir.empty() && !scan_file(L, entry_dir + "/" + entry, nullptr, true, rc_ctx)) return GLAVA_B200_ECONFIG;
    if (requests) {
        for (int i = 0; requests[i]; ++i) {                                   // render.c:1415-1435
            std::string r = requests[i];
            if (!apply_request(L, tokenize(r), "--request", i + 1)) return GLAVA_B200_ECONFIG;
        }
    }
    const int pre_module_smooth_pass = out->smooth_pass;   // what the module's stage-1 header will say (see below)
    int mod = module_from_name(L.module.c_str());
    if (mod < 0) {
        fail(GLAVA_B200_ECONFIG, "Could not find module '%s' (B200 path implements: bars radial circle graph wave test)", L.module.c_str());
        return GLAVA_B200_ECONFIG;
    }
    out->module = mod;
    // The module's shaders do `#include "@x.glsl"` then `#include ":x.glsl"` (e.g. bars/1.frag:9-10, util/smooth.glsl:6-7):
    // the defaults dir `dd` = LAST path (render.c:1327) first, then the config dir = where the entry was found
    // (render.c:1472-1476, shaderbuild(gl, shaders, data, dd, ...)), whose definitions win.
    Defs defs;
    if (!dirs.empty()) {
        const std::string dd = dirs.back();
        IncCtx dctx { dd, entry_dir, dd, true, true };
        IncCtx cctx { entry_dir, entry_dir, dd, true, true };
        for (const std::string& f : { std::string("smooth_parameters.glsl"), L.module + ".glsl" }) {
            // smooth_parameters.glsl's `#request`s only count when a MODULE shader includes util/smooth.glsl (all but
            // wave/1.frag do): util/smooth_pass.frag includes it too, but rd_new ignores them while loading that shader
            // (`loading_smooth_pass`, render.c:1186-1215, 1634-1642).  Its #defines reach the smoothing pass either way.
            const bool reqs = !(f == "smooth_parameters.glsl" && L.module == "wave");
            if (!scan_file(L, dd + "/" + f, &defs, reqs, dctx)) return GLAVA_B200_ECONFIG;
            if (!scan_file(L, entry_dir + "/" + f, &defs, reqs, cctx)) return GLAVA_B200_ECONFIG;
        }
    }
    else if (!bind_map.empty()) {
        // no config directory: the shipped module configs still carry their `@fg:` / `@bg:` macros, so a `--pipe` bind
        // reaches them (bars.glsl:18-22, radial.glsl:7,15-17, circle.glsl:6, graph.glsl:8-11,21, wave.glsl:6,10)
        switch (mod) {
            case GLAVA_B200_MOD_BARS:
                defs["GRADIENT"] = "80"; defs["COLOR"] = "@fg:mix(#3366b2, #a0a0b2, clamp(d / GRADIENT, 0, 1))";
                defs["BAR_OUTLINE"] = "@bg:vec4(COLOR.rgb * 1.5, COLOR.a)"; break;
            case GLAVA_B200_MOD_RADIAL:
                defs["GRADIENT"] = "95"; defs["COLOR"] = "@fg:mix(#cc3333, #cca0a0, clamp(d / GRADIENT, 0, 1))";
                defs["OUTLINE"] = "@bg:#333333"; break;
            case GLAVA_B200_MOD_CIRCLE: defs["OUTLINE"] = "@fg:#333333"; break;
            case GLAVA_B200_MOD_GRAPH:
                defs["GRADIENT"] = "75"; defs["COLOR"] = "@fg:mix(#802A2A, #4F4F92, clamp(pos / GRADIENT, 0, 1))";
                defs["OUTLINE"] = "@bg:#262626"; break;
            case GLAVA_B200_MOD_WAVE:
                defs["BASE_COLOR"] = "@fg:vec4(0.7, 0.2, 0.45, 1)"; defs["OUTLINE"] = "@bg:vec4(0.15, 0.15, 0.15, 1)"; break;
            default: break;
        }
    }
    // (CLI requests are NOT applied again: in the reference they run once, between rc.glsl and the module load
    // (render.c:1415-1435), so a request that smooth_parameters.glsl also makes — setsmoothfactor, setavgframes,
    // setsmoothpass, setfftscale ... — is overridden by it for every module but wave.  Checked against rd_new.)
    // The module's first shader is compiled with a header built BEFORE its own includes ran their requests
    // (shaderload: EBIND list at render.c:284-293, ext_process at :312): its `_PRE_SMOOTHED_AUDIO` is smooth_pass as of
    // the end of rc.glsl + CLI, while the K5 pass follows the final smooth_pass.
    if (!dirs.empty() && pre_module_smooth_pass != out->smooth_pass) out->shader_pre_smoothed = pre_module_smooth_pass ? 1 : 2;
    apply_defines(out, defs);
    if (has_error()) return GLAVA_B200_ECONFIG;
    out->ur = (float) out->rate_request / (float) (out->samplesize_request / 4);
    return validate_params(out);
}

int validate_params(const glava_b200_params* p) {
    auto bad = [&](const char* what) { fail(GLAVA_B200_EINVAL, "invalid parameter: %s", what); return GLAVA_B200_EINVAL; };
    if (p->n < 256 || p->n > 16384 || (p->n & (p->n - 1))) return bad("setbufsize must be a power of two in [256, 16384]");
    if (p->avg_frames < 1 || p->avg_frames > 16) return bad("setavgframes must be in [1, 16]");
    if (p->w < 4 || p->h < 1 || p->w > 16384 || p->h > 16384) return bad("geometry");
    if (p->module < 0 || p->module > GLAVA_B200_MOD_TEST) return bad("module");
    if (!(p->ur > 0.0f)) return bad("ur must be > 0");
    if (p->channels != 1 && p->channels != 2) return bad("channels");
    if (p->sample_mode < 0 || p->sample_mode > 2 || p->round_formula < 0 || p->round_formula > 2) return bad("sample mode / round formula");
    if (p->radial_nbars < 2) return bad("NBARS");
    if (p->shader_pre_smoothed < 0 || p->shader_pre_smoothed > 2 || (p->shader_pre_smoothed == 1 && p->smooth_pass) ||
        (p->shader_pre_smoothed == 2 && !p->smooth_pass)) return bad("shader_pre_smoothed contradicts smooth_pass");
    const glava_b200_color* cols[3] = { &p->bars_color, &p->radial_color, &p->graph_color };
    const glava_b200_color_prog* progs[4] = { &p->bars_color_prog, &p->radial_color_prog, &p->graph_color_prog, &p->bars_outline_prog };
    for (int i = 0; i < 3; ++i) {
        if (cols[i]->mode < 0 || cols[i]->mode > 2) return bad("colour mode");
        if (cols[i]->mode == 2 && progs[i]->n_ops < 1) return bad("colour mode 2 without a compiled expression");
    }
    if (p->bars_outline_mode < 0 || p->bars_outline_mode > 2 || (p->bars_outline_mode == 2 && p->bars_outline_prog.n_ops < 1)) return bad("bars_outline_mode");
    for (int i = 0; i < 4; ++i) {
        if (progs[i]->n_ops < 0 || progs[i]->n_ops > GLAVA_B200_COLOR_OPS || progs[i]->result < 0 || progs[i]->result >= GLAVA_B200_COLOR_REGS)
            return bad("colour program size");
        for (int k = 0; k < progs[i]->n_ops; ++k) {
            const glava_b200_color_op& o = progs[i]->ops[k];
            if (o.op >= GLAVA_B200_COP_COUNT || o.dst >= GLAVA_B200_COLOR_REGS || o.a >= GLAVA_B200_COLOR_REGS || o.b >= GLAVA_B200_COLOR_REGS)
                return bad("colour program instruction");
        }
    }
    if (!(p->bars_width + p->bars_gap > 0.0f)) return bad("BAR_WIDTH + BAR_GAP");
    if (p->bufscale < 1 || p->n % p->bufscale || p->n / p->bufscale < 256 || ((p->n / p->bufscale) & (p->n / p->bufscale - 1)))
        return bad("setbufsize / setbufscale must be a power of two >= 256");
    if (p->fr < 0.0f) return bad("fr must be >= 0");
    if (p->transform_smooth < 0 || p->transform_smooth > 2) return bad("transform_smooth must be 0, 1 (after \"fft\") or 2 (before \"fft\")");
    if (p->transform_smooth && !(p->smooth_ratio > 0.0f)) return bad("setsmoothratio must be > 0");
    return GLAVA_B200_OK;
}

}  // namespace glb
