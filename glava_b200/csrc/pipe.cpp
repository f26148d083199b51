// pipe.cpp — `--pipe NAME[:TYPE]` binds driven by text lines (host only).
//
//   argument parsing / validation      glava/glava.c:338-411
//   line parser `name = value`         glava/render.c:1861-1936 (prefix match of the name, first bind as the default target)
//   typed value parsers                glava/render.c:1938-1992 (+ ext_parse_color, glava/glsl_ext.c:88-122)
//   the uniform write                  glava/render.c:2071-2100 -> here: re-evaluation of the config with the binds' current
//                                      values and glava_b200_reconfigure
//
// A bound `@name:default` macro reads the uniform `_IN_name` (glsl_ext.c:571-576); a GL uniform nobody has written yet is
// zero, so a bind holds 0 / vec4(0) until its first line arrives — kept.
#include "internal.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace glb;

namespace {
enum { T_NONE = 0, T_INT = 1, T_FLOAT = 2, T_BOOL = 3, T_VEC2 = 4, T_VEC3 = 5, T_VEC4 = 6 };   // render.h:32-38
const char* const kTypeNames[] = { "NONE", "int", "float", "bool", "vec2", "vec3", "vec4" };   // render.c:24-33

struct Bind {
    std::string name;
    int type;
    bool b; int i; float f[4];      // value the uniform holds
};

}  // namespace

struct glava_b200_pipe {
    std::vector<std::string> paths, requests;
    std::string entry, module;
    bool has_paths, has_entry, has_module;
    std::vector<Bind> binds;
    std::string line;              // bytes of the current, still unterminated line
    bool overlong;
    bool dirty;                    // a bind changed since the last glava_b200_pipe_params / _apply
    union { bool b; int i; float f[4]; } parsed;   // render.c:1855-1859: persists across lines (sscanf may fill only a prefix)
};

static int parse_pipe_arg(const char* arg, std::vector<Bind>* binds) {
    std::string a = arg ? arg : "_";                                     // PIPE_DEFAULT (render.h:40)
    size_t sp = a.find(' ');
    if (sp != std::string::npos) a.resize(sp);
    std::string name = a, type;
    size_t sep = a.rfind(':');
    if (sep != std::string::npos) { name = a.substr(0, sep); type = a.substr(sep + 1); }
    if (name.empty())
        return fail(GLAVA_B200_ECONFIG, "Error: invalid pipe binding name: \"%s\"\nZero length names are not permitted.", name.c_str());
    for (size_t k = 0; k < name.size(); ++k) {
        const char c = name[k];
        const bool digit = c >= '0' && c <= '9';
        if (digit && k == 0)
            return fail(GLAVA_B200_ECONFIG, "Error: invalid pipe binding name: \"%s\" ('%c')\nValid names may not start with a number.", name.c_str(), c);
        if (!(digit || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'))
            return fail(GLAVA_B200_ECONFIG, "Error: invalid pipe binding name: \"%s\" ('%c')\nValid names may only contain [a..z], [A..Z], [0..9] "
                        "and '_' characters.", name.c_str(), c);
    }
    for (const Bind& b : *binds)
        if (b.name == name) return fail(GLAVA_B200_ECONFIG, "Error: attempted to re-bind pipe argument: \"%s\"", name.c_str());
    int t = -1;
    if (type.empty()) t = T_VEC4;
    else for (int k = 0; k <= T_VEC4; ++k) if (type == kTypeNames[k]) { t = k; break; }
    if (t == -1) return fail(GLAVA_B200_ECONFIG, "Error: Unsupported `--pipe` GLSL type: \"%s\"", type.c_str());
    Bind b; b.name = name; b.type = t; b.b = false; b.i = 0; b.f[0] = b.f[1] = b.f[2] = b.f[3] = 0.0f;
    binds->push_back(b);
    return 0;
}

extern "C" glava_b200_pipe* glava_b200_pipe_new(const char* const* paths, const char* entry, const char* const* requests,
                                                 const char* force_module, const char* const* pipe_args) {
    clear_error();
    glava_b200_pipe* p = new glava_b200_pipe();
    p->has_paths = paths != nullptr; p->has_entry = entry != nullptr; p->has_module = force_module != nullptr;
    if (paths) for (int i = 0; paths[i]; ++i) p->paths.push_back(paths[i]);
    if (requests) for (int i = 0; requests[i]; ++i) p->requests.push_back(requests[i]);
    if (entry) p->entry = entry;
    if (force_module) p->module = force_module;
    p->overlong = false; p->dirty = true;
    memset(&p->parsed, 0, sizeof(p->parsed));
    if (pipe_args) for (int i = 0; pipe_args[i]; ++i)
        if (parse_pipe_arg(pipe_args[i], &p->binds) != 0) { delete p; return nullptr; }
    return p;
}

extern "C" void glava_b200_pipe_free(glava_b200_pipe* p) { delete p; }

// one complete line (no terminator).  Returns 1 when a bind took a new value.
static int pipe_line(glava_b200_pipe* p, std::string text) {
    if (text.empty() || p->binds.empty()) return 0;                      // render.c:1876-1877
    size_t a = 0; while (a < text.size() && text[a] == ' ') ++a;         // :1885
    const std::string s = text.substr(a);
    // `name = value`: the first '=' ends the name (trailing spaces dropped), the value starts at the next non-space
    // character and loses its trailing spaces (:1886-1903)
    std::string name, value;
    bool saw_eq = false, valid = false;
    const size_t eq = s.find('=');
    if (eq != std::string::npos) {
        saw_eq = true;
        size_t e = eq; while (e > 0 && s[e - 1] == ' ') --e;
        name = s.substr(0, e);
        size_t v = eq + 1; while (v < s.size() && s[v] == ' ') ++v;
        if (v < s.size()) {
            size_t w = s.size(); while (w > v && s[w - 1] == ' ') --w;
            value = s.substr(v, w - v);
            valid = true;
        }
    }
    if (!saw_eq) { name.clear(); value = s; valid = true; }              // no assignment: a value for the default bind (:1905-1910)
    if (!valid) { fail(GLAVA_B200_ECONFIG, "Bad assignment format for \"%s\"", s.c_str()); return 0; }
    // the reference compares only the typed prefix: strncmp(bind, name, len(name)) — an empty name selects the first bind
    Bind* bd = nullptr;
    for (Bind& b : p->binds) if (b.name.compare(0, name.size(), name) == 0) { bd = &b; break; }
    if (!bd) { fail(GLAVA_B200_ECONFIG, "Variable name not bound: \"%s\"", name.c_str()); return 0; }

    const char* v = value.c_str();
    bool ready = false;
    switch (bd->type) {
        case T_BOOL:
            if (!strcmp("true", v) || !strcmp("TRUE", v) || !strcmp("True", v) || !strcmp("1", v)) { p->parsed.b = true; ready = true; }
            else if (!strcmp("false", v) || !strcmp("FALSE", v) || !strcmp("False", v) || !strcmp("0", v)) { p->parsed.b = false; ready = true; }
            else fail(GLAVA_B200_ECONFIG, "Bad format for boolean: \"%s\"", v);
            break;
        case T_INT:
            errno = 0; p->parsed.i = (int) strtol(v, nullptr, 10);
            ready = errno != ERANGE;
            break;
        case T_FLOAT:
            errno = 0; p->parsed.f[0] = strtof(v, nullptr);
            ready = errno != ERANGE;
            break;
        case T_VEC2: ready = EOF != sscanf(v, "%f,%f", &p->parsed.f[0], &p->parsed.f[1]); break;
        case T_VEC3: ready = EOF != sscanf(v, "%f,%f,%f", &p->parsed.f[0], &p->parsed.f[1], &p->parsed.f[2]); break;
        case T_VEC4:
            if (v[0] == '#') {
                p->parsed.f[0] = p->parsed.f[1] = p->parsed.f[2] = 0.0f; p->parsed.f[3] = 1.0f;
                float* ptrs[4] = { &p->parsed.f[0], &p->parsed.f[1], &p->parsed.f[2], &p->parsed.f[3] };
                if (parse_hex_components(v + 1, ptrs)) ready = true;
                else fail(GLAVA_B200_ECONFIG, "Bad format for color string: \"%s\"", v);
            } else ready = EOF != sscanf(v, "%f,%f,%f,%f", &p->parsed.f[0], &p->parsed.f[1], &p->parsed.f[2], &p->parsed.f[3]);
            break;
        default: break;
    }
    if (!ready) return 0;
    switch (bd->type) {                                                   // the glUniform* of render.c:2071-2099
        case T_BOOL: bd->b = p->parsed.b; break;
        case T_INT:  bd->i = p->parsed.i; break;
        case T_FLOAT: bd->f[0] = p->parsed.f[0]; break;
        default: for (int k = 0; k < bd->type - T_VEC2 + 2; ++k) bd->f[k] = p->parsed.f[k]; break;
    }
    p->dirty = true;
    return 1;
}

extern "C" int glava_b200_pipe_feed(glava_b200_pipe* p, const char* bytes, size_t len) {
    clear_error();
    if (!p || (!bytes && len)) return fail(GLAVA_B200_EINVAL, "glava_b200_pipe_feed: null argument");
    int updated = 0;
    for (size_t k = 0; k < len; ++k) {
        const char c = bytes[k];
        if (c != '\n') {
            // the reference's line buffer holds 127 characters (render.c:1846,1873); longer lines wedge its parser —
            // here they are reported and dropped
            if (p->line.size() >= 127) p->overlong = true; else p->line += c;
            continue;
        }
        if (p->overlong) fail(GLAVA_B200_ECONFIG, "pipe: line longer than 127 characters dropped");
        else updated += pipe_line(p, p->line);
        p->line.clear(); p->overlong = false;
    }
    return updated;
}

static std::string bind_value(const Bind& b) {
    char buf[160];
    switch (b.type) {
        case T_BOOL:  snprintf(buf, sizeof buf, "%d", b.b ? 1 : 0); break;
        case T_INT:   snprintf(buf, sizeof buf, "%d", b.i); break;
        case T_FLOAT: snprintf(buf, sizeof buf, "%.9g", (double) b.f[0]); break;
        case T_VEC2:  snprintf(buf, sizeof buf, "vec2(%.9g, %.9g)", (double) b.f[0], (double) b.f[1]); break;
        case T_VEC3:  snprintf(buf, sizeof buf, "vec3(%.9g, %.9g, %.9g)", (double) b.f[0], (double) b.f[1], (double) b.f[2]); break;
        default:      snprintf(buf, sizeof buf, "vec4(%.9g, %.9g, %.9g, %.9g)", (double) b.f[0], (double) b.f[1], (double) b.f[2], (double) b.f[3]); break;
    }
    return buf;
}

extern "C" int glava_b200_pipe_params(glava_b200_pipe* p, glava_b200_params* out) {
    clear_error();
    if (!p || !out) return fail(GLAVA_B200_EINVAL, "glava_b200_pipe_params: null argument");
    std::vector<const char*> paths, requests, binds;
    std::vector<std::string> bind_strs;
    for (const std::string& s : p->paths) paths.push_back(s.c_str());
    for (const std::string& s : p->requests) requests.push_back(s.c_str());
    for (const Bind& b : p->binds) bind_strs.push_back(b.name + "=" + bind_value(b));
    for (const std::string& s : bind_strs) binds.push_back(s.c_str());
    paths.push_back(nullptr); requests.push_back(nullptr); binds.push_back(nullptr);
    const int rc = load_config(out, p->has_paths ? paths.data() : nullptr, p->has_entry ? p->entry.c_str() : nullptr,
                               requests.data(), p->has_module ? p->module.c_str() : nullptr, binds.data());
    if (rc == 0) p->dirty = false;
    return rc;
}

extern "C" int glava_b200_pipe_bind_count(const glava_b200_pipe* p) { return p ? (int) p->binds.size() : 0; }

extern "C" int glava_b200_pipe_bind(const glava_b200_pipe* p, int index, const char** name, const char** type, float value[4]) {
    if (!p || index < 0 || index >= (int) p->binds.size()) return fail(GLAVA_B200_EINVAL, "glava_b200_pipe_bind: index out of range");
    const Bind& b = p->binds[index];
    if (name) *name = b.name.c_str();
    if (type) *type = kTypeNames[b.type];
    if (value) {
        for (int k = 0; k < 4; ++k) value[k] = b.f[k];
        if (b.type == T_BOOL) value[0] = b.b ? 1.0f : 0.0f;
        if (b.type == T_INT) value[0] = (float) b.i;
    }
    return 0;
}

extern "C" int glava_b200_pipe_apply(glava_b200_pipe* p, glava_b200* r) {
    clear_error();
    if (!p || !r) return fail(GLAVA_B200_EINVAL, "glava_b200_pipe_apply: null argument");
    if (!p->dirty) return 0;
    glava_b200_params cur, next;
    int rc = glava_b200_get_params(r, &cur);
    if (rc) return rc;
    rc = glava_b200_pipe_params(p, &next);
    if (rc) return rc;
    // what sizes device state is the renderer's: geometry may have been changed by glava_b200_sizereq, engine knobs are not
    // part of any config file
    next.w = cur.w; next.h = cur.h; next.fb_slots = cur.fb_slots; next.lazy_smooth = cur.lazy_smooth; next.ur = cur.ur; next.fr = cur.fr;
    return glava_b200_reconfigure(r, &next);
}
