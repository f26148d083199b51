// color_compile.h — host-side compiler: GLSL colour expression text -> glava_b200_color_prog (included by config.cpp).
//
// The reference pastes `COLOR` / `BAR_OUTLINE` into the fragment shader (bars/1.frag:118-129, radial/1.frag:89-93,
// graph/1.frag:117), so a user's macro may be any GLSL expression of the variables in scope there.  This path has no GLSL
// compiler: the expression is compiled once, when the config is read, into the straight-line program raster_core.h's
// eval_color_prog() runs (on the device when the row-colour tables / the polar geometry cache are built, per pixel in
// the generic kernels).  Supported: the per-pixel variable (`d`; `pos` in graph), int / float literals, `#rrggbb[aa]`,
// other object-like config macros (expanded textually, as the preprocessor does), + - * / and unary minus, swizzles,
// vec2 / vec3 / vec4 / float constructors, comparisons, && || ! and ?: on scalars, `true` / `false`, and mix clamp smoothstep
// min max mod step abs floor ceil fract sqrt sin cos log log2 exp exp2 pow tan atan sign.  Integer sub-expressions are folded with C semantics (`1 / 2` is 0) and convert to float where they meet one.
// Anything else is a config error naming the offending token.
#ifndef GLAVA_B200_COLOR_COMPILE_H
#define GLAVA_B200_COLOR_COMPILE_H

namespace glb {

struct ColorCompiler {
    const Defs* defs;
    std::vector<std::string> vars;       // names of the per-pixel variable
    std::string err;
    std::vector<glava_b200_color_op> ops;
    bool used[GLAVA_B200_COLOR_REGS];
    bool uses_var;

    struct Tok { int kind; std::string text; };          // kind: 'n' number, 'i' identifier, 'h' #hex, 'p' punctuation, 'e' end
    std::vector<Tok> toks;
    size_t pos;

    struct Val { int reg; int width; bool is_int; long ival; };   // reg < 0: compile-time integer constant

    ColorCompiler(const Defs* d, std::vector<std::string> v) : defs(d), vars(std::move(v)), uses_var(false), pos(0) {
        for (bool& u : used) u = false;
    }

    bool fail_at(const std::string& m) { if (err.empty()) err = m; return false; }

    // ---- lexer + textual macro expansion ---------------------------------------------------------------------------
    bool lex(const std::string& s, std::vector<Tok>* out, int depth, std::vector<std::string>* active) {
        if (depth > 16) return fail_at("macro expansion too deep");
        size_t i = 0;
        while (i < s.size()) {
            const char c = s[i];
            if (isspace((unsigned char) c)) { ++i; continue; }
            if (isdigit((unsigned char) c) || (c == '.' && i + 1 < s.size() && isdigit((unsigned char) s[i + 1]))) {
                size_t j = i;
                const bool hex = c == '0' && i + 1 < s.size() && (s[i + 1] == 'x' || s[i + 1] == 'X');
                while (j < s.size()) {
                    const char ch = s[j];
                    if (isalnum((unsigned char) ch) || ch == '.') { ++j; continue; }
                    if ((ch == '+' || ch == '-') && !hex && j > i && (s[j - 1] == 'e' || s[j - 1] == 'E')) { ++j; continue; }
                    break;
                }
                out->push_back({ 'n', s.substr(i, j - i) }); i = j; continue;
            }
            if (isalpha((unsigned char) c) || c == '_') {
                size_t j = i; while (j < s.size() && (isalnum((unsigned char) s[j]) || s[j] == '_')) ++j;
                const std::string id = s.substr(i, j - i);
                i = j;
                const bool after_dot = !out->empty() && out->back().kind == 'p' && out->back().text == ".";
                bool is_var = false; for (const std::string& v : vars) is_var |= v == id;
                auto it = defs->find(id);
                bool busy = false; for (const std::string& a : *active) busy |= a == id;
                if (!after_dot && !is_var && it != defs->end() && !busy) {
                    active->push_back(id);
                    const bool ok = lex(strip_bind(it->second), out, depth + 1, active);
                    active->pop_back();
                    if (!ok) return false;
                } else out->push_back({ 'i', id });
                continue;
            }
            if (c == '#') {
                size_t j = i + 1; while (j < s.size() && isxdigit((unsigned char) s[j]) && j - i - 1 < 8) ++j;
                out->push_back({ 'h', s.substr(i, j - i) }); i = j; continue;
            }
            if (i + 1 < s.size()) {
                const std::string two = s.substr(i, 2);
                if (two == "<=" || two == ">=" || two == "==" || two == "!=" || two == "&&" || two == "||") { out->push_back({ 'p', two }); i += 2; continue; }
            }
            if (strchr("()+-*/,.<>!?:", c)) { out->push_back({ 'p', std::string(1, c) }); ++i; continue; }
            return fail_at(std::string("unexpected character '") + c + "'");
        }
        return true;
    }
    const Tok& peek() const { static const Tok end = { 'e', "" }; return pos < toks.size() ? toks[pos] : end; }
    bool is_p(const char* t) const { return peek().kind == 'p' && peek().text == t; }
    bool eat(const char* t) { if (is_p(t)) { ++pos; return true; } return false; }

    // ---- registers / emission --------------------------------------------------------------------------------------
    int alloc() {
        for (int r = 0; r < GLAVA_B200_COLOR_REGS; ++r) if (!used[r]) { used[r] = true; return r; }
        fail_at("expression needs more than 8 live intermediate values");
        return -1;
    }
    void release(const Val& v) { if (v.reg >= 0) used[v.reg] = false; }
    bool emit(int op, int dst, int a, int b, float imm) {
        if (ops.size() >= GLAVA_B200_COLOR_OPS) return fail_at("expression longer than 64 operations");
        glava_b200_color_op o; o.op = (uint8_t) op; o.dst = (uint8_t) dst; o.a = (uint8_t) a; o.b = (uint8_t) b; o.imm = imm;
        ops.push_back(o);
        return true;
    }
    // integer constant -> float register (GLSL implicit int -> float conversion)
    bool materialise(Val* v) {
        if (v->reg >= 0) return true;
        const int r = alloc(); if (r < 0) return false;
        if (!emit(GLAVA_B200_COP_SPLAT, r, 0, 0, (float) v->ival)) return false;
        v->reg = r; v->width = 1; v->is_int = false;
        return true;
    }

    // ---- parser ----------------------------------------------------------------------------------------------------
    // precedence, lowest first: ?:  ||  &&  == !=  < > <= >=  + -  * /  unary  postfix
    bool expr(Val* out) {
        if (!lor(out)) return false;
        if (!eat("?")) return true;
        Val a, b;
        if (!expr(&a)) return false;
        if (!eat(":")) return fail_at("expected ':'");
        if (!expr(&b)) return false;
        if (out->reg < 0) {                                   // constant condition: the branch is chosen here
            const bool take_a = out->ival != 0;
            release(take_a ? b : a); *out = take_a ? a : b;
            return true;
        }
        if (out->width != 1) return fail_at("the condition of ?: is not a scalar");
        if (!materialise(&a) || !materialise(&b)) return false;
        if (a.width != b.width) return fail_at("the branches of ?: have different types");
        if (!emit(GLAVA_B200_COP_SELECT, a.reg, a.reg, b.reg, (float) out->reg)) return false;
        release(*out); release(b);
        *out = a;
        return true;
    }
    bool logic(int op, Val* l, Val r) {                       // && ||: folded when both sides are constants
        if (l->reg < 0 && r.reg < 0) { l->ival = op == GLAVA_B200_COP_AND ? (l->ival && r.ival) : (l->ival || r.ival); return true; }
        if (!materialise(l) || !materialise(&r)) return false;
        if (l->width != 1 || r.width != 1) return fail_at("&& / || on a vector");
        if (!emit(op, l->reg, l->reg, r.reg, 0.0f)) return false;
        release(r);
        return true;
    }
    bool lor(Val* out) {
        if (!land(out)) return false;
        while (eat("||")) { Val r; if (!land(&r) || !logic(GLAVA_B200_COP_OR, out, r)) return false; }
        return true;
    }
    bool land(Val* out) {
        if (!equality(out)) return false;
        while (eat("&&")) { Val r; if (!equality(&r) || !logic(GLAVA_B200_COP_AND, out, r)) return false; }
        return true;
    }
    bool compare(const char* op, Val* l, Val r) {
        const std::string o = op;
        if (l->reg < 0 && r.reg < 0) {
            const long a = l->ival, b = r.ival;
            l->ival = o == "<" ? a < b : o == ">" ? a > b : o == "<=" ? a <= b : o == ">=" ? a >= b : o == "==" ? a == b : a != b;
            return true;
        }
        if (!materialise(l) || !materialise(&r)) return false;
        if (l->width != 1 || r.width != 1) return fail_at(std::string("'") + op + "' on a vector");
        const bool swap = o == ">" || o == ">=";
        const int code = (o == "<" || o == ">") ? GLAVA_B200_COP_LT : (o == "<=" || o == ">=") ? GLAVA_B200_COP_LE
                       : o == "==" ? GLAVA_B200_COP_EQ : GLAVA_B200_COP_NE;
        if (!emit(code, l->reg, swap ? r.reg : l->reg, swap ? l->reg : r.reg, 0.0f)) return false;
        release(r);
        return true;
    }
    bool equality(Val* out) {
        if (!relational(out)) return false;
        for (;;) {
            const char* op = is_p("==") ? "==" : is_p("!=") ? "!=" : nullptr;
            if (!op) return true;
            ++pos;
            Val r; if (!relational(&r) || !compare(op, out, r)) return false;
        }
    }
    bool relational(Val* out) {
        if (!additive(out)) return false;
        for (;;) {
            const char* op = is_p("<=") ? "<=" : is_p(">=") ? ">=" : is_p("<") ? "<" : is_p(">") ? ">" : nullptr;
            if (!op) return true;
            ++pos;
            Val r; if (!additive(&r) || !compare(op, out, r)) return false;
        }
    }
    bool additive(Val* out) {
        if (!term(out)) return false;
        for (;;) {
            int op;
            if (eat("+")) op = GLAVA_B200_COP_ADD; else if (eat("-")) op = GLAVA_B200_COP_SUB; else return true;
            Val r; if (!term(&r) || !binary(op, out, r)) return false;
        }
    }
    bool term(Val* out) {
        if (!unary(out)) return false;
        for (;;) {
            int op;
            if (eat("*")) op = GLAVA_B200_COP_MUL; else if (eat("/")) op = GLAVA_B200_COP_DIV; else return true;
            Val r; if (!unary(&r) || !binary(op, out, r)) return false;
        }
    }
    bool binary(int op, Val* l, Val r) {
        if (l->reg < 0 && r.reg < 0) {                       // int op int: folded, C semantics
            long a = l->ival, b = r.ival;
            if (op == GLAVA_B200_COP_DIV && b == 0) return fail_at("integer division by zero");
            l->ival = op == GLAVA_B200_COP_ADD ? a + b : op == GLAVA_B200_COP_SUB ? a - b : op == GLAVA_B200_COP_MUL ? a * b : a / b;
            return true;
        }
        if (!materialise(l) || !materialise(&r)) return false;
        if (l->width != r.width && l->width != 1 && r.width != 1) return fail_at("operands of different vector sizes");
        if (!emit(op, l->reg, l->reg, r.reg, 0.0f)) return false;
        l->width = l->width > r.width ? l->width : r.width;
        release(r);
        return true;
    }
    bool unary(Val* out) {
        if (eat("-")) {
            if (!unary(out)) return false;
            if (out->reg < 0) { out->ival = -out->ival; return true; }
            return emit(GLAVA_B200_COP_NEG, out->reg, out->reg, 0, 0.0f);
        }
        if (eat("+")) return unary(out);
        if (eat("!")) {
            if (!unary(out)) return false;
            if (out->reg < 0) { out->ival = !out->ival; return true; }
            return emit(GLAVA_B200_COP_NOT, out->reg, out->reg, 0, 0.0f);
        }
        if (!primary(out)) return false;
        while (is_p(".")) {                                  // swizzle
            ++pos;
            const Tok t = peek();
            if (t.kind != 'i' || t.text.empty() || t.text.size() > 4) return fail_at("bad swizzle");
            ++pos;
            if (out->reg < 0) return fail_at("swizzle of an integer");
            const int len = (int) t.text.size();
            int lanes[4] = { 0, 0, 0, 0 };
            for (int k = 0; k < len; ++k) {
                const char* sets[3] = { "rgba", "xyzw", "stpq" };
                int lane = -1;
                for (const char* set : sets) { const char* q = strchr(set, t.text[k]); if (q) lane = (int) (q - set); }
                if (lane < 0 || lane >= out->width) return fail_at("swizzle '." + t.text + "' out of range");
                lanes[k] = lane;
            }
            int mask = 0;                                        // one component: splat it, so the value stays a scalar
            for (int k = 0; k < 4; ++k) mask |= (len == 1 ? lanes[0] : (k < len ? lanes[k] : 7)) << (3 * k);
            if (!emit(GLAVA_B200_COP_SHUF, out->reg, out->reg, 0, (float) mask)) return false;
            out->width = (int) t.text.size();
        }
        return true;
    }
    bool args(std::vector<Val>* out) {
        if (!eat("(")) return fail_at("expected '('");
        if (eat(")")) return true;
        for (;;) {
            Val v; if (!expr(&v)) return false;
            out->push_back(v);
            if (eat(")")) return true;
            if (!eat(",")) return fail_at("expected ',' or ')'");
        }
    }
    bool primary(Val* out) {
        const Tok t = peek();
        if (t.kind == 'p' && t.text == "(") {
            ++pos;
            if (!expr(out)) return false;
            return eat(")") ? true : fail_at("expected ')'");
        }
        if (t.kind == 'n') {
            ++pos;
            std::string s = t.text;
            bool is_float = false;
            const bool hex = s.size() > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X');
            if (!hex) for (char c : s) if (c == '.' || c == 'e' || c == 'E' || c == 'f' || c == 'F') is_float = true;
            if (!is_float) {
                while (!s.empty() && (s.back() == 'u' || s.back() == 'U')) s.pop_back();
                char* end; const long v = strtol(s.c_str(), &end, 0);
                if (*end) return fail_at("bad number '" + t.text + "'");
                *out = { -1, 1, true, v };
                return true;
            }
            while (!s.empty() && (s.back() == 'f' || s.back() == 'F')) s.pop_back();
            char* end; const float v = strtof(s.c_str(), &end);
            if (*end) return fail_at("bad number '" + t.text + "'");
            const int r = alloc(); if (r < 0) return false;
            *out = { r, 1, false, 0 };
            return emit(GLAVA_B200_COP_SPLAT, r, 0, 0, v);
        }
        if (t.kind == 'h') {                                  // "#rrggbb[aa]" = vec4 of the %.6f decimals (glsl_ext.c:489-514)
            ++pos;
            float c[4];
            if (!parse_hex_color(t.text.c_str(), c, true)) return fail_at("Invalid color format '" + t.text + "'");
            const int r = alloc(); if (r < 0) return false;
            for (int k = 0; k < 4; ++k) if (!emit(GLAVA_B200_COP_LANE, r, k, 0, c[k])) return false;
            *out = { r, 4, false, 0 };
            return true;
        }
        if (t.kind != 'i') return fail_at("unexpected '" + t.text + "'");
        ++pos;
        const std::string& id = t.text;
        for (const std::string& v : vars) if (v == id) {
            const int r = alloc(); if (r < 0) return false;
            uses_var = true;
            *out = { r, 1, false, 0 };
            return emit(GLAVA_B200_COP_VAR, r, 0, 0, 0.0f);
        }
        if (id == "true" || id == "false") { *out = { -1, 1, true, id == "true" ? 1 : 0 }; return true; }
        if (id == "PI" || id == "TWOPI") {                    // bars/1.frag:33-34 etc.: literals of the module shaders
            const int r = alloc(); if (r < 0) return false;
            *out = { r, 1, false, 0 };
            return emit(GLAVA_B200_COP_SPLAT, r, 0, 0, id == "PI" ? kPI : kTWOPI);
        }
        if (!is_p("(")) return fail_at("'" + id + "' is not available to a colour expression on this path");
        std::vector<Val> a;
        if (!args(&a)) return false;
        return call(id, a, out);
    }
    bool call(const std::string& fn, std::vector<Val>& a, Val* out) {
        // constructors
        int cw = fn == "vec4" ? 4 : fn == "vec3" ? 3 : fn == "vec2" ? 2 : fn == "float" ? 1 : 0;
        if (fn == "int") {
            if (a.size() != 1 || a[0].reg >= 0) return fail_at("int(<non-constant>) is not supported in a colour expression");
            *out = a[0]; return true;
        }
        if (cw) {
            if (a.empty()) return fail_at(fn + "() needs arguments");
            for (Val& v : a) if (!materialise(&v)) return false;
            if (a.size() == 1 && (a[0].width == 1 || a[0].width >= cw)) {      // splat / truncation
                *out = a[0]; out->width = cw;
                if (cw == 1 && a[0].width > 1) return emit(GLAVA_B200_COP_SHUF, out->reg, out->reg, 0, 0.0f);   // float(v) = v.x, splat
                return true;
            }
            int total = 0; for (const Val& v : a) total += v.width;
            if (total != cw) return fail_at(fn + "(): component count mismatch");
            const int r = alloc(); if (r < 0) return false;
            int off = 0;
            for (const Val& v : a) {
                int mask = 0;
                for (int k = 0; k < 4; ++k) mask |= ((k >= off && k < off + v.width) ? (v.width == 1 ? 0 : k - off) : 7) << (3 * k);
                if (!emit(GLAVA_B200_COP_SHUF, r, v.reg, 0, (float) mask)) return false;
                off += v.width;
                release(v);
            }
            *out = { r, cw, false, 0 };
            return true;
        }
        struct Fn { const char* name; int nargs; int op; };
        static const Fn fns[] = {
            { "abs", 1, GLAVA_B200_COP_ABS }, { "floor", 1, GLAVA_B200_COP_FLOOR }, { "ceil", 1, GLAVA_B200_COP_CEIL },
            { "fract", 1, GLAVA_B200_COP_FRACT }, { "sqrt", 1, GLAVA_B200_COP_SQRT }, { "sin", 1, GLAVA_B200_COP_SIN },
            { "cos", 1, GLAVA_B200_COP_COS }, { "log", 1, GLAVA_B200_COP_LOG }, { "sign", 1, GLAVA_B200_COP_SIGN },
            { "trunc", 1, GLAVA_B200_COP_TRUNC }, { "exp", 1, GLAVA_B200_COP_EXP }, { "exp2", 1, GLAVA_B200_COP_EXP2 },
            { "log2", 1, GLAVA_B200_COP_LOG2 }, { "pow", 2, GLAVA_B200_COP_POW }, { "tan", 1, GLAVA_B200_COP_TAN },
            { "atan", 2, GLAVA_B200_COP_ATAN2 },
            { "min", 2, GLAVA_B200_COP_MIN }, { "max", 2, GLAVA_B200_COP_MAX }, { "mod", 2, GLAVA_B200_COP_MOD },
            { "step", 2, GLAVA_B200_COP_STEP },
            { "mix", 3, GLAVA_B200_COP_MIX }, { "clamp", 3, GLAVA_B200_COP_CLAMP }, { "smoothstep", 3, GLAVA_B200_COP_SMOOTHSTEP },
        };
        if (fn == "atan" && a.size() == 1) {                     // atan(x) = atan(x, 1)
            const int r = alloc(); if (r < 0) return false;
            if (!emit(GLAVA_B200_COP_SPLAT, r, 0, 0, 1.0f)) return false;
            a.push_back({ r, 1, false, 0 });
        }
        for (const Fn& f : fns) {
            if (fn != f.name) continue;
            if ((int) a.size() != f.nargs) return fail_at(fn + "(): expected " + std::to_string(f.nargs) + " argument(s)");
            int w = 1;
            for (Val& v : a) { if (!materialise(&v)) return false; w = v.width > w ? v.width : w; }
            for (const Val& v : a) if (v.width != 1 && v.width != w) return fail_at(fn + "(): arguments of different vector sizes");
            const int dst = a[0].reg;
            if (!emit(f.op, dst, a[0].reg, f.nargs > 1 ? a[1].reg : 0, f.nargs > 2 ? (float) a[2].reg : 0.0f)) return false;
            for (size_t k = 1; k < a.size(); ++k) release(a[k]);
            *out = { dst, w, false, 0 };
            return true;
        }
        return fail_at("function '" + fn + "' is not available to a colour expression on this path");
    }

    // text -> program; the expression must be a vec4
    bool compile(const std::string& text, glava_b200_color_prog* prog) {
        std::vector<std::string> active;
        toks.clear(); pos = 0;
        if (!lex(text, &toks, 0, &active)) return false;
        Val v;
        if (!expr(&v)) return false;
        if (peek().kind != 'e') return fail_at("unexpected '" + peek().text + "'");
        if (v.reg < 0 || v.width != 4) return fail_at("the expression is not a vec4");
        memset(prog, 0, sizeof(*prog));
        prog->n_ops = (int) ops.size(); prog->result = v.reg;
        for (size_t k = 0; k < ops.size(); ++k) prog->ops[k] = ops[k];
        return true;
    }
};

}  // namespace glb
#endif
