"""TEST INFRASTRUCTURE — a small interpreter for the subset of GLSL that GLava's shipped shaders use.

Why: the raster half of the reference is GLSL.  In round 1 no OpenGL had been found in this image, so the C restatement in
glava_oracle.c could not be pinned by running the reference; this module closed most of that gap.  (Round 2 found a real one —
Mesa llvmpipe inside the Nsight Compute bundle, oracle/ref_gl.py — and the reference itself now renders the goldens that pin
the oracle, the kernels AND this interpreter: tests/golden/llvmpipe_golden.npz.  The interpreter stays for what a real GL
cannot give: exact single-stage evaluation on chosen textures and the 1000-seed config differential.)  It
reads the reference's OWN shader sources (shaders/glava/<module>/<n>.frag, util/*.frag, the module .glsl configs),
applies GLava's source extensions (glsl_ext.c: `#include` with ':' / '@', `#request`, `#expand`, `#rrggbb` colour
literals, `@name:default` binds, `#define` overriding) and the header GLava injects (render.c:284-327), runs a C
preprocessor, parses the result and EVALUATES it per fragment in IEEE float32 — so macro precedence, int / float typing,
operand order, stage chaining and RGBA8 / R16 quantisation all come from the reference's text, not from a restatement.

What it does not take from the reference (because GLSL leaves it to the implementation) is fixed as in DESIGN.md 4.3:
float = binary32 with every operation rounded; transcendentals = this host's libm (sinf, cosf, atan2f, logf, powf ...);
round() = nearest-even; texelFetch outside the texture = 0; unorm stores = rint(float32(clamp(c) * MAX)), ties to even
(what Mesa does: tests/golden/llvmpipe_golden.npz, the reference on a real llvmpipe, pins it).

Only tests/ and tests/golden/make_glsl_golden.py use this file; it needs /root/reference at run time, so everything that
must run elsewhere goes through the golden vectors that script commits.
"""
import ctypes
import ctypes.util
import os
import re

import numpy as np

F32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("sinf", "cosf", "tanf", "logf", "expf", "sqrtf", "floorf", "ceilf", "atanf", "rintf", "truncf", "exp2f", "log2f"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
for _n in ("atan2f", "powf", "fmodf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float, ctypes.c_float]


def _m1(name):
    fn = getattr(_libm, name)
    return lambda x: F32(fn(float(x)))


def _m2(name):
    fn = getattr(_libm, name)
    return lambda x, y: F32(fn(float(x), float(y)))


# ======================================================================================================================
# 1. GLava source extensions (glsl_ext.c) — text in, text out
# ======================================================================================================================
def hex_colour(lit):
    """#rrggbb[aa] -> 'vec4(r, g, b, a)' with the "%.6f" the reference prints (glsl_ext.c:489-514)"""
    h = lit.lstrip("#")
    vals = [int(h[i:i + 2], 16) / 255.0 for i in range(0, len(h), 2)]
    if len(vals) == 3:
        vals.append(1.0)
    return "vec4(%.6f, %.6f, %.6f, %.6f)" % tuple(float(F32(v)) for v in vals)


_COLOUR = re.compile(r"#([0-9a-fA-F]{8}|[0-9a-fA-F]{6})\b")
_BIND = re.compile(r"@([A-Za-z_][A-Za-z_0-9]*):")


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


class ExtCtx:
    """include context of glsl_ext.c:161-183.  `fallback`: a user config directory normally holds a full copy of the
    shader tree (glava --copy-config); tests give one with only the files they changed, the rest is taken from here."""

    def __init__(self, cd, cfd, dd, efuncs, fallback=None):
        self.cd, self.cfd, self.dd, self.efuncs, self.fallback = cd, cfd, dd, efuncs, fallback


def ext_process(path, ctx, depth=0):
    """returns a list of source lines with the glava extensions resolved"""
    assert depth < 32
    src = _strip_comments(open(path).read())
    out = []
    for line in src.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            body = s[1:].strip()
            word = re.match(r"[A-Za-z_]+", body)
            word = word.group(0) if word else ""
            if word == "request":
                continue
            if word == "include":
                target = re.search(r'"([^"]*)"', body).group(1)
                if target.startswith(":") and ctx.cfd:
                    target = target[1:]; ctx.cd = ctx.cfd
                if target.startswith("@"):
                    target = target[1:]; ctx.cd = ctx.dd
                path2 = os.path.join(ctx.cd, target)
                if not os.path.exists(path2) and ctx.fallback:
                    path2 = os.path.join(ctx.fallback, target)
                out += ext_process(path2, ctx, depth + 1)
                continue
            if word == "expand":
                _, macro, arg = body.split()[:3]
                out += ["%s(%d);" % (macro, t) for t in range(int(ctx.efuncs[arg]))]     # glsl_ext.c:327
                continue
            if word in ("define", "undef", "if", "ifdef", "ifndef", "else", "elif", "endif", "error", "version", "line", "pragma"):
                line = "#" + _COLOUR.sub(lambda m: hex_colour(m.group(0)), body)
                out.append(_BIND.sub("", line))                                             # @name:default -> default
                continue
        line = _COLOUR.sub(lambda m: hex_colour(m.group(0)), line)
        out.append(_BIND.sub("", line))
    return out


# ======================================================================================================================
# 2. C preprocessor
# ======================================================================================================================
_TOKEN = re.compile(r"""
    (?P<float>(?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?[fF]?|\d+[eE][-+]?\d+[fF]?|\d+[fF])
  | (?P<int>0[xX][0-9a-fA-F]+|\d+[uU]?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\#\#|\+\+|--|\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||[-+*/%<>=!?:;,.(){}\[\]&|^~\#])
  | (?P<ws>\s+)
""", re.X)


def tokenize(text):
    toks, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError("cannot tokenize %r" % text[pos:pos + 30])
        pos = m.end()
        if m.lastgroup != "ws":
            toks.append((m.lastgroup, m.group(0)))
    return toks


class DisabledStage(Exception):
    pass


class Preprocessor:
    def __init__(self):
        self.macros = {}       # name -> (params or None, body tokens)

    def define(self, name, body_text, params=None):
        self.macros[name] = (params, tokenize(body_text))

    # -- macro expansion ------------------------------------------------------------------------------------------------
    def expand(self, toks, hide=frozenset()):
        out, i = [], 0
        while i < len(toks):
            kind, val = toks[i]
            if kind == "id" and val in self.macros and val not in hide:
                params, body = self.macros[val]
                if params is None:
                    exp = self.expand(body, hide | {val})
                    # rescan with the following tokens: an object-like macro may expand to the NAME of a function-like one
                    # whose argument list follows the invocation (ROUND_FORMULA -> sinusoidal, then "(x)")
                    if exp and exp[-1][0] == "id" and exp[-1][1] in self.macros and self.macros[exp[-1][1]][0] is not None \
                            and i + 1 < len(toks) and toks[i + 1][1] == "(":
                        out += exp[:-1]
                        toks = [exp[-1]] + toks[i + 1:]
                        i = 0
                        continue
                    out += exp
                    i += 1
                    continue
                if i + 1 < len(toks) and toks[i + 1][1] == "(":
                    args, depth, j, cur = [], 0, i + 2, []
                    while True:
                        k, v = toks[j]
                        if v == "(":
                            depth += 1
                        if v == ")":
                            if depth == 0:
                                break
                            depth -= 1
                        if v == "," and depth == 0:
                            args.append(cur); cur = []
                        else:
                            cur.append((k, v))
                        j += 1
                    if cur or args:
                        args.append(cur)
                    sub = []
                    for k, v in body:
                        if k == "id" and v in params:
                            sub += [("arg", params.index(v))]
                        else:
                            sub.append((k, v))
                    # token pasting first (operands unexpanded), then ordinary parameters fully expanded
                    res, q = [], 0
                    while q < len(sub):
                        if q + 2 < len(sub) and sub[q + 1][1] == "##":
                            left = sub[q]; right = sub[q + 2]
                            ltxt = "".join(v for _, v in args[left[1]]) if left[0] == "arg" else left[1]
                            rtxt = "".join(v for _, v in args[right[1]]) if right[0] == "arg" else right[1]
                            res += tokenize(ltxt + rtxt)
                            q += 3
                        elif sub[q][0] == "arg":
                            res += self.expand(args[sub[q][1]], hide)
                            q += 1
                        else:
                            res.append(sub[q]); q += 1
                    out += self.expand(res, hide | {val})
                    i = j + 1
                    continue
            out.append((kind, val))
            i += 1
        return out

    # -- #if expressions --------------------------------------------------------------------------------------------------
    def eval_if(self, text):
        toks = tokenize(text)
        res, i = [], 0
        while i < len(toks):                       # defined X / defined(X)
            if toks[i][1] == "defined":
                if toks[i + 1][1] == "(":
                    res.append(("int", "1" if toks[i + 2][1] in self.macros else "0")); i += 4
                else:
                    res.append(("int", "1" if toks[i + 1][1] in self.macros else "0")); i += 2
            else:
                res.append(toks[i]); i += 1
        toks = self.expand(res)
        py = []
        for k, v in toks:
            if k == "id":
                py.append("0")                     # unknown identifiers are 0 in #if
            elif k in ("int", "float"):
                py.append(v.rstrip("fFuU"))
            elif v == "&&":
                py.append(" and ")
            elif v == "||":
                py.append(" or ")
            elif v == "!":
                py.append(" not ")
            elif v == "/":
                py.append("//")
            else:
                py.append(v)
        return bool(eval("".join(py), {"__builtins__": {}}, {}))

    # -- whole translation unit -------------------------------------------------------------------------------------------
    def run(self, lines):
        out, pending = [], []
        stack = []                                 # [taking, any_taken_so_far, parent_taking]

        def flush():
            if pending:
                out.extend(self.expand(tokenize("\n".join(pending))))
                pending.clear()

        active = lambda: all(s[0] for s in stack)
        for line in lines:
            s = line.strip()
            if not s.startswith("#"):
                if active():
                    pending.append(line)
                continue
            flush()
            m = re.match(r"#\s*([A-Za-z_]+)\s*(.*)", s)
            if not m:
                continue
            d, rest = m.group(1), m.group(2).strip()
            if d in ("ifdef", "ifndef", "if"):
                parent = active()
                cond = False
                if parent:
                    cond = (rest.split()[0] in self.macros) if d == "ifdef" else (rest.split()[0] not in self.macros) if d == "ifndef" else self.eval_if(rest)
                stack.append([cond and parent, cond, parent])
            elif d == "elif":
                top = stack[-1]
                cond = top[2] and not top[1] and self.eval_if(rest)
                top[0] = cond; top[1] = top[1] or cond
            elif d == "else":
                top = stack[-1]
                top[0] = top[2] and not top[1]; top[1] = True
            elif d == "endif":
                stack.pop()
            elif not active():
                continue
            elif d == "define":
                m2 = re.match(r"([A-Za-z_][A-Za-z_0-9]*)(\(([^)]*)\))?\s*(.*)", rest)
                name, has_params, plist, body = m2.group(1), m2.group(2), m2.group(3), m2.group(4)
                if has_params and rest[len(name)] == "(":
                    self.macros[name] = ([p.strip() for p in plist.split(",") if p.strip()], tokenize(body))
                else:
                    self.macros[name] = (None, tokenize(rest[len(name):].strip()))
            elif d == "undef":
                self.macros.pop(rest.split()[0], None)
            elif d == "error":
                if "__disablestage" in rest:
                    raise DisabledStage()
                raise SyntaxError("#error " + rest)
        flush()
        return out


# ======================================================================================================================
# 3. Parser (tokens -> AST tuples)
# ======================================================================================================================
TYPES = {"void", "float", "int", "bool", "vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4", "sampler1D", "sampler2D", "uint"}
QUALS = {"uniform", "in", "out", "inout", "const", "highp", "mediump", "lowp", "flat"}
ASSIGN = {"=", "+=", "-=", "*=", "/="}
BINPREC = [["||"], ["&&"], ["==", "!="], ["<", ">", "<=", ">="], ["+", "-"], ["*", "/", "%"]]


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0
        self.pixel_center_integer = False

    def peek(self, k=0):
        return self.t[self.i + k][1] if self.i + k < len(self.t) else None

    def next(self):
        v = self.t[self.i]; self.i += 1
        return v

    def expect(self, v):
        k, got = self.next()
        if got != v:
            raise SyntaxError("expected %r, got %r at token %d" % (v, got, self.i))

    def unit(self):
        globals_, funcs = [], {}
        while self.i < len(self.t):
            if self.peek() == ";":
                self.next(); continue
            if self.peek() == "layout":
                self.next(); self.expect("(")
                while self.peek() != ")":
                    if self.next()[1] == "pixel_center_integer":
                        self.pixel_center_integer = True
                self.expect(")")
            quals = []
            while self.peek() in QUALS:
                quals.append(self.next()[1])
            typ = self.next()[1]
            assert typ in TYPES, typ
            name = self.next()[1]
            if self.peek() == "(":
                self.next()
                params = []
                while self.peek() != ")":
                    while self.peek() in QUALS:
                        self.next()
                    pt = self.next()[1]
                    if pt == "void":
                        continue
                    pn = self.next()[1]
                    params.append((pt, pn))
                    if self.peek() == ",":
                        self.next()
                self.expect(")")
                funcs[name] = (typ, params, self.block())
            else:
                while True:
                    init = None
                    if self.peek() == "=":
                        self.next(); init = self.assign()
                    globals_.append((quals, typ, name, init))
                    if self.peek() == ",":
                        self.next(); name = self.next()[1]; continue
                    break
                self.expect(";")
        return globals_, funcs

    def block(self):
        self.expect("{")
        stmts = []
        while self.peek() != "}":
            stmts.append(self.stmt())
        self.expect("}")
        return ("block", stmts)

    def stmt(self):
        p = self.peek()
        if p == "{":
            return self.block()
        if p == ";":
            self.next(); return ("block", [])
        if p == "if":
            self.next(); self.expect("("); c = self.expr(); self.expect(")")
            a = self.stmt(); b = None
            if self.peek() == "else":
                self.next(); b = self.stmt()
            return ("if", c, a, b)
        if p == "for":
            self.next(); self.expect("(")
            init = self.stmt() if self.peek() != ";" else (self.next(), None)[1]
            cond = self.expr() if self.peek() != ";" else None
            self.expect(";")
            step = self.expr() if self.peek() != ")" else None
            self.expect(")")
            return ("for", init, cond, step, self.stmt())
        if p == "while":
            self.next(); self.expect("("); c = self.expr(); self.expect(")")
            return ("for", None, c, None, self.stmt())
        if p in ("break", "continue"):
            self.next(); self.expect(";")
            return (p,)
        if p == "return":
            self.next()
            e = None if self.peek() == ";" else self.expr()
            self.expect(";")
            return ("return", e)
        if p in QUALS or (p in TYPES and self.peek(1) != "("):
            while self.peek() in QUALS:
                self.next()
            typ = self.next()[1]
            decls = []
            while True:
                name = self.next()[1]; init = None
                if self.peek() == "=":
                    self.next(); init = self.assign()
                decls.append((name, init))
                if self.peek() == ",":
                    self.next(); continue
                break
            self.expect(";")
            return ("decl", typ, decls)
        e = self.expr()
        self.expect(";")
        return ("expr", e)

    def expr(self):
        e = self.assign()
        while self.peek() == ",":
            self.next(); e = ("comma", e, self.assign())
        return e

    def assign(self):
        lhs = self.ternary()
        if self.peek() in ASSIGN:
            op = self.next()[1]
            return ("assign", op, lhs, self.assign())
        return lhs

    def ternary(self):
        c = self.binary(0)
        if self.peek() == "?":
            self.next(); a = self.assign(); self.expect(":"); b = self.assign()
            return ("ternary", c, a, b)
        return c

    def binary(self, level):
        if level == len(BINPREC):
            return self.unary()
        e = self.binary(level + 1)
        while self.peek() in BINPREC[level]:
            op = self.next()[1]
            e = ("bin", op, e, self.binary(level + 1))
        return e

    def unary(self):
        p = self.peek()
        if p in ("-", "+", "!"):
            self.next(); return ("un", p, self.unary())
        if p in ("++", "--"):
            self.next(); return ("preinc", p, self.unary())
        return self.postfix()

    def postfix(self):
        k, v = self.next()
        if v == "(":
            e = self.expr(); self.expect(")")
        elif k == "float":
            e = ("lit", F32(float(v.rstrip("fF"))))
        elif k == "int":
            e = ("lit", int(v.rstrip("uU"), 0))
        elif v in ("true", "false"):
            e = ("lit", v == "true")
        elif k == "id":
            if self.peek() == "(":
                self.next(); args = []
                while self.peek() != ")":
                    args.append(self.assign())
                    if self.peek() == ",":
                        self.next()
                self.expect(")")
                e = ("call", v, args)
            else:
                e = ("var", v)
        else:
            raise SyntaxError("unexpected token %r" % v)
        while True:
            p = self.peek()
            if p == ".":
                self.next(); e = ("field", e, self.next()[1])
            elif p == "[":
                self.next(); idx = self.expr(); self.expect("]"); e = ("index", e, idx)
            elif p in ("++", "--"):
                self.next(); e = ("postinc", p, e)
            else:
                return e


# ======================================================================================================================
# 4. Values and evaluation
# ======================================================================================================================
class Vec:
    __slots__ = ("t", "v")

    def __init__(self, t, v):
        self.t, self.v = t, list(v)

    def __repr__(self):
        return "%s(%s)" % (self.t, ", ".join(str(x) for x in self.v))


SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3, "s": 0, "t": 1, "p": 2, "q": 3}


def is_float(x):
    return isinstance(x, np.floating)


def to_float(x):
    if isinstance(x, Vec):
        return Vec("vec%d" % len(x.v), [to_float(c) for c in x.v])
    return F32(x) if not is_float(x) else x


def to_int(x):
    if isinstance(x, Vec):
        return Vec("ivec%d" % len(x.v), [to_int(c) for c in x.v])
    if is_float(x):
        return int(_libm.truncf(float(x)))
    return int(x)


def convert(typ, x):
    if typ == "float":
        assert not isinstance(x, Vec), (typ, x)
        return to_float(x)
    if typ in ("int", "uint"):
        assert not isinstance(x, Vec)
        return to_int(x) if is_float(x) else int(x)
    if typ == "bool":
        return bool(x)
    if typ.startswith("vec"):
        assert isinstance(x, Vec) and len(x.v) == int(typ[3]), (typ, x)
        return to_float(x)
    if typ.startswith("ivec"):
        assert isinstance(x, Vec) and len(x.v) == int(typ[4])
        return to_int(x)
    return x


def _arith(op, a, b):
    if is_float(a) or is_float(b):
        a, b = to_float(a), to_float(b)
        if op == "+": return F32(a + b)
        if op == "-": return F32(a - b)
        if op == "*": return F32(a * b)
        if op == "/":
            with np.errstate(divide="ignore", invalid="ignore"):
                return F32(a / b)
        raise ValueError(op)
    if op == "+": return a + b
    if op == "-": return a - b
    if op == "*": return a * b
    if op == "/": return int(a / b) if b else 0            # C-style truncation
    if op == "%": return a - b * int(a / b)
    raise ValueError(op)


def arith(op, a, b):
    va, vb = isinstance(a, Vec), isinstance(b, Vec)
    if va or vb:
        n = len(a.v) if va else len(b.v)
        av = a.v if va else [a] * n
        bv = b.v if vb else [b] * n
        res = [_arith(op, x, y) for x, y in zip(av, bv)]
        return Vec(("vec%d" if any(is_float(c) for c in res) else "ivec%d") % n, res)
    return _arith(op, a, b)


def _round_even(x):
    return F32(_libm.rintf(float(x)))


def componentwise(fn, *args):
    n = max((len(a.v) for a in args if isinstance(a, Vec)), default=0)
    if n == 0:
        return fn(*[to_float(a) for a in args])
    cols = [a.v if isinstance(a, Vec) else [a] * n for a in args]
    return Vec("vec%d" % n, [fn(*[to_float(c) for c in row]) for row in zip(*cols)])


def _mod(x, y):
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(x - F32(y * F32(_libm.floorf(float(F32(x / y))))))


def _mix(a, b, t):
    return F32(F32(a * F32(F32(1.0) - t)) + F32(b * t))


def _clamp(x, lo, hi):
    return min(max(x, lo), hi)


def _smoothstep(e0, e1, x):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = _clamp(F32(F32(x - e0) / F32(e1 - e0)), F32(0.0), F32(1.0))
    return F32(F32(t * t) * F32(F32(3.0) - F32(F32(2.0) * t)))


def _sign(x):
    return F32(1.0) if x > 0 else (F32(-1.0) if x < 0 else F32(0.0))


BUILTINS = {
    "sin": lambda x: componentwise(_m1("sinf"), x), "cos": lambda x: componentwise(_m1("cosf"), x),
    "tan": lambda x: componentwise(_m1("tanf"), x),
    "log": lambda x: componentwise(_m1("logf"), x), "exp": lambda x: componentwise(_m1("expf"), x),
    "log2": lambda x: componentwise(_m1("log2f"), x), "exp2": lambda x: componentwise(_m1("exp2f"), x),
    "sqrt": lambda x: componentwise(lambda v: F32(np.sqrt(v)), x),
    "floor": lambda x: componentwise(_m1("floorf"), x), "ceil": lambda x: componentwise(_m1("ceilf"), x),
    "round": lambda x: componentwise(_round_even, x), "trunc": lambda x: componentwise(_m1("truncf"), x),
    "abs": lambda x: (abs(x) if isinstance(x, int) and not isinstance(x, bool) else componentwise(lambda v: F32(abs(v)), x)),
    "sign": lambda x: componentwise(_sign, x),
    "fract": lambda x: componentwise(lambda v: F32(v - F32(_libm.floorf(float(v)))), x),
    "mod": lambda x, y: componentwise(_mod, x, y),
    "min": lambda a, b: componentwise(lambda p, q: min(p, q), a, b),
    "max": lambda a, b: componentwise(lambda p, q: max(p, q), a, b),
    "clamp": lambda x, lo, hi: componentwise(_clamp, x, lo, hi),
    "mix": lambda a, b, t: componentwise(_mix, a, b, t),
    "smoothstep": lambda e0, e1, x: componentwise(_smoothstep, e0, e1, x),
    "pow": lambda a, b: componentwise(_m2("powf"), a, b),
    "atan": lambda *a: componentwise(_m2("atan2f"), *a) if len(a) == 2 else componentwise(_m1("atanf"), *a),
    "step": lambda e, x: componentwise(lambda p, q: F32(0.0) if q < p else F32(1.0), e, x),
}


class Sampler1D:
    """R16 1-D texture: texelFetch = u / 65535 (0 outside), texture() = NEAREST + REPEAT (render.c:510-524)"""

    def __init__(self, texels):
        self.u = np.asarray(texels, dtype=np.uint16)

    def fetch(self, i):
        r = F32(0.0) if (i < 0 or i >= len(self.u)) else F32(F32(self.u[i]) / F32(65535.0))
        return Vec("vec4", [r, F32(0.0), F32(0.0), F32(1.0)])

    def texture(self, coord):
        n = len(self.u)
        i = int(_libm.floorf(float(F32(to_float(coord) * F32(n))))) % n
        return self.fetch(i)


class Sampler2D:
    """previous stage's RGBA8 surface, produced on demand"""

    def __init__(self, w, h, pixel_fn):
        self.w, self.h, self.fn = w, h, pixel_fn

    def fetch(self, xy):
        x, y = xy.v
        if x < 0 or y < 0 or x >= self.w or y >= self.h:
            return Vec("vec4", [F32(0.0)] * 4)
        return Vec("vec4", [F32(F32(c) / F32(255.0)) for c in self.fn(x, y)])


class Break(Exception):
    pass


class Continue(Exception):
    pass


class Return(Exception):
    def __init__(self, v):
        self.v = v


class Shader:
    """one fragment-shader stage, parsed once, evaluated per fragment.  Variables are cells [declared type, value]."""

    def __init__(self, toks):
        p = Parser(toks)
        self.globals_decl, self.funcs = p.unit()
        self.pixel_center_integer = p.pixel_center_integer

    def run(self, uniforms, x, y):
        off = F32(0.0) if self.pixel_center_integer else F32(0.5)
        g = {k: [None, v] for k, v in uniforms.items()}
        g["gl_FragCoord"] = [None, Vec("vec4", [F32(F32(x) + off), F32(F32(y) + off), F32(0.5), F32(1.0)])]
        self.g, self.scopes = g, []
        for quals, typ, name, init in self.globals_decl:
            if name in g:
                g[name][0] = typ
                continue
            if init is not None:
                g[name] = [typ, convert(typ, self.ev(init))]
            elif typ.startswith("vec"):
                g[name] = [typ, Vec(typ, [F32(0.0)] * int(typ[3]))]   # `out vec4 fragment`: undefined until written; 0 here
            else:
                g[name] = [typ, None]
        self.call_user("main", [])
        return {k: c[1] for k, c in g.items()}

    # -- variables --------------------------------------------------------------------------------------------------------
    def cell(self, name):
        for s in reversed(self.scopes):
            if name in s:
                return s[name]
        if name in self.g:
            return self.g[name]
        raise NameError(name)

    def assign_to(self, node, val):
        kind = node[0]
        if kind == "var":
            c = self.cell(node[1])
            c[1] = convert(c[0], val) if c[0] else val
        elif kind == "field":
            base = self.ev(node[1])
            idx = [SWZ[ch] for ch in node[2]]
            conv = to_float if base.t.startswith("vec") else to_int
            if len(idx) == 1:
                base.v[idx[0]] = conv(val)
            else:
                for k, i in enumerate(idx):
                    base.v[i] = conv(val.v[k])
        else:
            raise SyntaxError("bad lvalue %r" % (node,))

    # -- calls --------------------------------------------------------------------------------------------------------------
    def call_user(self, name, args):
        rtype, params, body = self.funcs[name]
        scope = {}
        for (pt, pn), a in zip(params, args):
            scope[pn] = [pt, a if pt in ("sampler1D", "sampler2D") else convert(pt, copy_val(a))]
        saved = self.scopes
        self.scopes = [scope]
        ret = None
        try:
            self.exec(body)
        except Return as r:
            ret = r.v
        finally:
            self.scopes = saved
        if rtype == "void" or ret is None:
            return None
        return convert(rtype, ret)

    def construct(self, typ, args):
        if typ in ("float", "int", "bool", "uint"):
            a = args[0]
            if isinstance(a, Vec):
                a = a.v[0]
            return to_float(a) if typ == "float" else convert(typ, a)
        n = int(typ[-1])
        flat = []
        for a in args:
            flat += a.v if isinstance(a, Vec) else [a]
        if len(flat) == 1:
            flat = flat * n
        flat = flat[:n]
        assert len(flat) == n, (typ, args)
        return Vec(typ, [to_float(c) if typ.startswith("vec") else to_int(c) for c in flat])

    # -- statements -------------------------------------------------------------------------------------------------------
    def exec(self, node):
        kind = node[0]
        if kind == "block":
            self.scopes.append({})
            try:
                for st in node[1]:
                    self.exec(st)
            finally:
                self.scopes.pop()
        elif kind == "decl":
            _, typ, decls = node
            for name, init in decls:
                self.scopes[-1][name] = [typ, convert(typ, copy_val(self.ev(init))) if init is not None else None]
        elif kind == "expr":
            self.ev(node[1])
        elif kind == "if":
            if self.ev(node[1]):
                self.exec(node[2])
            elif node[3] is not None:
                self.exec(node[3])
        elif kind == "for":
            _, init, cond, step, body = node
            self.scopes.append({})
            try:
                if init is not None:
                    self.exec(init)
                while cond is None or self.ev(cond):
                    try:
                        self.exec(body)
                    except Break:
                        break
                    except Continue:
                        pass
                    if step is not None:
                        self.ev(step)
            finally:
                self.scopes.pop()
        elif kind == "return":
            raise Return(self.ev(node[1]) if node[1] is not None else None)
        elif kind == "break":
            raise Break()
        elif kind == "continue":
            raise Continue()
        else:
            raise SyntaxError(kind)

    # -- expressions ------------------------------------------------------------------------------------------------------
    def ev(self, node):
        kind = node[0]
        if kind == "lit":
            return node[1]
        if kind == "var":
            v = self.cell(node[1])[1]
            if v is None:
                raise ValueError("read of uninitialised variable %s" % node[1])
            return v
        if kind == "field":
            base = self.ev(node[1])
            idx = [SWZ[ch] for ch in node[2]]
            if len(idx) == 1:
                return base.v[idx[0]]
            return Vec(base.t[:-1] + str(len(idx)), [base.v[i] for i in idx])
        if kind == "index":
            return self.ev(node[1]).v[to_int(self.ev(node[2]))]
        if kind == "un":
            v = self.ev(node[2])
            if node[1] == "!":
                return not v
            if node[1] == "+":
                return v
            if isinstance(v, Vec):
                return Vec(v.t, [(F32(-c) if is_float(c) else -c) for c in v.v])
            return F32(-v) if is_float(v) else -v
        if kind == "bin":
            op = node[1]
            if op == "&&":
                return bool(self.ev(node[2])) and bool(self.ev(node[3]))
            if op == "||":
                return bool(self.ev(node[2])) or bool(self.ev(node[3]))
            a, b = self.ev(node[2]), self.ev(node[3])
            if op in ("+", "-", "*", "/", "%"):
                return arith(op, a, b)
            if is_float(a) or is_float(b):
                a, b = to_float(a), to_float(b)
            return {"==": a == b, "!=": a != b, "<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]
        if kind == "ternary":
            return self.ev(node[2]) if self.ev(node[1]) else self.ev(node[3])
        if kind == "assign":
            _, op, lhs, rhs = node
            val = self.ev(rhs)
            if op != "=":
                val = arith(op[0], self.ev(lhs), val)
            self.assign_to(lhs, copy_val(val))
            return self.ev(lhs)
        if kind in ("preinc", "postinc"):
            old = copy_val(self.ev(node[2]))
            self.assign_to(node[2], arith("+" if node[1] == "++" else "-", old, 1))
            return old if kind == "postinc" else self.ev(node[2])
        if kind == "comma":
            self.ev(node[1])
            return self.ev(node[2])
        if kind == "call":
            name = node[1]
            args = [self.ev(a) for a in node[2]]
            if name in self.funcs:
                return self.call_user(name, args)
            if name in TYPES:
                return self.construct(name, args)
            if name == "texelFetch":
                smp = args[0]
                return smp.fetch(args[1]) if isinstance(smp, Sampler2D) else smp.fetch(to_int(args[1]))
            if name == "texture":
                return args[0].texture(args[1])
            if name == "length":
                return F32(np.sqrt(sum((F32(c * c) for c in to_float(args[0]).v), F32(0.0))))
            return BUILTINS[name](*args)
        raise SyntaxError(kind)


def copy_val(v):
    return Vec(v.t, v.v) if isinstance(v, Vec) else v


# ======================================================================================================================
# 5. GLava program loading: header injection (render.c:284-327), stage chain, quantisation
# ======================================================================================================================
def unorm8(c):
    c = float(c)
    if not c > 0.0:
        return 0
    if not c < 1.0:
        return 255
    return int(np.rint(F32(F32(c) * F32(255.0))))              # one rounding of the float32 product, ties to even (Mesa)


def unorm16(c):
    c = float(c)
    if not c > 0.0:
        return 0
    if not c < 1.0:
        return 65535
    return int(np.rint(F32(F32(c) * F32(65535.0))))


def header_macros(pp, smooth_factor=0.025, avg_frames=5, avg_window=1, premultiply_alpha=1, channels=2, pre_smoothed=1):
    pp.define("_SMOOTH_FACTOR", "%.6f" % smooth_factor)                     # "#define _SMOOTH_FACTOR %.6f"
    pp.define("USE_STDIN", "0")
    pp.define("_AVG_FRAMES", str(avg_frames)); pp.define("_AVG_WINDOW", str(avg_window))
    pp.define("_USE_ALPHA", "1"); pp.define("_PREMULTIPLY_ALPHA", str(premultiply_alpha))
    pp.define("_CHANNELS", str(channels)); pp.define("_UNIFORM_LIMIT", "1024")
    pp.define("_PRE_SMOOTHED_AUDIO", str(pre_smoothed))


def load_stage(path, root, overrides=None, config_dir=None, **hdr):
    """preprocess + parse one shader file; overrides: {macro: text} applied AFTER the module's config includes, the way a
    user's copy of <module>.glsl would redefine them.  Raises DisabledStage for `#error __disablestage`."""
    efuncs = {"_AVG_FRAMES": hdr.get("avg_frames", 5)}
    ctx = ExtCtx(os.path.dirname(path), config_dir or root, root, efuncs, fallback=root)
    lines = ext_process(path, ctx)
    pp = Preprocessor()
    header_macros(pp, **hdr)
    if overrides:
        # a redefinition inside the config wins over an earlier one; emulate a user config appended to the default one
        patched = []
        for ln in lines:
            m = re.match(r"#\s*define\s+([A-Za-z_][A-Za-z_0-9]*)\b", ln.strip())
            if m and m.group(1) in overrides:
                patched.append("#define %s %s" % (m.group(1), overrides[m.group(1)]))
            else:
                patched.append(ln)
        lines = patched
    return Shader(pp.run(lines))


class ModuleProgram:
    """the stage chain of one module (render.c stage loading: 1.frag, 2.frag, ... until a file is missing)"""

    def __init__(self, root, module, w, h, tex_l, tex_r, overrides=None, config_dir=None, clear_color=(0.0, 0.0, 0.0, 0.0), **hdr):
        self.w, self.h = w, h
        # setopacity other than "native" (premultiply_alpha = 0): every stage is drawn with GL_BLEND enabled,
        # glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA), over the target glClear'd to `setbg` (render.c:1467-1470,
        # 1700, 2028).  Blending in the target's unorm8 fixed point, as llvmpipe does it (pinned by llvmpipe_golden.npz).
        self.blend = hdr.get("premultiply_alpha", 1) == 0
        self.clear8 = tuple(unorm8(F32(c)) for c in clear_color)
        self.stages = []
        k = 1
        while os.path.exists(os.path.join(root, module, "%d.frag" % k)):
            try:
                self.stages.append(load_stage(os.path.join(root, module, "%d.frag" % k), root, overrides, config_dir, **hdr))
            except DisabledStage:
                pass
            k += 1
        n = len(tex_l)
        self.base = {"screen": Vec("ivec2", [w, h]), "audio_sz": n, "audio_l": Sampler1D(tex_l), "audio_r": Sampler1D(tex_r),
                     "time": F32(0.0)}
        self.cache = [dict() for _ in self.stages]

    def stage_pixel(self, k, x, y):
        key = (x, y)
        c = self.cache[k]
        if key not in c:
            u = dict(self.base)
            if k > 0:
                u["tex"] = Sampler2D(self.w, self.h, lambda px, py: self.stage_pixel(k - 1, px, py))
            g = self.stages[k].run(u, x, y)
            frag = [to_float(v) for v in g["fragment"].v]
            if self.blend:
                # unorm8 fixed-point blend (llvmpipe): mul_norm(Cs, As) + mul_norm(Cd, 255 - As), saturating
                def mul_norm(p, q):
                    t = p * q + 128
                    return (t + (t >> 8)) >> 8
                s8 = [unorm8(v) for v in frag]
                a = s8[3]
                c[key] = tuple(min(mul_norm(s8[i], a) + mul_norm(self.clear8[i], 255 - a), 255) for i in range(4))
                return c[key]
            c[key] = tuple(unorm8(v) for v in frag)
        return c[key]

    def pixel(self, x, y):
        """final RGBA8 of pixel (x, y), y = 0 at the bottom (GL window coordinates)"""
        return self.stage_pixel(len(self.stages) - 1, x, y)


def run_1d_pass(path, root, n_out, uniforms, **hdr):
    """util/*_pass.frag over a 1-D R16 target: returns uint16[n_out] of the red channel"""
    sh = load_stage(path, root, None, **hdr)
    out = np.zeros(n_out, np.uint16)
    for x in range(n_out):
        g = sh.run(dict(uniforms), x, 0)
        out[x] = unorm16(g["fragment"].v[0])
    return out
