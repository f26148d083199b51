"""User COLOR / BAR_OUTLINE macros that are general GLSL expressions.  Reference side: the module's
own shaders with the user's <module>.glsl pasted in, evaluated by oracle/glsl_interp.py (tests/golden/color_expr_golden.npz,
made by tests/golden/make_color_expr_golden.py).  Product side: the config reader compiles the macro text into a colour
program (glava_b200/csrc/color_compile.h) that raster_core.h's eval_color_prog runs — checked here through the host build
of the product arithmetic (tests/emul), and on the device in tests/test_zz_gpu_color_expr.py."""
import os

import numpy as np
import pytest

import glava_b200 as g
from tests.conftest import GOLDEN


def gold():
    return np.load(os.path.join(GOLDEN, "color_expr_golden.npz"))


def cases():
    return [str(c) for c in gold()["case_names"]]


def load_case(z, case, tmp_path, n=512):
    module = str(z[f"{case}_module"]); w, h = (int(v) for v in z[f"{case}_size"])
    d = tmp_path / case
    d.mkdir()
    (d / "rc.glsl").write_text(f"#request mod {module}\n#request setbufsize {n}\n#request setgeometry 0 0 {w} {h}\n")
    (d / (module + ".glsl")).write_text(str(z[f"{case}_config"]))
    return g.load_config([str(d)]), z[f"{case}_tl"], z[f"{case}_tr"], z[f"{case}_frame"]


def lsb(a, b):
    return int(np.abs(a.astype(int) - b.astype(int)).max())


@pytest.mark.parametrize("case", cases())
def test_compiled_colour_expressions_equal_the_reference_shaders(case, tmp_path, built):
    from tests import emul
    p, tl, tr, want = load_case(gold(), case, tmp_path)
    prog = {"bars": p.bars_color_prog, "radial": p.radial_color_prog, "graph": p.graph_color_prog}[p.module_name]
    col = {"bars": p.bars_color, "radial": p.radial_color, "graph": p.graph_color}[p.module_name]
    assert col.mode == 2 and 0 < prog.n_ops <= 64
    got = emul.raster(p, tl, tr)
    assert want.any() and lsb(got, want) <= 1, (case, lsb(got, want))
    if case not in ("radial_expr", "graph_pow"):                         # no transcendental: every operation is exactly rounded
        assert np.array_equal(got, want), (case, int((got != want).any(axis=2).sum()))
    if p.module_name in ("bars", "graph") and not p.bars_mirror_yx:
        assert np.array_equal(emul.raster(p, tl, tr, fast=True), got)    # the kernels' hoisted (row table) evaluation


def _cfg(tmp_path, module, text, name="c"):
    d = tmp_path / name
    d.mkdir(exist_ok=True)
    (d / "rc.glsl").write_text(f"#request mod {module}\n")
    (d / (module + ".glsl")).write_text(text)
    return [str(d)]




def test_shipped_and_constant_forms_keep_their_closed_form(tmp_path, built):
    p = g.load_config(_cfg(tmp_path, "bars", "#define GRADIENT 80\n#define COLOR mix(#3366b2, #a0a0b2, clamp(d / GRADIENT, 0.0, 1))\n"))
    assert p.bars_color.mode == 0 and p.bars_color.gradient == 80 and p.bars_color_prog.n_ops == 0
    # constant expressions are folded at config time (mode 1), also for the plain OUTLINE colours
    p = g.load_config(_cfg(tmp_path, "bars", "#define COLOR vec4(#804020.rgb * 0.5, 2 / 4)\n#define BAR_OUTLINE COLOR * 2\n", "k"))
    assert p.bars_color.mode == 1 and p.bars_outline_mode == 1 and p.bars_color_prog.n_ops == 0
    assert np.allclose(list(p.bars_color.lo), [0.501961 * 0.5, 0.250980 * 0.5, 0.125490 * 0.5, 0.0], atol=1e-7)   # 2 / 4 is integer division
    assert np.allclose(list(p.bars_outline), [0.501961, 0.250980, 0.125490, 0.0], atol=1e-7)
    p = g.load_config(_cfg(tmp_path, "circle", "#define OUTLINE mix(#000000, #ffffff, 0.25)\n", "o"))
    assert list(p.circle_outline) == [0.25, 0.25, 0.25, 1.0]
    # a clamp() with other bounds, or another variable, is no longer mistaken for the shipped gradient
    p = g.load_config(_cfg(tmp_path, "radial", "#define COLOR mix(#cc3333, #cca0a0, clamp(d / 95, 0.5, 1))\n", "r"))
    assert p.radial_color.mode == 2


def test_textual_macro_expansion_and_integer_folding(tmp_path, built):
    from tests import emul
    # `K` expands textually: 6 * 1 + 1 = 7 (not 6 * 2); 7 / 2 = 3 in integers; `pos` and `d` name the same variable in graph
    text = "#define K 1 + 1\n#define COLOR vec4(float(6 * K) / 8, pos / 4, d / 4.0, float(7 / 2) / 4)\n#define DRAW_HIGHLIGHT 0\n#define VSCALE 1000\n"
    d = tmp_path / "g"; d.mkdir()
    (d / "rc.glsl").write_text("#request mod graph\n#request setbufsize 256\n#request setgeometry 0 0 8 4\n")
    (d / "graph.glsl").write_text(text)
    p = g.load_config([str(d)])
    assert p.graph_color.mode == 2
    full = np.full(256, 65535, np.uint16)
    frame = emul.raster(p, full, full)
    x = 2                                                                # a column whose line height covers every row
    for y in range(3):
        assert frame[y, x].tolist() == [223, int(y / 4 * 255 + 0.5), int(y / 4 * 255 + 0.5), 191], (y, frame[y, x])


@pytest.mark.parametrize("text,msg", [
    ("#define COLOR vec4(texture(audio_l, d).r, 0, 0, 1)", "'audio_l' is not available"),
    ("#define COLOR vec4(length(vec2(d)), 0, 0, 1)", "function 'length' is not available"),
    ("#define COLOR vec4(v, 0, 0, 1)", "'v' is not available to a colour expression"),
    ("#define COLOR vec3(d, 0, 0)", "not a vec4"),
    ("#define COLOR vec4(d, 0, 0)", "component count mismatch"),
    ("#define COLOR vec4(1, 0, 0, 1) + vec2(d)", "different vector sizes"),
    ("#define COLOR vec4(d % 2, 0, 0, 1)", "unexpected character '%'"),
    ("#define COLOR vec4(1) > vec4(0) ? vec4(1) : vec4(0)", "'>' on a vector"),
    ("#define COLOR d > 1 ? vec4(1) : vec3(0)", "have different types"),
    ("#define COLOR vec4(d.y, 0, 0, 1)", "swizzle '.y' out of range"),
    ("#define COLOR vec4(1 / 0, d, 0, 1)", "integer division by zero"),
])
def test_unsupported_expressions_are_config_errors(tmp_path, text, msg, built):
    with pytest.raises(g.GlavaError, match="unsupported colour expression.*" + msg.replace("(", r"\(").replace("?", r"\?")):
        g.load_config(_cfg(tmp_path, "bars", text + "\n"))


def test_hand_built_params_are_validated(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("validated identically with a device; this variant checks the message without one")
    p = g.default_params("bars")
    p.bars_color.mode = 2                                                # mode 2 without a program
    with pytest.raises(g.GlavaError):
        g.Renderer(p, batch=1)


# ---- differential fuzz: random expressions through the config compiler + eval_color_prog vs the GLSL interpreter ---------------
class _Gen:
    """random, well-typed GLSL colour expressions over the variable `d`, restricted to operations that are exactly
    rounded in both implementations (no transcendentals) and to denominators / radicands that stay away from 0"""

    def __init__(self, rng):
        self.r = rng

    def lit(self):
        k = self.r.integers(0, 4)
        if k == 0:
            return str(int(self.r.integers(1, 9)))                      # int literal
        if k == 1:
            return "%.3f" % self.r.uniform(0.05, 4.0)
        if k == 2:
            return "%d.%d" % (self.r.integers(0, 3), self.r.integers(1, 10))
        return "%.2f" % self.r.uniform(0.1, 2.0)

    def scalar(self, depth):
        if depth <= 0:
            return "d" if self.r.random() < 0.4 else self.lit()
        k = self.r.integers(0, 13)
        a, b = self.scalar(depth - 1), self.scalar(depth - 1)
        if k == 0: return f"({a} + {b})"
        if k == 1: return f"({a} - {b})"
        if k == 2: return f"({a} * {b})"
        if k == 3: return f"({a} / (abs({b}) + 0.5))"
        if k == 4: return f"min({self.fl(a)}, {self.fl(b)})"
        if k == 5: return f"clamp({self.fl(a)}, 0.25, 3)"
        if k == 6: return f"mix({self.fl(a)}, {self.fl(b)}, 0.375)"
        if k == 7: return f"{self.r.choice(['floor', 'fract', 'abs', 'ceil', 'sign'])}({self.fl(a)})"
        if k == 8: return f"sqrt(abs({self.fl(a)}) + 0.125)"
        if k == 9: return f"{self.vec(int(self.r.integers(2, 5)), depth - 1)}.{self.r.choice(['x', 'g', 'y', 'r'])}"
        if k == 10: return f"smoothstep(0.5, 2.5, {self.fl(a)})"
        if self.r.random() < 0.5:
            c, e = self.scalar(depth - 1), self.scalar(depth - 1)
            cond = self.r.choice([f"{a} {self.r.choice(['<', '>', '<=', '>=', '==', '!='])} {b}",
                                  f"({a} < {b} && {c} >= 1) || !({e} > 2)", f"!({a} <= {b})", "true", "2 > 3"])
            return f"(({cond}) ? {self.fl(c)} : {self.fl(e)})"
        return f"mod({self.fl(a)}, 1.75)"

    def fl(self, e):
        """function arguments are given as floats (GLSL 3.30 has no int overloads of these)"""
        return f"float({e})"

    def vec(self, n, depth):
        if depth <= 0:
            if n == 4 and self.r.random() < 0.4:
                return "#%06x" % int(self.r.integers(0, 1 << 24))
            return f"vec{n}(" + ", ".join(self.scalar(0) for _ in range(n)) + ")"
        k = self.r.integers(0, 7)
        if k == 0: return f"({self.vec(n, depth - 1)} * {self.scalar(depth - 1)})"
        if k == 1: return f"({self.vec(n, depth - 1)} + {self.vec(n, depth - 1)})"
        if k == 2: return f"mix({self.vec(n, depth - 1)}, {self.vec(n, depth - 1)}, {self.fl(self.scalar(depth - 1))})"
        if k == 3 and n >= 3: return f"vec{n}({self.vec(n - 2, depth - 1) if n > 3 else self.scalar(depth - 1)}, {self.vec(2, depth - 1)})"
        if k == 4: return f"vec{n}({self.fl(self.scalar(depth - 1))})"
        if k == 5 and n < 4: return f"{self.vec(4, depth - 1)}.{'bgra'[:n] if self.r.random() < 0.5 else 'wzyx'[:n]}"
        if k == 6: return f"clamp({self.vec(n, depth - 1)}, 0.0, 1.5)"
        return f"vec{n}(" + ", ".join(self.scalar(depth - 1) for _ in range(n)) + ")"


def test_random_expressions_agree_with_the_glsl_interpreter(tmp_path, built):
    from oracle import glsl_interp as gi
    from tests import emul
    rng = np.random.default_rng(20260923)
    gen = _Gen(rng)
    xs = [0.0, 0.5, 1.5, 7.25, 31.5, 99.5]
    checked = 0
    for i in range(120):
        expr = gen.vec(4, int(rng.integers(1, 4)))
        d = tmp_path / f"e{i}"
        d.mkdir()
        (d / "rc.glsl").write_text("#request mod bars\n")
        (d / "bars.glsl").write_text(f"#define COLOR {expr}\n")
        try:
            p = g.load_config([str(d)])
        except g.GlavaError as e:
            assert "more than 8 live" in str(e) or "longer than 64" in str(e), (expr, str(e))   # resource limits only
            continue
        frag = d / "e.frag"
        frag.write_text("uniform float d;\nout vec4 fragment;\nvoid main() {\n    fragment = %s;\n}\n" % expr)
        sh = gi.load_stage(str(frag), str(d))
        for x in xs:
            want = np.array([float(v) for v in sh.run({"d": gi.F32(x)}, 0, 0)["fragment"].v], np.float32)
            got = emul.eval_color(p.bars_color_prog, x) if p.bars_color.mode == 2 else np.array(list(p.bars_color.lo), np.float32)
            assert np.array_equal(got, want), (expr, x, got, want)
        checked += 1
    assert checked >= 90


def test_exp_exp2_log2_pow_accuracy(tmp_path, built):
    """gl_math.h's exp / exp2 (and log2 = log / ln 2, pow = exp(y log x)) through the colour VM against float64: <= 1 ulp for exp and exp2, <= 2 for log2,
    a few tens of ulp for pow with large results (GLSL derives pow's precision from exp2 / log2 as well)"""
    from tests import emul
    p = g.load_config(_cfg(tmp_path, "bars", "#define COLOR vec4(exp(d), exp2(d), log2(abs(d) + 0.25), pow(abs(d) + 0.25, 2.5))\n"))
    assert p.bars_color.mode == 2
    xs = np.concatenate([np.linspace(-100, 88, 1501), np.linspace(-2, 2, 801), np.linspace(-126, 127, 300)]).astype(np.float32)
    worst = np.zeros(4)
    for x in xs:
        got = emul.eval_color(p.bars_color_prog, float(x))
        a = np.float32(np.float32(abs(x)) + np.float32(0.25))
        want = [np.exp(np.float64(x)), np.exp2(np.float64(x)), np.log2(np.float64(a)), np.float64(a) ** 2.5]
        for k in range(4):
            if np.isfinite(want[k]) and 1e-37 < abs(want[k]) < 3e38:
                worst[k] = max(worst[k], abs(float(got[k]) - float(np.float32(want[k]))) / float(np.spacing(np.float32(abs(want[k])))))
    assert worst[0] <= 1 and worst[1] <= 1 and worst[2] <= 2 and worst[3] <= 32, worst
    assert emul.eval_color(p.bars_color_prog, 200.0)[0] == np.inf and emul.eval_color(p.bars_color_prog, -200.0)[0] == 0.0


def test_tan_and_atan(tmp_path, built):
    from tests import emul
    p = g.load_config(_cfg(tmp_path, "bars", "#define COLOR vec4(tan(d), atan(d), atan(d, 2.0), atan(2.0, d))\n"))
    for x in np.linspace(-1.4, 1.4, 57):
        got = emul.eval_color(p.bars_color_prog, float(x)).astype(np.float64)
        want = np.array([np.tan(x), np.arctan(x), np.arctan2(x, 2.0), np.arctan2(2.0, x)])
        assert np.allclose(got, want, rtol=3e-6, atol=3e-7), (x, got, want)
