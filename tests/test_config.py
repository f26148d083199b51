"""`#request` / `#define` config surface (glava_b200/csrc/config.cpp) — no GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

import glava_b200 as g
from tests.conftest import GOLDEN

REF_SHADERS = "/root/reference/shaders/glava"
OWN_CONFIG = os.path.join(os.path.dirname(g.__file__), "config")


def _same(a, b, skip=()):
    for name, _ in g.Params._fields_:
        if name in skip:
            continue
        va, vb = getattr(a, name), getattr(b, name)
        if isinstance(va, (C.Structure, C.Array)):
            assert bytes(va) == bytes(vb), name
        else:
            assert va == vb, name


def test_defaults_match_shipped_values(built):
    p = g.default_params("bars")
    assert (p.n, p.avg_frames, p.avg_window, p.accel_fft, p.smooth_pass) == (4096, 5, 1, 1, 1)
    assert p.fft_scale == np.float32(10.2) and p.fft_cutoff == np.float32(0.3) and p.gravity_step == np.float32(4.2)
    assert p.ur == np.float32(22050 / 256)
    assert (p.w, p.h) == (800, 600) and p.channels == 2 and p.premultiply_alpha == 1
    assert p.bars_amplify == 300 and p.radial_nbars == 160 and p.circle_amplify == 150 and p.graph_vscale == 300
    # colour literals are the "%.6f" strings glsl_ext.c:505 emits: #3366b2 -> (0.200000, 0.400000, 0.698039)
    assert list(p.bars_color.lo)[:3] == [np.float32(0.2), np.float32(0.4), np.float32(0.698039)]


def test_unknown_module_is_an_error(built):
    with pytest.raises(g.GlavaError, match="Could not find module"):
        g.default_params("spiral")


def test_hex_colours_match_reference_parser(built):
    # golden: ext_parse_color (glsl_ext.c:88-122) run from the compiled reference
    gold = np.load(os.path.join(GOLDEN, "colors.npz"))
    for name, rgba in zip(gold["names"], gold["rgba"]):
        d = os.path.join(os.environ.get("TMPDIR", "/tmp"), "glava_b200_cfg_col")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "rc.glsl"), "w") as f:
            f.write("#request mod circle\n")
        lit = str(name)
        if lit.startswith("0x") or len(lit) == 8:
            continue                                  # OUTLINE takes #rrggbb; alpha / 0x forms are setbg-only
        with open(os.path.join(d, "circle.glsl"), "w") as f:
            f.write(f"#define OUTLINE #{lit}\n")
        p = g.load_config([d])
        want = [np.float32(f"{float(v):.6f}") for v in rgba[:3]]
        assert list(p.circle_outline)[:3] == want, lit


@pytest.mark.skipif(not os.path.isdir(REF_SHADERS), reason="reference tree not present")
@pytest.mark.parametrize("module", ["bars", "radial", "circle", "graph", "wave"])
def test_reference_shipped_config_parses_to_builtin_defaults(module, built):
    """the reference's own rc.glsl + smooth_parameters.glsl + <module>.glsl, unmodified"""
    p = g.load_config([REF_SHADERS], force_module=module)
    _same(p, g.default_params(module), skip=_wave_skip(module, p))


def _wave_skip(module, p):
    """wave/1.frag is the one module shader that does not include util/smooth.glsl, so rd_new never takes
    smooth_parameters.glsl's requests for it (they are ignored while util/smooth_pass.frag loads, render.c:1186-1215):
    setavgframes stays at rd_new's initialiser 6 (render.c:913) — a value the wave chain never uses.  Pinned against the
    real rd_new in tests/test_ref_rd.py."""
    if module != "wave":
        return ()
    assert p.avg_frames == 6
    return ("avg_frames",)


@pytest.mark.parametrize("module", ["bars", "radial", "circle", "graph", "wave"])
def test_own_config_dir(module, built):
    p = g.load_config([OWN_CONFIG], force_module=module)
    _same(p, g.default_params(module), skip=_wave_skip(module, p))


def test_requests_and_user_override(tmp_path, built):
    user, system = tmp_path / "user", tmp_path / "sys"
    user.mkdir(); system.mkdir()
    (system / "rc.glsl").write_text('#request mod bars\n')
    # the first path that has the entry is THE config dir (render.c:1328-1351); `glava --copy-config` puts rc.glsl there
    (user / "rc.glsl").write_text('#request mod radial\n#request setbufsize 2048\n#request setgeometry 0 0 1280 720\n'
                                  '#request setmirror true\n#request setopacity "none"\n')
    (system / "radial.glsl").write_text("#define NBARS 120\n#define C_LINE 3\n#define COLOR @fg:#ff0000\n#define ROTATE (PI / 4)\n")
    (user / "radial.glsl").write_text("/* user copy wins */\n#define NBARS 96 // trailing comment\n#define AMPLIFY 250.5F\n")
    (system / "smooth_parameters.glsl").write_text("#define SAMPLE_MODE hybrid\n#request setavgframes 7\n#request setgravitystep 3.5\n")
    p = g.load_config([str(user), str(system)], requests=["setsmoothfactor 0.05", 'setaccelfft false'])
    assert p.module_name == "radial" and p.n == 2048 and (p.w, p.h) == (1280, 720)
    assert p.channels == 1 and p.premultiply_alpha == 0
    assert p.radial_nbars == 96 and p.radial_amplify == np.float32(250.5)
    assert p.radial_line == 3 and p.radial_line_half == 1            # (C_LINE / 2) is integer division for `3`
    assert p.radial_color.mode == 1 and list(p.radial_color.lo) == [1.0, 0.0, 0.0, 1.0]
    assert p.radial_rotate == np.float32(np.float32(3.14159265359) / np.float32(4))
    assert p.sample_mode == 2 and p.avg_frames == 7 and p.gravity_step == np.float32(3.5)
    assert p.smooth_factor == np.float32(0.05) and p.accel_fft == 0
    # -m / force_module beats `#request mod`
    assert g.load_config([str(user), str(system)], force_module="wave").module_name == "wave"


def test_include_rules_and_config_dir_selection(tmp_path, built):
    """glsl_ext.c:161-183: plain includes are relative to the current dir, ":x" to the config dir (= where the entry
    was found), "@x" to the defaults dir (= last path); a path without the entry is not a config dir at all."""
    user, system = tmp_path / "user", tmp_path / "sys"
    user.mkdir(); system.mkdir(); (user / "extra").mkdir()
    (system / "rc.glsl").write_text("#request mod bars\n")
    (system / "bars.glsl").write_text("#define BAR_WIDTH 7\n#define AMPLIFY 111\n")
    (user / "bars.glsl").write_text('#include "extra/tweaks.glsl"\n#define BAR_GAP 3\n#include "@shared.glsl"\n#include ":late.glsl"\n')
    (user / "extra" / "tweaks.glsl").write_text("#define AMPLIFY 222\n#request setavgframes 3\n")
    (system / "shared.glsl").write_text("#define BAR_OUTLINE_WIDTH 2\n")
    (user / "late.glsl").write_text("#define BAR_WIDTH 9\n")
    # user dir has no rc.glsl: skipped entirely, the system dir is both config and defaults dir
    p = g.load_config([str(user), str(system)])
    # (no smooth_parameters.glsl anywhere: setavgframes stays at rd_new's initialiser 6, render.c:913)
    assert (p.bars_width, p.bars_gap, p.bars_amplify, p.avg_frames) == (7, 1, 111, 6)
    # with the entry copied to the user dir its files take part, includes resolved per the three rules
    (user / "rc.glsl").write_text('#request mod bars\n#include "more_rc.glsl"\n')
    (user / "more_rc.glsl").write_text("#request setbufsize 1024\n")
    p = g.load_config([str(user), str(system)])
    assert (p.bars_width, p.bars_gap, p.bars_amplify, p.bars_outline_width, p.avg_frames, p.n) == (9, 3, 222, 2, 3, 1024)
    # errors: missing include target; '@' / ':' inside rc.glsl, which has no defaults / config dir (render.c:1356-1361)
    (user / "rc.glsl").write_text('#request mod bars\n#include "nope.glsl"\n')
    with pytest.raises(g.GlavaError, match="#include directive"):
        g.load_config([str(user), str(system)])
    (user / "rc.glsl").write_text('#request mod bars\n#include "@shared.glsl"\n')
    with pytest.raises(g.GlavaError, match="no default directory"):
        g.load_config([str(user), str(system)])
    (user / "rc.glsl").write_text('#request mod bars\n#include ":late.glsl"\n')
    with pytest.raises(g.GlavaError, match="#include directive"):          # no config dir: the ':' stays in the file name
        g.load_config([str(user), str(system)])
    (user / "rc.glsl").write_text('#request mod bars\n#include\n')
    with pytest.raises(g.GlavaError, match="No arguments provided to #include"):
        g.load_config([str(user), str(system)])


@pytest.mark.parametrize("text,match", [
    ("#request frobnicate 1\n", "unknown request type 'frobnicate'"),
    ("#request setbufsize\n", "failed to execute request 'setbufsize'"),
    ("#request setmirror maybe\n", "invalid raw string into a boolean"),
    ('#request setopacity "shiny"\n', "Invalid opacity option"),
    ("#request mod spiral\n", "Could not find module 'spiral'"),
    ("#request setbufsize 3000\n", "power of two"),
])
def test_config_errors_use_the_reference_wording(tmp_path, text, match, built):
    (tmp_path / "rc.glsl").write_text(text)
    with pytest.raises(g.GlavaError, match=match):
        g.load_config([str(tmp_path)])


def test_missing_entry(tmp_path, built):
    with pytest.raises(g.GlavaError, match="Could not find entry point"):
        g.load_config([str(tmp_path)])


def test_unsupported_colour_expression(tmp_path, built):
    (tmp_path / "rc.glsl").write_text("#request mod bars\n")
    (tmp_path / "bars.glsl").write_text("#define COLOR vec4(texture(audio_l, d).r, 0, 0, 1)\n")
    with pytest.raises(g.GlavaError, match="unsupported colour expression"):
        g.load_config([str(tmp_path)])


def test_pipe_binds_resolve_at_macros(tmp_path, built):
    """`@name:default` (glsl_ext.c:516-591): bound value when `--pipe` bound the name, else the default"""
    (tmp_path / "rc.glsl").write_text("#request mod bars\n")
    (tmp_path / "bars.glsl").write_text("#define GRADIENT 80\n#define COLOR @fg:mix(#3366b2, #a0a0b2, clamp(d / GRADIENT, 0, 1))\n"
                                        "#define BAR_OUTLINE @bg:vec4(COLOR.rgb * 1.5, COLOR.a)\n")
    p = g.load_config([str(tmp_path)])
    assert p.bars_color.mode == 0 and p.bars_outline_mode == 0
    q = g.load_config([str(tmp_path)], binds={"fg": "#ff8000", "bg": "vec4(0.1, 0.2, 0.3, 1.0)"})
    assert q.bars_color.mode == 1 and list(q.bars_color.lo)[:3] == [1.0, np.float32(0.501961), 0.0]
    assert q.bars_outline_mode == 1 and list(q.bars_outline) == [np.float32(0.1), np.float32(0.2), np.float32(0.3), 1.0]
    with pytest.raises(g.GlavaError, match="name=value"):
        g.lib().glava_b200_load_config_binds  # symbol exists
        import ctypes as C
        arr = (C.c_char_p * 2)(b"novalue", None)
        pp = g.Params()
        rc = g.lib().glava_b200_load_config_binds(C.byref(pp), None, None, None, None, arr)
        if rc != 0:
            raise g.GlavaError(g.lib().glava_b200_last_error().decode())


def test_setbg_and_setbgf(tmp_path, built):
    """render.c:1062-1099: setbg takes hex digits without '#'; components the string does not reach keep their value"""
    assert list(g.load_config().clear_color) == [0, 0, 0, 0]
    p = g.load_config(requests=["setbg ff8000"])
    assert list(p.clear_color) == [1.0, np.float32(128 / 255), 0.0, 0.0]              # alpha stays at its default 0
    p = g.load_config(requests=["setbg 0x10203040"])
    assert np.allclose(list(p.clear_color), [16 / 255, 32 / 255, 48 / 255, 64 / 255], atol=1e-7)
    p = g.load_config(requests=["setbgf 0.25 0.5 0.75 1.0", 'setopacity "none"'])
    assert list(p.clear_color) == [0.25, 0.5, 0.75, 1.0] and p.premultiply_alpha == 0
    with pytest.raises(g.GlavaError, match="Invalid value for `setbg` request: 'zz0000'"):
        g.load_config(requests=["setbg zz0000"])


def test_conditionals_select_defines_but_not_requests(tmp_path, built):
    """#if / #ifdef / #elif / #else / #endif / #undef are the GLSL compiler's business: they pick among #defines, while
    glsl_ext.c runs `#request` (and `#include`) lines wherever they stand"""
    (tmp_path / "rc.glsl").write_text("#request mod graph\n")
    (tmp_path / "graph.glsl").write_text(
        "#define MODE 3\n#if MODE == 1\n#define VSCALE 100\n#elif MODE >= 3\n#define VSCALE 250\n#else\n#define VSCALE 1\n#endif\n"
        "#ifdef NOPE\n#define DRAW_OUTLINE 1\n#request setavgframes 9\n#else\n#define DRAW_OUTLINE 0\n#endif\n"
        "#define INVERT 1\n#undef INVERT\n#if defined(INVERT) || !defined(MODE)\n#define DIRECTION -1\n#endif\n")
    p = g.load_config([str(tmp_path)])
    assert p.graph_vscale == 250 and p.graph_draw_outline == 0 and p.graph_invert == 0 and p.graph_direction == 1
    assert p.avg_frames == 9                                             # the request inside the dead branch still ran
    (tmp_path / "graph.glsl").write_text("#if VSCALE >\n#define VSCALE 2\n#endif\n")
    with pytest.raises(g.GlavaError, match="cannot evaluate '#if VSCALE >'"):
        g.load_config([str(tmp_path)])
    (tmp_path / "graph.glsl").write_text("#define VSCALE 2\n#endif\n")
    with pytest.raises(g.GlavaError, match="#endif without #if"):
        g.load_config([str(tmp_path)])


def test_setsmoothfactor_is_the_six_decimal_literal_of_the_injected_header(built):
    """render.c:315-324: the shaders get `#define _SMOOTH_FACTOR %.6f`, not the request's float"""
    assert g.load_config(requests=["setsmoothfactor 0.0438713878"]).smooth_factor == np.float32("0.043871")
    assert g.load_config(requests=["setsmoothfactor 0.0123456789"]).smooth_factor == np.float32("0.012346")
    assert g.load_config().smooth_factor == np.float32(0.025)


def test_float_spelling_of_a_preprocessor_tested_macro_is_a_config_error(built, tmp_path):
    """`#if BAR_OUTLINE_WIDTH > 0` (bars/1.frag:116): GLSL's preprocessor takes integer expressions only — with a float the
    reference's shader does not compile (seen on Mesa llvmpipe) and GLava aborts; here: ECONFIG naming the macro"""
    d = tmp_path / "cfg"; d.mkdir()
    (d / "rc.glsl").write_text("#request mod bars\n")
    (d / "bars.glsl").write_text("#define BAR_OUTLINE_WIDTH 0.5\n")
    with pytest.raises(g.GlavaError, match="BAR_OUTLINE_WIDTH.*preprocessor"):
        g.load_config([str(d)])
    (d / "bars.glsl").write_text("#define BAR_OUTLINE_WIDTH 2\n#define USE_ALPHA 1\n")
    p = g.load_config([str(d)])
    assert p.bars_outline_width == 2.0          # USE_ALPHA is read and ignored: bars/2.frag never sees bars.glsl (llvmpipe golden)
    (d / "rc.glsl").write_text("#request mod graph\n")
    (d / "graph.glsl").write_text("#define DRAW_OUTLINE 1.0\n")
    with pytest.raises(g.GlavaError, match="DRAW_OUTLINE"):
        g.load_config([str(d)])


def test_setmirror_reaches_the_audio_side_even_when_bars_disables_mono(built, tmp_path):
    """setmirror sets r->mirror_input (the backend mixes to mono, fifo.c:98-102) AND the shader's _CHANNELS (render.c:1054-1058,
    :290); bars' DISABLE_MONO 1 turns only the shader back to two sides (bars/1.frag:32-34, llvmpipe golden `bars_disable_mono`)"""
    d = tmp_path / "cfg"; d.mkdir()
    (d / "rc.glsl").write_text("#request mod bars\n#request setmirror true\n")
    p = g.load_config([str(d)])
    assert (p.channels, p.mirror_input) == (1, 1)
    (d / "bars.glsl").write_text("#define DISABLE_MONO 1\n")
    p = g.load_config([str(d)])
    assert (p.channels, p.mirror_input) == (2, 1)
    assert g.default_params("bars").mirror_input == 0


def test_params_survive_a_json_round_trip(built):
    import json
    from glava_b200.api import Params
    p = g.default_params("radial", n=2048, w=321, h=123, radial_nbars=44)
    q = Params.from_dict(json.loads(json.dumps(p.to_dict())))
    assert bytes(p) == bytes(q)
