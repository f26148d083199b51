"""Golden pixels computed FROM THE REFERENCE'S OWN SHADER SOURCES: oracle/glsl_interp.py reads
/root/reference/shaders/glava/<module>/<n>.frag (+ util/*.frag, <module>.glsl, smooth_parameters.glsl), applies GLava's
source extensions and injected header, and evaluates every fragment of a small surface in float32.  The frames written
here pin the C restatement (oracle/glava_oracle.c), the product arithmetic and the kernels to the reference's shader
text on machines where neither the reference tree nor OpenGL exists.  Run in the build container only:

    python tests/golden/make_glsl_golden.py

Cases: per module the shipped configuration scaled to the small surface, plus one variant exercising the module's
options.  Each case stores the GLSL macro overrides, the equivalent parameter overrides (JSON), the two R16 textures and
the interpreter's RGBA8 frame.  util passes: smooth_pass.frag (K5) in the three SAMPLE_MODEs, gravity_pass.frag,
average_pass.frag (windowed, 5 and 3 frames; unwindowed 2 frames) and pass.frag on small R16 textures.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import glsl_interp as gi  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHADERS = "/root/reference/shaders/glava"
W, H, N = 96, 54, 512

# (case name, module, GLSL macro overrides, parameter overrides, header overrides)
CASES = [
    ("bars", "bars", {"AMPLIFY": "40"}, dict(bars_amplify=40.0), {}),
    ("bars_opts", "bars", {"AMPLIFY": "45", "DIRECTION": "1", "INVERT": "1", "FLIP": "1", "BAR_WIDTH": "4", "BAR_GAP": "2", "BAR_OUTLINE_WIDTH": "0"},
     dict(bars_amplify=45.0, bars_direction=1, bars_invert=1, bars_flip=1, bars_width=4.0, bars_gap=2.0, bars_outline_width=0.0), {}),
    ("bars_mono", "bars", {"AMPLIFY": "40"}, dict(bars_amplify=40.0, channels=1), {"channels": 1}),
    ("radial", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "40"}, dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=40), {}),
    ("radial_opts", "radial", {"C_RADIUS": "10", "AMPLIFY": "28", "NBARS": "24", "INVERT": "1", "C_LINE": "3", "BAR_WIDTH": "3.5",
                               "CENTER_OFFSET_X": "5", "CENTER_OFFSET_Y": "-3", "ROTATE": "(PI / 3)"},
     dict(radial_radius=10.0, radial_amplify=28.0, radial_nbars=24, radial_invert=1, radial_line=3.0, radial_line_half=1.0,
          radial_bar_width=3.5, radial_off_x=5.0, radial_off_y=-3.0, radial_rotate=float(np.float32(np.float32(3.14159265359) / np.float32(3)))), {}),
    ("circle", "circle", {"C_RADIUS": "12", "AMPLIFY": "20"}, dict(circle_radius=12.0, circle_amplify=20.0), {}),
    ("circle_big", "circle", {"C_RADIUS": "22", "AMPLIFY": "30", "C_LINE": "2.5", "INVERT": "1"},
     dict(circle_radius=22.0, circle_amplify=30.0, circle_line=2.5, circle_invert=1), {}),
    ("circle_fill", "circle", {"C_RADIUS": "10", "AMPLIFY": "18", "C_FILL": "1", "C_SMOOTH": "0"},
     dict(circle_radius=10.0, circle_amplify=18.0, circle_fill=1, circle_smooth=0), {}),
    ("graph", "graph", {"VSCALE": "40"}, dict(graph_vscale=40.0), {}),
    ("graph_opts", "graph", {"VSCALE": "45", "INVERT": "1", "DIRECTION": "-1", "DRAW_OUTLINE": "1", "DRAW_HIGHLIGHT": "0"},
     dict(graph_vscale=45.0, graph_invert=1, graph_direction=-1, graph_draw_outline=1, graph_draw_highlight=0), {}),
    ("wave", "wave", {"AMPLIFY": "40"}, dict(wave_amplify=40.0), {}),
    ("wave_thick", "wave", {"AMPLIFY": "25", "MIN_THICKNESS": "2", "MAX_THICKNESS": "4"},
     dict(wave_amplify=25.0, wave_min_thickness=2.0, wave_max_thickness=4.0), {}),
    ("test", "test", {}, {}, {}),
    ("bars_mirror_yx", "bars", {"AMPLIFY": "60", "MIRROR_YX": "1"}, dict(bars_amplify=60.0, bars_mirror_yx=1), {}),
    # setsmoothpass false: the module shader itself runs smooth_audio()'s tap loop per fragment (_PRE_SMOOTHED_AUDIO 0)
    ("bars_nosmoothpass", "bars", {"AMPLIFY": "40"}, dict(bars_amplify=40.0, smooth_pass=0), {"pre_smoothed": 0}),
    ("radial_nosmoothpass", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "40"},
     dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=40, smooth_pass=0), {"pre_smoothed": 0}),
    ("bars_mono_invert", "bars", {"AMPLIFY": "40", "INVERT": "1"}, dict(bars_amplify=40.0, channels=1, bars_invert=1), {"channels": 1}),
    ("graph_both", "graph", {"VSCALE": "42", "DRAW_OUTLINE": "1", "DRAW_HIGHLIGHT": "1"},
     dict(graph_vscale=42.0, graph_draw_outline=1, graph_draw_highlight=1), {}),
    # odd surface sizes: screen.x / 2 and screen.y / 2 are INTEGER divisions in the shaders
    ("bars_odd", "bars", {"AMPLIFY": "40"}, dict(bars_amplify=40.0), {}, (95, 53)),
    ("radial_odd", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "40"}, dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=40), {}, (95, 53)),
    ("circle_odd", "circle", {"C_RADIUS": "12", "AMPLIFY": "20"}, dict(circle_radius=12.0, circle_amplify=20.0), {}, (93, 51)),
    ("graph_odd", "graph", {"VSCALE": "40"}, dict(graph_vscale=40.0), {}, (95, 53)),
    ("wave_odd", "wave", {"AMPLIFY": "40"}, dict(wave_amplify=40.0), {}, (95, 53)),
    # graph/3.frag: column-walking anti-alias stage (graph/4.frag never runs: its `#if ANTI_ALIAS == 0` sees an undefined macro)
    ("graph_aa", "graph", {"VSCALE": "42", "ANTI_ALIAS": "1"}, dict(graph_vscale=42.0, graph_anti_alias=1), {}),
    ("graph_aa_invert", "graph", {"VSCALE": "38", "ANTI_ALIAS": "1", "INVERT": "1", "DRAW_OUTLINE": "1"},
     dict(graph_vscale=38.0, graph_anti_alias=1, graph_invert=1, graph_draw_outline=1), {}),
    ("graph_join", "graph", {"VSCALE": "42", "JOIN_CHANNELS": "1"}, dict(graph_vscale=42.0, graph_join_channels=1), {}),
    # radial: `BAR_WIDTH / 2` is an integer division for an integer BAR_WIDTH; the deprecated bar outline (sides + end cap)
    ("radial_intwidth", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "24", "BAR_WIDTH": "5"},
     dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=24, radial_bar_width=5.0, radial_bar_width_int=1), {}),
    ("radial_outline", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "24", "BAR_WIDTH": "5.5", "BAR_OUTLINE_WIDTH": "1", "BAR_OUTLINE": "vec4(0.125490, 1.0, 0.250980, 1.0)"},
     dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=24, radial_bar_width=5.5, radial_bar_outline_width=1.0,
          radial_bar_outline=[0.125490, 1.0, 0.250980, 1.0]), {}),
    # setopacity "none": every stage blended (SRC_ALPHA, ONE_MINUS_SRC_ALPHA) over the glClear colour, premultiply stages skipped
    ("radial_nopremult", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "40"},
     dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=40, premultiply_alpha=0), {"premultiply_alpha": 0}),
    ("radial_blend", "radial", {"C_RADIUS": "12", "AMPLIFY": "30", "NBARS": "40"},
     dict(radial_radius=12.0, radial_amplify=30.0, radial_nbars=40, premultiply_alpha=0, clear_color=[0.1, 0.3, 0.2, 0.6]),
     {"premultiply_alpha": 0, "clear_color": (0.1, 0.3, 0.2, 0.6)}),
    ("bars_blend", "bars", {"AMPLIFY": "40"}, dict(bars_amplify=40.0, premultiply_alpha=0, clear_color=[0.2, 0.4, 0.6, 0.5]),
     {"premultiply_alpha": 0, "clear_color": (0.2, 0.4, 0.6, 0.5)}),
    ("circle_blend", "circle", {"C_RADIUS": "12", "AMPLIFY": "20"}, dict(circle_radius=12.0, circle_amplify=20.0, premultiply_alpha=0, clear_color=[0.5, 0.25, 0.75, 0.0]),
     {"premultiply_alpha": 0, "clear_color": (0.5, 0.25, 0.75, 0.0)}),
    ("circle_blend_opaque", "circle", {"C_RADIUS": "12", "AMPLIFY": "20"}, dict(circle_radius=12.0, circle_amplify=20.0, premultiply_alpha=0, clear_color=[0.0, 0.0, 0.5, 1.0]),
     {"premultiply_alpha": 0, "clear_color": (0.0, 0.0, 0.5, 1.0)}),
    ("graph_blend", "graph", {"VSCALE": "42", "DRAW_OUTLINE": "1", "DRAW_HIGHLIGHT": "1"},
     dict(graph_vscale=42.0, graph_draw_outline=1, graph_draw_highlight=1, premultiply_alpha=0, clear_color=[0.3, 0.3, 0.3, 0.0]),
     {"premultiply_alpha": 0, "clear_color": (0.3, 0.3, 0.3, 0.0)}),
    ("wave_blend", "wave", {"AMPLIFY": "40"}, dict(wave_amplify=40.0, premultiply_alpha=0, clear_color=[0.9, 0.8, 0.1, 0.25]),
     {"premultiply_alpha": 0, "clear_color": (0.9, 0.8, 0.1, 0.25)}),
    ("test_blend", "test", {}, dict(premultiply_alpha=0, clear_color=[0.0, 1.0, 0.0, 0.5]), {"premultiply_alpha": 0, "clear_color": (0.0, 1.0, 0.0, 0.5)}),
]


def textures(orc, p, module, seed):
    rng = np.random.default_rng(seed)
    tl = orc.smooth_pass(p, (rng.random(N) ** 2 * 65535).astype(np.uint16))
    tr = orc.smooth_pass(p, (rng.random(N) ** 3 * 65535).astype(np.uint16))
    if module == "wave":
        tl = np.clip(tl.astype(int) // 4 + 24576, 0, 65535).astype(np.uint16)
    return tl, tr


def main():
    orc = Oracle("libm")
    out = {"case_names": np.array([c[0] for c in CASES])}
    for i, case in enumerate(CASES):
        name, module, gl_over, p_over, hdr = case[:5]
        w, h = case[5] if len(case) > 5 else (W, H)
        p = orc.default_params(module, n=N, w=w, h=h)
        tl, tr = textures(orc, p, module, 100 + i)
        prog = gi.ModuleProgram(SHADERS, module, w, h, tl, tr, overrides=gl_over, **hdr)
        frame = np.zeros((h, w, 4), np.uint8)
        for y in range(h):
            for x in range(w):
                frame[y, x] = prog.pixel(x, y)
        out[f"{name}_size"] = np.array([w, h])
        out[f"{name}_module"] = np.array(module)
        out[f"{name}_params"] = np.array(json.dumps(p_over))
        out[f"{name}_tl"] = tl; out[f"{name}_tr"] = tr; out[f"{name}_frame"] = frame
        print(name, "stages", len(prog.stages), "lit", int(frame.any(axis=2).sum()), flush=True)
    # ---- util passes on 1-D R16 targets ---------------------------------------------------------------------------------
    n = 256
    rng = np.random.default_rng(7)
    tex = (rng.random(n) ** 2 * 65535).astype(np.uint16)
    out["k5_in"] = tex
    util = os.path.join(SHADERS, "util")
    for mode in ("average", "maximum", "hybrid"):
        for formula in ("sinusoidal", "linear", "circular"):
            if mode != "average" and formula != "sinusoidal":
                continue
            sh_over = {"SAMPLE_MODE": mode, "ROUND_FORMULA": formula}
            sh = gi.load_stage(os.path.join(util, "smooth_pass.frag"), SHADERS, sh_over)
            res = np.zeros(n, np.uint16)
            for x in range(n):
                g = sh.run({"tex": gi.Sampler1D(tex), "sz": n, "w": n}, x, 0)
                res[x] = gi.unorm16(g["fragment"].v[0])
            out[f"k5_{mode}_{formula}"] = res
            print("k5", mode, formula, flush=True)
    diff = np.float32(4.2) * (np.float32(1.0) / np.float32(86.1328125))
    out["gravity_diff"] = np.float32(diff)
    out["gravity_out"] = gi.run_1d_pass(os.path.join(util, "gravity_pass.frag"), SHADERS, n, {"tex": gi.Sampler1D(tex), "diff": diff})
    out["pass_out"] = gi.run_1d_pass(os.path.join(util, "pass.frag"), SHADERS, n, {"tex": gi.Sampler1D(tex)})
    frames = [(rng.random(n) ** 2 * 65535).astype(np.uint16) for _ in range(5)]
    out["avg_frames_in"] = np.stack(frames)
    for F, win in ((5, 1), (3, 1), (2, 1), (5, 0)):
        uni = {f"t{i}": gi.Sampler1D(frames[i]) for i in range(F)}                 # t0 = most recent (render.c:2250-2255)
        out[f"avg_F{F}_w{win}"] = gi.run_1d_pass(os.path.join(util, "average_pass.frag"), SHADERS, n, uni, avg_frames=F, avg_window=win)
    np.savez_compressed(os.path.join(HERE, "glsl_golden.npz"), **out)
    print("wrote glsl_golden.npz")


if __name__ == "__main__":
    main()
