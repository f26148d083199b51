"""Golden frames for user COLOR / BAR_OUTLINE macros that are general GLSL expressions: oracle/glsl_interp.py runs the
REFERENCE'S OWN module shaders with the user's <module>.glsl pasted in, exactly as GLava would compile them.  The product
side compiles the same macro text into a colour program (glava_b200/csrc/color_compile.h).  Run in the build container:

    python tests/golden/make_color_expr_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import glsl_interp as gi  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SHADERS = "/root/reference/shaders/glava"
N = 512

# name -> (module, (w, h), <module>.glsl text)
CASES = {
    "bars_expr": ("bars", (96, 54), """
#define AMPLIFY 40
#define GRADIENT 30
#define K 1 + 1
#define COLOR vec4(mix(#ff8000, #2040ff, smoothstep(0, GRADIENT, d)).rgb * (0.5 + 0.5 * step(10, mod(d, 20))), 1)
#define BAR_OUTLINE vec4(COLOR.bgr * K / 4, COLOR.a - 0.25)
"""),
    "bars_expr_yx": ("bars", (94, 54), """
#define AMPLIFY 60
#define MIRROR_YX 1
#define BAR_OUTLINE_WIDTH 0
#define COLOR vec4(vec2(d / 64.0, 1.0 - d / 64.0), 0.25 * float(3), 1)
"""),
    "radial_expr": ("radial", (96, 54), """
#define C_RADIUS 12
#define AMPLIFY 30
#define NBARS 40
#define GRADIENT 14
#define COLOR vec4(abs(sin(d / 7.0)), clamp(1 - d / GRADIENT, 0.2, 1), fract(d * 0.125), 1.0)
"""),
    "graph_pow": ("graph", (96, 54), """
#define VSCALE 42
#define DRAW_HIGHLIGHT 0
#define COLOR vec4(pow(pos / 30, 0.5), exp(-pos / 20), exp2(-pos / 16) * (pos > 12 ? 1.0 : 0.5), log2(pos + 2) / 5)
"""),
    "graph_expr": ("graph", (96, 54), """
#define VSCALE 42
#define GRADIENT 25
#define DRAW_OUTLINE 1
#define OUTLINE vec4(#ff00ff.rgb * 0.5, 1)
#define COLOR mix(vec4(#802A2A.rgb, 0.75), vec4(0.31, 0.31, 0.57, 1), min(sqrt(pos / GRADIENT), 1)) * vec4(vec2(1), max(0.5, floor(d / 10) / 4), 1)
"""),
}


def textures(orc, p, seed):
    rng = np.random.default_rng(seed)
    return (orc.smooth_pass(p, (rng.random(N) ** 2 * 65535).astype(np.uint16)),
            orc.smooth_pass(p, (rng.random(N) ** 3 * 65535).astype(np.uint16)))


def main():
    orc = Oracle("libm")
    out = {"case_names": np.array(sorted(CASES))}
    for i, name in enumerate(sorted(CASES)):
        module, (w, h), text = CASES[name]
        p = orc.default_params(module, n=N, w=w, h=h)
        tl, tr = textures(orc, p, 300 + i)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, module + ".glsl"), "w").write(text)
            prog = gi.ModuleProgram(SHADERS, module, w, h, tl, tr, config_dir=d)
            frame = np.zeros((h, w, 4), np.uint8)
            for y in range(h):
                for x in range(w):
                    frame[y, x] = prog.pixel(x, y)
        out[f"{name}_module"] = np.array(module); out[f"{name}_size"] = np.array([w, h]); out[f"{name}_config"] = np.array(text)
        out[f"{name}_tl"] = tl; out[f"{name}_tr"] = tr; out[f"{name}_frame"] = frame
        print(name, "lit", int(frame.any(axis=2).sum()), "colours", len(np.unique(frame.reshape(-1, 4), axis=0)), flush=True)
    np.savez_compressed(os.path.join(HERE, "color_expr_golden.npz"), **out)


if __name__ == "__main__":
    main()
