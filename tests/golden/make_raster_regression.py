"""Regression fixture for the RASTER half: frames of the C restatement (oracle, libm build) for every module at a small
size, from seeded textures.  This fixture does NOT pin the restatement to the reference (tests/golden/llvmpipe_golden.npz
does: the reference itself on Mesa llvmpipe) — it freezes the restatement's exact libm-build output, so
that neither the oracle nor the kernels can drift unnoticed between rounds.  Regenerate only together with a deliberate
change of the GLSL semantics (round 2 did: unorm rounding and blend arithmetic became llvmpipe's):

    python tests/golden/make_raster_regression.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, N = 160, 92, 1024
MODULES = ("bars", "radial", "circle", "graph", "wave", "test")


def textures(orc, p, module):
    rng = np.random.default_rng(2026)
    tl = orc.smooth_pass(p, (rng.random(N) ** 2 * 65535).astype(np.uint16))
    tr = orc.smooth_pass(p, (rng.random(N) ** 3 * 65535).astype(np.uint16))
    if module == "wave":
        tl = np.clip(tl.astype(int) // 4 + 24576, 0, 65535).astype(np.uint16)
    return tl, tr


def main():
    orc = Oracle("libm")
    out = {}
    for m in MODULES:
        over = dict(radial_radius=20.0, radial_amplify=40.0, circle_radius=18.0, circle_amplify=30.0, bars_amplify=70.0,
                    graph_vscale=60.0, wave_amplify=80.0)
        p = orc.default_params(m, n=N, w=W, h=H, **over)
        tl, tr = textures(orc, p, m)
        out[f"{m}_tl"] = tl; out[f"{m}_tr"] = tr
        out[f"{m}_frame"] = orc.raster(p, tl, tr)
    np.savez_compressed(os.path.join(HERE, "raster_regression.npz"), **out)
    print("wrote raster_regression.npz", {m: int(out[f"{m}_frame"].astype(bool).any(axis=2).sum()) for m in MODULES})


if __name__ == "__main__":
    main()
