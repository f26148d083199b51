"""Raster regression fixture (tests/golden/raster_regression.npz, made by tests/golden/make_raster_regression.py from the
oracle's libm build).  It freezes the GLSL restatement — the reference ships no golden pixels to pin it to — for the
oracle (both maths builds), for the product's arithmetic compiled for the host, and (-m gpu) for the kernels."""
import os

import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import params_from
from tests.conftest import GOLDEN

W, H, N = 160, 92, 1024
MODULES = ("bars", "radial", "circle", "graph", "wave", "test")
OVER = dict(radial_radius=20.0, radial_amplify=40.0, circle_radius=18.0, circle_amplify=30.0, bars_amplify=70.0,
            graph_vscale=60.0, wave_amplify=80.0)


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "raster_regression.npz"))


def _lsb(a, b):
    return int(np.abs(a.astype(int) - b.astype(int)).max())


@pytest.mark.parametrize("module", MODULES)
def test_oracle_matches_the_frozen_frames(orc, orc_pm, fx, module, built):
    p = orc.default_params(module, n=N, w=W, h=H, **OVER)
    tl, tr, want = fx[f"{module}_tl"], fx[f"{module}_tr"], fx[f"{module}_frame"]
    assert want.shape == (H, W, 4) and want.any()
    assert np.array_equal(orc.raster(p, tl, tr), want)                     # same build the fixture came from: exact
    assert _lsb(orc_pm.raster(p, tl, tr), want) <= 1                       # product-maths build: transcendental ulps only


@pytest.mark.parametrize("module", MODULES)
def test_product_arithmetic_on_the_host_matches_the_frozen_frames(orc_pm, fx, module, built):
    from tests import emul
    p = g.default_params(module, n=N, w=W, h=H, **OVER)
    tl, tr, want = fx[f"{module}_tl"], fx[f"{module}_tr"], fx[f"{module}_frame"]
    got = emul.raster(p, tl, tr)
    assert np.array_equal(got, orc_pm.raster(params_from(p), tl, tr)) and _lsb(got, want) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("module", MODULES)
def test_kernels_match_the_frozen_frames(orc_pm, fx, module, built):
    p = g.default_params(module, n=N, w=W, h=H, **OVER)
    tl, tr, want = fx[f"{module}_tl"], fx[f"{module}_tr"], fx[f"{module}_frame"]
    with g.Renderer(p, batch=2) as r:
        r.raster_textures(np.stack([tl, tl]), np.stack([tr, tr]))
        for s in range(2):
            got = r.readback(s)
            assert np.array_equal(got, orc_pm.raster(params_from(p), tl, tr)) and _lsb(got, want) <= 1
