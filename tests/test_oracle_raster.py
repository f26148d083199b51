"""Raster half of the oracle: the reference's only in-tree known answer (the `test` module must
render #55000055, shaders/glava/test_rc.glsl:27, render.c:2420-2453) and structural checks of
the GLSL restatement."""
import numpy as np
import pytest

from oracle.oracle import MODULES


def _tex(n, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return (rng.random(n) ** 2 * 65535 * scale).astype(np.uint16)


def test_test_module_known_answer(orc):
    # test_rc.glsl: 640x640, settesteval 55000055, margin 1/(255*2) (render.c:2425)
    p = orc.default_params("test", n=4096, w=640, h=640)
    img = orc.raster(p, _tex(4096), _tex(4096, 1))
    assert np.all(img == np.array([0x55, 0x00, 0x00, 0x55], dtype=np.uint8))


def test_bars_geometry(orc):
    n, w, h = 4096, 1920, 1080
    p = orc.default_params("bars", n=n, w=w, h=h)
    tl = np.full(n, 32768, np.uint16); tr = np.full(n, 16384, np.uint16)
    img = orc.raster(p, tl, tr, rows=(0, 200))
    a = img[:200, :, 3]
    # section = 6 px: 5 px bar + 1 px gap; left half from audio_l (v = 0.5*300), right from audio_r
    row = a[10]
    assert row[960 - 6:960].tolist() == [255, 255, 255, 255, 255, 0] or row[954:960].sum() == 5 * 255
    # 160 sections per side; the outermost one has p = 1 + 3.5/1920 > 1 and is discarded (bars/1.frag:85-88)
    assert (row[:960] > 0).sum() == 159 * 5 and (row[960:] > 0).sum() == 159 * 5
    hl = (a[:, 100 + (0 if a[0, 100] else 1)] > 0).sum()
    hr = (a[:, 1800 + (0 if a[0, 1800] else 1)] > 0).sum()
    assert abs(hl - 150) <= 1 and abs(hr - 75) <= 1
    # colour: COLOR at d=0.5 ~ #3366b2, outline = rgb * 1.5
    x = int(np.argmax(row[:20] > 0))
    assert img[0, x + 2, :3].tolist() == pytest.approx([0x33, 0x66, 0xb2], abs=2)


def test_bars_mirror_is_transpose(orc):
    n = 1024
    p = orc.default_params("bars", n=n, w=192, h=192)
    q = orc.default_params("bars", n=n, w=192, h=192, bars_mirror_yx=1)
    tl, tr = _tex(n, 3), _tex(n, 4)
    a = orc.raster(p, tl, tr); b = orc.raster(q, tl, tr)
    assert np.array_equal(a, b.transpose(1, 0, 2))


@pytest.mark.parametrize("module", ["radial", "circle", "graph", "wave"])
def test_modules_draw_something_and_silence_is_flat(orc, module):
    n, w, h = 2048, 640, 360
    p = orc.default_params(module, n=n, w=w, h=h)
    loud = orc.raster(p, _tex(n, 5), _tex(n, 6))
    assert 0.001 < (loud[..., 3] > 0).mean() < 0.9
    silent_tex = np.zeros(n, np.uint16) if module != "wave" else np.full(n, 32768, np.uint16)
    quiet = orc.raster(p, silent_tex, silent_tex)
    assert (quiet[..., 3] > 0).mean() < (loud[..., 3] > 0).mean() + 1e-9
    if module == "graph":
        assert not quiet.any()
    if module == "wave":
        ys = np.nonzero(quiet[..., 3].any(axis=1))[0]
        assert ys.min() >= h // 2 - 4 and ys.max() <= h // 2 + 4        # flat line + outline at mid height


def test_premultiply_stage_and_blending_over_transparent_black(orc):
    # native: radial stage 2 multiplies rgb by the alpha of the 8-bit quantised stage 1 (premultiply.frag:12-15).
    # setopacity "none" with the default clear colour 00000000: stage 1 is blended SRC_ALPHA / ONE_MINUS_SRC_ALPHA over
    # transparent black (render.c:1467-1470) — rgb * a as well (from the unquantised fragment), alpha a * a; stage 2 skipped.
    n = 1024
    p = orc.default_params("radial", n=n, w=400, h=400)
    q = orc.default_params("radial", n=n, w=400, h=400, premultiply_alpha=0)
    tl, tr = _tex(n, 7), _tex(n, 8)
    a = orc.raster(p, tl, tr).astype(np.float32); b = orc.raster(q, tl, tr).astype(np.float32)
    assert a.any() and np.abs(a[..., :3] - b[..., :3]).max() <= 1
    assert np.abs(b[..., 3] - np.floor((a[..., 3] / 255) ** 2 * 255 + 0.5)).max() <= 1
    # an opaque clear colour shows wherever nothing is drawn, and under translucent fragments
    q2 = orc.default_params("radial", n=n, w=400, h=400, premultiply_alpha=0, clear_color=[0.0, 0.0, 1.0, 1.0])
    c = orc.raster(q2, tl, tr)
    # (alpha uses the same factors: a * a + 1 * (1 - a) dips to 0.75 under half-transparent edge fragments)
    assert (c[..., 3] >= 191).all() and (c[..., 3] < 255).any() and (c[0, 0] == [0, 0, 255, 255]).all() and (c[..., 0] > 0).any()


def test_rows_api_matches_full_frame(orc):
    n = 1024
    for module in MODULES:
        p = orc.default_params(module, n=n, w=128, h=96)
        tl, tr = _tex(n, 9), _tex(n, 10)
        full = orc.raster(p, tl, tr)
        part = orc.raster(p, tl, tr, rows=(31, 64))
        assert np.array_equal(full[31:64], part[31:64])


def test_libm_and_product_math_agree_within_1lsb(orc, orc_pm):
    n = 2048
    tl, tr = orc.smooth_pass(orc.default_params("bars", n=n), _tex(n, 11)), _tex(n, 12)
    assert np.abs(orc.smooth_pass(orc.default_params("bars", n=n), tr).astype(int)
                  - orc_pm.smooth_pass(orc_pm.default_params("bars", n=n), tr).astype(int)).max() <= 1
    for module in ("radial", "circle"):
        p = orc.default_params(module, n=n, w=500, h=400)
        a = orc.raster(p, tl, tr).astype(int); b = orc_pm.raster(p, tl, tr).astype(int)
        diff = np.abs(a - b).max(axis=2)
        # transcendental implementations differ by a few ulp: at most 1 LSB, except a handful of
        # pixels sitting exactly on a hard edge (threshold flips)
        assert (diff > 1).sum() <= 8, (module, (diff > 1).sum())
