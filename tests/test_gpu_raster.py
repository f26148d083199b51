"""-m gpu: module raster kernels through the C ABI, bit-exact against the oracle on identical textures."""
import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import OracleChannel, params_from

pytestmark = pytest.mark.gpu


def _textures(orc, op, n, batch, seed):
    rng = np.random.default_rng(seed)
    tl = np.stack([orc.smooth_pass(op, (rng.random(n) ** 2 * 65535).astype(np.uint16)) for _ in range(batch)])
    tr = np.stack([orc.smooth_pass(op, (rng.random(n) ** 3 * 65535).astype(np.uint16)) for _ in range(batch)])
    if batch > 2:
        tl[1] = 0; tr[1] = 0                       # silence
        tl[2] = 65535; tr[2] = 65535               # saturated R16
    return tl, tr


def _check(orc, p, batch=3, seed=0):
    op = params_from(p)
    tl, tr = _textures(orc, op, p.n, batch, seed)
    if p.module == 4:
        tl = np.clip(tl.astype(int) // 4 + 24576, 0, 65535).astype(np.uint16)      # PCM-like values around 0.5
    with g.Renderer(p, batch=batch) as r:
        r.raster_textures(tl, tr)
        for s in range(batch):
            got = r.readback(s)
            want = orc.raster(op, tl[s], tr[s])
            assert np.array_equal(got, want), (p.module_name, s, int((got != want).any(axis=2).sum()))


@pytest.mark.parametrize("module", g.MODULES)
@pytest.mark.parametrize("w,h", [(640, 360), (332, 97), (1920, 1080)])
def test_modules_bit_exact(orc_pm, module, w, h, built):
    _check(orc_pm, g.default_params(module, n=2048 if w < 1920 else 4096, w=w, h=h), batch=3 if w < 1920 else 2)


@pytest.mark.parametrize("w", [333, 331, 6])
def test_widths_not_multiple_of_four(orc_pm, w, built):
    for module in ("bars", "graph", "wave", "radial"):
        _check(orc_pm, g.default_params(module, n=1024, w=w, h=40), batch=1)


@pytest.mark.parametrize("over", [dict(bars_direction=1), dict(bars_invert=1), dict(bars_flip=1), dict(bars_mirror_yx=1),
                                  dict(channels=1), dict(channels=1, bars_invert=1), dict(bars_outline_width=0.0),
                                  dict(bars_width=3.0, bars_gap=2.0), dict(smooth_pass=0)])
def test_bars_options(orc_pm, over, built):
    _check(orc_pm, g.default_params("bars", n=1024, w=320, h=200, **over), batch=2)


# (non-native opacity — premultiply_alpha = 0, GL blending over the clear colour — is covered by tests/test_zz_gpu_blend.py)
@pytest.mark.parametrize("module,over", [("radial", dict(radial_invert=1)),
                                         ("radial", dict(radial_off_x=40.0, radial_off_y=-25.0)),
                                         ("circle", dict(circle_fill=1)), ("circle", dict(circle_smooth=0)),
                                         ("circle", dict(circle_invert=1)), ("graph", dict(graph_direction=-1)),
                                         ("graph", dict(graph_invert=1)), ("graph", dict(graph_draw_outline=1)),
                                         ("graph", dict(graph_draw_highlight=0))])
def test_module_options(orc_pm, module, over, built):
    _check(orc_pm, g.default_params(module, n=1024, w=400, h=300, **over), batch=2)


def test_reference_known_answer_on_gpu(built):
    """the reference's only in-tree KAT: module `test` renders uniform #55000055 (test_rc.glsl:27)"""
    p = g.default_params("test", n=4096, w=640, h=640)
    rings = g.StreamRings(1, 4096)
    with g.Renderer(p, batch=1) as r:
        rings.advance(); r.update(rings.lb, rings.rb, True)
        img = r.readback(0)
    assert np.all(img == np.array([0x55, 0, 0, 0x55], np.uint8))


def test_libm_oracle_within_one_lsb(orc, built):
    """independent checker (libm transcendentals): <= 1 LSB per channel, bar a few hard-edge flips"""
    for module in ("radial", "circle"):
        p = g.default_params(module, n=2048, w=640, h=360); op = params_from(p)
        tl, tr = _textures(orc, op, 2048, 1, 9)
        with g.Renderer(p, batch=1) as r:
            r.raster_textures(tl, tr); got = r.readback(0).astype(int)
        want = orc.raster(op, tl[0], tr[0]).astype(int)
        assert (np.abs(got - want).max(axis=2) > 1).sum() <= 8


def test_fb_slots_ring_and_rerender(orc_pm, built):
    n, batch = 1024, 5
    p = g.default_params("bars", n=n, w=128, h=64, fb_slots=2); op = params_from(p)
    tl, tr = _textures(orc_pm, op, n, batch, 4)
    with g.Renderer(p, batch=batch) as r:
        r.raster_textures(tl, tr)
        # slot = stream % 2: the last stream mapped to each slot is what it holds
        assert np.array_equal(r.readback(4), orc_pm.raster(op, tl[4], tr[4]))
        assert np.array_equal(r.readback(3), orc_pm.raster(op, tl[3], tr[3]))


def test_modified_false_rerasters_last_spectrum(built):
    n = 1024
    p = g.default_params("graph", n=n, w=256, h=128)
    rings = g.StreamRings(2, n)
    with g.Renderer(p, batch=2) as r:
        for _ in range(6):
            rings.advance(); r.update(rings.lb, rings.rb, True)
        a = r.readback(1); t0 = r.textures()
        rings.advance(); r.update(rings.lb, rings.rb, False)          # no audio update: render.c:2268-2272
        b = r.readback(1); t1 = r.textures()
    assert np.array_equal(a, b) and np.array_equal(t0[0], t1[0])


@pytest.mark.parametrize("module", ["radial", "circle"])
def test_polar_geometry_cache_equals_direct_evaluation(orc_pm, module, monkeypatch, built):
    """radial / circle normally run from the per-renderer geometry cache; GLAVA_B200_NO_GEO=1 selects the
    kernels that evaluate the shader maths per pixel.  Both must give the oracle's pixels."""
    p = g.default_params(module, n=2048, w=900, h=700); op = params_from(p)
    tl, tr = _textures(orc_pm, op, 2048, 2, 21)
    frames = []
    for no_geo in ("", "1"):
        if no_geo:
            monkeypatch.setenv("GLAVA_B200_NO_GEO", "1")
        else:
            monkeypatch.delenv("GLAVA_B200_NO_GEO", raising=False)
        with g.Renderer(p, batch=2) as r:
            r.raster_textures(tl, tr)
            frames.append([r.readback(s) for s in range(2)])
    for s in range(2):
        want = orc_pm.raster(op, tl[s], tr[s])
        assert np.array_equal(frames[0][s], want) and np.array_equal(frames[1][s], want)


@pytest.mark.parametrize("module,over", [("radial", {}), ("radial", dict(radial_amplify=-120.0)), ("circle", {}),
                                         ("circle", dict(circle_fill=1)), ("circle", dict(circle_amplify=-150.0)),
                                         ("circle", dict(circle_line=9.0))])
def test_polar_per_stream_reach_cull(orc_pm, module, over, built):
    """radial / circle skip the cells beyond what THIS stream's largest (and, for circle, smallest) value can
    light.  Streams whose values sit in narrow, different bands exercise every bound of that cull."""
    p = g.default_params(module, n=1024, w=800, h=600)
    for k, v in over.items():
        setattr(p, k, v)
    op = params_from(p)
    rng = np.random.default_rng(5)
    bands = [(0, 1), (0, 700), (3000, 3400), (20000, 21000), (64000, 65535), (100, 40000)]
    tl = np.stack([rng.integers(lo, hi + 1, 1024).astype(np.uint16) for lo, hi in bands])
    tr = np.stack([rng.integers(lo, hi + 1, 1024).astype(np.uint16) for lo, hi in reversed(bands)])
    with g.Renderer(p, batch=len(bands)) as r:
        r.raster_textures(tl, tr)
        for s in range(len(bands)):
            want = orc_pm.raster(op, tl[s], tr[s])
            assert np.array_equal(r.readback(s), want), (module, over, s)


def test_tap_table_equals_direct_k5(built, monkeypatch):
    """lazy K5 runs from a precomputed (index, weight) table; GLAVA_B200_NO_TAPTAB=1 evaluates the weights
    in the kernel.  Same textures at the sampled texels, hence same frames."""
    n, batch = 4096, 2
    rings = g.StreamRings(batch, n)
    for _ in range(20):
        rings.advance()
    out = []
    for flag in ("", "1"):
        if flag:
            monkeypatch.setenv("GLAVA_B200_NO_TAPTAB", "1")
        else:
            monkeypatch.delenv("GLAVA_B200_NO_TAPTAB", raising=False)
        p = g.default_params("bars", n=n, w=1920, h=120, lazy_smooth=1)
        with g.Renderer(p, batch=batch) as r:
            r.update(rings.lb, rings.rb, True)
            out.append([r.readback(s) for s in range(batch)])
    for s in range(batch):
        assert np.array_equal(out[0][s], out[1][s])


def test_reconfigure_is_a_live_uniform_update(orc_pm, built):
    """glava_b200_reconfigure: colour / amplify change on a live renderer == a fresh renderer with those parameters"""
    n = 1024
    for module, change in (("bars", dict(bars_amplify=120.0)), ("radial", dict(radial_amplify=200.0)), ("graph", dict(graph_vscale=150.0))):
        p = g.default_params(module, n=n, w=320, h=240, lazy_smooth=1); op = params_from(p)
        tl, tr = _textures(orc_pm, op, n, 2, 33)
        with g.Renderer(p, batch=2) as r:
            r.raster_textures(tl, tr)
            assert np.array_equal(r.readback(0), orc_pm.raster(op, tl[0], tr[0]))
            q = p.copy()
            for k, v in change.items():
                setattr(q, k, v)
            if module == "bars":
                q.bars_color.mode = 1
                for i, c in enumerate((0.9, 0.2, 0.1, 1.0)):
                    q.bars_color.lo[i] = c
            r.reconfigure(q)
            r.raster_textures(tl, tr)
            oq = params_from(q)
            for s in range(2):
                assert np.array_equal(r.readback(s), orc_pm.raster(oq, tl[s], tr[s])), (module, s)
            bad = q.copy(); bad.w = 640
            with pytest.raises(g.GlavaError, match="cannot change on a live renderer"):
                r.reconfigure(bad)
