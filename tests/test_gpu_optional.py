"""-m gpu: the optional rd_update stages through the C ABI — bufscale (render.c:1765-1790), transform_smooth
(render.c:694-718), keyframe interpolation (render.c:1792-1809, 2347-2353) — against the oracle."""
import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import OracleStream, ext_from, params_from

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [256, 1024, 4096, 16384])
def test_transform_smooth_kernel_bit_exact(orc, n, built):
    rng = np.random.default_rng(n)
    for d, ratio in ((0.01, 4.0), (0.2, 2.0), (0.3, 1.0)):
        p = g.default_params("bars", n=n, w=64, h=64, transform_smooth=1, smooth_distance=d, smooth_ratio=ratio)
        planes = (rng.random((5, n)) ** 3).astype(np.float32)
        planes[rng.random((5, n)) < 0.25] = 0
        planes[3] = 0                                                   # silence: NaN over the whole head
        planes[4, 7] = np.nan
        with g.Renderer(p, batch=1) as r:
            got = r.transform_smooth(planes)
        for k in range(5):
            assert np.array_equal(got[k], orc.transform_smooth(planes[k], d, ratio), equal_nan=True), (n, d, ratio, k)


def _run(orc_pm, p, batch, steps, pattern=(True,)):
    """drive product and oracle with the same rings; returns per-frame (product textures, oracle textures, spectra)"""
    op = params_from(p); x = ext_from(p)
    rings = g.StreamRings(batch, p.n)
    streams = [OracleStream(orc_pm, op, x) for _ in range(batch)]
    out = []
    with g.Renderer(p, batch=batch) as r:
        assert r.nsz == streams[0].n
        for i in range(steps):
            mod = pattern[i % len(pattern)]
            if mod:
                rings.advance()
            r.update(rings.lb, rings.rb, mod)
            want = [st.update(rings.lb[s], rings.rb[s], mod) for s, st in enumerate(streams)]
            gl, gr = r.textures()
            sl, sr = r.spectrum()
            frames = [r.readback(s) for s in range(batch)]
            out.append((gl, gr, sl, sr, want, frames))
    return out, op


def _close_u16(a, b, lsb=2):
    return np.abs(a.astype(int) - b.astype(int)).max() <= lsb


@pytest.mark.parametrize("module,accel,k", [("bars", 1, 2), ("bars", 0, 4), ("wave", 1, 2), ("radial", 1, 2)])
def test_bufscale_end_to_end(orc_pm, module, accel, k, built):
    n = 2048
    p = g.default_params(module, n=n, w=640, h=360, accel_fft=accel, bufscale=k)
    out, op = _run(orc_pm, p, batch=3, steps=n // 256 + 4)
    gl, gr, sl, sr, want, frames = out[-1]
    assert gl.shape == (3, n // k)
    opk = params_from(p); opk.n = n // k
    for s in range(3):
        assert _close_u16(gl[s], want[s][2]), (module, s)
        if module != "wave":
            assert _close_u16(gr[s], want[s][3])
        peak = np.abs(want[s][0]).max()
        assert np.abs(sl[s] - want[s][0]).max() <= 2.5e-5 * max(peak, 1.0)
        # pixels: exact on the product's own textures
        assert np.array_equal(frames[s], orc_pm.raster(opk, gl[s], gr[s] if module != "wave" else gl[s]))


@pytest.mark.parametrize("module,accel", [("bars", 0), ("graph", 0), ("wave", 1), ("wave", 0)])
def test_keyframe_interpolation_end_to_end(orc_pm, module, accel, built):
    """4 frames per audio update: 1 modified + 3 interpolated; every frame's texture must follow the oracle's"""
    n = 1024
    p = g.default_params(module, n=n, w=320, h=200, accel_fft=accel, interpolate=1)
    p.fr = p.ur * 4
    out, op = _run(orc_pm, p, batch=2, steps=28, pattern=(True, False, False, False))
    moved = 0
    for i, (gl, gr, sl, sr, want, frames) in enumerate(out):
        for s in range(2):
            assert _close_u16(gl[s], want[s][2]), (module, i, s)
            assert np.array_equal(frames[s], orc_pm.raster(op, gl[s], gr[s] if module != "wave" else gl[s]))
        if i >= 9 and i % 4 in (1, 2, 3):
            moved += int(not np.array_equal(gl[0], out[i - 1][0][0]))
    assert moved >= 10                                                  # interpolated frames really change the texture


def test_interpolation_inactive_cases_match_plain_renderer(built):
    """uratio > 0.9 (render.c:1761-1763) and an fft module under setaccelfft (render.c:2161-2168): setinterpolate
    changes nothing"""
    n, batch = 1024, 2
    rings = g.StreamRings(batch, n)
    for kw in (dict(accel_fft=0, fr=0.0), dict(accel_fft=1, fr=400.0)):
        pa = g.default_params("bars", n=n, w=320, h=200, interpolate=1, **kw)
        pb = g.default_params("bars", n=n, w=320, h=200, interpolate=0, **kw)
        with g.Renderer(pa, batch=batch) as a, g.Renderer(pb, batch=batch) as b:
            for i in range(6):
                rings.advance()
                a.update(rings.lb, rings.rb, True); b.update(rings.lb, rings.rb, True)
                a.update(rings.lb, rings.rb, False); b.update(rings.lb, rings.rb, False)
            assert np.array_equal(a.textures()[0], b.textures()[0])
            assert np.array_equal(a.readback(1), b.readback(1))


@pytest.mark.parametrize("module,accel", [("bars", 1), ("bars", 0), ("wave", 1)])
def test_transform_smooth_end_to_end(orc_pm, module, accel, built):
    n = 2048
    p = g.default_params(module, n=n, w=640, h=360, accel_fft=accel, transform_smooth=1)
    out, op = _run(orc_pm, p, batch=3, steps=n // 256 + 4)
    gl, gr, sl, sr, want, frames = out[-1]
    for s in range(3):
        ws = want[s][0]
        assert np.array_equal(np.isnan(sl[s]), np.isnan(ws)) and np.isnan(sl[s][0])
        ok = ~np.isnan(ws)
        assert np.abs(sl[s][ok] - ws[ok]).max() <= 2.5e-5 * max(np.abs(ws[ok]).max(), 1.0)
        assert _close_u16(gl[s], want[s][2])
        assert np.array_equal(frames[s], orc_pm.raster(op, gl[s], gr[s] if module != "wave" else gl[s]))
    # the head of the spectrum really went through the transform: compare with a renderer without it
    q = g.default_params(module, n=n, w=640, h=360, accel_fft=0)
    out2, _ = _run(orc_pm, q, batch=3, steps=n // 256 + 4)
    assert not np.allclose(out2[-1][2][0][1: n // 4], sl[0][1: n // 4], atol=1e-7)
    assert np.allclose(out2[-1][2][0][n // 4 + 64:], sl[0][n // 4 + 64:], atol=1e-6)        # the tail is untouched


def test_all_three_together(orc_pm, built):
    n = 4096
    p = g.default_params("bars", n=n, w=640, h=360, accel_fft=1, bufscale=2, transform_smooth=1, interpolate=1)
    p.fr = 300.0
    out, op = _run(orc_pm, p, batch=2, steps=30, pattern=(True, False, False))
    for i, (gl, gr, sl, sr, want, frames) in enumerate(out[-6:]):
        for s in range(2):
            assert gl.shape[1] == 2048 and _close_u16(gl[s], want[s][2]) and _close_u16(gr[s], want[s][3]), (i, s)


def test_reconfigure_keeps_optional_stage_layout(built):
    p = g.default_params("bars", n=1024, w=320, h=200, transform_smooth=1)
    with g.Renderer(p, batch=1) as r:
        q = g.default_params("bars", n=1024, w=320, h=200, transform_smooth=1, smooth_distance=0.1)
        r.reconfigure(q)                                                    # table rebuilt
        x = np.random.default_rng(0).random((1, 1024)).astype(np.float32)
        from oracle.oracle import Oracle
        assert np.array_equal(r.transform_smooth(x)[0], Oracle("libm").transform_smooth(x[0], 0.1, 4.0), equal_nan=True)
        with pytest.raises(g.GlavaError):
            r.reconfigure(g.default_params("bars", n=1024, w=320, h=200, transform_smooth=0))
        with pytest.raises(g.GlavaError):
            r.reconfigure(g.default_params("bars", n=1024, w=320, h=200, transform_smooth=1, bufscale=2))


# ---- offscreen hand-off: glava_sizereq / glava_wait / glava_tex analogues (glava.h:22-24) -------------------------
@pytest.mark.parametrize("module", ["bars", "radial", "circle", "graph", "wave"])
def test_sizereq_resizes_at_the_next_update_and_keeps_the_spectrum_state(module, built):
    n, batch = 1024, 3
    small = g.default_params(module, n=n, w=320, h=200, lazy_smooth=1)
    big = g.default_params(module, n=n, w=804, h=601, lazy_smooth=1)
    rings = g.StreamRings(batch, n)
    with g.Renderer(small, batch=batch) as a, g.Renderer(big, batch=batch) as b:
        for i in range(7):
            rings.advance()
            if i == 5:
                a.sizereq(804, 601)
                assert a.params.w == 320                                  # not yet: applied by the next update
            a.update(rings.lb, rings.rb, True); b.update(rings.lb, rings.rb, True)
        assert (a.params.w, a.params.h) == (804, 601)
        for s in range(batch):
            assert np.array_equal(a.readback(s), b.readback(s)), (module, s)   # same gravity / average history
        a.sizereq(100000, 10)                                             # invalid geometry: the update reports it
        with pytest.raises(g.GlavaError):
            a.update(rings.lb, rings.rb, True)


def test_frame_event_wait_and_device_pointers(built):
    """a consumer stream ordered after the frame by the event alone (no host sync) copies a frame straight out of the
    framebuffer array — what a compositor / encoder does with glava_tex()'s texture in the reference"""
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    p = g.default_params("bars", n=1024, w=320, h=200, fb_slots=2)
    rings = g.StreamRings(4, 1024)
    frame = 320 * 200 * 4
    with g.Renderer(p, batch=4) as r:
        for _ in range(6):
            rings.advance(); r.update(rings.lb, rings.rb, True)
        ev = r.frame_event
        assert ev
        consumer = C.c_void_p()
        assert rt.cudaStreamCreateWithFlags(C.byref(consumer), 1) == 0            # cudaStreamNonBlocking
        assert rt.cudaStreamWaitEvent(consumer, C.c_void_p(ev), 0) == 0
        out = g.pinned_empty((200, 320, 4), np.uint8)
        assert rt.cudaMemcpyAsync(C.c_void_p(out.ctypes.data), C.c_void_p(r.frame_device(3)), C.c_size_t(frame), 2, consumer) == 0
        assert rt.cudaStreamSynchronize(consumer) == 0
        assert np.array_equal(out, r.readback(3))
        rt.cudaStreamDestroy(consumer)
        r.wait_frame()
        base = r.framebuffer_device
        assert r.frame_device(0) == base and r.frame_device(1) == base + frame and r.frame_device(3) == base + frame   # ring of 2
        assert r.frame_device(4) is None
        h = r.framebuffer_ipc()
        assert len(h) == 64 and any(h)


@pytest.mark.gpu
@pytest.mark.parametrize("accel", [1])       # (pipeline A feeds the NaN pattern into the stateful gravity pass: not comparable, see below)
def test_smooth_before_fft_equals_the_oracle_stream(orc_pm, built, accel):
    """transform_smooth = 2: "smooth" listed BEFORE "fft" in the module's bind (render.c:1218-1286) — transform_smooth on the
    PCM ring, then the module's chain as usual (pinned against the reference's rd_update in tests/test_ref_rd.py)"""
    from oracle.oracle import OracleStream, ext_from, params_from
    n, batch = 1024, 3
    p = g.default_params("bars", n=n, w=64, h=32, transform_smooth=2, smooth_distance=0.02, smooth_ratio=3.0, accel_fft=accel)
    rng = np.random.default_rng(31 + accel)
    sts = [OracleStream(orc_pm, params_from(p), ext_from(p)) for _ in range(batch)]
    with g.Renderer(p, batch=batch) as r:
        for _ in range(5):
            lb = (rng.standard_normal((batch, n)) * 0.2).astype(np.float32); rb = (rng.standard_normal((batch, n)) * 0.2).astype(np.float32)
            keep = lb.copy()
            r.update(lb, rb, True)
            want = [st.update(lb[s], rb[s], True) for s, st in enumerate(sts)]
            assert np.array_equal(lb, keep)                                    # the caller's rings are left alone
        sl, sr = r.spectrum(); tl, tr = r.textures()
        for s in range(batch):
            # transform_smooth leaves NaN in b[0] (0 / 0, render.c:694-718) and the FFT spreads it: WHICH outputs turn NaN
            # depends on the butterfly network (the reference's Danielson-Lanczos loop multiplies by every twiddle, trivial
            # ones included; the Stockham passes do not) — an artefact of a degenerate configuration, so only the bins that
            # are finite on both sides are compared
            ok = np.isfinite(want[s][0]) & np.isfinite(sl[s])
            if ok.any():
                assert np.abs(sl[s][ok] - want[s][0][ok]).max() <= 1e-5 * max(np.abs(want[s][0][ok]).max(), 1e-30)
