"""The reference itself on a real OpenGL — Mesa llvmpipe, GLava's stated software floor — against the oracle and the kernels.

tests/golden/llvmpipe_golden.npz (tests/golden/make_llvmpipe_golden.py) holds what the reference's own rd_update produced on
synthetic PCM for 58 configurations: the R16 texels of every upload, the textures its 1-D passes rendered (gravity store,
average, smooth) and the final frame.  Nothing in it was computed by code of this repository.

* CPU tier: the C oracle (libm build) replays every case pass by pass and end to end.  K1-K4 are integer / single-rounding
  arithmetic: bit-exact.  K5 and the module stages call sin / log / atan, whose precision GLSL leaves to the implementation:
  <= 1 LSB, plus a COUNTED handful of hard-edge flips (a tap entering a window, a pixel crossing a threshold) per case.
* live (where the llvmpipe harness loads): every case is re-run and must reproduce the committed golden bit for bit.
* -m gpu: the kernels, from the golden's textures (raster half alone) and from the PCM (whole path).
"""
import json
import os

import numpy as np
import pytest

import glava_b200 as g
from glava_b200.api import Params
from oracle.oracle import OracleStream, ext_from, params_from
from tests.conftest import GOLDEN
from tests.pcm import checksum, pcm_frames

PATH = os.path.join(GOLDEN, "llvmpipe_golden.npz")


def _names():
    z = np.load(PATH)
    return [str(c) for c in z["case_names"]]


@pytest.fixture(scope="module")
def gold():
    return np.load(PATH)


def case_of(gold, name):
    cfg = json.loads(str(gold[f"{name}_cfg"]))
    p = Params.from_dict(json.loads(str(gold[f"{name}_params"])))
    lb, rb = pcm_frames(cfg["seed"], p.n, cfg["frames"])
    assert checksum(lb) ^ checksum(rb) == cfg["pcm_checksum"], "tests/pcm.py no longer generates the PCM the golden was made from"
    return cfg, p, lb, rb


def frame_diff(a, b):
    d = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    return int((d > 0).sum()), int((d > 1).sum())


# What GLSL leaves to the implementation shows up as (a) 1-LSB differences and (b) a COUNTED handful of hard-edge flips
# (a bar top, a line edge: a threshold crossed behind an ulp in sin / log / atan, or behind a 1-LSB16 texel).  Everything
# else — uploads, the gravity / average passes, every non-native-opacity blend — is bit-exact.
FLIP_PIXELS = 8          # pixels off by more than 1 LSB, per case (the largest case has 8.3 M pixels)
FLIP_TEXELS = 4          # K5 texels off by more than 1 LSB16, per case and channel
TEXEL_LSB = {
    # ROUND_FORMULA circular is sqrt(1 - (x - 1)^2) (util/common.glsl:21): this llvmpipe evaluates sqrt through an approximate
    # reciprocal square root, libm's is correctly rounded — the weighted sums differ by up to 4 LSB16 (the frame is identical)
    "bars_k5_circular": 4,
}


def _close_textures(name, mine, want):
    d = np.abs(mine.astype(int) - want.astype(int))
    lsb = TEXEL_LSB.get(name, 1)
    assert (d > lsb).sum() <= FLIP_TEXELS, (name, "texels beyond %d LSB16" % lsb, int((d > lsb).sum()), int(d.max()))


def _uploads(orc, p, lb, rb, f):
    """the R16 texels GL makes of what rd_update uploads for frame f of this case: transform_fft's output under setaccelfft
    (the oracle's orc_fft_f32 is bit-identical to the reference's, tests/test_oracle_golden.py), wrange for wave"""
    if p.module == g.api.MODULES.index("wave"):
        vals = [((lb[f] + np.float32(1)) / np.float32(2)).astype(np.float32)]
    else:
        op = params_from(p)
        a = lb[f] if p.channels == 2 else lb[f]
        vals = [orc.fft_f32(op, lb[f]), orc.fft_f32(op, rb[f])]
    return [np.rint((np.clip(v, 0, 1).astype(np.float32) * np.float32(65535)).astype(np.float32)).astype(np.uint16) for v in vals]


@pytest.mark.parametrize("name", _names())
def test_oracle_replays_the_reference_on_llvmpipe(orc, gold, name, built):
    cfg, p, lb, rb = case_of(gold, name)
    op = params_from(p)
    st = OracleStream(orc, op, ext_from(p))
    for f in range(cfg["frames"]):
        sl, sr, tl, tr = st.update(lb[f], rb[f], True)
    tex = gold[f"{name}_tex"]; want = gold[f"{name}_frame"]
    is_test = p.module == g.api.MODULES.index("test"); is_wave = p.module == g.api.MODULES.index("wave")
    # 0. uploads: the float -> R16 conversion of GL, every frame, bit for bit (setaccelfft: transform_fft's output)
    if p.accel_fft and p.channels == 2 and not is_test:
        for f in range(cfg["frames"]):
            mine = _uploads(orc, p, lb, rb, f)
            for ch in range(len(mine)):
                assert np.array_equal(mine[ch], gold[f"{name}_upl"][f, ch]), (name, "upload", f, ch)
    # 1. K5 alone: the oracle's smooth pass on the reference's own average texture
    if p.smooth_pass and p.accel_fft and p.avg_frames > 1 and not is_test and not is_wave:
        for ch in range(2):
            _close_textures(name, orc.smooth_pass(op, gold[f"{name}_av"][ch]), gold[f"{name}_sm"][ch])
    # 2. the raster half on the reference's own textures
    off, bad = frame_diff(orc.raster(op, tex[0], tex[1]), want)
    assert bad <= FLIP_PIXELS, (name, "raster on llvmpipe's textures", off, bad)
    if not p.premultiply_alpha:
        assert off == 0, (name, "blended frames are reproduced exactly", off)
    # 3. the whole path from PCM: textures, then pixels
    if not is_test:
        _close_textures(name, tl, tex[0])
        if not is_wave:
            _close_textures(name, tr, tex[1])
    off, bad = frame_diff(orc.raster(op, tl, tr), want)
    assert bad <= FLIP_PIXELS, (name, "end to end", off, bad)


def test_the_reference_known_answer_on_llvmpipe(gold):
    """`test` module: every pixel #55000055 (glava.c test mode, render.c:2420-2456) — rendered by the real thing"""
    f = gold["test_small_frame"]
    assert (f == np.array([85, 0, 0, 85], np.uint8)).all()
    assert "llvmpipe" in str(gold["gl_strings"][1]) and str(gold["gl_strings"][0]).startswith("3.3")


def _live():
    from oracle import ref_gl
    return ref_gl if (ref_gl.available() and os.path.isdir(ref_gl.REF_SHADERS)) else None


@pytest.mark.skipif(_live() is None, reason="needs the reference tree and the llvmpipe harness (oracle/_ref/libglava_ref_gl.so + Nsight's Mesa)")
def test_live_llvmpipe_reproduces_the_committed_golden(gold):
    """re-run every case on the real thing: the committed file is what the reference renders today"""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_llvmpipe_golden as mk
    for name, rc_text, files, frames, seed in mk.cases():
        rec = mk.run_case(name, rc_text, files, frames, seed)
        for key, val in rec.items():
            if key.endswith("_params") or key.endswith("_cfg"):
                assert json.loads(str(val)) == json.loads(str(gold[key])), key
            else:
                assert np.array_equal(val, gold[key]), key


# ---- -m gpu: the kernels ------------------------------------------------------------------------------------------------
def _gpu_cases():
    return _names()


@pytest.mark.gpu
@pytest.mark.parametrize("name", _gpu_cases())
def test_kernels_render_what_llvmpipe_renders(gold, name, built):
    """(a) raster kernels on the reference's own textures; (b) the whole path from the PCM through glava_b200_update"""
    cfg, p, lb, rb = case_of(gold, name)
    tex = gold[f"{name}_tex"]; want = gold[f"{name}_frame"]
    is_test = p.module == g.api.MODULES.index("test"); is_wave = p.module == g.api.MODULES.index("wave")
    with g.Renderer(p, batch=2) as r:
        r.raster_textures(np.stack([tex[0], tex[0]]), np.stack([tex[1], tex[1]]))
        off, bad = frame_diff(r.readback(1), want)
        assert bad <= FLIP_PIXELS, (name, "raster kernels on llvmpipe's textures", off, bad)
    with g.Renderer(p, batch=2) as r:
        for f in range(cfg["frames"]):
            r.update(np.stack([lb[f], lb[f]]), np.stack([rb[f], rb[f]]), True)
        tl, tr = r.textures()
        if not is_test:
            d = np.abs(tl[1].astype(int) - tex[0].astype(int))
            lsb = max(TEXEL_LSB.get(name, 1), 2)                # + the FFT's own float32 rounding (<= 1e-5 of peak)
            assert (d > lsb).sum() <= 16, (name, "texture", int((d > lsb).sum()), int(d.max()))
        off, bad = frame_diff(r.readback(1), want)
        assert bad <= max(FLIP_PIXELS, 3e-4 * want.shape[0] * want.shape[1]), (name, "end to end", off, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", [0, 1])
def test_lazy_and_full_smoothing_both_match_llvmpipe_at_the_headline_geometry(gold, lazy, built):
    name = "bars_1080p_n4096"
    cfg, p, lb, rb = case_of(gold, name)
    p.lazy_smooth = lazy
    with g.Renderer(p, batch=3) as r:
        for f in range(cfg["frames"]):
            r.update(np.stack([lb[f]] * 3), np.stack([rb[f]] * 3), True)
        off, bad = frame_diff(r.readback(2), gold[f"{name}_frame"])
        assert bad <= 3e-4 * 1920 * 1080, (off, bad)
