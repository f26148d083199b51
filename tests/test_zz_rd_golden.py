"""tests/golden/rd_update_golden.npz: what the reference's OWN rd_update (on the null OpenGL driver) uploaded as audio
textures, frame by frame, for pipeline A with buffer scaling / keyframe interpolation (made by tests/golden/make_rd_golden.py).
Reference tree not needed: the oracle reproduces every frame exactly (CPU tier), the kernels' R16 textures follow within
2 LSB16 (-m gpu; north_star: spectrum within 1e-5 of peak)."""
import os

import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import OracleStream, OrcExt, params_from
from tests.conftest import GOLDEN


def _gold():
    return np.load(os.path.join(GOLDEN, "rd_update_golden.npz"))


def _cases():
    return [str(c) for c in _gold()["case_names"]]


def _unorm16(v):
    v = np.asarray(v, np.float32)
    q = np.rint((v * np.float32(65535.0)).astype(np.float32))       # GL's float -> R16: one rounding, ties to even (Mesa; llvmpipe goldens)
    with np.errstate(invalid="ignore"):
        return np.where(v > 0, np.where(v < 1, q.astype(np.int64), 65535), 0).astype(np.uint16)


def _params(z, case):
    req = [ln[len("#request "):] for ln in str(z[f"{case}_rc"]).splitlines() if ln.startswith("#request ") and not ln.startswith("#request mod")]
    p = g.load_config(requests=req + ["setsmoothpass false"], force_module="bars")       # no config dir: nothing overrides the requests
    ur, fr = (float(v) for v in z[f"{case}_rates"])
    p.ur, p.fr, p.w, p.h = ur, fr, 64, 16
    assert p.accel_fft == 0 and p.smooth_pass == 0 and p.n == int(z["n"])
    return p, ur, fr


def _frames(z, case, p, ur, fr):
    """(frame index, modified, pcm_l, pcm_r, reference upload l, r, comparable?)"""
    active = bool(p.interpolate) and ur / fr <= 0.9
    pushed = 0
    for k, m in enumerate(z[f"{case}_pattern"]):
        settled = pushed >= 2
        pushed += int(m)
        ok = not (active and not settled)                                  # the reference lerps uninitialised keyframes at first
        if p.bufscale > 1 and not m and not active:
            ok = False                                                     # ... and flashes raw scaled PCM here
        yield k, bool(m), z[f"{case}_pcm_l"][k], z[f"{case}_pcm_r"][k], z[f"{case}_up_l"][k], z[f"{case}_up_r"][k], ok


@pytest.mark.parametrize("case", _cases())
def test_oracle_reproduces_the_reference_uploads(orc, case, built):
    z = _gold()
    p, ur, fr = _params(z, case)
    st = OracleStream(orc, params_from(p), OrcExt(bufscale=p.bufscale, interpolate=p.interpolate, fr=fr, transform_smooth=0,
                                                  smooth_distance=0.01, smooth_ratio=4.0))
    checked = 0
    for k, m, pl, pr, ul, ur_, ok in _frames(z, case, p, ur, fr):
        sl, sr, tl, tr = st.update(pl, pr, m)
        if ok:
            assert np.array_equal(_unorm16(ul), tl) and np.array_equal(_unorm16(ur_), tr), (case, k)
            checked += 1
    assert checked >= 5


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases())
def test_kernels_follow_the_reference_uploads(case, built):
    z = _gold()
    p, ur, fr = _params(z, case)
    n = p.n
    checked = 0
    with g.Renderer(p, batch=2) as r:
        for k, m, pl, pr, ul, ur_, ok in _frames(z, case, p, ur, fr):
            r.update(np.stack([pl, pr]), np.stack([pr, pl]), m)            # stream 1 = the channels swapped
            tl, tr = r.textures()
            if ok:
                for got, want in ((tl[0], ul), (tr[0], ur_), (tl[1], ur_), (tr[1], ul)):
                    assert got.shape[0] == want.shape[0]
                    assert np.abs(got.astype(int) - _unorm16(want).astype(int)).max() <= 2, (case, k)
                checked += 1
    assert checked >= 5 and n == int(z["n"])
