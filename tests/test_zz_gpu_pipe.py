"""-m gpu: a `--pipe` line changes the colours of a RUNNING renderer (glava_b200_pipe_apply = the uniform write of
render.c:2071-2100); the frame equals the oracle's for the re-evaluated configuration, bit for bit."""
import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import params_from

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("module", ["bars", "radial", "graph", "wave", "circle"])
def test_pipe_apply_recolours_a_running_renderer(orc_pm, built, module):
    n, (w, h) = 1024, ((640, 480) if module in ("radial", "circle") else (256, 192))     # C_RADIUS 128 needs the room
    req = [f"setbufsize {n}", f"setgeometry 0 0 {w} {h}"]
    rng = np.random.default_rng(8)
    with g.Pipe(["fg", "bg"], requests=req, force_module=module) as pipe:
        p0 = pipe.params()
        op0 = params_from(p0)
        tl = orc_pm.smooth_pass(op0, (rng.random(n) ** 2 * 65535).astype(np.uint16))
        tr = orc_pm.smooth_pass(op0, (rng.random(n) ** 2 * 65535).astype(np.uint16))
        if module == "wave":
            tl = np.clip(tl.astype(int) // 8 + 28672, 0, 65535).astype(np.uint16)
        with g.Renderer(p0, batch=1) as r:
            r.raster_textures(tl[None], tr[None])
            before = r.readback(0)
            assert np.array_equal(before, orc_pm.raster(op0, tl, tr))             # unwritten binds: vec4(0) colours
            assert pipe.feed("fg = #ff8000\nbg = 0.1,0.2,0.3,1\n") == 2
            pipe.apply(r)
            p1 = pipe.params()
            r.raster_textures(tl[None], tr[None])
            after = r.readback(0)
            assert np.array_equal(after, orc_pm.raster(params_from(p1), tl, tr))
            assert not np.array_equal(after, before)
            launches = r.launch_count
            pipe.apply(r)                                                         # nothing changed: no reconfigure, no launches
            assert r.launch_count == launches
