"""Development aid: run the raster_textures-based -m gpu tests against a fake Renderer built on the host emulation (tests/emul),
to validate the test scripts themselves (golden keys, expectations, parameter plumbing) when no GPU is at hand.  The device
plumbing stays untested, of course.

    python tools/dryrun_gpu_tests.py
"""
import sys, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pytest
import glava_b200 as g
from tests import emul

class Fake:
    def __init__(self, params, batch=1, device=0):
        rc = g.lib().glava_b200_load_config  # noqa
        self.params = params.copy(); self.batch = batch; self.nsz = params.n; self._launch = 0; self._h = 1
        self.frames = [None] * batch
    def raster_textures(self, tl, tr=None):
        tl = np.asarray(tl); tr = tl if tr is None else np.asarray(tr)
        for s in range(self.batch):
            self.frames[s] = emul.raster(self.params, tl[s], tr[s])
        self._launch += 1
    def readback(self, s=0, out=None): return self.frames[s]
    # update() / textures(): the oracle's whole-rd_update restatement stands in for the spectrum kernels (script logic only)
    def update(self, lb, rb=None, modified=True):
        from oracle.oracle import Oracle, OracleStream, ext_from, params_from
        if not hasattr(self, "_streams"):
            o = Oracle("pm")
            self._streams = [OracleStream(o, params_from(self.params), ext_from(self.params)) for _ in range(self.batch)]
        rb = lb if rb is None else rb
        self._tex = [st.update(lb[s], rb[s], modified)[2:] for s, st in enumerate(self._streams)]
        self.nsz = self._streams[0].n
        self.frames = [emul.raster(self.params, t[0], t[1]) for t in self._tex]
    def textures(self):
        return np.stack([t[0] for t in self._tex]), np.stack([t[1] for t in self._tex])
    def reconfigure(self, p): self.params = p.copy(); self._launch += 1
    @property
    def launch_count(self): return self._launch
    def close(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): pass

g.Renderer = Fake
import glava_b200.api as api
api.Renderer = Fake
# Pipe.apply goes through the C ABI with a real handle: emulate with params re-evaluation
def fake_apply(self, renderer):
    if getattr(self, "_last", None) == self.binds():
        return
    self._last = self.binds(); renderer.reconfigure(self.params())
api.Pipe.apply = fake_apply
sys.exit(pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                      ROOT + "/tests/test_zz_gpu_blend.py", ROOT + "/tests/test_zz_gpu_color_expr.py", ROOT + "/tests/test_zz_gpu_pipe.py", ROOT + "/tests/test_glsl_golden.py", ROOT + "/tests/test_zz_rd_golden.py", ROOT + "/tests/test_llvmpipe_golden.py",
                      ]))
