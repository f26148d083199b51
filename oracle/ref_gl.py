"""TEST INFRASTRUCTURE / bench reference arm — the reference's OWN renderer on a REAL OpenGL.

oracle/_ref/libglava_ref_gl.so is glava/render.c + glsl_ext.c + glad.c compiled where they lie (oracle/Makefile), with a
window backend "mesa" (oracle/ref_shim.c) that creates an OpenGL 3.3 core context on the Mesa 18.1.9 *llvmpipe* software
GL bundled with Nsight Compute in this image — GLava's stated software floor (reference README.md:121) — through a
display-less Xlib stand-in (oracle/fakex/fake_x11.c).  rd_new reads the configuration, rd_update runs transform_fft on
the CPU, the pass / gravity / average / smooth fragment shaders and the module's stages: everything between PCM and
pixels is the reference's code, compiled by Mesa's GLSL compiler and rasterised by llvmpipe.

Used (a) to pin the raster half of the oracle and the product kernels to a GL implementation the builder did not write
(tests/golden/make_llvmpipe_golden.py, tests/test_llvmpipe_golden.py) and (b) as bench.py's reference arm / cpu_baseline
("CPU-FFT + software-GL (llvmpipe)", BASELINE.json north_star).  Never imported by the product.

Known properties of this GL, recorded where they matter:
* it does not advertise GL_NV_texture_barrier; render.c:2217 calls glTextureBarrierNV unconditionally, i.e. the shipped
  default (`setaccelfft true`) would jump to NULL on this Mesa.  The harness points the entry at glFinish (llvmpipe
  flushes a scene before a later draw samples its render target, so the in-place gravity pass reads what K1 wrote);
* gl->ur starts at 1.0 until the first measured second (render.c:906, 2380-2390): callers set the rates explicitly
  (ref_rd_set_rates), like the null-driver harness does.
"""
import ctypes as C
import glob
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SHADERS = "/root/reference/shaders/glava"
_lib = None
_shader_dir = None


def mesa_path():
    """the llvmpipe libGL of the Nsight Compute bundle (same path here and on the GPU box: same image)"""
    hits = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/*/Mesa/libGL.so.1"))
    return hits[-1] if hits else None


def lib():
    """load fakex + Mesa + the reference renderer; None when any piece is missing or does not load"""
    global _lib
    if _lib is not None:
        return _lib or None
    so = os.path.join(HERE, "_ref", "libglava_ref_gl.so")
    mesa = mesa_path()
    if not (os.path.exists(so) and mesa and os.path.exists(os.path.join(HERE, "_ref", "fakex", "libX11.so.6"))):
        _lib = False
        return None
    try:
        L = C.CDLL(so)                                   # executable stack (GNU nested-function trampolines in rd_new)
    except OSError:
        _lib = False
        return None
    cp, vp, i32 = C.c_char_p, C.c_void_p, C.c_int
    L.ref_gl_error.restype = cp
    L.ref_gl_strings.restype = cp
    L.ref_gl_load.argtypes = [cp, cp]
    L.ref_gl_new.restype = vp
    L.ref_gl_new.argtypes = [C.POINTER(cp), cp, C.POINTER(cp)]
    L.ref_gl_frame.argtypes = [vp, vp, vp, C.c_size_t, i32, vp]
    L.ref_gl_texture.argtypes = [vp, i32, vp, i32]
    L.ref_gl_pass_texture.argtypes = [vp, i32, i32, vp, i32]
    L.ref_gl_size.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.ref_gl_destroy.argtypes = [vp]
    L.ref_gl_unpack_shaders.argtypes = [cp]
    L.ref_gl_finish.restype = None
    L.ref_rd_set_rates.argtypes = [vp, C.c_float, C.c_float]
    L.ref_rd_config.argtypes = [vp, vp, vp]
    if L.ref_gl_load(os.path.join(HERE, "_ref", "fakex").encode(), mesa.encode()) != 0:
        _lib = False
        return None
    _lib = L
    return L


def available():
    return lib() is not None


def shader_dir():
    """GLava's installed shader tree: the reference checkout when present, else the copy packed into the library"""
    global _shader_dir
    if _shader_dir is None:
        if os.path.isdir(REF_SHADERS):
            _shader_dir = REF_SHADERS
        else:
            d = tempfile.mkdtemp(prefix="glava_ref_shaders_")
            if lib().ref_gl_unpack_shaders(d.encode()) < 1:
                raise RuntimeError("could not unpack the packed shader tree")
            _shader_dir = d
    return _shader_dir


def user_dir(path, files):
    """a user configuration directory the way `glava --copy-config` lays it out: the user's own files, symlinks to the
    installed shaders / modules for everything else (render.c:1318-1320)"""
    os.makedirs(path)
    sd = shader_dir()
    for name, text in files.items():
        with open(os.path.join(path, name), "w") as f:
            f.write(text)
    for entry in os.listdir(sd):
        if entry not in files:
            os.symlink(os.path.join(sd, entry), os.path.join(path, entry))
    return path


class ReferenceGL:
    """One GLava renderer (rd_new .. rd_destroy) on llvmpipe.

    rc: text of the user's rc.glsl (None = the shipped one); files: other user files {name: text} (e.g. "bars.glsl",
    "smooth_parameters.glsl"); ur: the update rate handed to gl->ur / gl->fr (render.c:2386-2387 measure it).
    Only ONE `requests` entry is safe (a second --request corrupts the reference's heap: render.c:1415-1435 reuses a
    `struct glsl_ext` whose destructor list ext_free leaves dangling) — put everything else into rc."""

    def __init__(self, rc=None, files=None, requests=(), ur=22050.0 / 256.0, fr=None):
        L = lib()
        assert L is not None, "llvmpipe harness unavailable"
        assert len(requests) <= 1
        self.L = L
        self._tmp = None
        paths = [shader_dir()]
        if rc is not None or files:
            self._tmp = tempfile.mkdtemp(prefix="glava_ref_user_")
            fl = dict(files or {})
            if rc is not None:
                fl["rc.glsl"] = rc
            paths = [user_dir(os.path.join(self._tmp, "cfg"), fl), shader_dir()]
        pa = (C.c_char_p * (len(paths) + 1))(*[p.encode() for p in paths], None)
        rq = (C.c_char_p * (len(requests) + 1))(*[r.encode() for r in requests], None)
        self.h = L.ref_gl_new(pa, b"rc.glsl", rq)
        if not self.h:
            raise ValueError("the reference aborted in rd_new (shader compile errors are printed by the reference)")
        w, h = C.c_int(), C.c_int()
        L.ref_gl_size(self.h, C.byref(w), C.byref(h))
        self.w, self.hh = w.value, h.value
        ints = (C.c_int * 16)(); floats = (C.c_float * 12)()
        L.ref_rd_config(self.h, ints, floats)
        self.n = int(ints[0]) // max(int(ints[9]), 1)
        self.bufsize = int(ints[0])
        L.ref_rd_set_rates(self.h, ur, ur if fr is None else fr)
        self._lb = np.zeros(self.bufsize, np.float32); self._rb = np.zeros(self.bufsize, np.float32)

    @property
    def gl_strings(self):
        return tuple(self.L.ref_gl_strings(i).decode() for i in range(3))

    def frame(self, pcm_l=None, pcm_r=None, want_frame=True):
        """one iteration of glava.c:523-539: new PCM (modified) or none, rd_update; -> RGBA8 [h][w][4], row 0 = bottom"""
        modified = pcm_l is not None
        if modified:
            self._lb[:] = pcm_l; self._rb[:] = pcm_r
        lb, rb = self._lb.copy(), self._rb.copy()            # rd_update transforms its arguments in place
        img = np.empty((self.hh, self.w, 4), np.uint8) if want_frame else None
        rc = self.L.ref_gl_frame(self.h, lb.ctypes.data, rb.ctypes.data, lb.shape[0], 1 if modified else 0,
                                 img.ctypes.data if want_frame else None)
        if rc != 0:
            raise ValueError("the reference aborted in rd_update")
        return img

    def finish(self):
        self.L.ref_gl_finish()

    def texture(self, which):
        """the R16 texels stage 1 samples for audio_l (0) / audio_r (1)"""
        t = np.zeros(self.bufsize, np.uint16)
        w = self.L.ref_gl_texture(self.h, which, t.ctypes.data, t.shape[0])
        return t[:w].copy() if w > 0 else None

    def pass_texture(self, which, what):
        """what: 0 upload (transform chain output), 1 gr_store (K1 / K2), 2 av (K4), 3 sm (K5), 4 + i: ring slot i (K3)"""
        t = np.zeros(self.bufsize, np.uint16)
        w = self.L.ref_gl_pass_texture(self.h, which, what, t.ctypes.data, t.shape[0])
        return t[:w].copy() if w > 0 else None

    def close(self):
        if self.h:
            self.L.ref_gl_destroy(self.h); self.h = None
        if self._tmp:
            shutil.rmtree(self._tmp, ignore_errors=True); self._tmp = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
