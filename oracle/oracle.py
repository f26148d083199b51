"""TEST INFRASTRUCTURE — ctypes loader for the oracle libraries.

  libglava_oracle.so     our C restatement, libm transcendentals (the independent checker)
  libglava_oracle_pm.so  same restatement, transcendentals from glava_b200/csrc/gl_math.h
                         (bit-exact comparisons against the CUDA kernels)
  _ref/libglava_ref.so   the reference's own render.c transforms, compiled where they lie
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODULES = ("bars", "radial", "circle", "graph", "wave", "test")


class OrcColor(C.Structure):
    _fields_ = [("mode", C.c_int), ("lo", C.c_float * 4), ("hi", C.c_float * 4), ("gradient", C.c_float)]


class OrcParams(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("fft_scale", C.c_float), ("fft_cutoff", C.c_float), ("gravity_step", C.c_float),
        ("ur", C.c_float), ("avg_frames", C.c_int), ("avg_window", C.c_int), ("accel_fft", C.c_int),
        ("smooth_pass", C.c_int), ("smooth_factor", C.c_float), ("sample_range", C.c_float),
        ("sample_scale", C.c_float), ("hybrid_weight", C.c_float), ("sample_mode", C.c_int),
        ("round_formula", C.c_int),
        ("module", C.c_int), ("w", C.c_int), ("h", C.c_int), ("channels", C.c_int), ("premultiply_alpha", C.c_int),
        ("bars_width", C.c_float), ("bars_gap", C.c_float), ("bars_outline_width", C.c_float), ("bars_amplify", C.c_float),
        ("bars_color", OrcColor), ("bars_outline_mode", C.c_int), ("bars_outline", C.c_float * 4),
        ("bars_direction", C.c_int), ("bars_invert", C.c_int), ("bars_flip", C.c_int), ("bars_mirror_yx", C.c_int),
        ("radial_radius", C.c_float), ("radial_line", C.c_float), ("radial_line_half", C.c_float),
        ("radial_outline", C.c_float * 4), ("radial_nbars", C.c_int), ("radial_bar_width", C.c_float),
        ("radial_amplify", C.c_float), ("radial_color", OrcColor), ("radial_rotate", C.c_float), ("radial_invert", C.c_int),
        ("radial_bar_alias", C.c_float), ("radial_c_alias", C.c_float), ("radial_off_x", C.c_float), ("radial_off_y", C.c_float),
        ("circle_radius", C.c_float), ("circle_line", C.c_float), ("circle_outline", C.c_float * 4),
        ("circle_amplify", C.c_float), ("circle_rotate", C.c_float), ("circle_invert", C.c_int),
        ("circle_fill", C.c_int), ("circle_smooth", C.c_int),
        ("graph_vscale", C.c_float), ("graph_direction", C.c_int), ("graph_color", OrcColor),
        ("graph_draw_outline", C.c_int), ("graph_draw_highlight", C.c_int), ("graph_outline", C.c_float * 4),
        ("graph_invert", C.c_int),
        ("wave_min_thickness", C.c_float), ("wave_max_thickness", C.c_float), ("wave_base_color", C.c_float * 4),
        ("wave_amplify", C.c_float), ("wave_outline", C.c_float * 4),
        ("graph_join_channels", C.c_int), ("graph_anti_alias", C.c_int), ("shader_pre_smoothed", C.c_int), ("radial_bar_width_int", C.c_int), ("radial_bar_outline_width", C.c_float), ("radial_bar_outline", C.c_float * 4),
        ("clear_color", C.c_float * 4),
    ]


def _copy_fields(dst, src):
    """field-by-name copy between two ctypes structures (nested structs / arrays by bytes)."""
    names = {f[0] for f in src._fields_}
    for name, typ in dst._fields_:
        if name not in names:
            continue
        v = getattr(src, name)
        if isinstance(v, (C.Structure, C.Array)):
            d = getattr(dst, name)
            assert C.sizeof(d) == C.sizeof(v), name
            C.memmove(C.byref(d), C.byref(v), C.sizeof(v))
        else:
            setattr(dst, name, v)
    return dst


def params_from(product_params):
    """OrcParams with the same settings as a glava_b200.Params."""
    return _copy_fields(OrcParams(), product_params)


class OrcExt(C.Structure):
    """optional rd_update stages (orc_ext in glava_oracle.h)"""
    _fields_ = [("bufscale", C.c_int), ("interpolate", C.c_int), ("fr", C.c_float), ("transform_smooth", C.c_int),
                ("smooth_distance", C.c_float), ("smooth_ratio", C.c_float)]


def ext_from(product_params=None, **over):
    x = OrcExt(1, 0, 0.0, 0, 0.01, 4.0)                      # render.c:908,912(rc.glsl:131),917,918
    if product_params is not None:
        for name, _ in OrcExt._fields_:
            if hasattr(product_params, name):
                setattr(x, name, getattr(product_params, name))
    for k, v in over.items():
        setattr(x, k, v)
    return x


def build(force=False):
    """make oracle (+ ref when /root/reference exists).  Building the checker is not using it."""
    need = force or not all(os.path.exists(os.path.join(HERE, f)) for f in ("libglava_oracle.so", "libglava_oracle_pm.so"))
    if need:
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if os.path.exists("/root/reference/glava/render.c") and (force or not os.path.exists(os.path.join(HERE, "_ref", "libglava_ref.so"))):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class Oracle:
    def __init__(self, kind="libm"):
        build()
        name = {"libm": "libglava_oracle.so", "pm": "libglava_oracle_pm.so"}[kind]
        L = C.CDLL(os.path.join(HERE, name))
        vp, i32 = C.c_void_p, C.c_int
        PP = C.POINTER(OrcParams)
        L.orc_default_params.argtypes = [PP, i32, i32, i32, i32]
        L.orc_window.argtypes = [vp, i32]
        L.orc_fft_f32.argtypes = [PP, vp]
        L.orc_fft_f64.argtypes = [PP, vp, vp]
        L.orc_chan_new.restype = vp
        L.orc_chan_new.argtypes = [PP]
        L.orc_chan_free.argtypes = [vp]
        L.orc_chan_update.argtypes = [vp, PP, vp, i32, vp, vp]
        L.orc_smooth_pass.argtypes = [PP, vp, vp]
        L.orc_raster.argtypes = [PP, vp, vp, vp]
        L.orc_raster_rows.argtypes = [PP, vp, vp, vp, i32, i32]
        L.orc_fifo_ingest.argtypes = [vp, vp, i32, vp, i32, i32]
        L.orc_math_kind.restype = C.c_char_p
        L.orc_bufscale.argtypes = [vp, i32, i32, vp]
        L.orc_transform_smooth.argtypes = [vp, i32, C.c_float, C.c_float]
        L.orc_interp.argtypes = [vp, vp, i32, C.c_float, C.c_float, i32, vp]
        L.orc_stream_new.restype = vp
        L.orc_stream_new.argtypes = [PP, C.POINTER(OrcExt)]
        L.orc_stream_free.argtypes = [vp]
        L.orc_stream_n.argtypes = [vp]
        L.orc_stream_update.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
        self.L = L
        self.kind = kind

    def default_params(self, module="bars", n=4096, w=800, h=600, **over):
        p = OrcParams()
        self.L.orc_default_params(C.byref(p), MODULES.index(module), n, w, h)
        for k, v in over.items():
            if isinstance(v, (list, tuple)):
                v = type(getattr(p, k))(*v)
            setattr(p, k, v)
        return p

    def window(self, n):
        w = np.empty(n, np.float64)
        self.L.orc_window(w.ctypes.data, n)
        return w

    def fft_f32(self, p, pcm):
        b = np.array(pcm, dtype=np.float32, copy=True)
        self.L.orc_fft_f32(C.byref(p), b.ctypes.data)
        return b

    def fft_f64(self, p, pcm):
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        out = np.empty(p.n, np.float64)
        self.L.orc_fft_f64(C.byref(p), x.ctypes.data, out.ctypes.data)
        return out

    def smooth_pass(self, p, tex):
        tex = np.ascontiguousarray(tex, dtype=np.uint16)
        out = np.empty_like(tex)
        self.L.orc_smooth_pass(C.byref(p), tex.ctypes.data, out.ctypes.data)
        return out

    def raster(self, p, tex_l, tex_r, rows=None):
        tl = np.ascontiguousarray(tex_l, dtype=np.uint16)
        tr = np.ascontiguousarray(tex_r if tex_r is not None else tex_l, dtype=np.uint16)
        out = np.zeros((p.h, p.w, 4), dtype=np.uint8)
        y0, y1 = rows if rows else (0, p.h)
        self.L.orc_raster_rows(C.byref(p), tl.ctypes.data, tr.ctypes.data, out.ctypes.data, y0, y1)
        return out

    def bufscale(self, pcm, k):
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        out = np.empty(x.shape[0] // k, np.float32)
        self.L.orc_bufscale(x.ctypes.data, x.shape[0], k, out.ctypes.data)
        return out

    def transform_smooth(self, buf, smooth_distance=0.01, smooth_ratio=4.0):
        b = np.array(buf, dtype=np.float32, copy=True)
        self.L.orc_transform_smooth(b.ctypes.data, b.shape[0], smooth_distance, smooth_ratio)
        return b

    def interp(self, start, end, ur, fr, kcounter):
        s = np.ascontiguousarray(start, dtype=np.float32); e = np.ascontiguousarray(end, dtype=np.float32)
        out = np.empty_like(s)
        self.L.orc_interp(s.ctypes.data, e.ctypes.data, s.shape[0], ur, fr, kcounter, out.ctypes.data)
        return out

    def fifo_ingest(self, ring_l, ring_r, chunk, channels=2):
        chunk = np.ascontiguousarray(chunk, dtype=np.int16)
        self.L.orc_fifo_ingest(ring_l.ctypes.data, ring_r.ctypes.data, ring_l.shape[0], chunk.ctypes.data,
                               chunk.shape[0] // 2, channels)


class OracleChannel:
    """persistent per-(stream, channel) state + one-update entry point"""

    def __init__(self, oracle, p):
        self.o, self.p = oracle, p
        self.h = oracle.L.orc_chan_new(C.byref(p))

    def update(self, pcm, is_fft=True):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        spec = np.empty(self.p.n, np.float32); tex = np.empty(self.p.n, np.uint16)
        self.o.L.orc_chan_update(self.h, C.byref(self.p), pcm.ctypes.data, 1 if is_fft else 0, spec.ctypes.data, tex.ctypes.data)
        return spec, tex

    def __del__(self):
        try: self.o.L.orc_chan_free(self.h)
        except Exception: pass


class OracleStream:
    """one stream through whole rd_update calls, optional stages included (orc_stream_*)"""

    def __init__(self, oracle, p, ext):
        self.o, self.p, self.x = oracle, p, ext
        self.h = oracle.L.orc_stream_new(C.byref(p), C.byref(ext))
        self.n = oracle.L.orc_stream_n(self.h)

    def update(self, lb, rb, modified=True):
        lb = np.ascontiguousarray(lb, dtype=np.float32); rb = np.ascontiguousarray(rb, dtype=np.float32)
        sl = np.empty(self.n, np.float32); sr = np.empty(self.n, np.float32)
        tl = np.empty(self.n, np.uint16); tr = np.empty(self.n, np.uint16)
        self.o.L.orc_stream_update(self.h, lb.ctypes.data, rb.ctypes.data, 1 if modified else 0,
                                   sl.ctypes.data, sr.ctypes.data, tl.ctypes.data, tr.ctypes.data)
        return sl, sr, tl, tr

    def __del__(self):
        try: self.o.L.orc_stream_free(self.h)
        except Exception: pass


class Reference:
    """The reference's own compiled transforms (oracle/_ref/libglava_ref.so)."""

    @staticmethod
    def available():
        """built (or shipped prebuilt) AND loadable — a library that does not load must read as "absent", not kill the
        worker pools of bench.py's reference arm"""
        build()
        path = os.path.join(HERE, "_ref", "libglava_ref.so")
        if not os.path.exists(path):
            return False
        try:
            C.CDLL(path)
            return True
        except OSError:
            return False

    def __init__(self):
        L = C.CDLL(os.path.join(HERE, "_ref", "libglava_ref.so"))
        vp = C.c_void_p
        L.ref_chan_new.restype = vp
        L.ref_chan_new.argtypes = [C.c_float] * 4 + [C.c_int] * 2
        L.ref_chan_free.argtypes = [vp]
        for f in (L.ref_fft, L.ref_gravity, L.ref_average, L.ref_update_a):
            f.argtypes = [vp, vp, C.c_size_t]
        L.ref_wrange.argtypes = [vp, C.c_size_t]
        L.ref_parse_color.argtypes = [C.c_char_p, vp]
        L.ref_smooth.argtypes = [vp, C.c_size_t, C.c_float, C.c_float]
        if hasattr(L, "ref_ext_process"):
            L.ref_ext_process.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int,
                                          C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        self.L = L

    def ext_process(self, path, cd, cfd, dd, binds=None, avg_frames=5):
        """the reference's own glsl_ext.c on one file -> (processed source, [[request, arg, ...], ...]); raises ValueError
        when the reference reports a parse error (it would have called glava_abort)"""
        out = C.create_string_buffer(1 << 21); req = C.create_string_buffer(1 << 16)
        arr = (C.c_char_p * (len(binds) + 1))(*[b.encode() for b in binds], None) if binds else None
        rc = self.L.ref_ext_process(path.encode(), cd.encode(), cfd.encode() if cfd else None, dd.encode(), arr, avg_frames,
                                    out, len(out), req, len(req))
        if rc == 1:
            raise ValueError("the reference rejected %s (parse_error -> glava_abort)" % path)
        assert rc == 0, rc
        return out.value.decode(), [ln.split("|") for ln in req.value.decode().splitlines()]

    def chan(self, p):
        return self.L.ref_chan_new(p.fft_scale, p.fft_cutoff, p.gravity_step, p.ur, p.avg_frames, p.avg_window)

    def fft(self, ch, pcm):
        b = np.array(pcm, dtype=np.float32, copy=True)
        self.L.ref_fft(ch, b.ctypes.data, b.shape[0])
        return b

    def update_a(self, ch, pcm):
        b = np.array(pcm, dtype=np.float32, copy=True)
        self.L.ref_update_a(ch, b.ctypes.data, b.shape[0])
        return b

    def wrange(self, pcm):
        b = np.array(pcm, dtype=np.float32, copy=True)
        self.L.ref_wrange(b.ctypes.data, b.shape[0])
        return b

    def smooth(self, buf, smooth_distance=0.01, smooth_ratio=4.0):
        b = np.array(buf, dtype=np.float32, copy=True)
        self.L.ref_smooth(b.ctypes.data, b.shape[0], smooth_distance, smooth_ratio)
        return b

    def parse_color(self, s):
        out = np.zeros(4, np.float32); out[3] = 1.0
        ok = self.L.ref_parse_color(s.encode(), out.ctypes.data)
        return bool(ok), out


class ReferenceRenderer:
    """The reference's own rd_new / rd_update (render.c), run on a null OpenGL driver and a null window backend
    (oracle/ref_shim.c): what rd_new reads from a configuration, and the float buffers rd_update uploads as the audio
    textures — the outcome of its transform chain, buffer scaling and keyframe interpolation as IT orchestrates them."""

    INTS = ("bufsize", "rate", "samplesize", "mirror_input", "avg_frames", "avg_window", "smooth_pass", "accel_fft",
            "interpolate", "bufscale", "premultiply_alpha", "framerate", "w", "h", "stages", "copy_desktop")
    FLOATS = ("fft_scale", "fft_cutoff", "gravity_step", "smooth_factor", "smooth_distance", "smooth_ratio",
              "clear_r", "clear_g", "clear_b", "clear_a", "ur", "fr")
    _L = None

    @classmethod
    def lib(cls):
        if cls._L is None:
            path = os.path.join(HERE, "_ref", "libglava_ref_rd.so")
            if not os.path.exists(path):
                return None
            try:
                L = C.CDLL(path)                       # needs an executable stack (GNU nested-function trampolines)
            except OSError:
                return None
            cp, vp = C.c_char_p, C.c_void_p
            L.ref_rd_new.restype = vp
            L.ref_rd_new.argtypes = [C.POINTER(cp), cp, C.POINTER(cp)]
            L.ref_rd_config.argtypes = [vp, vp, vp]
            L.ref_rd_update.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
            L.ref_rd_upload.argtypes = [vp, C.c_int, C.POINTER(C.c_int), vp, C.c_int]
            L.ref_rd_set_rates.argtypes = [vp, C.c_float, C.c_float]
            L.ref_rd_destroy.argtypes = [vp]
            L.ref_rd_source.restype = C.c_char_p
            L.ref_rd_source.argtypes = [C.c_int]
            cls._L = L
        return cls._L

    def __init__(self, paths, entry="rc.glsl", requests=()):
        L = self.lib()
        assert L is not None
        pa = (C.c_char_p * (len(paths) + 1))(*[p.encode() for p in paths], None)
        rq = (C.c_char_p * (len(requests) + 1))(*[r.encode() for r in requests], None)
        self.L = L
        L.ref_rd_clear_sources()
        self.h = L.ref_rd_new(pa, entry.encode(), rq)
        if not self.h:
            raise ValueError("the reference aborted in rd_new")
        # every text rd_new handed to glShaderSource, in load order: vertex + fragment per module stage, then the util passes
        self.sources = [L.ref_rd_source(i).decode() for i in range(L.ref_rd_source_count())]
        self.cfg = self.config()
        self.lb = np.zeros(self.cfg["bufsize"], np.float32)            # glava.c:487-494: lb / rb persist across frames
        self.rb = np.zeros(self.cfg["bufsize"], np.float32)

    def config(self):
        ints = (C.c_int * 16)(); floats = (C.c_float * 12)()
        self.L.ref_rd_config(self.h, ints, floats)
        d = dict(zip(self.INTS, list(ints)))
        d.update(zip(self.FLOATS, [float(np.float32(v)) for v in floats]))
        return d

    def set_rates(self, ur, fr):
        self.L.ref_rd_set_rates(self.h, ur, fr)

    def frame(self, pcm_l=None, pcm_r=None):
        """one iteration of glava.c:523-539: copy new PCM into lb / rb when there is some (modified), rd_update;
        -> {0: floats uploaded for audio_l, 1: ... audio_r} (the LAST upload of each texture in this frame)"""
        modified = pcm_l is not None
        if modified:
            self.lb[:] = pcm_l; self.rb[:] = pcm_r
        k = self.L.ref_rd_update(self.h, self.lb.ctypes.data, self.rb.ctypes.data, self.lb.shape[0], 1 if modified else 0)
        if k < 0:
            raise ValueError("the reference aborted in rd_update")
        out = {}
        for i in range(k):
            which = C.c_int(); buf = np.zeros(self.lb.shape[0], np.float32)
            w = self.L.ref_rd_upload(self.h, i, C.byref(which), buf.ctypes.data, buf.shape[0])
            out[which.value] = buf[:w].copy()
        return out

    def close(self):
        if self.h:
            self.L.ref_rd_destroy(self.h); self.h = None
