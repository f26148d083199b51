"""ctypes binding of include/glava_b200.h."""
import ctypes as C
import os

import numpy as np

MODULES = ("bars", "radial", "circle", "graph", "wave", "test")
_HERE = os.path.dirname(os.path.abspath(__file__))


class GlavaError(RuntimeError):
    pass


class Color(C.Structure):
    _fields_ = [("mode", C.c_int), ("lo", C.c_float * 4), ("hi", C.c_float * 4), ("gradient", C.c_float)]


class ColorOp(C.Structure):
    _fields_ = [("op", C.c_uint8), ("dst", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8), ("imm", C.c_float)]


class ColorProg(C.Structure):
    """glava_b200_color_prog: a compiled colour expression (mode 2 of the colour it belongs to)"""
    _fields_ = [("n_ops", C.c_int), ("result", C.c_int), ("ops", ColorOp * 64)]


class Params(C.Structure):
    """struct glava_b200_params (include/glava_b200.h) — same field order."""
    _fields_ = [
        ("n", C.c_int), ("fft_scale", C.c_float), ("fft_cutoff", C.c_float), ("gravity_step", C.c_float),
        ("ur", C.c_float), ("avg_frames", C.c_int), ("avg_window", C.c_int), ("accel_fft", C.c_int),
        ("smooth_pass", C.c_int), ("smooth_factor", C.c_float), ("sample_range", C.c_float),
        ("sample_scale", C.c_float), ("hybrid_weight", C.c_float), ("sample_mode", C.c_int),
        ("round_formula", C.c_int),
        ("module", C.c_int), ("w", C.c_int), ("h", C.c_int), ("channels", C.c_int), ("premultiply_alpha", C.c_int),
        ("bars_width", C.c_float), ("bars_gap", C.c_float), ("bars_outline_width", C.c_float), ("bars_amplify", C.c_float),
        ("bars_color", Color), ("bars_outline_mode", C.c_int), ("bars_outline", C.c_float * 4),
        ("bars_direction", C.c_int), ("bars_invert", C.c_int), ("bars_flip", C.c_int), ("bars_mirror_yx", C.c_int),
        ("radial_radius", C.c_float), ("radial_line", C.c_float), ("radial_line_half", C.c_float),
        ("radial_outline", C.c_float * 4), ("radial_nbars", C.c_int), ("radial_bar_width", C.c_float),
        ("radial_amplify", C.c_float), ("radial_color", Color), ("radial_rotate", C.c_float), ("radial_invert", C.c_int),
        ("radial_bar_alias", C.c_float), ("radial_c_alias", C.c_float), ("radial_off_x", C.c_float), ("radial_off_y", C.c_float),
        ("circle_radius", C.c_float), ("circle_line", C.c_float), ("circle_outline", C.c_float * 4),
        ("circle_amplify", C.c_float), ("circle_rotate", C.c_float), ("circle_invert", C.c_int),
        ("circle_fill", C.c_int), ("circle_smooth", C.c_int),
        ("graph_vscale", C.c_float), ("graph_direction", C.c_int), ("graph_color", Color),
        ("graph_draw_outline", C.c_int), ("graph_draw_highlight", C.c_int), ("graph_outline", C.c_float * 4),
        ("graph_invert", C.c_int),
        ("wave_min_thickness", C.c_float), ("wave_max_thickness", C.c_float), ("wave_base_color", C.c_float * 4),
        ("wave_amplify", C.c_float), ("wave_outline", C.c_float * 4),
        ("rate_request", C.c_int), ("samplesize_request", C.c_int),
        ("fb_slots", C.c_int), ("lazy_smooth", C.c_int),
        ("bufscale", C.c_int), ("interpolate", C.c_int), ("fr", C.c_float), ("transform_smooth", C.c_int),
        ("smooth_distance", C.c_float), ("smooth_ratio", C.c_float),
        ("bars_color_prog", ColorProg), ("bars_outline_prog", ColorProg), ("radial_color_prog", ColorProg),
        ("graph_color_prog", ColorProg),
        ("clear_color", C.c_float * 4),
        ("radial_bar_width_int", C.c_int), ("radial_bar_outline_width", C.c_float), ("radial_bar_outline", C.c_float * 4),
        ("graph_join_channels", C.c_int), ("graph_anti_alias", C.c_int), ("shader_pre_smoothed", C.c_int),
        ("mirror_input", C.c_int),
    ]

    def copy(self):
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q

    @property
    def module_name(self):
        return MODULES[self.module]

    def to_dict(self):
        """JSON-serialisable copy (nested structs as dicts, arrays as lists); floats survive exactly"""
        def conv(v):
            if isinstance(v, ColorProg):                         # only the instructions in use
                return {"n_ops": v.n_ops, "result": v.result, "ops": [conv(v.ops[i]) for i in range(v.n_ops)]}
            if isinstance(v, C.Structure):
                return {n: conv(getattr(v, n)) for n, _t in v._fields_}
            if isinstance(v, C.Array):
                return [conv(x) for x in v]
            return v
        return conv(self)

    @classmethod
    def from_dict(cls, d):
        def fill(dst, src):
            for n, _t in dst._fields_:
                if n not in src:
                    continue
                cur = getattr(dst, n)
                if isinstance(cur, C.Structure):
                    fill(cur, src[n])
                elif isinstance(cur, C.Array):
                    for i, x in enumerate(src[n]):
                        if isinstance(cur[i], C.Structure):
                            fill(cur[i], x)
                        else:
                            cur[i] = x
                else:
                    setattr(dst, n, src[n])
        q = cls()
        fill(q, d)
        return q


_lib = None


def lib_path():
    return os.path.join(_HERE, "libglava_b200.so")


def lib():
    """Load libglava_b200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise GlavaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(the B200 path has no CPU fallback)")
    L = C.CDLL(path)
    vp, cp, i32 = C.c_void_p, C.c_char_p, C.c_int
    L.glava_b200_last_error.restype = cp
    L.glava_b200_version.restype = cp
    L.glava_b200_default_params.argtypes = [C.POINTER(Params), cp]
    L.glava_b200_load_config.argtypes = [C.POINTER(Params), C.POINTER(cp), cp, C.POINTER(cp), cp]
    L.glava_b200_load_config_binds.argtypes = [C.POINTER(Params), C.POINTER(cp), cp, C.POINTER(cp), cp, C.POINTER(cp)]
    L.glava_b200_reconfigure.argtypes = [vp, C.POINTER(Params)]
    L.glava_b200_new.restype = vp
    L.glava_b200_new.argtypes = [C.POINTER(Params), i32, i32]
    L.glava_b200_destroy.argtypes = [vp]
    L.glava_b200_get_params.argtypes = [vp, C.POINTER(Params)]
    L.glava_b200_batch.argtypes = [vp]
    L.glava_b200_module_name.argtypes = [vp]
    L.glava_b200_module_name.restype = cp
    L.glava_b200_host_alloc.restype = vp
    L.glava_b200_host_alloc.argtypes = [C.c_size_t]
    L.glava_b200_host_free.argtypes = [vp]
    L.glava_b200_host_alloc_on.restype = vp
    L.glava_b200_host_alloc_on.argtypes = [C.c_size_t, i32]
    L.glava_b200_device_numa_node.argtypes = [i32]
    L.glava_b200_bind_thread_to_device.argtypes = [i32]
    L.glava_b200_shard_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.glava_b200_new_sharded.restype = vp
    L.glava_b200_new_sharded.argtypes = [C.POINTER(Params), i32, C.c_uint64]
    L.glava_b200_new_sharded_devices.restype = vp
    L.glava_b200_new_sharded_devices.argtypes = [C.POINTER(Params), i32, C.POINTER(i32), i32]
    L.glava_b200_sharded_destroy.argtypes = [vp]
    L.glava_b200_sharded_shards.argtypes = [vp]
    L.glava_b200_sharded_batch.argtypes = [vp]
    L.glava_b200_sharded_shard.restype = vp
    L.glava_b200_sharded_shard.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.glava_b200_sharded_update.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.glava_b200_sharded_rerender.argtypes = [vp]
    L.glava_b200_sharded_ingest_fifo.argtypes = [vp, vp, i32]
    L.glava_b200_sharded_sync.argtypes = [vp]
    L.glava_b200_sharded_readback.argtypes = [vp, i32, vp]
    L.glava_b200_sharded_frame_device.restype = vp
    L.glava_b200_sharded_frame_device.argtypes = [vp, i32, C.POINTER(i32)]
    L.glava_b200_sharded_textures.argtypes = [vp, vp, vp]
    L.glava_b200_set_async_input.argtypes = [vp, i32]
    L.glava_b200_wait_input.argtypes = [vp]
    L.glava_b200_update.argtypes = [vp, vp, vp, C.c_size_t, i32]
    L.glava_b200_update_device.argtypes = [vp, vp, vp, C.c_size_t, i32]
    L.glava_b200_update_device_after.argtypes = [vp, vp, vp, C.c_size_t, i32, vp]
    L.glava_b200_input_event.restype = vp
    L.glava_b200_input_event.argtypes = [vp]
    L.glava_b200_update_masked.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.glava_b200_update_device_masked.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.glava_b200_update_rings_masked.argtypes = [vp, vp]
    L.glava_b200_ingest_fifo.argtypes = [vp, vp, i32]
    L.glava_b200_ingest_float.argtypes = [vp, vp, i32]
    L.glava_b200_update_rings.argtypes = [vp, i32]
    L.glava_b200_sync.argtypes = [vp]
    L.glava_b200_readback.argtypes = [vp, i32, vp]
    L.glava_b200_readback_async.argtypes = [vp, i32, vp]
    L.glava_b200_readback_fence.argtypes = [vp]
    L.glava_b200_spectrum.argtypes = [vp, vp, vp]
    L.glava_b200_textures.argtypes = [vp, vp, vp]
    L.glava_b200_framebuffer_device.argtypes = [vp]
    L.glava_b200_framebuffer_device.restype = vp
    L.glava_b200_cuda_stream.argtypes = [vp]
    L.glava_b200_cuda_stream.restype = vp
    L.glava_b200_smooth_pass.argtypes = [vp, vp, vp, i32]
    L.glava_b200_raster_textures.argtypes = [vp, vp, vp]
    L.glava_b200_transform_smooth.argtypes = [vp, vp, i32]
    L.glava_b200_timeline.argtypes = [vp, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.glava_b200_sizereq.argtypes = [vp, i32, i32]
    L.glava_b200_wait_frame.argtypes = [vp]
    L.glava_b200_frame_event.argtypes = [vp]
    L.glava_b200_frame_event.restype = vp
    L.glava_b200_frame_device.argtypes = [vp, i32]
    L.glava_b200_frame_device.restype = vp
    L.glava_b200_framebuffer_ipc.argtypes = [vp, vp, C.c_size_t]
    L.glava_b200_spectrum_size.argtypes = [vp]
    L.glava_b200_launch_count.argtypes = [vp]
    L.glava_b200_launch_count.restype = C.c_uint64
    L.glava_b200_set_abort_hook.argtypes = [vp]
    L.glava_b200_set_timing.argtypes = [vp, i32]
    L.glava_b200_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(i32)]
    L.glava_b200_pipe_new.restype = vp
    L.glava_b200_pipe_new.argtypes = [C.POINTER(cp), cp, C.POINTER(cp), cp, C.POINTER(cp)]
    L.glava_b200_pipe_feed.argtypes = [vp, cp, C.c_size_t]
    L.glava_b200_pipe_params.argtypes = [vp, C.POINTER(Params)]
    L.glava_b200_pipe_bind_count.argtypes = [vp]
    L.glava_b200_pipe_bind.argtypes = [vp, i32, C.POINTER(cp), C.POINTER(cp), C.POINTER(C.c_float)]
    L.glava_b200_pipe_apply.argtypes = [vp, vp]
    L.glava_b200_pipe_free.argtypes = [vp]
    _lib = L
    # errors surface as Python exceptions; keep stderr quiet
    _quiet = C.CFUNCTYPE(None, cp)(lambda msg: None)
    L._quiet_hook = _quiet
    L.glava_b200_set_abort_hook(C.cast(_quiet, vp))
    return L


def _check(rc):
    if rc != 0:
        raise GlavaError(lib().glava_b200_last_error().decode(errors="replace") or f"error {rc}")


def default_params(module="bars", **overrides):
    p = Params()
    _check(lib().glava_b200_default_params(C.byref(p), module.encode()))
    for k, v in overrides.items():
        if isinstance(v, (list, tuple)):
            v = type(getattr(p, k))(*v)
        setattr(p, k, v)
    return p


def _cstr_array(items):
    if items is None:
        return None
    arr = (C.c_char_p * (len(items) + 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode()
    arr[len(items)] = None
    return arr


def load_config(paths=None, entry="rc.glsl", requests=None, force_module=None, binds=None):
    """rd_new's config half (render.c:1322-1435): read `entry` from the first of `paths`, apply
    `#request`s and `requests` (CLI --request strings), then the module's `#define`s.
    binds: {"fg": "#ff0000", ...} — `--pipe` binds resolving `@name:default` macros."""
    p = Params()
    bl = [f"{k}={v}" for k, v in binds.items()] if binds else None
    _check(lib().glava_b200_load_config_binds(C.byref(p), _cstr_array(paths), entry.encode() if entry else None,
                                              _cstr_array(requests), force_module.encode() if force_module else None,
                                              _cstr_array(bl)))
    return p


class Pipe:
    """live `--pipe NAME[:TYPE]` binds fed with stdin text (glava.c:338-411, render.c:1846-2100)"""

    def __init__(self, pipe_args, paths=None, entry="rc.glsl", requests=None, force_module=None):
        self._L = lib()
        self._h = self._L.glava_b200_pipe_new(_cstr_array(paths), entry.encode() if entry else None, _cstr_array(requests),
                                              force_module.encode() if force_module else None, _cstr_array(list(pipe_args)))
        if not self._h:
            raise GlavaError(self._L.glava_b200_last_error().decode(errors="replace"))
        self.messages = []

    def feed(self, text):
        """-> number of binds that took a new value; parse complaints (reference wording) are appended to .messages"""
        data = text.encode() if isinstance(text, str) else bytes(text)
        n = self._L.glava_b200_pipe_feed(self._h, data, len(data))
        if n < 0:
            raise GlavaError(self._L.glava_b200_last_error().decode(errors="replace"))
        return n

    def params(self):
        p = Params()
        _check(self._L.glava_b200_pipe_params(self._h, C.byref(p)))
        return p

    def binds(self):
        out = {}
        for i in range(self._L.glava_b200_pipe_bind_count(self._h)):
            name, typ, val = C.c_char_p(), C.c_char_p(), (C.c_float * 4)()
            _check(self._L.glava_b200_pipe_bind(self._h, i, C.byref(name), C.byref(typ), val))
            out[name.value.decode()] = (typ.value.decode(), tuple(val))
        return out

    def apply(self, renderer):
        _check(self._L.glava_b200_pipe_apply(self._h, renderer._h))
        _check(self._L.glava_b200_get_params(renderer._h, C.byref(renderer.params)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.glava_b200_pipe_free(self._h)
            self._h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()


_pinned = []


def pinned_empty(shape, dtype, device=None):
    """numpy array backed by pinned host memory on the NUMA node of `device` (default: the current CUDA device);
    lives until process exit."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = lib().glava_b200_host_alloc(nbytes) if device is None else lib().glava_b200_host_alloc_on(nbytes, int(device))
    if not ptr:
        raise GlavaError("pinned allocation failed")
    buf = (C.c_byte * nbytes).from_address(ptr)
    _pinned.append(buf)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Renderer:
    """Batched renderer handle: the role of struct glava_renderer (render.h:8-30)."""

    def __init__(self, params, batch=1, device=0):
        self._L = lib()
        self._h = self._L.glava_b200_new(C.byref(params), int(batch), int(device))
        if not self._h:
            raise GlavaError(self._L.glava_b200_last_error().decode(errors="replace"))
        self.params = Params()
        _check(self._L.glava_b200_get_params(self._h, C.byref(self.params)))
        self.batch = int(batch)
        self.device = int(device)
        self.nsz = int(self._L.glava_b200_spectrum_size(self._h))      # params.n / bufscale: spectrum / texture entries

    # -- rd_update ---------------------------------------------------------------------------------
    def update(self, lb, rb=None, modified=True):
        """lb, rb: host float32 arrays [batch][n] (ring contents oldest-first)."""
        lb = np.ascontiguousarray(lb, dtype=np.float32)
        assert lb.shape == (self.batch, self.params.n), lb.shape
        rp = None
        if rb is not None:
            rb = np.ascontiguousarray(rb, dtype=np.float32)
            assert rb.shape == lb.shape
            rp = rb.ctypes.data
        _check(self._L.glava_b200_update(self._h, lb.ctypes.data, rp, self.params.n, 1 if modified else 0))
        self._after_update()

    def _mask(self, modified):
        m = np.ascontiguousarray(np.asarray(modified) != 0, dtype=np.uint8)
        assert m.shape == (self.batch,), m.shape
        return m

    def update_masked(self, lb, rb, modified):
        """per-stream `modified` flags ([batch] booleans): glava.c:528-537 per stream"""
        lb = np.ascontiguousarray(lb, dtype=np.float32); rb = np.ascontiguousarray(rb, dtype=np.float32)
        assert lb.shape == (self.batch, self.params.n) and rb.shape == lb.shape
        m = self._mask(modified)
        _check(self._L.glava_b200_update_masked(self._h, lb.ctypes.data, rb.ctypes.data, self.params.n, m.ctypes.data))
        self._after_update()

    def update_device_masked(self, d_lb, d_rb, modified):
        m = self._mask(modified)
        _check(self._L.glava_b200_update_device_masked(self._h, d_lb, d_rb, self.params.n, m.ctypes.data))
        self._after_update()

    def update_rings_masked(self, modified):
        m = self._mask(modified)
        _check(self._L.glava_b200_update_rings_masked(self._h, m.ctypes.data))
        self._after_update()

    def _after_update(self):
        if getattr(self, "_resize_pending", False):          # a glava_b200_sizereq was applied by this update
            self._resize_pending = False
            self.refresh_params()

    def reconfigure(self, params):
        """live parameter update (colours, amplify, smoothing...): the `--pipe` analogue"""
        _check(self._L.glava_b200_reconfigure(self._h, C.byref(params)))
        _check(self._L.glava_b200_get_params(self._h, C.byref(self.params)))

    def update_device(self, d_lb, d_rb, modified=True):
        """d_lb, d_rb: integer device addresses of [batch][n] float32 (e.g. torch tensor.data_ptr())."""
        _check(self._L.glava_b200_update_device(self._h, d_lb, d_rb, self.params.n, 1 if modified else 0))
        self._after_update()

    def update_device_after(self, d_lb, d_rb, modified, ready_event):
        """update_device ordered after `ready_event` (a cudaEvent_t handle, e.g. torch.cuda.Event().cuda_event)"""
        _check(self._L.glava_b200_update_device_after(self._h, d_lb, d_rb, self.params.n, 1 if modified else 0, ready_event))
        self._after_update()

    @property
    def input_event(self):
        return self._L.glava_b200_input_event(self._h)

    def ingest_fifo(self, chunks):
        chunks = np.ascontiguousarray(chunks, dtype=np.int16)
        assert chunks.ndim == 2 and chunks.shape[0] == self.batch and chunks.shape[1] % 2 == 0
        _check(self._L.glava_b200_ingest_fifo(self._h, chunks.ctypes.data, chunks.shape[1] // 2))

    def ingest_float(self, chunks):
        """PulseAudio semantics (pulse_input.c:146-174): float samples as they are, [batch][frames*2] interleaved"""
        chunks = np.ascontiguousarray(chunks, dtype=np.float32)
        assert chunks.ndim == 2 and chunks.shape[0] == self.batch and chunks.shape[1] % 2 == 0
        _check(self._L.glava_b200_ingest_float(self._h, chunks.ctypes.data, chunks.shape[1] // 2))

    def update_rings(self, modified=True):
        _check(self._L.glava_b200_update_rings(self._h, 1 if modified else 0))
        self._after_update()

    def set_async_input(self, enable=True):
        _check(self._L.glava_b200_set_async_input(self._h, 1 if enable else 0))

    def wait_input(self):
        _check(self._L.glava_b200_wait_input(self._h))

    def sync(self):
        _check(self._L.glava_b200_sync(self._h))

    # -- outputs -----------------------------------------------------------------------------------
    def readback(self, stream=0, out=None):
        p = self.params
        if out is None:
            out = np.empty((p.h, p.w, 4), dtype=np.uint8)
        _check(self._L.glava_b200_readback(self._h, int(stream), out.ctypes.data))
        return out

    def readback_async(self, stream, out):
        """enqueue the D2H copy of one frame into `out` (pinned uint8 [h][w][4]); valid after sync()"""
        _check(self._L.glava_b200_readback_async(self._h, int(stream), out.ctypes.data))

    def readback_fence(self):
        """order later work on cuda_stream after the read-backs issued so far (their D2H runs on a separate stream)"""
        _check(self._L.glava_b200_readback_fence(self._h))

    def spectrum(self):
        l = np.empty((self.batch, self.nsz), dtype=np.float32); r = np.empty_like(l)
        _check(self._L.glava_b200_spectrum(self._h, l.ctypes.data, r.ctypes.data))
        return l, r

    def textures(self):
        l = np.empty((self.batch, self.nsz), dtype=np.uint16); r = np.empty_like(l)
        _check(self._L.glava_b200_textures(self._h, l.ctypes.data, r.ctypes.data))
        return l, r

    def smooth_pass(self, tex):
        tex = np.ascontiguousarray(tex, dtype=np.uint16)
        assert tex.ndim == 2 and tex.shape[1] == self.nsz
        out = np.empty_like(tex)
        _check(self._L.glava_b200_smooth_pass(self._h, tex.ctypes.data, out.ctypes.data, tex.shape[0]))
        return out

    def transform_smooth(self, planes):
        """transform_smooth (render.c:694-718) on host float32 [count][nsz]; returns the transformed copy"""
        b = np.array(planes, dtype=np.float32, copy=True)
        assert b.ndim == 2 and b.shape[1] == self.nsz
        _check(self._L.glava_b200_transform_smooth(self._h, b.ctypes.data, b.shape[0]))
        return b

    def raster_textures(self, tex_l, tex_r=None):
        tex_l = np.ascontiguousarray(tex_l, dtype=np.uint16)
        assert tex_l.shape == (self.batch, self.nsz)
        rp = None
        if tex_r is not None:
            tex_r = np.ascontiguousarray(tex_r, dtype=np.uint16)
            rp = tex_r.ctypes.data
        _check(self._L.glava_b200_raster_textures(self._h, tex_l.ctypes.data, rp))

    def timeline(self, cap=4096):
        """(spectrum [k][2], raster [m][2]) start/end ms of the timed launches (set_timing(True) first)"""
        out = np.zeros((cap, 2), np.float64); ns, nq = C.c_int(), C.c_int()
        _check(self._L.glava_b200_timeline(self._h, out.ctypes.data, cap, C.byref(ns), C.byref(nq)))
        return out[: ns.value].copy(), out[ns.value: ns.value + nq.value].copy()

    # -- offscreen hand-off (glava_sizereq / glava_wait / glava_tex) ---------------------------------
    def sizereq(self, w, h):
        """resize request, applied at the start of the next update"""
        _check(self._L.glava_b200_sizereq(self._h, int(w), int(h)))
        self._resize_pending = True

    def refresh_params(self):
        _check(self._L.glava_b200_get_params(self._h, C.byref(self.params)))
        return self.params

    def wait_frame(self):
        _check(self._L.glava_b200_wait_frame(self._h))

    @property
    def frame_event(self):
        return self._L.glava_b200_frame_event(self._h)

    def frame_device(self, stream):
        return self._L.glava_b200_frame_device(self._h, int(stream))

    def framebuffer_ipc(self):
        buf = (C.c_ubyte * 64)()
        _check(self._L.glava_b200_framebuffer_ipc(self._h, buf, 64))
        return bytes(buf)

    @property
    def framebuffer_device(self):
        return self._L.glava_b200_framebuffer_device(self._h)

    @property
    def cuda_stream(self):
        return self._L.glava_b200_cuda_stream(self._h)

    @property
    def launch_count(self):
        return int(self._L.glava_b200_launch_count(self._h))

    def set_timing(self, enable=True):
        _check(self._L.glava_b200_set_timing(self._h, 1 if enable else 0))

    def kernel_times(self):
        """-> dict(spectrum_ms, spectrum_launches, raster_ms, raster_launches) since set_timing(True)"""
        s, q = C.c_double(), C.c_double()
        ns, nq = C.c_int(), C.c_int()
        _check(self._L.glava_b200_kernel_times(self._h, C.byref(s), C.byref(ns), C.byref(q), C.byref(nq)))
        return dict(spectrum_ms=s.value, spectrum_launches=ns.value, raster_ms=q.value, raster_launches=nq.value)

    # -- rd_destroy --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.glava_b200_destroy(self._h)
            self._h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
    def __del__(self):
        try: self.close()
        except Exception: pass


class ShardedRenderer:
    """One handle for a batch spread over several GPUs of a node (glava_b200_new_sharded): contiguous stream blocks, one
    worker thread per device, no inter-device traffic.  devices: list of CUDA ordinals (may repeat), or None = all visible."""

    def __init__(self, params, batch, devices=None):
        self._L = lib()
        if devices is None:
            self._h = self._L.glava_b200_new_sharded(C.byref(params), int(batch), 0)
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._h = self._L.glava_b200_new_sharded_devices(C.byref(params), int(batch), arr, len(devices))
        if not self._h:
            raise GlavaError(self._L.glava_b200_last_error().decode(errors="replace"))
        self.params = params.copy()
        self.batch = int(batch)

    @property
    def shards(self):
        """[(device, first_stream, count)]"""
        out = []
        for k in range(self._L.glava_b200_sharded_shards(self._h)):
            d, f, c = C.c_int(), C.c_int(), C.c_int()
            self._L.glava_b200_sharded_shard(self._h, k, C.byref(d), C.byref(f), C.byref(c))
            out.append((d.value, f.value, c.value))
        return out

    def update(self, lb, rb, modified=None):
        lb = np.ascontiguousarray(lb, dtype=np.float32); rb = np.ascontiguousarray(rb, dtype=np.float32)
        assert lb.shape == (self.batch, self.params.n) and rb.shape == lb.shape
        m = None
        if modified is not None:
            m = np.ascontiguousarray(np.asarray(modified) != 0, dtype=np.uint8)
            assert m.shape == (self.batch,)
        _check(self._L.glava_b200_sharded_update(self._h, lb.ctypes.data, rb.ctypes.data, self.params.n, m.ctypes.data if m is not None else None))

    def rerender(self):
        _check(self._L.glava_b200_sharded_rerender(self._h))

    def ingest_fifo(self, chunks):
        chunks = np.ascontiguousarray(chunks, dtype=np.int16)
        assert chunks.ndim == 2 and chunks.shape[0] == self.batch and chunks.shape[1] % 2 == 0
        _check(self._L.glava_b200_sharded_ingest_fifo(self._h, chunks.ctypes.data, chunks.shape[1] // 2))

    def sync(self):
        _check(self._L.glava_b200_sharded_sync(self._h))

    def readback(self, stream, out=None):
        p = self.params
        if out is None:
            out = np.empty((p.h, p.w, 4), dtype=np.uint8)
        _check(self._L.glava_b200_sharded_readback(self._h, int(stream), out.ctypes.data))
        return out

    def textures(self):
        n = self.params.n // max(self.params.bufscale, 1)
        l = np.empty((self.batch, n), dtype=np.uint16); r = np.empty_like(l)
        _check(self._L.glava_b200_sharded_textures(self._h, l.ctypes.data, r.ctypes.data))
        return l, r

    def close(self):
        if getattr(self, "_h", None):
            self._L.glava_b200_sharded_destroy(self._h)
            self._h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
    def __del__(self):
        try: self.close()
        except Exception: pass
