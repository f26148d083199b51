"""ctypes binding of include/glava_b200_audio.h: GLava's audio plug-in ABI (fifo.h:9-26), the batch of backend
threads with host rings (glava.c:487-537) and the batched FIFO reader that feeds glava_b200_ingest_fifo."""
import ctypes as C

import numpy as np

from .api import GlavaError, _check, _cstr_array, lib


class _Mutex(C.Structure):                     # pthread_mutex_t: 40 bytes, 8-aligned on x86-64 / aarch64 glibc
    _fields_ = [("_opaque", C.c_long * 5)]


class AudioData(C.Structure):
    """struct audio_data (glava/fifo.h:9-20) — same field order."""
    _fields_ = [("audio_out_r", C.POINTER(C.c_float)), ("audio_out_l", C.POINTER(C.c_float)),
                ("modified", C.c_bool), ("audio_buf_sz", C.c_size_t), ("sample_sz", C.c_size_t),
                ("format", C.c_int), ("rate", C.c_uint), ("source", C.c_char_p), ("channels", C.c_int),
                ("terminate", C.c_int), ("mutex", _Mutex)]


class AudioImpl(C.Structure):
    """struct audio_impl (glava/fifo.h:22-26)."""
    _fields_ = [("name", C.c_char_p), ("init", C.c_void_p), ("entry", C.c_void_p)]


_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        vp, cp, i32 = C.c_void_p, C.c_char_p, C.c_int
        L.glava_b200_audio_register.argtypes = [vp]
        L.glava_b200_audio_find.argtypes = [cp]
        L.glava_b200_audio_find.restype = vp
        L.glava_b200_audio_start.argtypes = [cp, C.POINTER(cp), i32, C.c_size_t, C.c_size_t, C.c_uint, i32]
        L.glava_b200_audio_start.restype = vp
        L.glava_b200_audio_collect.argtypes = [vp, vp, vp, vp]
        L.glava_b200_audio_frame.argtypes = [vp, vp]
        L.glava_b200_audio_stream.argtypes = [vp, i32]
        L.glava_b200_audio_stream.restype = C.POINTER(AudioData)
        L.glava_b200_audio_stop.argtypes = [vp]
        L.glava_b200_fifo_open.argtypes = [C.POINTER(cp), i32, C.c_size_t]
        L.glava_b200_fifo_open.restype = vp
        L.glava_b200_fifo_gather.argtypes = [vp, vp, vp]
        L.glava_b200_fifo_timeout_ms.argtypes = [vp]
        L.glava_b200_fifo_pump.argtypes = [vp, vp]
        L.glava_b200_fifo_close.argtypes = [vp]
        _bound = True
    return L


def _err():
    return GlavaError(lib().glava_b200_last_error().decode(errors="replace"))


def register_backend(impl_address):
    """register_audio_impl (fifo.h:33): `impl_address` = address of a `struct audio_impl` that outlives the process"""
    _check(_L().glava_b200_audio_register(C.c_void_p(impl_address)))


def find_backend(name):
    """the `-a NAME` lookup (glava.c:469-479); raises with the reference's message when absent"""
    p = _L().glava_b200_audio_find(name.encode())
    if not p:
        raise _err()
    return p


class AudioBatch:
    """`batch` backend threads filling host rings — glava.c:487-520 per stream."""

    def __init__(self, backend, sources, batch, bufsz, samplesz=1024, rate=22050, channels=2):
        self._Lib = _L()
        self.batch, self.bufsz = int(batch), int(bufsz)
        self._h = self._Lib.glava_b200_audio_start(backend.encode(), _cstr_array(sources), self.batch, self.bufsz,
                                                   int(samplesz), int(rate), int(channels))
        if not self._h:
            raise _err()

    def collect(self, lb, rb):
        """glava.c:528-537 for every stream; returns the per-stream modified flags"""
        assert lb.dtype == np.float32 and lb.shape == (self.batch, self.bufsz) and lb.flags.c_contiguous
        assert rb.dtype == np.float32 and rb.shape == lb.shape and rb.flags.c_contiguous
        m = np.zeros(self.batch, np.uint8)
        n = self._Lib.glava_b200_audio_collect(self._h, lb.ctypes.data, rb.ctypes.data, m.ctypes.data)
        if n < 0:
            raise _err()
        return m.astype(bool)

    def frame(self, renderer):
        _check(self._Lib.glava_b200_audio_frame(self._h, renderer._h))

    def stream(self, s):
        return self._Lib.glava_b200_audio_stream(self._h, int(s)).contents

    def stop(self):
        if self._h:
            h, self._h = self._h, None
            _check(self._Lib.glava_b200_audio_stop(h))

    def __enter__(self): return self
    def __exit__(self, *a): self.stop()


class FifoReader:
    """One poll() over every stream's FIFO per tick -> [batch][samplesz / 2] int16 chunks for ingest_fifo."""

    def __init__(self, sources, samplesz=1024):
        self._Lib = _L()
        self.batch, self.samplesz = len(sources), int(samplesz)
        self._h = self._Lib.glava_b200_fifo_open(_cstr_array(sources), self.batch, self.samplesz)
        if not self._h:
            raise _err()

    def gather(self):
        """-> (chunks int16 [batch][samplesz / 2], fresh bool [batch])"""
        chunks = np.empty((self.batch, self.samplesz // 2), np.int16)
        fresh = np.zeros(self.batch, np.uint8)
        if self._Lib.glava_b200_fifo_gather(self._h, chunks.ctypes.data, fresh.ctypes.data) < 0:
            raise _err()
        return chunks, fresh.astype(bool)

    @property
    def timeout_ms(self):
        return self._Lib.glava_b200_fifo_timeout_ms(self._h)

    def pump(self, renderer):
        _check(self._Lib.glava_b200_fifo_pump(self._h, renderer._h))

    def close(self):
        if self._h:
            self._Lib.glava_b200_fifo_close(self._h)
            self._h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()
