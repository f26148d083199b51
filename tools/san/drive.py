"""Drives the sanitizer build of the host-only sources (see tools/san_host_check.sh): config reader + colour compiler over
the fuzz generators' configs and over hostile text, the --pipe line parser over random bytes, the audio feeder and the
batched FIFO gather over real named pipes.  Any ASan / UBSan report aborts the process."""
import ctypes as C
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from glava_b200.api import Params, _cstr_array          # noqa: E402  (ctypes mirrors only; the product library is not loaded)

L = C.CDLL(sys.argv[1])
vp, cp = C.c_void_p, C.c_char_p
L.glava_b200_last_error.restype = cp
L.glava_b200_load_config_binds.argtypes = [C.POINTER(Params), C.POINTER(cp), cp, C.POINTER(cp), cp, C.POINTER(cp)]
L.glava_b200_pipe_new.restype = vp
L.glava_b200_pipe_new.argtypes = [C.POINTER(cp), cp, C.POINTER(cp), cp, C.POINTER(cp)]
L.glava_b200_pipe_feed.argtypes = [vp, cp, C.c_size_t]
L.glava_b200_pipe_params.argtypes = [vp, C.POINTER(Params)]
L.glava_b200_pipe_apply.argtypes = [vp, vp]
L.glava_b200_pipe_free.argtypes = [vp]
L.san_renderer_new.restype = vp
L.san_renderer_new.argtypes = [C.c_int]
L.san_renderer_free.argtypes = [vp]
L.san_renderer_updates.argtypes = [vp]
L.san_renderer_updates.restype = C.c_long
L.glava_b200_audio_start.restype = vp
L.glava_b200_audio_start.argtypes = [cp, C.POINTER(cp), C.c_int, C.c_size_t, C.c_size_t, C.c_uint, C.c_int]
L.glava_b200_audio_frame.argtypes = [vp, vp]
L.glava_b200_audio_stop.argtypes = [vp]
L.glava_b200_fifo_open.restype = vp
L.glava_b200_fifo_open.argtypes = [C.POINTER(cp), C.c_int, C.c_size_t]
L.glava_b200_fifo_pump.argtypes = [vp, vp]
L.glava_b200_fifo_close.argtypes = [vp]


def load(paths, binds=None, requests=None, module=None):
    p = Params()
    rc = L.glava_b200_load_config_binds(C.byref(p), _cstr_array(paths), b"rc.glsl", _cstr_array(requests),
                                        module.encode() if module else None, _cstr_array(binds))
    return rc, p


def configs():
    import fuzz_module_configs as fz
    from tests.test_color_expr import _Gen
    rng = np.random.default_rng(1)
    n = 0
    for seed in range(150):
        module = ["bars", "radial", "circle", "graph", "wave"][seed % 5]
        d = tempfile.mkdtemp()
        open(d + "/rc.glsl", "w").write(f"#request mod {module}\n#request setgeometry 0 0 64 48\n")
        open(f"{d}/{module}.glsl", "w").write(fz.gen(np.random.default_rng(seed), module))
        rc, _ = load([d]); assert rc == 0, L.glava_b200_last_error()
        n += 1
    gen = _Gen(rng)
    for i in range(150):                                    # random colour expressions, truncated ones too
        expr = gen.vec(4, int(rng.integers(1, 5)))
        for text in (expr, expr[: int(rng.integers(1, len(expr)))], expr + ")", "((" + expr):
            d = tempfile.mkdtemp()
            open(d + "/rc.glsl", "w").write("#request mod bars\n")
            open(d + "/bars.glsl", "w").write(f"#define K 2 +\n#define COLOR {text}\n#define BAR_OUTLINE vec4(COLOR.rgb * K 1, COLOR.a)\n")
            load([d]); n += 1
    hostile = ["#define COLOR " + "(" * 300, "#define COLOR " + "vec4(" * 40 + "1" + ")" * 40, "#define COLOR #", "#define COLOR @:",
               "#define COLOR @fg", "#define A A\n#define COLOR vec4(A)", "#define A B\n#define B A\n#define COLOR vec4(A, B, A, B)",
               "#if\n#endif", "#if ((((\n#endif", "#elif 1", "#else", "#if 1 / 0\n#endif", "#ifdef\n#endif", "#define\n#undef",
               "#define COLOR vec4(1.e+, 0, 0, 1)", "#define COLOR vec4(0x, 0, 0, 1)", "#define COLOR vec4(1,,1)", "#define COLOR d.rgbargb",
               "#request", "#request setbg", "#request setbg " + "f" * 100, "#request setbgf 1 2 3", '#request settitle "unterminated',
               "#include", '#include ""', '#include "@x"', "#request transform a", "#request setgeometry 1 2 3 999999999999",
               "#define BAR_WIDTH 1e400\n#define BAR_GAP -1e400", "#define NBARS 0", "#define GRADIENT 0", "\x00\xff#define COLOR \xfe"]
    for h in hostile:
        for module in ("bars", "radial", "graph"):
            d = tempfile.mkdtemp()
            open(d + "/rc.glsl", "w").write(f"#request mod {module}\n")
            open(f"{d}/{module}.glsl", "wb").write(h.encode("latin-1") + b"\n")
            load([d], binds=["fg=#112233", "bg=vec4(1,"]); n += 1
    return n


def pipes():
    rng = np.random.default_rng(2)
    r = L.san_renderer_new(1)
    n = 0
    for args in (["fg", "bg", "amp:float", "on:bool", "k:int", "p2:vec2", "p3:vec3"], ["_"], []):
        h = L.glava_b200_pipe_new(None, None, None, b"bars", _cstr_array(args))
        assert h
        for _ in range(400):
            k = int(rng.integers(0, 6))
            if k == 0:
                data = bytes(rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8))
            elif k == 1:
                data = ("%s = %s\n" % (rng.choice(["fg", "bg", "amp", "on", "k", "p2", "p3", "", "f", "zz"]),
                                       rng.choice(["#ff8000", "#ff80", "#", "1,2,3,4", "1,,", "true", "nan", "1e99", "-", " " * 50]))).encode()
            elif k == 2:
                data = b"=" * int(rng.integers(0, 200)) + b"\n"
            elif k == 3:
                data = b"fg = #" + bytes(rng.integers(48, 103, int(rng.integers(0, 140)), dtype=np.uint8)) + b"\n"
            else:
                data = b"amp=" + str(float(rng.normal())).encode() + b"\nk = " + str(int(rng.integers(-2**40, 2**40))).encode() + b"\n"
            L.glava_b200_pipe_feed(h, data, len(data)); n += 1
            if rng.random() < 0.1:
                L.glava_b200_pipe_apply(h, r)
        p = Params(); L.glava_b200_pipe_params(h, C.byref(p))
        L.glava_b200_pipe_free(h)
    L.san_renderer_free(r)
    return n


def audio():
    d = tempfile.mkdtemp()
    batch, n, ssz = 5, 2048, 512
    paths = [f"{d}/s{i}.fifo" for i in range(batch)]
    for p in paths:
        os.mkfifo(p)
    r = L.san_renderer_new(batch)
    rng = np.random.default_rng(3)
    # host rings: native fifo backend threads + frame loop
    a = L.glava_b200_audio_start(b"fifo", _cstr_array(paths), batch, n, ssz, 22050, 2)
    assert a
    fds = [os.open(p, os.O_WRONLY) for p in paths]
    for _ in range(30):
        for fd in fds:
            os.write(fd, rng.integers(-32768, 32767, int(rng.integers(1, ssz)), dtype=np.int16).tobytes())   # ragged writes
        time.sleep(0.002)
        assert L.glava_b200_audio_frame(a, r) == 0
    for fd in fds:
        os.close(fd)
    assert L.glava_b200_audio_stop(a) == 0
    # device-ring path: batched gather with writers that come, stall and go
    f = L.glava_b200_fifo_open(_cstr_array(paths), batch, ssz)
    assert f
    stop = threading.Event()

    def writer(path, seed):
        g = np.random.default_rng(seed)
        fd = os.open(path, os.O_WRONLY)
        while not stop.is_set():
            os.write(fd, g.integers(-32768, 32767, int(g.integers(1, 700)), dtype=np.int16).tobytes())
            time.sleep(float(g.uniform(0, 0.004)))
        os.close(fd)
    ths = [threading.Thread(target=writer, args=(p, i)) for i, p in enumerate(paths[:-1])]      # the last stream stays silent
    for t in ths:
        t.start()
    for _ in range(60):
        assert L.glava_b200_fifo_pump(f, r) == 0
    stop.set()
    while any(t.is_alive() for t in ths):                        # a writer may sit in write() on a full pipe: keep draining
        assert L.glava_b200_fifo_pump(f, r) == 0
    for t in ths:
        t.join()
    for _ in range(3):
        assert L.glava_b200_fifo_pump(f, r) == 0                 # writers gone: silence, no spin
    L.glava_b200_fifo_close(f)
    ups = L.san_renderer_updates(r)
    L.san_renderer_free(r)
    return ups


if __name__ == "__main__":
    print("configs", configs(), flush=True)
    print("pipe feeds", pipes(), flush=True)
    print("audio updates", audio(), flush=True)
    print("sanitizer run clean")
