"""-m gpu: the whole path PCM -> pixels through glava_b200_update, against the oracle chain."""
import numpy as np
import pytest

import glava_b200 as g
from oracle.oracle import OracleChannel, params_from

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("module,n,w,h", [("bars", 4096, 1920, 1080), ("radial", 2048, 800, 600), ("circle", 2048, 800, 600),
                                          ("graph", 2048, 1280, 720), ("wave", 2048, 1280, 720)])
def test_end_to_end_pixels(orc_pm, module, n, w, h, built):
    """GPU spectrum and oracle spectrum agree to ~1e-6, so after R16 quantisation a few texels differ by
    1 LSB16 and a handful of bar-top / edge pixels may flip; everything else must be identical.
    (Raster parity proper is bit-exact on identical textures: tests/test_gpu_raster.py.)"""
    batch = 4
    p = g.default_params(module, n=n, w=w, h=h); op = params_from(p)
    rings = g.StreamRings(batch, n)
    fft = module != "wave"
    chans = [[OracleChannel(orc_pm, op), OracleChannel(orc_pm, op)] for _ in range(batch)]
    with g.Renderer(p, batch=batch) as r:
        for _ in range(n // 256 + 6):
            rings.advance()
            r.update(rings.lb, rings.rb, True)
            tex = [[chans[s][0].update(rings.lb[s], fft)[1], chans[s][1].update(rings.rb[s], fft)[1]] for s in range(batch)]
        gl, gr = r.textures()
        for s in range(batch):
            got = r.readback(s).astype(int)
            same_tex = orc_pm.raster(op, gl[s], gr[s] if fft else gl[s])
            assert np.array_equal(got, same_tex.astype(int))                  # exact on the GPU's own textures
            want = orc_pm.raster(op, tex[s][0], tex[s][1]).astype(int)
            bad = (np.abs(got - want).max(axis=2) > 1).sum()
            assert bad <= 2e-4 * w * h, (module, s, bad)


def test_lazy_smooth_gives_identical_frames(built):
    for module, n, w, h in (("bars", 4096, 1920, 1080), ("radial", 2048, 640, 480), ("graph", 2048, 640, 360), ("wave", 1024, 640, 360),
                            ("circle", 1024, 320, 240)):
        batch = 3
        frames = []
        for lazy in (0, 1):
            p = g.default_params(module, n=n, w=w, h=h, lazy_smooth=lazy)
            rings = g.StreamRings(batch, n)
            with g.Renderer(p, batch=batch) as r:
                for _ in range(8):
                    rings.advance(); r.update(rings.lb, rings.rb, True)
                frames.append([r.readback(s) for s in range(batch)])
        for s in range(batch):
            assert np.array_equal(frames[0][s], frames[1][s]), (module, s)


def test_device_pointer_entry_point_matches_host_entry_point(built):
    import torch
    n, batch = 2048, 8
    p = g.default_params("bars", n=n, w=640, h=360)
    rings = g.StreamRings(batch, n)
    for _ in range(10):
        rings.advance()
    with g.Renderer(p, batch=batch) as a, g.Renderer(p, batch=batch) as b:
        a.update(rings.lb, rings.rb, True)
        dl, dr = torch.from_numpy(rings.lb).cuda(), torch.from_numpy(rings.rb).cuda()
        torch.cuda.synchronize()
        b.update_device(dl.data_ptr(), dr.data_ptr(), True)
        for s in (0, 7):
            assert np.array_equal(a.readback(s), b.readback(s))
        assert a.launch_count >= 2 and b.launch_count >= 2


def test_async_readback_snapshots_each_frame_before_the_next_update_overwrites_it(built):
    """readback_async copies out on its own stream while the next update's kernels run: every step's frame must still be
    the frame of THAT step (snapshot on the raster stream), for more read-backs in flight than staging buffers"""
    n, batch, steps = 1024, 2, 7
    p = g.default_params("bars", n=n, w=640, h=360, lazy_smooth=1)
    rings_a, rings_b = g.StreamRings(batch, n), g.StreamRings(batch, n)
    bufs = [g.pinned_empty((360, 640, 4), np.uint8) for _ in range(steps)]
    want = []
    with g.Renderer(p, batch=batch) as a, g.Renderer(p, batch=batch) as b:
        for i in range(steps):
            rings_a.advance(); rings_b.advance()
            a.update(rings_a.lb, rings_a.rb, True)
            a.readback_async(i % batch, bufs[i])
            b.update(rings_b.lb, rings_b.rb, True)
            want.append(b.readback(i % batch).copy())
        a.readback_fence()
        a.sync()
        for i in range(steps):
            assert np.array_equal(bufs[i], want[i]), i
        assert not np.array_equal(want[0], want[2])                        # the frames do differ from step to step


def _cudart():
    import ctypes, glob, os, torch
    for pat in ("libcudart.so", os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"), "/usr/local/cuda/lib64/libcudart.so*"):
        for cand in ([pat] if "*" not in pat else sorted(glob.glob(pat))):
            try:
                return ctypes.CDLL(cand)
            except OSError:
                continue
    pytest.skip("libcudart not loadable from Python")


@pytest.mark.gpu
def test_update_device_is_ordered_by_the_callers_events(built):
    """glava_b200_update_device_after / glava_b200_input_event: a producer that fills the PCM buffers on its OWN stream and
    overwrites them right after the call — without the two events the spectrum kernel (internal non-blocking stream) would
    race with it; with them every update sees exactly the buffer content of its turn"""
    import ctypes
    import torch
    n, batch, steps = 2048, 16, 12
    p = g.default_params("bars", n=n, w=128, h=64)
    rng = np.random.default_rng(77)
    seq = [((rng.random((batch, n), np.float32) - 0.5) * 0.4, (rng.random((batch, n), np.float32) - 0.5) * 0.4) for _ in range(steps)]
    prod = torch.cuda.Stream()
    dl = torch.zeros(batch, n, device="cuda"); dr = torch.zeros(batch, n, device="cuda")
    host = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in seq]
    with g.Renderer(p, batch=batch) as r, g.Renderer(p, batch=batch) as ref:
        for k in range(steps):
            with torch.cuda.stream(prod):
                # (torch has no wrapper for a foreign cudaEvent_t: wait on it through the runtime)
                rt = _cudart()
                rt.cudaStreamWaitEvent(ctypes.c_void_p(prod.cuda_stream), ctypes.c_void_p(r.input_event), 0)
                dl.copy_(host[k][0], non_blocking=True); dr.copy_(host[k][1], non_blocking=True)
                # burn some time on the producer stream so that an unordered consumer would read half-written buffers
                for _ in range(4):
                    dl.mul_(1.0); dr.mul_(1.0)
                ready = torch.cuda.Event(); ready.record(prod)
            r.update_device_after(dl.data_ptr(), dr.data_ptr(), True, ready.cuda_event)
            ref.update(seq[k][0], seq[k][1], True)
        r.sync(); ref.sync(); torch.cuda.synchronize()
        a, b = r.textures(), ref.textures()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].any()
