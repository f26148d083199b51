"""-m gpu: the two conveniences of include/glava_b200_audio.h that touch the device.
glava_b200_fifo_pump (gather -> ingest kernel -> update on the resident rings) and glava_b200_audio_frame (locked
collect -> glava_b200_update) must give exactly what glava_b200_update gives on host rings built by the oracle's
fifo.c restatement from the same bytes."""
import os
import time

import numpy as np
import pytest

import glava_b200 as g
from glava_b200 import audio
from tests.test_audio_fifo import PyBackend, _keep, _pipes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_fifo_pump_equals_update_on_oracle_rings(orc_pm, built, tmp_path):
    n, batch, samplesz, ticks = 1024, 3, 1024, 7
    hop = samplesz // 4
    p = g.default_params("bars", n=n, w=64, h=16)
    paths = _pipes(tmp_path, batch)
    rng = np.random.default_rng(21)
    sent = rng.integers(-32768, 32767, size=(ticks, batch, hop * 2), dtype=np.int16)
    rl = np.zeros((batch, n), np.float32); rr = np.zeros_like(rl)
    with audio.FifoReader(paths, samplesz) as fr, g.Renderer(p, batch=batch) as r, g.Renderer(p, batch=batch) as r2:
        fds = [os.open(q, os.O_WRONLY) for q in paths]
        try:
            for t in range(ticks):
                for s in range(batch):
                    if not (t == 3 and s == 1):                          # stream 1 is silent for one tick: zeros slide in
                        os.write(fds[s], sent[t, s].tobytes())
                    else:
                        sent[t, s] = 0
                fr.pump(r)
                for s in range(batch):
                    orc_pm.fifo_ingest(rl[s], rr[s], sent[t, s], 2)
                r2.update(rl, rr, True)
        finally:
            for fd in fds:
                os.close(fd)
        r.sync(); r2.sync()
        a, b = r.spectrum(), r2.spectrum()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for s in range(batch):
            assert np.array_equal(r.readback(s), r2.readback(s))
        assert a[0].any()


def test_audio_frame_equals_update_with_the_collected_rings(built):
    n, batch = 1024, 2
    p = g.default_params("graph", n=n, w=64, h=32)
    be = PyBackend("pytest_rings_gpu", n); _keep.append(be)
    rng = np.random.default_rng(4)
    lb = np.zeros((batch, n), np.float32); rb = np.zeros_like(lb)
    with audio.AudioBatch("pytest_rings_gpu", [f"stream-{s}" for s in range(batch)], batch, n) as ab, \
            g.Renderer(p, batch=batch) as r, g.Renderer(p, batch=batch) as r2:
        while len(be.streams) < batch:
            time.sleep(0.001)
        for step in range(7):
            mask = np.zeros(batch, bool)
            for s in range(batch):
                if step != 2 and (step + s) % 3 != 1:                     # step 2: nobody publishes -> modified = 0 re-raster
                    lb[s] = (rng.random(n, np.float32) - 0.5) * 0.2; rb[s] = (rng.random(n, np.float32) - 0.5) * 0.2
                    be.publish(s, lb[s], rb[s]); mask[s] = True
            ab.frame(r)
            r2.update_masked(lb, rb, mask)                                # glava.c:528-537 per stream: only those that ticked run the chain
        r.sync(); r2.sync()
        a, b = r.spectrum(), r2.spectrum()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].any()
        for s in range(batch):
            assert np.array_equal(r.readback(s), r2.readback(s))


def test_float_ingest_follows_the_pulseaudio_ring_update(built):
    """glava_b200_ingest_float = pulse_input.c:146-174: the samples are already float (no / 65535), slide by `frames`,
    mono = (l + r) / 2 in float; the rings it leaves in HBM must render exactly what the same rings handed to
    glava_b200_update render"""
    n, batch, frames = 1024, 3, 256
    rng = np.random.default_rng(17)
    for channels in (2, 1):
        p = g.default_params("bars", n=n, w=64, h=16, channels=channels)
        rl = np.zeros((batch, n), np.float32); rr = np.zeros_like(rl)
        with g.Renderer(p, batch=batch) as r, g.Renderer(p, batch=batch) as r2:
            for _ in range(6):
                chunk = ((rng.random((batch, frames * 2), np.float32) - 0.5) * 0.8).astype(np.float32)
                r.ingest_float(chunk); r.update_rings(True)
                a, b = chunk[:, 0::2], chunk[:, 1::2]
                rl[:, :-frames] = rl[:, frames:]; rr[:, :-frames] = rr[:, frames:]
                if channels == 1:
                    m = ((a + b) / np.float32(2)).astype(np.float32)
                    rl[:, -frames:] = m; rr[:, -frames:] = m
                else:
                    rl[:, -frames:] = a; rr[:, -frames:] = b
                r2.update(rl, rr, True)
            r.sync(); r2.sync()
            x, y = r.spectrum(), r2.spectrum()
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[0].any()
            for s in range(batch):
                assert np.array_equal(r.readback(s), r2.readback(s))
