"""Audio side of the boundary (SURVEY §8 row a1, §8b "Audio plug-in ABI"), CPU tier, over real named pipes:

* the reference's OWN FIFO backend (glava/fifo.c compiled where it lies -> oracle/_ref/libglava_ref_fifo.so) plugs into
  the batch feeder through the unchanged `struct audio_impl`, and its rings pin the oracle's orc_fifo_ingest;
* the native "fifo" backend produces the same rings from the same bytes;
* the batched FIFO gather (chunks for glava_b200_ingest_fifo) returns the written chunks, zero chunks for silent
  streams, keeps partial chunks queued, and follows the producer's cadence with its deadline."""
import ctypes as C
import fcntl
import os
import termios
import time

import numpy as np
import pytest

import glava_b200 as g
from glava_b200 import audio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FIFO = os.path.join(ROOT, "oracle", "_ref", "libglava_ref_fifo.so")

pytestmark = pytest.mark.timeout(120)
_keep = []


@pytest.fixture(scope="module")
def ref_backend(built):
    """register the reference's fifo.c backend under the name "fifo_ref" (its own name "fifo" is taken by the native one)"""
    if not os.path.exists(REF_FIFO):
        pytest.skip("oracle/_ref/libglava_ref_fifo.so not built (reference tree absent and no prebuilt copy)")
    L = C.CDLL(REF_FIFO)
    assert C.c_size_t.in_dll(L, "audio_impls_idx").value == 1            # AUDIO_ATTACH(fifo) ran (fifo.c:129)
    impl = C.cast((C.c_void_p * 4).in_dll(L, "audio_impls")[0], C.POINTER(audio.AudioImpl)).contents
    assert impl.name == b"fifo"
    mine = audio.AudioImpl(b"fifo_ref", impl.init, impl.entry)
    _keep.extend([L, mine])
    audio.register_backend(C.addressof(mine))
    return "fifo_ref"


def _pipes(tmp_path, k):
    paths = [str(tmp_path / f"s{i}.fifo") for i in range(k)]
    for p in paths:
        os.mkfifo(p)
    return paths


def _unread(fd):
    buf = C.c_int(0)
    fcntl.ioctl(fd, termios.FIONREAD, buf)
    return buf.value


def _run_backend(backend, tmp_path, data, n, samplesz, channels):
    """feed data[s] ([chunks][samplesz / 2] int16) into stream s's FIFO in one write; -> rings after everything was consumed"""
    batch = len(data)
    paths = _pipes(tmp_path, batch)
    lb = np.zeros((batch, n), np.float32); rb = np.zeros_like(lb)
    with audio.AudioBatch(backend, paths, batch, n, samplesz=samplesz, channels=channels) as ab:
        fds = [os.open(p, os.O_WRONLY) for p in paths]                   # returns once the backend thread has opened its end
        try:
            for fd, d in zip(fds, data):
                assert os.write(fd, d.tobytes()) == d.nbytes
            deadline = time.time() + 20
            while any(_unread(fd) for fd in fds):
                assert time.time() < deadline, "backend did not drain its FIFO"
                time.sleep(0.002)
            time.sleep(0.01)                                             # the last chunk is being pushed into the ring
            modified = ab.collect(lb, rb)
            assert modified.all()
            st = ab.stream(0)
            assert st.audio_buf_sz == n and st.sample_sz == samplesz and st.channels == channels and st.format == -1
            assert st.source == paths[0].encode()
        finally:
            for fd in fds:
                os.close(fd)
    return lb, rb


def _expected(orc, chunks, ring_l, n, hop, channels):
    """oracle ring after `chunks`, then as many zero chunks as the backend slid in while the FIFO stayed silent"""
    tail = np.flatnonzero(ring_l)
    assert tail.size, "ring is empty: the backend zero-slid everything out (machine too slow?)"
    k = (n - 1 - tail[-1]) // hop
    el = np.zeros(n, np.float32); er = np.zeros(n, np.float32)
    for c in chunks:
        orc.fifo_ingest(el, er, c, channels)
    zero = np.zeros(hop * 2, np.int16)
    for _ in range(k):
        orc.fifo_ingest(el, er, zero, channels)
    return el, er, k


def _data(batch, chunks, hop, seed):
    rng = np.random.default_rng(seed)
    d = rng.integers(200, 32767, size=(batch, chunks, hop * 2)).astype(np.int16)
    d *= rng.choice(np.array([-1, 1], np.int16), size=(batch, chunks, 1))   # one sign per chunk: no sample and no L/R mean is 0
    return d


@pytest.mark.parametrize("channels", [2, 1])
def test_reference_fifo_backend_plugs_in_and_pins_the_oracle(ref_backend, orc, tmp_path, channels):
    n, samplesz, batch, chunks = 16384, 256, 3, 40
    hop = samplesz // 4
    data = _data(batch, chunks, hop, 11 + channels)
    lb, rb = _run_backend(ref_backend, tmp_path, data, n, samplesz, channels)
    for s in range(batch):
        el, er, k = _expected(orc, data[s], lb[s], n, hop, channels)
        assert chunks + k < n // hop, "silence slid data out of the ring before the snapshot"
        assert np.array_equal(lb[s], el) and np.array_equal(rb[s], er), (s, k)
        if channels == 1:
            assert np.array_equal(lb[s], rb[s])


@pytest.mark.parametrize("channels", [2, 1])
def test_native_fifo_backend_matches_the_oracle(built, orc, tmp_path, channels):
    n, samplesz, batch, chunks = 16384, 256, 4, 40
    hop = samplesz // 4
    data = _data(batch, chunks, hop, 23 + channels)
    lb, rb = _run_backend("fifo", tmp_path, data, n, samplesz, channels)
    for s in range(batch):
        el, er, k = _expected(orc, data[s], lb[s], n, hop, channels)
        assert chunks + k < n // hop
        assert np.array_equal(lb[s], el) and np.array_equal(rb[s], er), (s, k)


def test_unknown_backend_and_missing_source_fail_with_the_reference_messages(built, tmp_path):
    with pytest.raises(g.GlavaError, match=r'The specified audio backend \("pulseaudio"\) is not available\.'):
        audio.find_backend("pulseaudio")
    with pytest.raises(g.GlavaError, match="is not available"):
        audio.AudioBatch("nope", None, 1, 1024)
    with pytest.raises(g.GlavaError, match=r'failed to open FIFO audio source ".*absent\.fifo": No such file'):
        audio.FifoReader([str(tmp_path / "absent.fifo")])
    with pytest.raises(g.GlavaError, match="bad arguments"):
        audio.AudioBatch("fifo", None, 1, 64, samplesz=1024)             # a hop longer than the ring
    assert audio.find_backend("fifo")


def test_fifo_gather_chunks_silence_partial_and_cadence(built, orc, tmp_path):
    batch, samplesz = 3, 1024
    hop = samplesz // 4
    paths = _pipes(tmp_path, batch)
    rng = np.random.default_rng(5)
    with audio.FifoReader(paths, samplesz) as fr:
        assert fr.timeout_ms == 50                                       # fifo.c:40
        # no writer anywhere: one tick of silence, after the full deadline
        t0 = time.time()
        chunks, fresh = fr.gather()
        assert 0.045 <= time.time() - t0 < 1.0
        assert not fresh.any() and not chunks.any() and chunks.shape == (batch, hop * 2)
        fds = [os.open(p, os.O_WRONLY) for p in paths]
        try:
            sent = rng.integers(-32768, 32767, size=(4, batch, hop * 2), dtype=np.int16)
            # tick 1: streams 0 and 2 deliver, stream 1 stays silent -> zeros for it
            os.write(fds[0], sent[0, 0].tobytes()); os.write(fds[2], sent[0, 2].tobytes())
            chunks, fresh = fr.gather()
            assert fresh.tolist() == [True, False, True]
            assert np.array_equal(chunks[0], sent[0, 0]) and np.array_equal(chunks[2], sent[0, 2]) and not chunks[1].any()
            # tick 2: everyone delivers, stream 1 in two pieces written before the tick -> returns without waiting
            os.write(fds[0], sent[1, 0].tobytes()); os.write(fds[2], sent[1, 2].tobytes())
            raw = sent[1, 1].tobytes()
            os.write(fds[1], raw[:300]); os.write(fds[1], raw[300:])
            t0 = time.time()
            chunks, fresh = fr.gather()
            assert time.time() - t0 < 0.045 and fresh.all() and np.array_equal(chunks, sent[1])
            assert 1 <= fr.timeout_ms <= 500                             # deadline now follows the producer (fifo.c:82-87)
            # tick 3: half a chunk only -> silence this tick, the bytes stay queued ...
            raw = sent[2, 1].tobytes()
            os.write(fds[1], raw[:1000])
            chunks, fresh = fr.gather()
            assert not fresh.any() and not chunks.any()
            # ... and complete the chunk in the next one; two chunks queued on stream 0 come out one per tick, in order
            os.write(fds[1], raw[1000:])
            os.write(fds[0], sent[2, 0].tobytes() + sent[3, 0].tobytes())
            chunks, fresh = fr.gather()
            assert fresh.tolist() == [True, True, False]
            assert np.array_equal(chunks[1], sent[2, 1]) and np.array_equal(chunks[0], sent[2, 0])
            chunks, fresh = fr.gather()
            assert fresh.tolist() == [True, False, False] and np.array_equal(chunks[0], sent[3, 0])
        finally:
            for fd in fds:
                os.close(fd)
        # writers gone (POLLHUP): still just silence, no spin, no error
        chunks, fresh = fr.gather()
        assert not fresh.any() and not chunks.any()


def test_gathered_chunks_through_the_ingest_equal_a_backend_thread_on_the_same_bytes(built, orc, tmp_path):
    """device-ring path == host-ring path: chunks from the batched gather, pushed through the (pinned) ingest arithmetic,
    give the rings the per-stream backend thread builds from the same FIFO bytes"""
    n, samplesz, batch, nch = 8192, 512, 2, 12
    hop = samplesz // 4
    data = _data(batch, nch, hop, 77)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    lb, rb = _run_backend("fifo", tmp_path / "a", data, n, samplesz, 2)
    paths = _pipes(tmp_path / "b", batch)
    rl = np.zeros((batch, n), np.float32); rr = np.zeros_like(rl)
    with audio.FifoReader(paths, samplesz) as fr:
        fds = [os.open(p, os.O_WRONLY) for p in paths]
        try:
            for fd, d in zip(fds, data):
                os.write(fd, d.tobytes())
            for _ in range(nch):
                chunks, fresh = fr.gather()
                assert fresh.all()
                for s in range(batch):
                    orc.fifo_ingest(rl[s], rr[s], chunks[s], 2)
        finally:
            for fd in fds:
                os.close(fd)
    zero = np.zeros(hop * 2, np.int16)
    for s in range(batch):
        k = (n - 1 - np.flatnonzero(lb[s])[-1]) // hop                   # silence the backend thread slid in afterwards
        for _ in range(k):
            orc.fifo_ingest(rl[s], rr[s], zero, 2)
        assert np.array_equal(rl[s], lb[s]) and np.array_equal(rr[s], rb[s])


class PyBackend:
    """a backend written against the plug-in ABI in Python (ctypes callbacks): each stream's entry thread publishes
    `rings[k]` on demand — deterministic stand-in for an audio thread in the feeder tests (CPU and GPU tier)"""

    def __init__(self, name, n):
        self.n, self.pending, self.streams = n, {}, {}
        self._init = C.CFUNCTYPE(None, C.POINTER(audio.AudioData))(self.init)
        self._entry = C.CFUNCTYPE(C.c_void_p, C.c_void_p)(self.entry)
        self.impl = audio.AudioImpl(name.encode(), C.cast(self._init, C.c_void_p), C.cast(self._entry, C.c_void_p))
        audio.register_backend(C.addressof(self.impl))

    def init(self, d):
        if not d.contents.source:
            d.contents.source = b"stream-0"                              # a backend default, like "/tmp/mpd.fifo"

    def entry(self, p):
        d = C.cast(p, C.POINTER(audio.AudioData)).contents
        sid = int(d.source.decode().split("-")[1])
        self.streams[sid] = d
        while not d.terminate:
            job = self.pending.pop(sid, None)
            if job is not None:
                l, r = job
                C.memmove(d.audio_out_l, l.ctypes.data, self.n * 4); C.memmove(d.audio_out_r, r.ctypes.data, self.n * 4)
                d.modified = True
            time.sleep(0.001)
        return None

    def publish(self, sid, l, r):
        self.pending[sid] = (np.ascontiguousarray(l, np.float32), np.ascontiguousarray(r, np.float32))
        while sid in self.pending or not self.streams[sid].modified:
            time.sleep(0.001)


def test_batch_feeder_collect_copies_only_modified_streams(built):
    n, batch = 512, 3
    be = PyBackend("pytest_rings", n); _keep.append(be)
    rng = np.random.default_rng(9)
    with audio.AudioBatch("pytest_rings", [f"stream-{s}" for s in range(batch)], batch, n, samplesz=256) as ab:
        while len(be.streams) < batch:
            time.sleep(0.001)
        lb = np.full((batch, n), 7.0, np.float32); rb = np.full((batch, n), 8.0, np.float32)
        assert not ab.collect(lb, rb).any() and (lb == 7).all() and (rb == 8).all()      # nothing modified: rows untouched
        l1, r1 = rng.random(n, np.float32), rng.random(n, np.float32)
        be.publish(1, l1, r1)
        assert ab.collect(lb, rb).tolist() == [False, True, False]
        assert np.array_equal(lb[1], l1) and np.array_equal(rb[1], r1) and (lb[0] == 7).all() and (rb[2] == 8).all()
        assert not ab.collect(lb, rb).any()                                              # the flag was cleared (glava.c:535)
        assert ab.stream(2).rate == 22050 and ab.stream(2).source == b"stream-2"


def test_a_silent_stream_costs_one_deadline_then_stops_holding_the_batch_up(built, tmp_path):
    """every stream of the reference has its own thread and timeout; in the batched gather a stream that went silent is
    awaited for one deadline and then contributes zeros without delaying the others — and rejoins when it has a chunk"""
    batch, samplesz, ticks = 3, 512, 12
    paths = _pipes(tmp_path, batch)
    rng = np.random.default_rng(8)
    sent = rng.integers(-32768, 32767, size=(ticks, 2, samplesz // 2), dtype=np.int16)
    with audio.FifoReader(paths, samplesz) as fr:
        fds = [os.open(p, os.O_WRONLY) for p in paths]
        try:
            for s in range(2):
                os.write(fds[s], sent[:, s].tobytes())                   # streams 0 and 1: all their chunks queued; 2: nothing
            t0 = time.time()
            chunks, fresh = fr.gather()                                  # first tick: everybody is awaited -> one full deadline
            first = time.time() - t0
            assert fresh.tolist() == [True, True, False] and 0.045 <= first < 1.0
            for t in range(1, ticks - 1):
                t0 = time.time()
                chunks, fresh = fr.gather()
                assert time.time() - t0 < 0.045, t                       # no waiting for the silent stream any more
                assert fresh.tolist() == [True, True, False] and np.array_equal(chunks[:2], sent[t]) and not chunks[2].any()
            late = rng.integers(-32768, 32767, samplesz // 2, dtype=np.int16)
            os.write(fds[2], late.tobytes())                             # the silent stream comes back
            chunks, fresh = fr.gather()
            assert fresh.all() and np.array_equal(chunks[2], late) and np.array_equal(chunks[:2], sent[ticks - 1])
        finally:
            for fd in fds:
                os.close(fd)


def test_native_backend_without_a_writer_is_silence_and_stays_joinable(built, tmp_path):
    """the reference's thread sits in open() until a writer appears (and cannot be joined until then), and spins on a FIFO
    whose writer has exited; the native backend reads both situations as silence at the poll cadence"""
    paths = _pipes(tmp_path, 2)
    n = 1024
    lb = np.ones((2, n), np.float32); rb = np.ones_like(lb)
    t0 = time.time()
    with audio.AudioBatch("fifo", paths, 2, n, samplesz=256) as ab:
        time.sleep(0.13)                                                 # > 2 poll timeouts of 50 ms, nobody writes
        assert ab.collect(lb, rb).all() and not lb.any() and not rb.any()
        fd = os.open(paths[0], os.O_WRONLY)
        os.write(fd, np.full(128, 1000, np.int16).tobytes())
        deadline = time.time() + 5
        while not (ab.collect(lb, rb)[0] and lb[0].any()):
            assert time.time() < deadline
            time.sleep(0.002)
        os.close(fd)                                                     # writer gone: POLLHUP from now on
        time.sleep(0.05)
        c0 = time.process_time()
        time.sleep(0.2)
        assert time.process_time() - c0 < 0.15                           # no busy loop on the hung-up FIFO
    assert time.time() - t0 < 3.0                                        # stop() joined both threads promptly
