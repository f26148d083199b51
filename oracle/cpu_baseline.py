"""TEST INFRASTRUCTURE / bench reference arm — the reference's own CPU path timed on the host cores.

kind "reference" (default wherever oracle/_ref/libglava_ref_gl.so and the Nsight Compute Mesa load — this image, i.e. also
the GPU box): every worker process IS a GLava renderer.  rd_new reads the benchmark's configuration, and one frame is one
rd_update(modified = true): transform_fft on the CPU, the pass / gravity / average / smooth fragment shaders and the
module's stages on Mesa llvmpipe — "CPU-FFT + software-GL (llvmpipe)", the baseline BASELINE.json's north_star names.
Nothing on that path is code of this repository (oracle/ref_gl.py, oracle/ref_shim.c only host the reference).

kind "port" (fallback, e.g. a box without that Mesa): reference transform_fft where oracle/_ref has it + the C
restatement of the GL passes and the fragment shader (oracle/glava_oracle.c).

One worker per host core, pinned (sched_setaffinity), llvmpipe's own rasteriser threads off (LP_NUM_THREADS=0: the cores
are already taken by the other streams); a step = every worker renders `frames_per_worker` frames between two barriers;
frames/s = frames / wall time of the step.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _rc_for(pdict):
    """the shipped rc.glsl with the benchmark's module / buffer size / geometry appended (later requests win)"""
    from oracle import ref_gl
    with open(os.path.join(ref_gl.shader_dir(), "rc.glsl")) as f:
        rc = f.read()
    return rc + "\n#request mod %s\n#request setbufsize %d\n#request setgeometry 0 0 %d %d\n" % (
        pdict["module"], pdict["n"], pdict["w"], pdict["h"])


def _worker(wid, core, pdict, hop, frames_per_worker, kind, start, done, stop, out):
    try:
        os.sched_setaffinity(0, {core})
    except OSError:
        pass
    os.environ["LP_NUM_THREADS"] = "0"
    dn = os.open(os.devnull, os.O_WRONLY)
    os.dup2(dn, 1)                                                # rd_new prints deprecation warnings to stdout (C stdio):
    os.close(dn)                                                  # the bench's stdout carries one JSON line only
    from glava_b200.synth import StreamRings
    rings = StreamRings(1, pdict["n"], hop=hop, first_stream=wid)
    for _ in range(pdict["n"] // hop):
        rings.advance()
    if kind == "reference":
        from oracle import ref_gl
        r = ref_gl.ReferenceGL(rc=_rc_for(pdict))
        assert (r.w, r.hh, r.bufsize) == (pdict["w"], pdict["h"], pdict["n"])

        def frame():
            r.frame(rings.lb[0], rings.rb[0], want_frame=False)
            r.finish()
    else:
        from oracle.oracle import Oracle, OracleChannel, Reference
        o = Oracle("libm")
        p = o.default_params(pdict["module"], n=pdict["n"], w=pdict["w"], h=pdict["h"])
        ref = Reference() if Reference.available() else None
        chans = [OracleChannel(o, p), OracleChannel(o, p)]
        rch = [ref.chan(p), ref.chan(p)] if ref else None
        is_fft = p.module != 4

        def frame():
            texs = []
            for ch, pcm in enumerate((rings.lb[0], rings.rb[0])):
                if not is_fft:
                    texs.append(chans[0].update(pcm, is_fft=0)[1] if ch == 0 else texs[0])
                elif ref is not None:
                    texs.append(chans[ch].update(ref.fft(rch[ch], pcm), is_fft=2)[1])
                else:
                    texs.append(chans[ch].update(pcm, is_fft=1)[1])
            o.raster(p, texs[0], texs[1])
    frame()                                                       # first frame: shader JIT, allocations
    while True:
        start.wait()
        if stop.value:
            break
        t0 = time.perf_counter()
        for _ in range(frames_per_worker):
            rings.advance()
            frame()
        out[wid] = time.perf_counter() - t0
        done.wait()
    if kind == "reference":
        r.close()


class CpuBaseline:
    """persistent pinned workers; step() = every worker renders `frames_per_worker` frames"""

    def __init__(self, pdict, cores=None, frames_per_worker=1, hop=256, kind=None):
        avail = sorted(os.sched_getaffinity(0))
        self.cores = min(cores or len(avail), len(avail))
        self.fpw = frames_per_worker
        if kind is None:
            from oracle import ref_gl
            kind = "reference" if ref_gl.available() else "port"
        self.kind = kind
        if kind == "reference":
            from oracle import ref_gl
            ref_gl.shader_dir()                                   # unpack once, before the workers fork
        ctx = mp.get_context("fork")
        self.start = ctx.Barrier(self.cores + 1); self.done = ctx.Barrier(self.cores + 1)
        self.stop = ctx.Value("i", 0); self.out = ctx.Array("d", self.cores)
        self.procs = [ctx.Process(target=_worker, daemon=True,
                                  args=(i, avail[i], pdict, hop, frames_per_worker, kind, self.start, self.done, self.stop, self.out))
                      for i in range(self.cores)]
        for p in self.procs:
            p.start()

    def step(self):
        """returns (wall seconds, frames rendered)"""
        self.start.wait(timeout=600)
        t0 = time.perf_counter()
        self.done.wait(timeout=600)
        return time.perf_counter() - t0, self.cores * self.fpw

    def close(self):
        self.stop.value = 1
        try:
            self.start.wait(timeout=60)
        except Exception:
            pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()


def describe(kind):
    if kind == "reference":
        return ("the reference itself: rd_update = transform_fft on the CPU + its GL passes and module shaders on Mesa llvmpipe "
                "(oracle/_ref/libglava_ref_gl.so), one GLava renderer per pinned host core, LP_NUM_THREADS=0")
    return ("reference render.c transform_fft where oracle/_ref has it + C restatement of the GL passes K1-K5 and the module "
            "fragment shader (no loadable OpenGL on this box), one process per pinned host core")
