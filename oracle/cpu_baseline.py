"""TEST INFRASTRUCTURE / bench reference arm — the reference's CPU path timed on host cores.

One frame = what one rd_update(modified=true) does for one stream under the shipped config:
  * transform_fft per channel — the reference's OWN compiled code (oracle/_ref) when present,
    otherwise the bit-identical C restatement;
  * the GL 1-D passes K1-K5 and the module fragment shader — our C restatement of the GLSL
    (no OpenGL / llvmpipe exists in this image), gcc -O2, no hand SIMD.
Streams are spread over worker processes, one per host core; frames/s are summed.
"""
import multiprocessing as mp
import os
import time

import numpy as np

_state = {}


def _worker_init(pdict, first_stream, streams_per_worker, hop, use_ref):
    import ctypes as C
    from oracle.oracle import Oracle, OracleChannel, OrcParams, Reference
    o = Oracle("libm")
    p = o.default_params(pdict["module"], n=pdict["n"], w=pdict["w"], h=pdict["h"])
    for k, v in pdict.items():
        if k not in ("module",):
            setattr(p, k, v)
    ident = mp.current_process()._identity
    wid = (ident[0] - 1) if ident else 0
    from glava_b200.synth import StreamRings
    rings = StreamRings(streams_per_worker, p.n, hop=hop, first_stream=first_stream + wid * streams_per_worker)
    for _ in range(p.n // hop):              # fill the rings before timing
        rings.advance()
    ref = Reference() if (use_ref and Reference.available()) else None
    is_fft = p.module != 4
    _state.update(o=o, p=p, rings=rings, ref=ref, is_fft=is_fft,
                  chans=[[OracleChannel(o, p), OracleChannel(o, p)] for _ in range(streams_per_worker)],
                  rchans=[[ref.chan(p), ref.chan(p)] for _ in range(streams_per_worker)] if ref else None,
                  img=np.zeros((p.h, p.w, 4), np.uint8))


def _worker_step(_):
    st = _state
    o, p, rings = st["o"], st["p"], st["rings"]
    rings.advance()                           # input production is outside the timed part
    t0 = time.perf_counter()
    for s in range(rings.batch):
        texs = []
        for ch, pcm in enumerate((rings.lb[s], rings.rb[s])):
            if not st["is_fft"]:
                if ch == 1:
                    texs.append(texs[0]); continue
                texs.append(st["chans"][s][ch].update(pcm, is_fft=0)[1])
            elif st["ref"] is not None:
                f = st["ref"].fft(st["rchans"][s][ch], pcm)                  # reference's own transform_fft
                texs.append(st["chans"][s][ch].update(f, is_fft=2)[1])
            else:
                texs.append(st["chans"][s][ch].update(pcm, is_fft=1)[1])
        img = o.raster(p, texs[0], texs[1])
    return time.perf_counter() - t0, int(img[..., 3].sum() & 0xFFFF)


class CpuBaseline:
    """persistent worker pool; step() = every worker renders `streams_per_worker` frames"""

    def __init__(self, pdict, cores=None, streams_per_worker=1, hop=256, use_ref=True):
        self.cores = cores or len(os.sched_getaffinity(0))
        self.spw = streams_per_worker
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.cores, initializer=_worker_init, initargs=(pdict, 0, streams_per_worker, hop, use_ref))
        from oracle.oracle import Reference
        self.kind = "reference" if (use_ref and Reference.available()) else "port"

    def step(self):
        """returns (wall seconds, frames rendered)"""
        t0 = time.perf_counter()
        self.pool.map(_worker_step, range(self.cores), chunksize=1)
        return time.perf_counter() - t0, self.cores * self.spw

    def close(self):
        self.pool.close(); self.pool.join()


def describe(kind):
    fft = "reference render.c transform_fft (oracle/_ref)" if kind == "reference" else "C restatement of transform_fft"
    return (f"{fft} + C restatement of the GL passes K1-K5 and the module fragment shader "
            "(no OpenGL/llvmpipe in this image), one process per host core")
