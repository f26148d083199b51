"""Generates tests/golden/*.npz from the REFERENCE'S OWN compiled code (oracle/_ref, built from
/root/reference/glava/render.c by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures (all seeded, small):
  spectrum_a_n{N}.npz   N in 512/1024/4096: 12 updates of one synthetic stereo stream through
                        transform_fft -> transform_gravity -> transform_average (render.c:2149-2156);
                        stores the int16 FIFO chunks and the left-channel result after updates 1, 6, 12,
                        plus the raw transform_fft output of update 12.
  fft_kat.npz           transform_fft on a 64-cycle sine and on an impulse, N = 1024
  wrange.npz            transform_wrange on a ramp
  smooth.npz            transform_smooth (render.c:694-718) on a seeded spectrum-like buffer with zeros, N = 1024 and
                        4096, default (0.01, 4) and a wide (0.2, 2) setting
  colors.npz            ext_parse_color (glsl_ext.c:88-122) on the colour literals the shipped modules use
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.oracle import Reference, Oracle  # noqa: E402
from glava_b200.synth import synth_pcm_int16, fifo_to_float  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert Reference.available(), "needs /root/reference to build oracle/_ref"
    ref = Reference()
    orc = Oracle("libm")
    hop = 256
    for n in (512, 1024, 4096):
        p = orc.default_params("bars", n=n, accel_fft=0)
        ch = ref.chan(p)
        ring = np.zeros(n, np.float32)
        chunks, outs, raw = [], {}, None
        for u in range(1, 13):
            c = synth_pcm_int16(7, (u - 1) * hop, hop)
            chunks.append(c)
            l, _ = fifo_to_float(c)
            ring = np.concatenate([ring[hop:], l])
            if u == 12:
                raw = ref.fft(ref.chan(p), ring)
            out = ref.update_a(ch, ring)
            if u in (1, 6, 12):
                outs[f"out_{u}"] = out
        np.savez_compressed(os.path.join(HERE, f"spectrum_a_n{n}.npz"), chunks=np.stack(chunks), raw_fft_12=raw,
                            ur=np.float32(p.ur), **outs)
    p = orc.default_params("bars", n=1024, accel_fft=0)
    i = np.arange(1024)
    sine = np.sin(2 * np.pi * 64 * i / 1024).astype(np.float32) * np.float32(0.25)
    imp = np.zeros(1024, np.float32); imp[10] = 0.5
    np.savez_compressed(os.path.join(HERE, "fft_kat.npz"), sine=sine, sine_out=ref.fft(ref.chan(p), sine),
                        impulse=imp, impulse_out=ref.fft(ref.chan(p), imp))
    ramp = np.linspace(-0.5, 0.5, 1024, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "wrange.npz"), ramp=ramp, out=ref.wrange(ramp))
    rng = np.random.default_rng(99)
    sm = {}
    for n in (1024, 4096):
        x = (rng.random(n) ** 3).astype(np.float32); x[rng.random(n) < 0.2] = 0
        sm[f"in_{n}"] = x
        sm[f"out_{n}_default"] = ref.smooth(x, 0.01, 4.0)
        sm[f"out_{n}_wide"] = ref.smooth(x, 0.2, 2.0)
    np.savez_compressed(os.path.join(HERE, "smooth.npz"), **sm)
    names = ["3366b2", "a0a0b2", "333333", "cc3333", "cca0a0", "802A2A", "4F4F92", "262626", "00000000", "55000055", "0xff8000"]
    vals = np.stack([ref.parse_color(s)[1] for s in names])
    np.savez_compressed(os.path.join(HERE, "colors.npz"), names=np.array(names), rgba=vals)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
