"""Spectrum-kernel probe (development aid): the kernel ALONE (tiny frame, so the raster is negligible) at the need-list of a
given frame width, per buffer size.

    python tools/spec_probe.py                     # table over n, pipelines, lazy / full K5
    python tools/spec_probe.py --one 8192          # a few updates at one size (profiling target for ncu -k regex:spectrum)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g


def run(mod, n, w, batch, lazy, accel, smooth=1, F=5, reps=10):
    p = g.default_params(mod, n=n, w=w, h=16, lazy_smooth=lazy, accel_fft=accel, smooth_pass=smooth, avg_frames=F)
    r = g.Renderer(p, batch=batch)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    y = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    for _ in range(3):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    r.sync(); r.set_timing(True)
    for _ in range(reps):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    kt = r.kernel_times()
    r.close()
    return kt["spectrum_ms"] / kt["spectrum_launches"] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", type=int, default=0)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--module", default="bars")
    ap.add_argument("--sizes", default="512,1024,2048,4096,8192,16384")
    ap.add_argument("--quick", action="store_true", help="lazy pipeline-B only")
    a = ap.parse_args()
    if a.one:
        print(f"n={a.one}: {run(a.module, a.one, a.width, a.batch, 1, 1, reps=4):.1f} us")
        return
    rows = []
    for n in [int(v) for v in a.sizes.split(",")]:
        for lazy, accel, smooth in (((1, 1, 1),) if a.quick else ((1, 1, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1))):
            us = run(a.module, n, a.width, a.batch, lazy, accel, smooth)
            pcm_mb = a.batch * 2 * n * 4 / 1e6
            rows.append(dict(n=n, lazy=lazy, accel=accel, smooth=smooth, us=us, pcm_gbs=pcm_mb / us * 1e3))
            print(f"n={n:5d} lazy={lazy} pipeline={'B' if accel else 'A'} K5={smooth}: {us:7.1f} us  (PCM alone would stream at {pcm_mb / us * 1e3:6.0f} GB/s)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/spec_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
