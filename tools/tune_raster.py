"""Sweep rows-per-CTA / block width of the bars raster kernel (development aid)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os; sys.path.insert(0, %r)
import torch, glava_b200 as g
p = g.default_params("bars", n=4096, w=1920, h=1080, lazy_smooth=1)
r = g.Renderer(p, batch=1024)
x = (torch.rand(1024, 4096, device="cuda") - 0.5) * 0.2
for _ in range(3): r.update_device(x.data_ptr(), x.data_ptr(), True)
r.sync(); r.set_timing(True)
for _ in range(10): r.update_device(x.data_ptr(), x.data_ptr(), False)
kt = r.kernel_times(); ms = kt["raster_ms"] / kt["raster_launches"]
print("%%.4f ms %%.0f GB/s" %% (ms, 1024*1920*1080*4/ms/1e6))
''' % ROOT
for rows in (30, 54, 90, 135, 270, 540, 1080):
    for bx in (96, 160, 256):
        env = dict(os.environ, GLAVA_B200_ROWS=str(rows), GLAVA_B200_BX=str(bx))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"rows={rows:5d} bx={bx:4d}: {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
