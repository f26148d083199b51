"""Quick GPU probe: smoke + rough kernel timings (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import glava_b200 as g
import __graft_entry__ as ge

ge.smoke()
for mod, n, w, h, batch in (("bars", 4096, 1920, 1080, 256), ("bars", 4096, 1920, 1080, 1024), ("radial", 4096, 1920, 1080, 64),
                            ("circle", 4096, 1920, 1080, 64), ("graph", 2048, 1280, 720, 256), ("wave", 2048, 1280, 720, 256)):
    p = g.default_params(mod, n=n, w=w, h=h)
    r = g.Renderer(p, batch=batch)
    st = torch.cuda.ExternalStream(r.cuda_stream)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    y = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    torch.cuda.synchronize()
    for _ in range(3):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    r.sync()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(st)
    for _ in range(5):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    e[1].record(st)
    for _ in range(5):
        r.update_device(x.data_ptr(), y.data_ptr(), False)   # raster only
    e[2].record(st)
    r.sync()
    full = e[0].elapsed_time(e[1]) / 5; ras = e[1].elapsed_time(e[2]) / 5
    gb = batch * w * h * 4 / 1e9
    print(f"{mod} n={n} {w}x{h} batch={batch}: step {full:.3f} ms ({batch/full*1e3:.0f} fps), raster {ras:.3f} ms = {gb/ras*1e3:.0f} GB/s, spectrum ~{full-ras:.3f} ms", flush=True)
    r.close()
