"""Quick GPU probe: rough kernel timings per module (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g

cfgs = [("bars", 4096, 1920, 1080, 1024, 0), ("bars", 4096, 1920, 1080, 1024, 1), ("radial", 8192, 3840, 2160, 64, 1),
        ("radial", 4096, 1920, 1080, 256, 1), ("circle", 4096, 1920, 1080, 256, 1), ("graph", 2048, 1280, 720, 256, 1),
        ("wave", 2048, 1280, 720, 256, 1), ("bars", 16384, 1280, 720, 512, 1), ("bars", 512, 1280, 720, 512, 1)]
for mod, n, w, h, batch, lazy in cfgs:
    p = g.default_params(mod, n=n, w=w, h=h, lazy_smooth=lazy)
    r = g.Renderer(p, batch=batch)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    y = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    torch.cuda.synchronize()
    for _ in range(3):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    r.sync()
    r.set_timing(True)
    for _ in range(5):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    kt = r.kernel_times()
    ras = kt["raster_ms"] / kt["raster_launches"]; spec = kt["spectrum_ms"] / kt["spectrum_launches"]
    gb = batch * w * h * 4 / 1e9
    print(f"{mod:7s} n={n:5d} {w}x{h} batch={batch} lazy={lazy}: raster {ras:.3f} ms = {gb/ras*1e3:6.0f} GB/s ({gb/ras*1e3/6569.6:.2f}), "
          f"spectrum {spec:.3f} ms, step fps {batch/(ras+spec)*1e3:.0f}", flush=True)
    r.close()
