"""Spectrum kernel at the big buffer sizes (global tap table path), whole step next to it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g
for mod, n, w, h, batch in (("bars", 8192, 1920, 1080, 1024), ("bars", 16384, 1920, 1080, 1024), ("bars", 8192, 1280, 720, 1024), ("graph", 4096, 1920, 1080, 1024)):
    p = g.default_params(mod, n=n, w=w, h=h, lazy_smooth=1)
    r = g.Renderer(p, batch=batch)
    st = torch.cuda.ExternalStream(r.cuda_stream)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    for _ in range(4): r.update_device(x.data_ptr(), x.data_ptr(), True)
    r.sync(); r.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20): r.update_device(x.data_ptr(), x.data_ptr(), True)
    e1.record(st)
    kt = r.kernel_times()
    print(f"{mod} n={n} {w}x{h}: step {e0.elapsed_time(e1)/20:.3f} ms  spectrum(co-run) {kt['spectrum_ms']/kt['spectrum_launches']*1e3:.0f} us  raster(co-run) {kt['raster_ms']/kt['raster_launches']:.3f} ms", flush=True)
    r.close()
