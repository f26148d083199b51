"""Spectrum-kernel breakdown probe at the headline's need-list (w=1920): with / without K5, L2-warm (tiny raster)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g
for n, batch, lazy, smooth, F in ((4096, 1024, 1, 1, 5), (4096, 1024, 1, 0, 5), (4096, 1024, 0, 0, 5), (4096, 1024, 1, 1, 1), (4096, 1024, 1, 0, 1),
                                  (2048, 1024, 1, 1, 5), (8192, 1024, 1, 1, 5), (8192, 1024, 1, 0, 5), (16384, 512, 1, 1, 5), (16384, 512, 1, 0, 5)):
    p = g.default_params("bars", n=n, w=1920, h=16, lazy_smooth=lazy, smooth_pass=smooth, avg_frames=F)
    r = g.Renderer(p, batch=batch)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    for _ in range(3): r.update_device(x.data_ptr(), x.data_ptr(), True)
    r.sync(); r.set_timing(True)
    for _ in range(10): r.update_device(x.data_ptr(), x.data_ptr(), True)
    kt = r.kernel_times()
    print(f"n={n} batch={batch} lazy={lazy} smooth={smooth} F={F}: spectrum {kt['spectrum_ms']/kt['spectrum_launches']*1e3:.1f} us", flush=True)
    r.close()
