"""Spectrum-kernel timing probe (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g
for mod, n, batch, lazy, accel in (("bars", 4096, 1024, 1, 1), ("bars", 4096, 1024, 1, 0), ("bars", 4096, 1024, 0, 1), ("bars", 16384, 512, 1, 1), ("bars", 1024, 1024, 1, 1)):
    p = g.default_params(mod, n=n, w=64, h=16, lazy_smooth=lazy, accel_fft=accel)
    r = g.Renderer(p, batch=batch)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    for _ in range(3): r.update_device(x.data_ptr(), x.data_ptr(), True)
    r.sync(); r.set_timing(True)
    for _ in range(10): r.update_device(x.data_ptr(), x.data_ptr(), True)
    kt = r.kernel_times()
    print(f"{mod} n={n} batch={batch} lazy={lazy} accel={accel}: spectrum {kt['spectrum_ms']/kt['spectrum_launches']*1e3:.1f} us", flush=True)
    r.close()
