"""Quick GPU probe: kernel timings per module (development aid, not the bench).

raster_iso = raster kernel alone (modified=0), raster_co / spectrum_co = inside full updates, where the
spectrum kernel of update i+1 co-runs with the raster kernel of update i; step = wall per update."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g

cfgs = [("bars", 4096, 1920, 1080, 1024, 0), ("bars", 4096, 1920, 1080, 1024, 1), ("radial", 8192, 3840, 2160, 64, 1),
        ("radial", 8192, 3840, 2160, 512, 1), ("radial", 4096, 1920, 1080, 256, 1), ("circle", 4096, 1920, 1080, 256, 1),
        ("graph", 2048, 1280, 720, 256, 1), ("wave", 2048, 1280, 720, 256, 1), ("graph", 2048, 1280, 720, 1024, 1),
        ("bars", 16384, 1280, 720, 512, 1), ("bars", 512, 1280, 720, 512, 1)]
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if c[0] in sys.argv[1:]]
out = []
for mod, n, w, h, batch, lazy in cfgs:
    p = g.default_params(mod, n=n, w=w, h=h, lazy_smooth=lazy)
    if batch * w * h * 4 > 40e9:
        p.fb_slots = int(40e9 // (w * h * 4))
    r = g.Renderer(p, batch=batch)
    st = torch.cuda.ExternalStream(r.cuda_stream)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    y = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    torch.cuda.synchronize()
    for _ in range(3):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    r.sync()
    r.set_timing(True)
    for _ in range(10):
        r.update_device(x.data_ptr(), y.data_ptr(), False)
    iso = r.kernel_times()
    r.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    e1.record(st)
    co = r.kernel_times()
    step = e0.elapsed_time(e1) / 10
    ras_iso = iso["raster_ms"] / iso["raster_launches"]
    ras_co = co["raster_ms"] / co["raster_launches"]; spec = co["spectrum_ms"] / co["spectrum_launches"]
    gb = batch * w * h * 4 / 1e9
    row = dict(module=mod, n=n, w=w, h=h, batch=batch, lazy=lazy, raster_iso_ms=ras_iso, raster_iso_gbs=gb / ras_iso * 1e3,
               raster_iso_frac=gb / ras_iso * 1e3 / 6569.6, raster_co_ms=ras_co, spectrum_co_ms=spec, step_ms=step, fps=batch / step * 1e3)
    out.append(row)
    print(f"{mod:7s} n={n:5d} {w}x{h} batch={batch:4d} lazy={lazy}: raster_iso {ras_iso:.3f} ms = {row['raster_iso_gbs']:5.0f} GB/s ({row['raster_iso_frac']:.2f}) | "
          f"co-run raster {ras_co:.3f} spectrum {spec:.3f} | step {step:.3f} ms = {row['fps']:.0f} fps", flush=True)
    r.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
