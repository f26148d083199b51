"""BASELINE configs[2] and configs[4] on ONE GPU's share of the batch.

  configs[2]: radial, N=8192, 3840x2160, 4096 streams over 8 GPUs  -> 512 streams on this GPU
  configs[4]: sweep N in 512..16384 x {720p, 1080p, 4K, 8K}, 8192 streams over 8 GPUs -> 1024 streams on this GPU
              (bars; 8K frames do not fit 1024x resident -> framebuffer ring, fb_slots reported)

Per point: whole-step frames/s (device-resident PCM, spectrum + raster, two-stream overlap), the raster
kernel's isolated duration and its fraction of the measured HBM peak.  Writes gpurun_out/sweep.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g

try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
FB_BUDGET = 48e9          # bytes of framebuffer kept resident per GPU; beyond that a ring of slots


def run(module, n, w, h, batch, steps=16):
    p = g.default_params(module, n=n, w=w, h=h, lazy_smooth=1)
    frame = w * h * 4
    if batch * frame > FB_BUDGET:
        p.fb_slots = max(1, int(FB_BUDGET // frame))
    r = g.Renderer(p, batch=batch)
    st = torch.cuda.ExternalStream(r.cuda_stream)
    x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    y = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
    torch.cuda.synchronize()
    for _ in range(5):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    r.sync()
    r.set_timing(True)
    for _ in range(10):
        r.update_device(x.data_ptr(), y.data_ptr(), False)
    iso = r.kernel_times()
    r.set_timing(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        r.update_device(x.data_ptr(), y.data_ptr(), True)
    e1.record(st)
    r.sync()
    step = e0.elapsed_time(e1) / steps
    ras = iso["raster_ms"] / iso["raster_launches"]
    gbs = batch * frame / ras / 1e6
    row = dict(module=module, bufsize=n, width=w, height=h, streams_this_gpu=batch, fb_slots=int(p.fb_slots) or batch,
               step_ms=step, frames_per_s=batch / step * 1e3, raster_ms_isolated=ras, raster_gbs_isolated=gbs,
               raster_frac_of_hbm_peak=gbs / PEAK)
    r.close()
    del x, y
    torch.cuda.empty_cache()
    return row


def main():
    out = {"hbm_peak_gbs": PEAK, "points": []}
    only = sys.argv[1:]          # optional filters "module:n:WxH" (e.g. bars:8192:1920x1080)
    pts = [("radial", 8192, 3840, 2160, 512)]
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        for w, h in ((1280, 720), (1920, 1080), (3840, 2160), (7680, 4320)):
            pts.append(("bars", n, w, h, 1024))
    for mod in ("radial", "circle", "graph", "wave"):
        pts.append((mod, 4096, 1920, 1080, 1024))
    if only:
        pts = [pt for pt in pts if f"{pt[0]}:{pt[1]}:{pt[2]}x{pt[3]}" in only]
    for pt in pts:
        row = run(*pt)
        row["whole_step_frac_of_hbm_peak"] = pt[4] * pt[2] * pt[3] * 4 / row["step_ms"] / 1e6 / PEAK
        out["points"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
