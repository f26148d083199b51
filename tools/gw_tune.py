"""graph / wave raster: rows per CTA x registers (GLAVA_B200_ROWS, GLAVA_B200_GW_MINB), with and without the per-stream
column table, in ONE process (launch_raster reads those two knobs at every launch).  Development aid; prints raster-kernel
time (CUDA events around the raster launches, spectrum path co-running) and the whole step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, glava_b200 as g

PEAK = float(os.environ.get("PEAK_GBS", "6569.6"))
ROWS = tuple(int(v) for v in os.environ.get("GW_ROWS", "30,45,60,90,135,180,270").split(","))
CASES = [("graph", 4096, 1920, 1080, 1024), ("wave", 4096, 1920, 1080, 1024), ("graph", 2048, 1280, 720, 256), ("wave", 2048, 1280, 720, 256)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if f"{c[0]}{c[3]}" in sys.argv[1:]]

def measure(r, x, K=20):
    for _ in range(3): r.update_device(x.data_ptr(), x.data_ptr(), True)
    r.sync(); r.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(K): r.update_device(x.data_ptr(), x.data_ptr(), True)
    r.sync(); e1.record(); torch.cuda.synchronize()
    kt = r.kernel_times(); r.set_timing(False)
    return kt["raster_ms"] / kt["raster_launches"], e0.elapsed_time(e1) / K

def main():
    for module, n, w, h, batch in CASES:
      x = (torch.rand(batch, n, device="cuda") - 0.5) * 0.2
      gb = batch * w * h * 4 / 1e6
      p = g.default_params(module, n=n, w=w, h=h, lazy_smooth=1)
      for tab in (0, 1):
          if tab: os.environ.pop("GLAVA_B200_NO_COLTAB", None)
          else: os.environ["GLAVA_B200_NO_COLTAB"] = "1"
          with g.Renderer(p, batch=batch) as r:
              for minb in ((4, 5) if tab and not os.environ.get("GW_MINB4") else (4,)):
                  for rows in (ROWS if tab else (0,)):
                      os.environ["GLAVA_B200_GW_MINB"] = str(minb)
                      if rows: os.environ["GLAVA_B200_ROWS"] = str(rows)
                      else: os.environ.pop("GLAVA_B200_ROWS", None)
                      ras, step = measure(r, x)
                      print(f"{module:5s} {w}x{h} x{batch} table={tab} minb={minb} rows={rows:4d}: raster {ras:.4f} ms = {gb / ras / PEAK:.3f}   step {step:.4f} ms = {gb / step / PEAK:.3f}", flush=True)
      os.environ.pop("GLAVA_B200_ROWS", None)


if __name__ == "__main__":
    main()
