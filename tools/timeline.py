"""Timeline of the two streams over a few headline steps: where do the ~75 us per step between the raster kernel's
own duration and the step time go?  (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g

p = g.default_params("bars", n=4096, w=1920, h=1080, lazy_smooth=1)
batch = 1024
r = g.Renderer(p, batch=batch)
x = (torch.rand(batch, 4096, device="cuda") - 0.5) * 0.2
y = (torch.rand(batch, 4096, device="cuda") - 0.5) * 0.2
torch.cuda.synchronize()
for _ in range(5):
    r.update_device(x.data_ptr(), y.data_ptr(), True)
r.sync()
r.set_timing(True)
for _ in range(12):
    r.update_device(x.data_ptr(), y.data_ptr(), True)
sp, ra = r.timeline()
print("step  spec[start end]   raster[start end]   raster_dur  gap_to_prev_raster  spec_end-prev_raster_end")
for i in range(len(ra)):
    gap = ra[i][0] - ra[i - 1][1] if i else 0.0
    lag = sp[i][1] - ra[i - 1][1] if i else 0.0
    print(f"{i:3d}  {sp[i][0]:8.3f} {sp[i][1]:8.3f}   {ra[i][0]:8.3f} {ra[i][1]:8.3f}   {ra[i][1]-ra[i][0]:6.3f}   {gap:7.3f}   {lag:7.3f}")
print("mean step", (ra[-1][1] - ra[1][1]) / (len(ra) - 2))
