"""A few spectrum-dominated updates at the headline need-list (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import glava_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = g.default_params("bars", n=n, w=1920, h=16, lazy_smooth=1)
r = g.Renderer(p, batch=1024)
x = (torch.rand(1024, n, device="cuda") - 0.5) * 0.2
for _ in range(4): r.update_device(x.data_ptr(), x.data_ptr(), True)
r.sync()
