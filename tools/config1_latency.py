"""BASELINE configs[0]: single 22050 Hz stereo FIFO stream, N=1024, pipeline A (setaccelfft false),
spectrum only — latency of one update through the C ABI (host rings in -> spectrum back on the host),
next to the reference's own compiled transforms (oracle/_ref, when built) on one host core.
Parity for this config is tests/test_gpu_spectrum.py::test_config1_pipeline_a_vs_reference_golden."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glava_b200 as g

N, REPS = 1024, 2000
p = g.default_params("bars", n=N, w=64, h=4, accel_fft=0, smooth_pass=0)
rings = g.StreamRings(1, N, pinned=True)
for _ in range(N // 256 + 2):
    rings.advance()
res = {"config": "configs[0]: 1 stream, N=1024, pipeline A, no raster consumer (a 64x4 dummy frame is rendered)"}
with g.Renderer(p, batch=1) as r:
    for _ in range(50):
        r.update(rings.lb, rings.rb, True); r.sync()
    lat = []
    for _ in range(REPS):
        t0 = time.perf_counter()
        r.update(rings.lb, rings.rb, True)
        r.sync()
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e6
    r.set_timing(True)
    for _ in range(200):
        r.update(rings.lb, rings.rb, True)
    kt = r.kernel_times()
    res["gpu"] = {"update_plus_sync_us_median": float(np.median(lat)), "p99_us": float(np.percentile(lat, 99)),
                  "spectrum_kernel_us": kt["spectrum_ms"] / kt["spectrum_launches"] * 1e3}
try:
    from oracle.oracle import Reference, Oracle
    if Reference.available():
        ref = Reference(); o = Oracle("libm")
        op = o.default_params("bars", n=N, accel_fft=0, smooth_pass=0)
        cl, cr = ref.chan(op), ref.chan(op)
        t0 = time.perf_counter()
        for _ in range(REPS):
            ref.update_a(cl, rings.lb[0]); ref.update_a(cr, rings.rb[0])
        res["reference_cpu_1core"] = {"stereo_update_us": (time.perf_counter() - t0) / REPS * 1e6,
                                      "what": "render.c transform_fft + transform_gravity + transform_average, both channels (incl. ctypes call overhead)"}
except Exception as e:  # noqa: BLE001
    res["reference_cpu_1core"] = {"error": str(e)}
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/config1_latency.json", "w"), indent=1)
