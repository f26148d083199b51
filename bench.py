#!/usr/bin/env python
"""bench.py — headline benchmark of the PCM -> spectrum -> pixels hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): module `bars`, setbufsize 4096 ("4096-pt FFT"),
1920x1080 RGBA8, 1024 independent synthetic 22050 Hz stereo streams PER GPU (weak scaling: the
batch is sharded stream-wise, no inter-GPU collective on the data path).

One step = one rd_update(modified=true) for every stream of the batch: fused spectrum kernel
(window + FFT + log + gravity + average + smoothing) and one bars frame per stream.

  value  frames/s, inputs already resident in HBM (glava_b200_update_device)
  e2e    frames/s through the reference-facing C-ABI call with HOST rings
         (glava_b200_update: H2D of lb/rb inside the timed region, then a D2H read-back of one
         stream's framebuffer, the "optional cudaMemcpy readback for inspection")
  roofline  raster kernel: W*H*4 algorithmic bytes per frame / its CUDA-event duration, vs the
         measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU path on the host cores (oracle/cpu_baseline.py), bounded sample

`--impl reference` times the reference's own CPU implementation of the same path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODULE, N, W, H, HOP = "bars", 4096, 1920, 1080, 256
METRIC = "spectrum frames/sec @4096-pt FFT, 1920x1080 bars"
UNIT = "frames/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        # "under load": upper half of the samples (the sampler also sees the idle gaps around the region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the raster kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "raster_bars_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def params_dict():
    return dict(module=MODULE, n=N, w=W, h=H)


def run_reference(args, rank, world):
    """reference arm: the reference's CPU path on all host cores, bounded sample per step"""
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuBaseline, describe
    base = CpuBaseline(params_dict(), streams_per_worker=1)
    for _ in range(args.warmup):
        base.step()
    t, frames = 0.0, 0
    for _ in range(args.steps):
        dt, f = base.step()
        t += dt; frames += f
    base.close()
    value = frames / t
    sample = f"{base.cores} frames per step (1 stream per core) x {args.steps} steps; " + describe(base.kind)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "bars module, 4096-sample buffer (2048-pt complex FFT), 1920x1080 RGBA8", "module": MODULE,
                   "bufsize": N, "width": W, "height": H, "frames_per_step": frames // args.steps, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base.cores, "kind": base.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=env_int("GLAVA_BENCH_BATCH", 1024), help="streams per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = env_int("RANK", 0); world = env_int("WORLD_SIZE", 1); local = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import glava_b200 as g
    from glava_b200.synth import synth_batch_int16

    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner ("NCCL version ...") straight to fd 1 when the communicator is created; stdout must
        # carry the one JSON line only, so communicator creation (init + first collective) runs with fd 1 -> fd 2
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    batch, K, Wm = args.batch, args.steps, args.warmup
    p = g.default_params(MODULE, n=N, w=W, h=H)
    p.lazy_smooth = 1
    r = g.Renderer(p, batch=batch, device=local)
    stream = torch.cuda.ExternalStream(r.cuda_stream, device=local)

    # ---- synthetic input: this rank's shard of the global stream set -----------------------------
    nsnap = min(K + Wm, 4)                                   # distinct ring snapshots cycled through
    total = N + nsnap * HOP
    pcm = synth_batch_int16(rank * batch, batch, 0, total)  # [batch][total][2] int16
    pcm_f = pcm.astype(np.float32) / np.float32(65535)      # fifo.c:104-107
    host_l = [g.pinned_empty((batch, N), np.float32) for _ in range(nsnap)]
    host_r = [g.pinned_empty((batch, N), np.float32) for _ in range(nsnap)]
    dev_l, dev_r = [], []
    for i in range(nsnap):
        host_l[i][:] = pcm_f[:, (i + 1) * HOP:(i + 1) * HOP + N, 0]
        host_r[i][:] = pcm_f[:, (i + 1) * HOP:(i + 1) * HOP + N, 1]
        dev_l.append(torch.from_numpy(host_l[i]).to(f"cuda:{local}"))
        dev_r.append(torch.from_numpy(host_r[i]).to(f"cuda:{local}"))
    torch.cuda.synchronize()

    def barrier():
        if distributed:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not distributed:
            return ms
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg ("value") --------------------------------------------------------------
    for i in range(Wm):
        r.update_device(dev_l[i % nsnap].data_ptr(), dev_r[i % nsnap].data_ptr(), True)
    r.sync()
    sampler = ClockSampler(local); sampler.start()
    t_sampler = time.perf_counter()
    time.sleep(0.25)
    r.set_timing(True)
    launches0 = r.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        r.update_device(dev_l[(Wm + i) % nsnap].data_ptr(), dev_r[(Wm + i) % nsnap].data_ptr(), True)
    e1.record(stream)
    r.sync()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    kt = r.kernel_times()
    launches = r.launch_count - launches0
    # the raster kernel alone (no spectrum kernel co-running): modified=0 re-rasters the last spectrum
    r.set_timing(True)
    for _ in range(10):
        r.update_device(dev_l[0].data_ptr(), dev_r[0].data_ptr(), False)
    kt_iso = r.kernel_times()
    r.set_timing(False)

    # ---- end-to-end leg: host rings in, one framebuffer out, every step -------------------------------
    frame = np.empty((H, W, 4), np.uint8)
    frame_pinned = g.pinned_empty((H, W, 4), np.uint8)
    for i in range(2):
        r.update(host_l[i % nsnap], host_r[i % nsnap], True); r.readback(i % batch, frame_pinned)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for i in range(K):
        r.update(host_l[i % nsnap], host_r[i % nsnap], True)     # H2D (copy stream) + kernels
        r.readback_async(i % batch, frame_pinned)                # D2H of one stream's frame (snapshot + separate copy-out stream)
    r.readback_fence()                                           # e3 is ordered after the last frame's D2H
    e3.record(stream)
    r.sync()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))

    # ---- end-to-end through the FIFO ingest entry point (fifo.c semantics: only the 256 new frames per
    #      stream cross PCIe, the rings live in HBM) -------------------------------------------------------
    fifo_chunks = [g.pinned_empty((batch, HOP * 2), np.int16) for _ in range(nsnap)]
    for i in range(nsnap):
        fifo_chunks[i][:] = pcm[:, N + i * HOP:N + (i + 1) * HOP, :].reshape(batch, HOP * 2)
    for i in range(2):
        r.ingest_fifo(fifo_chunks[i % nsnap]); r.update_rings(True); r.readback_async(i % batch, frame_pinned)
    r.sync()
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record(stream)
    for i in range(K):
        r.ingest_fifo(fifo_chunks[i % nsnap])
        r.update_rings(True)
        r.readback_async(i % batch, frame_pinned)
    r.readback_fence()
    e5.record(stream)
    r.sync()
    barrier()
    ms_fifo = max_over_ranks(e4.elapsed_time(e5))
    # nvidia-smi samples every 100 ms and short runs (small --steps) end sooner than that: keep the SAME load
    # running, untimed, until the sampler has seen about 2 s of it, so the clocks line describes the load
    soak = 0
    while time.perf_counter() - t_sampler < 2.25:
        for i in range(16):
            r.update_device(dev_l[i % nsnap].data_ptr(), dev_r[i % nsnap].data_ptr(), True)
        r.sync(); soak += 16
    time.sleep(0.1)
    clocks = sampler.stop()
    clocks["untimed_soak_steps"] = soak
    frame[:] = frame_pinned
    checksum = int(frame.astype(np.uint32).sum())

    total_frames = batch * world * K
    value = total_frames / (ms_dev / 1e3)
    e2e_value = total_frames / (ms_e2e / 1e3)
    peak, peak_src = hbm_peak()
    ras_ms = kt["raster_ms"] / max(kt["raster_launches"], 1)
    spec_ms = kt["spectrum_ms"] / max(kt["spectrum_launches"], 1)
    ras_iso_ms = kt_iso["raster_ms"] / max(kt_iso["raster_launches"], 1)
    alg_bytes = batch * W * H * 4
    achieved = alg_bytes / (ras_ms / 1e3) / 1e9

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuBaseline, describe
        base = CpuBaseline(params_dict(), streams_per_worker=1)
        base.step()
        t, f = 0.0, 0
        t_end = time.perf_counter() + 12.0
        steps = 0
        while steps < 3 or (time.perf_counter() < t_end and steps < 40):
            dt, fr = base.step(); t += dt; f += fr; steps += 1
        base.close()
        cpu = {"value": f / t, "unit": UNIT, "cores": base.cores, "kind": base.kind,
               "sample": f"{f} frames ({base.cores} per step, {steps} steps) of the same workload; " + describe(base.kind)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "bars module, 4096-sample buffer (2048-pt complex FFT), 1920x1080 RGBA8, "
                                   f"batch={batch} streams per GPU", "module": MODULE, "bufsize": N, "width": W, "height": H,
                       "batch_per_gpu": batch, "streams_total": batch * world, "parallelism": f"stream-sharded x{world}, no collective",
                       "pipeline": "B (setaccelfft true: R16 gravity/average/smooth passes)", "lazy_smooth": int(p.lazy_smooth),
                       "l2": "every step writes batch*8.29 MB of framebuffer (>> 126 MB L2) and cycles 4 input snapshots; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": "raster_bars_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ras_ms,
                         "spectrum_kernel_ms": spec_ms, "raster_share_of_step": ras_ms / (ms_dev / K),
                         "note": "kernel_ms is measured inside the timed steps, where the spectrum kernel of update i+1 "
                                 "co-runs with the raster kernel of update i (two streams); *_isolated = the raster kernel alone",
                         "kernel_ms_isolated": ras_iso_ms, "achieved_isolated": alg_bytes / (ras_iso_ms / 1e3) / 1e9,
                         "frac_isolated": alg_bytes / (ras_iso_ms / 1e3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * batch * N * 4, "d2h_bytes_per_step": W * H * 4,
                    "ms_per_step": ms_e2e / K, "readback_checksum": checksum,
                    "fifo_path": {"value": total_frames / (ms_fifo / 1e3), "unit": UNIT, "h2d_bytes_per_step": batch * HOP * 2 * 2,
                                  "d2h_bytes_per_step": W * H * 4, "ms_per_step": ms_fifo / K,
                                  "note": "glava_b200_ingest_fifo + glava_b200_update_rings: raw int16 FIFO chunks in, rings resident in HBM"}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    r.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
