"""Turn gpurun_out/{launches*.csv, prof*.ncu-rep} into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <tag> <launches.csv> <prof.ncu-rep>

Writes profiles/<tag>_launches.csv (per-launch device times), profiles/<tag>_ncu_summary.json/.md
(per-kernel key metrics of the `ncu --set full` capture) and, for the raster kernel,
profiles/raster_bars_traffic.json (dram bytes per launch, read by bench.py's roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    tag, launches = sys.argv[1:3]
    reps = sys.argv[3:]
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    # ---- launch list -----------------------------------------------------------------------------
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(list)
    with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
        f.write("id,kernel,block,grid,gpu_time_ns\n")
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "").replace("glb::", "")
            f.write(f'{r[0]},"{name}","{r[7]}","{r[8]}",{r[-1]}\n')
            agg[name].append(float(r[-1]))
    total = sum(sum(v) for v in agg.values())
    share = {k: {"launches": len(v), "avg_us": sum(v) / len(v) / 1e3, "share_of_listed_time": sum(v) / total} for k, v in agg.items()}
    # ---- full capture ------------------------------------------------------------------------------
    kernels = []
    rows_all = []
    for rep in reps:
        # a capture (.ncu-rep), or its `ncu -i … --page raw --csv` export made on the GPU box (the reports themselves are too
        # large to bring back: gpurun_out is capped at 64 MiB)
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        if len(rr) < 3:
            continue
        rows_all.append((rr[0], rr[1], rr[2:], os.path.basename(rep)))
    for hdr, units, body, repname in rows_all:
      for r in body:
          d = {"capture": repname, "kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("glb::", "")}
          for k in KEYS:
              if k in hdr:
                  v, u = r[hdr.index(k)], units[hdr.index(k)]
                  try:
                      v = float(v.replace(",", ""))
                  except ValueError:
                      continue
                  if k.startswith("dram__bytes") or k.startswith("lts__t_bytes"):
                      v *= SCALE.get(u, 1)
                      u = "byte"
                  if k == "gpu__time_duration.sum":
                      v *= SCALE.get(u, 1)
                      u = "ms"
                  d[k] = {"value": v, "unit": u}
          kernels.append(d)
    summary = {"tag": tag, "launch_list": share, "ncu_full": kernels,
               "note": "ncu times are cold-cache and serialised; compare kernel SHARES with bench.py, not absolutes"}
    json.dump(summary, open(os.path.join(out, f"{tag}_ncu_summary.json"), "w"), indent=1)
    with open(os.path.join(out, f"{tag}_ncu_summary.md"), "w") as f:
        f.write(f"# ncu summary {tag}\n\nLaunch list (`ncu --metrics gpu__time_duration.sum --clock-control none`):\n\n")
        f.write("| kernel | launches | avg µs | share |\n|---|---|---|---|\n")
        for k, v in share.items():
            f.write(f"| `{k}` | {v['launches']} | {v['avg_us']:.1f} | {v['share_of_listed_time']:.3f} |\n")
        f.write("\nFull capture (`ncu --set full --clock-control none --import-source on`):\n\n")
        for d in kernels:
            f.write(f"## `{d['kernel']}`  ({d.get('capture', '')})\n\n")
            for k in KEYS:
                if k in d:
                    f.write(f"- {k} = {d[k]['value']:.6g} {d[k]['unit']}\n")
            if "dram__bytes_write.sum" in d and "gpu__time_duration.sum" in d:
                tr = d["dram__bytes_read.sum"]["value"] + d["dram__bytes_write.sum"]["value"]
                f.write(f"- dram traffic per launch = {tr / 1e9:.4f} GB -> {tr / d['gpu__time_duration.sum']['value'] / 1e6:.0f} GB/s under ncu\n")
            f.write("\n")
    for d in kernels:
        if "raster_bars" in d["kernel"] and "dram__bytes_write.sum" in d:
            tr = d["dram__bytes_read.sum"]["value"] + d["dram__bytes_write.sum"]["value"]
            json.dump({"dram_bytes_per_launch": tr, "source": f"profiles/{tag}_ncu_summary.json", "kernel": d["kernel"],
                       "grid": d.get("launch__grid_size", {}).get("value")}, open(os.path.join(out, "raster_bars_traffic.json"), "w"))
            break
    print("wrote", out)


if __name__ == "__main__":
    main()
