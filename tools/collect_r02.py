#!/usr/bin/env python
"""Copy what tools/gpu_final_r02.sh brought back (gpurun_out/) into profiles/ and derive
profiles/roofline_traffic.json from the ncu launch list.  usage: tools/collect_r02.py"""
import csv, json, os, shutil, statistics, subprocess, sys
for f in ("bench_r02", "bench_ref_r02", "config1_r02", "config3_r02", "selfplay_r02"):
    shutil.copy("gpurun_out/%s.json" % f, "profiles/%s.json" % f)
shutil.copy("gpurun_out/quick_bench_r02.txt", "profiles/r02_quick_bench.txt")
open("profiles/r02_launches_bench.csv", "w").write("".join(l for l in open("gpurun_out/r02_launches_bench.csv") if not l.startswith("==")))
print(subprocess.run([sys.executable, "tools/refresh_profiles_r02.py", "r02"], capture_output=True, text=True).stdout[:600])
rows = [r for r in csv.reader(open("profiles/r02_launches_bench.csv")) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
per = {}
for r in rows[1:]:
    per.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
play = [v for k, v in per.items() if "playout_group_kernel" in k[1]]
fill = [v for k, v in per.items() if "FillFunctor" in k[1]]
m = json.load(open("profiles/r02_mappings.json"))["group8"]
dur = statistics.mean(v["gpu__time_duration.sum"] for v in play) / 1e6
fd = statistics.mean(v["gpu__time_duration.sum"] for v in fill) / 1e6 if fill else 0.0
json.dump({"kernel": "playout_group_kernel<GCfg<19, 8 lanes per playout, 16 draws per lane>, 25 warps, 1 block per SM>",
           "launches": len(play),
           "dram_bytes_per_launch": statistics.median(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in play),
           "ms_per_launch_under_ncu": dur, "l2_flush_fill_ms": fd, "kernel_share_of_step": dur / (dur + fd),
           "source": "profiles/r02_launches_bench.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras)",
           "algorithmic_bytes_per_launch": 4.0e6, "warp_instructions_per_playout": m["warp_instructions_per_playout"],
           "warp_instructions_source": "profiles/r02_group8_metrics.txt (ncu smsp__inst_executed.sum / 100000 playouts)",
           "ipc_per_sm": m["ipc_per_sm"], "alu_pipe_active_pct": m["alu_pipe_active_pct"], "alu_share_of_instructions": 0.583}, open("profiles/roofline_traffic.json", "w"), indent=1)
d = json.load(open("profiles/bench_r02.json"))
print("bench value %.4g e2e %.4g cpu %.4g  int_issue %.3f issue_slots %s smem %.3f genmove %.1f ms strong %.4g" % (
    d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["roofline"]["frac"],
    d["roofline"].get("issue_slots", {}).get("frac"), d["roofline"]["smem"]["frac"], d["genmove_1M"]["device_ms_median"],
    d["strong_16M"]["playouts_per_s"]))
for f in ("config1_r02", "config3_r02", "selfplay_r02"):
    c = json.load(open("profiles/%s.json" % f))
    print(f, "%.4g" % c["value"], "cpu", (c.get("cpu_baseline") or {}).get("value"), (c.get("cpu_baseline") or {}).get("one_thread", {}).get("value"),
          (c.get("genmove") or {}).get("device_ms_median"))
