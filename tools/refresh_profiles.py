#!/usr/bin/env python
"""Turn the raw files a gpurun call brought back (gpurun_out/) into the committed summaries under
profiles/.  usage: tools/refresh_profiles.py <version-tag e.g. v12> <ncu-rep>"""
import csv, json, os, statistics, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
K = "playout_warp_kernelILi19ELi8ELi4ELb0"
lib = "agentgo_b200/libagentgo_b200.so"
os.system("cp gpurun_out/bench_r01.json profiles/bench_r01.json")
os.system("grep -v '^==' gpurun_out/launches_r01.csv > profiles/r01_launches_bench.csv")
open("profiles/r01_warp_%s_metrics.txt" % tag, "w").write(subprocess.run([sys.executable, "tools/ncu_metrics.py", rep], capture_output=True, text=True).stdout)
open("profiles/r01_warp_%s_by_line.txt" % tag, "w").write(subprocess.run([sys.executable, "tools/ncu_by_line.py", rep, lib, K, "2000"], capture_output=True, text=True).stdout)
sass = subprocess.run([sys.executable, "tools/ncu_sass.py", rep, lib, K], capture_output=True, text=True).stdout
open("/tmp/sass_%s.txt" % tag, "w").write(sass)
open("profiles/r01_warp_%s_regions.txt" % tag, "w").write(subprocess.run([sys.executable, "tools/regions_inclusive.py", "/tmp/sass_%s.txt" % tag], capture_output=True, text=True).stdout)
rows = [r for r in csv.reader(open("profiles/r01_launches_bench.csv")) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    per.setdefault((int(r[ix["ID"]]), r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]])
play = [v for k, v in per.items() if "playout_warp_kernel" in k[1]]
fill = [v for k, v in per.items() if "FillFunctor" in k[1]]
m = dict(l.strip().split(" = ") for l in open("profiles/r01_warp_%s_metrics.txt" % tag) if " = " in l)
inst = [float(v) for k, v in m.items() if k.startswith("smsp__inst_executed.sum")][0]
ipc = [float(v) for k, v in m.items() if k.startswith("sm__inst_executed.avg.per_cycle_elapsed")][0]
# median: a launch that follows the 256 MiB L2-flush fill sometimes absorbs the write-back of its dirty lines
traffic = statistics.median(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in play)
dur = statistics.mean(v["gpu__time_duration.sum"] for v in play) / 1e6
fd = statistics.mean(v["gpu__time_duration.sum"] for v in fill) / 1e6
json.dump({"kernel": "playout_warp_kernel<19,8,4,false>", "launches": len(play), "dram_bytes_per_launch": traffic,
           "ms_per_launch_under_ncu": dur, "l2_flush_fill_ms": fd, "kernel_share_of_step": dur / (dur + fd),
           "source": "profiles/r01_launches_bench.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu-baseline)",
           "algorithmic_bytes_per_launch": 4.0e6, "warp_instructions_per_playout": inst / 100000.0,
           "warp_instructions_source": "profiles/r01_warp_%s_metrics.txt (ncu smsp__inst_executed.sum / 100000 playouts)" % tag,
           "ipc_per_sm": ipc}, open("profiles/roofline_traffic.json", "w"), indent=1)
d = json.load(open("profiles/bench_r01.json"))
print("value %.0f e2e %.0f cpu %.0f traffic %.0f instr/playout %.0f ipc %.2f ms/launch(ncu) %.1f" %
      (d["value"], d["e2e"]["value"], d["cpu_baseline"]["value"], traffic, inst / 1e5, ipc, dur))
print(open("profiles/r01_warp_%s_regions.txt" % tag).read())
