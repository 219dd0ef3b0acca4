#!/usr/bin/env python
"""Turn the ncu reports a `bash tools/gpu_profiles_r02.sh` call brought back (gpurun_out/r02_*.ncu-rep)
into the committed summaries under profiles/.  usage: tools/refresh_profiles_r02.py [tag]"""
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = "agentgo_b200/libagentgo_b200.so"
KERN = {"group8": "playout_group_kernelINS_3grp4GCfgILi19ELi8E", "group16": "playout_group_kernelINS_3grp4GCfgILi19ELi16E",
        "group32": "playout_group_kernelINS_3grp4GCfgILi19ELi32E", "warp": "playout_warp_kernelILi19ELi8ELi4ELb0"}


def run(*a):
    return subprocess.run([sys.executable] + list(a), capture_output=True, text=True).stdout


summary = {}
for k, mangled in KERN.items():
    rep = "gpurun_out/r02_%s.ncu-rep" % k
    m = run("tools/ncu_metrics.py", rep)
    open("profiles/%s_%s_metrics.txt" % (tag, k), "w").write(m)
    bl = run("tools/ncu_by_line.py", rep, lib, mangled, "400")
    open("profiles/%s_%s_by_line.txt" % (tag, k), "w").write(bl)
    if k != "warp":
        open("profiles/%s_%s_regions.txt" % (tag, k), "w").write(run("tools/regions_group.py", "profiles/%s_%s_by_line.txt" % (tag, k)))
        open("profiles/%s_%s_smem_by_line.txt" % (tag, k), "w").write(run("tools/ncu_by_line.py", rep, lib, mangled, "40", "--smem"))
    d = dict(l.strip().split(" = ") for l in m.splitlines() if " = " in l)
    g = lambda key: float([v for kk, v in d.items() if kk.startswith(key)][0])
    summary[k] = {"ms_100k_playouts_under_ncu": g("gpu__time_duration.sum"),
                  "warp_instructions_per_playout": g("smsp__inst_executed.sum") / 1e5,
                  "ipc_per_sm": g("sm__inst_executed.avg.per_cycle_elapsed"),
                  "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "alu_pipe_active_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                  "warps_per_sm": g("sm__warps_active.avg.per_cycle_active"), "registers": g("launch__registers_per_thread"),
                  "lanes_active_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
                  "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                  "smem_bank_conflict_wavefronts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}
json.dump(summary, open("profiles/%s_mappings.json" % tag, "w"), indent=1)
for k, v in summary.items():
    print("%-8s %6.1f K instr/playout  IPC %.2f  issue %.1f%%  warps %.1f  regs %d  %.2f ms/100k" % (
        k, v["warp_instructions_per_playout"] / 1e3, v["ipc_per_sm"], v["issue_active_pct"], v["warps_per_sm"],
        v["registers"], v["ms_100k_playouts_under_ncu"]))

# the peak micro-kernels: which pipe each saturates
raw = subprocess.run(["ncu", "-i", "gpurun_out/r02_peaks.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
with open("profiles/%s_peak_kernels.txt" % tag, "w") as f:
    f.write("ncu --set full of the micro-kernels behind agb_measure_peaks (diag.cu): launches in order are\n"
            "int_peak_kernel(mixed lop3 + mad.lo 1:1), int_peak_kernel(lop3 only), smem_peak_kernel, three times.\n\n")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        f.write("launch %s\n" % r[ix["ID"]])
        for w in want:
            if w in ix:
                f.write("  %s = %s\n" % (w, r[ix[w]]))
print(open("profiles/%s_peak_kernels.txt" % tag).read()[:1800])
