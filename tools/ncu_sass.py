#!/usr/bin/env python
"""Dump the SASS of the first kernel in an ncu report with per-instruction execution counts and
the CUDA source line of each instruction.  usage: ncu_sass.py <rep> <lib.so> <mangled-substr> [file:lo-hi]"""
import csv, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
flt = sys.argv[4] if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in out: continue
    insec, cur = False, ("?", 0)
    for ln in out.splitlines():
        if ln.startswith("\t.section"):
            insec = ".text." in ln and kern in ln; continue
        if not insec: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m: lines.append((cur, m.group(2).strip()))
        elif re.match(r"^\.L_x_\d+:", ln): lines.append((cur, ln.strip()))
    if lines: break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
k = 0
ffile, lo, hi = None, 0, 10**9
if flt:
    ffile, rng = flt.split(":"); lo, hi = [int(x) for x in rng.split("-")]
for cur, text in lines:
    if text.startswith(".L_x"):
        print("%12s %s" % ("", text)); continue
    cnt = int(data[k][ix["Instructions Executed"]]) if k < len(data) else -1
    smp = int(data[k][ix["# Samples"]]) if k < len(data) else -1
    k += 1
    if ffile and not (cur[0] == ffile and lo <= cur[1] <= hi): continue
    print("%12d %6d %-22s %s" % (cnt, smp, "%s:%d" % cur, text))
