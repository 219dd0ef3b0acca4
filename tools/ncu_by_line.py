#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with nvdisasm line info and aggregate by
CUDA source line (inlined code is attributed to the line where it was written).

usage: tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top] [--smem]
--smem: rank the lines by shared-memory wavefronts instead and show, per line, the wavefronts
the accesses needed, the ideal number and the excess (bank conflicts).
Needs -lineinfo at compile time.  Works offline (no GPU)."""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    smem_mode = "--smem" in sys.argv
    argv = [a for a in sys.argv if a != "--smem"]
    rep, lib, kern = argv[1], argv[2], argv[3]
    top = int(argv[4]) if len(argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    lines = []   # (file, line) per instruction of the kernel, in order
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern not in out:
            continue
        insec, cur = False, ("?", 0)
        for ln in out.splitlines():
            if ln.startswith("\t.section") or ln.startswith("//-----"):
                insec = (".text." in ln and kern in ln) if ln.startswith("\t.section") else insec
                continue
            if not insec:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        if lines:
            break
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # find the section of the wanted kernel
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name":      # first kernel of the report (names are demangled here)
            start = i
            break
    hdr = rows[start + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    n = min(len(data), len(lines))
    if len(data) != len(lines):
        sys.stderr.write("warning: %d ncu rows vs %d disassembled instructions\n" % (len(data), len(lines)))
    inst = collections.Counter(); smp = collections.Counter(); static = collections.Counter()
    wav = collections.Counter(); exc = collections.Counter(); ideal = collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.defaultdict(collections.Counter)
    for k in range(n):
        key = lines[k]
        inst[key] += int(data[k][ix["Instructions Executed"]])
        smp[key] += int(data[k][ix["# Samples"]])
        static[key] += 1
        if "L1 Wavefronts Shared" in ix:
            wav[key] += int(data[k][ix["L1 Wavefronts Shared"]] or 0)
            exc[key] += int(data[k][ix["L1 Wavefronts Shared Excessive"]] or 0)
            ideal[key] += int(data[k][ix["L1 Wavefronts Shared Ideal"]] or 0)
        for h in stall_cols:
            v = data[k][ix[h]]
            if v and v != "0":
                stalls[key][h] += int(v)
    ti, ts = sum(inst.values()), sum(smp.values())
    print("kernel instructions: static %d, executed %d, samples %d" % (n, ti, ts))
    src_cache = {}

    def src(key):
        f, l = key
        if f not in src_cache:
            for d in ("agentgo_b200/csrc", "."):
                p = os.path.join(d, f)
                if os.path.exists(p):
                    src_cache[f] = open(p).read().splitlines()
                    break
            else:
                src_cache[f] = []
        s = src_cache[f]
        return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""

    if smem_mode:
        tw, te = sum(wav.values()), sum(exc.values())
        print("shared-memory wavefronts %d, ideal %d, excessive (bank conflicts) %d = %.1f%% of all wavefronts" %
              (tw, sum(ideal.values()), te, 100.0 * te / max(1, tw)))
        print("%7s %7s %9s  %-22s %s" % ("wavef%", "excess%", "exc/ideal", "where", "source"))
        for key, v in sorted(wav.items(), key=lambda kv: -kv[1])[:top]:
            if v == 0:
                break
            print("%7.2f %7.2f %9.2f  %-22s %s" % (100.0 * v / tw, 100.0 * exc[key] / max(1, te), exc[key] / max(1, ideal[key]),
                                                    "%s:%d" % key, src(key)))
        return
    print("%6s %6s %6s  %-22s %s" % ("inst%", "smp%", "static", "where", "source / top stalls"))
    for key, v in sorted(smp.items(), key=lambda kv: -kv[1])[:top]:
        st = ", ".join("%s %d%%" % (h[6:], 100 * c / max(1, smp[key])) for h, c in stalls[key].most_common(3))
        print("%6.2f %6.2f %6d  %-22s %s   [%s]" % (100.0 * inst[key] / ti, 100.0 * v / ts, static[key],
                                                     "%s:%d" % key, src(key), st))
    agg = collections.Counter()
    for h in stall_cols:
        agg[h] = sum(stalls[k][h] for k in stalls)
    print("stall totals:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * c / max(1, ts)) for h, c in agg.most_common(8)))


if __name__ == "__main__":
    main()
