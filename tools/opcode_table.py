#!/usr/bin/env python
"""Executed instructions of a kernel by opcode, from an ncu report joined with the SASS of the library
(tools/ncu_sass.py).  usage: tools/opcode_table.py <rep> <lib.so> <mangled-substr> <playouts>"""
import collections, re, subprocess, sys
rep, lib, kern, npl = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
out = subprocess.run([sys.executable, "tools/ncu_sass.py", rep, lib, kern], capture_output=True, text=True).stdout
tot = 0; by = collections.Counter(); st = collections.Counter()
for ln in out.splitlines():
    p = ln.split(None, 3)
    if len(p) < 4 or not p[0].isdigit():
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', p[3])
    op = m.group(2) if m else '?'
    op = 'BRA.DIV' if op.startswith('BRA.DIV') else op.split('.')[0]
    by[op] += int(p[0]); st[op] += 1; tot += int(p[0])
scaf = ('BSSY', 'BSYNC', 'WARPSYNC', 'ENDCOLLECTIVE', 'UMOV', 'BRA.DIV', 'NOP')
print("kernel %s: %d static instructions, %.0f executed per playout" % (kern, sum(st.values()), tot / npl))
print("convergence scaffolding (%s): %.2f %% of the executed instructions" % (", ".join(scaf), 100.0 * sum(by[k] for k in scaf) / tot))
for op, c in by.most_common(40):
    print("%-14s %6.2f%% %8.0f /playout  static %d" % (op, 100.0 * c / tot, c / npl, st[op]))
