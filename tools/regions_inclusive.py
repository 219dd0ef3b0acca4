#!/usr/bin/env python
"""Inclusive cost per code region: walks tools/ncu_sass.py output in address order and attributes
inlined helper / eval_cells instructions to the enclosing region (the last instruction seen whose
source line lies in a top-level function body)."""
import re, sys
src = open('agentgo_b200/csrc/playout_warp.cu').read().splitlines()
marks = [('refill', 'void refill('), ('eval_cells', 'Eval eval_cells('), ('complex', 'int random_piece_complex('),
         ('random_piece', 'int random_piece('), ('kill', 'void kill_string('), ('put', 'void put('),
         ('score', 'int score_diff('), ('kernel', 'playout_warp_kernel(const')]
starts = []
for name, pat in marks:
    for i, l in enumerate(src):
        if pat in l and '__' in l:
            starts.append((i + 1, name)); break
starts.sort()
def region(f, l):
    if f != 'playout_warp.cu': return None
    r = None
    for s, name in starts:
        if l >= s: r = name
    return r
P = float(sys.argv[2]) if len(sys.argv) > 2 else 100000.0
cur = 'kernel'; tot = {}; sub = {}
for ln in open(sys.argv[1]):
    m = re.match(r'\s*(\d+)\s+(\d+)\s+(\S+):(\d+)\s+(.*)', ln)
    if not m: continue
    cnt, smp, f, l = int(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4))
    r = region(f, l)
    inner = None
    if r == 'eval_cells': inner = 'eval'
    elif r is None: inner = 'helper'
    else: cur = r
    tot[cur] = tot.get(cur, 0) + cnt
    key = (cur, inner or 'own')
    sub[key] = sub.get(key, 0) + cnt
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    parts = ', '.join('%s %.0f' % (kk[1], vv / P) for kk, vv in sorted(sub.items()) if kk[0] == k)
    print('%-14s %7.0f /playout  %5.1f%%   (%s)' % (k, v / P, 100.0 * v / T, parts))
print('total %.0f' % (T / P))
