#!/usr/bin/env python
"""Aggregate tools/ncu_by_line.py output (all lines) by code region of playout_warp.cu."""
import re, sys
src = open('agentgo_b200/csrc/playout_warp.cu').read().splitlines()
marks = [('refill', 'void refill('), ('eval_cells', 'Eval eval_cells('), ('complex', 'int random_piece_complex('), ('random_piece', 'int random_piece('),
         ('kill', 'void kill_string('), ('put', 'void put('), ('score', 'int score_diff('), ('kernel', 'playout_warp_kernel(const')]
starts = []
for name, pat in marks:
    for i, l in enumerate(src):
        if pat in l and '__' in l:
            starts.append((i + 1, name)); break
starts.sort()
def region(f, l):
    if f == 'board_ops.h': return 'tinymt'
    if 'intrinsics' in f: return 'shfl/ballot'
    if f != 'playout_warp.cu': return f
    r = 'helpers'
    for s, name in starts:
        if l >= s: r = name
    return r
tot = {}
total_inst = None
for ln in open(sys.argv[1]):
    m = re.match(r'kernel instructions: static (\d+), executed (\d+)', ln)
    if m: total_inst = int(m.group(2))
    m = re.match(r'\s*([\d.]+)\s+([\d.]+)\s+(\d+)\s+(\S+):(\d+)', ln)
    if not m: continue
    inst, smp, st, f, l = float(m.group(1)), float(m.group(2)), int(m.group(3)), m.group(4), int(m.group(5))
    t = tot.setdefault(region(f, l), [0, 0, 0])
    t[0] += inst; t[1] += smp; t[2] += st
P = float(sys.argv[2]) if len(sys.argv) > 2 else 100000.0
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print('%-14s inst %5.1f%%  (%7.0f /playout)  samples %5.1f%%  static %d' % (k, v[0], v[0] / 100 * total_inst / P, v[1], v[2]))
print('total %.0f warp-instr/playout' % (total_inst / P))
