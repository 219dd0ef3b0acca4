#!/usr/bin/env python
"""Aggregate the per-source-line table of tools/ncu_by_line.py by code region of the group kernel
(playout_group.cuh).  usage: tools/regions_group.py <by_line.txt> [playouts-in-the-capture]"""
import collections
import re
import sys

rows = []
for ln in open(sys.argv[1]):
    m = re.match(r'\s*([\d.]+)\s+([\d.]+)\s+(\d+)\s+(\S+):(\d+)\s', ln)
    if m:
        rows.append((float(m.group(1)), float(m.group(2)), int(m.group(3)), m.group(4), int(m.group(5))))
    m = re.match(r'kernel instructions: static (\d+), executed (\d+)', ln)
    if m:
        executed = int(m.group(2))
npl = float(sys.argv[2]) if len(sys.argv) > 2 else 100000.0
src = open('agentgo_b200/csrc/playout_group.cuh').read().splitlines()


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    raise KeyError(s)


marks = [('eval_point', find('AGB_DEV Ev eval_point')), ('refill', find('struct GenState')),
         ('compact', find('AGB_NOINLINE int compact')), ('complex_pick', find('AGB_NOINLINE int complex_pick')),
         ('score', find('AGB_DEV int score_diff')), ('body_init', find('AGB_DEV void playout_group_body')),
         ('service', find('// ---- service')), ('select_hdr', find('#define AGB_ENSURE')),
         ('select (scan, decide)', find('const bool nbact = search && prev >= 0;')),
         ('complex_call', find('// complex mode, :413-478')), ('put', find('// ---- Board::put')),
         ('kill', find('// Board::killSetNode, :293-339')), ('pair', find('// ---- pair loop'))]
agg = collections.Counter(); smp = collections.Counter(); st = collections.Counter()
for i, s, n, f, l in rows:
    if f == 'playout_group.cuh':
        key = 'accessors' if l < marks[0][1] else [m[0] for m in marks if m[1] <= l][-1]
    elif f == 'board_ops.h':
        key = 'tinymt (refill)'
    else:
        key = f + ':' + str(l)
    agg[key] += i; smp[key] += s; st[key] += n
print("%-30s %8s %8s %8s %7s" % ("region", "inst%", "/playout", "samples%", "static"))
for k, v in agg.most_common():
    if v >= 0.05:
        print("%-30s %7.2f%% %8.0f %7.2f%% %7d" % (k, v, v / 100.0 * executed / npl, smp[k], st[k]))
print("total %.0f warp-instructions per playout" % (executed / npl))
