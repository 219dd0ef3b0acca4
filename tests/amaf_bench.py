"""Throughput of candidate-mode playouts with and without AMAF marks (19x19, 361 candidates)."""
import json, sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
n, P = 19, 1000
e = ag.Engine(n)
cands = [(i // n, i % n) for i in range(n * n)]
e.playouts(1, cands, 50)
w, sp, sn, _, st0 = e.playouts(1, cands, P, avrg_win=0)
e.playouts_amaf(1, cands, 50)
w2, sp2, sn2, a, st1 = e.playouts_amaf(1, cands, P, avrg_win=0)
assert (w == w2).all() and (sp == sp2).all() and (sn == sn2).all()
print(json.dumps({"config": "19x19 empty board, all 361 candidates x 1000 playouts, candidate mode",
                  "plain_ms": st0["device_ms"], "plain_playouts_per_s": st0["playouts"] / st0["device_ms"] * 1e3,
                  "amaf_ms": st1["device_ms"], "amaf_playouts_per_s": st1["playouts"] / st1["device_ms"] * 1e3,
                  "amaf_nonzero_cells": int((a != 0).sum())}))
