"""BASELINE.json configs[0]: 9x9 empty board, all 81 points as candidates, 10k playouts each (the
reference's CPU-runnable case) on the GPU; prints one JSON line.  usage: python tests/run_config1.py"""
import json
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
e = ag.Engine(9)
cands = [(i // 9, i % 9) for i in range(81)]
e.playouts(ag.BLACK, cands, 200)
w, sp, sn, _, st = e.playouts(ag.BLACK, cands, 10000)
print(json.dumps({"config": "BASELINE configs[0]: 9x9 empty board, 81 candidates x 10k playouts, candidate mode",
                  "device_ms": st["device_ms"], "playouts_per_s": st["playouts"] / st["device_ms"] * 1e3,
                  "command": "python tests/run_config1.py"}))
