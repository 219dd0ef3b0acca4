"""One-off parity soak (not part of the default suites): many random positions of every length,
ALL empty points as candidates (including suicides and ko points), both colours, the group kernel
(8 lanes per playout: AGB_KERNEL_AUTO's choice for large launches), AMAF marks included, against the
oracle.  usage: python tests/soak_parity.py [positions] [P] [group8|warp]"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np
import agentgo_b200 as ag
from oracle import pyoracle as po

NPOS = int(sys.argv[1]) if len(sys.argv) > 1 else 120
P = int(sys.argv[2]) if len(sys.argv) > 2 else 24
KERNEL = {'group8': ag.KERNEL_GROUP8, 'warp': ag.KERNEL_WARP}[sys.argv[3] if len(sys.argv) > 3 else 'group8']
rng = np.random.RandomState(7)
bad, total, t0 = 0, 0, time.time()
for k in range(NPOS):
    n = [9, 13, 19][k % 3]
    length = int(rng.randint(0, 3 * n * n))
    mv = po.OracleBoard(n, "best").random_game(1 + k % 2, length, 1000 + k)
    ob = po.OracleBoard(n, "best").replay(mv)
    st = ob.export_state()
    empties = [i for i in range(n * n) if st[16 + i] == 0]
    if not empties:
        continue
    cands = [(i // n, i % n) for i in empties]
    colour = 1 + (k // 3) % 2
    avrg = int(rng.randint(-5, 6))
    e = ag.Engine(n, kernel=KERNEL).replay(mv)
    w, sp, sn, d, _ = e.playouts(colour, cands, P, avrg_win=avrg, seed_base=k, position_id=k, want_diffs=True)
    ow, osp, osn, od = ob.playouts(colour, cands, P, avrg_win=avrg, seed_base=k, position_id=k, threads=0)
    r = e.playouts(colour, None, 4 * P, seed_base=k, position_id=k, want_diffs=True)
    orr = ob.playouts(colour, None, 4 * P, seed_base=k, position_id=k, threads=0)
    good = (d == od).all() and (w == ow).all() and (sp == osp).all() and (sn == osn).all() and (r[3] == orr[3]).all()
    a = e.playouts_amaf(colour, cands[:12], P, avrg_win=avrg, seed_base=k)
    oa = ob.playouts_amaf(colour, cands[:12], P, avrg_win=avrg, seed_base=k)
    good = good and (a[3] == oa[3]).all()
    total += len(cands) * P + 4 * P
    if not good:
        bad += 1
        print("MISMATCH position", k, "n", n, "moves", length, file=sys.stderr)
print(json.dumps({"positions": NPOS, "playouts_compared": total, "mismatching_positions": bad,
                  "seconds": time.time() - t0, "kernel": sys.argv[3] if len(sys.argv) > 3 else "group8", "oracle": po.OracleBoard(9, "best").kind}))
sys.exit(1 if bad else 0)
