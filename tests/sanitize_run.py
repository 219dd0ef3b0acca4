"""Tiny run of both kernels for compute-sanitizer (memcheck / racecheck): 9x9 and 19x19, a few
hundred playouts, root and candidate mode, plus a near-full board that exercises complex mode."""
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
for kern in (ag.KERNEL_WARP, ag.KERNEL_THREAD):
    for n, P in ((9, 96), (19, 48)):
        e = ag.Engine(n, kernel=kern)
        w, sp, sn, d, st = e.playouts(1, None, P)
        e.play(1, 2, 2); e.play(2, 3, 3)
        w2, *_ = e.playouts(2, [(0, 0), (4, 4), (n - 1, n - 1)], P // 3, avrg_win=2)
        print(kern, n, int(w[0]), w2.tolist(), st['moves'])
