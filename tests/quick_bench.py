"""Ad-hoc throughput probe of both kernels (not the contract bench; see bench.py)."""
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
for kern in (ag.KERNEL_THREAD, ag.KERNEL_WARP):
    for n, P in ((9, 400000), (19, 100000)):
        e = ag.Engine(n, kernel=kern)
        e.playouts(1, None, 4000)
        w, sp, sn, d, st = e.playouts(1, None, P)
        print('kernel', kern, 'n', n, 'P', P, 'ms %.2f' % st['device_ms'],
              'playouts/s %.0f' % (P / st['device_ms'] * 1e3), (w, sp, sn))
