"""Short single-kernel run for ncu: 19x19 root-mode playouts (same workload as bench.py, smaller P).
usage: python tests/prof_run.py [warp|thread|group8|group16|group32] [P] [N]"""
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
kern = {'warp': ag.KERNEL_WARP, 'thread': ag.KERNEL_THREAD, 'group8': ag.KERNEL_GROUP8, 'group16': ag.KERNEL_GROUP16,
        'group32': ag.KERNEL_GROUP32}[sys.argv[1] if len(sys.argv) > 1 else 'group8']
P = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 19
e = ag.Engine(n, kernel=kern)
e.playouts(1, None, 2000)
w, sp, sn, d, st = e.playouts(1, None, P)
print('P', P, 'ms %.2f' % st['device_ms'], 'playouts/s %.0f' % (P / st['device_ms'] * 1e3), int(w[0]), int(sp[0]), int(sn[0]))
