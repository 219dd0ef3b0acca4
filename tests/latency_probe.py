"""Ad-hoc: device time of one launch as a function of its size, warp kernel against group kernel
(how AGB_KERNEL_AUTO's threshold was chosen).  usage: python tests/latency_probe.py"""
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
for n in (13, 19):
    eng = {k: ag.Engine(n, kernel=v) for k, v in (('warp', ag.KERNEL_WARP), ('group8', ag.KERNEL_GROUP8))}
    for e in eng.values():
        e.playouts(1, None, 2000)
    for P in (2000, 8000, 16000, 32000, 64000, 128000, 256000, 512000):
        ms = {}
        for k, e in eng.items():
            ms[k] = min(e.playouts(1, None, P, seed_base=7 + r)[4]['device_ms'] for r in range(3))
        print('n %2d P %7d warp %8.3f ms group8 %8.3f ms ratio %.2f' % (n, P, ms['warp'], ms['group8'], ms['warp'] / ms['group8']), flush=True)
