"""Ad-hoc throughput probe of the kernels (not the contract bench; see bench.py).
usage: python tests/quick_bench.py [kernel ...]   kernels: thread warp group8 group16 group32"""
import sys
sys.path.insert(0, '.')
import agentgo_b200 as ag
KERN = {'thread': ag.KERNEL_THREAD, 'warp': ag.KERNEL_WARP, 'group8': ag.KERNEL_GROUP8,
        'group16': ag.KERNEL_GROUP16, 'group32': ag.KERNEL_GROUP32}
names = sys.argv[1:] or ['warp', 'group8', 'group16', 'group32']
ref = {}
for name in names:
    for n, P in ((9, 400000), (19, 1000000 if name != 'thread' else 100000)):
        e = ag.Engine(n, kernel=KERN[name])
        e.playouts(1, None, 4000)
        w, sp, sn, d, st = e.playouts(1, None, P)
        key = (n, P)
        res = (int(w[0]), int(sp[0]), int(sn[0]), st['moves'], st['rng_draws'])
        same = ref.setdefault(key, res) == res
        print('kernel %-8s n %2d P %7d ms %8.2f playouts/s %9.0f %s %s' % (
            name, n, P, st['device_ms'], P / st['device_ms'] * 1e3, res, 'same' if same else 'DIFFERENT'), flush=True)
        e.close()
