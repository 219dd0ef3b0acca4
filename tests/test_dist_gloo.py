"""N>1 path on CPU: world_size 2 over gloo.

Each rank evaluates its shard of the global playout ids (g % world == rank — the kernel's rule,
restated in agentgo_b200.dist.shard_ids) with the oracle, the int64 counters are combined with one
allreduce, and the sum must equal the unsharded evaluation.  This covers the host-side sharding
arithmetic and the collective plumbing; the CUDA kernels apply the same rule on the GPU
(tests/test_gpu_api.py::test_sharded_engines_partition_the_ids)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from agentgo_b200 import dist as agdist
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, P, avrg = 9, 120, 1
    cands = [(2, 2), (4, 4), (6, 5)]
    ob = po.OracleBoard(n, "port")
    ob.put(1, 4, 3)
    # every rank computes all per-playout diffs of ITS ids only
    total = len(cands) * P
    ids = agdist.shard_ids(total, rank, world)
    assert len(ids) == agdist.local_units(total, rank, world)
    w = np.zeros(len(cands), dtype=np.int64); sp = w.copy(); sn = w.copy()
    full = ob.playouts(2, cands, P, avrg_win=avrg, threads=1)       # reference result (diffs per id)
    diffs = full[3].reshape(-1).astype(np.int64)
    for g in ids:
        c = g // P
        v = diffs[g] - avrg
        if v >= 0:
            w[c] += 1; sp[c] += v
        else:
            sn[c] += v
    rw, rsp, rsn = agdist.allreduce_host_counts([w, sp, sn])
    ok = (rw == full[0]).all() and (rsp == full[1]).all() and (rsn == full[2]).all()
    # partial counts differ from the total (the test would be vacuous otherwise)
    partial_differs = not (w == full[0]).all()
    q.put((rank, bool(ok), bool(partial_differs)))
    dist.destroy_process_group()


def test_world_size_2_gloo(port_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]


def _batch_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from agentgo_b200 import dist as agdist
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, P, B, first = 9, 60, 5, 3
    games = [po.OracleBoard(n, "port").random_game(1, 10 + 3 * b, 77 + b) for b in range(B)]
    full = np.zeros((B, 3), dtype=np.int64)
    for b in range(B):
        r = po.OracleBoard(n, "port").replay(games[b]).playouts(1 + len(games[b]) % 2, None, P, position_id=first + b,
                                                                 want_diffs=False, threads=1)
        full[b] = [r[0][0], r[1][0], r[2][0]]
    # this rank's table: [3][stride], its positions in batch order, zero padded (Engine::playouts_batch)
    stride, where = agdist.batch_layout(B, first, world)
    table = np.zeros((3, stride), dtype=np.int64)
    for b, (r, k) in enumerate(where):
        if r == rank:
            table[:, k] = full[b]
    tables = agdist.gather_host_tables(table)
    got = np.array([tables[r][:, k] for r, k in where])
    q.put((rank, bool((got == full).all()), int(stride)))
    dist.destroy_process_group()


def test_batch_gather_world_size_2_gloo(port_lib):
    """Config 3 on N ranks: positions sharded by id, one allgather of the padded tables, every rank
    reassembles the whole [B x 3] table (last sentence)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 3), (1, True, 3)]


def test_shard_rule():
    from agentgo_b200 import dist as agdist
    for total in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            allids = np.concatenate([agdist.shard_ids(total, r, world) for r in range(world)])
            assert sorted(allids.tolist()) == list(range(total))
            assert [agdist.local_units(total, r, world) for r in range(world)] == \
                   [len(agdist.shard_ids(total, r, world)) for r in range(world)]
