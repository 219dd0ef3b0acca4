"""GPU tests of the rest of the C ABI: batch, genmove, sharding, edge cases, full-size parity."""
import numpy as np
import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po
from util import golden, split_moves

pytestmark = pytest.mark.gpu

KERNELS = [ag.KERNEL_THREAD, ag.KERNEL_WARP, ag.KERNEL_GROUP8, ag.KERNEL_GROUP16, ag.KERNEL_GROUP32]


@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_root_prefix(built_lib, kernel):
    for n in (9, 13, 19):
        g = golden("root_%d.npz" % n)
        P = int(g["P"])
        w, sp, sn, d, st = ag.Engine(n, kernel=kernel).playouts(1, None, P, seed_base=int(g["seed_base"]), want_diffs=True)
        assert (d[0] == g["diffs"]).all()
        assert (w[0], sp[0], sn[0]) == (g["wins"][0], g["sum_pos"][0], g["sum_neg"][0])
        assert st["moves"] == int(g["moves"]) and st["rng_draws"] == int(g["rng_draws"])


@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_candidates(built_lib, kernel):
    for name in ("cands_9.npz", "cands_13_midgame.npz"):
        g = golden(name)
        e = ag.Engine(int(g["n"]), kernel=kernel)
        if "moves" in g.files:
            e.replay(g["moves"])
        pid = int(g["position_id"]) if "position_id" in g.files else 0
        w, sp, sn, d, _ = e.playouts(int(g["colour"]), [tuple(int(x) for x in rc) for rc in g["cands"]], int(g["P"]),
                                     avrg_win=int(g["avrg_win"]), seed_base=int(g["seed_base"]), position_id=pid,
                                     want_diffs=True)
        assert (d == g["diffs"]).all()
        assert (w == g["wins"]).all() and (sp == g["sum_pos"]).all() and (sn == g["sum_neg"]).all()


@pytest.mark.parametrize("kernel", KERNELS)
def test_golden_midgame_batch(built_lib, kernel):
    """config 3 subset: 16 mid-game 19x19 positions through agb_playouts_batch."""
    g = golden("midgame_19.npz")
    moves = split_moves(g)
    ids = [int(x) for x in g["position_ids"]]
    P = int(g["P"])
    e = ag.Engine(19, kernel=kernel)
    for bi in range(len(moves)):     # position ids are not contiguous: one call per position id
        w, sp, sn, d, _ = e.playouts_batch([(moves[bi], int(g["to_move"][bi]))], P, seed_base=int(g["seed_base"]),
                                           first_position_id=ids[bi], want_diffs=True)
        assert (d[0] == g["diffs"][bi]).all(), bi
        assert (w[0], sp[0], sn[0]) == tuple(g["sums"][bi])


def test_batch_many_positions_vs_oracle(built_lib):
    n, B, P = 19, 24, 400
    pos = []
    for k in range(B):
        mv = po.OracleBoard(n, "best").random_game(1, 30 + 9 * k, 7000 + k)
        pos.append((mv, 1 + len(mv) % 2))
    w, sp, sn, d, st = ag.Engine(n).playouts_batch(pos, P, seed_base=5, first_position_id=100, want_diffs=True)
    assert st["playouts"] == B * P
    for k in (0, 7, 23):
        ow, osp, osn, od = po.OracleBoard(n, "best").replay(pos[k][0]).playouts(pos[k][1], None, P, seed_base=5,
                                                                                position_id=100 + k, threads=0)
        assert (d[k] == od[0]).all() and (w[k], sp[k], sn[k]) == (ow[0], osp[0], osn[0])


def test_sharded_engines_partition_the_ids(built_lib):
    n, P = 13, 600
    cands = [(3, 3), (6, 6), (9, 2)]
    full = ag.Engine(n).playouts(ag.BLACK, cands, P, avrg_win=-1, want_diffs=True)
    acc = [np.zeros(3, dtype=np.int64) for _ in range(3)]
    merged = np.full((3, P), -32768, dtype=np.int16)
    for r in range(3):
        e = ag.Engine(n, shard_rank=r, shard_count=3)
        w, sp, sn, d, st = e.playouts(ag.BLACK, cands, P, avrg_win=-1, want_diffs=True)
        assert st["playouts"] == len(range(r, 3 * P, 3))
        for a, x in zip(acc, (w, sp, sn)):
            a += x
        mine = d.reshape(-1)[r::3]
        assert (mine != -32768).all()
        other = np.delete(d.reshape(-1), np.arange(r, 3 * P, 3))
        assert (other == -32768).all()          # entries of other shards are left untouched
        merged.reshape(-1)[r::3] = mine
    assert (merged == full[3]).all()
    for a, x in zip(acc, full[:3]):
        assert (a == x).all()


def test_allreduce_hook_is_called_on_the_counters(built_lib):
    import ctypes
    calls = []
    e = ag.Engine(9, shard_rank=0, shard_count=2)
    e.set_allreduce(lambda ptr, count, stream: calls.append((ptr, count, stream)) or 0)
    e.playouts(ag.BLACK, [(4, 4), (2, 2)], 50)
    assert len(calls) == 1 and calls[0][1] == 6 and calls[0][0] == e.counters_dev() and calls[0][2] == e.stream()


def test_edge_cases(built_lib):
    e = ag.Engine(9)
    # P = 1
    w, sp, sn, d, st = e.playouts(ag.WHITE, None, 1, want_diffs=True)
    od = po.OracleBoard(9, "best").playouts(2, None, 1, threads=1)[3]
    assert d[0, 0] == od[0, 0] and st["playouts"] == 1
    # occupied candidate is refused
    e.play(ag.BLACK, 4, 4)
    with pytest.raises(ag.AgbError) as ex:
        e.playouts(ag.WHITE, [(4, 4)], 10)
    assert ex.value.status == -2
    with pytest.raises(ag.AgbError):
        e.playouts(ag.WHITE, [(9, 0)], 10)
    with pytest.raises(ag.AgbError):
        e.playouts(3, None, 10)
    # nearly full board with a ko on the root: game ends quickly, passes happen
    n = 9
    mv = po.OracleBoard(n, "best").random_game(1, 3 * n * n, 31337)
    ob = po.OracleBoard(n, "best").replay(mv)
    e2 = ag.Engine(n).replay(mv)
    for kern in KERNELS:
        e3 = ag.Engine(n, kernel=kern).replay(mv)
        w, sp, sn, d, _ = e3.playouts(1, None, 300, want_diffs=True)
        assert (d == ob.playouts(1, None, 300, threads=0)[3]).all()
    assert (e2.export_state() == ob.export_state()).all()


def _oracle_genmove(ob, n, colour, P_cand, P_root, seed_base, use_amaf=False, moves=()):
    w, sp, sn, _ = ob.playouts(colour, None, P_root, seed_base=seed_base, want_diffs=False, threads=0)
    tot = int(sp[0] + sn[0])
    avrg = abs(tot) // P_root * (1 if tot >= 0 else -1)          # C truncation, AgentGo.cpp:447
    st = ob.export_state()
    cands = []
    for i in range(n):
        for j in range(n):
            if (ob.check("true_eye", colour, i, j) or st[16 + i * n + j] != 0 or ob.check("suicide", colour, i, j)
                    or ob.check("no_sense", colour, i, j) or ob.check("compete", colour, i, j)):
                continue
            cands.append((i, j))
    marks = {}
    amaf12 = np.zeros(n * n)
    if cands:
        cnsv = 2.0 ** (avrg * 0.01)
        if use_amaf:
            w, sp, sn, Ac = ob.playouts_amaf(colour, cands, P_cand, avrg_win=avrg, seed_base=seed_base + 1,
                                             per_candidate=True)
            for (ci, cj), A in zip(cands, Ac):
                # the decay table is recomputed per candidate from the stones AFTER it is played
                # (AgentGo.cpp:84-94): a candidate that captures sees fewer
                after = po.OracleBoard(n, "best").replay(moves)
                after.put(colour, ci, cj)
                sa = after.export_state()
                num_piece = int(sa[1] + sa[2])
                num_a = 1.0 / (40.0 - (39.0 / 180) * num_piece)
                num_b = 0.3 - (0.3 / 180) * num_piece
                for s in range(2, 41):
                    k = s - 1
                    base = num_a * k
                    v = 1 - (base ** num_b if base >= 0 else float("nan"))
                    v = 0.0 if not (v >= 0) else min(v, 1.0)
                    amaf12 += 100.0 * v * ((A[0, 0, s] + cnsv * A[0, 1, s]) - (A[1, 0, s] + cnsv * A[1, 1, s]))
        else:
            w, sp, sn, _ = ob.playouts(colour, cands, P_cand, avrg_win=avrg, seed_base=seed_base + 1, want_diffs=False, threads=0)
        for c, a, b in zip(cands, sp, sn):
            marks[c] = 2.0 * 100.0 * (float(a) + cnsv * float(b))
    best, arg = None, None
    for i in range(n):
        for j in range(n):
            if st[16 + i * n + j] == 0 and not ob.check("true_eye", colour, i, j) and \
                    not ob.check("suicide", colour, i, j) and not ob.check("compete", colour, i, j):
                m = marks.get((i, j), 0.0) + amaf12[i * n + j]
                if best is None or m > best:
                    best, arg = m, (i, j)
    return arg      # also without candidates: the first eligible point competes with mark 0 (AgentGo.cpp:647-676)


def test_genmove_matches_oracle_restatement(built_lib):
    for n, seed in ((9, 11), (13, 12)):
        mv = po.OracleBoard(n, "best").random_game(1, n * n // 2, seed)
        colour = 1 + len(mv) % 2
        ob = po.OracleBoard(n, "best").replay(mv)
        e = ag.Engine(n).replay(mv)
        got, st = e.genmove(colour, 60, 200, seed_base=777)
        assert got == _oracle_genmove(ob, n, colour, 60, 200, 777)
        assert st["playouts"] > 200
        # with the AMAF terms of AgentGo.cpp:329-338,654-655
        e2 = ag.Engine(n, genmove_amaf=1).replay(mv)
        got2, _ = e2.genmove(colour, 60, 200, seed_base=777)
        assert got2 == _oracle_genmove(ob, n, colour, 60, 200, 777, use_amaf=True, moves=mv)


def test_full_size_config2_bit_exact(built_lib):
    """BASELINE.json configs[1]: 19x19 empty board, 1 000 000 root-mode playouts, every per-playout
    diff and the three counters against the oracle (the compiled reference when present)."""
    P = 1000000
    w, sp, sn, d, st = ag.Engine(19).playouts(ag.BLACK, None, P, want_diffs=True)
    assert st["playouts"] == P and st["faults"] == 0
    # size-independent properties
    dd = d[0].astype(np.int64)
    assert w[0] == (dd >= 0).sum() and sp[0] == dd[dd >= 0].sum() and sn[0] == dd[dd < 0].sum()
    assert np.abs(dd).max() <= 361
    g = golden("root_19.npz")
    assert (d[0][:int(g["P"])] == g["diffs"]).all()
    ow, osp, osn, od, ost = po.OracleBoard(19, "best").playouts(1, None, P, threads=0, with_stats=True)
    assert (d == od).all()
    assert (w[0], sp[0], sn[0]) == (ow[0], osp[0], osn[0])
    assert st["moves"] == ost["moves"] and st["rng_draws"] == ost["rng_draws"]


@pytest.mark.parametrize("kernel", [ag.KERNEL_WARP, ag.KERNEL_GROUP8])
def test_amaf_marks_match_oracle(built_lib, kernel):
    """'next' row 1: AMAF first-play marks, integer contract of agb_playouts_amaf, by both
    kernels that record them (the warp kernel and the group kernel with 8 lanes per playout)."""
    for n, P in ((9, 200), (13, 120), (19, 80)):
        mv = po.OracleBoard(n, "best").random_game(1, n * n // 3, 321)
        ob = po.OracleBoard(n, "best").replay(mv)
        st = ob.export_state()
        emp = [i for i in range(n * n) if st[16 + i] == 0][::5][:9]
        cands = [(i // n, i % n) for i in emp]
        ow, osp, osn, oa = ob.playouts_amaf(2, cands, P, avrg_win=1, seed_base=5, position_id=2)
        e = ag.Engine(n, kernel=kernel).replay(mv)
        w, sp, sn, a, stt = e.playouts_amaf(2, cands, P, avrg_win=1, seed_base=5, position_id=2)
        assert (w == ow).all() and (sp == osp).all() and (sn == osn).all()
        assert (a == oa).all(), np.argwhere(a != oa)[:5]
        assert a[:, :, :2, :].sum() == 0 and np.count_nonzero(a) > 100
        # same counters as the plain evaluation
        w2, sp2, sn2, _, _ = e.playouts(2, cands, P, avrg_win=1, seed_base=5, position_id=2)
        assert (w2 == w).all() and (sp2 == sp).all() and (sn2 == sn).all()
    with pytest.raises(ag.AgbError):
        ag.Engine(9, kernel=ag.KERNEL_THREAD).playouts_amaf(1, [(1, 1)], 10)


@pytest.mark.parametrize("kernel", KERNELS)
def test_move_cap_is_reported_as_fault(built_lib, kernel):
    """The per-playout move cap is a fault detector only (the reference has no cap): hitting it
    must surface as AGB_ERR_FAULT, never as a silent wrong score."""
    e = ag.Engine(9, kernel=kernel, max_moves=10)
    with pytest.raises(ag.AgbError) as ex:
        e.playouts(ag.BLACK, None, 64)
    assert ex.value.status == -5
    # the engine stays usable afterwards
    e2 = ag.Engine(9, kernel=kernel)
    assert e2.playouts(ag.BLACK, None, 64)[4]["faults"] == 0


def test_ragged_batch_and_wide_candidate_sets(built_lib):
    """Ragged inputs and the widest candidate set: a batch mixing an empty move list, a list with
    passes and long games; every empty point of a 19x19 position as a candidate with P = 1 and 2."""
    n = 19
    games = [np.zeros(0, dtype=np.uint32),
             np.array([ag.move(1, 3, 3), ag.move_pass(2), ag.move(1, 15, 15), ag.move_pass(2), ag.move_pass(1)], dtype=np.uint32),
             po.OracleBoard(n, "best").random_game(1, 1, 5),
             po.OracleBoard(n, "best").random_game(2, 333, 6),
             po.OracleBoard(n, "best").random_game(1, 700, 7)]
    pos = [(mv, 1 + k % 2) for k, mv in enumerate(games)]
    e = ag.Engine(n)
    P = 200
    w, sp, sn, d, st = e.playouts_batch(pos, P, seed_base=4242, first_position_id=10, want_diffs=True)
    assert st["playouts"] == len(pos) * P and st["faults"] == 0
    for k, (mv, colour) in enumerate(pos):
        ob = po.OracleBoard(n, "best").replay(mv)
        ow, osp, osn, od = ob.playouts(colour, None, P, seed_base=4242, position_id=10 + k, threads=0)
        assert (d[k] == od[0]).all(), k
        assert (w[k], sp[k], sn[k]) == (ow[0], osp[0], osn[0])
    # every empty point of a mid-game position as a candidate (suicides included: the candidate
    # stone is placed by the host's put, which handles self-capture like the reference)
    mv = po.OracleBoard(n, "best").random_game(2, 150, 8)
    ob = po.OracleBoard(n, "best").replay(mv)
    st0 = ob.export_state()
    cands = [(i // n, i % n) for i in range(n * n) if st0[16 + i] == 0]
    assert len(cands) > 100
    e = ag.Engine(n).replay(mv)
    for P in (1, 2):
        w, sp, sn, d, st = e.playouts(ag.WHITE, cands, P, avrg_win=-2, seed_base=9, want_diffs=True)
        ow, osp, osn, od = ob.playouts(2, cands, P, avrg_win=-2, seed_base=9, threads=0)
        assert (d == od).all() and (w == ow).all() and (sp == osp).all() and (sn == osn).all()


def test_batch_gather_on_two_gpus(built_lib):
    """Config 3 on two GPUs of one box (agb_attach_nccl: positions sharded by id, one ncclAllGather of the
    per-position table): every engine returns the table a single engine computes."""
    import ctypes as C
    import threading
    l = ag.lib()
    if l.agb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden("midgame_19.npz")
    pos = [(mv, int(g["to_move"][i])) for i, mv in enumerate(split_moves(g))][:7]
    P = 256
    ref = ag.Engine(19).playouts_batch(pos, P, seed_base=5, first_position_id=3)
    engines = [ag.Engine(19, device=i, shard_rank=i, shard_count=2) for i in range(2)]
    arr = (C.c_void_p * 2)(*[e._h for e in engines])
    grp = C.c_void_p()
    assert l.agb_attach_nccl(arr, 2, C.byref(grp)) == 0
    out = [None, None]

    def run(i):
        out[i] = engines[i].playouts_batch(pos, P, seed_base=5, first_position_id=3)
    ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    l.agb_detach_nccl(grp)
    for i in range(2):
        assert (out[i][0] == ref[0]).all() and (out[i][1] == ref[1]).all() and (out[i][2] == ref[2]).all()
        assert out[i][4]["playouts"] == P * len([b for b in range(7) if (3 + b) % 2 == i])
