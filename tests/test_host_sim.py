"""The sequential playout logic the thread-per-playout kernel executes (board_ops.h), run on the
CPU through a test-only harness and compared with the golden vectors.  This is NOT a product
path: the library has no CPU fallback."""
import numpy as np

from util import golden, host_sim, sim_playouts, split_moves


def test_root_mode_golden():
    sim = host_sim()
    for n, P in ((9, 1024), (13, 512), (19, 512)):
        g = golden("root_%d.npz" % n)
        o, d, s = sim_playouts(sim, n, [], 1, None, P, seed_base=int(g["seed_base"]))
        assert (d == g["diffs"][:P]).all()


def test_candidates_golden():
    sim = host_sim()
    g = golden("cands_13_midgame.npz")
    for ci in range(0, len(g["cands"]), 4):
        c = tuple(int(x) for x in g["cands"][ci])
        o, d, s = sim_playouts(sim, 13, g["moves"], int(g["colour"]), c, int(g["P"]), avrg_win=int(g["avrg_win"]),
                               seed_base=int(g["seed_base"]), position_id=int(g["position_id"]), cand_idx=ci)
        assert (d == g["diffs"][ci]).all()
        assert (o[0], o[1], o[2]) == (g["wins"][ci], g["sum_pos"][ci], g["sum_neg"][ci])


def test_midgame_golden():
    sim = host_sim()
    g = golden("midgame_19.npz")
    moves = split_moves(g)
    for bi in range(0, len(moves), 2):
        P = 64
        o, d, s = sim_playouts(sim, 19, moves[bi], int(g["to_move"][bi]), None, P, seed_base=int(g["seed_base"]),
                               position_id=int(g["position_ids"][bi]))
        assert (d == g["diffs"][bi][:P]).all()
