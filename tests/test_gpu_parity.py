"""GPU parity: CUDA playouts through the C ABI vs the oracle (bit-exact, integers)."""
import numpy as np
import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

KERNELS = [ag.KERNEL_THREAD, ag.KERNEL_WARP, ag.KERNEL_GROUP8, ag.KERNEL_GROUP16, ag.KERNEL_GROUP32]


def oracle_board(n, moves=()):
    return po.OracleBoard(n, "best").replay(moves)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [9, 13, 19])
def test_root_mode_empty_board(built_lib, n, kernel):
    P = 2048 if n < 19 else 1024
    e = ag.Engine(n, kernel=kernel)
    w, sp, sn, d, st = e.playouts(ag.BLACK, None, P, want_diffs=True)
    ow, osp, osn, od, ost = oracle_board(n).playouts(1, None, P, threads=0, with_stats=True)
    assert (d == od).all(), np.nonzero(d != od)
    assert (w, sp, sn) == (ow, osp, osn)
    assert st["moves"] == ost["moves"] and st["rng_draws"] == ost["rng_draws"]
    assert st["playouts"] == P and st["faults"] == 0


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("n", [9, 19])
def test_candidates_midgame(built_lib, n, kernel):
    mv = oracle_board(n).random_game(1, n * n // 2, 4242)
    ob = oracle_board(n, mv)
    st = ob.export_state()
    empties = [i for i in range(n * n) if st[16 + i] == 0]
    cands = [(i // n, i % n) for i in empties[::max(1, len(empties) // 12)]]
    P = 96
    e = ag.Engine(n, kernel=kernel).replay(mv)
    w, sp, sn, d, _ = e.playouts(ag.WHITE, cands, P, avrg_win=3, seed_base=99, position_id=7, want_diffs=True)
    ow, osp, osn, od = ob.playouts(2, cands, P, avrg_win=3, seed_base=99, position_id=7, threads=0)
    assert (d == od).all()
    assert (w == ow).all() and (sp == osp).all() and (sn == osn).all()
