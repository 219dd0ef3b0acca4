"""agb_genmove with AMAF + static marks + book against the reference's own MyGame::onMove
(AgentGo.cpp:451-692, compiled into oracle/_ref/liboracle_ref_genmove_13.so): same move and the same
total_mark table at the reference's settings (13x13, 150 playouts per candidate, 500 root playouts).

The integer part (every playout's score, the AMAF marks, the static units) is bit-exact and tested
elsewhere; total_mark itself is a double the reference accumulates term by term while the engine
forms it from integer sums, so it is compared to 1e-9 relative."""
import numpy as np
import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po
from util import golden, split_moves

pytestmark = pytest.mark.gpu


def check(e, colour, seed, want_move, want_marks):
    mv, st = e.genmove(colour, 150, 500, int(seed))
    marks = e.last_marks()
    scale = np.abs(want_marks).max()
    assert np.allclose(marks, want_marks, rtol=1e-9, atol=1e-9 * scale), np.abs(marks - want_marks).max()
    assert mv == want_move


def test_genmove_matches_reference_fixture(built_lib):
    g = golden("genmove_13.npz")
    for moves, colour, seed, ref_mv, tm in zip(split_moves(g), g["colours"], g["seeds"], g["ref_moves"], g["total_mark"]):
        e = ag.Engine(13, genmove_amaf=1, genmove_static=1, genmove_book=1).replay(moves)
        e.genmove_step = 10
        check(e, int(colour), seed, tuple(int(v) for v in ref_mv), tm.reshape(13, 13))


@pytest.mark.skipif(not po.genmove_ref_available(), reason="oracle/_ref/liboracle_ref_genmove_13.so not built")
def test_genmove_matches_reference_live(built_lib):
    moves = po.OracleBoard(13, "reference").random_game(2, 64, 31337)
    ref = po.RefGame().replay(moves)
    ref.step = 7
    want_mv, t = ref.genmove(1, 424242)
    e = ag.Engine(13, genmove_amaf=1, genmove_static=1, genmove_book=1).replay(moves)
    e.genmove_step = 7
    check(e, 1, 424242, want_mv, t["total_mark"])


@pytest.mark.skipif(not po.genmove_ref_available(), reason="oracle/_ref/liboracle_ref_genmove_13.so not built")
def test_genmove_with_other_tunables_matches_reference(built_lib):
    """The eight tunables of the reference's command line (AgentGo.cpp:18-25,709-719), set to something
    else than the defaults on both sides: same move, same total_mark table."""
    tun = (700.0, 3.0, 0.5, 35.0, 0.15, 0.45, 0.001, 0.03)
    moves = po.OracleBoard(13, "reference").random_game(1, 70, 2718)
    ref = po.RefGame().replay(moves)
    if not hasattr(ref.lib, "orc_gm_set_tunables"):
        pytest.skip("oracle/_ref was built before the tunables setter existed")
    ref.step = 9
    ref.set_tunables(tun)
    try:
        want_mv, t = ref.genmove(2, 99991)
    finally:
        ref.set_tunables(None)
    e = ag.Engine(13, genmove_amaf=1, genmove_static=1, genmove_book=1, tunables=tun).replay(moves)
    e.genmove_step = 9
    check(e, 2, 99991, want_mv, t["total_mark"])
    # and the defaults are not what was just compared
    e2 = ag.Engine(13, genmove_amaf=1, genmove_static=1, genmove_book=1).replay(moves)
    e2.genmove_step = 9
    e2.genmove(2, 150, 500, 99991)
    assert not np.allclose(e2.last_marks(), t["total_mark"], rtol=1e-6)


def test_game_opening_then_playouts(built_lib):
    """A game from the empty board: four book moves per engine, then the Monte-Carlo branch."""
    e = ag.Engine(13, genmove_amaf=1, genmove_static=1, genmove_book=1)
    seq = []
    for s in range(5):
        mv, st = e.genmove(1, 150, 500, 99 + s)
        seq.append(mv)
        e.play(1, *mv)
        assert (st["playouts"] == 0) == (s < 4)
    assert seq[:4] == [(3, 3), (9, 9), (3, 9), (9, 3)]
    assert e.last_marks() is not None
