"""Row f3 of SURVEY.md §8: the static pattern marks of MyWorker::work (AgentGo.cpp:104-282) and the
13x13 opening book of MyGame::onMove (:468-572).  Host-only code behind the C ABI, so these run
without a GPU (device = -1) against fixtures generated from the reference's own AgentGo.cpp
(tests/golden/make_golden_genmove.py) and, when oracle/_ref is built, against it live."""
import numpy as np
import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po
from util import golden, split_moves

BONUS_A, BONUS_B = 150 * 1000, 150        # int(numset*index_a), int(numset*index_b): AgentGo.cpp:18-19,85,95-96


def host_engine(moves=(), **kw):
    return ag.Engine(13, device=-1, **kw).replay(moves)


def test_static_units_match_reference_fixture(built_lib):
    g = golden("static_13.npz")
    checked = last_row = 0
    for moves, marks in zip(split_moves(g), g["marks"]):
        e = host_engine(moves)
        for colour in (1, 2):
            for q in np.nonzero(~np.isnan(marks[colour - 1]))[0]:
                ua, ub = e.static_units(colour, int(q) // 13, int(q) % 13)
                assert ua * BONUS_A + ub * BONUS_B == marks[colour - 1][q], (colour, divmod(int(q), 13))
                checked += 1
                last_row += q >= 12 * 13
    assert checked > 2000 and last_row > 100


@pytest.mark.skipif(not po.genmove_ref_available(), reason="oracle/_ref/liboracle_ref_genmove_13.so not built")
def test_static_units_match_reference_live(built_lib):
    """Fresh positions, including full-board endgames where the reads past Board::data matter."""
    for k, length in enumerate((15, 70, 160, 220)):
        moves = po.OracleBoard(13, "reference").random_game(1, length, 7000 + k)
        ref = po.RefGame().replay(moves)
        ob = po.OracleBoard(13, "reference").replay(moves)
        st = ob.export_state()
        e = host_engine(moves)
        for colour in (1, 2):
            for q in range(169):
                i, j = divmod(q, 13)
                if st[16 + q] != 0 or ob.check("suicide", colour, i, j):
                    continue
                ua, ub = e.static_units(colour, i, j)
                assert ua * BONUS_A + ub * BONUS_B == ref.static_mark(colour, i, j), (k, colour, i, j)


def test_static_units_errors(built_lib):
    e = host_engine()
    e.play(1, 3, 3)
    with pytest.raises(ag.AgbError) as ei:
        e.static_units(1, 3, 3)
    assert ei.value.status == -2      # AGB_ERR_OCCUPIED
    with pytest.raises(ag.AgbError):
        e.static_units(1, 13, 0)
    with pytest.raises(ag.AgbError):
        e.static_units(3, 0, 0)


def test_opening_book_matches_reference(built_lib):
    """Steps 1..4 of MyGame::onMove need no playout, so the book answers on a position-only engine."""
    book = golden("genmove_13.npz")["book"]
    e = host_engine(genmove_book=1)
    for s, want in enumerate(book):
        colour = 1 + s % 2
        mv, _ = e.genmove(colour)
        assert mv == tuple(int(v) for v in want)
        assert e.last_marks() is None
        e.play(colour, *mv)
    assert e.genmove_step == 4
    # the fifth call is the Monte-Carlo branch: no device, no fallback
    with pytest.raises(ag.AgbError) as ei:
        e.genmove(1)
    assert ei.value.status == -3      # AGB_ERR_CUDA
    # clear_board resets MyGame::step (onClear, AgentGo.cpp:375-385)
    e.clear_board()
    assert e.genmove_step == 0
    assert e.genmove(1)[0] == (3, 3)


def test_opening_book_region_order(built_lib):
    """Each book point is taken when its region is free, in the reference's order (:537-581)."""
    order = [(3, 3), (9, 9), (3, 9), (9, 3), (3, 5), (9, 7), (7, 3), (5, 9), (6, 6)]
    e = host_engine(genmove_book=1)
    for k, want in enumerate(order):
        e.genmove_step = 0
        assert e.genmove(1 + k % 2)[0] == want
        e.play(1 + k % 2, *want)
    # other sizes have no book: the call goes to the playout path (and fails without a device)
    e9 = ag.Engine(9, device=-1, genmove_book=1)
    with pytest.raises(ag.AgbError):
        e9.genmove(1)


def test_calc_result_matches_reference(built_lib):
    """Row f4 of SURVEY.md §8: the scorer of the reference's match harness (CalcResults,
    GeneHatchery.cpp:393-436), against the fixture and the compiled reference."""
    g = golden("calc_result_13.npz")
    for moves, want in zip(split_moves(g), g["dscore"]):
        assert host_engine(moves).calc_result() == int(want)
    assert host_engine().calc_result() == -169          # empty board: nothing is black (:430)
    if po.genmove_ref_available():
        for k in range(12):
            moves = po.OracleBoard(13, "reference").random_game(1 + k % 2, 30 + 31 * k, 8800 + k)
            assert host_engine(moves).calc_result() == po.RefGame().replay(moves).calc_result()


def test_random_piece_matches_oracle(built_lib, port_lib):
    """agb_random_piece = Board::getRandomPiece on the host (the referee's fallback move,
    GeneHatchery.cpp:474,520) against every oracle available, first move of a seeded stream."""
    from util import oracle_kinds
    for kind in oracle_kinds(13):
        for k in range(40):
            moves = po.OracleBoard(13, kind).random_game(1 + k % 2, [0, 5, 40, 90, 150, 200, 260, 300][k % 8], 6000 + k)
            colour = 1 + (k // 8) % 2
            m = int(po.OracleBoard(13, kind).replay(moves).random_game(colour, 1, 777 + k)[0])
            _, r, c = po.decode_move(m)
            assert host_engine(moves).random_piece(colour, 777 + k) == (None if r is None else (r, c)), (kind, k)
