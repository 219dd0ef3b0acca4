"""Host-side replay board of the product (position-only engine, no GPU needed) against the
golden states/predicates and the live oracle."""
import numpy as np
import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po
from util import golden, oracle_kinds, split_moves


def test_golden_states_and_predicates(built_lib):
    g = golden("midgame_19.npz")
    moves = split_moves(g)
    e = ag.Engine(19, device=-1)
    for bi, mv in enumerate(moves):
        e.clear_board()
        e.replay(mv)
        assert (e.export_state() == g["states"][bi]).all()
        assert [e.count_score(1), e.count_score(2)] == list(g["scores"][bi])
        tab = g["checks"][bi]
        board = e.board().reshape(-1)
        for which in range(8):
            for c in (1, 2):
                for pt in range(361):
                    if board[pt] == 0 or which >= 5:   # 4-neighbour predicates are only asked of empty points
                        assert e.check(which, c, pt // 19, pt % 19) == tab[which, c - 1, pt], (bi, which, c, pt)


@pytest.mark.parametrize("n", [9, 13, 19])
def test_replay_matches_oracle_move_by_move(built_lib, port_lib, n):
    kind = oracle_kinds(n)[-1]
    e = ag.Engine(n, device=-1)
    for gi in range(8):
        mv = po.OracleBoard(n, kind).random_game(1 + gi % 2, [n * n // 2, n * n, 2 * n * n][gi % 3], 900 + gi)
        ob = po.OracleBoard(n, kind)
        e.clear_board()
        for i, m in enumerate(mv):
            ob.replay([m]); e.replay([m])
            if i % 5 == 0 or i == len(mv) - 1:
                assert (ob.export_state() == e.export_state()).all(), (n, gi, i)


def test_error_behaviour(built_lib):
    e = ag.Engine(9, device=-1)
    e.play(ag.BLACK, 0, 0)
    with pytest.raises(ag.AgbError) as ex:
        e.play(ag.WHITE, 0, 0)            # reference: AG_PANIC (Board.cpp:75-78)
    assert ex.value.status == -2
    with pytest.raises(ag.AgbError) as ex:
        e.play(ag.WHITE, 9, 0)            # Piece::legal (Piece.cpp:28-30)
    assert ex.value.status == -1
    with pytest.raises(ag.AgbError) as ex:
        e.play(3, 1, 1)
    assert ex.value.status == -1
    with pytest.raises(ag.AgbError):
        e.boardsize(10)
    e.boardsize(13)
    assert e.board().shape == (13, 13) and e.board().sum() == 0
    # no CPU fallback: compute calls fail loudly on a position-only engine
    with pytest.raises(ag.AgbError) as ex:
        e.playouts(ag.BLACK, None, 10)
    assert ex.value.status == -3
    with pytest.raises(ag.AgbError):
        ag.Engine(11, device=-1)


def test_pass_clears_ko_and_last_move(built_lib):
    e = ag.Engine(13, device=-1)
    e.play(2, 0, 0); e.play(1, 0, 1); e.play(1, 1, 0)     # KAT3: single-stone capture
    st = e.export_state()
    assert st[4] == 1 and list(st[5:8]) == [2, 0, 0]
    assert e.check("compete", 2, 0, 0) == 1 and e.check("compete", 1, 0, 0) == 0
    e.pass_(2)                                             # Board::pass, Board.cpp:200-203
    st = e.export_state()
    assert st[4] == 0 and list(st[11:14]) == [0, -1, -1]
    assert e.check("compete", 2, 0, 0) == 0
