"""The oracle against the golden vectors the reference itself produced (tests/golden/).

`port` = oracle/agb_oracle.c (plain-C restatement); `reference` = the reference's own Board.cpp
compiled into oracle/_ref/ (present when oracle/build_ref.sh ran).  Everything is integer and
must match bit for bit.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from util import GOLDEN, golden, oracle_kinds, split_moves


def kinds(n):
    return oracle_kinds(n)


def test_tinymt32_known_answer():
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    # RFC 8682 parameter set, seed 1
    assert kat["tinymt32_seed1"] == [2545341989, 981918433, 3715302833, 2387538352, 3591001365]


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_known_answer_positions(port_lib, kind):
    if kind not in kinds(13):
        pytest.skip("oracle/_ref not built")
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    b = po.OracleBoard(13, kind)
    for c, r, cc in [(2, 1, 0), (2, 1, 1), (2, 0, 1), (1, 2, 0), (1, 2, 1), (1, 2, 2), (1, 1, 2), (1, 0, 2)]:
        b.put(c, r, cc)
    assert [b.check("suicide", 1, 0, 0), b.check("kill", 1, 0, 0), b.check("suicide", 2, 0, 0),
            b.check("true_eye", 2, 0, 0)] == kat["kat1_checks"] == [0, 1, 1, 1]
    b.put(1, 0, 0)
    st = b.export_state()
    snext = st[16 + 6 * 169:16 + 7 * 169]
    lst, i = [], int(st[3])
    for _ in range(6):
        lst.append(i)
        i = int(snext[i])
    assert lst == kat["kat1_after"]["empty_list_head6"] == [1, 14, 13, 3, 4, 5]
    assert [b.count_score(1), b.count_score(2)] == kat["kat1_after"]["score"] == [10, 0]
    b = po.OracleBoard(13, kind)
    assert [b.count_score(1), b.count_score(2)] == [0, 0]
    for r, c in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        b.put(1, r, c)
    assert [b.count_score(1), b.count_score(2)] == kat["kat2_score"] == [6, 0]
    b = po.OracleBoard(13, kind)
    b.put(2, 0, 0); b.put(1, 0, 1); b.put(1, 1, 0)
    st = b.export_state()
    assert int(st[4]) == 1 and [int(x) for x in st[5:8]] == [2, 0, 0] and int(st[3]) == 0


@pytest.mark.parametrize("kind", ["port", "reference"])
@pytest.mark.parametrize("n", [9, 13, 19])
def test_root_mode_matches_golden(port_lib, n, kind):
    if kind not in kinds(n):
        pytest.skip("oracle/_ref not built")
    g = golden("root_%d.npz" % n)
    P = 1024                      # prefix of the golden run: seeds depend only on the playout index
    w, sp, sn, d, st = po.OracleBoard(n, kind).playouts(1, None, P, seed_base=int(g["seed_base"]),
                                                        threads=0, with_stats=True)
    assert (d[0] == g["diffs"][:P]).all()
    dd = g["diffs"][:P].astype(np.int64)
    assert w[0] == (dd >= 0).sum() and sp[0] == dd[dd >= 0].sum() and sn[0] == dd[dd < 0].sum()


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_candidates_match_golden(port_lib, kind):
    for name in ("cands_9.npz", "cands_13_midgame.npz"):
        g = golden(name)
        n = int(g["n"])
        if kind not in kinds(n):
            pytest.skip("oracle/_ref not built")
        mv = g["moves"] if "moves" in g.files else []
        pid = int(g["position_id"]) if "position_id" in g.files else 0
        cands = [tuple(int(x) for x in rc) for rc in g["cands"]][::5]
        ob = po.OracleBoard(n, kind).replay(mv)
        w, sp, sn, d = ob.playouts(int(g["colour"]), [tuple(int(x) for x in rc) for rc in g["cands"]],
                                   int(g["P"]), avrg_win=int(g["avrg_win"]), seed_base=int(g["seed_base"]),
                                   position_id=pid, threads=0)
        assert (d == g["diffs"]).all()
        assert (w == g["wins"]).all() and (sp == g["sum_pos"]).all() and (sn == g["sum_neg"]).all()
        assert len(cands) > 0


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_midgame_positions_match_golden(port_lib, kind):
    g = golden("midgame_19.npz")
    if kind not in kinds(19):
        pytest.skip("oracle/_ref not built")
    moves = split_moves(g)
    for bi in range(0, len(moves), 3):
        ob = po.OracleBoard(19, kind).replay(moves[bi])
        assert (ob.export_state() == g["states"][bi]).all()
        assert [ob.count_score(1), ob.count_score(2)] == list(g["scores"][bi])
        tab = g["checks"][bi]
        for which in range(8):
            for c in (1, 2):
                for pt in range(0, 361, 3):
                    assert ob.check(which, c, pt // 19, pt % 19) == tab[which, c - 1, pt]
        P = 64
        w, sp, sn, d = ob.playouts(int(g["to_move"][bi]), None, P, seed_base=int(g["seed_base"]),
                                   position_id=int(g["position_ids"][bi]), threads=0)
        assert (d[0] == g["diffs"][bi][:P]).all()


def test_port_equals_reference_live(port_lib):
    """State after every few moves of random games, all predicates, and move sequences."""
    for n in (9, 13, 19):
        if "reference" not in kinds(n):
            pytest.skip("oracle/_ref not built")
        for gi in range(6):
            mv = po.OracleBoard(n, "reference").random_game(1 + gi % 2, [n * n // 3, n * n, 2 * n * n][gi % 3], 500 + gi)
            mv2 = po.OracleBoard(n, "port").random_game(1 + gi % 2, len(mv), 500 + gi)
            assert (mv == mv2).all()
            r, p = po.OracleBoard(n, "reference"), po.OracleBoard(n, "port")
            for i, m in enumerate(mv):
                r.replay([m]); p.replay([m])
                if i % 11 == 0 or i == len(mv) - 1:
                    assert (r.export_state() == p.export_state()).all()
            st = r.export_state()
            for which in range(8):
                for c in (1, 2):
                    for pt in range(n * n):
                        if st[16 + pt] == 0 or which >= 5:
                            assert r.check(which, c, pt // n, pt % n) == p.check(which, c, pt // n, pt % n)


def test_thread_count_does_not_change_results(port_lib):
    b = po.OracleBoard(9, "port")
    a1 = b.playouts(1, [(2, 2), (4, 4)], 300, avrg_win=1, threads=1)
    a7 = b.playouts(1, [(2, 2), (4, 4)], 300, avrg_win=1, threads=7)
    for x, y in zip(a1, a7):
        assert (x == y).all()


def test_amaf_marks_port_equals_reference(port_lib):
    """AMAF first-play marks (AgentGo.cpp:286-338): the C restatement against the compiled reference."""
    for n in (9, 13):
        if "reference" not in kinds(n):
            pytest.skip("oracle/_ref not built")
        mv = po.OracleBoard(n, "reference").random_game(1, n * n // 3, 321)
        r = po.OracleBoard(n, "reference").replay(mv)
        p = po.OracleBoard(n, "port").replay(mv)
        st = r.export_state()
        emp = [i for i in range(n * n) if st[16 + i] == 0][::7][:5]
        cands = [(i // n, i % n) for i in emp]
        A = r.playouts_amaf(2, cands, 50, avrg_win=1, seed_base=5)
        B = p.playouts_amaf(2, cands, 50, avrg_win=1, seed_base=5)
        for x, y in zip(A, B):
            assert (x == y).all()
        plain = r.playouts(2, cands, 50, avrg_win=1, seed_base=5, threads=1)
        for i in range(3):
            assert (A[i] == plain[i]).all()
        # every playout contributes w to at most one table per point, steps 2..40 only
        assert A[3][:, :, :2, :].sum() == 0
