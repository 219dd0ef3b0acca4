"""Row f4 of SURVEY.md §8: the match referee over PROCESSES.  `agentgo_gtp --match` starts the two
players as child processes and speaks GTP to them over pipes, like the reference's GA tuner does
(GeneHatchery.cpp:221-281 pipes and processes, :337-375 Genmove / IsReplyOK, :438-568
SimulateOneGame), so any GTP engine can be a player.  Here both players are scripted engines
(tests/gtp_script_engine.py) and the referee's game — every move, every substituted move, the end of
the game and the score — is compared with a restatement of SimulateOneGame on the oracle's board.
No GPU involved: the referee only keeps a board."""
import json
import os
import subprocess
import sys

import pytest

import agentgo_b200 as ag
from oracle import pyoracle as po

GTP = os.path.join(ag.HERE, "agentgo_gtp")
SCRIPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gtp_script_engine.py")
LETTERS = "ABCDEFGHJKLMNOPQRST"


class Scripted:
    """the engine of tests/gtp_script_engine.py, in process"""

    def __init__(self, n, stride, passes):
        self.n, self.stride, self.passes, self.seen, self.k = n, stride, set(passes), set(), 0

    def saw(self, mv):
        self.seen.add(mv)

    def genmove(self):
        mv = None
        if self.k not in self.passes:
            for i in range(self.n * self.n):
                q = (i * self.stride) % (self.n * self.n)
                if (q // self.n, q % self.n) not in self.seen:
                    mv = (q // self.n, q % self.n)
                    break
        self.k += 1
        if mv is not None:
            self.seen.add(mv)
        return mv


def simulate_one_game(n, players, seed, game, max_moves=800):
    """SimulateOneGame, GeneHatchery.cpp:438-568, on the oracle's board.  -> (log, moves played)"""
    bd = po.OracleBoard(n, "port")
    log, moves = [], []
    played = [True, True]           # bplay / lastwplay (:461,:511)
    for t in range(2 * max_moves):
        c, o = t & 1, 1 - (t & 1)
        colour = c + 1
        mv = players[c].genmove()                                       # Genmove, :337-367
        moved = mv is not None
        if moved:
            bd.put(colour, *mv)                                         # :477 / :513
            players[o].saw(mv)                                          # play <c> <vertex> to the other side
        else:
            # the referee's own policy move (:484,:520): getRandomPiece on its board
            m = int(bd.random_game(colour, 1, (seed + 977 * game + t) & 0xFFFFFFFF)[0])
            _, r, cc = po.decode_move(m)
            if r is None:
                if not played[o]:
                    break                                               # :486-491,:522-527: the game is over
                # (random_game has passed on the oracle's board already, like bd.pass at :494,:530)
            else:
                mv = (r, cc)                                            # random_game has put the stone
                players[c].saw(mv); players[o].saw(mv)
                moved = True
        played[c] = moved
        log.append("%3d %s %s" % (t + 1, "bw"[c], "pass" if not moved else "%s%d" % (LETTERS[mv[0]], mv[1] + 1)))
        if moved:
            moves.append(po.move(colour, *mv))
        else:
            moves.append(po.move_pass(colour))
    return log, moves


@pytest.mark.parametrize("n,strides,passes", [(13, (7, 11), ((3, 4, 20), (5, 6, 7))), (9, (5, 7), ((0, 9), (2,)))])
def test_referee_over_pipes_follows_simulate_one_game(built_lib, port_lib, n, strides, passes):
    cmd = lambda w: "%s %s %d %d %s" % (sys.executable, SCRIPT, n, strides[w], ",".join(str(x) for x in passes[w]))
    r = subprocess.run([GTP, "--device", "-1", "--boardsize", str(n), "--match", "1", "--seed", "41",
                        "--black-cmd", cmd(0), "--white-cmd", cmd(1)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got_log = [l for l in r.stderr.splitlines() if l[:3].strip().isdigit()]
    m = json.loads(r.stdout.strip().splitlines()[-1])["match"]
    want_log, moves = simulate_one_game(n, [Scripted(n, strides[0], passes[0]), Scripted(n, strides[1], passes[1])], 41, 0)
    assert got_log == want_log
    assert any(l.endswith("pass") for l in got_log) and len(got_log) > n * n // 2
    assert m["players"] == "child processes over pipes" and m["stones_played"] == sum(not l.endswith("pass") for l in got_log)
    # the score: CalcResults of the reference itself where it is compiled (13x13), else of the host board
    if n == 13 and po.genmove_ref_available():
        want = po.RefGame().replay([x for x in moves if po.decode_move(int(x))[1] is not None]).calc_result()
        assert m["dscore"] == [want]
    host = ag.Engine(n, device=-1).replay(moves)
    assert m["dscore"] == [host.calc_result()]


def test_referee_stops_on_a_bad_reply(built_lib):
    """IsReplyOK / Genmove throw on a reply that is not '=' (GeneHatchery.cpp:360,375): the match ends
    with an error instead of going on."""
    r = subprocess.run([GTP, "--device", "-1", "--boardsize", "9", "--match", "1", "--black-cmd",
                        "%s %s 9 5 ''" % (sys.executable, SCRIPT), "--white-cmd", "sed -u 's/.*/? no/;G'"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "bad reply" in r.stderr
