"""Generate tests/golden/static_13.npz and genmove_13.npz from the REFERENCE ITSELF.

Runs the reference's own move selection (AgentGo.cpp:15-705 compiled by oracle/build_ref.sh into
oracle/_ref/liboracle_ref_genmove_13.so, serial schedule, injected per-playout TinyMT32) and stores
  * static_13.npz  — the static pattern marks of MyWorker::work (AgentGo.cpp:104-282) for every
                     empty, non-suicide point of 12 positions, both colours;
  * genmove_13.npz — MyGame::onMove on 6 positions: the move and the whole total_mark table, plus
                     the opening-book sequence of a game's first eight genmoves;
  * calc_result_13.npz — CalcResults of the match harness (GeneHatchery.cpp:393-436) on 24 positions.
Re-run:  python tests/golden/make_golden_genmove.py   (about a minute)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402

N = 13
SEED_BASE = 20151001


def positions(count, lens, seed0):
    out = []
    for k in range(count):
        out.append(po.OracleBoard(N, "reference").random_game(1 + k % 2, lens[k % len(lens)], seed0 + k))
    return out


def static_marks(moves):
    g = po.RefGame().replay(moves)
    ob = po.OracleBoard(N, "reference").replay(moves)
    st = ob.export_state()
    out = np.full((2, N * N), np.nan)
    for colour in (1, 2):
        for i in range(N):
            for j in range(N):
                if st[16 + i * N + j] != 0 or ob.check("suicide", colour, i, j):
                    continue
                out[colour - 1, i * N + j] = g.static_mark(colour, i, j)
    return out


def calc_results():
    ps = positions(24, [5, 40, 90, 150, 200, 260, 330, 400], 4000)
    np.savez_compressed(os.path.join(HERE, "calc_result_13.npz"),
                        moves=np.concatenate(ps).astype(np.uint32), lens=np.array([len(p) for p in ps]),
                        dscore=np.array([po.RefGame().replay(p).calc_result() for p in ps]))


def main():
    calc_results()
    if "--calc-only" in sys.argv:
        return
    ps = positions(12, [6, 25, 50, 80, 110, 140], 500)
    np.savez_compressed(os.path.join(HERE, "static_13.npz"),
                        moves=np.concatenate(ps).astype(np.uint32), lens=np.array([len(p) for p in ps]),
                        marks=np.stack([static_marks(p) for p in ps]))
    gp = positions(6, [8, 24, 48, 72, 100, 130], 900)
    colours, moves_out, marks = [], [], []
    for k, p in enumerate(gp):
        colour = 1 + k % 2
        g = po.RefGame().replay(p)
        g.step = 10
        mv, t = g.genmove(colour, SEED_BASE + k)
        colours.append(colour)
        moves_out.append(mv if mv else (-1, -1))
        marks.append(t["total_mark"].reshape(-1))
        print("position", k, "colour", colour, "->", mv)
    # opening book: the engine plays both sides from the empty board; steps 1..4 come from the
    # book on each side's own counter in the reference (one MyGame per process); here one MyGame
    # answers eight genmoves in a row, so the first four are book moves.
    g = po.RefGame()
    book = []
    for s in range(4):
        colour = 1 + s % 2
        mv, _ = g.genmove(colour, SEED_BASE)
        book.append(mv)
        g.play(colour, *mv)
    np.savez_compressed(os.path.join(HERE, "genmove_13.npz"),
                        moves=np.concatenate(gp).astype(np.uint32), lens=np.array([len(p) for p in gp]),
                        colours=np.array(colours), seeds=np.array([SEED_BASE + k for k in range(len(gp))], dtype=np.uint64),
                        ref_moves=np.array(moves_out), total_mark=np.stack(marks), book=np.array(book))
    print("written")


if __name__ == "__main__":
    main()
