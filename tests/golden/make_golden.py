"""Generate tests/golden/*.npz|json from the REFERENCE ITSELF.

Runs the reference's own Board.cpp (compiled by oracle/build_ref.sh from /root/reference into
oracle/_ref/) and stores small input/output vectors.  The reference ships no tests or golden
vectors, so these are the pin for the oracle restatement (oracle/agb_oracle.c),
the host replay board and the CUDA kernels.  Re-run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
import agentgo_b200 as ag  # noqa: E402  (only for playout_seed, a pure Python function)

SEED_BASE = 20151001
KIND = "reference"


def tinymt_kat():
    """RFC 8682 parameter set, seed 1: first five outputs (from-spec Python implementation)."""
    M = 0xFFFFFFFF
    mat1, mat2, tmat = 0x8F7011EE, 0xFC78FF1F, 0x3793FDFF
    st = [1, mat1, mat2, tmat]

    def nxt():
        y = st[3]
        x = (st[0] & 0x7FFFFFFF) ^ st[1] ^ st[2]
        x ^= (x << 1) & M
        y ^= (y >> 1) ^ x
        st[0], st[1], st[2], st[3] = st[1], st[2], x ^ ((y << 10) & M), y
        if y & 1:
            st[1] ^= mat1
            st[2] ^= mat2

    for i in range(1, 8):
        st[i & 3] ^= (i + 1812433253 * (st[(i - 1) & 3] ^ (st[(i - 1) & 3] >> 30))) & M
    for _ in range(8):
        nxt()
    out = []
    for _ in range(5):
        nxt()
        t1 = (st[0] + (st[2] >> 8)) & M
        t0 = st[3] ^ t1
        if t1 & 1:
            t0 ^= tmat
        out.append(t0)
    return out


def kats():
    k = {"tinymt32_seed1": tinymt_kat()}
    # KAT1: the commented-out position of AgentGo.cpp:728-738 (13x13)
    b = po.OracleBoard(13, KIND)
    for c, r, cc in [(2, 1, 0), (2, 1, 1), (2, 0, 1), (1, 2, 0), (1, 2, 1), (1, 2, 2), (1, 1, 2), (1, 0, 2)]:
        b.put(c, r, cc)
    k["kat1_checks"] = [b.check("suicide", 1, 0, 0), b.check("kill", 1, 0, 0),
                        b.check("suicide", 2, 0, 0), b.check("true_eye", 2, 0, 0)]
    b.put(1, 0, 0)
    st = b.export_state()
    NN = 169
    snext = st[16 + 6 * NN:16 + 7 * NN]
    lst, i = [], int(st[3])
    for _ in range(6):
        lst.append(i)
        i = int(snext[i])
    k["kat1_after"] = {"num_black": int(st[1]), "num_white": int(st[2]), "exist_compete": int(st[4]),
                       "empty_list_head6": lst, "score": [b.count_score(1), b.count_score(2)]}
    # KAT2: GoContest.cpp:12-17 — four black corner stones
    b = po.OracleBoard(13, KIND)
    k["kat2_empty_score"] = [b.count_score(1), b.count_score(2)]
    for r, c in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        b.put(1, r, c)
    k["kat2_score"] = [b.count_score(1), b.count_score(2)]
    # KAT3: single-stone capture sets the ko flag
    b = po.OracleBoard(13, KIND)
    b.put(2, 0, 0); b.put(1, 0, 1); b.put(1, 1, 0)
    st = b.export_state()
    k["kat3"] = {"exist_compete": int(st[4]), "compete": [int(x) for x in st[5:8]], "space_head": int(st[3])}
    k["sizeof_board"] = {str(n): po.load(KIND, n)[0].orc_sizeof_board() for n in (9, 13, 19)}
    return k


def midgame_moves(n, k):
    """position k = M_k = 40 + (k mod 200) random moves from the empty
    board, black first, ONE TinyMT32 stream seeded seed(seed_base+1, k, 0, 0)."""
    m = 40 + (k % 200)
    if n < 19:
        m = min(m, n * n // 2)
    seed = ag.playout_seed(SEED_BASE + 1, k, 0, 0)
    return po.OracleBoard(n, KIND).random_game(1, m, seed)


def main():
    if not all(po.ref_available(n) for n in (9, 13, 19)):
        assert po.build_ref(), "cannot build oracle/_ref (needs /root/reference)"
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kats(), f, indent=1, sort_keys=True)

    # root mode, empty board (config 2 prefix; also 9 and 13)
    for n, P in ((9, 4096), (13, 2048), (19, 4096)):
        b = po.OracleBoard(n, KIND)
        w, sp, sn, d, st = b.playouts(1, None, P, seed_base=SEED_BASE, threads=0, with_stats=True)
        np.savez_compressed(os.path.join(HERE, "root_%d.npz" % n), n=n, colour=1, P=P, seed_base=SEED_BASE,
                            diffs=d[0], wins=w, sum_pos=sp, sum_neg=sn,
                            moves=st["moves"], rng_draws=st["rng_draws"])

    # config 1: 9x9 empty, all 81 candidates, black to play (white answers first)
    n, P = 9, 128
    cands = [(i // n, i % n) for i in range(n * n)]
    b = po.OracleBoard(n, KIND)
    w, sp, sn, d = b.playouts(1, cands, P, avrg_win=0, seed_base=SEED_BASE, threads=0)
    np.savez_compressed(os.path.join(HERE, "cands_9.npz"), n=n, colour=1, P=P, seed_base=SEED_BASE, avrg_win=0,
                        cands=np.array(cands, dtype=np.int16), diffs=d, wins=w, sum_pos=sp, sum_neg=sn)

    # config 3 subset: 16 mid-game 19x19 positions, root mode, side to move
    n, P = 19, 256
    ks = list(range(0, 208, 13))
    moves, lens, to_move = [], [], []
    diffs = np.zeros((len(ks), P), dtype=np.int16)
    sums = np.zeros((len(ks), 3), dtype=np.int64)
    states, checks = [], []
    for bi, k in enumerate(ks):
        mv = midgame_moves(n, k)
        colour = 1 + (len(mv) % 2)               # black moved first
        ob = po.OracleBoard(n, KIND).replay(mv)
        w, sp, sn, d = ob.playouts(colour, None, P, seed_base=SEED_BASE, position_id=k, threads=0)
        diffs[bi] = d[0]
        sums[bi] = (w[0], sp[0], sn[0])
        moves.append(mv); lens.append(len(mv)); to_move.append(colour)
        states.append(ob.export_state())
        tab = np.zeros((8, 2, n * n), dtype=np.int8)
        for which in range(8):
            for c in (1, 2):
                for pt in range(n * n):
                    tab[which, c - 1, pt] = ob.check(which, c, pt // n, pt % n)
        checks.append(tab)
    flat = np.concatenate(moves).astype(np.uint32)
    np.savez_compressed(os.path.join(HERE, "midgame_19.npz"), n=n, P=P, seed_base=SEED_BASE,
                        position_ids=np.array(ks), moves=flat, lens=np.array(lens),
                        to_move=np.array(to_move), diffs=diffs, sums=sums,
                        states=np.stack(states), checks=np.stack(checks),
                        scores=np.array([[po.OracleBoard(n, KIND).replay(m).count_score(c) for c in (1, 2)]
                                         for m in moves]))

    # a candidate-mode case on a mid-game position with a ko and captures around (13x13)
    n, P = 13, 64
    mv = midgame_moves(n, 77)
    ob = po.OracleBoard(n, KIND).replay(mv)
    st = ob.export_state()
    empties = [i for i in range(n * n) if st[16 + i] == 0]
    cands = [(i // n, i % n) for i in empties]
    w, sp, sn, d = ob.playouts(2, cands, P, avrg_win=-2, seed_base=SEED_BASE + 7, position_id=5, threads=0)
    np.savez_compressed(os.path.join(HERE, "cands_13_midgame.npz"), n=n, colour=2, P=P, seed_base=SEED_BASE + 7,
                        avrg_win=-2, position_id=5, moves=mv, cands=np.array(cands, dtype=np.int16),
                        diffs=d, wins=w, sum_pos=sp, sum_neg=sn, state=st)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
