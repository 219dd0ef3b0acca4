#!/usr/bin/env python
"""TEST-ONLY scripted GTP engine for the match-referee test (tests/test_match_referee.py).

usage: gtp_script_engine.py <board size> <stride> <pass-at,pass-at,...>
It answers `genmove` number k with `pass` when k is in the pass list, else with the first point it
has never seen a stone on, walking the board with the given stride (coprime with size*size); every
other command is acknowledged.  It knows nothing about captures or legality.
"""
import sys

n, stride = int(sys.argv[1]), int(sys.argv[2])
passes = set(int(x) for x in sys.argv[3].split(",") if x)
LETTERS = "ABCDEFGHJKLMNOPQRST"
seen = set()
k = 0
for line in sys.stdin:
    w = line.split()
    if not w:
        continue
    cmd = w[0].lower()
    if cmd == "quit":
        sys.stdout.write("=\n\n"); sys.stdout.flush()
        break
    if cmd == "clear_board":
        seen.clear()
        sys.stdout.write("=\n\n")
    elif cmd == "play" and len(w) >= 3 and w[2].lower() != "pass":
        seen.add((LETTERS.index(w[2][0].upper()), int(w[2][1:]) - 1))
        sys.stdout.write("= \n\n")
    elif cmd == "genmove":
        mv = None
        if k not in passes:
            for i in range(n * n):
                q = (i * stride) % (n * n)
                if (q // n, q % n) not in seen:
                    mv = (q // n, q % n)
                    break
        k += 1
        if mv is None:
            sys.stdout.write("= pass\n\n")
        else:
            seen.add(mv)
            sys.stdout.write("= %s%d\n\n" % (LETTERS[mv[0]], mv[1] + 1))
    else:
        sys.stdout.write("= \n\n")
    sys.stdout.flush()
