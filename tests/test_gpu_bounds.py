"""The pool has no compute-sanitizer: the warp kernel's shared-memory indexing is checked by its
own AGB_CHECK_BOUNDS build (python -m agentgo_b200.build --debug), run in a subprocess so the
release library of this process is left alone."""
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import sys
sys.path.insert(0, %r)
sys.path.insert(1, sys.path[0] + '/tests')
import numpy as np
import agentgo_b200 as ag
from oracle import pyoracle as po
assert ag.LIB_PATH.endswith('_dbg.so')
for n, P in ((9, 600), (13, 300), (19, 300)):
    e = ag.Engine(n, kernel=ag.KERNEL_WARP)
    w, sp, sn, d, st = e.playouts(1, None, P, want_diffs=True)
    od = po.OracleBoard(n, 'best').playouts(1, None, P, threads=0)[3]
    assert (d == od).all()
    mv = po.OracleBoard(n, 'best').random_game(1, 2 * n * n, 99)
    e.replay(mv)
    e.playouts(2, None, P)
# mid-game positions with many captures ahead: the empty-list array must fill up and be compacted
from util import golden, split_moves
pos = [(mv, 1 + len(mv) %% 2) for mv in split_moves(golden('midgame_19.npz'))]
e = ag.Engine(19, kernel=ag.KERNEL_WARP)
w, sp, sn, d, st = e.playouts_batch(pos, 400, want_diffs=True)
for k in (0, 7, 15):
    od = po.OracleBoard(19, 'best').replay(pos[k][0]).playouts(pos[k][1], None, 400, position_id=k, threads=0)[3]
    assert (d[k] == od[0]).all()
print('events', [ag.debug_event(k) for k in range(3)])
assert ag.debug_event(1) > 0 and ag.debug_event(2) > 0
assert ag.debug_event(0) > 0, 'the full-array compaction was never exercised'
print('violations', ag.debug_violations())
assert ag.debug_violations() == 0
"""


def test_no_shared_memory_bounds_violations(built_lib):
    import importlib
    lib = importlib.import_module("agentgo_b200.build").build(debug=True)
    env = dict(os.environ, AGB_LIB=lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "violations 0" in r.stdout
