"""A plain-C client of the ABI (examples/capi_client.c, gcc -std=c99) against the Python mirror."""
import os
import subprocess

import pytest

import agentgo_b200 as ag
from util import BUILD, ROOT


def build_client():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "capi_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "capi_client.c"), "-L", ag.HERE, "-lagentgo_b200",
                           "-Wl,-rpath," + ag.HERE, "-o", exe])
    return exe


def test_c_client_compiles_and_links(built_lib):
    assert os.path.exists(build_client())


@pytest.mark.gpu
def test_c_client_matches_python_mirror(built_lib):
    out = subprocess.run([build_client(), "9"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    n = 9
    e = ag.Engine(n)
    e.play(ag.BLACK, 4, 4); e.play(ag.WHITE, 3, 4)
    w, sp, sn, _, _ = e.playouts(ag.BLACK, None, 500, seed_base=20151001)
    avrg = int((sp[0] + sn[0]) / 500) if (sp[0] + sn[0]) >= 0 else -int(-(sp[0] + sn[0]) / 500)
    assert lines[0] == "root wins=%d sum_pos=%d sum_neg=%d avrg_win=%d" % (w[0], sp[0], sn[0], avrg)
    cands = [tuple(int(x) for x in l.split()[1:3]) for l in lines if l.startswith("cand ")]
    w, sp, sn, _, _ = e.playouts(ag.BLACK, cands, 150, avrg_win=avrg, seed_base=20151002)
    for l, a, b, c in zip([l for l in lines if l.startswith("cand ")], w, sp, sn):
        assert l.endswith("wins=%d sum_pos=%d sum_neg=%d" % (a, b, c))
    ua, ub = e.static_units(ag.BLACK, *cands[0])
    assert lines[-2] == "static %d %d units_a=%d units_b=%d calc_result=%d" % (cands[0] + (ua, ub, e.calc_result()))
    mv, _ = e.genmove(ag.BLACK, 150, 500, seed_base=20151001)
    assert lines[-1] == "genmove row=%d col=%d pass=0" % mv
