"""The C-ABI library loads and exports every function include/agentgo_b200.h declares."""
import ctypes
import os
import re
import subprocess

import agentgo_b200 as ag
from util import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "agentgo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(agb_[a-z_0-9]+)\s*\(", src))
    names.discard("agb_playout_seed")      # static inline in the header
    return names


def test_header_symbols_exported(built_lib):
    names = declared_symbols()
    assert names == set(ag.SYMBOLS), names ^ set(ag.SYMBOLS)
    l = ctypes.CDLL(ag.LIB_PATH)
    for name in sorted(names):
        assert hasattr(l, name), name


def test_no_torch_or_cxx_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "agentgo_b200.h")).read()
    assert "std::" not in src and "torch" not in src and "at::" not in src


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "agentgo_b200.h"\nint main(void){agb_config c; (void)c; '
                 'return (int)(agb_playout_seed(1,2,3,4) & 0);}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_seed_function_matches_python_restatement(port_lib):
    # the oracle uses the header's inline function; the product restates it in board_ops.h
    assert ag.playout_seed(20151001, 0, ag.ROOT_CANDIDATE, 0) == ag.playout_seed(20151001, 0, 0xFFFFFFFF, 0)
    vals = {ag.playout_seed(20151001, pid, c, p) for pid in range(3) for c in range(3) for p in range(50)}
    assert len(vals) == 450


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under agentgo_b200/ may import, link or run it."""
    pkg = os.path.join(ROOT, "agentgo_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cpp", ".cu", ".cuh")):
                txt = open(os.path.join(d, f)).read()
                assert "pyoracle" not in txt and "oracle/" not in txt.replace("the oracle/", ""), os.path.join(d, f)
                assert "liboracle" not in txt
