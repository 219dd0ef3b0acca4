"""Shared helpers for the tests."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "_build")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def split_moves(g):
    """midgame_19.npz stores all move lists back to back."""
    out, o = [], 0
    for ln in g["lens"]:
        out.append(g["moves"][o:o + int(ln)].astype(np.uint32))
        o += int(ln)
    return out


def oracle_kinds(n):
    """The oracles available on this machine: always the C port; the compiled reference when
    oracle/_ref was built (here from /root/reference, on the GPU box from the shipped .so)."""
    from oracle import pyoracle as po
    kinds = ["port"]
    if po.ref_available(n):
        kinds.append("reference")
    return kinds


def host_sim():
    """Test-only CPU harness that runs the __host__ __device__ playout logic of
    agentgo_b200/csrc/board_ops.h (tests/csrc/host_sim.cpp)."""
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libhost_sim.so")
    srcs = [os.path.join(ROOT, "tests", "csrc", "host_sim.cpp"),
            os.path.join(ROOT, "agentgo_b200", "csrc", "host_board.cpp")]
    deps = srcs + [os.path.join(ROOT, "agentgo_b200", "csrc", "board_ops.h"),
                   os.path.join(ROOT, "agentgo_b200", "csrc", "host_board.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                               "-I", os.path.join(ROOT, "agentgo_b200", "csrc"),
                               "-I", os.path.join(ROOT, "include")] + srcs + ["-o", lib])
    return C.CDLL(lib)


def sim_playouts(sim, n, moves, colour, cand, P, avrg_win=0, seed_base=20151001, position_id=0, cand_idx=0):
    mv = np.ascontiguousarray(moves, dtype=np.uint32)
    d = np.zeros(P, dtype=np.int16)
    o = np.zeros(3, dtype=np.int64)
    s = np.zeros(2, dtype=np.int64)
    r = sim.sim_playouts(n, mv.ctypes.data_as(C.c_void_p), len(mv), colour,
                         cand[0] if cand is not None else -1, cand[1] if cand is not None else -1,
                         C.c_uint32(position_id), C.c_uint32(cand_idx), C.c_int64(P), C.c_uint64(seed_base),
                         avrg_win, d.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                         s.ctypes.data_as(C.c_void_p))
    assert r == 0
    return o, d, s


def emu_group():
    """Test-only CPU build of the group kernel's device code on the warp emulator
    (tests/csrc/emu_group.cpp, tests/csrc/simt_emu.{h,cpp})."""
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libemu_group.so")
    csrc = os.path.join(ROOT, "agentgo_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "csrc", "emu_group.cpp"), os.path.join(ROOT, "tests", "csrc", "simt_emu.cpp"),
            os.path.join(csrc, "host_board.cpp"), os.path.join(csrc, "host_layout.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in ("playout_group.cuh", "kernel_args.h", "board_ops.h", "host_board.h")]
    deps.append(os.path.join(ROOT, "tests", "csrc", "simt_emu.h"))
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-I", csrc,
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "csrc"),
                               "-I", cuda_inc] + srcs + ["-o", lib])
    so = C.CDLL(lib)
    so.emu_group_playouts.restype = C.c_longlong
    return so


def emu_playouts(emu, n, lanes, moves, colour, cand, P, avrg_win=0, seed_base=20151001, position_id=0,
                 cand_idx=0, max_moves=0, amaf=None):
    mv = np.ascontiguousarray(moves, dtype=np.uint32)
    d = np.zeros(P, dtype=np.int16)
    o = np.zeros(3, dtype=np.int64)
    s = np.zeros(4, dtype=np.int64)
    r = emu.emu_group_playouts(n, lanes, mv.ctypes.data_as(C.c_void_p), len(mv), colour,
                               cand[0] if cand is not None else -1, cand[1] if cand is not None else -1,
                               C.c_uint32(position_id), C.c_uint32(cand_idx), C.c_int64(P), C.c_uint64(seed_base),
                               avrg_win, max_moves, d.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                               s.ctypes.data_as(C.c_void_p), amaf.ctypes.data_as(C.c_void_p) if amaf is not None else None)
    assert r >= 0
    return o, d, s, r
