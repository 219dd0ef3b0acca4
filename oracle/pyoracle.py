"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over the `orc_*` API exported by
  * oracle/_ref/liboracle_ref_<N>.so  — the reference's own Board.cpp compiled by
    oracle/build_ref.sh (kind "reference"), and
  * oracle/liboracle_port.so          — the plain-C restatement oracle/agb_oracle.c
    (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (agentgo_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_LIB = os.path.join(HERE, "liboracle_port.so")

CHECKS = {"suicide": 0, "kill": 1, "survive": 2, "dying": 3, "chase": 4,
          "true_eye": 5, "compete": 6, "no_sense": 7}
ROOT_CANDIDATE = 0xFFFFFFFF
AMAF_RANGE = 40          # GoContest/AgentGo.h:16


def move(colour, row, col):
    return (colour << 16) | ((row & 0xFF) << 8) | (col & 0xFF)


def move_pass(colour):
    return move(colour, 0xFF, 0xFF)


def decode_move(m):
    colour, row, col = (m >> 16) & 0xFF, (m >> 8) & 0xFF, m & 0xFF
    if row == 0xFF:
        return colour, None, None
    return colour, row, col


def build_port(force=False):
    """Compile the C restatement (gcc) into oracle/liboracle_port.so."""
    src = os.path.join(HERE, "agb_oracle.c")
    deps = [src, os.path.join(HERE, "tinymt32.h"),
            os.path.join(HERE, "..", "include", "agentgo_b200.h")]
    if (not force and os.path.exists(PORT_LIB)
            and all(os.path.getmtime(PORT_LIB) >= os.path.getmtime(d) for d in deps)):
        return PORT_LIB
    cmd = ["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-pthread", "-Wall",
           "-I", HERE, "-I", os.path.join(HERE, "..", "include"), "-o", PORT_LIB, src]
    subprocess.check_call(cmd)
    return PORT_LIB


def build_ref():
    """Compile the reference itself (needs /root/reference); returns True when built."""
    r = subprocess.run([os.path.join(HERE, "build_ref.sh")], capture_output=True, text=True)
    return r.returncode == 0


def ref_available(n):
    return os.path.exists(os.path.join(REF_DIR, "liboracle_ref_%d.so" % n))


def _bind(lib):
    i16p, i32p, i64p, u32p = (C.POINTER(C.c_int16), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int64), C.POINTER(C.c_uint32))
    lib.orc_kind.restype = C.c_char_p
    lib.orc_board_size.restype = C.c_int
    lib.orc_new.restype = C.c_void_p
    lib.orc_new.argtypes = [C.c_int]
    lib.orc_free.argtypes = [C.c_void_p]
    lib.orc_clear.argtypes = [C.c_void_p]
    lib.orc_put.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.orc_pass.argtypes = [C.c_void_p, C.c_int]
    lib.orc_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_count_score.argtypes = [C.c_void_p, C.c_int]
    lib.orc_export_state.argtypes = [C.c_void_p, i32p, C.c_int]
    lib.orc_random_game.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, u32p]
    lib.orc_playouts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                 C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int,
                                 i16p, i64p, i64p]
    lib.orc_playouts_amaf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int, i64p, i64p]
    lib.orc_playouts_mt.argtypes = [C.c_void_p, C.c_int, i16p, C.c_int, C.c_int64, C.c_uint32,
                                    C.c_uint64, C.c_int, C.c_int, i16p, i64p, i64p]
    return lib


_libs = {}


def load(kind, n):
    """kind: 'reference' | 'port' | 'best' (reference when built, else port)."""
    if kind == "best":
        kind = "reference" if ref_available(n) else "port"
    key = (kind, n if kind == "reference" else 0)
    if key not in _libs:
        if kind == "reference":
            path = os.path.join(REF_DIR, "liboracle_ref_%d.so" % n)
        elif kind == "port":
            path = build_port()
        else:
            raise ValueError(kind)
        _libs[key] = _bind(C.CDLL(path))
    return _libs[key], kind


class OracleBoard:
    """A board living inside one of the oracle libraries."""

    def __init__(self, n, kind="best"):
        self.lib, self.kind = load(kind, n)
        self.n = n
        self.h = self.lib.orc_new(n)
        if not self.h:
            raise RuntimeError("orc_new(%d) failed for kind %s" % (n, self.kind))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_free(self.h)
            self.h = None

    def clear(self):
        self.lib.orc_clear(self.h)

    def put(self, colour, row, col):
        r = self.lib.orc_put(self.h, colour, row, col)
        if r:
            raise ValueError("orc_put(%d,%d,%d) -> %d" % (colour, row, col, r))

    def pass_(self, colour):
        self.lib.orc_pass(self.h, colour)

    def replay(self, moves):
        for m in moves:
            colour, row, col = decode_move(int(m))
            if row is None:
                self.pass_(colour)
            else:
                self.put(colour, row, col)
        return self

    def check(self, which, colour, row, col):
        return int(self.lib.orc_check(self.h, CHECKS[which] if isinstance(which, str) else which,
                                      colour, row, col))

    def count_score(self, colour):
        return int(self.lib.orc_count_score(self.h, colour))

    def export_state(self):
        words = 16 + 7 * self.n * self.n
        out = np.zeros(words, dtype=np.int32)
        r = self.lib.orc_export_state(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), words)
        if r != words:
            raise RuntimeError("orc_export_state -> %d" % r)
        return out

    def random_game(self, first_colour, n_moves, seed):
        out = np.zeros(n_moves, dtype=np.uint32)
        self.lib.orc_random_game(self.h, first_colour, n_moves, seed,
                                 out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def playouts_amaf(self, colour, cands, P, avrg_win=0, seed_base=20151001, position_id=0, per_candidate=False):
        """Candidate-mode playouts with AMAF first-play marks (AgentGo.cpp:286-338).
        returns (wins, sum_pos, sum_neg, amaf) with amaf int64 [2 which][2 sign][41 step][NN],
        accumulated over all candidates like the reference's global amaf1/amaf2, or one such table
        per candidate ([C, 2, 2, 41, NN]) when per_candidate is set."""
        nn = self.n * self.n
        out3 = np.zeros((len(cands), 3), dtype=np.int64)
        amaf = np.zeros((len(cands) if per_candidate else 1, 2, 2, AMAF_RANGE + 1, nn), dtype=np.int64)
        for c, (row, col) in enumerate(cands):
            tab = amaf[c if per_candidate else 0]
            r = self.lib.orc_playouts_amaf(self.h, colour, row, col, position_id, c, 0, P, 1, seed_base, avrg_win,
                                           out3[c].ctypes.data_as(C.POINTER(C.c_int64)),
                                           tab.ctypes.data_as(C.POINTER(C.c_int64)))
            if r:
                raise RuntimeError("orc_playouts_amaf -> %d" % r)
        return out3[:, 0].copy(), out3[:, 1].copy(), out3[:, 2].copy(), (amaf if per_candidate else amaf[0])

    def playouts(self, colour, cands, P, avrg_win=0, seed_base=20151001, position_id=0,
                 want_diffs=True, threads=1, with_stats=False):
        """cands: list of (row, col) or None/[] for root mode.
        returns (wins, sum_pos, sum_neg, diffs) as int64 arrays [max(C,1)], diffs int16 [max(C,1), P]."""
        cands = list(cands) if cands else []
        Cn = len(cands)
        nc = max(Cn, 1)
        out3 = np.zeros((nc, 3), dtype=np.int64)
        stat2 = np.zeros(2, dtype=np.int64)
        diffs = np.zeros((nc, P), dtype=np.int16) if want_diffs else None
        dptr = diffs.ctypes.data_as(C.POINTER(C.c_int16)) if want_diffs else None
        rc = np.array(cands, dtype=np.int16).reshape(-1) if Cn else np.zeros(2, dtype=np.int16)
        r = self.lib.orc_playouts_mt(self.h, colour, rc.ctypes.data_as(C.POINTER(C.c_int16)), Cn, P,
                                     position_id, seed_base, avrg_win, threads, dptr,
                                     out3.ctypes.data_as(C.POINTER(C.c_int64)),
                                     stat2.ctypes.data_as(C.POINTER(C.c_int64)))
        if r:
            raise RuntimeError("orc_playouts_mt -> %d" % r)
        res = (out3[:, 0].copy(), out3[:, 1].copy(), out3[:, 2].copy(), diffs)
        if with_stats:
            return res + ({"moves": int(stat2[0]), "rng_draws": int(stat2[1])},)
        return res


GENMOVE_LIB = os.path.join(REF_DIR, "liboracle_ref_genmove_13.so")


def genmove_ref_available():
    return os.path.exists(GENMOVE_LIB)


class RefGame:
    """The reference's own MyGame (AgentGo.cpp:351-705: opening book, testAvrgWin, candidate filter,
    MyWorker::work with its static pattern heuristics, AMAF, argmax), 13x13 only, run with a serial
    schedule and the injected per-playout TinyMT32 (oracle/ref_genmove_shim.h)."""
    N = 13

    def __init__(self):
        key = ("genmove", 13)
        if key not in _libs:
            lib = C.CDLL(GENMOVE_LIB)
            dp = C.POINTER(C.c_double)
            lib.orc_gm_new.restype = C.c_void_p
            lib.orc_gm_free.argtypes = [C.c_void_p]
            lib.orc_gm_clear.argtypes = [C.c_void_p]
            lib.orc_gm_play.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
            lib.orc_gm_step.argtypes = [C.c_void_p]
            lib.orc_gm_set_step.argtypes = [C.c_void_p, C.c_int]
            lib.orc_gm_genmove.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), dp, dp, dp, dp]
            lib.orc_gm_static_mark.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64]
            lib.orc_gm_static_mark.restype = C.c_double
            lib.orc_gm_calc_result.argtypes = [C.c_void_p]
            if hasattr(lib, "orc_gm_set_tunables"):
                lib.orc_gm_set_tunables.argtypes = [dp]
            _libs[key] = lib
        self.lib = _libs[key]
        self.h = self.lib.orc_gm_new()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_gm_free(self.h)
            self.h = None

    def clear(self):
        self.lib.orc_gm_clear(self.h)

    def play(self, colour, row, col):
        r = self.lib.orc_gm_play(self.h, colour, row, col)
        if r:
            raise ValueError("orc_gm_play(%d,%d,%d) -> %d" % (colour, row, col, r))

    def replay(self, moves):
        for m in moves:
            colour, row, col = decode_move(int(m))
            if row is not None:
                self.play(colour, row, col)
        return self

    @property
    def step(self):
        return int(self.lib.orc_gm_step(self.h))

    @step.setter
    def step(self, v):
        self.lib.orc_gm_set_step(self.h, int(v))

    def genmove(self, colour, seed_base=20151001):
        """returns (move or None, tables) — tables: dict of 13x13 double arrays total_mark, mc, amaf1, amaf2"""
        row, col = C.c_int(-1), C.c_int(-1)
        t = {k: np.zeros((13, 13)) for k in ("total_mark", "mc", "amaf1", "amaf2")}
        dp = C.POINTER(C.c_double)
        moved = self.lib.orc_gm_genmove(self.h, colour, seed_base, C.byref(row), C.byref(col),
                                        *[t[k].ctypes.data_as(dp) for k in ("total_mark", "mc", "amaf1", "amaf2")])
        return ((row.value, col.value) if moved else None), t

    DEFAULT_TUNABLES = (1000.0, 1.0, 1.0, 40.0, 39. / 180, 0.3, 0.3 / 180, 0.01)    # AgentGo.cpp:18-25

    def set_tunables(self, v=None):
        """the eight doubles of the reference's command line (AgentGo.cpp:709-719); they are globals of
        the reference, so this holds for every RefGame of the process until it is called again"""
        arr = (C.c_double * 8)(*[float(x) for x in (v or self.DEFAULT_TUNABLES)])
        self.lib.orc_gm_set_tunables(arr)

    def calc_result(self):
        """CalcResults of the match harness (GeneHatchery.cpp:393-436): black - white."""
        return int(self.lib.orc_gm_calc_result(self.h))

    def static_mark(self, colour, row, col):
        """total_mark[row][col] - mc[row][col] after MyWorker::work at avrg_win = 0: the static
        pattern marks of AgentGo.cpp:104-282 (an integer combination of bonus_a, bonus_b)."""
        return float(self.lib.orc_gm_static_mark(self.h, colour, row, col, 1))
