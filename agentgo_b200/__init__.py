"""agentgo_b200 — B200-native flat Monte-Carlo playout engine (drop-in for AgentGo's playout path).

This package is only a thin ctypes mirror of the C ABI in include/agentgo_b200.h, used by the
tests and bench.py; the product is the shared library agentgo_b200/libagentgo_b200.so (CUDA
kernels for sm_100a + C/C++ host code) and the GTP front end agentgo_b200/agentgo_gtp.

There is no CPU fallback: if the library is missing, import fails; if no GPU is present,
every compute call raises AgbError(AGB_ERR_CUDA).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGB_LIB", os.path.join(HERE, "libagentgo_b200.so"))

EMPTY, BLACK, WHITE = 0, 1, 2
KERNEL_AUTO, KERNEL_THREAD, KERNEL_WARP, KERNEL_GROUP8, KERNEL_GROUP16, KERNEL_GROUP32 = 0, 1, 2, 3, 4, 5
ROOT_CANDIDATE = 0xFFFFFFFF
AMAF_RANGE = 40
CHECKS = {"suicide": 0, "kill": 1, "survive": 2, "dying": 3, "chase": 4,
          "true_eye": 5, "compete": 6, "no_sense": 7}
STATUS = {0: "AGB_OK", -1: "AGB_ERR_INVALID", -2: "AGB_ERR_OCCUPIED", -3: "AGB_ERR_CUDA",
          -4: "AGB_ERR_NOMEM", -5: "AGB_ERR_FAULT", -6: "AGB_ERR_STATE"}

# every symbol include/agentgo_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "agb_default_config", "agb_create", "agb_destroy", "agb_last_error", "agb_version",
    "agb_device_count", "agb_boardsize", "agb_clear_board", "agb_play", "agb_pass",
    "agb_get_board", "agb_check", "agb_count_score", "agb_export_state", "agb_playouts",
    "agb_playouts_amaf", "agb_playouts_launch", "agb_playouts_finish", "agb_counters_dev", "agb_stream",
    "agb_set_allreduce", "agb_playouts_batch", "agb_genmove", "agb_measure_peaks",
    "agb_debug_violations", "agb_attach_nccl", "agb_detach_nccl", "agb_set_allgather", "agb_static_units", "agb_last_marks",
    "agb_genmove_step", "agb_set_genmove_step", "agb_calc_result", "agb_random_piece", "agb_debug_event",
]


class AgbError(RuntimeError):
    def __init__(self, status, msg):
        RuntimeError.__init__(self, "%s: %s" % (STATUS.get(status, status), msg))
        self.status = status


class Config(C.Structure):
    _fields_ = [("board_size", C.c_int), ("device", C.c_int), ("kernel", C.c_int),
                ("shard_rank", C.c_int), ("shard_count", C.c_int),
                ("mat1", C.c_uint32), ("mat2", C.c_uint32), ("tmat", C.c_uint32),
                ("max_moves", C.c_int), ("genmove_amaf", C.c_int),
                ("genmove_static", C.c_int), ("genmove_book", C.c_int),
                ("index_a", C.c_double), ("index_b", C.c_double), ("index_c", C.c_double),
                ("index_amaf_a1", C.c_double), ("index_amaf_a2", C.c_double),
                ("index_amaf_b1", C.c_double), ("index_amaf_b2", C.c_double), ("index_cnsv", C.c_double)]

# the eight tunables of the reference's command line, in argv order (AgentGo.cpp:709-719)
TUNABLES = ("index_a", "index_b", "index_c", "index_amaf_a1", "index_amaf_a2", "index_amaf_b1", "index_amaf_b2",
            "index_cnsv")


class Stats(C.Structure):
    _fields_ = [("device_ms", C.c_float), ("playouts", C.c_int64), ("moves", C.c_int64),
                ("rng_draws", C.c_int64), ("kernel_launches", C.c_int32), ("faults", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PositionC(C.Structure):
    _fields_ = [("moves", C.POINTER(C.c_uint32)), ("n_moves", C.c_int32), ("to_move", C.c_int32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)

_lib = None


def build_lib(force=False):
    """Compile libagentgo_b200.so in-tree (see agentgo_b200/build.py)."""
    import importlib
    return importlib.import_module("agentgo_b200.build").build(force=force)


def lib():
    """Load libagentgo_b200.so (no fallback: raises if it was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `python -m agentgo_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    i16p, i32p, i64p = C.POINTER(C.c_int16), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    vp = C.c_void_p
    l.agb_default_config.argtypes = [C.POINTER(Config)]
    l.agb_default_config.restype = None
    l.agb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    l.agb_destroy.argtypes = [vp]
    l.agb_destroy.restype = None
    l.agb_last_error.argtypes = [vp]
    l.agb_last_error.restype = C.c_char_p
    l.agb_version.restype = C.c_char_p
    l.agb_boardsize.argtypes = [vp, C.c_int]
    l.agb_clear_board.argtypes = [vp]
    l.agb_play.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    l.agb_pass.argtypes = [vp, C.c_int]
    l.agb_get_board.argtypes = [vp, C.POINTER(C.c_uint8), C.POINTER(C.c_int)]
    l.agb_check.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    l.agb_count_score.argtypes = [vp, C.c_int]
    l.agb_export_state.argtypes = [vp, i32p, C.c_int]
    l.agb_playouts.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_uint32,
                               i64p, i64p, i64p, i16p, C.POINTER(Stats)]
    l.agb_playouts_amaf.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int64, C.c_int, C.c_uint64, C.c_uint32,
                                    i64p, i64p, i64p, i64p, C.POINTER(Stats)]
    l.agb_playouts_launch.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                      C.c_uint32, C.c_int]
    l.agb_playouts_finish.argtypes = [vp, i64p, i64p, i64p, i16p, C.POINTER(Stats)]
    l.agb_counters_dev.argtypes = [vp]
    l.agb_counters_dev.restype = vp
    l.agb_stream.argtypes = [vp]
    l.agb_stream.restype = vp
    l.agb_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    l.agb_set_allgather.argtypes = [vp, ALLGATHER_FN, vp]
    l.agb_playouts_batch.argtypes = [vp, C.POINTER(PositionC), C.c_int, C.c_int64, C.c_uint64,
                                     C.c_uint32, i64p, i64p, i64p, i16p, C.POINTER(Stats)]
    l.agb_genmove.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, C.c_uint64,
                              C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                              C.POINTER(Stats)]
    l.agb_random_piece.argtypes = [vp, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.agb_calc_result.argtypes = [vp, C.POINTER(C.c_int)]
    l.agb_static_units.argtypes = [vp, C.c_int, C.c_int, C.c_int, i64p, i64p]
    l.agb_last_marks.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    l.agb_genmove_step.argtypes = [vp]
    l.agb_set_genmove_step.argtypes = [vp, C.c_int]
    l.agb_measure_peaks.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_int)]
    l.agb_debug_violations.restype = C.c_longlong
    l.agb_debug_event.restype = C.c_longlong
    l.agb_debug_event.argtypes = [C.c_int]
    _lib = l
    return l


def debug_violations():
    """Bounds violations counted by the AGB_CHECK_BOUNDS build (0 in the release build)."""
    return int(lib().agb_debug_violations())


def debug_event(which):
    """Coverage counters of the AGB_CHECK_BOUNDS build (see include/agentgo_b200.h)."""
    return int(lib().agb_debug_event(which))


def measure_peaks(device=0):
    """-> dict(int_ops_per_s, int_ops_per_s_alu_only, smem_bytes_per_s, sm_count), measured live."""
    a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    st = lib().agb_measure_peaks(device, C.byref(a), C.byref(b), C.byref(c), C.byref(n))
    if st:
        raise AgbError(st, "agb_measure_peaks")
    return {"int_ops_per_s": a.value, "int_ops_per_s_alu_only": b.value,
            "smem_bytes_per_s": c.value, "sm_count": n.value}


def move(colour, row, col):
    return (colour << 16) | ((row & 0xFF) << 8) | (col & 0xFF)


def move_pass(colour):
    return move(colour, 0xFF, 0xFF)


def playout_seed(seed_base, position_id, candidate_idx, playout_idx):
    """Python restatement of agb_playout_seed() (include/agentgo_b200.h)."""
    M = (1 << 64) - 1
    z = seed_base & M
    z = (z + 0x9E3779B97F4A7C15 * (playout_idx + 1)) & M
    z ^= (0xBF58476D1CE4E5B9 * (candidate_idx + 1)) & M
    z = (z + 0x94D049BB133111EB * (position_id + 1)) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z ^= z >> 31
    return (z ^ (z >> 32)) & 0xFFFFFFFF


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype)) if arr is not None else None


class Engine:
    """Mirror of the reference's playout job interface for this path.

    reference                                   here
    Scheduler<MyWorker> + MyJob (AgentGo.cpp)   Engine.playouts(colour, cands, P, avrg_win)
    MyGame::testAvrgWin                         Engine.playouts(colour, None, P)
    GTPAdapter::bd put/clear                    Engine.play / clear_board / boardsize
    """

    def __init__(self, board_size=13, device=0, kernel=KERNEL_AUTO, shard_rank=0, shard_count=1,
                 max_moves=0, genmove_amaf=0, genmove_static=0, genmove_book=0, tunables=None):
        self._l = lib()
        cfg = Config()
        self._l.agb_default_config(C.byref(cfg))
        cfg.board_size, cfg.device, cfg.kernel = board_size, device, kernel
        cfg.shard_rank, cfg.shard_count, cfg.max_moves = shard_rank, shard_count, max_moves
        cfg.genmove_amaf, cfg.genmove_static, cfg.genmove_book = genmove_amaf, genmove_static, genmove_book
        if tunables is not None:            # eight doubles in the order of the reference's argv, or a dict
            for k, v in (tunables.items() if isinstance(tunables, dict) else zip(TUNABLES, tunables)):
                setattr(cfg, k, float(v))
        h = C.c_void_p()
        st = self._l.agb_create(C.byref(cfg), C.byref(h))
        if st:
            raise AgbError(st, self._l.agb_last_error(None).decode())
        self._h = h
        self.n = board_size
        self.shard_rank, self.shard_count = shard_rank, shard_count
        self._cb = None

    def close(self):
        if getattr(self, "_h", None):
            self._l.agb_destroy(self._h)
            self._h = None

    __del__ = close

    def _chk(self, st):
        if st < 0:
            raise AgbError(st, self._l.agb_last_error(self._h).decode())
        return st

    # ---- position
    def boardsize(self, n):
        self._chk(self._l.agb_boardsize(self._h, n))
        self.n = n

    def clear_board(self):
        self._chk(self._l.agb_clear_board(self._h))

    def play(self, colour, row, col):
        self._chk(self._l.agb_play(self._h, colour, row, col))

    def pass_(self, colour):
        self._chk(self._l.agb_pass(self._h, colour))

    def replay(self, moves):
        for m in moves:
            m = int(m)
            colour, row, col = (m >> 16) & 0xFF, (m >> 8) & 0xFF, m & 0xFF
            if row == 0xFF:
                self.pass_(colour)
            else:
                self.play(colour, row, col)
        return self

    def board(self):
        out = np.zeros(self.n * self.n, dtype=np.uint8)
        n = C.c_int()
        self._chk(self._l.agb_get_board(self._h, _p(out, C.c_uint8), C.byref(n)))
        return out.reshape(self.n, self.n)

    def check(self, which, colour, row, col):
        w = CHECKS[which] if isinstance(which, str) else which
        return self._chk(self._l.agb_check(self._h, w, colour, row, col))

    def count_score(self, colour):
        return self._chk(self._l.agb_count_score(self._h, colour))

    def export_state(self):
        words = 16 + 7 * self.n * self.n
        out = np.zeros(words, dtype=np.int32)
        r = self._l.agb_export_state(self._h, _p(out, C.c_int32), words)
        if r != words:
            raise AgbError(r, "agb_export_state")
        return out

    # ---- hot path
    def playouts_launch(self, colour, cands, P, avrg_win=0, seed_base=20151001, position_id=0,
                        want_diffs=False):
        cands = list(cands) if cands is not None else []
        self._C = len(cands)
        self._P = P
        rc = np.array(cands, dtype=np.int16).reshape(-1) if cands else None
        self._chk(self._l.agb_playouts_launch(self._h, colour, _p(rc, C.c_int16), self._C, P, avrg_win,
                                              seed_base, position_id, 1 if want_diffs else 0))
        self._want = want_diffs

    def playouts_finish(self):
        nc = max(self._C, 1)
        wins, sp, sn = (np.zeros(nc, dtype=np.int64) for _ in range(3))
        diffs = np.full((nc, self._P), -32768, dtype=np.int16) if self._want else None
        st = Stats()
        self._chk(self._l.agb_playouts_finish(self._h, _p(wins, C.c_int64), _p(sp, C.c_int64),
                                              _p(sn, C.c_int64), _p(diffs, C.c_int16), C.byref(st)))
        return wins, sp, sn, diffs, st.as_dict()

    def playouts(self, colour, cands, P, avrg_win=0, seed_base=20151001, position_id=0,
                 want_diffs=False):
        """-> (wins, sum_pos, sum_neg, diffs|None, stats) for this engine's shard."""
        self.playouts_launch(colour, cands, P, avrg_win, seed_base, position_id, want_diffs)
        return self.playouts_finish()

    def playouts_amaf(self, colour, cands, P, avrg_win=0, seed_base=20151001, position_id=0):
        """Candidate-mode playouts with AMAF first-play marks.
        -> (wins, sum_pos, sum_neg, amaf int64 [2 which][2 sign][41 step][N*N], stats)"""
        cands = list(cands)
        Cn = len(cands)
        rc = np.array(cands, dtype=np.int16).reshape(-1)
        wins, sp, sn = (np.zeros(Cn, dtype=np.int64) for _ in range(3))
        amaf = np.zeros((2, 2, AMAF_RANGE + 1, self.n * self.n), dtype=np.int64)
        st = Stats()
        self._chk(self._l.agb_playouts_amaf(self._h, colour, _p(rc, C.c_int16), Cn, P, avrg_win, seed_base,
                                            position_id, _p(wins, C.c_int64), _p(sp, C.c_int64),
                                            _p(sn, C.c_int64), _p(amaf, C.c_int64), C.byref(st)))
        return wins, sp, sn, amaf, st.as_dict()

    def playouts_batch(self, positions, P, seed_base=20151001, first_position_id=0, want_diffs=False):
        """positions: list of (moves uint32 array, to_move)."""
        B = len(positions)
        arr = (PositionC * B)()
        keep = []
        for i, (mv, to_move) in enumerate(positions):
            mv = np.ascontiguousarray(mv, dtype=np.uint32)
            keep.append(mv)
            arr[i].moves = _p(mv, C.c_uint32)
            arr[i].n_moves = len(mv)
            arr[i].to_move = to_move
        wins, sp, sn = (np.zeros(B, dtype=np.int64) for _ in range(3))
        diffs = np.full((B, P), -32768, dtype=np.int16) if want_diffs else None
        st = Stats()
        self._chk(self._l.agb_playouts_batch(self._h, arr, B, P, seed_base, first_position_id,
                                             _p(wins, C.c_int64), _p(sp, C.c_int64), _p(sn, C.c_int64),
                                             _p(diffs, C.c_int16), C.byref(st)))
        return wins, sp, sn, diffs, st.as_dict()

    def genmove(self, colour, P_per_candidate=150, P_root=500, seed_base=20151001):
        """-> (row, col) or None for pass, plus stats.  Defaults are the reference's
        numset = 150 (AgentGo.cpp:85) and 500 root playouts (AgentGo.cpp:413)."""
        row, col, is_pass = C.c_int(), C.c_int(), C.c_int()
        st = Stats()
        self._chk(self._l.agb_genmove(self._h, colour, P_per_candidate, P_root, seed_base,
                                      C.byref(row), C.byref(col), C.byref(is_pass), C.byref(st)))
        return (None if is_pass.value else (row.value, col.value)), st.as_dict()

    def random_piece(self, colour, seed):
        """Board::getRandomPiece on the host (the match referee's fallback move); (row, col) or None."""
        row, col, is_pass = C.c_int(), C.c_int(), C.c_int()
        self._chk(self._l.agb_random_piece(self._h, colour, seed, C.byref(row), C.byref(col), C.byref(is_pass)))
        return None if is_pass.value else (row.value, col.value)

    def calc_result(self):
        """CalcResults of the reference's match harness (GeneHatchery.cpp:393-436): black - white."""
        d = C.c_int()
        self._chk(self._l.agb_calc_result(self._h, C.byref(d)))
        return d.value

    def static_units(self, colour, row, col):
        """-> (units_a, units_b): the static pattern marks of MyWorker::work (AgentGo.cpp:104-282)
        for one candidate, total_mark += units_a*bonus_a + units_b*bonus_b.  Host only."""
        ua, ub = C.c_int64(), C.c_int64()
        self._chk(self._l.agb_static_units(self._h, colour, row, col, C.byref(ua), C.byref(ub)))
        return ua.value, ub.value

    def last_marks(self):
        """total_mark of the last genmove as an (N, N) array, or None after a book move."""
        out = np.zeros((self.n, self.n), dtype=np.float64)
        r = self._l.agb_last_marks(self._h, _p(out, C.c_double), out.size)
        self._chk(r)
        return out if r else None

    @property
    def genmove_step(self):
        return int(self._l.agb_genmove_step(self._h))

    @genmove_step.setter
    def genmove_step(self, v):
        self._chk(self._l.agb_set_genmove_step(self._h, int(v)))

    def counters_dev(self):
        return self._l.agb_counters_dev(self._h)

    def stream(self):
        return self._l.agb_stream(self._h)

    def set_allreduce(self, fn):
        """fn(dev_ptr:int, count:int, stream:int) -> 0; sums int64 counters over all shards."""
        if fn is None:
            self._cb = None
            self._chk(self._l.agb_set_allreduce(self._h, C.cast(None, ALLREDUCE_FN), None))
            return
        self._cb = ALLREDUCE_FN(lambda ctx, ptr, count, stream: int(fn(ptr, count, stream) or 0))
        self._chk(self._l.agb_set_allreduce(self._h, self._cb, None))

    def set_allgather(self, fn):
        """fn(send_ptr:int, recv_ptr:int, count_per_shard:int, stream:int) -> 0; gathers the int64 tables of
        all shards in shard order (the collective of playouts_batch on sharded engines)."""
        if fn is None:
            self._cbg = None
            self._chk(self._l.agb_set_allgather(self._h, C.cast(None, ALLGATHER_FN), None))
            return
        self._cbg = ALLGATHER_FN(lambda ctx, send, recv, count, stream: int(fn(send, recv, count, stream) or 0))
        self._chk(self._l.agb_set_allgather(self._h, self._cbg, None))
