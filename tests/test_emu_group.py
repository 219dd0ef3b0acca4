"""The group kernel's device code (agentgo_b200/csrc/playout_group.cuh: 32/G playouts per warp, G
lanes each) compiled for the CPU on the test-only warp emulator (tests/csrc/simt_emu.h) and compared
with the golden vectors generated from the reference.  The emulator also enforces the property the
kernel is built on: every warp collective is reached by all 32 lanes at the same call site.
This is NOT a product path: the library has no CPU fallback."""
import os

import numpy as np
import pytest

from util import emu_group, emu_playouts, golden, split_moves


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_root_mode_golden(lanes):
    emu = emu_group()
    for n, P in ((9, 192), (13, 96), (19, 64)):
        g = golden("root_%d.npz" % n)
        o, d, s, _ = emu_playouts(emu, n, lanes, [], 1, None, P, seed_base=int(g["seed_base"]))
        assert (d == g["diffs"][:P]).all()
        assert s[3] == P and s[2] == 0


@pytest.mark.parametrize("lanes", [8, 16])
def test_candidates_golden(lanes):
    emu = emu_group()
    g = golden("cands_13_midgame.npz")
    for ci in range(0, len(g["cands"]), 5):
        c = tuple(int(x) for x in g["cands"][ci])
        o, d, s, _ = emu_playouts(emu, 13, lanes, g["moves"], int(g["colour"]), c, int(g["P"]), avrg_win=int(g["avrg_win"]),
                                  seed_base=int(g["seed_base"]), position_id=int(g["position_id"]), cand_idx=ci)
        assert (d == g["diffs"][ci]).all()
        assert (o[0], o[1], o[2]) == (g["wins"][ci], g["sum_pos"][ci], g["sum_neg"][ci])


def test_amaf_marks_match_oracle(port_lib):
    """The first-play marks (AgentGo.cpp:289-290,304,319,329-338) recorded by the group kernel, against
    the oracle's restatement of the reference's loop, candidate by candidate."""
    emu = emu_group()
    po = port_lib
    for n, nmoves, P in ((9, 20, 24), (13, 60, 12)):
        mv = po.OracleBoard(n, "port").random_game(1, nmoves, 4711 + n)
        ob = po.OracleBoard(n, "port").replay(mv)
        st = ob.export_state()
        empties = [i for i in range(n * n) if st[16 + i] == 0]
        cands = [(i // n, i % n) for i in empties[::max(1, len(empties) // 4)]][:4]
        ow, osp, osn, oa = ob.playouts_amaf(2, cands, P, avrg_win=2, seed_base=31, position_id=5, per_candidate=True)
        for ci, c in enumerate(cands):
            amaf = np.zeros((2, 2, 41, n * n), dtype=np.int64)
            o, d, s, _ = emu_playouts(emu, n, 8, mv, 2, c, P, avrg_win=2, seed_base=31, position_id=5, cand_idx=ci, amaf=amaf)
            assert (o[0], o[1], o[2]) == (ow[ci], osp[ci], osn[ci])
            assert (amaf == oa[ci]).all() and amaf.any()


def test_midgame_golden():
    emu = emu_group()
    g = golden("midgame_19.npz")
    moves = split_moves(g)
    for bi in range(0, len(moves), 3):
        P = 48
        o, d, s, _ = emu_playouts(emu, 19, 8, moves[bi], int(g["to_move"][bi]), None, P, seed_base=int(g["seed_base"]),
                                  position_id=int(g["position_ids"][bi]))
        assert (d == g["diffs"][bi][:P]).all()


def test_random_positions_match_oracle(port_lib):
    """Random positions of every length (openings with all eight neighbourhood cells empty, crowded
    endgames in complex mode, stones on edges and corners of the one-border-column board), root and
    candidate mode, against the oracle — the CPU-side counterpart of tests/soak_parity.py."""
    emu = emu_group()
    po = port_lib
    rng = np.random.RandomState(11)
    for k in range(18):
        n = [9, 13, 19][k % 3]
        length = int(rng.randint(0, 3 * n * n))
        mv = po.OracleBoard(n, "port").random_game(1 + k % 2, length, 2000 + k)
        ob = po.OracleBoard(n, "port").replay(mv)
        st = ob.export_state()
        empties = [i for i in range(n * n) if st[16 + i] == 0]
        if not empties:
            continue
        colour = 1 + (k // 3) % 2
        P = 12
        orr = ob.playouts(colour, None, P, seed_base=k, position_id=k, threads=0)
        o, d, s, _ = emu_playouts(emu, n, 8, mv, colour, None, P, seed_base=k, position_id=k)
        assert (d == orr[3]).all(), (k, n, length)
        cands = [(i // n, i % n) for i in empties[::max(1, len(empties) // 3)]][:3]
        ow, osp, osn, od = ob.playouts(colour, cands, P, avrg_win=1, seed_base=k, position_id=k, threads=0)
        for ci, c in enumerate(cands):
            o, d, s, _ = emu_playouts(emu, n, 8, mv, colour, c, P, avrg_win=1, seed_base=k, position_id=k, cand_idx=ci)
            assert (d == od[ci]).all() and (o[0], o[1], o[2]) == (ow[ci], osp[ci], osn[ci]), (k, n, length, c)


def test_device_layout_of_a_position(port_lib):
    """The one-word-per-point layout the group kernel starts from (kernel_args.h): ONE border
    column — the point right of a row's last point is the point left of the next row's first —, colours
    and string structure of the oracle's board, liberties at the roots, the string's last stone at its
    second stone, the empty-point array in the order of the reference's space_list."""
    import ctypes as C
    emu = emu_group()
    po = port_lib
    for n, nmoves in ((9, 40), (13, 90), (19, 230)):
        mv = np.ascontiguousarray(po.OracleBoard(n, "port").random_game(1, nmoves, 77 + n), dtype=np.uint32)
        st = po.OracleBoard(n, "port").replay(mv).export_state()
        W, NP = n + 1, (n + 1) * (n + 2) + 1
        words = np.zeros(NP, dtype=np.uint32)
        elist = np.zeros(n * n, dtype=np.uint32)
        scal = np.zeros(8, dtype=np.int32)
        assert emu.emu_group_layout(n, mv.ctypes.data_as(C.c_void_p), len(mv), words.ctypes.data_as(C.c_void_p),
                                    elist.ctypes.data_as(C.c_void_p), scal.ctypes.data_as(C.c_void_p)) == NP
        pad = lambda i: (i // n + 1) * W + (i % n + 1)
        col, A, nxt, F = words & 3, (words >> 2) & 511, (words >> 11) & 511, words >> 21
        on_board = {pad(i) for i in range(n * n)}
        data, root, hp, nis = st[16:16 + n * n], st[16 + n * n:16 + 2 * n * n], st[16 + 2 * n * n:16 + 3 * n * n], st[16 + 4 * n * n:16 + 5 * n * n]
        for q in range(NP):
            if q not in on_board:
                assert col[q] == 3 and A[q] == q and F[q] == 1           # border: its own root, "one liberty"
        for i in range(n * n):
            q = pad(i)
            assert col[q] == data[i]
            r, c = divmod(i, n)
            # the four neighbours are one word away in each direction; off the board they are border
            for dr, dc, off in ((-1, 0, -W), (0, -1, -1), (1, 0, W), (0, 1, 1)):
                inside = 0 <= r + dr < n and 0 <= c + dc < n
                assert (col[q + off] != 3) == inside
                if inside:
                    assert q + off == pad((r + dr) * n + c + dc)
            if data[i] == 0:
                assert A[q] == q and F[q] == 1 and elist[(words[q] >> 11) & 1023] == q
            else:
                assert A[q] == pad(int(root[i]))
                assert nxt[q] == (511 if nis[i] < 0 or nis[i] >= n * n else pad(int(nis[i])))
                if root[i] == i:
                    assert F[q] == hp[i]
                    if nxt[q] != 511:                                       # the second stone names the last one
                        last = q
                        while nxt[last] != 511:
                            last = nxt[last]
                        assert F[nxt[q]] == last
        assert scal[0] == (data == 0).sum() and scal[1] == (data != 0).sum()
        # 2 224 B per playout at 19x19: ring 128 + board 421 + candidates 5 words, bank-spreading stride
        if n == 19:
            assert scal[7] * 4 == 2224


def test_lane_order_does_not_matter():
    """EMU_SHUFFLE runs the lanes of the emulated warp in a random order between two collectives: a
    shared-memory hand-over between lanes without a __syncwarp() would change results."""
    emu = emu_group()
    g = golden("root_19.npz")
    old = os.environ.get("EMU_SHUFFLE")
    try:
        for seed in ("5", "6"):
            os.environ["EMU_SHUFFLE"] = seed
            o, d, s, _ = emu_playouts(emu, 19, 8, [], 1, None, 32, seed_base=int(g["seed_base"]))
            assert (d == g["diffs"][:32]).all()
    finally:
        if old is None:
            os.environ.pop("EMU_SHUFFLE", None)
        else:
            os.environ["EMU_SHUFFLE"] = old


def test_move_cap_is_a_fault():
    emu = emu_group()
    o, d, s, _ = emu_playouts(emu, 9, 8, [], 1, None, 8, max_moves=10)
    assert s[2] == 8 and (d == -32768).all() and o.sum() == 0


def test_memory_bounds_under_address_sanitizer(tmp_path):
    """The same source under -fsanitize=address, the warp's shared memory and its empty-point arrays
    allocated exactly: an index that leaves them (on the GPU: into a neighbour's board, or past the
    block's shared memory — an illegal address only when the block fills the SM) aborts the run."""
    import subprocess
    from util import ROOT
    csrc = os.path.join(ROOT, "agentgo_b200", "csrc")
    exe = str(tmp_path / "emu_asan")
    srcs = [os.path.join(ROOT, "tests", "csrc", f) for f in ("emu_main.cpp", "emu_group.cpp", "simt_emu.cpp")] + \
           [os.path.join(csrc, f) for f in ("host_board.cpp", "host_layout.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address", "-w", "-I", csrc, "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "csrc"), "-I", os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")]
                          + srcs + ["-o", exe])
    for n, lanes, P in ((19, 8, 10), (9, 8, 7), (19, 16, 4)):
        r = subprocess.run([exe, str(n), str(lanes), str(P)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "ERROR" not in r.stderr, r.stderr[-1500:]
        assert ("playouts %d" % P) in r.stdout
