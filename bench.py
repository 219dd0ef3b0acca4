#!/usr/bin/env python
"""bench.py — playouts/sec of the flat Monte-Carlo playout path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1|2|3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

--config 2 (default, the headline): a "step" is one evaluation of BASELINE.json configs[1]: 19x19
empty board, root mode (MyGame::testAvrgWin order, black first), 1 000 000 playouts per GPU,
TinyMT32 RFC 8682 parameters, seed_base 20151001 + step.  With N GPUs the job is N x 1M playouts
(weak scaling); playout ids are sharded g % N == rank and the int64 counters are combined by ONE
NCCL allreduce inside every step.
--config 1: configs[0], 9x9 empty board, all 81 points as candidates x 10 000 playouts (the
reference's CPU-runnable case), with the CPU figures at 1 thread and at 2*cores+2 threads.
--config 3: configs[2], 1024 seeded random-legal mid-game 19x19 positions x 100 000 playouts through
agb_playouts_batch; with N GPUs the positions are sharded and the table is gathered.
--config 4: configs[3], a GTP self-play game at 1 M playouts per move through the GTP engine
(agentgo_gtp --selfplay), device-timed genmove latency.

value     : playouts/s from CUDA-event time on the engine's stream (position already uploaded,
            kernel + collective), max over ranks.
e2e       : playouts/s through the C-ABI call with HOST buffers: positions are copied host->device
            and the counters device->host inside the timed region (barrier + synchronize on both
            sides, wall clock, max over ranks).
roofline  : the resource that binds the path — SM integer issue (algorithmic 7.0e5 int ops per
            19x19 playout against the lane-op rate MEASURED live by agb_measure_peaks) — with the
            shared-memory, issue-slot and HBM lines nested (the HBM line is ~0 by construction).
Extras on the config-2 line, outside the timed headline: `shard_check` (N > 1: a 65 536-playout
slice evaluated unsharded by rank 0 and sharded + all-reduced by the group must agree), `strong_16M`
(16 M playouts in total split over the N GPUs), `genmove_1M` (agb_genmove at a fixed budget of 1 M
playouts per move over 24 mid-game 19x19 positions).
--impl reference : the reference's own CPU implementation (oracle/_ref, the reference's Board.cpp
            compiled with g++; else the C port) on all host threads with the reference's thread
            policy 2*cores+2, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_json_fd = 1
UNIT = "playouts/s"
SEED_BASE = 20151001
# algorithmic work per 19x19 root-mode playout (SURVEY.md §8d / BASELINE.md §2); 9x9: measured ratio
W_INT_OPS = {19: 7.0e5, 9: 2.6e5}
W_SMEM_BYTES = {19: 4.4e5, 9: 1.6e5}
W_HBM_BYTES = 4.0
L2_FLUSH_BYTES = 256 << 20
KERNELS = ["auto", "thread", "warp", "group8", "group16", "group32"]


def workload(cfg_id, per_gpu, n_gpus, extra=None):
    """The `config` object of the JSON line; byte-identical for both arms."""
    if cfg_id == 1:
        d = {"workload": "BASELINE configs[0]: 9x9 empty board, black to move, all 81 points as candidates, "
                         "10000 playouts per candidate (candidate mode: the candidate is played, white moves first), "
                         "TinyMT32 RFC8682, seed_base %d+step" % SEED_BASE,
             "board_size": 9, "playouts_per_step": 810000}
    elif cfg_id == 3:
        d = {"workload": "BASELINE configs[2]: %d seeded random-legal mid-game 19x19 positions (40 + k mod 200 moves of the "
                         "policy's own getRandomPiece from the empty board) x %d root-mode playouts each, "
                         "agb_playouts_batch, TinyMT32 RFC8682, seed_base %d+step" % (extra["B"], extra["P"], SEED_BASE),
             "board_size": 19, "playouts_per_step": extra["B"] * extra["P"], "positions": extra["B"]}
    elif cfg_id == 4:
        d = {"workload": "BASELINE configs[3]: GTP self-play 19x19 through agentgo_gtp, 1000000 playouts per move split "
                         "over the candidates + 100000 root playouts, game to two passes",
             "board_size": 19, "playouts_per_step": 1100000}
    else:
        d = {"workload": "19x19 empty board, root mode (black first), %d playouts per GPU per step, "
                         "TinyMT32 RFC8682, seed_base %d+step" % (per_gpu, SEED_BASE),
             "board_size": 19, "playouts_per_step": per_gpu * n_gpus, "playouts_per_gpu": per_gpu}
    d["parallelism"] = ("playout ids sharded g %% %d == rank, one NCCL allreduce of 3 int64 per step" % n_gpus
                        if cfg_id in (1, 2, 4) else
                        "positions sharded id %% %d == rank, one NCCL allgather of the [positions x 3] int64 table" % n_gpus) \
        if n_gpus > 1 else "single GPU"
    d["l2"] = "L2 flushed between timed steps (256 MiB write); inputs are a few KB per position"
    return d


def metric_name(cfg_id):
    return {1: "9x9 playouts/sec", 2: "19x19 playouts/sec", 3: "19x19 playouts/sec (mid-game batch)",
            4: "19x19 playouts/sec (GTP self-play)"}[cfg_id]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v == "Active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path (the one place outside tests/ and smoke() that executes oracle/)

def cpu_reference(cfg_id, sample_seconds, n_steps=1, warmup=0, threads=None):
    """Times the reference's CPU implementation of the configuration's workload on this box's host
    cores, on a bounded sample.  Returns (list of (rate, seconds) per step, info)."""
    from oracle import pyoracle as po
    n = 9 if cfg_id == 1 else 19
    kind = "reference" if po.ref_available(n) else "port"
    cores = os.cpu_count() or 1
    if threads is None:
        threads = 2 * cores + 2                 # TinyMT.h:543-546
    if cfg_id == 1:
        cands = [(i // 9, i % 9) for i in range(81)]
        b = po.OracleBoard(9, kind)

        def run(P, i):
            b.playouts(1, cands, P, seed_base=SEED_BASE + i, want_diffs=False, threads=threads)
            return 81 * P
        probe_P, what = 50, "all 81 candidates x %d playouts (candidate mode, 9x9 empty board)"
    elif cfg_id == 3:
        import agentgo_b200 as ag
        pos = midgame_positions(ag, 200)[::25]          # 8 positions spread over the batch's 40..239 moves
        boards = [po.OracleBoard(19, kind).replay(mv) for mv, _ in pos]

        def run(P, i):
            for k, ob in enumerate(boards):
                ob.playouts(pos[k][1], None, P, seed_base=SEED_BASE + i, position_id=25 * k, want_diffs=False, threads=threads)
            return len(boards) * P
        probe_P, what = 1000, "positions 0, 25, ..., 175 of the batch x %d root-mode playouts each"
    else:
        b = po.OracleBoard(19, kind)

        def run(P, i):
            b.playouts(1, None, P, seed_base=SEED_BASE + i, want_diffs=False, threads=threads)
            return P
        probe_P, what = 4000, "%d root-mode 19x19 playouts per step (same seeds as the GPU arm's first ids)"
    t0 = time.perf_counter()
    done = run(probe_P, 0)
    rate = done / (time.perf_counter() - t0)
    P = max(probe_P, int(probe_P * rate * sample_seconds / done))
    rates, total = [], 0
    for i in range(warmup + n_steps):
        t0 = time.perf_counter()
        total = run(P, i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            rates.append((total / dt, dt))
    info = {"kind": kind, "cores": cores, "threads": threads,
            "sample": (what % P) + ", std::thread pool of %d threads on %d host cores" % (threads, cores),
            "sample_playouts": total}
    return rates, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg_id = args.config if args.config != 4 else 2     # self-play has no CPU arm of its own: the playout rate
    rates, info = cpu_reference(cfg_id, sample_seconds=3.0, n_steps=args.steps, warmup=args.warmup)
    total_p = info["sample_playouts"] * len(rates)
    total_t = sum(dt for _, dt in rates)
    value = total_p / total_t
    cfg = workload(args.config, args.playouts, args.gpus, {"B": args.positions, "P": args.batch_playouts})
    line = {"metric": metric_name(args.config), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * cfg["playouts_per_step"] / value, "sampled": True,
            "ms_per_step_note": "extrapolated from the sample's rate to the configuration's playouts_per_step",
            "ms_per_sample": 1e3 * total_t / len(rates),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": cfg, "impl": "reference",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "threads": info["threads"],
                             "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------

def midgame_positions(ag, B):
    """configs[2]: position k = 40 + (k mod 200) moves from the empty board, black first, each move drawn
    by the policy's own getRandomPiece on the host (agb_random_piece), seeded from (seed_base+1, k, move)."""
    eng = ag.Engine(19, device=-1)
    out = []
    import numpy as np
    for k in range(B):
        eng.clear_board()
        mv = []
        colour = 1
        for t in range(40 + (k % 200)):
            p = eng.random_piece(colour, ag.playout_seed(SEED_BASE + 1, k, 0, t))
            if p is None:
                eng.pass_(colour)
                mv.append((colour << 16) | (0xFF << 8) | 0xFF)
            else:
                eng.play(colour, p[0], p[1])
                mv.append((colour << 16) | (p[0] << 8) | p[1])
            colour = 3 - colour
        out.append((np.array(mv, dtype=np.uint32), colour))
    eng.close()
    return out


def genmove_positions():
    """24 mid-game 19x19 positions from the committed fixture tests/golden/midgame_19.npz: its 16 move
    lists (40..235 moves) and the first halves of the eight longest ones."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "midgame_19.npz"))
    lists, o = [], 0
    for ln in g["lens"]:
        lists.append(g["moves"][o:o + int(ln)].astype(np.uint32))
        o += int(ln)
    pos = list(lists) + [mv[:len(mv) // 2] for mv in lists[8:]]
    return [(mv, 3 - ((int(mv[-1]) >> 16) & 0xFF)) for mv in pos]


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * (len(v) - 1) + 0.5))] if v else None


def run_b200(args):
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner) get stderr
    global _json_fd
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import agentgo_b200 as ag
    from agentgo_b200 import dist as agdist

    rank, world, local_rank = agdist.env_rank_world()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.stderr.write("bench.py: --gpus %d needs torchrun (one process per GPU)\n" % args.gpus)
            return 2
        args.gpus = world
    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device; the product has no CPU fallback\n")
        return 3
    torch.cuda.set_device(local_rank)
    tdist = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg_id = args.config
    board = 9 if cfg_id == 1 else 19
    kernel = {"auto": ag.KERNEL_AUTO, "thread": ag.KERNEL_THREAD, "warp": ag.KERNEL_WARP, "group8": ag.KERNEL_GROUP8,
              "group16": ag.KERNEL_GROUP16, "group32": ag.KERNEL_GROUP32}[args.kernel]
    eng = ag.Engine(board, device=local_rank, kernel=kernel, shard_rank=rank, shard_count=world)
    if world > 1:
        agdist.attach_allreduce(eng)
        agdist.attach_allgather(eng)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def bracket():
        torch.cuda.synchronize()
        if tdist is not None:
            tdist.barrier()
        torch.cuda.synchronize()

    # one pass of the configuration's workload through the C ABI with host buffers in and out
    if cfg_id == 1:
        cands = [(i // 9, i % 9) for i in range(81)]
        P_step = 81 * 10000
        h2d = 81 * ((121 + 41) * 4 + 16 + 20)
        d2h = 81 * 3 * 8 + 8 * 8

        def evaluate(i, small=False):
            w, sp, sn, _, st = eng.playouts(ag.BLACK, cands, 200 if small else 10000, seed_base=SEED_BASE + i)
            return st, (int(w.sum()), int(sp.sum()), int(sn.sum()))
    elif cfg_id == 3:
        B, PB = args.positions, args.batch_playouts
        positions = midgame_positions(ag, B)
        P_step = B * PB
        h2d = sum(len(mv) for mv, _ in positions) * 4 + B * 16
        d2h = B * 3 * 8 + 8 * 8

        def evaluate(i, small=False):
            if small:
                w, sp, sn, _, st = eng.playouts_batch(positions[:8 * world], 1000, seed_base=SEED_BASE + i)
            else:
                w, sp, sn, _, st = eng.playouts_batch(positions, PB, seed_base=SEED_BASE + i)
            return st, (int(w.sum()), int(sp.sum()), int(sn.sum()))
    else:
        P_step = args.playouts * world            # whole-job playouts per step
        h2d = ((441 + 181) * 4 + 16 + 20) * world     # GroupPosition (board words, empty-point array, scalars) + StartDesc
        d2h = (3 * 8 + 8 * 8) * world

        def evaluate(i, small=False):
            w, sp, sn, _, st = eng.playouts(ag.BLACK, None, P_step, seed_base=SEED_BASE + i)
            return st, (int(w[0]), int(sp[0]), int(sn[0]))

    def step(i, small=False):
        flush.fill_(i & 0xFF)            # L2 flush, outside the timed bracket
        bracket()
        t0 = time.perf_counter()
        st, counts = evaluate(i, small)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        bracket()
        return wall, st, counts

    for i in range(args.warmup):
        step(i, small=cfg_id == 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    walls, devms, launches, counts, local_playouts = [], [], 0, None, 0
    for i in range(args.steps):
        wall, st, counts = step(args.warmup + i)
        walls.append(wall)
        devms.append(st["device_ms"])
        launches += st["kernel_launches"]
        local_playouts += st["playouts"]
        assert st["faults"] == 0
        if world == 1:
            assert st["playouts"] == P_step and counts[0] <= P_step
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([sum(walls), sum(devms) * 1e-3, float(launches), float(local_playouts)], dtype=torch.float64, device=dev)
    if tdist is not None:
        tmax = t.clone(); tdist.all_reduce(tmax, op=tdist.ReduceOp.MAX)
        tsum = t.clone(); tdist.all_reduce(tsum, op=tdist.ReduceOp.SUM)
        wall_total, dev_total, launches_total, played = tmax[0].item(), tmax[1].item(), int(tsum[2].item()), int(tsum[3].item())
    else:
        wall_total, dev_total, launches_total, played = t[0].item(), t[1].item(), int(t[2].item()), int(t[3].item())
    K = args.steps
    assert played == K * P_step, "the ranks together played %d playouts, the job is %d" % (played, K * P_step)

    extras = {}
    if cfg_id == 2:
        # ---- shard_check: the all-reduced counters of a sharded evaluation equal the unsharded ones ----
        if world > 1:
            w, sp, sn, _, _ = eng.playouts(ag.BLACK, None, 65536, seed_base=SEED_BASE - 1)
            got = (int(w[0]), int(sp[0]), int(sn[0]))
            ok = True
            if rank == 0:
                solo = ag.Engine(board, device=local_rank, kernel=kernel)
                w1, sp1, sn1, _, _ = solo.playouts(ag.BLACK, None, 65536, seed_base=SEED_BASE - 1)
                solo.close()
                ok = got == (int(w1[0]), int(sp1[0]), int(sn1[0]))
            flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=dev)
            tdist.all_reduce(flag, op=tdist.ReduceOp.SUM)
            extras["shard_check"] = "ok" if flag.item() == 0 else "MISMATCH"
            extras["shard_check_counts"] = {"wins": got[0], "sum_pos": got[1], "sum_neg": got[2], "playouts": 65536}
        # ---- strong scaling: 16 M playouts in total, split over the GPUs ----
        if not args.no_extras:
            bracket()
            t0 = time.perf_counter()
            _, _, _, _, st = eng.playouts(ag.BLACK, None, 16000000, seed_base=SEED_BASE + 1000)
            torch.cuda.synchronize()
            wl = time.perf_counter() - t0
            bracket()
            t16 = torch.tensor([st["device_ms"], wl * 1e3], dtype=torch.float64, device=dev)
            if tdist is not None:
                tdist.all_reduce(t16, op=tdist.ReduceOp.MAX)
            extras["strong_16M"] = {"playouts": 16000000, "n_gpus": world, "device_ms": t16[0].item(), "wall_ms": t16[1].item(),
                                    "playouts_per_s": 16e6 / (t16[0].item() * 1e-3),
                                    "limits": "kernel time of 16M/N playouts per GPU + one allreduce (24 B); the last wave of "
                                              "resident playouts (76 per SM) and the launch are what keeps it from 1/N"}
            # ---- genmove latency at a fixed budget of 1 M playouts per move (configs[3]) ----
            gm_dev, gm_wall, gm_play = [], [], []
            for k, (mv, colour) in enumerate(genmove_positions()):
                eng.clear_board()
                eng.replay(mv)
                bracket()
                t0 = time.perf_counter()
                _, st = eng.genmove(colour, P_per_candidate=-1000000, P_root=65536, seed_base=SEED_BASE + 2000 + k)
                wl = time.perf_counter() - t0
                gm_dev.append(st["device_ms"]); gm_wall.append(wl * 1e3); gm_play.append(st["playouts"])
            eng.clear_board()
            tg = torch.tensor([gm_dev, gm_wall], dtype=torch.float64, device=dev)
            if tdist is not None:
                tdist.all_reduce(tg, op=tdist.ReduceOp.MAX)
            gd, gw = tg[0].tolist(), tg[1].tolist()
            extras["genmove_1M"] = {"positions": len(gd), "n_gpus": world, "budget": "1000000 playouts split over the candidates "
                                    "+ 65536 root playouts (avrg_win)", "device_ms_median": pct(gd, 0.5), "device_ms_p95": pct(gd, 0.95),
                                    "wall_ms_median": pct(gw, 0.5), "wall_ms_p95": pct(gw, 0.95),
                                    "limits": "two kernels (root 65536, candidates 1M) of 1/N of the playouts per GPU, each followed by its "
                                              "allreduce; the last wave of resident playouts of each kernel"}

    if rank != 0:
        if tdist is not None:
            tdist.destroy_process_group()
        return 0

    value = K * P_step / dev_total
    e2e = K * P_step / wall_total
    per_gpu_rate = value / world
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    mp = ag.measure_peaks(local_rank)
    traffic, winst, kname = None, None, None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof) and cfg_id == 2 and args.kernel == "auto":
        pj = json.load(open(prof))
        traffic, winst, kname = pj.get("dram_bytes_per_launch"), pj.get("warp_instructions_per_playout"), pj.get("kernel")
    launch_ms = 1e3 * dev_total / K
    hbm_achieved = (P_step / world) * W_HBM_BYTES / (launch_ms * 1e-3) / 1e9
    roofline = {
        "bound": "int_issue", "achieved": per_gpu_rate * W_INT_OPS[board], "peak": mp["int_ops_per_s"],
        "unit": "int32 lane-ops/s", "frac": per_gpu_rate * W_INT_OPS[board] / mp["int_ops_per_s"],
        "traffic": traffic, "algorithmic_ops_per_playout": W_INT_OPS[board],
        "peak_source": "agb_measure_peaks, live (mixed LOP3/IADD3/IMAD micro-kernel; profiles/r02_peak_kernels.txt)",
        "peak_alu_pipe_only": mp["int_ops_per_s_alu_only"],
        "kernel": kname or ("playout_group_kernel<%d>, 8 lanes per playout" % board if args.kernel in ("auto", "group8") else args.kernel),
        "launch_ms": launch_ms,
        "note": "the path is SM integer-issue bound, then shared-memory bound; not HBM or tensor (SURVEY.md §8d)",
        "smem": {"achieved": per_gpu_rate * W_SMEM_BYTES[board] / 1e9, "peak": mp["smem_bytes_per_s"] / 1e9, "unit": "GB/s",
                 "frac": per_gpu_rate * W_SMEM_BYTES[board] / mp["smem_bytes_per_s"],
                 "algorithmic_bytes_per_playout": W_SMEM_BYTES[board], "peak_source": "agb_measure_peaks, live"},
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "traffic": traffic, "peak_source": hbm_src, "algorithmic_bytes_per_playout": W_HBM_BYTES,
                "note": "reported because the contract asks for it; ~0 by construction"},
    }
    if winst and clocks and clocks.get("sm_mhz"):
        # issued side: warp-instructions per playout from ncu (profiles/), against 4 issue slots
        # per clock per SM at the SM clock sampled during this run
        peak_issue = mp["sm_count"] * 4 * clocks["sm_mhz"] * 1e6
        roofline["issue_slots"] = {"achieved": per_gpu_rate * winst, "peak": peak_issue, "unit": "warp-instructions/s",
                                   "frac": per_gpu_rate * winst / peak_issue, "warp_instructions_per_playout": winst,
                                   "source": "ncu smsp__inst_executed.sum (profiles/roofline_traffic.json)"}
        if pj.get("alu_pipe_active_pct"):
            # what binds the issue rate: the ALU pipe (logic, compares, shifts, selects) takes one warp
            # instruction every second cycle per scheduler; utilisation from the ncu capture of the same kernel
            roofline["alu_pipe"] = {"frac": pj["alu_pipe_active_pct"] / 100.0, "alu_share_of_instructions": pj.get("alu_share_of_instructions"),
                                    "source": "ncu sm__pipe_alu_cycles_active (profiles/r02_group8_metrics.txt, 100 k playouts); "
                                              "opcode shares: profiles/r02_group8_opcodes.txt"}
            roofline["note"] = ("the path is SM integer-issue bound — the ALU pipe first, then the issue slots, then shared memory; "
                                "not HBM or tensor (SURVEY.md §8d)")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rates, info = cpu_reference(cfg_id, sample_seconds=args.cpu_seconds)
        cpu = {"value": rates[0][0], "unit": UNIT, "cores": info["cores"], "threads": info["threads"],
               "kind": info["kind"], "sample": info["sample"], "seconds": rates[0][1]}
        if cfg_id == 1:
            r1, i1 = cpu_reference(cfg_id, sample_seconds=args.cpu_seconds / 2, threads=1)
            cpu["one_thread"] = {"value": r1[0][0], "unit": UNIT, "threads": 1, "sample": i1["sample"]}
    line = {"metric": metric_name(cfg_id), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": workload(cfg_id, args.playouts, world, {"B": args.positions, "P": args.batch_playouts}),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * wall_total / K},
            "gpu_launches": launches_total, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "last_counts": {"wins": counts[0], "sum_pos": counts[1], "sum_neg": counts[2]},
            "impl": "b200"}
    line.update(extras)
    os.write(_json_fd, (json.dumps(line) + "\n").encode())
    if tdist is not None:
        tdist.destroy_process_group()
    return 4 if extras.get("shard_check") == "MISMATCH" else 0


def run_selfplay(args):
    """configs[3] through the GTP engine itself: agentgo_gtp --selfplay at 1 M playouts per move."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import agentgo_b200 as ag
    ag.build_lib()
    gtp = os.path.join(ROOT, "agentgo_b200", "agentgo_gtp")
    cmd = [gtp, "--boardsize", "19", "--total-playouts", "1000000", "--root-playouts", "100000", "--selfplay",
           "--gpus", str(args.gpus), "--kernel", args.kernel, "--max-moves", str(args.selfplay_moves)]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if out.returncode != 0:
        sys.stderr.write(out.stderr[-2000:])
        return out.returncode
    sp = json.loads(out.stdout.strip().splitlines()[-1])["selfplay"]
    # playouts per second of device time, from the median genmove (the engine reports the median, not the sum)
    value = sp["playouts"] / max(1e-9, sp["genmoves"] * sp["device_ms_median"] * 1e-3)
    line = {"metric": metric_name(4), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": sp["genmoves"], "warmup": 0,
            "ms_per_step": sp["device_ms_median"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload(4, args.playouts, args.gpus),
            "e2e": {"value": sp["playouts"] / wall, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "note": "whole process wall clock incl. start-up"},
            "genmove": sp, "impl": "b200", "command": " ".join(cmd[1:])}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4])
    ap.add_argument("--playouts", type=int, default=1000000, help="config 2: playouts per GPU per step")
    ap.add_argument("--positions", type=int, default=1024, help="config 3: positions in the batch")
    ap.add_argument("--batch-playouts", type=int, default=100000, help="config 3: playouts per position")
    ap.add_argument("--selfplay-moves", type=int, default=800, help="config 4: genmove cap")
    ap.add_argument("--kernel", default="auto", choices=KERNELS)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="config 2: skip strong_16M and genmove_1M")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1 if args.config == 3 else 5
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_selfplay(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
