"""GTP command surface: wire format of AgentGo/GTPAdapter.cpp (position-only engine on CPU)."""
import os
import subprocess

import pytest

import agentgo_b200 as ag

GTP = os.path.join(ag.HERE, "agentgo_gtp")


def gtp(cmds, *args):
    r = subprocess.run([GTP] + list(args), input="\n".join(cmds) + "\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_wire_format_without_gpu(built_lib):
    out = gtp(["name", "VERSION", "protocol_version", "list_commands", "", "komi 6.5", "boardsize 19",
               "clear_board", "play b D4", "play W q16", "play b pass", "play b D4", "play x D4",
               "frobnicate 1 2", "genmove x", "quit", "name"], "--device", "-1")
    exp = ("= AgentGo\n\n" "= 0.0.1\n\n" "= 2\n\n"
           "= protocol_version\nname\nversion\nlist_commands\nquit\nboardsize\nclear_board\nkomi\nplay\ngenmove\n\n"
           "= \n\n"            # komi
           "=\n\n"             # boardsize (no space, GTPAdapter.cpp:147)
           "=\n\n"             # clear_board
           "= \n\n"            # play b D4
           "= \n\n"            # play W q16 (colour and letter are case-insensitive)
           "= \n\n"            # play b pass: acknowledged, board untouched (GTPAdapter.cpp:107-113)
           "? unknown command : play b D4\n\n"      # occupied: the reference would AG_PANIC
           "? unknown command : play x D4\n\n"
           "? unknown command : frobnicate 1 2\n\n"
           "? unknown command : genmove x\n\n"
           "=\n\n")            # quit; nothing after it
    assert out == exp


def test_failed_genmove_is_an_error_not_a_pass(built_lib):
    """An engine that cannot run playouts (no device: --device -1; on a GPU a CUDA error or the move
    cap) answers GTP's failure line and leaves the board alone; it used to answer `= pass`."""
    out = gtp(["boardsize 9", "play b E5", "genmove w", "play w E5", "quit"], "--device", "-1")
    replies = [l for l in out.split("\n\n") if l]
    assert replies[0] == "=" and replies[1] == "= "
    assert replies[2].startswith("? genmove failed: ") and "no CPU fallback" in replies[2]
    assert replies[3] == "? unknown command : play w E5"     # E5 still holds the black stone


def test_reference_command_line_tunables(built_lib):
    """agentgo_gtp takes the reference's own command line: eight doubles (AgentGo.cpp:709-719)."""
    out = gtp(["name", "quit"], "--device", "-1", "1000", "1", "1", "40", "0.2166", "0.3", "0.00166", "0.01")
    assert out.startswith("= AgentGo")
    r = subprocess.run([GTP, "--device", "-1", "1", "2", "3"], input="quit\n", capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "eight tunables" in r.stderr


def test_coordinates_skip_letter_i(built_lib):
    # GaCharToInt / GaIntToChar (GTPAdapter.cpp:19-20): a..h -> 0..7, j -> 8; the LETTER is the row
    e = ag.Engine(19, device=-1)
    out = gtp(["boardsize 19", "play b J1", "play w H19", "play b T3", "quit"], "--device", "-1")
    assert out.count("= \n\n") == 3


@pytest.mark.gpu
def test_genmove_and_selfplay_on_gpu(built_lib):
    out = gtp(["boardsize 9", "clear_board", "play b E5", "genmove w", "genmove b", "quit"],
              "--playouts", "64", "--root-playouts", "128")
    lines = [l for l in out.split("\n\n") if l]
    assert lines[0] == "=" and lines[1] == "=" and lines[2] == "= "
    for l in lines[3:5]:
        assert l.startswith("= ") and (l[2:] == "pass" or (l[2] in "ABCDEFGHJ" and 1 <= int(l[3:]) <= 9))
    r = subprocess.run([GTP, "--boardsize", "9", "--selfplay", "--playouts", "32", "--root-playouts", "64",
                        "--max-moves", "400"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    import json
    sp = json.loads(r.stdout.strip().splitlines()[-1])["selfplay"]
    assert sp["genmoves"] >= 20 and sp["device_ms_median"] > 0


@pytest.mark.gpu
def test_multi_gpu_selfplay_equals_single_gpu(built_lib):
    """agentgo_gtp --gpus 2 (one process, two engines, NCCL allreduce through agb_attach_nccl) must
    play exactly the game of --gpus 1: results do not depend on how playouts are partitioned."""
    if ag.lib().agb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    games = []
    for gpus in ("1", "2"):
        r = subprocess.run([GTP, "--boardsize", "9", "--selfplay", "--playouts", "48", "--root-playouts", "96",
                            "--max-moves", "60", "--gpus", gpus, "--amaf"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        games.append([l for l in r.stderr.splitlines() if l[:3].strip().isdigit()])
    assert len(games[0]) == 60 and games[0] == games[1]
    # the same with the reference's complete policy (book, static marks, AMAF tables split by the
    # stones a candidate leaves) at the reference's board size
    games = []
    for gpus in ("1", "2"):
        r = subprocess.run([GTP, "--boardsize", "13", "--selfplay", "--playouts", "30", "--root-playouts", "60",
                            "--max-moves", "50", "--gpus", gpus, "--reference-policy"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        games.append([l for l in r.stderr.splitlines() if l[:3].strip().isdigit()])
    assert len(games[0]) == 50 and games[0] == games[1]


@pytest.mark.gpu
def test_match_driver_on_gpu(built_lib):
    """Row f4 of SURVEY.md §8: two configurations under the referee of the reference's GA tuner
    (SimulateOneGame, GeneHatchery.cpp:438-568), scored by CalcResults; deterministic for a seed."""
    import json
    args = ["--boardsize", "9", "--playouts", "40", "--root-playouts", "100", "--match", "2",
            "--white-playouts", "4", "--seed", "5"]
    runs = []
    for _ in range(2):
        r = subprocess.run([GTP] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        runs.append(json.loads(r.stdout.strip().splitlines()[-1])["match"])
    m = runs[0]
    assert runs[0] == runs[1]
    assert m["games"] == 2 and len(m["dscore"]) == 2 and m["stones_played"] > 60
    assert all(-81 <= d <= 81 and d % 2 != 0 for d in m["dscore"])      # 81 points: black - rest is odd
    assert m["black_wins"] == sum(d > 0 for d in m["dscore"])
