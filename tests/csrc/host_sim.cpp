// tests/csrc/host_sim.cpp — TEST-ONLY harness (never part of the product library).
// Runs the `__host__ __device__` playout logic of agentgo_b200/csrc/board_ops.h on the CPU,
// one playout at a time, so the sequential device logic can be checked against the oracle
// on a machine without a GPU.  The thread-per-playout kernel executes exactly this code.
#include <stdint.h>
#include <string.h>

#include "agentgo_b200.h"
#include "host_board.h"

using namespace agb;

extern "C" int sim_playouts(int n, const uint32_t* moves, int n_moves, int colour, int cand_row,
                            int cand_col, uint32_t position_id, uint32_t cand_idx, int64_t P,
                            uint64_t seed_base, int avrg_win, int16_t* diffs, int64_t* out3,
                            int64_t* stat2) {
    HostBoard hb(n);
    if (hb.replay(moves, n_moves)) return -1;
    Position start = hb.position();
    int first = colour;
    if (cand_row >= 0) {
        start = hb.with_move(colour, cand_row * n + cand_col);
        first = 3 - colour;
    } else {
        cand_idx = AGB_ROOT_CANDIDATE;
    }
    RngParams prm = {AGB_TINYMT_MAT1, AGB_TINYMT_MAT2, AGB_TINYMT_TMAT};
    Position work;
    for (int64_t p = 0; p < P; p++) {
        work = start;
        BoardRef<1, 0> bd;
        bd.w0 = work.w0; bd.w1 = work.w1;
        bd.load_scalars(work);
        Rng rng;
        rng.seed(playout_seed(seed_base, position_id, cand_idx, (uint64_t)p), prm);
        uint32_t mv = 0;
        int diff = bd.run_playout(rng, colour, first, 4096, &mv);
        if (diffs) diffs[p] = (int16_t)diff;
        int w = diff - avrg_win;
        if (w >= 0) { out3[0]++; out3[1] += w; } else out3[2] += w;
        stat2[0] += mv; stat2[1] += rng.draws;
    }
    return 0;
}
