// capi.cpp — extern "C" surface declared in include/agentgo_b200.h.
// No C++ type or exception crosses this boundary.
#include <string.h>

#include <new>
#include <string>

#include "agentgo_b200.h"
#include "engine.h"
#include "kernel_args.h"

using agb::Engine;

static thread_local std::string g_create_error;

#define AGB_GUARD_BEGIN try {
#define AGB_GUARD_END(h)                                                        \
    }                                                                           \
    catch (const std::bad_alloc&) {                                             \
        if (h) (h)->e->err = "out of host memory";                              \
        return AGB_ERR_NOMEM;                                                   \
    }                                                                           \
    catch (...) {                                                               \
        if (h) (h)->e->err = "unexpected C++ exception";                        \
        return AGB_ERR_STATE;                                                   \
    }

extern "C" {

void agb_default_config(agb_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->board_size = 13;            /* the reference's BOARD_SIZE, GoContest/AgentGo.h:15 */
    cfg->device = 0;
    cfg->kernel = AGB_KERNEL_AUTO;
    cfg->shard_rank = 0;
    cfg->shard_count = 1;
    cfg->mat1 = AGB_TINYMT_MAT1;
    cfg->mat2 = AGB_TINYMT_MAT2;
    cfg->tmat = AGB_TINYMT_TMAT;
    cfg->max_moves = 0;
    /* AgentGo.cpp:18-25 */
    cfg->index_a = 1000; cfg->index_b = 1; cfg->index_c = 1;
    cfg->index_amaf_a1 = 40; cfg->index_amaf_a2 = 39. / 180; cfg->index_amaf_b1 = 0.3; cfg->index_amaf_b2 = 0.3 / 180;
    cfg->index_cnsv = 0.01;
}

const char* agb_version(void) { return "agentgo_b200 0.1 (sm_100a)"; }

int agb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int agb_create(const agb_config* cfg, agb_engine** out) {
    if (!cfg || !out) return AGB_ERR_INVALID;
    *out = nullptr;
    try {
        Engine* e = nullptr;
        int st = Engine::create(*cfg, &e, &g_create_error);
        if (st) return st;
        agb_engine* h = new agb_engine;
        h->e = e;
        *out = h;
        return AGB_OK;
    } catch (...) {
        g_create_error = "out of host memory";
        return AGB_ERR_NOMEM;
    }
}

void agb_destroy(agb_engine* h) {
    if (!h) return;
    delete h->e;
    delete h;
}

const char* agb_last_error(const agb_engine* h) {
    if (!h) return g_create_error.c_str();
    return h->e->err.c_str();
}

int agb_boardsize(agb_engine* h, int n) {
    if (!h) return AGB_ERR_INVALID;
    if (!agb::HostBoard::size_supported(n)) return h->e->fail(AGB_ERR_INVALID, "unsupported board size");
    if (h->e->busy()) return h->e->fail(AGB_ERR_STATE, "a launch is in flight: finish it before changing the position");
    h->e->board.reset(n);
    h->e->genmove_step = 0;
    return AGB_OK;
}

int agb_clear_board(agb_engine* h) {
    if (!h) return AGB_ERR_INVALID;
    if (h->e->busy()) return h->e->fail(AGB_ERR_STATE, "a launch is in flight: finish it before changing the position");
    h->e->board.clear();
    h->e->genmove_step = 0;                 /* MyGame::onClear, AgentGo.cpp:377 */
    return AGB_OK;
}

int agb_play(agb_engine* h, int colour, int row, int col) {
    if (!h) return AGB_ERR_INVALID;
    if (h->e->busy()) return h->e->fail(AGB_ERR_STATE, "a launch is in flight: finish it before changing the position");
    int st = h->e->board.play(colour, row, col);
    if (st == -1) return h->e->fail(AGB_ERR_INVALID, "illegal colour or coordinates");
    if (st == -2) return h->e->fail(AGB_ERR_OCCUPIED, "point occupied");
    return AGB_OK;
}

int agb_pass(agb_engine* h, int colour) {
    if (!h) return AGB_ERR_INVALID;
    if (colour != AGB_BLACK && colour != AGB_WHITE) return h->e->fail(AGB_ERR_INVALID, "bad colour");
    if (h->e->busy()) return h->e->fail(AGB_ERR_STATE, "a launch is in flight: finish it before changing the position");
    h->e->board.pass(colour);
    return AGB_OK;
}

int agb_get_board(const agb_engine* h, uint8_t* colours, int* n) {
    if (!h) return AGB_ERR_INVALID;
    const int N = h->e->board.n();
    if (n) *n = N;
    if (colours)
        for (int i = 0; i < N * N; i++) colours[i] = (uint8_t)h->e->board.colour_at(i);
    return AGB_OK;
}

int agb_check(const agb_engine* h, int which, int colour, int row, int col) {
    if (!h) return AGB_ERR_INVALID;
    int r = h->e->board.check(which, colour, row, col);
    return r < 0 ? AGB_ERR_INVALID : r;
}

int agb_count_score(const agb_engine* h, int colour) {
    if (!h || (colour != AGB_BLACK && colour != AGB_WHITE)) return AGB_ERR_INVALID;
    return h->e->board.count_score(colour);
}

int agb_export_state(const agb_engine* h, int32_t* out, int cap) {
    if (!h || !out) return AGB_ERR_INVALID;
    return h->e->board.export_state(out, cap);
}

int agb_playouts_launch(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P,
                        int avrg_win, uint64_t seed_base, uint32_t position_id, int want_diffs) {
    if (!h) return AGB_ERR_INVALID;
    AGB_GUARD_BEGIN
    return h->e->playouts_launch(colour, cand_rc, C, P, avrg_win, seed_base, position_id, want_diffs != 0);
    AGB_GUARD_END(h)
}

int agb_playouts_finish(agb_engine* h, int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                        int16_t* per_playout_diff, agb_stats* stats) {
    if (!h) return AGB_ERR_INVALID;
    AGB_GUARD_BEGIN
    return h->e->finish(wins, sum_pos, sum_neg, per_playout_diff, stats);
    AGB_GUARD_END(h)
}

int agb_playouts(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P, int avrg_win,
                 uint64_t seed_base, uint32_t position_id, int64_t* wins, int64_t* sum_pos,
                 int64_t* sum_neg, int16_t* per_playout_diff, agb_stats* stats) {
    int st = agb_playouts_launch(h, colour, cand_rc, C, P, avrg_win, seed_base, position_id,
                                 per_playout_diff != nullptr);
    if (st) return st;
    return agb_playouts_finish(h, wins, sum_pos, sum_neg, per_playout_diff, stats);
}

int agb_playouts_amaf(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P, int avrg_win,
                      uint64_t seed_base, uint32_t position_id, int64_t* wins, int64_t* sum_pos,
                      int64_t* sum_neg, int64_t* amaf, agb_stats* stats) {
    if (!h || !amaf) return AGB_ERR_INVALID;
    AGB_GUARD_BEGIN
    int st = h->e->playouts_launch(colour, cand_rc, C, P, avrg_win, seed_base, position_id, false, true);
    if (st) return st;
    return h->e->finish(wins, sum_pos, sum_neg, nullptr, stats, amaf);
    AGB_GUARD_END(h)
}

void* agb_counters_dev(agb_engine* h) { return h ? h->e->counters_dev() : nullptr; }
void* agb_stream(agb_engine* h) { return h ? (void*)h->e->stream() : nullptr; }

int agb_set_allreduce(agb_engine* h, agb_allreduce_fn fn, void* ctx) {
    if (!h) return AGB_ERR_INVALID;
    h->e->set_allreduce((agb::AllreduceFn)fn, ctx);
    return AGB_OK;
}

int agb_set_allgather(agb_engine* h, agb_allgather_fn fn, void* ctx) {
    if (!h) return AGB_ERR_INVALID;
    h->e->set_allgather((agb::AllgatherFn)fn, ctx);
    return AGB_OK;
}

int agb_playouts_batch(agb_engine* h, const agb_position* pos, int B, int64_t P, uint64_t seed_base,
                       uint32_t first_position_id, int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                       int16_t* per_playout_diff, agb_stats* stats) {
    if (!h) return AGB_ERR_INVALID;
    AGB_GUARD_BEGIN
    return h->e->playouts_batch(pos, B, P, seed_base, first_position_id, wins, sum_pos, sum_neg,
                                per_playout_diff, stats);
    AGB_GUARD_END(h)
}

int agb_calc_result(const agb_engine* h, int* black_minus_white) {
    if (!h || !black_minus_white) return AGB_ERR_INVALID;
    *black_minus_white = h->e->board.calc_result();
    return AGB_OK;
}

int agb_random_piece(agb_engine* h, int colour, uint32_t seed, int* row, int* col, int* is_pass) {
    if (!h || !row || !col || !is_pass) return AGB_ERR_INVALID;
    if (colour != AGB_BLACK && colour != AGB_WHITE) return h->e->fail(AGB_ERR_INVALID, "bad colour");
    const agb_config& c = h->e->cfg;
    agb::RngParams prm = {c.mat1 ? c.mat1 : AGB_TINYMT_MAT1, c.mat2 ? c.mat2 : AGB_TINYMT_MAT2, c.tmat ? c.tmat : AGB_TINYMT_TMAT};
    const int idx = h->e->board.random_piece(colour, seed, prm);
    const int n = h->e->board.n();
    *is_pass = idx < 0;
    *row = idx < 0 ? -1 : idx / n;
    *col = idx < 0 ? -1 : idx % n;
    return AGB_OK;
}

int agb_static_units(const agb_engine* h, int colour, int row, int col, int64_t* units_a, int64_t* units_b) {
    if (!h || !units_a || !units_b) return AGB_ERR_INVALID;
    const int n = h->e->board.n();
    if ((colour != AGB_BLACK && colour != AGB_WHITE) || row < 0 || col < 0 || row >= n || col >= n)
        return AGB_ERR_INVALID;
    if (h->e->board.colour_at(row * n + col) != 0) return AGB_ERR_OCCUPIED;
    h->e->board.static_units(colour, row, col, units_a, units_b);
    return AGB_OK;
}

int agb_last_marks(const agb_engine* h, double* total_mark, int cap) {
    if (!h || !total_mark) return AGB_ERR_INVALID;
    const int words = (int)h->e->last_marks.size();
    if (cap < words) return AGB_ERR_INVALID;
    memcpy(total_mark, h->e->last_marks.data(), sizeof(double) * (size_t)words);
    return words;
}

int agb_genmove_step(const agb_engine* h) { return h ? h->e->genmove_step : AGB_ERR_INVALID; }

int agb_set_genmove_step(agb_engine* h, int step) {
    if (!h || step < 0) return AGB_ERR_INVALID;
    h->e->genmove_step = step;
    return AGB_OK;
}

int agb_genmove(agb_engine* h, int colour, int64_t P_per_candidate, int64_t P_root, uint64_t seed_base,
                int* row, int* col, int* is_pass, agb_stats* stats) {
    if (!h || !row || !col || !is_pass) return AGB_ERR_INVALID;
    if (P_per_candidate == 0 || P_root <= 0) return h->e->fail(AGB_ERR_INVALID, "bad playout counts");
    AGB_GUARD_BEGIN
    return h->e->genmove(colour, P_per_candidate, P_root, seed_base, row, col, is_pass, stats);
    AGB_GUARD_END(h)
}

long long agb_debug_violations(void) { return (long long)agb::debug_violations(); }
long long agb_debug_event(int which) { return (long long)agb::debug_event(which); }

}  /* extern "C" */
