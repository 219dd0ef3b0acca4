/*
 * oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * C-callable harness around the REFERENCE's own Board class (GoContest/Board.cpp,
 * SetNode.cpp, Piece.cpp), compiled by oracle/build_ref.sh as one translation unit:
 *     ref_shim.h + <reference sources, streamed from /root/reference> + this file.
 * Nothing of the reference is copied into the repo; only the built
 * oracle/_ref/liboracle_ref_<N>.so exists (git-ignored).
 *
 * The two playout loops are restated from AgentGo.cpp:285-327 (MyWorker::work,
 * enemy moves first) and AgentGo.cpp:418-444 (MyGame::testAvrgWin, agent moves
 * first); they contain no arithmetic beyond calling Board methods.  AgentGo.cpp
 * itself cannot be compiled here (<Windows.h>, _tmain, Win32 threads).
 *
 * Exports the `orc_*` API shared with the C restatement (oracle/agb_oracle.c) so
 * tests drive both through one ctypes wrapper (oracle/pyoracle.py).
 */
#include "agentgo_b200.h"
#include "tinymt32.h"

static thread_local oracle_tinymt32_t g_rng;
static thread_local int64_t g_draws;

extern "C" int oracle_rand(void) {
    g_draws++;
    return (int)(oracle_tinymt32_generate_uint32(&g_rng) >> 17);
}

static void rng_seed(uint32_t seed) {
    g_rng.mat1 = AGB_TINYMT_MAT1;
    g_rng.mat2 = AGB_TINYMT_MAT2;
    g_rng.tmat = AGB_TINYMT_TMAT;
    oracle_tinymt32_init(&g_rng, seed);
}

#define NN (BOARD_SIZE * BOARD_SIZE)

extern "C" {

const char* orc_kind(void) { return "reference"; }
int orc_board_size(void) { return BOARD_SIZE; }
int orc_sizeof_board(void) { return (int)sizeof(Board); }

void* orc_new(int n) {
    if (n != BOARD_SIZE) return NULL;
    return new Board();
}
void orc_free(void* b) { delete (Board*)b; }
void orc_clear(void* b) { ((Board*)b)->clear(); }

int orc_put(void* vb, int agent, int row, int col) {
    Board* b = (Board*)vb;
    if (!Piece(agent, row, col).legal()) return -1;
    if (b->data[row][col] != GO_NULL) return -2;
    b->put(agent, row, col);
    return 0;
}
void orc_pass(void* b, int agent) { ((Board*)b)->pass(agent); }

int orc_check(void* vb, int which, int agent, int row, int col) {
    Board* b = (Board*)vb;
    switch (which) {
        case 0: return b->checkSuicide(agent, row, col);
        case 1: return b->checkKill(agent, row, col);
        case 2: return b->checkSurvive(agent, row, col);
        case 3: return b->checkDying(agent, row, col);
        case 4: return b->checkChase(agent, row, col);
        case 5: return b->checkTrueEye(agent, row, col);
        case 6: return b->checkCompete(agent, row, col);
        case 7: return b->checkNoSense(agent, row, col);
    }
    return -1;
}
int orc_count_score(void* b, int agent) { return ((Board*)b)->countScore(agent); }

/* canonical state dump, 16 + 7*NN int32 words (layout: DESIGN.md §3) */
int orc_export_state(void* vb, int32_t* out, int cap) {
    Board* b = (Board*)vb;
    const int words = 16 + 7 * NN;
    if (cap < words) return -1;
    for (int i = 0; i < words; i++) out[i] = 0;
    out[0] = BOARD_SIZE;
    out[1] = b->num_black;
    out[2] = b->num_white;
    out[3] = b->space_head;
    out[4] = b->exist_compete ? 1 : 0;
    out[5] = b->exist_compete ? b->compete[0] : 0;
    out[6] = b->exist_compete ? b->compete[1] : -1;
    out[7] = b->exist_compete ? b->compete[2] : -1;
    for (int a = 0; a < 2; a++) {
        out[8 + 3 * a] = b->last_move[a].agent;
        out[9 + 3 * a] = b->last_move[a].row;
        out[10 + 3 * a] = b->last_move[a].col;
    }
    out[14] = b->game_ending ? 1 : 0;
    int32_t* data = out + 16;
    int32_t* root = data + NN;
    int32_t* hp = root + NN;
    int32_t* size = hp + NN;
    int32_t* next = size + NN;
    int32_t* sprev = next + NN;
    int32_t* snext = sprev + NN;
    for (int idx = 0; idx < NN; idx++) {
        int c = b->data[idx / BOARD_SIZE][idx % BOARD_SIZE];
        data[idx] = c;
        next[idx] = -1;
        if (c == GO_NULL) {
            root[idx] = -1;
            sprev[idx] = b->space_list[idx][0];
            snext[idx] = b->space_list[idx][1];
        } else {
            root[idx] = b->set_nodes[idx].pnode;
            sprev[idx] = snext[idx] = -1;
        }
    }
    for (int idx = 0; idx < NN; idx++) {
        if (data[idx] == GO_NULL || root[idx] != idx) continue;
        SetNode* sn = &b->set_nodes[idx];
        hp[idx] = sn->hp;
        size[idx] = sn->size;
        for (int k = 0; k < sn->size; k++) {
            int p = sn->pieces[k].row * BOARD_SIZE + sn->pieces[k].col;
            if (k == 0 && p != idx) return -3; /* invariant: the root is the first piece */
            if (k + 1 < sn->size)
                next[p] = sn->pieces[k + 1].row * BOARD_SIZE + sn->pieces[k + 1].col;
        }
    }
    return words;
}

/* Play n_moves random moves alternately from the current position with ONE TinyMT32
 * stream (seed), starting with first_colour; moves_out[i] uses the AGB_MOVE encoding
 * (pass allowed).  Used to generate the mid-game fixtures.
 * The board is left in the played position but with game_ending as getRandomPiece set
 * it; fixtures are defined by the move list replayed with put/pass.               */
int orc_random_game(void* vb, int first_colour, int n_moves, uint32_t seed, uint32_t* moves_out) {
    Board* b = (Board*)vb;
    rng_seed(seed);
    int colour = first_colour;
    for (int i = 0; i < n_moves; i++) {
        Piece r = b->getRandomPiece(colour);
        if (!r.isEmpty()) {
            b->put(r);
            moves_out[i] = AGB_MOVE(colour, r.row, r.col);
        } else {
            b->pass(colour);
            moves_out[i] = AGB_MOVE_PASS(colour);
        }
        colour = 3 - colour;
    }
    return 0;
}

/* one playout from `start` on scratch board `bd`; returns diff.
 * first/second = order inside the pair loop (AgentGo.cpp:300/310 or :425/434). */
static int one_playout(Board* bd, const Board& start, int agent, int first, int second,
                       int64_t* moves) {
    bd->clone(start);
    bool first_go = 1, second_go = 1;
    while (first_go || second_go) {
        Piece r = bd->getRandomPiece(first);
        if (!r.isEmpty()) { bd->put(r); first_go = true; (*moves)++; }
        else { first_go = false; bd->pass(first); }
        r = bd->getRandomPiece(second);
        if (!r.isEmpty()) { bd->put(r); second_go = true; (*moves)++; }
        else { second_go = false; bd->pass(second); }
    }
    int enemy = 3 - agent;
    return bd->countScore(agent) - bd->countScore(enemy);
}

/* Playouts p = p_begin, p_begin+p_stride, ... (< p_end) for one candidate (or root mode
 * when cand_row < 0).  Accumulates into out3 = {wins,sum_pos,sum_neg} and
 * stat2 = {moves, rand draws}; diffs (nullable) is indexed by absolute p.          */
int orc_playouts(void* vroot, int colour, int cand_row, int cand_col, uint32_t position_id,
                 uint32_t cand_idx, int64_t p_begin, int64_t p_end, int64_t p_stride,
                 uint64_t seed_base, int avrg_win, int16_t* diffs, int64_t* out3,
                 int64_t* stat2) {
    Board* root = (Board*)vroot;
    int agent = colour, enemy = 3 - colour;
    Board start(*root);
    int first, second;
    if (cand_row >= 0) {
        if (start.data[cand_row][cand_col] != GO_NULL) return -2;
        start.put(agent, cand_row, cand_col);     /* AgentGo.cpp:83 */
        first = enemy; second = agent;            /* AgentGo.cpp:300,310 */
    } else {
        cand_idx = AGB_ROOT_CANDIDATE;
        first = agent; second = enemy;            /* AgentGo.cpp:425,434 */
    }
    Board* bd = new Board();
    int64_t wins = 0, sp = 0, sn = 0, moves = 0;
    g_draws = 0;
    for (int64_t p = p_begin; p < p_end; p += p_stride) {
        rng_seed(agb_playout_seed(seed_base, position_id, cand_idx, (uint64_t)p));
        int diff = one_playout(bd, start, agent, first, second, &moves);
        if (diffs) diffs[p] = (int16_t)diff;
        int w = diff - avrg_win;                  /* AgentGo.cpp:327 */
        if (w >= 0) { wins++; sp += w; } else sn += w;   /* sign test of AgentGo.cpp:328 */
    }
    delete bd;
    if (out3) { out3[0] += wins; out3[1] += sp; out3[2] += sn; }
    if (stat2) { stat2[0] += moves; stat2[1] += g_draws; }
    return 0;
}

/* AMAF first-play marks, restating AgentGo.cpp:286-338 (candidate mode only): per playout
 * mark[q] / mark2[q] = cstep of the first agent / enemy play at q (cstep starts at 1 and is
 * incremented at the top of every pair, so the first pair has cstep 2); after the playout, with
 * w = diff - avrg_win, point q contributes w to the agent table when 0 < mark[q] <= AMAF_RANGE,
 * else to the enemy table when 0 < mark2[q] <= AMAF_RANGE (the `else if` of :334).
 * Integer contract: amaf[which][sign][step][q] (int64, which 0 = agent / 1 = enemy first play,
 * sign 0 = sum of w >= 0 / 1 = sum of w < 0, step 0..AMAF_RANGE, q = row*N+col), accumulated
 * over all playouts; the reference's doubles are
 *   amaf1[q] =  sum_s 100*amaf[s-1]*index_c*(A[0][0][s][q] + cnsv*A[0][1][s][q])   (:332)
 *   amaf2[q] = -sum_s 100*amaf[s-1]*index_c*(A[1][0][s][q] + cnsv*A[1][1][s][q])   (:335)  */
int orc_playouts_amaf(void* vroot, int colour, int cand_row, int cand_col, uint32_t position_id,
                      uint32_t cand_idx, int64_t p_begin, int64_t p_end, int64_t p_stride,
                      uint64_t seed_base, int avrg_win, int64_t* out3, int64_t* amaf) {
    Board* root = (Board*)vroot;
    const int agent = colour, enemy = 3 - colour;
    if (cand_row < 0) return -1;
    Board start(*root);
    if (start.data[cand_row][cand_col] != GO_NULL) return -2;
    start.put(agent, cand_row, cand_col);                      /* AgentGo.cpp:83 */
    Board* bd = new Board();
    static thread_local int mark[NN], mark2[NN];
    const int S = AMAF_RANGE + 1;
    for (int64_t p = p_begin; p < p_end; p += p_stride) {
        rng_seed(agb_playout_seed(seed_base, position_id, cand_idx, (uint64_t)p));
        bd->clone(start);
        memset(mark, 0, sizeof(mark));
        memset(mark2, 0, sizeof(mark2));
        bool agent_go = 1, enemy_go = 1;
        int cstep = 1;
        while (agent_go || enemy_go) {
            cstep++;
            Piece r = bd->getRandomPiece(enemy);
            if (!r.isEmpty()) {
                bd->put(r); enemy_go = true;
                if (mark2[r.row * BOARD_SIZE + r.col] == 0) mark2[r.row * BOARD_SIZE + r.col] = cstep;
            } else { enemy_go = false; bd->pass(enemy); }
            r = bd->getRandomPiece(agent);
            if (!r.isEmpty()) {
                bd->put(r); agent_go = true;
                if (mark[r.row * BOARD_SIZE + r.col] == 0) mark[r.row * BOARD_SIZE + r.col] = cstep;
            } else { agent_go = false; bd->pass(agent); }
        }
        const int w = bd->countScore(agent) - bd->countScore(enemy) - avrg_win;
        const int sign = w >= 0 ? 0 : 1;
        if (out3) { if (w >= 0) { out3[0]++; out3[1] += w; } else out3[2] += w; }
        for (int q = 0; q < NN; q++) {
            if (mark[q] != 0 && mark[q] <= AMAF_RANGE) amaf[((0 * 2 + sign) * S + mark[q]) * NN + q] += w;
            else if (mark2[q] != 0 && mark2[q] <= AMAF_RANGE) amaf[((1 * 2 + sign) * S + mark2[q]) * NN + q] += w;
        }
    }
    delete bd;
    return 0;
}

/* CPU baseline: the reference's thread policy (2*cores+2 threads, TinyMT.h:543-546;
 * one job per candidate, AgentGo.cpp:595-600) re-created with std::thread.  Jobs are
 * (candidate, chunk) pairs pulled from an atomic counter so root mode also scales.
 * cand_rc: C pairs; C == 0 => root mode.  out3 is [max(C,1)][3], diffs [max(C,1)*P]. */
int orc_playouts_mt(void* vroot, int colour, const int16_t* cand_rc, int C, int64_t P,
                    uint32_t position_id, uint64_t seed_base, int avrg_win, int n_threads,
                    int16_t* diffs, int64_t* out3, int64_t* stat2) {
    const int nc = C > 0 ? C : 1;
    if (n_threads <= 0) n_threads = 2 * (int)std::thread::hardware_concurrency() + 2;
    const int64_t chunk = 256;
    const int64_t chunks_per_cand = (P + chunk - 1) / chunk;
    const int64_t n_jobs = chunks_per_cand * nc;
    std::atomic<int64_t> next_job(0);
    std::vector<std::vector<int64_t> > acc(n_threads, std::vector<int64_t>(3 * nc + 2, 0));
    std::atomic<int> err(0);
    auto body = [&](int t) {
        for (;;) {
            int64_t j = next_job.fetch_add(1);
            if (j >= n_jobs) break;
            int c = (int)(j / chunks_per_cand);
            int64_t b = (j % chunks_per_cand) * chunk;
            int64_t e = b + chunk < P ? b + chunk : P;
            int r = orc_playouts(vroot, colour, C > 0 ? cand_rc[2 * c] : -1,
                                 C > 0 ? cand_rc[2 * c + 1] : -1, position_id, (uint32_t)c, b, e, 1,
                                 seed_base, avrg_win, diffs ? diffs + (int64_t)c * P : NULL,
                                 &acc[t][3 * c], &acc[t][3 * nc]);
            if (r) err = r;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(body, t);
    for (auto& x : th) x.join();
    for (int t = 0; t < n_threads; t++) {
        for (int i = 0; i < 3 * nc; i++) out3[i] += acc[t][i];
        if (stat2) { stat2[0] += acc[t][3 * nc]; stat2[1] += acc[t][3 * nc + 1]; }
    }
    return err;
}

} /* extern "C" */
