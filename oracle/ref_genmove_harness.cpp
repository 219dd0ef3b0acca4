/*
 * oracle/ref_genmove_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * C-callable harness around the REFERENCE's own MyGame (AgentGo.cpp:351-705), compiled by
 * oracle/build_ref.sh as the tail of the liboracle_ref_genmove_13.so translation unit.  BOARD_SIZE
 * is 13 here and nowhere else: the move-selection code is full of 13x13 literals.
 *
 * orc_gm_* drive the game the way GTPAdapter::MainLoop does (GTPAdapter.cpp:78-97,151): `play` puts
 * the stone on `bd` and calls onPlay; `genmove` calls onMove and, when it returns a move, puts it
 * and calls onMoved.
 */
static uint64_t g_gm_seed_base = 0;

extern "C" void oracle_seed_root(int playout) {
    rng_seed(agb_playout_seed(g_gm_seed_base, 0, AGB_ROOT_CANDIDATE, (uint64_t)playout));
}
extern "C" void oracle_seed_candidate(int job, int playout) {
    rng_seed(agb_playout_seed(g_gm_seed_base + 1, 0, (uint32_t)job, (uint64_t)playout));
}

extern "C" {

void* orc_gm_new(void) {
    MyGame* g = new MyGame();
    g->bd.clear();          /* clear_board, GTPAdapter.cpp:151-152 */
    g->onClear();
    return g;
}
void orc_gm_free(void* g) { delete (MyGame*)g; }
void orc_gm_clear(void* vg) {
    MyGame* g = (MyGame*)vg;
    g->bd.clear();
    g->onClear();
}
int orc_gm_play(void* vg, int colour, int row, int col) {
    MyGame* g = (MyGame*)vg;
    if (!Piece(colour, row, col).legal()) return -1;
    if (g->bd.data[row][col] != GO_NULL) return -2;
    g->bd.put(colour, row, col);
    g->onPlay(colour - 1, row, col);
    return 0;
}
/* the eight tunables of the reference's command line (AgentGo.cpp:18-25, set from argv at :709-719) */
void orc_gm_set_tunables(const double* v) {
    index_a = v[0]; index_b = v[1]; index_c = v[2];
    index_amaf_a1 = v[3]; index_amaf_a2 = v[4]; index_amaf_b1 = v[5]; index_amaf_b2 = v[6];
    index_cnsv = v[7];
}
int orc_gm_step(void* vg) { return ((MyGame*)vg)->step; }
void orc_gm_set_step(void* vg, int step) { ((MyGame*)vg)->step = step; }

/* onMove for `colour`; returns 1 and the move, or 0 for pass.  The move is NOT played.
 * total_mark_out / mc_out / amaf1_out / amaf2_out (169 doubles each, may be NULL) receive the
 * reference's global tables as onMove left them. */
int orc_gm_genmove(void* vg, int colour, uint64_t seed_base, int* row, int* col,
                   double* total_mark_out, double* mc_out, double* amaf1_out, double* amaf2_out) {
    MyGame* g = (MyGame*)vg;
    g_gm_seed_base = seed_base;
    int a = -1, b = -1;
    const bool moved = g->onMove(colour - 1, a, b);
    *row = a; *col = b;
    if (total_mark_out) memcpy(total_mark_out, total_mark, sizeof(double) * 169);
    if (mc_out) memcpy(mc_out, mc, sizeof(double) * 169);
    if (amaf1_out) memcpy(amaf1_out, amaf1, sizeof(double) * 169);
    if (amaf2_out) memcpy(amaf2_out, amaf2, sizeof(double) * 169);
    return moved ? 1 : 0;
}

/* The static part of MyWorker::work (AgentGo.cpp:75-282) for one candidate: runs work() with the
 * playout count... no: work() has numset = 150 as a literal, so the static marks are isolated by
 * difference instead — total_mark[row][col] after work() minus mc[row][col] (both receive exactly
 * the same playout terms, :339-340).  Exact: the static part is an integer < 2^53 and the playout
 * terms are added to total_mark after it, so the difference is formed from the returned tables by
 * the caller when index_c == 1 and all terms are integers (cnsv == 1, i.e. avrg_win == 0). */
double orc_gm_static_mark(void* vg, int colour, int row, int col, uint64_t seed_base) {
    MyGame* g = (MyGame*)vg;
    g_gm_seed_base = seed_base;
    memset(total_mark, 0, sizeof(total_mark));
    memset(mc, 0, sizeof(mc));
    memset(amaf1, 0, sizeof(amaf1));
    memset(amaf2, 0, sizeof(amaf2));
    MyJob* j = new MyJob(g->bd);
    j->i = row; j->j = col; j->isWh = colour - 1; j->avrg_win = 0;
    oracle_job_index = 0;
    MyWorker w;
    w.work(j);
    delete j;
    return total_mark[row][col] - mc[row][col];
}

/* CalcResults (GeneHatchery.cpp:393-436) on a COPY of the game's board (the reference scores the
 * referee's scratch board in place): black minus white after the iterated neighbour fill. */
int orc_gm_calc_result(void* vg) {
    Board copy(((MyGame*)vg)->bd);
    return CalcResults(copy);
}

}  /* extern "C" */
