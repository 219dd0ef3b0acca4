/* examples/capi_client.c — a plain-C client of include/agentgo_b200.h (no C++, no Python).
 *
 *   gcc -std=c99 -I include examples/capi_client.c -L agentgo_b200 -lagentgo_b200 \
 *       -Wl,-rpath,$PWD/agentgo_b200 -o capi_client && ./capi_client 9
 *
 * Does what MyGame::onMove does around the playout path (AgentGo.cpp:573-688): root playouts
 * for avrg_win, the candidate filter, playouts per candidate, and prints the integer counters.
 */
#include <stdio.h>
#include <stdlib.h>

#include "agentgo_b200.h"

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 9;
    agb_config cfg;
    agb_engine* h = NULL;
    agb_stats st;
    int64_t w, sp, sn;
    int16_t cand[2 * AGB_MAX_POINTS];
    static int64_t wins[AGB_MAX_POINTS], sum_pos[AGB_MAX_POINTS], sum_neg[AGB_MAX_POINTS];
    int C = 0, i, j, avrg_win, row, col, is_pass;

    agb_default_config(&cfg);
    cfg.board_size = n;
    if (agb_create(&cfg, &h) != AGB_OK) {
        fprintf(stderr, "agb_create: %s\n", agb_last_error(NULL));
        return 1;
    }
    agb_play(h, AGB_BLACK, n / 2, n / 2);
    agb_play(h, AGB_WHITE, n / 2 - 1, n / 2);

    /* testAvrgWin: 500 root playouts, black to move */
    if (agb_playouts(h, AGB_BLACK, NULL, 0, 500, 0, 20151001ull, 0, &w, &sp, &sn, NULL, &st) != AGB_OK) {
        fprintf(stderr, "agb_playouts: %s\n", agb_last_error(h));
        return 1;
    }
    avrg_win = (int)((sp + sn) / 500);
    printf("root wins=%lld sum_pos=%lld sum_neg=%lld avrg_win=%d\n", (long long)w, (long long)sp, (long long)sn, avrg_win);

    /* candidate filter of AgentGo.cpp:586-592 */
    for (i = 0; i < n; i++)
        for (j = 0; j < n; j++) {
            if (agb_check(h, 5, AGB_BLACK, i, j) || agb_check(h, 0, AGB_BLACK, i, j) ||
                agb_check(h, 7, AGB_BLACK, i, j) || agb_check(h, 6, AGB_BLACK, i, j))
                continue;
            {
                uint8_t board[AGB_MAX_POINTS];
                agb_get_board(h, board, NULL);
                if (board[i * n + j] != AGB_EMPTY) continue;
            }
            cand[2 * C] = (int16_t)i; cand[2 * C + 1] = (int16_t)j; C++;
        }
    if (agb_playouts(h, AGB_BLACK, cand, C, 150, avrg_win, 20151002ull, 0, wins, sum_pos, sum_neg, NULL, &st) != AGB_OK) {
        fprintf(stderr, "agb_playouts: %s\n", agb_last_error(h));
        return 1;
    }
    printf("candidates=%d playouts=%lld device_ms=%.3f\n", C, (long long)st.playouts, st.device_ms);
    for (i = 0; i < C; i++)
        printf("cand %d %d wins=%lld sum_pos=%lld sum_neg=%lld\n", cand[2 * i], cand[2 * i + 1],
               (long long)wins[i], (long long)sum_pos[i], (long long)sum_neg[i]);
    if (C > 0) {
        /* static pattern marks of MyWorker::work for the first candidate (AgentGo.cpp:104-282) and the
         * score of the position the way the reference's match harness counts it (GeneHatchery.cpp:393) */
        int64_t ua = 0, ub = 0;
        int dscore = 0;
        if (agb_static_units(h, AGB_BLACK, cand[0], cand[1], &ua, &ub) != AGB_OK || agb_calc_result(h, &dscore) != AGB_OK)
            return 1;
        printf("static %d %d units_a=%lld units_b=%lld calc_result=%d\n", cand[0], cand[1], (long long)ua, (long long)ub, dscore);
    }
    if (agb_genmove(h, AGB_BLACK, 150, 500, 20151001ull, &row, &col, &is_pass, NULL) != AGB_OK) return 1;
    printf("genmove row=%d col=%d pass=%d\n", row, col, is_pass);
    agb_destroy(h);
    return 0;
}
