/*
 * oracle/agb_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the flat Monte-Carlo playout path of Menooker/AgentGo,
 * board size chosen at run time (the reference is compile-time 13x13,
 * GoContest/AgentGo.h:15).  Every function cites the reference lines it follows
 * (paths relative to the reference checkout, AgentGo/...).
 *
 * PINNING: the reference ships no tests and no golden vectors, so
 * parity is pinned against outputs of the reference ITSELF: oracle/build_ref.sh
 * compiles the reference's own Board.cpp with g++ (oracle/_ref/), and
 * tests/test_oracle.py checks this restatement against it — state after every move
 * of random games, all predicates at all points, and per-playout diffs — plus the
 * committed fixtures under tests/golden/ that the reference build generated
 * (tests/golden/make_golden.py).
 *
 * Representation differences (behaviour-preserving):
 *   - a point is idx = row*N + col instead of a Piece{agent,row,col} (Piece.h:7-21);
 *   - SetNode's ordered `pieces` array (SetNode.h:19-20, merge_pieces SetNode.cpp:78-82)
 *     is a singly linked list in piece order: first piece = the root stone (invariant of
 *     unionSetNode, Board.cpp:279-291), `snext[]` = following piece, `tail[root]` = last;
 *   - dead state is dropped: distance[] (its only writer is commented out, Board.cpp:169),
 *     true_eyes, deepth, GO_HISTORY, GO_BOARD_TIME.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "agentgo_b200.h"
#include "tinymt32.h"

#define MAXN 19
#define MAXNN (MAXN * MAXN)
#define UNKNOWN_IDX 999          /* SetNode.h:6 */
#define SPACE_HEAD_NULL (-1)     /* Board.h:12 */
#define ENDING_THRESHOLD 100     /* Board.h:13 — kept at every board size */
#define ENDING_STONES 120        /* Board.cpp:349 — kept at every board size */

typedef struct {
    int n, nn;
    int data[MAXNN];             /* Board.h:42 */
    int pnode[MAXNN];            /* SetNode::pnode */
    int hp[MAXNN];               /* SetNode::hp (pseudo-liberties), valid at roots */
    int size[MAXNN];             /* SetNode::size, valid at roots */
    int snext[MAXNN];            /* next piece of the string in SetNode::pieces order */
    int tail[MAXNN];             /* last piece, valid at roots */
    int space_list[MAXNN][2];    /* Board.h:46 */
    int space_head;
    unsigned char reserve[MAXNN];
    int reserve_total;
    int num_black, num_white;
    int exist_compete, compete[3];
    int last_move[2];            /* idx or -1 (Piece empty) */
    int game_ending;
} OBoard;

typedef struct {
    oracle_tinymt32_t mt;
    int64_t draws;
} ORng;

static int o_rand(ORng* r) {      /* rand() := tinymt32 >> 17 */
    r->draws++;
    return (int)(oracle_tinymt32_generate_uint32(&r->mt) >> 17);
}
static void o_seed(ORng* r, uint32_t seed) {
    r->mt.mat1 = AGB_TINYMT_MAT1; r->mt.mat2 = AGB_TINYMT_MAT2; r->mt.tmat = AGB_TINYMT_TMAT;
    oracle_tinymt32_init(&r->mt, seed);
}

/* SetNode::clear, SetNode.cpp:46-57 */
static void set_clear(OBoard* b, int idx) {
    b->pnode[idx] = UNKNOWN_IDX; b->hp[idx] = 0; b->size[idx] = 0;
    b->snext[idx] = -1; b->tail[idx] = -1;
}

/* Board::clear, Board.cpp:27-66 */
static void o_clear(OBoard* b) {
    int i, nn = b->nn;
    for (i = 0; i < nn; i++) { b->data[i] = AGB_EMPTY; set_clear(b, i); }
    b->num_black = b->num_white = 0;
    for (i = 0; i < nn; i++) {
        b->space_list[i][0] = i - 1;
        b->space_list[i][1] = i + 1;
        b->reserve[i] = 0;
    }
    b->space_list[0][0] = 0;
    b->space_list[nn - 1][1] = nn - 1;
    b->space_head = 0;
    b->reserve_total = 0;
    b->exist_compete = 0;
    b->compete[0] = AGB_EMPTY; b->compete[1] = b->compete[2] = -1;
    b->game_ending = 0;
    b->last_move[0] = b->last_move[1] = -1;
}

/* Board::getSet, Board.cpp:263-266: root index of the string at idx */
static int get_set(const OBoard* b, int idx) { return b->pnode[idx]; }

/* Board::unionSetNode, Board.cpp:279-291 (+ SetNode::merge_pieces :78-82, drop :59-69) */
static void union_set(OBoard* b, int s1, int s2) {
    int k;
    if (s1 == s2) return;
    b->hp[s1] += b->hp[s2];
    b->snext[b->tail[s1]] = s2;               /* pieces of s1, then pieces of s2 */
    b->tail[s1] = b->tail[s2];
    b->size[s1] += b->size[s2];
    for (k = s2; k != -1; k = b->snext[k]) b->pnode[k] = s1;
    b->hp[s2] = 0; b->size[s2] = 0; b->tail[s2] = -1;
}

/* Board::killSetNode, Board.cpp:293-339 */
static void kill_set(OBoard* b, int sn) {
    int n = b->n, k, nx;
    int kill_single = (b->size[sn] == 1);
    for (k = sn; k != -1; k = nx) {
        int i = k / n, j = k % n, agent = b->data[k];
        nx = b->snext[k];
        b->data[k] = AGB_EMPTY;
        if (agent == AGB_BLACK) b->num_black--; else b->num_white--;
        if (k != sn) set_clear(b, k);
        if (b->space_head != SPACE_HEAD_NULL) {          /* push at the head, :311-316 */
            b->space_list[k][0] = k;
            b->space_list[k][1] = b->space_head;
            b->space_list[b->space_head][0] = k;
            b->space_head = k;
        } else {                                         /* :317-321 */
            b->space_list[k][0] = k;
            b->space_list[k][1] = k;
            b->space_head = k;
        }
        /* one pseudo-liberty back to every adjacent stone of the other colour, :323-326 */
        if (i - 1 >= 0 && b->data[k - n] != AGB_EMPTY && b->data[k - n] != agent) b->hp[get_set(b, k - n)]++;
        if (j - 1 >= 0 && b->data[k - 1] != AGB_EMPTY && b->data[k - 1] != agent) b->hp[get_set(b, k - 1)]++;
        if (i + 1 < n && b->data[k + n] != AGB_EMPTY && b->data[k + n] != agent) b->hp[get_set(b, k + n)]++;
        if (j + 1 < n && b->data[k + 1] != AGB_EMPTY && b->data[k + 1] != agent) b->hp[get_set(b, k + 1)]++;
        if (kill_single) {                               /* simple-ko flag, :328-333 */
            b->exist_compete = 1;
            b->compete[0] = agent; b->compete[1] = i; b->compete[2] = j;
        }
    }
    set_clear(b, sn);
}

/* one neighbour step of Board::put, Board.cpp:129-137 (x4) */
static void put_touch(OBoard* b, int agent, int idx, int nb) {
    int sn;
    if (b->data[nb] == AGB_EMPTY) return;
    sn = get_set(b, nb);
    b->hp[sn] -= 1;
    if (b->data[nb] != agent) {
        if (b->hp[sn] == 0) kill_set(b, sn);
    } else {
        union_set(b, sn, get_set(b, idx));
    }
}

/* Board::put, Board.cpp:68-177 */
static void o_put(OBoard* b, int agent, int row, int col) {
    int n = b->n, i = row, j = col, idx = row * n + col, hp = 0, sp, sx, self;
    b->data[idx] = agent;
    if (agent == AGB_BLACK) b->num_black++; else b->num_white++;
    b->exist_compete = 0;
    /* SetNode::init, SetNode.cpp:31-44 (`hp = hp;` is a self-assignment: hp stays 0) */
    b->pnode[idx] = idx; b->size[idx] = 1; b->snext[idx] = -1; b->tail[idx] = idx;
    /* unlink from the empty list, :105-118 */
    sp = b->space_list[idx][0]; sx = b->space_list[idx][1];
    if (b->num_white + b->num_black == b->nn) b->space_head = SPACE_HEAD_NULL;
    else if (sp == idx) { b->space_list[sx][0] = sx; b->space_head = sx; }
    else if (sx == idx) { b->space_list[sp][1] = sp; }
    else { b->space_list[sp][1] = sx; b->space_list[sx][0] = sp; }
    /* own pseudo-liberties counted before any capture, :123-127 */
    if (i - 1 >= 0 && b->data[idx - n] == AGB_EMPTY) hp++;
    if (j - 1 >= 0 && b->data[idx - 1] == AGB_EMPTY) hp++;
    if (i + 1 < n && b->data[idx + n] == AGB_EMPTY) hp++;
    if (j + 1 < n && b->data[idx + 1] == AGB_EMPTY) hp++;
    /* neighbours in the order up, left, down, right, :129-164 */
    if (i - 1 >= 0) put_touch(b, agent, idx, idx - n);
    if (j - 1 >= 0) put_touch(b, agent, idx, idx - 1);
    if (i + 1 < n) put_touch(b, agent, idx, idx + n);
    if (j + 1 < n) put_touch(b, agent, idx, idx + 1);
    self = get_set(b, idx);
    b->hp[self] += hp;
    if (b->hp[self] == 0) kill_set(b, self);             /* :167 */
    b->last_move[agent - 1] = idx;                       /* :171 */
}

/* Board::pass, Board.cpp:200-203 */
static void o_pass(OBoard* b, int agent) {
    b->exist_compete = 0;
    b->last_move[agent - 1] = -1;
}

/* The five 4-neighbour predicates share one scan: neighbours in the order up, left,
 * down, right; strings seen twice are merged and their multiplicity becomes minus_hp
 * (Board.cpp:482-515 and the four copies of it).                                    */
typedef struct { int n_empty; int s[4]; int enemy[4]; int minus[4]; } NbScan;

static void nb_scan(const OBoard* b, int agent, int row, int col, NbScan* q) {
    int n = b->n, idx = row * n + col, m, k;
    int nb[4], ok[4];
    nb[0] = idx - n; ok[0] = row - 1 >= 0;
    nb[1] = idx - 1; ok[1] = col - 1 >= 0;
    nb[2] = idx + n; ok[2] = row + 1 < n;
    nb[3] = idx + 1; ok[3] = col + 1 < n;
    q->n_empty = 0;
    for (m = 0; m < 4; m++) {
        q->s[m] = -1; q->enemy[m] = 0; q->minus[m] = 1;
        if (!ok[m]) continue;
        if (b->data[nb[m]] == AGB_EMPTY) { q->n_empty++; continue; }
        if (b->data[nb[m]] != agent) q->enemy[m] = 1;
        q->s[m] = get_set(b, nb[m]);
    }
    for (m = 0; m < 3; m++) {
        if (q->s[m] < 0) continue;
        for (k = m + 1; k < 4; k++)
            if (q->s[m] == q->s[k]) { q->s[k] = -1; q->minus[m] += q->minus[k]; }
    }
}

/* Board::checkSuicide, Board.cpp:480-524 */
static int check_suicide(const OBoard* b, int agent, int row, int col) {
    NbScan q; int m;
    nb_scan(b, agent, row, col, &q);
    if (q.n_empty > 0) return 0;
    for (m = 0; m < 4; m++) {
        if (q.s[m] < 0) continue;
        if (q.enemy[m] && b->hp[q.s[m]] - q.minus[m] == 0) return 0;
        if (!q.enemy[m] && b->hp[q.s[m]] - q.minus[m] > 0) return 0;
    }
    return 1;
}

/* Board::checkKill, Board.cpp:526-563 */
static int check_kill(const OBoard* b, int agent, int row, int col) {
    NbScan q; int m;
    nb_scan(b, agent, row, col, &q);
    for (m = 0; m < 4; m++)
        if (q.s[m] >= 0 && q.enemy[m] && b->hp[q.s[m]] - q.minus[m] == 0) return 1;
    return 0;
}

/* Board::checkSurvive, Board.cpp:565-626 */
static int check_survive(const OBoard* b, int agent, int row, int col) {
    NbScan q; int m, hp_new, hp_min = 9999;
    if (!check_kill(b, 3 - agent, row, col)) return 0;
    nb_scan(b, agent, row, col, &q);
    hp_new = q.n_empty;
    for (m = 0; m < 4; m++) {
        if (q.s[m] < 0) continue;
        if (q.enemy[m] && b->hp[q.s[m]] - q.minus[m] == 0) return 1;
        if (!q.enemy[m]) {
            hp_new += b->hp[q.s[m]] - q.minus[m];
            if (b->hp[q.s[m]] < hp_min) hp_min = b->hp[q.s[m]];
        }
    }
    return hp_new > hp_min;
}

/* Board::checkDying, Board.cpp:628-679 */
static int check_dying(const OBoard* b, int agent, int row, int col) {
    NbScan q; int m, hp_new;
    nb_scan(b, agent, row, col, &q);
    hp_new = q.n_empty;
    for (m = 0; m < 4; m++) {
        if (q.s[m] < 0) continue;
        if (q.enemy[m] && b->hp[q.s[m]] - q.minus[m] == 0) return 0;
        if (!q.enemy[m]) hp_new += b->hp[q.s[m]] - q.minus[m];
    }
    return hp_new == 1;
}

/* Board::checkChase, Board.cpp:681-733 */
static int check_chase(const OBoard* b, int agent, int row, int col) {
    NbScan q; int m, hp_new, chase = 0;
    nb_scan(b, agent, row, col, &q);
    hp_new = q.n_empty;
    for (m = 0; m < 4; m++) {
        if (q.s[m] < 0) continue;
        if (q.enemy[m] && b->hp[q.s[m]] - q.minus[m] == 1) chase = 1;
        if (!q.enemy[m]) hp_new += b->hp[q.s[m]] - q.minus[m];
    }
    return chase && hp_new > 1;
}

/* Board::checkTrueEye, Board.cpp:735-751 */
static int check_true_eye(const OBoard* b, int agent, int row, int col) {
    int n = b->n, i = row, j = col, idx = row * n + col, cnt = 0;
    if (i - 1 >= 0 && b->data[idx - n] != agent) return 0;
    if (j - 1 >= 0 && b->data[idx - 1] != agent) return 0;
    if (i + 1 < n && b->data[idx + n] != agent) return 0;
    if (j + 1 < n && b->data[idx + 1] != agent) return 0;
    if (i - 1 >= 0 && j - 1 >= 0 && b->data[idx - n - 1] != AGB_EMPTY && b->data[idx - n - 1] != agent) cnt++;
    if (i - 1 >= 0 && j + 1 < n && b->data[idx - n + 1] != AGB_EMPTY && b->data[idx - n + 1] != agent) cnt++;
    if (i + 1 < n && j - 1 >= 0 && b->data[idx + n - 1] != AGB_EMPTY && b->data[idx + n - 1] != agent) cnt++;
    if (i + 1 < n && j + 1 < n && b->data[idx + n + 1] != AGB_EMPTY && b->data[idx + n + 1] != agent) cnt++;
    if (i == 0 || i == n - 1 || j == 0 || j == n - 1) { if (cnt >= 1) return 0; }
    else if (cnt >= 2) return 0;
    return 1;
}

/* Board::checkCompete, Board.cpp:753-755 */
static int check_compete(const OBoard* b, int agent, int row, int col) {
    return b->exist_compete && row == b->compete[1] && col == b->compete[2] && agent == b->compete[0];
}

/* Board::checkPiece, Board.cpp:193-198 */
static int check_piece(const OBoard* b, int agent, int row, int col) {
    if (row < 0 || col < 0 || row > b->n - 1 || col > b->n - 1) return 0;
    return b->data[row * b->n + col] == agent;
}

/* Board::checkNoSense, Board.cpp:883-908.  The reference falls off the end on the false
 * path (undefined behaviour); `false` is taken there, as in oracle/build_ref.sh edit 5. */
static int check_no_sense(const OBoard* b, int agent, int row, int col) {
    int enemy = 3 - agent, n = b->n, i, j, range;
    int self1 = 0, enemy1 = 0, self2 = 0, enemy2 = 0, at_border, near_border;
    range = 1;
    for (i = row - range; i <= row + range; i++)
        for (j = col - range; j <= col + range; j++) {
            if (i == row && j == col) continue;
            if (check_piece(b, agent, i, j)) self1++;
            else if (check_piece(b, enemy, i, j)) enemy1++;
        }
    at_border = (row == 0 || row == n - 1 || col == 0 || col == n - 1);
    range = 2;
    for (i = row - range; i <= row + range; i++)
        for (j = col - range; j <= col + range; j++) {
            if (i == row && j == col) continue;
            if (check_piece(b, agent, i, j)) self2++;
            else if (check_piece(b, enemy, i, j)) enemy2++;
        }
    near_border = (row <= 1 || row >= n - 2 || col <= 1 || col >= n - 2);
    return ((enemy1 + self1 == 0) && at_border) || ((enemy2 + self2 == 0) && near_border);
}

/* Board::getRandomPieceComplex, Board.cpp:413-478; returns idx or -1 (pass) */
static int random_piece_complex(OBoard* b, ORng* rng, int agent) {
    int n = b->n, idx, range, rnd, k;
    int spaces = b->nn - b->num_black - b->num_white;
    if (spaces == 0) return -1;
    idx = b->space_head;
    for (;;) {
        int row = idx / n, col = idx % n;
        if (check_true_eye(b, agent, row, col) || check_suicide(b, agent, row, col) ||
            check_compete(b, agent, row, col)) {
            b->reserve[idx] = 1;                         /* addReserve, :757-761 */
            b->reserve_total++;
        }
        if (b->space_list[idx][1] == idx) break;
        idx = b->space_list[idx][1];
    }
    range = b->nn - b->num_black - b->num_white - b->reserve_total;
    if (range == 0) {
        b->reserve_total = 0;                            /* resetReserve, :764-768 */
        for (k = 0; k < b->nn; k++) b->reserve[k] = 0;
        return -1;
    }
    rnd = o_rand(rng) % range;
    idx = b->space_head;
    for (;;) {
        if (b->reserve[idx]) { idx = b->space_list[idx][1]; continue; }
        if (rnd == 0) break;
        idx = b->space_list[idx][1];
        rnd--;
    }
    b->reserve_total = 0;
    for (k = 0; k < b->nn; k++) b->reserve[k] = 0;
    return idx;
}

/* Board::getRandomPiece, Board.cpp:341-411; returns idx or -1 (pass) */
static int random_piece(OBoard* b, ORng* rng, int agent) {
    int n = b->n, row = 0, col = 0, flag = 1, count = 0, k;
    int enemy = 3 - agent, lm;
    if (b->game_ending) {                                /* :348-351 */
        if (b->num_black + b->num_white <= ENDING_STONES) b->game_ending = 0;
        else return random_piece_complex(b, rng, agent);
    }
    lm = b->last_move[enemy - 1];
    if (lm >= 0 && b->data[lm] == enemy) {               /* :355-380 */
        int er = lm / n, ec = lm % n;
        for (k = 0; k < 9; k++) {
            int i = er - 1 + (o_rand(rng) % 3);
            int j = ec - 1 + (o_rand(rng) % 3);
            if ((i >= 0 && i < n && j >= 0 && j < n) && b->data[i * n + j] == AGB_EMPTY &&
                (check_kill(b, agent, i, j) || (k < 6 && check_survive(b, agent, i, j)) ||
                 (k < 4 && check_chase(b, agent, i, j))) &&
                !check_dying(b, agent, i, j) && !check_true_eye(b, agent, i, j) &&
                !check_suicide(b, agent, i, j) && !check_compete(b, agent, i, j)) {
                row = i; col = j; flag = 0;
                break;
            }
        }
    }
    while (flag) {                                       /* :382-406 */
        flag = 0;
        count++;
        row = o_rand(rng) % n;
        col = o_rand(rng) % n;
        if (b->data[row * n + col] != AGB_EMPTY || check_dying(b, agent, row, col) ||
            check_true_eye(b, agent, row, col) || check_suicide(b, agent, row, col) ||
            check_compete(b, agent, row, col))
            flag = 1;
        if (count == ENDING_THRESHOLD) {                 /* even if the 100th draw passed */
            b->game_ending = 1;
            return random_piece_complex(b, rng, agent);
        }
    }
    return row * n + col;
}

/* Board::countScore, Board.cpp:910-925 */
static int count_score(const OBoard* b, int agent) {
    int n = b->n, i, j, count = 0;
    for (i = 0; i < n; i++)
        for (j = 0; j < n; j++) {
            int c = b->data[i * n + j];
            if (c == agent) count++;
            else if (c == AGB_EMPTY) {
                if ((j > 0 && b->data[i * n + j - 1] == agent) ||
                    (j == 0 && b->data[i * n + j + 1] == agent)) count++;
            }
        }
    return count;
}

/* pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443) */
static int one_playout(OBoard* bd, const OBoard* start, ORng* rng, int agent, int first,
                       int second, int64_t* moves) {
    int first_go = 1, second_go = 1, r;
    *bd = *start;                                        /* Board::clone, Board.cpp:216-254 */
    while (first_go || second_go) {
        r = random_piece(bd, rng, first);
        if (r >= 0) { o_put(bd, first, r / bd->n, r % bd->n); first_go = 1; (*moves)++; }
        else { first_go = 0; o_pass(bd, first); }
        r = random_piece(bd, rng, second);
        if (r >= 0) { o_put(bd, second, r / bd->n, r % bd->n); second_go = 1; (*moves)++; }
        else { second_go = 0; o_pass(bd, second); }
    }
    return count_score(bd, agent) - count_score(bd, 3 - agent);   /* AgentGo.cpp:327 / :444 */
}

/* ------------------------------------------------------------------ orc_* API */
const char* orc_kind(void) { return "port"; }
int orc_board_size(void) { return 0; /* any size up to 19 */ }
int orc_sizeof_board(void) { return (int)sizeof(OBoard); }

void* orc_new(int n) {
    OBoard* b;
    if (n < 2 || n > MAXN) return NULL;
    b = (OBoard*)malloc(sizeof(OBoard));
    if (!b) return NULL;
    memset(b, 0, sizeof(*b));
    b->n = n; b->nn = n * n;
    o_clear(b);
    return b;
}
void orc_free(void* b) { free(b); }
void orc_clear(void* b) { o_clear((OBoard*)b); }

int orc_put(void* vb, int agent, int row, int col) {
    OBoard* b = (OBoard*)vb;
    if ((agent != AGB_BLACK && agent != AGB_WHITE) || row < 0 || col < 0 || row >= b->n || col >= b->n)
        return -1;                                       /* Piece::legal, Piece.cpp:28-30 */
    if (b->data[row * b->n + col] != AGB_EMPTY) return -2;
    o_put(b, agent, row, col);
    return 0;
}
void orc_pass(void* b, int agent) { o_pass((OBoard*)b, agent); }

int orc_check(void* vb, int which, int agent, int row, int col) {
    const OBoard* b = (const OBoard*)vb;
    switch (which) {
        case 0: return check_suicide(b, agent, row, col);
        case 1: return check_kill(b, agent, row, col);
        case 2: return check_survive(b, agent, row, col);
        case 3: return check_dying(b, agent, row, col);
        case 4: return check_chase(b, agent, row, col);
        case 5: return check_true_eye(b, agent, row, col);
        case 6: return check_compete(b, agent, row, col);
        case 7: return check_no_sense(b, agent, row, col);
    }
    return -1;
}
int orc_count_score(void* b, int agent) { return count_score((const OBoard*)b, agent); }

int orc_export_state(void* vb, int32_t* out, int cap) {
    const OBoard* b = (const OBoard*)vb;
    int nn = b->nn, n = b->n, words = 16 + 7 * nn, idx, a;
    int32_t *data, *root, *hp, *size, *next, *sprev, *snext;
    if (cap < words) return -1;
    memset(out, 0, sizeof(int32_t) * (size_t)words);
    out[0] = n; out[1] = b->num_black; out[2] = b->num_white; out[3] = b->space_head;
    out[4] = b->exist_compete ? 1 : 0;
    out[5] = b->exist_compete ? b->compete[0] : 0;
    out[6] = b->exist_compete ? b->compete[1] : -1;
    out[7] = b->exist_compete ? b->compete[2] : -1;
    for (a = 0; a < 2; a++) {
        int lm = b->last_move[a];
        out[8 + 3 * a] = lm >= 0 ? a + 1 : 0;
        out[9 + 3 * a] = lm >= 0 ? lm / n : -1;
        out[10 + 3 * a] = lm >= 0 ? lm % n : -1;
    }
    out[14] = b->game_ending ? 1 : 0;
    data = out + 16; root = data + nn; hp = root + nn; size = hp + nn; next = size + nn;
    sprev = next + nn; snext = sprev + nn;
    for (idx = 0; idx < nn; idx++) {
        data[idx] = b->data[idx];
        if (b->data[idx] == AGB_EMPTY) {
            root[idx] = -1; next[idx] = -1;
            sprev[idx] = b->space_list[idx][0]; snext[idx] = b->space_list[idx][1];
        } else {
            root[idx] = b->pnode[idx]; next[idx] = b->snext[idx];
            sprev[idx] = snext[idx] = -1;
            if (b->pnode[idx] == idx) { hp[idx] = b->hp[idx]; size[idx] = b->size[idx]; }
        }
    }
    return words;
}

int orc_random_game(void* vb, int first_colour, int n_moves, uint32_t seed, uint32_t* moves_out) {
    OBoard* b = (OBoard*)vb;
    ORng rng; int i, colour = first_colour;
    rng.draws = 0;
    o_seed(&rng, seed);
    for (i = 0; i < n_moves; i++) {
        int r = random_piece(b, &rng, colour);
        if (r >= 0) { o_put(b, colour, r / b->n, r % b->n); moves_out[i] = AGB_MOVE(colour, r / b->n, r % b->n); }
        else { o_pass(b, colour); moves_out[i] = AGB_MOVE_PASS(colour); }
        colour = 3 - colour;
    }
    return 0;
}

int orc_playouts(void* vroot, int colour, int cand_row, int cand_col, uint32_t position_id,
                 uint32_t cand_idx, int64_t p_begin, int64_t p_end, int64_t p_stride,
                 uint64_t seed_base, int avrg_win, int16_t* diffs, int64_t* out3, int64_t* stat2) {
    const OBoard* root = (const OBoard*)vroot;
    int agent = colour, enemy = 3 - colour, first, second;
    int64_t wins = 0, sp = 0, sn = 0, moves = 0, p;
    OBoard *start, *bd;
    ORng rng;
    start = (OBoard*)malloc(sizeof(OBoard));
    bd = (OBoard*)malloc(sizeof(OBoard));
    if (!start || !bd) { free(start); free(bd); return -4; }
    *start = *root;
    if (cand_row >= 0) {
        if (start->data[cand_row * start->n + cand_col] != AGB_EMPTY) { free(start); free(bd); return -2; }
        o_put(start, agent, cand_row, cand_col);         /* AgentGo.cpp:83 */
        first = enemy; second = agent;                   /* AgentGo.cpp:300,310 */
    } else {
        cand_idx = AGB_ROOT_CANDIDATE;
        first = agent; second = enemy;                   /* AgentGo.cpp:425,434 */
    }
    rng.draws = 0;
    for (p = p_begin; p < p_end; p += p_stride) {
        int diff, w;
        o_seed(&rng, agb_playout_seed(seed_base, position_id, cand_idx, (uint64_t)p));
        diff = one_playout(bd, start, &rng, agent, first, second, &moves);
        if (diffs) diffs[p] = (int16_t)diff;
        w = diff - avrg_win;                             /* AgentGo.cpp:327 */
        if (w >= 0) { wins++; sp += w; } else sn += w;   /* sign test of AgentGo.cpp:328 */
    }
    free(start); free(bd);
    if (out3) { out3[0] += wins; out3[1] += sp; out3[2] += sn; }
    if (stat2) { stat2[0] += moves; stat2[1] += rng.draws; }
    return 0;
}

#define AMAF_RANGE 40            /* GoContest/AgentGo.h:16 */

/* AMAF first-play marks, AgentGo.cpp:286-338 (candidate mode); contract: see ref_harness.cpp.
 * amaf[which][sign][step 0..40][q], which 0 = agent / 1 = enemy first play, sign 0: w >= 0. */
int orc_playouts_amaf(void* vroot, int colour, int cand_row, int cand_col, uint32_t position_id,
                      uint32_t cand_idx, int64_t p_begin, int64_t p_end, int64_t p_stride,
                      uint64_t seed_base, int avrg_win, int64_t* out3, int64_t* amaf) {
    const OBoard* root = (const OBoard*)vroot;
    int agent = colour, enemy = 3 - colour, nn = root->nn, n = root->n, S = AMAF_RANGE + 1, q;
    int mark[MAXNN], mark2[MAXNN];
    int64_t p;
    OBoard *start, *bd;
    ORng rng;
    if (cand_row < 0) return -1;
    start = (OBoard*)malloc(sizeof(OBoard));
    bd = (OBoard*)malloc(sizeof(OBoard));
    if (!start || !bd) { free(start); free(bd); return -4; }
    *start = *root;
    if (start->data[cand_row * n + cand_col] != AGB_EMPTY) { free(start); free(bd); return -2; }
    o_put(start, agent, cand_row, cand_col);                      /* AgentGo.cpp:83 */
    rng.draws = 0;
    for (p = p_begin; p < p_end; p += p_stride) {
        int agent_go = 1, enemy_go = 1, cstep = 1, r, w, sign;
        o_seed(&rng, agb_playout_seed(seed_base, position_id, cand_idx, (uint64_t)p));
        *bd = *start;                                             /* :288 */
        memset(mark, 0, sizeof(mark));                            /* :289-290 */
        memset(mark2, 0, sizeof(mark2));
        while (agent_go || enemy_go) {                            /* :295-325 */
            cstep++;
            r = random_piece(bd, &rng, enemy);
            if (r >= 0) { o_put(bd, enemy, r / n, r % n); enemy_go = 1; if (mark2[r] == 0) mark2[r] = cstep; }
            else { enemy_go = 0; o_pass(bd, enemy); }
            r = random_piece(bd, &rng, agent);
            if (r >= 0) { o_put(bd, agent, r / n, r % n); agent_go = 1; if (mark[r] == 0) mark[r] = cstep; }
            else { agent_go = 0; o_pass(bd, agent); }
        }
        w = count_score(bd, agent) - count_score(bd, enemy) - avrg_win;     /* :327 */
        sign = w >= 0 ? 0 : 1;
        if (out3) { if (w >= 0) { out3[0]++; out3[1] += w; } else out3[2] += w; }
        for (q = 0; q < nn; q++) {                                /* :329-338 */
            if (mark[q] != 0 && mark[q] <= AMAF_RANGE) amaf[((0 * 2 + sign) * S + mark[q]) * nn + q] += w;
            else if (mark2[q] != 0 && mark2[q] <= AMAF_RANGE) amaf[((1 * 2 + sign) * S + mark2[q]) * nn + q] += w;
        }
    }
    free(start); free(bd);
    return 0;
}

/* CPU baseline thread pool: the reference's policy (2*cores+2 threads, TinyMT.h:543-546;
 * one job per candidate, AgentGo.cpp:595-600) with pthreads; jobs are (candidate, chunk). */
typedef struct {
    void* root; int colour; const int16_t* cand_rc; int C; int64_t P; uint32_t position_id;
    uint64_t seed_base; int avrg_win; int16_t* diffs;
    int64_t chunk, chunks_per_cand, n_jobs;
    int64_t* next_job; pthread_mutex_t* mu;
    int64_t* acc; int err;
} MtJob;

static void* mt_body(void* arg) {
    MtJob* j = (MtJob*)arg;
    int nc = j->C > 0 ? j->C : 1;
    for (;;) {
        int64_t id, b, e; int c, r;
        pthread_mutex_lock(j->mu);
        id = (*j->next_job)++;
        pthread_mutex_unlock(j->mu);
        if (id >= j->n_jobs) break;
        c = (int)(id / j->chunks_per_cand);
        b = (id % j->chunks_per_cand) * j->chunk;
        e = b + j->chunk < j->P ? b + j->chunk : j->P;
        r = orc_playouts(j->root, j->colour, j->C > 0 ? j->cand_rc[2 * c] : -1,
                         j->C > 0 ? j->cand_rc[2 * c + 1] : -1, j->position_id, (uint32_t)c, b, e, 1,
                         j->seed_base, j->avrg_win, j->diffs ? j->diffs + (int64_t)c * j->P : NULL,
                         &j->acc[3 * c], &j->acc[3 * nc]);
        if (r) j->err = r;
    }
    return NULL;
}

int orc_playouts_mt(void* vroot, int colour, const int16_t* cand_rc, int C, int64_t P,
                    uint32_t position_id, uint64_t seed_base, int avrg_win, int n_threads,
                    int16_t* diffs, int64_t* out3, int64_t* stat2) {
    int nc = C > 0 ? C : 1, t, i, err = 0;
    int64_t next_job = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_t* th; MtJob* jobs;
    if (n_threads <= 0) n_threads = 2 * (int)sysconf(_SC_NPROCESSORS_ONLN) + 2;
    th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    jobs = (MtJob*)calloc((size_t)n_threads, sizeof(MtJob));
    for (t = 0; t < n_threads; t++) {
        MtJob* j = &jobs[t];
        j->root = vroot; j->colour = colour; j->cand_rc = cand_rc; j->C = C; j->P = P;
        j->position_id = position_id; j->seed_base = seed_base; j->avrg_win = avrg_win;
        j->diffs = diffs; j->chunk = 256; j->chunks_per_cand = (P + 255) / 256;
        j->n_jobs = j->chunks_per_cand * nc; j->next_job = &next_job; j->mu = &mu;
        j->acc = (int64_t*)calloc((size_t)(3 * nc + 2), sizeof(int64_t));
        pthread_create(&th[t], NULL, mt_body, j);
    }
    for (t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        for (i = 0; i < 3 * nc; i++) out3[i] += jobs[t].acc[i];
        if (stat2) { stat2[0] += jobs[t].acc[3 * nc]; stat2[1] += jobs[t].acc[3 * nc + 1]; }
        if (jobs[t].err) err = jobs[t].err;
        free(jobs[t].acc);
    }
    free(th); free(jobs);
    return err;
}
