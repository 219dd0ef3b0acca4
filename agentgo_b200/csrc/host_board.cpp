// host_board.cpp — see host_board.h
#include "host_board.h"

#include <string.h>

namespace agb {

HostBoard::HostBoard(int n) { reset(n); }

HostBoard::Ref HostBoard::open() const {
    Ref r;
    r.w0 = const_cast<uint32_t*>(pos_.w0);
    r.w1 = const_cast<uint32_t*>(pos_.w1);
    r.load_scalars(pos_);
    return r;
}

void HostBoard::close(const Ref& r) { r.store_scalars(pos_); }

void HostBoard::reset(int n) {
    memset(&pos_, 0, sizeof(pos_));
    pos_.n = (int16_t)n;
    clear();
}

void HostBoard::clear() {
    const int n = pos_.n;
    memset(pos_.w0, 0, sizeof(pos_.w0));
    memset(pos_.w1, 0, sizeof(pos_.w1));
    Ref r = open();
    r.n_rt = n;
    r.clear();
    close(r);
    memset(space_shadow_, 0, sizeof(space_shadow_));
    sync_space_shadow();
}

// see space_shadow_: called after every change of the position
void HostBoard::sync_space_shadow() {
    const Ref r = open();
    const int NN = nn();
    for (int idx = 0; idx < NN; idx++)
        if (r.col(idx) == 0) {
            space_shadow_[idx][0] = (int16_t)r.A(idx);
            space_shadow_[idx][1] = (int16_t)r.B(idx);
        }
}

int HostBoard::play(int colour, int row, int col) {
    const int n = pos_.n;
    // Piece::legal, Piece.cpp:28-30
    if ((colour != 1 && colour != 2) || row < 0 || col < 0 || row >= n || col >= n) return -1;
    Ref r = open();
    const int idx = row * n + col;
    if (r.col(idx) != 0) return -2;   // reference: AG_PANIC, Board.cpp:75-78
    r.put(colour, idx);
    close(r);
    sync_space_shadow();
    return 0;
}

void HostBoard::pass(int colour) {
    Ref r = open();
    r.pass(colour);
    close(r);
}

int HostBoard::replay(const uint32_t* moves, int n_moves) {
    for (int i = 0; i < n_moves; i++) {
        const int colour = (int)((moves[i] >> 16) & 0xff);
        const int row = (int)((moves[i] >> 8) & 0xff), col = (int)(moves[i] & 0xff);
        if (colour != 1 && colour != 2) return -1;
        if (row == 0xff) { pass(colour); continue; }
        const int st = play(colour, row, col);
        if (st) return st;
    }
    return 0;
}

int HostBoard::colour_at(int idx) const { return open().col(idx); }

// Board::checkNoSense, Board.cpp:883-908.  The reference has no return on the false path
// (undefined behaviour); false is taken there.
bool HostBoard::no_sense(int colour, int row, int col) const {
    const Ref r = open();
    const int n = pos_.n, enemy = 3 - colour;
    int self1 = 0, enemy1 = 0, self2 = 0, enemy2 = 0;
    for (int range = 1; range <= 2; range++) {
        int s = 0, e = 0;
        for (int i = row - range; i <= row + range; i++)
            for (int j = col - range; j <= col + range; j++) {
                if (i == row && j == col) continue;
                if (i < 0 || j < 0 || i >= n || j >= n) continue;     // checkPiece, :193-198
                const int c = r.col(i * n + j);
                if (c == colour) s++;
                else if (c == enemy) e++;
            }
        if (range == 1) { self1 = s; enemy1 = e; } else { self2 = s; enemy2 = e; }
    }
    const bool at_border = (row == 0 || row == n - 1 || col == 0 || col == n - 1);
    const bool near_border = (row <= 1 || row >= n - 2 || col <= 1 || col >= n - 2);
    return ((enemy1 + self1 == 0) && at_border) || ((enemy2 + self2 == 0) && near_border);
}

int HostBoard::check(int which, int colour, int row, int col) const {
    const int n = pos_.n;
    if ((colour != 1 && colour != 2) || row < 0 || col < 0 || row >= n || col >= n) return -1;
    const Ref r = open();
    const int idx = row * n + col;
    const Ref::Nb q = r.scan(colour, idx);
    switch (which) {
        case 0: return Ref::is_suicide(q);
        case 1: return Ref::is_kill(q);
        case 2: return Ref::is_survive(q);
        case 3: return Ref::is_dying(q);
        case 4: return Ref::is_chase(q);
        case 5: return r.true_eye(colour, idx);
        case 6: return r.compete(colour, idx);
        case 7: return no_sense(colour, row, col);
    }
    return -1;
}

int HostBoard::count_score(int colour) const { return open().count_score(colour); }

// CalcResults, GeneHatchery.cpp:393-436.  SET_SCORE (:379-390) `continue`s the column loop at the
// first neighbour that is a stone, in the order up, down, left, right; the sweep is in place.
int HostBoard::calc_result() const {
    const int n = pos_.n;
    int data[kMaxNN];
    for (int idx = 0; idx < n * n; idx++) data[idx] = (int)((pos_.w0[idx] >> 20) & 3u);
    bool changed = true;
    while (changed) {
        changed = false;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (data[i * n + j] != 0) continue;
                const int nb[4] = {i > 0 ? data[(i - 1) * n + j] : 0, i < n - 1 ? data[(i + 1) * n + j] : 0,
                                   j > 0 ? data[i * n + j - 1] : 0, j < n - 1 ? data[i * n + j + 1] : 0};
                for (int d = 0; d < 4; d++)
                    if (nb[d] == 1 || nb[d] == 2) { data[i * n + j] = nb[d]; changed = true; break; }
            }
    }
    int dscore = 0;
    for (int idx = 0; idx < n * n; idx++) dscore += data[idx] == 1 ? 1 : -1;     // :430
    return dscore;
}

int HostBoard::random_piece(int colour, uint32_t seed, const RngParams& prm) {
    Ref r = open();
    Rng rng;
    rng.seed(seed, prm);
    const int idx = r.random_piece(rng, colour);
    close(r);                                   // game_ending may have been latched, Board.cpp:407
    return idx;
}

bool HostBoard::is_candidate(int colour, int row, int col) const {
    const Ref r = open();
    const int idx = row * pos_.n + col;
    if (r.true_eye(colour, idx)) return false;
    if (r.col(idx) != 0) return false;
    if (Ref::is_suicide(r.scan(colour, idx))) return false;
    if (no_sense(colour, row, col)) return false;
    if (r.compete(colour, idx)) return false;
    return true;
}

bool HostBoard::is_eligible(int colour, int row, int col) const {
    const Ref r = open();
    const int idx = row * pos_.n + col;
    return r.col(idx) == 0 && !r.true_eye(colour, idx) && !Ref::is_suicide(r.scan(colour, idx)) &&
           !r.compete(colour, idx);
}

Position HostBoard::with_move(int colour, int idx) const {
    HostBoard copy(*this);
    Ref r = copy.open();
    r.put(colour, idx);
    copy.close(r);
    return copy.pos_;
}

// `mj->bd->data[row][col]` as the reference's compiled code reads it.  Two of the knight's-move
// tests of MyWorker::work guard one point and read another (AgentGo.cpp:185-191, 217-223), so
// `col` reaches -2..N and `row` reaches N: the read lands on the flat int array Board::data
// (Board.h:41) or, past its end, on the members that follow it: num_black, num_white, space_head,
// space_list[0..][2] (Board.h:42-45).  That is undefined behaviour in C++, but it is what both
// the MSVC build and the g++ oracle do, and it decides marks on the last row.
int HostBoard::data_at(int row, int col) const {
    const int n = pos_.n, NN = n * n;
    const int f = row * n + col;
    if (f < 0) return 0;                                // not reachable from static_units
    if (f < NN) return (int)((pos_.w0[f] >> 20) & 3u);
    const int k = f - NN;
    if (k == 0) return pos_.num_black;
    if (k == 1) return pos_.num_white;
    if (k == 2) return pos_.head;
    const int e = (k - 3) >> 1;
    return e < NN ? space_shadow_[e][(k - 3) & 1] : 0;
}

void HostBoard::static_units(int colour, int row, int col, int64_t* units_a, int64_t* units_b) const {
    const int agent = colour, enemy = 3 - colour;
    const int last = pos_.n - 1;                        // the literal 12 of the reference
    int64_t ua = 0, ub = 0;
    // on the board before the candidate is played (bd_copy, :81): :104-109
    if (check(3, agent, row, col) == 1) ua -= 1;        // checkDying
    if (no_sense(agent, row, col)) ua -= 1;             // checkNoSense
    if (check(2, agent, row, col) == 1) ub += 1;        // checkSurvive
    // everything below reads the board AFTER the candidate (and its captures), mj->bd (:83)
    HostBoard after(*this);
    after.play(agent, row, col);
    auto in = [&](int v) { return v >= 0 && v <= last; };
    // contact / diagonal / knight's-move terms, :111-240.  "Friend" is tested as data == agent+1:
    // for black that is a WHITE stone (so the enemy branch is never reached), for white it is 3
    // and never matches.  gr/gc: the point the guard tests; dr/dc: the point that is read.
    struct Term { int gr, gc, dr, dc; };
    static const Term terms[] = {
        {0, -1, 0, -1}, {-1, 0, -1, 0}, {0, 1, 0, 1}, {1, 0, 1, 0},          // :111-141
        {1, 1, 1, 1}, {1, -1, 1, -1}, {-1, 1, -1, 1}, {-1, -1, -1, -1},      // :143-174
        {-1, 2, -1, 2},                                                      // :176-183
        {-2, 1, 1, -2},                                                      // :184-191 (guard and read differ)
        {-1, -2, -1, -2}, {-2, -1, -2, -1}, {1, 2, 1, 2},                    // :192-215
        {2, -2, 2, 1},                                                       // :216-223 (guard and read differ)
        {1, -2, 1, -2}, {2, -1, 2, -1},                                      // :224-239
    };
    for (const Term& t : terms) {
        if (!(in(row + t.gr) && in(col + t.gc))) continue;
        const int v = after.data_at(row + t.dr, col + t.dc);
        if (v == agent + 1) ub += 1;
        else if (v == enemy) ub -= 1;
    }
    // cut patterns, :241-282 (these compare with `agent`), each worth 20 * bonus_b
    if (row - 1 >= 0 && row + 1 <= last && col - 1 >= 0 && col + 1 <= last) {
        auto D = [&](int dr, int dc) { return after.data_at(row + dr, col + dc); };
        if ((D(0, 1) == agent || D(0, -1) == agent || (D(-1, 1) == agent && D(-1, -1) == agent) ||
             (D(1, 1) == agent && D(1, -1) == agent)) && D(-1, 0) == enemy && D(1, 0) == enemy) ub += 20;
        if ((D(-1, 0) == agent || D(1, 0) == agent || (D(-1, 1) == agent && D(1, 1) == agent) ||
             (D(-1, -1) == agent && D(1, -1) == agent)) && D(0, 1) == enemy && D(0, -1) == enemy) ub += 20;
        if (D(1, 1) == agent && D(1, 0) == enemy && D(0, 1) == enemy && (D(0, -1) == agent || D(-1, 0) == agent)) ub += 20;
        if (D(-1, 1) == agent && D(-1, 0) == enemy && D(0, 1) == enemy && (D(0, -1) == agent || D(1, 0) == agent)) ub += 20;
        if (D(1, -1) == agent && D(1, 0) == enemy && D(0, -1) == enemy && (D(-1, 0) == agent || D(0, 1) == agent)) ub += 20;
        if (D(-1, -1) == agent && D(-1, 0) == enemy && D(0, -1) == enemy && (D(0, 1) == agent || D(1, 0) == agent)) ub += 20;
    }
    *units_a = ua;
    *units_b = ub;
}

bool HostBoard::book_move(int* row, int* col) const {
    if (pos_.n != 13) return false;
    // nine regions cut at rows/cols 4 and 9, :482-536
    bool taken[9] = {false, false, false, false, false, false, false, false, false};
    for (int i = 0; i < 13; i++)
        for (int j = 0; j < 13; j++)
            if (((pos_.w0[i * 13 + j] >> 20) & 3u) != 0)
                taken[(i < 4 ? 0 : (i < 9 ? 1 : 2)) * 3 + (j < 4 ? 0 : (j < 9 ? 1 : 2))] = true;
    // first free region in this order, :537-581
    static const int order[8][3] = {{0, 3, 3}, {8, 9, 9}, {2, 3, 9}, {6, 9, 3}, {1, 3, 5}, {7, 9, 7}, {3, 7, 3}, {5, 5, 9}};
    *row = 6; *col = 6;
    for (const auto& o : order)
        if (!taken[o[0]]) { *row = o[1]; *col = o[2]; break; }
    return true;
}

int HostBoard::export_state(int32_t* out, int cap) const {
    const int n = pos_.n, NN = n * n, words = 16 + 7 * NN;
    if (cap < words) return -1;
    const Ref r = open();
    memset(out, 0, sizeof(int32_t) * (size_t)words);
    out[0] = n; out[1] = r.num_black; out[2] = r.num_white; out[3] = r.head;
    out[4] = r.ko ? 1 : 0;
    out[5] = r.ko ? r.ko_col : 0;
    out[6] = r.ko ? r.ko_idx / n : -1;
    out[7] = r.ko ? r.ko_idx % n : -1;
    const int lm[2] = {r.last0, r.last1};
    for (int a = 0; a < 2; a++) {
        out[8 + 3 * a] = lm[a] >= 0 ? a + 1 : 0;
        out[9 + 3 * a] = lm[a] >= 0 ? lm[a] / n : -1;
        out[10 + 3 * a] = lm[a] >= 0 ? lm[a] % n : -1;
    }
    out[14] = r.ending ? 1 : 0;
    int32_t* data = out + 16;
    int32_t* root = data + NN;
    int32_t* hp = root + NN;
    int32_t* size = hp + NN;
    int32_t* next = size + NN;
    int32_t* sprev = next + NN;
    int32_t* snext = sprev + NN;
    for (int idx = 0; idx < NN; idx++) {
        const int c = r.col(idx);
        data[idx] = c;
        if (c == 0) {
            root[idx] = -1; next[idx] = -1;
            sprev[idx] = r.A(idx); snext[idx] = r.B(idx);
        } else {
            root[idx] = r.A(idx);
            next[idx] = r.B(idx) == kNil ? -1 : r.B(idx);
            sprev[idx] = snext[idx] = -1;
        }
    }
    for (int idx = 0; idx < NN; idx++) {
        if (data[idx] == 0 || root[idx] != idx) continue;
        hp[idx] = r.HP(idx);
        int sz = 0;
        for (int k = idx; k != kNil; k = r.B(k)) sz++;
        size[idx] = sz;
    }
    return words;
}

}  // namespace agb
