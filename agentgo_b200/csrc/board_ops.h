// board_ops.h — compact Go position + rules, shared by the host replay board and the
// thread-per-playout kernel (`__host__ __device__`).
//
// Replaces, for the playout path, the reference's Board / SetNode / Piece classes
// (GoContest/Board.{h,cpp}, SetNode.{h,cpp}, Piece.{h,cpp}); 126 064 B per 19x19 board
// there, 2 888 B + a few scalars here.  Behaviour is bit-exact, including the history
// dependent orders the playout policy observes:
//   * strings keep their stones in SetNode::pieces order (SetNode.cpp:78-82) as a singly
//     linked list that starts at the root stone (invariant of unionSetNode, Board.cpp:279-291);
//   * the empty-point list keeps the reference's prev/next links with self-links as end
//     markers (Board.cpp:46-54), removal (:105-118) and head insertion on capture (:311-321).
//
// Packed state, two 32-bit words per point so that a point's colour and links arrive in
// one load and a warp of independent boards is bank-conflict free when the words are
// interleaved by lane (stride template parameter):
//   w0 = A | B<<10 | colour<<20 | reserve<<22
//        stone:  A = root index of its string,      B = next stone in the string (NIL = end)
//        empty:  A = previous empty point (self = head), B = next empty point (self = last)
//   w1 = hp | tail<<16        valid at root stones: pseudo-liberties, last stone of the string
#ifndef AGB_BOARD_OPS_H_
#define AGB_BOARD_OPS_H_

#include <stdint.h>

#ifdef __CUDACC__
#define AGB_HD __host__ __device__ __forceinline__
#else
#define AGB_HD inline
#endif

namespace agb {

constexpr int kMaxN = 19;
constexpr int kMaxNN = kMaxN * kMaxN;
constexpr int kNil = 1023;
constexpr int kEndingThreshold = 100;  // Board.h:13, kept at every board size
constexpr int kEndingStones = 120;     // Board.cpp:349, kept at every board size
constexpr int kDefaultMaxMoves = 4096; // fault detector only; the reference has no cap

// Position as it travels host -> device (plain words, stride 1).
struct Position {
    uint32_t w0[kMaxNN];
    uint32_t w1[kMaxNN];
    int16_t n;
    int16_t head;        // first empty point or -1 (SPACE_HEAD_NULL, Board.h:12)
    int16_t num_black, num_white;
    int16_t ko;          // exist_compete
    int16_t ko_col;      // compete[0]: colour that may not play at ko_idx
    int16_t ko_idx;      // compete[1]*N + compete[2]
    int16_t last[2];     // last_move[colour-1] as idx, -1 = empty Piece
    int16_t ending;      // game_ending
    int16_t pad[2];
};

// Where one playout starts and how it is scored/seeded.
struct StartDesc {
    uint32_t position_id;
    uint32_t cand_idx;
    int16_t agent;       // side the diff is computed for
    int16_t first;       // colour that moves first in the pair loop
    int32_t avrg_win;
    int32_t amaf_class;  // which AMAF table this start accumulates into (Engine::genmove), else 0
};

// TinyMT32 (RFC 8682), one stream per playout; rand() := uint32 >> 17.
struct TinyMT {
    uint32_t s0, s1, s2, s3;
    AGB_HD void next_state(uint32_t mat1, uint32_t mat2) {
        uint32_t y = s3;
        uint32_t x = (s0 & 0x7fffffffu) ^ s1 ^ s2;
        x ^= (x << 1);
        y ^= (y >> 1) ^ x;
        s0 = s1;
        s1 = s2;
        s2 = x ^ (y << 10);
        s3 = y;
        // (selects, not masks: one test and two predicated XORs instead of negate, two ANDs, two XORs)
        const bool odd = (y & 1u) != 0;
        s1 = odd ? s1 ^ mat1 : s1;
        s2 = odd ? s2 ^ mat2 : s2;
    }
    AGB_HD uint32_t temper(uint32_t tmat) const {
        uint32_t t1 = s0 + (s2 >> 8);
        uint32_t t0 = s3 ^ t1;
        return (t1 & 1u) ? t0 ^ tmat : t0;
    }
    AGB_HD void init(uint32_t seed, uint32_t mat1, uint32_t mat2, uint32_t tmat) {
        uint32_t st[4] = {seed, mat1, mat2, tmat};
#pragma unroll
        for (int i = 1; i < 8; i++)
            st[i & 3] ^= (uint32_t)i + 1812433253u * (st[(i - 1) & 3] ^ (st[(i - 1) & 3] >> 30));
        if ((st[0] & 0x7fffffffu) == 0 && st[1] == 0 && st[2] == 0 && st[3] == 0) {
            st[0] = 'T'; st[1] = 'I'; st[2] = 'N'; st[3] = 'Y';
        }
        s0 = st[0]; s1 = st[1]; s2 = st[2]; s3 = st[3];
#pragma unroll
        for (int i = 0; i < 8; i++) next_state(mat1, mat2);
    }
};

struct RngParams {
    uint32_t mat1, mat2, tmat;
};

struct Rng {
    TinyMT mt;
    RngParams p;
    uint32_t draws;
    AGB_HD void seed(uint32_t s, const RngParams& prm) {
        p = prm;
        draws = 0;
        mt.init(s, p.mat1, p.mat2, p.tmat);
    }
    // n rand() calls whose values cannot matter: only the state advances
    AGB_HD void skip(int n) {
        draws += (uint32_t)n;
        for (int i = 0; i < n; i++) mt.next_state(p.mat1, p.mat2);
    }
    // the reference's rand(): 15-bit value
    AGB_HD uint32_t rand15() {
        draws++;
        mt.next_state(p.mat1, p.mat2);
        return mt.temper(p.tmat) >> 17;
    }
};

AGB_HD uint32_t playout_seed(uint64_t seed_base, uint32_t position_id, uint32_t candidate_idx,
                             uint64_t playout_idx) {
    // identical to agb_playout_seed() in include/agentgo_b200.h
    uint64_t z = seed_base;
    z += 0x9E3779B97F4A7C15ull * (playout_idx + 1);
    z ^= 0xBF58476D1CE4E5B9ull * ((uint64_t)candidate_idx + 1);
    z += 0x94D049BB133111EBull * ((uint64_t)position_id + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z ^ (z >> 32));
}

// A board whose words live at w0[i*STRIDE], w1[i*STRIDE]; scalars are members
// (registers on the device).  NFIX > 0 fixes the size at compile time.
template <int STRIDE, int NFIX = 0>
struct BoardRef {
    uint32_t* w0;
    uint32_t* w1;
    int n_rt;
    int head, num_black, num_white, ko, ko_col, ko_idx, last0, last1, ending;

    AGB_HD int n() const { return NFIX > 0 ? NFIX : n_rt; }
    AGB_HD int nn() const { return n() * n(); }

    AGB_HD uint32_t W0(int i) const { return w0[i * STRIDE]; }
    AGB_HD void setW0(int i, uint32_t v) { w0[i * STRIDE] = v; }
    AGB_HD static int fA(uint32_t w) { return (int)(w & 1023u); }
    AGB_HD static int fB(uint32_t w) { return (int)((w >> 10) & 1023u); }
    AGB_HD static int fCol(uint32_t w) { return (int)((w >> 20) & 3u); }
    AGB_HD static uint32_t pack0(int a, int b, int col) {
        return (uint32_t)a | ((uint32_t)b << 10) | ((uint32_t)col << 20);
    }
    AGB_HD int col(int i) const { return fCol(W0(i)); }
    AGB_HD int A(int i) const { return fA(W0(i)); }
    AGB_HD int B(int i) const { return fB(W0(i)); }
    AGB_HD void setA(int i, int a) { setW0(i, (W0(i) & ~1023u) | (uint32_t)a); }
    AGB_HD void setB(int i, int b) { setW0(i, (W0(i) & ~(1023u << 10)) | ((uint32_t)b << 10)); }
    AGB_HD bool reserved(int i) const { return (W0(i) >> 22) & 1u; }
    AGB_HD void set_reserved(int i) { setW0(i, W0(i) | (1u << 22)); }
    AGB_HD void clear_reserved(int i) { setW0(i, W0(i) & ~(1u << 22)); }

    AGB_HD int HP(int i) const { return (int)(w1[i * STRIDE] & 0xffffu); }
    AGB_HD int TL(int i) const { return (int)(w1[i * STRIDE] >> 16); }
    AGB_HD void setHP(int i, int hp) {
        w1[i * STRIDE] = (w1[i * STRIDE] & 0xffff0000u) | ((uint32_t)hp & 0xffffu);
    }
    AGB_HD void setHT(int i, int hp, int tl) { w1[i * STRIDE] = ((uint32_t)hp & 0xffffu) | ((uint32_t)tl << 16); }
    AGB_HD void addHP(int i, int d) { setHP(i, HP(i) + d); }

    AGB_HD int last(int colour) const { return colour == 1 ? last0 : last1; }
    AGB_HD void set_last(int colour, int idx) {
        if (colour == 1) last0 = idx; else last1 = idx;
    }

    // Board::clear, Board.cpp:27-66
    AGB_HD void clear() {
        const int NN = nn();
        for (int i = 0; i < NN; i++) {
            int prev = i == 0 ? 0 : i - 1;
            int next = i == NN - 1 ? NN - 1 : i + 1;
            setW0(i, pack0(prev, next, 0));
            w1[i * STRIDE] = 0;
        }
        head = 0;
        num_black = num_white = 0;
        ko = 0; ko_col = 0; ko_idx = -1;
        last0 = last1 = -1;
        ending = 0;
    }

    AGB_HD void load_scalars(const Position& p) {
        n_rt = p.n; head = p.head; num_black = p.num_black; num_white = p.num_white;
        ko = p.ko; ko_col = p.ko_col; ko_idx = p.ko_idx; last0 = p.last[0]; last1 = p.last[1];
        ending = p.ending;
    }
    AGB_HD void store_scalars(Position& p) const {
        p.n = (int16_t)n(); p.head = (int16_t)head; p.num_black = (int16_t)num_black;
        p.num_white = (int16_t)num_white; p.ko = (int16_t)ko; p.ko_col = (int16_t)ko_col;
        p.ko_idx = (int16_t)ko_idx; p.last[0] = (int16_t)last0; p.last[1] = (int16_t)last1;
        p.ending = (int16_t)ending; p.pad[0] = p.pad[1] = 0;
    }

    // Board::unionSetNode, Board.cpp:279-291: s1 (neighbour's string) survives; its stones come
    // first, then those of s2 (SetNode::merge_pieces, SetNode.cpp:78-82); all of s2 is relabelled.
    AGB_HD void union_strings(int s1, int s2) {
        if (s1 == s2) return;
        const int t1 = TL(s1);
        setHT(s1, HP(s1) + HP(s2), TL(s2));
        setB(t1, s2);
        for (int k = s2; k != kNil;) {
            uint32_t w = W0(k);
            setW0(k, (w & ~1023u) | (uint32_t)s1);
            k = fB(w);
        }
    }

    // Board::killSetNode, Board.cpp:293-339
    AGB_HD void kill_string(int root) {
        const int N = n();
        const bool single = (TL(root) == root);
        for (int k = root; k != kNil;) {
            const uint32_t w = W0(k);
            const int nx = fB(w);
            const int agent = fCol(w);
            if (agent == 1) num_black--; else num_white--;
            // becomes empty and is pushed at the head of the empty list (:311-321)
            if (head >= 0) {
                setW0(k, pack0(k, head, 0));
                setA(head, k);
            } else {
                setW0(k, pack0(k, k, 0));
            }
            head = k;
            // one pseudo-liberty back per adjacent stone of the other colour (:323-326)
            const int i = k / N, j = k - i * N;
            if (i > 0) { uint32_t v = W0(k - N); if (fCol(v) != 0 && fCol(v) != agent) addHP(fA(v), 1); }
            if (j > 0) { uint32_t v = W0(k - 1); if (fCol(v) != 0 && fCol(v) != agent) addHP(fA(v), 1); }
            if (i + 1 < N) { uint32_t v = W0(k + N); if (fCol(v) != 0 && fCol(v) != agent) addHP(fA(v), 1); }
            if (j + 1 < N) { uint32_t v = W0(k + 1); if (fCol(v) != 0 && fCol(v) != agent) addHP(fA(v), 1); }
            if (single) { ko = 1; ko_col = agent; ko_idx = k; }   // :328-333
            k = nx;
        }
    }

    // one neighbour step of Board::put, Board.cpp:129-137
    AGB_HD void put_touch(int agent, int idx, int nb) {
        const uint32_t v = W0(nb);
        if (fCol(v) == 0) return;
        const int sn = fA(v);
        const int hp = HP(sn) - 1;
        setHP(sn, hp);
        if (fCol(v) != agent) {
            if (hp == 0) kill_string(sn);
        } else {
            union_strings(sn, A(idx));
        }
    }

    // Board::put, Board.cpp:68-177 (the caller guarantees the point is empty)
    AGB_HD void put(int agent, int idx) {
        const int N = n();
        const int i = idx / N, j = idx - i * N;
        if (agent == 1) num_black++; else num_white++;
        ko = 0;
        // unlink from the empty list (:105-118)
        {
            const uint32_t w = W0(idx);
            const int sp = fA(w), sx = fB(w);
            if (num_black + num_white == N * N) head = -1;
            else if (sp == idx) { setA(sx, sx); head = sx; }
            else if (sx == idx) { setB(sp, sp); }
            else { setB(sp, sx); setA(sx, sp); }
        }
        // SetNode::init (SetNode.cpp:31-44): a new one-stone string, hp starts at 0
        setW0(idx, pack0(idx, kNil, agent));
        setHT(idx, 0, idx);
        // own pseudo-liberties before any capture (:123-127)
        int hp = 0;
        if (i > 0 && col(idx - N) == 0) hp++;
        if (j > 0 && col(idx - 1) == 0) hp++;
        if (i + 1 < N && col(idx + N) == 0) hp++;
        if (j + 1 < N && col(idx + 1) == 0) hp++;
        // up, left, down, right (:129-164)
        if (i > 0) put_touch(agent, idx, idx - N);
        if (j > 0) put_touch(agent, idx, idx - 1);
        if (i + 1 < N) put_touch(agent, idx, idx + N);
        if (j + 1 < N) put_touch(agent, idx, idx + 1);
        const int self = A(idx);
        const int h = HP(self) + hp;
        setHP(self, h);
        if (h == 0) kill_string(self);            // :167
        set_last(agent, idx);                     // :171
    }

    // Board::pass, Board.cpp:200-203
    AGB_HD void pass(int agent) {
        ko = 0;
        set_last(agent, -1);
    }

    // Everything the five 4-neighbour predicates need, from one scan (the reference repeats
    // the scan in each of checkSuicide/Kill/Survive/Dying/Chase, Board.cpp:480-733).
    struct Nb {
        int n_empty;
        int friend_sum;      // sum over distinct own strings of hp - multiplicity
        int friend_min_hp;   // min un-decremented hp over own strings (9999 if none)
        bool enemy_cap;      // an enemy string with hp - multiplicity == 0
        bool enemy_atari;    // an enemy string with hp - multiplicity == 1
        bool friend_cap;     // an own string with hp - multiplicity == 0
        bool friend_safe;    // an own string with hp - multiplicity > 0
    };

    AGB_HD Nb scan(int agent, int idx) const {
        const int N = n();
        const int i = idx / N, j = idx - i * N;
        int root[4], colr[4];
        Nb q;
        q.n_empty = 0; q.friend_sum = 0; q.friend_min_hp = 9999;
        q.enemy_cap = q.enemy_atari = q.friend_cap = q.friend_safe = false;
        const int nb[4] = {idx - N, idx - 1, idx + N, idx + 1};
        const bool ok[4] = {i > 0, j > 0, i + 1 < N, j + 1 < N};
#pragma unroll
        for (int m = 0; m < 4; m++) {
            root[m] = -1; colr[m] = 0;
            if (ok[m]) {
                const uint32_t v = W0(nb[m]);
                colr[m] = fCol(v);
                if (colr[m] == 0) q.n_empty++;
                else root[m] = fA(v);
            }
        }
#pragma unroll
        for (int m = 0; m < 4; m++) {
            if (root[m] < 0) continue;
            bool dup = false;
            int mult = 1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < m && root[k] == root[m]) dup = true;
                if (k > m && root[k] == root[m]) mult++;
            }
            if (dup) continue;
            const int hp = HP(root[m]);
            const int left = hp - mult;
            if (colr[m] != agent) {
                if (left == 0) q.enemy_cap = true;
                if (left == 1) q.enemy_atari = true;
            } else {
                q.friend_sum += left;
                if (hp < q.friend_min_hp) q.friend_min_hp = hp;
                if (left == 0) q.friend_cap = true;
                if (left > 0) q.friend_safe = true;
            }
        }
        return q;
    }

    AGB_HD static bool is_suicide(const Nb& q) { return q.n_empty == 0 && !q.enemy_cap && !q.friend_safe; } // :480-524
    AGB_HD static bool is_kill(const Nb& q) { return q.enemy_cap; }                                           // :526-563
    AGB_HD static bool is_survive(const Nb& q) {                                                              // :565-626
        return q.friend_cap && (q.enemy_cap || q.n_empty + q.friend_sum > q.friend_min_hp);
    }
    AGB_HD static bool is_dying(const Nb& q) { return !q.enemy_cap && q.n_empty + q.friend_sum == 1; }        // :628-679
    AGB_HD static bool is_chase(const Nb& q) { return q.enemy_atari && q.n_empty + q.friend_sum > 1; }        // :681-733

    // Board::checkTrueEye, Board.cpp:735-751
    AGB_HD bool true_eye(int agent, int idx) const {
        const int N = n();
        const int i = idx / N, j = idx - i * N;
        if (i > 0 && col(idx - N) != agent) return false;
        if (j > 0 && col(idx - 1) != agent) return false;
        if (i + 1 < N && col(idx + N) != agent) return false;
        if (j + 1 < N && col(idx + 1) != agent) return false;
        const int enemy = 3 - agent;
        int cnt = 0;
        if (i > 0 && j > 0 && col(idx - N - 1) == enemy) cnt++;
        if (i > 0 && j + 1 < N && col(idx - N + 1) == enemy) cnt++;
        if (i + 1 < N && j > 0 && col(idx + N - 1) == enemy) cnt++;
        if (i + 1 < N && j + 1 < N && col(idx + N + 1) == enemy) cnt++;
        const bool edge = (i == 0 || i == N - 1 || j == 0 || j == N - 1);
        return edge ? cnt < 1 : cnt < 2;
    }

    // Board::checkCompete, Board.cpp:753-755
    AGB_HD bool compete(int agent, int idx) const { return ko && idx == ko_idx && agent == ko_col; }

    // the filter shared by both sampling phases: dying / true eye / suicide / ko
    AGB_HD bool rejected(int agent, int idx, const Nb& q) const {
        return is_dying(q) || true_eye(agent, idx) || is_suicide(q) || compete(agent, idx);
    }

    // Board::getRandomPieceComplex, Board.cpp:413-478; returns idx or -1 (pass)
    template <class R>
    AGB_HD int random_piece_complex(R& rng, int agent) {
        const int NN = nn();
        const int spaces = NN - num_black - num_white;
        if (spaces == 0) return -1;
        int reserved_total = 0;
        for (int idx = head;;) {
            const Nb q = scan(agent, idx);
            if (true_eye(agent, idx) || is_suicide(q) || compete(agent, idx)) {
                set_reserved(idx);
                reserved_total++;
            }
            const int nx = B(idx);
            if (nx == idx) break;
            idx = nx;
        }
        const int range = spaces - reserved_total;
        int pick = -1;
        if (range != 0) {
            int rnd = (int)(rng.rand15() % (uint32_t)range);
            for (int idx = head;;) {
                if (!reserved(idx)) {
                    if (rnd == 0) { pick = idx; break; }
                    rnd--;
                }
                idx = B(idx);
            }
        }
        for (int idx = head;;) {                 // resetReserve, :764-768
            clear_reserved(idx);
            const int nx = B(idx);
            if (nx == idx) break;
            idx = nx;
        }
        return pick;
    }

    // Board::getRandomPiece, Board.cpp:341-411; returns idx or -1 (pass)
    template <class R>
    AGB_HD int random_piece(R& rng, int agent) {
        const int N = n();
        if (ending) {                                           // :348-351
            if (num_black + num_white <= kEndingStones) ending = 0;
            else return random_piece_complex(rng, agent);
        }
        const int enemy = 3 - agent;
        const int lm = last(enemy);
        if (lm >= 0 && col(lm) == enemy) {                      // :355-380
            const int er = lm / N, ec = lm - er * N;
            for (int k = 0; k < 9; k++) {
                const int i = er - 1 + (int)(rng.rand15() % 3u);
                const int j = ec - 1 + (int)(rng.rand15() % 3u);
                if (i < 0 || i >= N || j < 0 || j >= N) continue;
                const int idx = i * N + j;
                if (col(idx) != 0) continue;
                const Nb q = scan(agent, idx);
                if ((is_kill(q) || (k < 6 && is_survive(q)) || (k < 4 && is_chase(q))) &&
                    !rejected(agent, idx, q))
                    return idx;
            }
        }
        for (int count = 1;; count++) {                         // :382-406
            const int row = (int)(rng.rand15() % (uint32_t)N);
            const int cl = (int)(rng.rand15() % (uint32_t)N);
            const int idx = row * N + cl;
            bool bad = col(idx) != 0;
            if (!bad) {
                const Nb q = scan(agent, idx);
                bad = rejected(agent, idx, q);
            }
            if (count == kEndingThreshold) {                    // even if the 100th draw passed
                ending = 1;
                return random_piece_complex(rng, agent);
            }
            if (!bad) return idx;
        }
    }

    // Board::countScore, Board.cpp:910-925
    AGB_HD int count_score(int agent) const {
        const int N = n();
        int count = 0;
        for (int i = 0; i < N; i++) {
            int left = col(i * N + 1);                          // j == 0 looks right (:918)
            for (int j = 0; j < N; j++) {
                const int c = col(i * N + j);
                if (c == agent) count++;
                else if (c == 0 && left == agent) count++;
                left = c;
            }
        }
        return count;
    }

    // pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443).
    // returns diff = countScore(agent) - countScore(enemy); *moves counts stones placed;
    // max_moves is a fault detector (returns INT16_MIN when hit).
    template <class R>
    AGB_HD int run_playout(R& rng, int agent, int first, int max_moves, uint32_t* moves) {
        const int second = 3 - first;
        bool go1 = true, go2 = true;
        uint32_t mv = 0;
        while (go1 || go2) {
            int r = random_piece(rng, first);
            if (r >= 0) { put(first, r); go1 = true; mv++; }
            else { go1 = false; pass(first); }
            r = random_piece(rng, second);
            if (r >= 0) { put(second, r); go2 = true; mv++; }
            else { go2 = false; pass(second); }
            if ((int)mv > max_moves) { *moves += mv; return -32768; }
        }
        *moves += mv;
        return count_score(agent) - count_score(3 - agent);
    }
};

}  // namespace agb

#endif  // AGB_BOARD_OPS_H_
