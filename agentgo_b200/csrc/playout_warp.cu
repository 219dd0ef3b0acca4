// playout_warp.cu — one WARP per playout (sm_100a).
//
// One board per warp in shared memory, so 32 playouts are resident per SM at 19x19 with 64
// registers per thread (the thread-per-playout variant fits 64 boards but pays 32-way
// divergence and has 2 warps per SM to hide latency with).
//
//  * The board is padded to (N+2)x(N+2) with a border of colour 3 (PaddedPosition,
//    kernel_args.h): neighbour and diagonal reads need no bounds checks, and every stored index
//    (links, roots, tails, head, last moves, ko point) is a padded index.
//  * State updates (put / capture / merge / pass) run WARP-UNIFORM: every lane executes the
//    same scalar code on the same shared-memory words (same address, same value), so there is
//    no divergence and no intra-warp communication.
//  * The policy's predicates are evaluated LANE-PARALLEL: lane = (cell, direction) with
//    8 cells x 4 directions = 32 lanes.  In the neighbourhood phase of Board::getRandomPiece
//    (Board.cpp:354-380) the 8 cells are the 3x3 block around the opponent's last move (its
//    centre is occupied by precondition), evaluated ONCE per move into three 8-bit masks
//    (kill / survive / chase, already filtered by dying / true eye / suicide / ko).  The
//    reference's 9-iteration loop then degenerates to two RNG draws + a bit test per
//    iteration, consuming exactly the reference's draw sequence (the predicates are pure
//    functions of the state, which does not change inside the loop).
//  * When all three masks are empty (75 % of the moves) none of the 18 draws can be accepted;
//    the generator is advanced by 18 steps at once: TinyMT32's transition is linear over GF(2),
//    so T^18 is applied as 32 per-lane nibble-table lookups (one LDG.128 per lane) folded by
//    four redux.xor — about a dozen instructions instead of 18 x 12.
//  * Clone, scoring and the complex-mode legality pass are strided over the lanes; complex mode
//    first compacts the empty points with ballots.
//  * The RNG (TinyMT32, one stream per playout) is warp-uniform in registers.
//
// The code is written for a SMALL INSTRUCTION FOOTPRINT: 32 warps per SM sit at 32 different
// program counters, and the first version of this kernel (every helper force-inlined, both
// movers unrolled: 6 150 SASS instructions, 98 KB) spent 70 % of its stall samples waiting for
// instruction fetch (profiles/r01_warp_v1_*.txt).  Hence: one move loop for both colours, one
// capture site, one complex-mode site, rolled loops.  Since then the kernel is bound by the
// integer ALU pipe (profiles/r01_warp_v3_*.txt), so the remaining lever is instruction count.
#include "kernel_args.h"

namespace agb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kBorder = 3;

// Shared-memory array view.  In the AGB_CHECK_BOUNDS build (agentgo_b200/build.py --debug; the
// pool has no compute-sanitizer) every index is checked and violations are counted in
// g_bounds_violations, which agb_debug_violations() returns; release builds index directly.
#ifdef AGB_CHECK_BOUNDS
__device__ unsigned long long g_bounds_violations;
template <typename T>
struct SmemArr {
    T* p;
    int n;
    __device__ __forceinline__ T& operator[](int i) const {
        if ((unsigned)i >= (unsigned)n) { atomicAdd(&g_bounds_violations, 1ull); i = 0; }
        return p[i];
    }
};
#else
template <typename T>
struct SmemArr {
    T* p;
    __device__ __forceinline__ T& operator[](int i) const { return p[i]; }
};
#endif

__device__ __forceinline__ int fA(uint32_t w) { return (int)(w & 1023u); }
__device__ __forceinline__ int fB(uint32_t w) { return (int)((w >> 10) & 1023u); }
__device__ __forceinline__ int fCol(uint32_t w) { return (int)((w >> 20) & 3u); }
__device__ __forceinline__ bool is_stone(int col) { return ((col + 1) & 2) != 0; }   // 1 or 2
__device__ __forceinline__ uint32_t pack0(int a, int b, int col) {
    return (uint32_t)a | ((uint32_t)b << 10) | ((uint32_t)col << 20);
}

template <int N>
struct WarpBoard {
    static constexpr int W = N + 2;
    static constexpr int NP = W * W;
    static constexpr int NN = N * N;
    SmemArr<uint32_t> w0;        // A | B<<10 | colour<<20 | reserve<<22   (field meaning: board_ops.h)
    SmemArr<uint32_t> w1;        // hp | tail<<16 at root stones
    SmemArr<uint16_t> scratch;   // NN entries: compacted list of empty points (complex mode)
    int lane;
    int off_nb;          // offset of this lane's orthogonal neighbour (dir = lane & 3: up, left, down, right)
    int off_dg;          // offset of this lane's diagonal
    int off_cell;        // offset of this lane's cell in the 3x3 block around a point (cell = lane >> 2)
    int head;            // first empty point, -1 = none
    int stones;          // num_black + num_white (only the sum is ever read on this path)
    int ko_key;          // exist_compete ? compete colour<<16 | idx : -1
    int last0, last1;    // last_move[colour-1] as idx, -1 = none
    int ending;          // game_ending
    TinyMT mt;
    uint32_t draws;

    __device__ __forceinline__ void init_lane(int l) {
        lane = l;
        const int dir = l & 3;
        off_nb = dir == 0 ? -W : (dir == 1 ? -1 : (dir == 2 ? W : 1));
        off_dg = (dir < 2 ? -W : W) + ((dir & 1) ? 1 : -1);
        const int cell = l >> 2;
        const int cc = cell + (cell >= 4);                          // skip the centre of the 3x3 block
        off_cell = (cc / 3 - 1) * W + (cc % 3 - 1);
    }

    __device__ __forceinline__ uint32_t rand15(const RngParams& p) {
        mt.next_state(p.mat1, p.mat2);
        return mt.temper(p.tmat) >> 17;
    }
    // rand() % M for a 15-bit value and a small constant M: one IMAD.HI and one IMAD
    // (q = floor(v * ceil(2^32 / M) / 2^32) is exact for v < 2^15)
    template <uint32_t M>
    __device__ __forceinline__ uint32_t rand_mod(const RngParams& p) {
        const uint32_t v = rand15(p);
        const uint32_t q = __umulhi(v, (uint32_t)((0x100000000ull + M - 1) / M));
        return v - q * M;
    }

    // advance the generator by 18 steps (values not needed): s' = T^18 s over GF(2)
    __device__ __forceinline__ void jump18(const JumpTable* __restrict__ jt) {
        const uint32_t word = lane < 16 ? (lane < 8 ? mt.s0 : mt.s1) : (lane < 24 ? mt.s2 : mt.s3);
        const uint32_t nib = (word >> ((lane & 7) * 4)) & 15u;
        const uint4 t = __ldg(&jt->e[lane * 16 + nib]);
        mt.s0 = __reduce_xor_sync(kFull, t.x);
        mt.s1 = __reduce_xor_sync(kFull, t.y);
        mt.s2 = __reduce_xor_sync(kFull, t.z);
        mt.s3 = __reduce_xor_sync(kFull, t.w);
    }

    // Lane-parallel evaluation of up to 8 points at once: lane = 4*cell + dir evaluates
    // neighbour `dir` (up, left, down, right) of its cell `cidx` (padded index; a border cell is
    // never `empty`); the four lanes of a cell hold the same result.  Restates
    // checkSuicide/Kill/Survive/Dying/Chase/TrueEye/Compete (Board.cpp:480-755).
    //   bad     = dying | true eye | suicide | ko   (the filter of both sampling phases)
    //   reserve = true eye | suicide | ko           (complex mode: no checkDying, :425-427)
    // Force-inlined at three call sites.  MODE selects what is computed, because the warp-sync
    // shuffles and ballots have side effects the compiler will not remove on its own:
    //   kEvalAll     neighbourhood phase: everything
    //   kEvalBad     rejection phase: `bad` only (no survive / chase: no fmin, fcap, eatari)
    //   kEvalReserve complex mode: `reserve` only (additionally no friend sum)
    struct Eval {
        bool empty, kill, survive, chase, bad, reserve;
    };
    enum { kEvalAll = 0, kEvalBad = 1, kEvalReserve = 2 };
    template <int MODE>
    __device__ __forceinline__ Eval eval_cells(int agent, int cidx) const {
        const unsigned nib = 15u << (lane & ~3);                    // this cell's four lanes
        const uint32_t wc = w0[cidx];
        const bool cell_on = fCol(wc) != kBorder;
        const uint32_t wn = w0[cell_on ? cidx + off_nb : 0];        // w0[0] is a border cell
        const int ncol = fCol(wn);
        const bool nb_stone = is_stone(ncol);
        const int root = nb_stone ? fA(wn) : -1 - (lane & 3);       // distinct negatives per direction
        const int hp = (int)(w1[nb_stone ? root : 0] & 0xffffu);

        // strings seen from several directions count once, with their multiplicity as minus_hp
        // (Board.cpp:506-515): lanes of this cell holding the same root
        const unsigned same = __match_any_sync(kFull, root) & nib;
        const int left = hp - __popc(same);
        const bool first = (same & ((1u << lane) - 1u)) == 0;
        const bool enemy = ncol == 3 - agent;
        const bool frnd = ncol == agent;

        const int n_empty = __popc(__ballot_sync(kFull, ncol == 0) & nib);
        const bool ecap = (__ballot_sync(kFull, enemy && left == 0) & nib) != 0;
        const bool fsafe = (__ballot_sync(kFull, frnd && left > 0) & nib) != 0;
        const bool suicide = n_empty == 0 && !ecap && !fsafe;       // Board.cpp:480-524
        bool dying = false, survive = false, chase = false;
        if (MODE != kEvalReserve) {
            int fsum = (frnd && first) ? left : 0;
            fsum += __shfl_xor_sync(kFull, fsum, 1);
            fsum += __shfl_xor_sync(kFull, fsum, 2);
            const int tot = n_empty + fsum;
            dying = !ecap && tot == 1;                              // :628-679
            if (MODE == kEvalAll) {
                const bool eatari = (__ballot_sync(kFull, enemy && left == 1) & nib) != 0;
                const bool fcap = (__ballot_sync(kFull, frnd && left == 0) & nib) != 0;
                int fmin = frnd ? hp : 9999;
                fmin = min(fmin, __shfl_xor_sync(kFull, fmin, 1));
                fmin = min(fmin, __shfl_xor_sync(kFull, fmin, 2));
                survive = fcap && (ecap || tot > fmin);             // :565-626
                chase = eatari && tot > 1;                          // :681-733
            }
        }

        // Board::checkTrueEye, :735-751 — lane `dir` also looks at one diagonal; a cell is on the
        // edge iff one of its orthogonal neighbours is a border cell
        const bool edge = (__ballot_sync(kFull, ncol == kBorder) & nib) != 0;
        const bool own4 = (__ballot_sync(kFull, ncol != kBorder && ncol != agent) & nib) == 0;
        const int dcol = fCol(w0[cell_on ? cidx + off_dg : 0]);
        const int cnt = __popc(__ballot_sync(kFull, dcol == 3 - agent) & nib);
        const bool eye = own4 && cnt < (edge ? 1 : 2);
        const bool ko = ko_key == ((agent << 16) | cidx);           // :753-755

        Eval e;
        e.empty = fCol(wc) == 0;
        e.kill = ecap;                                              // :526-563
        e.survive = survive;
        e.chase = chase;
        e.reserve = eye || suicide || ko;
        e.bad = e.reserve || dying;
        return e;
    }

    // Board::getRandomPieceComplex, Board.cpp:413-478.  Pass 1 (reserve flags): the empty points
    // are compacted with ballots, then evaluated 8 per round with eval_cells; the ordered walk
    // that maps rnd to a point follows the empty list and is warp-uniform.
    __device__ __forceinline__ int random_piece_complex(const RngParams& prm, int agent) {
        const int spaces = NN - stones;
        if (spaces == 0) return -1;
        int n_e = 0;
#pragma unroll 1
        for (int base = 0; base < NP; base += 32) {
            const int idx = base + lane;
            const bool e = idx < NP && fCol(w0[idx]) == 0;
            const unsigned m = __ballot_sync(kFull, e);
            if (e) scratch[n_e + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idx;
            n_e += __popc(m);
        }
        __syncwarp();
        int reserved_total = 0;
#pragma unroll 1
        for (int base = 0; base < n_e; base += 8) {
            const int k = base + (lane >> 2);
            const int idx = k < n_e ? (int)scratch[k] : 0;          // 0 is a border cell
            const Eval f = eval_cells<kEvalReserve>(agent, idx);
            const bool res = f.empty && f.reserve;
            if (res && (lane & 3) == 0) w0[idx] |= (1u << 22);      // addReserve, :757-761
            reserved_total += __popc(__ballot_sync(kFull, res) & 0x11111111u);
        }
        __syncwarp();
        const int range = spaces - reserved_total;
        int pick = -1;
        if (range != 0) {
            draws++;
            int rnd = (int)(rand15(prm) % (uint32_t)range);
#pragma unroll 1
            for (int idx = head;;) {
                const uint32_t w = w0[idx];
                if (!((w >> 22) & 1u)) {
                    if (rnd == 0) { pick = idx; break; }
                    rnd--;
                }
                idx = fB(w);
            }
        }
        if (reserved_total) {                                       // resetReserve, :764-768
#pragma unroll 1
            for (int k = lane; k < n_e; k += 32) w0[scratch[k]] &= ~(1u << 22);
            __syncwarp();
        }
        return pick;
    }

    // Board::getRandomPiece, Board.cpp:341-411; returns the (padded) point or -1 (pass)
    __device__ __forceinline__ int random_piece(const RngParams& prm, const JumpTable* jt, int agent) {
        bool complex_mode = false;
        if (ending) {                                               // :348-351
            if (stones <= kEndingStones) ending = 0;
            else complex_mode = true;
        }
        if (!complex_mode) {
            const int lm = agent == 1 ? last1 : last0;              // opponent's last move
            if (lm >= 0 && fCol(w0[lm]) == 3 - agent) {             // :355-380
                const Eval f = eval_cells<kEvalAll>(agent, lm + off_cell);
                const bool ok = f.empty && !f.bad;
                const unsigned K = __ballot_sync(kFull, ok && f.kill);
                const unsigned S = __ballot_sync(kFull, ok && f.survive);
                const unsigned C = __ballot_sync(kFull, ok && f.chase);
                draws += 18;
                if ((K | S | C) == 0) {
                    // no cell can be accepted: the loop runs its 9 iterations, 2 draws each
                    jump18(jt);
                } else {
#pragma unroll 1
                    for (int k = 0; k < 9; k++) {
                        const uint32_t a = rand_mod<3>(prm);
                        const uint32_t b = rand_mod<3>(prm);
                        const uint32_t c9 = a * 3u + b;
                        const uint32_t c8 = c9 - (c9 > 4u);
                        const unsigned m = K | (k < 6 ? S : 0u) | (k < 4 ? C : 0u);
                        if (c9 != 4 && ((m >> (4 * c8)) & 1u)) {    // centre (c9 == 4) is occupied
                            draws -= 2 * (8 - k);
                            return lm + (int)(a * W + b) - (W + 1);
                        }
                    }
                }
            }
#pragma unroll 1
            for (int count = 1;; count++) {                         // :382-406
                draws += 2;
                const uint32_t row = rand_mod<N>(prm);
                const uint32_t cl = rand_mod<N>(prm);
                const int idx = (int)(row * W + cl + (W + 1));
                bool bad = fCol(w0[idx]) != 0;
                if (count == kEndingThreshold) {                    // even if the 100th draw passed
                    ending = 1;
                    break;
                }
                if (!bad) {
                    bad = eval_cells<kEvalBad>(agent, idx).bad;
                    if (!bad) return idx;
                }
            }
        }
        return random_piece_complex(prm, agent);
    }

    // Board::killSetNode, Board.cpp:293-339 (warp-uniform)
    __device__ __forceinline__ void kill_string(int root) {
        const bool single = (int)(w1[root] >> 16) == root;
#pragma unroll 1
        for (int k = root; k != kNil;) {
            const uint32_t w = w0[k];
            const int nx = fB(w);
            const int agent = fCol(w);
            stones--;
            if (head >= 0) {                                        // push at the list head, :311-321
                w0[k] = pack0(k, head, 0);
                w0[head] = (w0[head] & ~1023u) | (uint32_t)k;
            } else {
                w0[k] = pack0(k, k, 0);
            }
            head = k;
#pragma unroll
            for (int d = 0; d < 4; d++) {                           // one liberty back, :323-326
                const uint32_t v = w0[k + (d == 0 ? -W : (d == 1 ? -1 : (d == 2 ? W : 1)))];
                if (fCol(v) == 3 - agent) w1[fA(v)] += 1u;
            }
            if (single) ko_key = (agent << 16) | k;                 // :328-333
            k = nx;
        }
    }

    // Board::put, Board.cpp:68-177 (warp-uniform; the point is empty)
    __device__ __forceinline__ void put(int agent, int idx) {
        stones++;
        ko_key = -1;
        {                                                           // unlink from the empty list, :105-118
            const uint32_t w = w0[idx];
            const int sp = fA(w), sx = fB(w);
            if (stones == NN) head = -1;
            else if (sp == idx) { w0[sx] = (w0[sx] & ~1023u) | (uint32_t)sx; head = sx; }
            else if (sx == idx) { w0[sp] = (w0[sp] & ~(1023u << 10)) | ((uint32_t)sp << 10); }
            else {
                w0[sp] = (w0[sp] & ~(1023u << 10)) | ((uint32_t)sx << 10);
                w0[sx] = (w0[sx] & ~1023u) | (uint32_t)sp;
            }
        }
        w0[idx] = pack0(idx, kNil, agent);                          // SetNode::init, SetNode.cpp:31-44
        w1[idx] = (uint32_t)idx << 16;
        // lane d (mod 4) looks at neighbour d (up, left, down, right) once: own liberties before
        // any capture (:123-127) and the set of directions that hold a stone
        const int nb_l = idx + off_nb;
        const int col_l = fCol(w0[nb_l]);
        const int hp_self = __popc(__ballot_sync(kFull, col_l == 0) & 15u);
        unsigned todo = (__ballot_sync(kFull, is_stone(col_l)) & 15u) | 16u;
        // up, left, down, right (:129-164), then the stone's own string (:165-167) as step 4
#pragma unroll 1
        while (todo) {
            const int d = __ffs(todo) - 1;
            todo &= todo - 1;
            int victim = -1;
            if (d < 4) {
                const int nb = __shfl_sync(kFull, nb_l, d);
                const uint32_t v = w0[nb];
                if (fCol(v) == 0) continue;                         // captured earlier in this put
                const int sn = fA(v);
                const uint32_t h = w1[sn] - 1u;                     // sn->hp -= 1
                w1[sn] = h;
                if (fCol(v) != agent) {
                    if ((h & 0xffffu) == 0) victim = sn;
                } else {
                    // Board::unionSetNode(sn, getSet(idx)), :279-291: sn survives, its stones first
                    const int s2 = fA(w0[idx]);
                    if (s2 != sn) {
                        const uint32_t h2 = w1[s2];
                        const int t1 = (int)(h >> 16);
                        w1[sn] = ((h + h2) & 0xffffu) | (h2 & 0xffff0000u);
                        w0[t1] = (w0[t1] & ~(1023u << 10)) | ((uint32_t)s2 << 10);
#pragma unroll 1
                        for (int k = s2; k != kNil;) {
                            const uint32_t w = w0[k];
                            w0[k] = (w & ~1023u) | (uint32_t)sn;
                            k = fB(w);
                        }
                    }
                }
            } else {
                const int self = fA(w0[idx]);
                const uint32_t h = w1[self] + (uint32_t)hp_self;
                w1[self] = h;
                if ((h & 0xffffu) == 0) victim = self;
            }
            if (victim >= 0) kill_string(victim);
        }
        if (agent == 1) last0 = idx; else last1 = idx;             // :171
    }

    // Board::countScore(agent) - countScore(enemy), Board.cpp:910-925, strided over the lanes:
    // an empty point counts for the colour of its LEFT neighbour (right one in column 0)
    __device__ __forceinline__ int score_diff(int agent) const {
        int d = 0;
#pragma unroll 1
        for (int idx = W + lane; idx < NP - W; idx += 32) {
            const int c = fCol(w0[idx]);
            if (c == kBorder) continue;
            int owner = c;
            if (c == 0) {
                const int l = fCol(w0[idx - 1]);
                owner = l != kBorder ? l : fCol(w0[idx + 1]);
            }
            d += (owner == agent) - (owner == 3 - agent);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
        return d;
    }
};

// per-warp shared memory: w0[NP], w1[NP], scratch[NN] (16 bit), rounded to 32 words
template <int N>
struct BoardWords {
    static constexpr int NP = (N + 2) * (N + 2);
    static constexpr int value = ((2 * NP + (N * N + 1) / 2 + 31) / 32) * 32;
};

template <int N, int WARPS, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, BLOCKS_PER_SM) playout_warp_kernel(const KernelArgs a) {
    extern __shared__ uint32_t smem[];
    typedef WarpBoard<N> WB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    WB bd;
    uint32_t* base = smem + (size_t)warp * BoardWords<N>::value;
    bd.w0.p = base;
    bd.w1.p = base + WB::NP;
    bd.scratch.p = reinterpret_cast<uint16_t*>(base + 2 * WB::NP);
#ifdef AGB_CHECK_BOUNDS
    bd.w0.n = WB::NP; bd.w1.n = WB::NP; bd.scratch.n = WB::NN;
#endif
    bd.init_lane(lane);

    unsigned long long moves = 0, draws = 0;
    unsigned int faults = 0, played = 0;

#pragma unroll 1
    for (;;) {
        unsigned long long l = 0;
        if (lane == 0) l = atomicAdd(a.work_counter, 1ull);
        l = __shfl_sync(kFull, l, 0);
        if (l >= (unsigned long long)a.local_units) break;
        const int64_t g = (int64_t)l * a.shard_count + a.shard_rank;
        const int s = (int)(g / a.P);
        const int64_t p = g - (int64_t)s * a.P;
        const PaddedPosition& pos = a.pstarts[s];
        const StartDesc d = a.descs[s];

        __syncwarp();
#pragma unroll 1
        for (int i = lane; i < WB::NP; i += 32) {                   // Board::clone, Board.cpp:216-254
            bd.w0[i] = pos.w0[i];
            bd.w1[i] = pos.w1[i];
        }
        bd.head = pos.head;
        bd.stones = pos.stones;
        bd.ko_key = pos.ko_key_col > 0 ? ((pos.ko_key_col << 16) | pos.ko_idx) : -1;
        bd.last0 = pos.last[0];
        bd.last1 = pos.last[1];
        bd.ending = pos.ending;
        bd.draws = 0;
        __syncwarp();
        {
            // tinymt32_init (RFC 8682)
            uint32_t st[4] = {playout_seed(a.seed_base, d.position_id, d.cand_idx, (uint64_t)p),
                              a.rng.mat1, a.rng.mat2, a.rng.tmat};
#pragma unroll
            for (int i = 1; i < 8; i++)
                st[i & 3] ^= (uint32_t)i + 1812433253u * (st[(i - 1) & 3] ^ (st[(i - 1) & 3] >> 30));
            if ((st[0] & 0x7fffffffu) == 0 && st[1] == 0 && st[2] == 0 && st[3] == 0) {
                st[0] = 'T'; st[1] = 'I'; st[2] = 'N'; st[3] = 'Y';
            }
            bd.mt.s0 = st[0]; bd.mt.s1 = st[1]; bd.mt.s2 = st[2]; bd.mt.s3 = st[3];
#pragma unroll 1
            for (int i = 0; i < 8; i++) bd.mt.next_state(a.rng.mat1, a.rng.mat2);
        }

        // pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443), one move
        // per iteration: the loop ends after a pair in which neither side moved.
        const int first = d.first;
        bool first_moved = true, fault = false;
        uint32_t mv = 0;
#pragma unroll 1
        for (int t = 0;; t++) {
            const int colour = (t & 1) ? 3 - first : first;
            const int r = bd.random_piece(a.rng, a.jump18, colour);
            if (r >= 0) {
                bd.put(colour, r);
                mv++;
            } else {                                                // Board::pass, Board.cpp:200-203
                bd.ko_key = -1;
                if (colour == 1) bd.last0 = -1; else bd.last1 = -1;
            }
            if (!(t & 1)) first_moved = r >= 0;
            else if (!first_moved && r < 0) break;
            if ((int)mv > a.max_moves) { fault = true; break; }
        }
        __syncwarp();
        const int diff = fault ? -32768 : bd.score_diff(d.agent);
        moves += mv;
        draws += bd.draws;
        played++;
        if (fault) faults++;
        if (lane == 0) {
            if (a.diffs) a.diffs[g] = (int16_t)diff;
            if (!fault) {
                const int w = diff - d.avrg_win;                     // AgentGo.cpp:327
                unsigned long long* c = (unsigned long long*)a.counters;
                if (w >= 0) {                                        // sign test of AgentGo.cpp:328
                    atomicAdd(&c[s], 1ull);
                    atomicAdd(&c[a.n_starts + s], (unsigned long long)w);
                } else {
                    atomicAdd(&c[2 * a.n_starts + s], (unsigned long long)(long long)w);
                }
            }
        }
    }
    if (lane == 0) {
        atomicAdd(&a.stats[kStatMoves], moves);
        atomicAdd(&a.stats[kStatDraws], draws);
        if (faults) atomicAdd(&a.stats[kStatFaults], (unsigned long long)faults);
        atomicAdd(&a.stats[kStatPlayouts], (unsigned long long)played);
    }
}

template <int N, int WARPS, int BLOCKS_PER_SM>
cudaError_t launch_w(const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    const size_t smem = (size_t)WARPS * BoardWords<N>::value * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(playout_warp_kernel<N, WARPS, BLOCKS_PER_SM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t want = (a.local_units + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)sm_count * BLOCKS_PER_SM;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (info) { info->grid = grid; info->block = WARPS * 32; info->smem = smem; info->name = "playout_warp_kernel"; }
    playout_warp_kernel<N, WARPS, BLOCKS_PER_SM><<<grid, WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

unsigned long long debug_violations() {
#ifdef AGB_CHECK_BOUNDS
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_bounds_violations, sizeof(v)) != cudaSuccess) return ~0ull;
    return v;
#else
    return 0;
#endif
}

// Host side of the padded layout: re-index a Position (board_ops.h) into a PaddedPosition.
void to_padded(const Position& p, PaddedPosition* out) {
    const int n = p.n, w = n + 2;
    auto pad = [&](int idx) { return (idx / n + 1) * w + (idx % n + 1); };
    for (int i = 0; i < kMaxNP; i++) { out->w0[i] = (uint32_t)kBorder << 20; out->w1[i] = 0; }
    for (int idx = 0; idx < n * n; idx++) {
        const uint32_t v0 = p.w0[idx], v1 = p.w1[idx];
        const int col = (int)((v0 >> 20) & 3u);
        const int a = (int)(v0 & 1023u), b = (int)((v0 >> 10) & 1023u);
        const int pa = pad(a), pb = (col != 0 && b == kNil) ? kNil : pad(b);
        out->w0[pad(idx)] = (uint32_t)pa | ((uint32_t)pb << 10) | ((uint32_t)col << 20);
        const bool is_root = col != 0 && a == idx;                  // hp/tail are meaningful at roots only
        out->w1[pad(idx)] = is_root ? ((v1 & 0xffffu) | ((uint32_t)pad((int)(v1 >> 16)) << 16)) : 0u;
    }
    out->head = (int16_t)(p.head >= 0 ? pad(p.head) : -1);
    out->stones = (int16_t)(p.num_black + p.num_white);
    out->ko_key_col = (int16_t)(p.ko ? p.ko_col : 0);
    out->ko_idx = (int16_t)(p.ko ? pad(p.ko_idx) : 0);
    out->last[0] = (int16_t)(p.last[0] >= 0 ? pad(p.last[0]) : -1);
    out->last[1] = (int16_t)(p.last[1] >= 0 ? pad(p.last[1]) : -1);
    out->ending = p.ending;
    out->pad = 0;
}

// T^18 of TinyMT32 as a nibble table (see JumpTable): column b of T^18 is the state reached from
// the unit vector e_b after 18 transitions (the transition is linear over GF(2)).
void build_jump18(const RngParams& prm, JumpTable* out) {
    uint32_t col[128][4];
    for (int b = 0; b < 128; b++) {
        TinyMT t;
        t.s0 = t.s1 = t.s2 = t.s3 = 0;
        (b < 32 ? t.s0 : (b < 64 ? t.s1 : (b < 96 ? t.s2 : t.s3))) = 1u << (b & 31);
        for (int i = 0; i < 18; i++) t.next_state(prm.mat1, prm.mat2);
        col[b][0] = t.s0; col[b][1] = t.s1; col[b][2] = t.s2; col[b][3] = t.s3;
    }
    for (int l = 0; l < 32; l++)
        for (int v = 0; v < 16; v++) {
            uint32_t acc[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++)
                if (v & (1 << k))
                    for (int wd = 0; wd < 4; wd++) acc[wd] ^= col[4 * l + k][wd];
            out->e[l * 16 + v] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        }
}

cudaError_t launch_playout_warp(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    // 8 warps per block, 4 blocks per SM: 32 resident playouts per SM, 64 registers per thread
    switch (n) {
        case 9: return launch_w<9, 8, 4>(a, sm_count, s, info);
        case 13: return launch_w<13, 8, 4>(a, sm_count, s, info);
        case 19: return launch_w<19, 8, 4>(a, sm_count, s, info);
    }
    return cudaErrorInvalidValue;
}

}  // namespace agb
