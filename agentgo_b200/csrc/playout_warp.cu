// playout_warp.cu — one WARP per playout (sm_100a).
//
// One board (2*N*N words) per warp in shared memory, so 32+ playouts are resident per SM at
// 19x19 (the thread-per-playout variant fits 64 but pays 32-way divergence).
//
//  * State updates (put / capture / merge / pass, board_ops.h) run WARP-UNIFORM: every lane
//    executes the same scalar code on the same shared-memory words (same address, same value),
//    so there is no divergence and no intra-warp communication.
//  * The policy's predicates are evaluated LANE-PARALLEL: lane = (cell, direction) with
//    8 cells x 4 directions = 32 lanes.  In the neighbourhood phase of Board::getRandomPiece
//    (Board.cpp:354-380) the 8 cells are the 3x3 block around the opponent's last move (its
//    centre is occupied by precondition), evaluated ONCE per move into three 8-bit masks
//    (kill / survive / chase, already filtered by dying / true eye / suicide / ko).  The
//    reference's 9-iteration loop then degenerates to two RNG draws + a bit test per
//    iteration, consuming exactly the reference's draw sequence (the predicates are pure
//    functions of the state, which does not change inside the loop).  When all three masks
//    are empty the 18 draws cannot be accepted and the generator is only stepped.
//  * Clone, scoring and the complex-mode legality pass are strided over the lanes.
//  * The RNG (TinyMT32, one stream per playout) is warp-uniform in registers.
#include "kernel_args.h"

namespace agb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

template <int N>
struct WarpBoard : BoardRef<1, N> {
    typedef BoardRef<1, N> Base;
    int lane;

    struct Flags {
        bool empty;      // cell on board and empty
        bool kill, survive, chase;
        bool dying, eye, suicide, ko;
    };

    // Lane-parallel evaluation of up to 8 points at once: lane = 4*cell + dir evaluates
    // neighbour `dir` (up, left, down, right) of its cell (ci, cj).  All four lanes of a cell
    // return the same Flags.  Equivalent to scan()/true_eye()/compete() of board_ops.h.
    __device__ __forceinline__ Flags eval_cells(int agent, int ci, int cj) const {
        const int dir = lane & 3;
        const int sh = lane & ~3;
        const bool cell_on = ci >= 0 && ci < N && cj >= 0 && cj < N;
        const int cidx = ci * N + cj;
        Flags f;
        f.empty = cell_on && Base::fCol(this->w0[cell_on ? cidx : 0]) == 0;

        const int ni = ci + (dir == 0 ? -1 : (dir == 2 ? 1 : 0));
        const int nj = cj + (dir == 1 ? -1 : (dir == 3 ? 1 : 0));
        const bool nb_on = cell_on && ni >= 0 && ni < N && nj >= 0 && nj < N;
        const uint32_t wn = this->w0[nb_on ? ni * N + nj : 0];
        const int ncol = nb_on ? Base::fCol(wn) : -1;
        const bool nb_stone = ncol > 0;
        const int root = nb_stone ? Base::fA(wn) : -1 - dir;     // distinct negatives never match
        const int hp = (int)(this->w1[nb_stone ? root : 0] & 0xffffu);

        const int r1 = __shfl_xor_sync(kFull, root, 1);
        const int r2 = __shfl_xor_sync(kFull, root, 2);
        const int r3 = __shfl_xor_sync(kFull, root, 3);
        const int mult = 1 + (r1 == root) + (r2 == root) + (r3 == root);
        const bool first = !((r1 == root && (dir ^ 1) < dir) || (r2 == root && (dir ^ 2) < dir) ||
                             (r3 == root && (dir ^ 3) < dir));
        const int left = hp - mult;
        const bool enemy = nb_stone && ncol != agent;
        const bool frnd = nb_stone && ncol == agent;

        const unsigned b_empty = __ballot_sync(kFull, ncol == 0);
        const unsigned b_ecap = __ballot_sync(kFull, enemy && left == 0);
        const unsigned b_eatari = __ballot_sync(kFull, enemy && left == 1);
        const unsigned b_fcap = __ballot_sync(kFull, frnd && left == 0);
        const unsigned b_fsafe = __ballot_sync(kFull, frnd && left > 0);
        int fsum = (frnd && first) ? left : 0;
        fsum += __shfl_xor_sync(kFull, fsum, 1);
        fsum += __shfl_xor_sync(kFull, fsum, 2);
        int fmin = frnd ? hp : 9999;
        fmin = min(fmin, __shfl_xor_sync(kFull, fmin, 1));
        fmin = min(fmin, __shfl_xor_sync(kFull, fmin, 2));

        const int n_empty = __popc((b_empty >> sh) & 15u);
        const bool ecap = ((b_ecap >> sh) & 15u) != 0;
        const bool eatari = ((b_eatari >> sh) & 15u) != 0;
        const bool fcap = ((b_fcap >> sh) & 15u) != 0;
        const bool fsafe = ((b_fsafe >> sh) & 15u) != 0;
        const int tot = n_empty + fsum;
        f.suicide = n_empty == 0 && !ecap && !fsafe;          // Board.cpp:480-524
        f.kill = ecap;                                        // :526-563
        f.survive = fcap && (ecap || tot > fmin);             // :565-626
        f.dying = !ecap && tot == 1;                          // :628-679
        f.chase = eatari && tot > 1;                          // :681-733

        // Board::checkTrueEye, :735-751 — lane `dir` also looks at one diagonal
        const unsigned b_own = __ballot_sync(kFull, !nb_on || ncol == agent);
        const int di = ci + (dir < 2 ? -1 : 1), dj = cj + ((dir & 1) ? 1 : -1);
        const bool d_on = cell_on && di >= 0 && di < N && dj >= 0 && dj < N;
        const int dcol = Base::fCol(this->w0[d_on ? di * N + dj : 0]);
        const unsigned b_diag = __ballot_sync(kFull, d_on && dcol == 3 - agent);
        const int cnt = __popc((b_diag >> sh) & 15u);
        const bool edge = ci == 0 || ci == N - 1 || cj == 0 || cj == N - 1;
        f.eye = ((b_own >> sh) & 15u) == 15u && (edge ? cnt < 1 : cnt < 2);
        f.ko = this->ko && cidx == this->ko_idx && agent == this->ko_col;   // :753-755
        return f;
    }

    // Board::getRandomPieceComplex, Board.cpp:413-478.  Pass 1 (reserve flags) is strided over
    // the lanes (one point per lane per round); the ordered walk that maps rnd to a point
    // follows the empty list and is warp-uniform.
    __device__ int random_piece_complex_w(Rng& rng, int agent) {
        constexpr int NN = N * N;
        const int spaces = NN - this->num_black - this->num_white;
        if (spaces == 0) return -1;
        int reserved_total = 0;
        for (int base = 0; base < NN; base += 32) {
            const int idx = base + lane;
            bool res = false;
            if (idx < NN && Base::fCol(this->w0[idx]) == 0) {
                const typename Base::Nb q = this->scan(agent, idx);
                res = this->true_eye(agent, idx) || Base::is_suicide(q) || this->compete(agent, idx);
                if (res) this->w0[idx] |= (1u << 22);
            }
            reserved_total += __popc(__ballot_sync(kFull, res));
        }
        __syncwarp();
        const int range = spaces - reserved_total;
        int pick = -1;
        if (range != 0) {
            int rnd = (int)(rng.rand15() % (uint32_t)range);
            for (int idx = this->head;;) {
                const uint32_t w = this->w0[idx];
                if (!((w >> 22) & 1u)) {
                    if (rnd == 0) { pick = idx; break; }
                    rnd--;
                }
                idx = Base::fB(w);
            }
        }
        if (reserved_total) {
            for (int idx = lane; idx < NN; idx += 32) this->w0[idx] &= ~(1u << 22);
            __syncwarp();
        }
        return pick;
    }

    // Board::getRandomPiece, Board.cpp:341-411
    __device__ int random_piece_w(Rng& rng, int agent) {
        if (this->ending) {                                     // :348-351
            if (this->num_black + this->num_white <= kEndingStones) this->ending = 0;
            else return random_piece_complex_w(rng, agent);
        }
        const int enemy = 3 - agent;
        const int lm = this->last(enemy);
        if (lm >= 0 && this->col(lm) == enemy) {                // :355-380
            const int er = lm / N, ec = lm - er * N;
            const int cell = lane >> 2;
            const int cc = cell + (cell >= 4);                  // skip the (occupied) centre
            const int ci = er - 1 + cc / 3, cj = ec - 1 + cc % 3;
            const Flags f = eval_cells(agent, ci, cj);
            const bool ok = f.empty && !f.dying && !f.eye && !f.suicide && !f.ko;
            const unsigned K = __ballot_sync(kFull, ok && f.kill);
            const unsigned S = __ballot_sync(kFull, ok && f.survive);
            const unsigned C = __ballot_sync(kFull, ok && f.chase);
            if ((K | S | C) == 0) {
                // no cell can be accepted: the loop runs its 9 iterations, 2 draws each
                rng.skip(18);
            } else {
                for (int k = 0; k < 9; k++) {
                    const int a = (int)(rng.rand15() % 3u);
                    const int b = (int)(rng.rand15() % 3u);
                    const int c9 = a * 3 + b;
                    if (c9 == 4) continue;                      // centre: occupied
                    const int c8 = c9 - (c9 > 4);
                    const unsigned m = K | (k < 6 ? S : 0u) | (k < 4 ? C : 0u);
                    if ((m >> (4 * c8)) & 1u) return (er - 1 + a) * N + (ec - 1 + b);
                }
            }
        }
        for (int count = 1;; count++) {                         // :382-406
            const int row = (int)(rng.rand15() % (uint32_t)N);
            const int cl = (int)(rng.rand15() % (uint32_t)N);
            const int idx = row * N + cl;
            bool bad = this->col(idx) != 0;
            if (!bad) {
                const Flags f = eval_cells(agent, row, cl);
                bad = f.dying || f.eye || f.suicide || f.ko;
            }
            if (count == kEndingThreshold) {                    // even if the 100th draw passed
                this->ending = 1;
                return random_piece_complex_w(rng, agent);
            }
            if (!bad) return idx;
        }
    }

    // Board::countScore(agent) - countScore(enemy), Board.cpp:910-925, strided over the lanes
    __device__ int score_diff(int agent) const {
        constexpr int NN = N * N;
        int d = 0;
        for (int idx = lane; idx < NN; idx += 32) {
            const int i = idx / N, j = idx - i * N;
            const int c = Base::fCol(this->w0[idx]);
            const int owner = c != 0 ? c : Base::fCol(this->w0[j > 0 ? idx - 1 : idx + 1]);
            d += (owner == agent) - (owner == 3 - agent);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
        return d;
    }
};

template <int N>
struct BoardWords { static constexpr int value = ((2 * N * N + 31) / 32) * 32; };

template <int N, int WARPS, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, BLOCKS_PER_SM) playout_warp_kernel(const KernelArgs a) {
    extern __shared__ uint32_t smem[];
    constexpr int NN = N * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    WarpBoard<N> bd;
    bd.w0 = smem + (size_t)warp * BoardWords<N>::value;
    bd.w1 = bd.w0 + NN;
    bd.n_rt = N;
    bd.lane = lane;

    unsigned long long moves = 0, draws = 0;
    unsigned int faults = 0, played = 0;

    for (;;) {
        unsigned long long l = 0;
        if (lane == 0) l = atomicAdd(a.work_counter, 1ull);
        l = __shfl_sync(kFull, l, 0);
        if (l >= (unsigned long long)a.local_units) break;
        const int64_t g = (int64_t)l * a.shard_count + a.shard_rank;
        const int s = (int)(g / a.P);
        const int64_t p = g - (int64_t)s * a.P;
        const Position& pos = a.starts[s];
        const StartDesc d = a.descs[s];

        __syncwarp();
        for (int i = lane; i < NN; i += 32) {                   // Board::clone, Board.cpp:216-254
            bd.w0[i] = pos.w0[i];
            bd.w1[i] = pos.w1[i];
        }
        bd.load_scalars(pos);
        __syncwarp();

        Rng rng;
        rng.seed(playout_seed(a.seed_base, d.position_id, d.cand_idx, (uint64_t)p), a.rng);

        // pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443)
        const int first = d.first, second = 3 - d.first;
        bool go1 = true, go2 = true, fault = false;
        uint32_t mv = 0;
        while (go1 || go2) {
            int r = bd.random_piece_w(rng, first);
            if (r >= 0) { bd.put(first, r); go1 = true; mv++; }
            else { go1 = false; bd.pass(first); }
            r = bd.random_piece_w(rng, second);
            if (r >= 0) { bd.put(second, r); go2 = true; mv++; }
            else { go2 = false; bd.pass(second); }
            if ((int)mv > a.max_moves) { fault = true; break; }
        }
        __syncwarp();
        const int diff = fault ? -32768 : bd.score_diff(d.agent);
        moves += mv;
        draws += rng.draws;
        played++;
        if (fault) faults++;
        if (lane == 0) {
            if (a.diffs) a.diffs[g] = (int16_t)diff;
            if (!fault) {
                const int w = diff - d.avrg_win;                 // AgentGo.cpp:327
                unsigned long long* c = (unsigned long long*)a.counters;
                if (w >= 0) {                                    // sign test of AgentGo.cpp:328
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
