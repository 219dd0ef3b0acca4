// playout_warp.cu — one WARP per playout (sm_100a).
//
// One board per warp in shared memory, so 32 playouts are resident per SM at 19x19 with 64
// registers per thread (the thread-per-playout variant fits 64 boards but pays 32-way
// divergence and has 2 warps per SM to hide latency with).
//
//  * The board is padded to (N+2)x(N+2) with a border of colour 3 (PaddedPosition,
//    kernel_args.h): neighbour and diagonal reads need no bounds checks, and every stored index
//    (links, roots, tails, last moves, ko point) is a padded index.  The reference's linked list
//    of empty points is an append-only array here (kernel_args.h): placing a stone costs nothing.
//  * State updates (put / capture / merge / pass): the 88 % of moves that capture nothing take a
//    lane-parallel path (one lane per direction); a capture makes the reference's sequential
//    order observable and runs WARP-UNIFORM — every lane executes the same scalar code on the
//    same shared-memory words (same address, same value), no divergence, no communication.
//  * The policy's predicates are evaluated LANE-PARALLEL: lane = (cell, direction) with
//    8 cells x 4 directions = 32 lanes (eval_cells).
//  * RANDOM NUMBERS ARE DECOUPLED FROM THEIR CONSUMPTION.  The reference consumes one serial
//    rand() stream per playout; stepping one TinyMT32 in all 32 lanes costs 20 instructions per
//    draw with one useful lane (it was 35 % of the kernel).  Instead each warp keeps a ring of
//    1 024 pre-generated rand() values in shared memory.  A refill produces the next 512: the
//    transition of TinyMT32 is linear over GF(2), so T^16 is applied 32 times as a per-lane
//    nibble-table lookup (one LDG.128) folded by four redux.xor, giving every lane the start
//    state of its own 16-draw segment, which it then generates by itself: ~1.8 instructions
//    per draw.  Draw number p of a playout always lands in ring[p & 1023], so the stream the
//    policy sees is exactly the reference's.
//  * With the draws in shared memory both sampling loops of Board::getRandomPiece
//    (Board.cpp:354-406) become lane-parallel too: they consume two draws per iteration
//    whatever happens inside, and their predicates are pure functions of a state that does not
//    change inside the loop, so lane j examines iteration j.  Neighbourhood phase: the 8 cells
//    around the opponent's last move (the centre is occupied by precondition) are evaluated
//    once into three 8-bit masks (kill / survive / chase, filtered by dying / true eye /
//    suicide / ko); lanes 0..8 test their iteration against them; empty masks (75 % of the
//    moves) just advance the read position by 18.  Rejection phase: 32 (row, col) pairs per
//    round, the empty ones are evaluated in draw order, 8 per eval_cells call; the 100th draw
//    is never accepted (Board.cpp:402-405).
//  * Clone and scoring are strided over the lanes.  Complex mode (Board::getRandomPieceComplex)
//    looks at every empty point, 50 on average: ONE LANE PER POINT (reserved_point, no
//    collectives), and its ordered walk counts 32 entries of the empty-point array per ballot.
//  * 64 registers per thread is what 32 resident warps leave.  Nothing but board state lives in
//    registers across a playout, the generator state is parked in shared memory between refills,
//    and the lane constants are pinned with opaque moves: the compiler otherwise recomputes them
//    inside the move loop.
//
// The code is written for a SMALL INSTRUCTION FOOTPRINT: 32 warps per SM sit at 32 different
// program counters, and the first version of this kernel (every helper force-inlined, both
// movers unrolled: 6 150 SASS instructions, 98 KB) spent 70 % of its stall samples waiting for
// instruction fetch (profiles/r01_warp_v1_*.txt).  Hence: one move loop for both colours, one
// capture site, one complex-mode site, rolled loops.  Since then the kernel is bound by the
// integer ALU pipe and the issue slots (80 % busy), so the remaining lever is instruction count.
#include "kernel_args.h"

namespace agb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kBorder = 3;
constexpr int kAmafRange = 40;      // AMAF_RANGE, GoContest/AgentGo.h:16

// Shared-memory array view.  In the AGB_CHECK_BOUNDS build (agentgo_b200/build.py --debug; the
// pool has no compute-sanitizer) every index is checked and violations are counted in
// g_bounds_violations, which agb_debug_violations() returns; release builds index directly.
#ifdef AGB_CHECK_BOUNDS
__device__ unsigned long long g_bounds_violations;
// coverage counters of the checked build (agb_debug_event): 0 = elist compactions forced by a full
// array, 1 = compactions on entering / during complex mode, 2 = puts through the slow path
__device__ unsigned long long g_debug_events[4];
#define AGB_DEBUG_EVENT(k) do { if (lane == 0) atomicAdd(&g_debug_events[k], 1ull); } while (0)
#else
#define AGB_DEBUG_EVENT(k) do { } while (0)
#endif
#ifdef AGB_CHECK_BOUNDS
template <typename T>
struct SmemArr {
    T* p;
    int n;
    __device__ __forceinline__ T& operator[](int i) const {
        if ((unsigned)i >= (unsigned)n) { atomicAdd(&g_bounds_violations, 1ull); i = 0; }
        return p[i];
    }
    // pointer to the two elements at byte offset `off` (both checked)
    __device__ __forceinline__ T* pair_at(uint32_t off) const {
        if (off / sizeof(T) + 1u >= (unsigned)n) { atomicAdd(&g_bounds_violations, 1ull); off = 0; }
        return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + off);
    }
};
#else
template <typename T>
struct SmemArr {
    T* p;
    __device__ __forceinline__ T& operator[](int i) const { return p[i]; }
    __device__ __forceinline__ T* pair_at(uint32_t off) const { return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + off); }
};
#endif

// device word layout (kernel_args.h): A | 4B<<16 | colour<<30.  A and B are the 16-bit halves of the
// word, so labels and links are rewritten with 16-bit stores; the colour sits in the top two bits so
// that it is one shift away and `empty` / `border` are plain compares on the word.
__device__ __forceinline__ int fA(uint32_t w) { return (int)(w & 0xffffu); }
// B is stored times four (the byte offset of the next stone's word), so a walk along a string
// turns the loaded half word into the next address with one AND
__device__ __forceinline__ int fBh(uint32_t hi) { return (int)((hi >> 2) & 511u); }   // from the high half
__device__ __forceinline__ uint32_t offB(uint32_t hi) { return hi & 0x7fcu; }         // byte offset of B's word
constexpr uint32_t kNilOff = 4u * kNilP;
__device__ __forceinline__ int fCol(uint32_t w) { return (int)(w >> 30); }
__device__ __forceinline__ bool isEmpty(uint32_t w) { return w < 0x40000000u; }
__device__ __forceinline__ bool isBorder(uint32_t w) { return w >= 0xc0000000u; }
// colour 1 or 2: adding a quarter turn moves exactly those two into the upper half of the range
__device__ __forceinline__ bool isStone(uint32_t w) { return (int32_t)(w + 0x40000000u) < 0; }
__device__ __forceinline__ uint32_t pack0(int a, int b, int col) {
    return (uint32_t)a | ((uint32_t)b << 18) | ((uint32_t)col << 30);
}
// high half of a stone's word: colour and next-in-string link
__device__ __forceinline__ uint16_t packHi(int b, int col) { return (uint16_t)(((uint32_t)b << 2) | ((uint32_t)col << 14)); }
// CONVERGENCE.  The scalars of a playout (the move, the position in the rand() stream, walk cursors) hold
// the same value in all 32 lanes, but ptxas cannot see that once they have been through a shuffle or
// shared memory: a branch on them counts as divergent, and then EVERY collective of the kernel gets a
// UMOV + BRA.DIV guard and a WARPSYNC.COLLECTIVE slow path (15 % of the executed instructions, a third
// of the code).  uni() takes the branch on a vote, which is provably uniform.  The sites below are the
// minimal set (found by removing them one at a time while the kernel stayed free of BRA.DIV); together
// with the __syncwarp() that ends the playout loop — it gives the loop a single latch behind the
// lane-0 epilogue — they take the 19x19 kernel from 2 496 to 1 560 SASS instructions
// (tests/test_sass_convergence.py keeps it that way).
__device__ __forceinline__ bool uni(bool p) { return __any_sync(kFull, p) != 0; }
constexpr uint32_t kResFlag = 0x8000u;     // elist entry: point is reserved in the current complex-mode call

template <int N>
struct WarpBoard {
    static constexpr int W = N + 2;
    static constexpr int NP = W * W;
    static constexpr int NN = N * N;
    SmemArr<uint32_t> w0;        // A | B<<16 | colour<<30   (field meaning: kernel_args.h)
    SmemArr<uint16_t> h0;        // the same words as 16-bit halves: h0[2i] = A, h0[2i+1] = B | colour<<14
    SmemArr<uint32_t> w1;        // hp | tail<<16 at root stones
    SmemArr<uint16_t> elist;     // kElistCap entries: the empty list as an append-only array (kernel_args.h)
    SmemArr<uint16_t> cbuf;      // 8 entries: candidates of one rejection round
    int lane;
    int off_nb;          // offset of this lane's orthogonal neighbour (dir = lane & 3: up, left, down, right)
    int off_dg;          // offset of this lane's diagonal
    int off_cell;        // offset of this lane's cell in the 3x3 block around a point (cell = lane >> 2)
    int n_el;            // entries in elist (dead ones included)
    int n_dead;          // >= 0: complex mode lasts, elist holds no stale entries other than n_dead
                         //       zeroed ones (put maintains it); -1: stale entries are validated lazily
    int stones;          // num_black + num_white (only the sum is ever read on this path)
    int ko_key;          // exist_compete ? compete colour<<16 | idx : -1
    int prev;            // last_move[] of the side that is NOT to move, as idx, -1 = none.  The two sides
                         // alternate strictly, so after the first move of a playout this is simply the
                         // previous move (or -1 after a pass, Board.cpp:200-203): one register
                         // instead of the reference's pair indexed by colour.
    int ending;          // game_ending
    // --- random numbers: a ring of pre-generated rand() values (see refill) ---
    SmemArr<uint16_t> ring;      // kRing entries, entry (p & (kRing-1)) holds draw number p of this playout;
                                 // then kMirror more that repeat the first kMirror, so that a read of
                                 // up to kMirror draws ahead of pos needs no wrap-around per lane
    uint4* gstate;               // generator state at draw number gen_end: only a refill needs it, so it
                                 // lives in shared memory, not in four registers of the move loop
    uint32_t pos;                // next draw to be consumed = rand() calls made so far
    uint32_t gen_end;            // draws [pos, gen_end) are in the ring

    static constexpr int kRing = 1024;
    static constexpr int kMirror = 64;              // peek(i) is used with i < 64
    static constexpr int kSeg = kJumpSteps;         // draws generated per lane per refill
    static constexpr int kRefill = 32 * kSeg;       // 512 draws per refill
    // One move consumes at most 18 (neighbourhood) + 200 (rejection) + 1 (complex) draws and the
    // lane-parallel phases never read past that, so this much is made available before each move.
    static constexpr int kLookahead = 224;

    __device__ __forceinline__ void init_lane(int l) {
        lane = l;
        const int dir = l & 3;
        off_nb = dir == 0 ? -W : (dir == 1 ? -1 : (dir == 2 ? W : 1));
        off_dg = (dir < 2 ? -W : W) + ((dir & 1) ? 1 : -1);
        const int cell = l >> 2;
        const int cc = cell + (cell >= 4);                          // skip the centre of the 3x3 block
        off_cell = (cc / 3 - 1) * W + (cc % 3 - 1);
        // Opaque moves: without them the compiler recomputes these lane constants (two divisions
        // by three for off_cell) at every use inside the move loop rather than hold a register.
        asm volatile("mov.s32 %0, %0;" : "+r"(off_cell));
        asm volatile("mov.s32 %0, %0;" : "+r"(off_nb));
        asm volatile("mov.s32 %0, %0;" : "+r"(off_dg));
        asm volatile("mov.s32 %0, %0;" : "+r"(lane));
    }

    // rand() number pos + i (i < kLookahead), without consuming it
    __device__ __forceinline__ uint32_t peek(uint32_t i) const { return ring[(pos & (kRing - 1)) + i]; }
    // v % M for a 15-bit value and a small constant M: one IMAD.HI and one IMAD
    // (q = floor(v * ceil(2^32 / M) / 2^32) is exact for v < 2^15)
    template <uint32_t M>
    __device__ __forceinline__ static uint32_t mod_small(uint32_t v) {
        const uint32_t q = __umulhi(v, (uint32_t)((0x100000000ull + M - 1) / M));
        return v - q * M;
    }

    // Generate the next kRefill draws of this playout's TinyMT32 stream into the ring.
    // All 32 lanes used to step the one generator redundantly (20 instructions per draw, a
    // quarter of the kernel); here the serial part is only the 32 segment-start states: the
    // transition is linear over GF(2), so T^16 is applied as a per-lane nibble-table lookup
    // (LDG.128) folded by four redux.xor, and then every lane generates its own 16 draws.
    __device__ __forceinline__ void refill(const RngParams& prm, const JumpTable* __restrict__ jt) {
        const bool lo16 = lane < 16, lo8 = (lane & 8) == 0;
        const uint32_t sh = (lane & 7) * 4;
        const uint4* __restrict__ tab = jt->e;
        const uint32_t row = (uint32_t)lane * 16u;                  // 32-bit index: one IMAD.WIDE per lookup
        // The half of the ring about to be written is free (at most kLookahead draws are pending
        // and they sit in the other half), so the segment-start states are parked there: state l
        // in the first 16 bytes of lane l's own 32-byte segment, which that lane reads back before
        // it writes its draws over it.  (A register capture costs a compare and four selects per
        // step in all lanes; this is one store.)
        uint4* park = reinterpret_cast<uint4*>(ring.p + (gen_end & (kRing - 1)));
        uint32_t g0, g1, g2, g3;
        {
            const uint4 g = *gstate;
            g0 = g.x; g1 = g.y; g2 = g.z; g3 = g.w;
        }
#pragma unroll 4
        for (int l = 0; l < 32; l++) {
            park[2 * l] = make_uint4(g0, g1, g2, g3);
            const uint32_t word = lo16 ? (lo8 ? g0 : g1) : (lo8 ? g2 : g3);
            const uint4 t = __ldg(tab + (row | ((word >> sh) & 15u)));
            g0 = __reduce_xor_sync(kFull, t.x);
            g1 = __reduce_xor_sync(kFull, t.y);
            g2 = __reduce_xor_sync(kFull, t.z);
            g3 = __reduce_xor_sync(kFull, t.w);
        }
        TinyMT mine;
        {
            const uint4 st = park[2 * lane];
            mine.s0 = st.x; mine.s1 = st.y; mine.s2 = st.z; mine.s3 = st.w;
        }
        const uint32_t base = (gen_end + (uint32_t)lane * kSeg) & (kRing - 1);   // segments do not wrap
        const bool mirrored = base < kMirror;
#pragma unroll 2
        for (int j = 0; j < kSeg; j++) {
            mine.next_state(prm.mat1, prm.mat2);
            const uint16_t v = (uint16_t)(mine.temper(prm.tmat) >> 17);
            ring[base + j] = v;
            if (mirrored) ring[kRing + base + j] = v;
        }
        *gstate = make_uint4(g0, g1, g2, g3);
        gen_end += kRefill;
        __syncwarp();
    }

    // Lane-parallel evaluation of up to 8 points at once: lane = 4*cell + dir evaluates
    // neighbour `dir` (up, left, down, right) of its cell `cidx` (padded index; a border cell is
    // never `empty`); the four lanes of a cell hold the same result.  Restates
    // checkSuicide/Kill/Survive/Dying/Chase/TrueEye/Compete (Board.cpp:480-755).
    //   bad     = dying | true eye | suicide | ko   (the filter of both sampling phases)
    //   reserve = true eye | suicide | ko           (complex mode: no checkDying, :425-427)
    // Force-inlined at two call sites.  MODE selects what is computed, because the warp-sync
    // shuffles and ballots have side effects the compiler will not remove on its own:
    //   kEvalAll     neighbourhood phase: everything
    //   kEvalBad     rejection phase: `bad` only (no survive / chase: no fmin, fcap, eatari)
    // (Complex mode has many points to look at and uses one lane per point: reserved_point.)
    struct Eval {
        bool empty, kill, survive, chase, bad, reserve;
        bool live;      // kEvalAll only, warp-uniform: false = the early exit was taken, all masks are empty
        bool easy0;     // kEvalBad only, warp-uniform: cell 0 was accepted by the early test
    };
    enum { kEvalAll = 0, kEvalBad = 1 };
    template <int MODE>
    __device__ __forceinline__ Eval eval_cells(int agent, int cidx) const {
        const unsigned nib = 15u << (lane & ~3);                    // this cell's four lanes
        // kEvalBad is only ever handed points of the board (candidates of the rejection phase; cbuf
        // starts out holding one), so it neither loads the cell nor guards the neighbour offsets
        const uint32_t wc = MODE == kEvalBad ? 0u : w0[cidx];
        const bool cell_on = MODE == kEvalBad || !isBorder(wc);
        const uint32_t wn = w0[cell_on ? cidx + off_nb : 0];        // w0[0] is a border cell
        const int ncol = fCol(wn);
        if (MODE == kEvalBad) {
            // rejection phase: a point with two or more empty neighbours that is not the ko point
            // cannot be dying (its string keeps >= 2 liberties), an eye or a suicide.  When the
            // FIRST candidate in draw order (cell 0) is of that kind it is accepted right away,
            // before any string / liberty is looked at.
            const int ne = __popc(__ballot_sync(kFull, ncol == 0) & nib);
            const bool easy = ne >= 2 && ko_key != ((agent << 16) | cidx);
            if (__ballot_sync(kFull, easy) & 1u) {
                Eval z;
                z.empty = true; z.kill = z.survive = z.chase = z.reserve = false;
                z.bad = lane >= 4;
                z.easy0 = true;
                return z;
            }
        }
        const bool nb_stone = isStone(wn);
        const int root = nb_stone ? fA(wn) : -1 - (lane & 3);       // distinct negatives per direction
        const int hp = (int)(w1[nb_stone ? root : 0] & 0xffffu);

        // strings seen from several directions count once, with their multiplicity as minus_hp
        // (Board.cpp:506-515): lanes of this cell holding the same root
        const unsigned same = __match_any_sync(kFull, root) & nib;
        const int left = hp - __popc(same);
        const bool first = (same & ((1u << lane) - 1u)) == 0;
        const bool enemy = ncol == 3 - agent;
        const bool frnd = ncol == agent;

        if (MODE == kEvalAll) {
            // kill, survive and chase all need a string next to an EMPTY cell that would be left
            // with 0 or 1 liberties (ecap / fcap: 0, eatari: 1); without one the three masks of
            // the neighbourhood phase are empty and nothing else matters (~70 % of the calls)
            if (__ballot_sync(kFull, isEmpty(wc) && nb_stone && left <= 1) == 0) {
                Eval z;
                z.empty = z.kill = z.survive = z.chase = z.bad = z.reserve = z.live = z.easy0 = false;
                return z;
            }
        }

        const int n_empty = __popc(__ballot_sync(kFull, ncol == 0) & nib);
        const bool ecap = (__ballot_sync(kFull, enemy && left == 0) & nib) != 0;
        const bool fsafe = (__ballot_sync(kFull, frnd && left > 0) & nib) != 0;
        const bool suicide = n_empty == 0 && !ecap && !fsafe;       // Board.cpp:480-524
        bool dying = false, survive = false, chase = false;
        {
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
        e.empty = isEmpty(wc);
        e.kill = ecap;                                              // :526-563
        e.survive = survive;
        e.chase = chase;
        e.reserve = eye || suicide || ko;
        e.bad = e.reserve || dying;
        e.live = true;
        e.easy0 = false;
        return e;
    }

    // Drop the dead entries of elist, keeping the order.  Entry k naming cell c is alive iff c is
    // empty and A(c) == k (a zeroed entry names the border cell 0).  In place: a round reads 32
    // entries and then writes at or below where it read.
    __device__ __forceinline__ void compact() {
        int out = 0;
#pragma unroll 1
        for (int base = 0; base < n_el; base += 32) {
            const int k = base + lane;
            const uint32_t c = k < n_el ? ((uint32_t)elist[k] & 0x7fffu) : 0u;
            const uint32_t w = w0[(int)c];
            const bool live = isEmpty(w) && fA(w) == k;
            const unsigned m = __ballot_sync(kFull, live);
            __syncwarp();
            if (live) {
                const int o = out + __popc(m & ((1u << lane) - 1u));
                elist[o] = (uint16_t)c;
                h0[2 * (int)c] = (uint16_t)o;
            }
            out += __popc(m);
        }
        n_el = out;
        if (n_dead > 0) n_dead = 0;
        __syncwarp();
    }

    // true eye | suicide | ko of the empty point idx for `agent` (checkTrueEye :735-751,
    // checkSuicide :480-524, checkCompete :753-755; the reserve test of complex mode, :425-427),
    // computed by ONE lane.  Complex mode looks at every empty point (50 on average), so a lane
    // per point keeps all 32 lanes busy where eval_cells' four lanes per point would take four
    // times as many rounds; nothing here is a warp collective.
    __device__ __forceinline__ bool reserved_point(int agent, int idx) const {
        const int offs[4] = {-W, -1, W, 1};
        int root[4], hp[4], col[4];
        int n_empty = 0;
        bool edge = false, own4 = true;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const uint32_t w = w0[idx + offs[d]];
            col[d] = fCol(w);
            n_empty += col[d] == 0;
            edge |= col[d] == kBorder;
            own4 &= col[d] == kBorder || col[d] == agent;
            const bool stone = isStone(w);
            root[d] = stone ? fA(w) : -1 - d;                       // distinct negatives per direction
            hp[d] = (int)(w1[stone ? root[d] : 0] & 0xffffu);
        }
        // a string seen from several directions counts once, less its multiplicity (:506-515)
        bool ecap = false, fsafe = false;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            int m = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) m += root[e] == root[d];
            ecap |= col[d] == 3 - agent && hp[d] == m;
            fsafe |= col[d] == agent && hp[d] > m;
        }
        const bool suicide = n_empty == 0 && !ecap && !fsafe;
        const int enemy = 3 - agent;
        const int cnt = (fCol(w0[idx - W - 1]) == enemy) + (fCol(w0[idx - W + 1]) == enemy) +
                        (fCol(w0[idx + W - 1]) == enemy) + (fCol(w0[idx + W + 1]) == enemy);
        const bool eye = own4 && cnt < (edge ? 1 : 2);
        return eye || suicide || ko_key == ((agent << 16) | idx);
    }

    // Board::getRandomPieceComplex, Board.cpp:413-478.  Pass 1 (reserve flags, :419-433): the
    // entries of elist are evaluated 32 per round, one lane each, and flagged in place.  Pass 2
    // (:445-470, the rnd-th unreserved point in list order) counts 32 entries per round from the
    // end of the array.
    __device__ __forceinline__ int random_piece_complex(int agent) {
        const int spaces = NN - stones;
        if (spaces == 0) return -1;
        // Complex mode is sticky (game_ending, :348-351): the array is made dense once and then
        // kept current by put (zeroes the entry) and kill_string (appends).
        if (n_dead < 0 || n_dead >= 8) { AGB_DEBUG_EVENT(1); n_dead = 0; compact(); }
        int reserved_total = 0;
#pragma unroll 1
        for (int base = 0; base < n_el; base += 32) {
            const int k = base + lane;
            const int ent = k < n_el ? (int)((uint32_t)elist[k] & 0x7fffu) : 0;    // 0: a dead entry
            // dead entries look at some point of the board; their result is dropped
            const bool res = reserved_point(agent, ent ? ent : W + 1) && ent != 0;
            if (ent) elist[k] = (uint16_t)((uint32_t)ent | (res ? kResFlag : 0u));  // addReserve, :757-761
            reserved_total += __popc(__ballot_sync(kFull, res));
        }
        __syncwarp();
        const int range = spaces - reserved_total;
        int pick = -1;
        if (range != 0) {
            int rnd = (int)(peek(0) % (uint32_t)range);
            pos++;
#pragma unroll 1
            for (int top = n_el - 1; top >= 0; top -= 32) {
                const int k = top - lane;                           // lane order = list order
                const uint32_t ent = k >= 0 ? (uint32_t)elist[k] : 0u;
                const bool ok = ent - 1u < 0x7fffu;                 // alive and not reserved
                const unsigned m = __ballot_sync(kFull, ok);
                const int c = __popc(m);
                if (uni(rnd < c)) {
                    const bool hit = ok && __popc(m & ((1u << lane) - 1u)) == rnd;
                    pick = (int)__shfl_sync(kFull, ent, __ffs(__ballot_sync(kFull, hit)) - 1);
                    break;
                }
                rnd -= c;
            }
        }
        return pick;
    }

    // Board::getRandomPiece, Board.cpp:341-411; returns the (padded) point or -1 (pass).
    // Both sampling loops of the reference consume two rand() per iteration whatever happens
    // inside, and the predicates are pure functions of a state that does not change inside the
    // loops, so lane j can examine iteration j directly from the ring: the first accepted
    // iteration wins and exactly its draws (and those before it) are consumed.
    __device__ __forceinline__ int random_piece(int agent) {
        bool complex_mode = false;
        if (uni(ending != 0)) {                                     // :348-351
            if (stones <= kEndingStones) ending = 0;
            else complex_mode = true;
        }
        if (!complex_mode) {
            n_dead = -1;                                            // put stops maintaining elist
            // opponent's last move.  The reference also requires the stone to be still there
            // (:356); inside a playout it always is (it was placed one move ago), and for the first
            // move of a playout the kernel settles that before the loop.
            const int lm = prev;
            if (uni(lm >= 0)) {                                     // :355-380
                const Eval f = eval_cells<kEvalAll>(agent, lm + off_cell);
                unsigned K = 0, S = 0, C = 0;
                if (f.live) {
                    const bool ok = f.empty && !f.bad;
                    K = __ballot_sync(kFull, ok && f.kill);
                    S = __ballot_sync(kFull, ok && f.survive);
                    C = __ballot_sync(kFull, ok && f.chase);
                }
                if ((K | S | C) != 0) {
                    // lane k = iteration k of `for (k = 0; k < 9; k++)`
                    const uint32_t a = mod_small<3>(peek(2 * lane));
                    const uint32_t b = mod_small<3>(peek(2 * lane + 1));
                    const uint32_t c9 = a * 3u + b;
                    const uint32_t c8 = c9 - (c9 > 4u);
                    const unsigned m = K | (lane < 6 ? S : 0u) | (lane < 4 ? C : 0u);
                    const bool hit = lane < 9 && c9 != 4u && ((m >> (4 * c8)) & 1u);   // centre is occupied
                    const unsigned hm = __ballot_sync(kFull, hit);
                    if (hm) {
                        const int k = __ffs(hm) - 1;
                        pos += 2 * (k + 1);
                        return lm + (int)__shfl_sync(kFull, a * W + b, k) - (W + 1);
                    }
                }
                pos += 18;      // nine iterations, two draws each, none accepted
            }
            // rejection loop, :382-406: up to 32 pairs are examined in one round.  lim = the pairs
            // that may still be accepted: the 100th draw is consumed but never accepted (:402-405)
#pragma unroll 1
            for (int lim = kEndingThreshold - 1; lim >= 0;) {
                const uint32_t row = mod_small<N>(peek(2 * lane));
                const uint32_t cl = mod_small<N>(peek(2 * lane + 1));
                const int idx = (int)(row * W + cl + (W + 1));
                const bool can = lane < lim && isEmpty(w0[idx]);
                unsigned cand = __ballot_sync(kFull, can);
                // evaluate the empty candidates in draw order, 8 per round
#pragma unroll 1
                while (cand) {
                    const int rank = __popc(cand & ((1u << lane) - 1u));
                    const bool mine = ((cand >> lane) & 1u) && rank < 8;
                    if (mine) cbuf[rank] = (uint16_t)(idx | (lane << 9));
                    __syncwarp();
                    const int ncand = min(8, __popc(cand));
                    const int cell = lane >> 2;
                    const uint32_t entry = cbuf[cell];          // cells >= ncand: a stale or the initial point, unused
                    const Eval e = eval_cells<kEvalBad>(agent, (int)(entry & 511u));
                    if (e.easy0) {                              // the first candidate in draw order, held by lane 0
                        const uint32_t win = __shfl_sync(kFull, entry, 0);
                        pos += 2 * ((win >> 9) + 1);
                        return (int)(win & 511u);
                    }
                    const unsigned good = __ballot_sync(kFull, cell < ncand && !e.bad);
                    if (good) {
                        const uint32_t win = __shfl_sync(kFull, entry, __ffs(good) - 1);
                        pos += 2 * ((win >> 9) + 1);
                        return (int)(win & 511u);
                    }
                    cand &= ~__ballot_sync(kFull, mine);
                    __syncwarp();
                }
                const int B = min(32, lim + 1);                         // pairs consumed by this round
                pos += 2 * B;
                lim -= B;
            }
            ending = 1;
        }
        return random_piece_complex(agent);
    }

    // Board::killSetNode, Board.cpp:293-339.  The walk is warp-uniform; the four neighbours of each
    // captured stone are looked at by lanes 0..3 (increments commute, so they are atomics).
    __device__ __forceinline__ void kill_string(int root) {
        const bool single = (int)(w1[root] >> 16) == root;
        const int agent = fCol(w0[root]);                           // colour of the captured string
#pragma unroll 1
        for (int k = root; uni(k != kNilP);) {
            const int nx = fBh(h0[2 * k + 1]);
            stones--;
            if (n_el == kElistCap) { AGB_DEBUG_EVENT(0); compact(); }   // stale entries are dropped
            elist[n_el] = (uint16_t)k;                              // push at the list head, :311-321
            w0[k] = (uint32_t)n_el;                                 // empty, A = slot
            n_el++;
            const uint32_t v = w0[k + off_nb];                      // one liberty back, :323-326
            if (lane < 4 && fCol(v) == 3 - agent) atomicAdd(&w1[fA(v)], 1u);
            if (single) ko_key = (agent << 16) | k;                 // :328-333
            k = nx;
        }
        __syncwarp();
    }

    // Board::put, Board.cpp:68-177 (warp-uniform; the point is empty).  Nothing is unlinked: the
    // entry of the point in elist dies by itself (kernel_args.h); only while complex mode lasts it
    // is zeroed, so that the array stays free of stale entries.
    __device__ __forceinline__ void put(int agent, int idx) {
        stones++;
        ko_key = -1;
        if (n_dead >= 0) {
            elist[h0[2 * idx]] = 0;
            n_dead++;
            __syncwarp();       // (also keeps this rare block a branch instead of five predicated instructions)
        }
        // lane d (mod 4) looks at neighbour d (up, left, down, right): colour, string, liberties
        const int dl = lane & 3;
        const unsigned nib = 15u << (lane & ~3);
        const int nb_l = idx + off_nb;
        const uint32_t v_l = w0[nb_l];
        const int col_l = fCol(v_l);
        const bool st_l = isStone(v_l);
        const int root_l = st_l ? fA(v_l) : -1 - dl;
        const uint32_t h_l = w1[st_l ? root_l : 0];
        const unsigned same = __match_any_sync(kFull, root_l) & nib;    // directions seeing the same string
        const int mult = __popc(same);
        const bool first = (same & ((1u << lane) - 1u)) == 0;
        const bool enemy_l = col_l == 3 - agent, frnd_l = col_l == agent;
        const int left_l = (int)(h_l & 0xffffu) - mult;
        const int hp_self = __popc(__ballot_sync(kFull, col_l == 0) & 15u);   // own liberties, :123-127
        const unsigned captures = __ballot_sync(kFull, enemy_l && left_l == 0) & 15u;
        const unsigned fr_mask = __ballot_sync(kFull, frnd_l && first) & 15u;  // distinct own strings

        // A move of the policy is never a suicide (checkSuicide, i.e. no capture, no liberty of its
        // own and no own string with a liberty to spare, is filtered by both sampling phases and by
        // complex mode), so "no capture" alone selects the fast path; the slow path below would
        // also handle a self-capture, like the reference does.
#ifdef AGB_CHECK_BOUNDS
        {   // the invariant relied on above, checked: no capture => the new string has a liberty
            int t = (frnd_l && first) ? left_l : 0;
            t += __shfl_xor_sync(kFull, t, 1);
            t += __shfl_xor_sync(kFull, t, 2);
            if (captures == 0 && t + hp_self == 0 && lane == 0) atomicAdd(&g_bounds_violations, 1ull);
            if (captures != 0) AGB_DEBUG_EVENT(2);
        }
#endif
        if (captures == 0) {
            // FAST PATH (no capture, ~88 % of the moves): everything the four sequential neighbour
            // steps of :129-167 do commutes except the unions, whose net effect is known in closed
            // form.  With F1..Fk the distinct own strings in direction order, the reference ends
            // with root Fk, piece order Fk, ..., F1, new stone (each union appends the mover's
            // current string to the neighbour's, SetNode.cpp:78-82), hp = sum(hp_i - mult_i) + own.
            if (lane < 4 && enemy_l && first) w1[root_l] = h_l - (uint32_t)mult;   // sn->hp -= 1 per adjacency
            if (fr_mask == 0) {
                w0[idx] = pack0(idx, kNilP, agent);                 // SetNode::init, SetNode.cpp:31-44
                w1[idx] = (uint32_t)hp_self | ((uint32_t)idx << 16);
            } else if ((fr_mask & (fr_mask - 1u)) == 0) {
                // one own string (most merges): the new stone is appended to it, nothing is relabelled
                const int top = 31 - __clz(fr_mask);
                const int fk = __shfl_sync(kFull, root_l, top);
                // tail of that string and its liberties after the move (its own, less the
                // adjacencies the stone takes, plus the stone's)
                const uint32_t ht = __shfl_sync(kFull, (h_l & 0xffff0000u) | (uint32_t)(left_l + hp_self), top);
                h0[2 * (int)(ht >> 16) + 1] = packHi(idx, agent);
                w0[idx] = pack0(fk, kNilP, agent);
                w1[fk] = (ht & 0xffffu) | ((uint32_t)idx << 16);
            } else {
                int hp_total = (frnd_l && first) ? left_l : 0;
                hp_total += __shfl_xor_sync(kFull, hp_total, 1);
                hp_total += __shfl_xor_sync(kFull, hp_total, 2);
                hp_total += hp_self;
                const int fk = __shfl_sync(kFull, root_l, 31 - __clz(fr_mask));
                // lane of F_i links its tail to the head of F_(i-1), the lane of F_1 to the new stone
                const unsigned lower = fr_mask & ((1u << dl) - 1u);
                const int prev = __shfl_sync(kFull, root_l, lower ? 31 - __clz(lower) : 0);
                const int target = lower ? prev : idx;
                const int tail_l = (int)(h_l >> 16);
                const bool own_l = lane < 4 && frnd_l && first;
                if (own_l) h0[2 * tail_l + 1] = packHi(target, agent);
                w0[idx] = pack0(fk, kNilP, agent);
                w1[fk] = (uint32_t)hp_total | ((uint32_t)idx << 16);
                // strings F_1..F_(k-1) are relabelled to Fk (:284-289)
                const unsigned others = fr_mask & ~(1u << (31 - __clz(fr_mask)));
                if (!(others & (others - 1u))) {
                    // the usual case of a merge: one string, walked warp-uniformly
                    const int d1 = __ffs(others) - 1;
                    const uint32_t t1 = 4u * (__shfl_sync(kFull, h_l, d1) >> 16);
#pragma unroll 1
                    for (uint32_t off = 4u * (uint32_t)__shfl_sync(kFull, root_l, d1);;) {
                        uint16_t* c = h0.pair_at(off);              // both halves of the stone's word
                        c[0] = (uint16_t)fk;
                        if (off == t1) break;
                        off = offB(c[1]);
                    }
                } else {
                    // two or three strings: each is walked by its own lane
                    bool walk = own_l && root_l != fk;
                    uint32_t off = walk ? 4u * (uint32_t)root_l : 0u;
#pragma unroll 1
                    while (__ballot_sync(kFull, walk)) {
                        if (walk) {
                            uint16_t* c = h0.pair_at(off);
                            c[0] = (uint16_t)fk;
                            walk = off != 4u * (uint32_t)tail_l;
                            off = offB(c[1]);
                        }
                    }
                }
            }
            __syncwarp();
        } else {
        // SLOW PATH (a capture happens): the reference's sequential order matters
        w0[idx] = pack0(idx, kNilP, agent);                         // SetNode::init, SetNode.cpp:31-44
        w1[idx] = (uint32_t)idx << 16;
        unsigned todo = (__ballot_sync(kFull, st_l) & 15u) | 16u;
        // up, left, down, right (:129-164), then the stone's own string (:165-167) as step 4
#pragma unroll 1
        while (todo) {
            const int d = __ffs(todo) - 1;
            todo &= todo - 1;
            int victim = -1;
            if (d < 4) {
                const int nb = __shfl_sync(kFull, nb_l, d);
                const uint32_t v = w0[nb];
                if (uni(isEmpty(v))) continue;                    // captured earlier in this put
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
                        h0[2 * t1 + 1] = packHi(s2, agent);
#pragma unroll 1
                        for (uint32_t off = 4u * (uint32_t)s2; off != kNilOff;) {
                            uint16_t* c = h0.pair_at(off);
                            c[0] = (uint16_t)sn;
                            off = offB(c[1]);
                        }
                    }
                }
            } else {
                const int self = fA(w0[idx]);
                const uint32_t h = w1[self] + (uint32_t)hp_self;
                w1[self] = h;
                if ((h & 0xffffu) == 0) victim = self;
            }
            if (uni(victim >= 0)) kill_string(victim);
        }
        }
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

// per-warp shared memory in words: w0[NP], w1[NP], elist[kElistCap] and cand[8] (16 bit), the
// generator state (4 words), rounded to 32 words, then the ring of kRing 16-bit draws
template <int N>
struct BoardWords {
    static constexpr int NP = (N + 2) * (N + 2);
    static constexpr int gstate = ((2 * NP + kElistCap / 2 + 4 + 3) / 4) * 4;     // 4 words, 16-byte aligned
    static constexpr int board = ((gstate + 4 + 31) / 32) * 32;
    static constexpr int value = board + (WarpBoard<N>::kRing + WarpBoard<N>::kMirror) / 2;
    static constexpr int kAmafLog = 80;                     // first plays with cstep <= 40: at most 78
    static constexpr int value_amaf = value + kAmafLog / 2;
};

// AMAF = true additionally records the first-play marks of MyWorker::work (AgentGo.cpp:289-290,
// 304,319) and accumulates KernelArgs::amaf (:329-338); candidate mode only.
template <int N, int WARPS, int BLOCKS_PER_SM, bool AMAF>
__global__ void __launch_bounds__(WARPS * 32, BLOCKS_PER_SM) playout_warp_kernel(const KernelArgs a) {
    extern __shared__ uint32_t smem[];
    typedef WarpBoard<N> WB;
    constexpr int kWarpWords = AMAF ? BoardWords<N>::value_amaf : BoardWords<N>::value;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    WB bd;
    uint32_t* base = smem + (size_t)warp * kWarpWords;
    uint16_t* amaf_log = reinterpret_cast<uint16_t*>(base + BoardWords<N>::value);   // AMAF only
    bd.w0.p = base;
    bd.w1.p = base + WB::NP;
    bd.h0.p = reinterpret_cast<uint16_t*>(base);
    bd.elist.p = reinterpret_cast<uint16_t*>(base + 2 * WB::NP);
    bd.cbuf.p = bd.elist.p + kElistCap;
    bd.gstate = reinterpret_cast<uint4*>(base + BoardWords<N>::gstate);
    bd.ring.p = reinterpret_cast<uint16_t*>(base + BoardWords<N>::board);
#ifdef AGB_CHECK_BOUNDS
    bd.w0.n = WB::NP; bd.w1.n = WB::NP; bd.h0.n = 2 * WB::NP; bd.elist.n = kElistCap; bd.cbuf.n = 8; bd.ring.n = WB::kRing + WB::kMirror;
#endif
    bd.init_lane(lane);

    // Nothing but the board state stays in registers across a playout: the statistics go out
    // with one RED each per playout and the start descriptor is read again at the end (64
    // registers per thread is what 32 resident warps leave; every register held here makes the
    // compiler recompute lane constants inside the move loop instead).
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
        const int first = a.descs[s].first;

        __syncwarp();
#pragma unroll 1
        for (int i = lane; i < WB::NP; i += 32) {                   // Board::clone, Board.cpp:216-254
            bd.w0[i] = pos.w0[i];
            bd.w1[i] = pos.w1[i];
        }
        {
            uint32_t* ew = base + 2 * WB::NP;                       // elist, two entries per word
#pragma unroll 1
            for (int i = lane; i < (WB::NN + 1) / 2; i += 32) ew[i] = pos.elist_w[i];
        }
        bd.n_el = pos.n_el;
        bd.stones = pos.stones;
        bd.ko_key = pos.ko_key_col > 0 ? ((pos.ko_key_col << 16) | pos.ko_idx) : -1;
        bd.prev = pos.last[2 - first];                              // last_move[enemy of the first mover]
        bd.ending = pos.ending;
        if (lane < 8) bd.cbuf[lane] = (uint16_t)(WB::W + 1);        // any point of the board (see eval_cells)
        __syncwarp();
        // Board.cpp:356 for the first move: the opponent's last stone may have been captured since
        if (bd.prev >= 0 && fCol(bd.w0[bd.prev]) != 3 - first) bd.prev = -1;
        bd.pos = 0;
        bd.gen_end = 0;
        bd.n_dead = -1;
        __syncwarp();
        {
            // tinymt32_init (RFC 8682)
            const StartDesc d0 = a.descs[s];
            uint32_t st[4] = {playout_seed(a.seed_base, d0.position_id, d0.cand_idx, (uint64_t)p),
                              a.rng.mat1, a.rng.mat2, a.rng.tmat};
#pragma unroll
            for (int i = 1; i < 8; i++)
                st[i & 3] ^= (uint32_t)i + 1812433253u * (st[(i - 1) & 3] ^ (st[(i - 1) & 3] >> 30));
            if ((st[0] & 0x7fffffffu) == 0 && st[1] == 0 && st[2] == 0 && st[3] == 0) {
                st[0] = 'T'; st[1] = 'I'; st[2] = 'N'; st[3] = 'Y';
            }
            TinyMT t;
            t.s0 = st[0]; t.s1 = st[1]; t.s2 = st[2]; t.s3 = st[3];
#pragma unroll 1
            for (int i = 0; i < 8; i++) t.next_state(a.rng.mat1, a.rng.mat2);
            *bd.gstate = make_uint4(t.s0, t.s1, t.s2, t.s3);
        }

        // pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443), one move
        // per iteration: the loop ends after a pair in which neither side moved.
        int colour = first;
        const int amaf_agent = AMAF ? (int)a.descs[s].agent : 0;
        bool fault = false;
        uint32_t mv = 0;
        int n_log = 0, cstep = 2;       // cstep is 2 in the first pair (:293,:297)
#pragma unroll 1
        for (;;) {
            while (bd.gen_end - bd.pos < (uint32_t)WB::kLookahead) bd.refill(a.rng, a.jump);
            const int r = bd.random_piece(colour);
            if (uni(r >= 0)) {
                bd.put(colour, r);
                mv++;
                if (AMAF) {
                    // mark[r] / mark2[r] = cstep of the FIRST play of that side at r (:304,:319);
                    // only cstep <= AMAF_RANGE is ever used
                    if ((cstep <= kAmafRange)) {
                        const uint32_t key = (uint32_t)r | (colour == amaf_agent ? 512u : 0u);
                        bool dup = false;
                        for (int b0 = 0; uni(b0 < n_log); b0 += 32) {
                            const uint32_t e = b0 + lane < n_log ? amaf_log[b0 + lane] : 0xffffu;
                            dup |= __ballot_sync(kFull, (e & 1023u) == key) != 0;
                        }
                        if ((!dup)) { amaf_log[n_log] = (uint16_t)(key | ((uint32_t)cstep << 10)); n_log++; }
                    }
                }
            } else {                                                // Board::pass, Board.cpp:200-203
                bd.ko_key = -1;
            }
            // the move that closes a pair (the second mover's) ends the playout when neither side
            // moved: the first mover's move is what prev still holds, -1 for a pass
            const int before = bd.prev;
            bd.prev = r;                                            // Board.cpp:171 / :202
            if (colour != first) {
                if (uni((before & r) < 0)) break;
                if ((int)mv > a.max_moves) { fault = true; break; } // fault detector, once per pair
                if (AMAF) cstep++;
            }
            colour = 3 - colour;
        }
        __syncwarp();
        const StartDesc d = a.descs[s];
        const int diff = fault ? -32768 : bd.score_diff(d.agent);
        if (lane == 0) {
            atomicAdd(&a.stats[kStatMoves], (unsigned long long)mv);
            atomicAdd(&a.stats[kStatDraws], (unsigned long long)bd.pos);
            atomicAdd(&a.stats[kStatPlayouts], 1ull);
            if (fault) atomicAdd(&a.stats[kStatFaults], 1ull);
        }
        if (AMAF && uni(!fault)) {
            // :329-338 — point q counts for the agent table when the agent's first play there has
            // cstep <= 40, else for the enemy table when the enemy's has (the `else if`)
            __syncwarp();
            const int w = diff - (a.avrg_dev ? __ldg(a.avrg_dev) : d.avrg_win);
            uint32_t agent_bits = 0;                // bit (q & 31) of lane (q >> 5): agent marked q
            for (int k = 0; uni(k < n_log); k++) {
                const uint32_t e = amaf_log[k];
                agent_bits |= ((e & 512u) && lane == (int)((e & 511u) >> 5)) ? 1u << (e & 31u) : 0u;
            }
            for (int b0 = 0; uni(b0 < n_log); b0 += 32) {
                const bool valid = b0 + lane < n_log;
                const uint32_t e = valid ? amaf_log[b0 + lane] : 0u;
                const uint32_t q = e & 511u;
                const bool is_agent = (e & 512u) != 0;
                const bool agent_marked = (__shfl_sync(kFull, agent_bits, (int)(q >> 5)) >> (q & 31u)) & 1u;
                if (valid && w != 0 && (is_agent || !agent_marked)) {
                    const uint32_t row = q / WB::W - 1, col = q % WB::W - 1;
                    const uint32_t slot = ((is_agent ? 0u : 2u) + (w < 0 ? 1u : 0u)) * (kAmafRange + 1) + (e >> 10);
                    atomicAdd((unsigned long long*)&a.amaf[((size_t)d.amaf_class * (4 * (kAmafRange + 1)) + slot) * WB::NN + row * N + col],
                              (unsigned long long)(long long)w);
                }
                __syncwarp();   // one latch (see CONVERGENCE)
            }
        }
        if (lane == 0) {
            if (a.diffs) a.diffs[g] = (int16_t)diff;
            if (!fault) {
                const int w = diff - (a.avrg_dev ? __ldg(a.avrg_dev) : d.avrg_win);       // AgentGo.cpp:327
                unsigned long long* c = (unsigned long long*)a.counters;
                if (w >= 0) {                                        // sign test of AgentGo.cpp:328
                    atomicAdd(&c[s], 1ull);
                    atomicAdd(&c[a.counter_stride + s], (unsigned long long)w);
                } else {
                    atomicAdd(&c[2 * a.counter_stride + s], (unsigned long long)(long long)w);
                }
            }
        }
        __syncwarp();       // one latch for the playout loop (see CONVERGENCE)
    }
}

template <int N, int WARPS, int BLOCKS_PER_SM, bool AMAF>
cudaError_t launch_w(const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    const size_t smem = (size_t)WARPS * (AMAF ? BoardWords<N>::value_amaf : BoardWords<N>::value) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(playout_warp_kernel<N, WARPS, BLOCKS_PER_SM, AMAF>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t want = (a.local_units + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)sm_count * BLOCKS_PER_SM;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (info) { info->grid = grid; info->block = WARPS * 32; info->smem = smem; info->name = "playout_warp_kernel"; }
    playout_warp_kernel<N, WARPS, BLOCKS_PER_SM, AMAF><<<grid, WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

unsigned long long debug_event(int which) {
#ifdef AGB_CHECK_BOUNDS
    unsigned long long v[4] = {0, 0, 0, 0};
    if (which < 0 || which >= 4 || cudaMemcpyFromSymbol(v, g_debug_events, sizeof(v)) != cudaSuccess) return ~0ull;
    return v[which];
#else
    (void)which;
    return 0;
#endif
}

unsigned long long debug_violations() {
#ifdef AGB_CHECK_BOUNDS
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_bounds_violations, sizeof(v)) != cudaSuccess) return ~0ull;
    return v;
#else
    return 0;
#endif
}

cudaError_t launch_playout_warp(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    // 8 warps per block, 4 blocks per SM: 32 resident playouts per SM, 64 registers per thread
    switch (n) {
        case 9: return a.amaf ? launch_w<9, 8, 4, true>(a, sm_count, s, info) : launch_w<9, 8, 4, false>(a, sm_count, s, info);
        case 13: return a.amaf ? launch_w<13, 8, 4, true>(a, sm_count, s, info) : launch_w<13, 8, 4, false>(a, sm_count, s, info);
        case 19: return a.amaf ? launch_w<19, 8, 4, true>(a, sm_count, s, info) : launch_w<19, 8, 4, false>(a, sm_count, s, info);
    }
    return cudaErrorInvalidValue;
}

}  // namespace agb
