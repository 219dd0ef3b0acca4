// playout_group.cuh — SEVERAL playouts per warp: a group of G lanes (8 or 16; 32 for reference) owns
// one playout, 32/G playouts advance in lock step in one warp (sm_100a).
//
// Why: the warp-per-playout kernel (playout_warp.cu) issues on 80 % of the scheduler cycles, but
// only 13 % of its lane slots do algorithmic work — the board update, the loop glue and the scalar
// parts of the move selection keep one lane busy and 31 waiting.  Here
// those sections are shared by 32/G playouts, and the predicates of the policy are evaluated ONE
// LANE PER POINT (all four directions of a point in one lane, no warp collectives inside), so a
// neighbourhood evaluation (8 cells) or a rejection round (G draws pairs) of 32/G playouts is one
// pass over the warp.
//
// Rules of this file:
//  * CONTROL FLOW IS WARP-UNIFORM.  Every loop bound and every branch around a collective is
//    derived from a ballot / any over the whole warp; what differs between the playouts of a
//    warp is DATA (predicates per group).  Every collective therefore runs with the full mask in
//    converged code.  tests/csrc/simt_emu.h runs this very source on the CPU and aborts if any lane
//    reaches a collective at another site.
//    CONVERGENCE NOTE.  ptxas has to PROVE the convergence, or it guards every collective of the kernel
//    with UMOV + BRA.DIV and a WARPSYNC.COLLECTIVE slow path (8 % of the executed instructions, a
//    quarter of the code).  Two things defeated the proof and are avoided (found by cutting the kernel
//    down until the guards went away, profiles/r02_convergence.txt): (1) a loop whose body ENDS in a
//    lane-dependent `if` gets two back edges, one of them out of divergent code, and the loop header —
//    hence everything in the loop — counts as divergent: such loops end in a __syncwarp() (a NOP once
//    convergence is proven), which gives them a single latch behind the reconvergence point; (2) values
//    that are warp-uniform by construction but arrive through a shuffle, shared memory or a function
//    argument make the branches on them "divergent": they are passed through a redux (its result is
//    a uniform register) or the branch is taken on a vote.
//  * RARE BULK WORK BORROWS THE WHOLE WARP.  Clone, scoring, complex mode
//    (Board::getRandomPieceComplex) and the compaction of the empty-point array serve one group at
//    a time with all 32 lanes (the board of any group of the warp is in reach of every lane).
//  * One 32-bit word per point (the warp kernel had two) of a padded board with ONE border column (GCfg;
//    word layout: kernel_args.h, gw::):
//        bits 0-1   colour (0 empty, 1 black, 2 white, 3 border)
//        bits 2-10  A     stone: root of its string; empty / border point: its own index
//        bits 11-19 next  stone: next stone of the string in SetNode::pieces order (511 = end);
//                         empty point: its slot in the empty-point array `elist`
//        bits 21-31 F     root stone: pseudo-liberties (SetNode::hp); SECOND stone of a string: the
//                         string's last stone (the tail is needed only when strings are joined, the
//                         liberties on every lookup, so the root keeps the field that is one shift
//                         away); 1 on empty and border points
//    With A = own index on empty and border points the root of "whatever is next to a point" is
//    one AND away with no case split (w & 0x7fc is even its byte offset), and equal roots mean
//    the same string.
//  * The empty-point list is the append-only array of the warp kernel (kernel_args.h) — in GLOBAL
//    memory (L2): only complex mode and compaction read it, all 32 lanes at a time, and the 752 B
//    it took per playout were five warps per SM.  The rand() ring is smaller than the warp kernel's
//    (R = 2*G*SEG entries) and refilled whenever a phase is about to read past its end.  State per
//    playout in shared memory: 2 224 B at 19x19, 100 playouts = 25 warps per SM.
//
// Restates Board.cpp:68-177 (put), :293-339 (killSetNode), :341-411 (getRandomPiece), :413-478
// (getRandomPieceComplex), :480-755 (predicates), :910-925 (countScore) and the pair loops of
// AgentGo.cpp:285-327 / :412-448, bit for bit (tests/test_emu_group.py, tests/test_gpu_parity.py).
#ifndef AGB_PLAYOUT_GROUP_CUH_
#define AGB_PLAYOUT_GROUP_CUH_

#include "kernel_args.h"

#ifdef AGB_EMU
#define AGB_DEV inline
#define AGB_NOINLINE __attribute__((noinline))
#else
#define AGB_DEV __device__ __forceinline__
#define AGB_NOINLINE __device__ __noinline__
#endif

// tunables of the move selection (measured: profiles/r02_tuning.txt)
#ifndef AGB_KSCAN
#define AGB_KSCAN 32
#endif
#ifndef AGB_KWANT
#define AGB_KWANT 3
#endif
#ifndef AGB_NBFILTER
#define AGB_NBFILTER 1
#endif
#ifndef AGB_WALK_UNROLL
#define AGB_WALK_UNROLL 4
#endif

namespace agb {
namespace grp {

constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kNilG = 511u;
constexpr uint32_t kHpOne = gw::kOne;
constexpr uint32_t kHpMask = 0x7ffu << gw::kFShift;
constexpr uint32_t kNextMask = 511u << gw::kNextShift;
constexpr uint32_t kAMask = 511u << gw::kAShift;
constexpr uint32_t kResFlag = 0x8000u;     // elist entry: point is reserved in the current complex-mode call

AGB_DEV uint32_t fA(uint32_t w) { return (w >> gw::kAShift) & 511u; }
AGB_DEV uint32_t offA(uint32_t w) { return w & kAMask; }                 // byte offset of the root's word
AGB_DEV uint32_t fNext(uint32_t w) { return (w >> gw::kNextShift) & 511u; }
AGB_DEV uint32_t fSlot(uint32_t w) { return (w >> gw::kNextShift) & 1023u; }         // empty point: its entry in elist
AGB_DEV uint32_t fHp(uint32_t w) { return w >> gw::kFShift; }
AGB_DEV uint32_t fCol(uint32_t w) { return w & 3u; }
AGB_DEV bool isEmpty(uint32_t w) { return (w & 3u) == 0; }
AGB_DEV uint32_t setA(uint32_t w, uint32_t a) { return (w & ~kAMask) | (a << gw::kAShift); }
AGB_DEV uint32_t setNext(uint32_t w, uint32_t n) { return (w & ~kNextMask) | (n << gw::kNextShift); }
AGB_DEV uint32_t setF(uint32_t w, uint32_t f) { return (w & ~kHpMask) | (f << gw::kFShift); }
// the empty-point array lives in global memory and is handed from lane to lane: L2 accesses
#ifdef AGB_EMU
AGB_DEV uint32_t eld(const uint16_t* p) { return *p; }
AGB_DEV void est(uint16_t* p, uint32_t v) { *p = (uint16_t)v; }
AGB_DEV void est32(uint32_t* p, uint32_t v) { *p = v; }
#else
AGB_DEV uint32_t eld(const uint16_t* p) { return __ldcg(p); }
AGB_DEV void est(uint16_t* p, uint32_t v) { __stcg(p, (unsigned short)v); }
AGB_DEV void est32(uint32_t* p, uint32_t v) { __stcg(p, v); }
#endif
AGB_DEV uint32_t at(const uint32_t* w, uint32_t byte_off) {              // the word at a byte offset
    return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(w) + byte_off);
}

// AMAF_: the playouts also record the first-play marks of MyWorker::work (AgentGo.cpp:289-290,304,319)
// and accumulate KernelArgs::amaf (:329-338); candidate mode only.
template <int N_, int G_, int SEG_, bool AMAF_ = false>
struct GCfg {
    static constexpr int N = N_, G = G_, SEG = SEG_;
    static constexpr bool AMAF = AMAF_;
    static constexpr int kAmafLog = 80;             // first plays with cstep <= 40: at most 78
    // The padded board has ONE border column: the point right of a row's last point is the point left of
    // the next row's first, and both are border.  Rows -1 .. N of N + 1 points each, and one more point
    // for the lower right diagonal of the last point: (N + 1)(N + 2) + 1 words (421 at 19x19; two border
    // columns took 441, and the 80 B are the difference between 24 and 25 warps per SM).
    static constexpr int W = N + 1, NP = W * (N + 2) + 1, NN = N * N, NG = 32 / G;
    static constexpr int BATCH = G * SEG;           // draws per refill of one playout
    static constexpr int R = 2 * BATCH;             // ring entries
    static constexpr int E = kGroupElistCap;        // elist entries, in global memory (kernel_args.h)
    // shared-memory words of one playout: ring | board | cbuf, a multiple of 4
    static constexpr int oRing = 0, oBoard = R / 2, oCbuf = oBoard + NP;
    // cbuf: G + 1 16-bit entries: the candidates of a move, and the first one that found no room (where
    // the pass ends).  The stride from one playout of a warp to the next is a multiple of 4 words (the
    // ring is written with 128-bit stores) and such that the same field of the playouts of a warp
    // (cbuf, ring position, board word) sits in different banks: k * words is no multiple of 32 for 0 < k < NG.
    static constexpr int oLog = oCbuf + (G + 2) / 2;    // AMAF only: kAmafLog 16-bit entries
    static constexpr int words4 = ((oLog + (AMAF ? kAmafLog / 2 : 0) + 3) / 4) * 4;
    static constexpr bool spread(int w) {
        for (int k = 1; k < NG; k++) if ((k * w) % 32 == 0) return false;
        return true;
    }
    static constexpr int pick_words(int w) { return spread(w) ? w : pick_words(w + 4); }
    static constexpr int words = pick_words(words4);
    static constexpr int warp_words = NG * words;
    static constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
};

// v % M for a 15-bit value and a small constant M: one IMAD.HI and one IMAD
template <uint32_t M>
AGB_DEV uint32_t mod_small(uint32_t v) {
    const uint32_t q = __umulhi(v, (uint32_t)((0x100000000ull + M - 1) / M));
    return v - q * M;
}

// ---- the policy's predicates, one lane per point -------------------------------------------------
// checkSuicide/Kill/Survive/Dying/Chase/TrueEye/Compete (Board.cpp:480-755) of the EMPTY on-board
// point idx for `agent`, computed by one lane from the 4 + 4 words around it.
//   reserve = true eye | suicide | ko            (complex mode: no checkDying, :425-427)
//   bad     = reserve | dying                    (the filter of both sampling phases)
// The four directions are the four BYTES of a register: colours, liberties (clamped to 255, which
// no test below can tell from the true value) and multiplicities are packed once, classes come out
// of the carry-free "byte != 0" trick as 0x80 flags, sums out of dp4a.  Thirty booleans kept in
// predicates made the compiler spill them into a register bit by bit (220 instructions per call).
enum { kEvAll = 0, kEvRes = 2 };
struct Ev {
    bool kill, survive, chase, bad;     // kEvRes: bad = reserve
};
constexpr uint32_t k80 = 0x80808080u, k7f = 0x7f7f7f7fu, k01 = 0x01010101u;
// 0x80 in every byte of x that is not zero; the bytes of x are below 0x80
AGB_DEV uint32_t nz_small(uint32_t x) { return (x + k7f) & k80; }
// the same for any bytes
AGB_DEV uint32_t nz_bytes(uint32_t x) { return (((x & k7f) + k7f) | x) & k80; }
template <int W, int MODE>
AGB_DEV Ev eval_point(const uint32_t* w, uint32_t agent, int idx, int ko_key) {
    const uint32_t agent4 = agent * k01;
    const uint32_t v0 = w[idx - W], v1 = w[idx - 1], v2 = w[idx + W], v3 = w[idx + 1];
    // colours of up, left, down, right in bytes 0..3
    const uint32_t cb = __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410) & 0x03030303u;
    const uint32_t xf = cb ^ agent4;                        // zero byte: own stone
    const uint32_t nF = nz_small(xf), nE = nz_small(xf ^ 0x03030303u), nZ = nz_small(cb);    // 0x80: NOT own / enemy / empty
    const uint32_t F80 = nF ^ k80, E80 = nE ^ k80;
    const int n_empty = 4 - __popc(nZ);
    // liberties of the strings around (1 on a point without a stone, see gw::), less the number of
    // directions that see the same string (minus_hp, :506-515)
    const uint32_t o0 = offA(v0), o1 = offA(v1), o2 = offA(v2), o3 = offA(v3);
    const uint32_t h0 = min(fHp(at(w, o0)), 255u), h1 = min(fHp(at(w, o1)), 255u);
    const uint32_t h2 = min(fHp(at(w, o2)), 255u), h3 = min(fHp(at(w, o3)), 255u);
    const uint32_t H = h0 | (h1 << 8) | (h2 << 16) | (h3 << 24);
    const bool e01 = o0 == o1, e02 = o0 == o2, e03 = o0 == o3, e12 = o1 == o2, e13 = o1 == o3, e23 = o2 == o3;
    const uint32_t M = k01 + (e01 ? 0x00000101u : 0u) + (e02 ? 0x00010001u : 0u) + (e03 ? 0x01000001u : 0u) +
                       (e12 ? 0x00010100u : 0u) + (e13 ? 0x01000100u : 0u) + (e23 ? 0x01010000u : 0u);
    const uint32_t L = H - M;                               // left = hp - multiplicity, per byte
    const uint32_t nL0 = nz_bytes(L);                       // 0x80: left != 0
    const bool ecap = (E80 & ~nL0) != 0;                    // an enemy string is captured
    const bool fsafe = (F80 & nL0) != 0;                    // an own string keeps a liberty
    const bool suicide = n_empty == 0 && !ecap && !fsafe;                       // :480-524
    // Board::checkTrueEye, :735-751: nothing but own stones and border around, and fewer than two
    // enemy stones on the diagonals (none on the edge: an orthogonal neighbour is border there)
    const bool own4 = (E80 | (nZ ^ k80)) == 0;              // no enemy stone, no empty point
    const bool edge = (nF & nE & nZ) != 0;                  // neither own nor enemy nor empty: border
    bool eye = false;
    if (own4) {
        const uint32_t d0 = w[idx - W - 1], d1 = w[idx - W + 1], d2 = w[idx + W - 1], d3 = w[idx + W + 1];
        const uint32_t db = __byte_perm(__byte_perm(d0, d1, 0x0040), __byte_perm(d2, d3, 0x0040), 0x5410) & 0x03030303u;
        const int cnt = 4 - __popc(nz_small(db ^ agent4 ^ 0x03030303u));
        eye = cnt < (edge ? 1 : 2);
    }
    const bool ko = ko_key == (int)((agent << 16) | (uint32_t)idx);             // :753-755
    Ev e;
    e.kill = e.survive = e.chase = false;
    e.bad = eye || suicide || ko;
    if (MODE == kEvRes) return e;
    // own strings count once, where they are first seen
    const uint32_t first = k01 & ~((e01 ? 0x00000100u : 0u) | ((e02 || e12) ? 0x00010000u : 0u) |
                                   ((e03 || e13 || e23) ? 0x01000000u : 0u));
    const int tot = n_empty + (int)__dp4a(first & (F80 >> 7), L, 0u);
    e.bad = e.bad || (!ecap && tot == 1);                                       // checkDying, :628-679
    if (MODE == kEvAll) {
        const bool eatari = (E80 & ~nz_bytes(L ^ k01)) != 0;                    // an enemy string is left with one
        const bool fcap = (F80 & ~nL0) != 0;                                    // an own string has none left
        e.kill = ecap;                                                          // :526-563
        e.chase = eatari && tot > 1;                                            // :681-733
        if (fcap) {                                                             // :565-626
            const uint32_t hf = H | ((nF >> 7) * 255u);     // 255 where the neighbour is not an own stone
            const uint32_t fmin = min(min(hf & 255u, (hf >> 8) & 255u), min((hf >> 16) & 255u, hf >> 24));
            e.survive = ecap || tot > (int)fmin;
        }
    }
    return e;
}

// ---- rand(): ring refill ------------------------------------------------------------------------
// Lane j of a playout's G lanes generates draws [j*SEG, (j+1)*SEG) of every batch of BATCH = G*SEG, and
// keeps for that a generator state of its OWN: the state its NEXT segment starts from.  TinyMT32's
// transition is linear over GF(2), so that state moves from one batch to the next by the same matrix
// for every lane, T^BATCH (a nibble table in shared memory), no other lane involved.  When a playout is
// set up lane j jumps from the seed state by T^(SEG*j) (matrix j of JumpBytes).
// A refill TOPS UP every live playout of the warp, segment by segment: a lane generates its next segment
// iff the ring has room for it (the slots it overwrites hold consumed draws), so the frontier gen_end of a
// playout moves to the last segment boundary the ring can hold, whatever the playout had left.  (Batches
// taken or left as a whole served 1.8 of the 4 playouts per call: the one that ran dry and those that
// happened to have half a ring free.  The call costs the same for any number of takers.)
struct GenState {
    uint32_t g0, g1, g2, g3;    // this lane's state at the start of its next segment: the first draw at or
                                // after gen_end that is SEG * (lane in group) modulo BATCH
    uint32_t gen_end;           // draws [pos, gen_end) of the playout are in the ring; a multiple of SEG
};
AGB_DEV uint4 jump_lookup(const uint4* __restrict__ tab, uint32_t g0, uint32_t g1, uint32_t g2, uint32_t g3) {
    uint4 t = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const uint32_t word = b < 4 ? g0 : (b < 8 ? g1 : (b < 12 ? g2 : g3));
        const uint4 u = __ldg(tab + b * 256 + ((word >> (8 * (b & 3))) & 255u));
        t.x ^= u.x; t.y ^= u.y; t.z ^= u.z; t.w ^= u.w;
    }
    return t;
}
template <class C>
AGB_NOINLINE GenState refill(bool live, uint32_t pos, uint16_t* ring, GenState st, RngParams prm, const uint4* tab, int gl) {
    // this lane's next segment, and how far the ring can be filled: a segment at s needs s + SEG <= pos + R;
    // at most one segment per lane and call
    const uint32_t mine_at = st.gen_end + (((uint32_t)(C::SEG * gl) - st.gen_end) & (uint32_t)(C::BATCH - 1));
    // (gen_end <= pos + R and is a multiple of SEG, so the boundary below is never behind it)
    const uint32_t frontier = min((pos + (uint32_t)C::R) & ~(uint32_t)(C::SEG - 1), st.gen_end + (uint32_t)C::BATCH);
    const bool part = live && mine_at < frontier;
    // tab: T^BATCH split by state NIBBLE (JumpTable: entry l*16 + v is the image of nibble l having value
    // v), 8 KB in SHARED memory — the byte-split matrix in L2 cost 16 lookups instead of 32, but their
    // latency was 5 % of all stall samples and their traffic 2.9 TB/s
    uint4 t = make_uint4(0, 0, 0, 0);
#pragma unroll 1
    for (int q = 0; q < 4; q++) {       // (rolled: eight 128-bit loads in flight are what the registers hold)
        const uint32_t word = q == 0 ? st.g0 : (q == 1 ? st.g1 : (q == 2 ? st.g2 : st.g3));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 u = tab[(8 * q + k) * 16 + ((word >> (4 * k)) & 15u)];
            t.x ^= u.x; t.y ^= u.y; t.z ^= u.z; t.w ^= u.w;
        }
    }
    if (part) {         // (no collective inside: the lanes of a group that does not take part just wait)
        TinyMT mine;
        mine.s0 = st.g0; mine.s1 = st.g1; mine.s2 = st.g2; mine.s3 = st.g3;
        uint32_t* out = reinterpret_cast<uint32_t*>(ring + (mine_at & (uint32_t)(C::R - 1)));
#pragma unroll 2
        for (int j = 0; j < C::SEG / 2; j++) {
            mine.next_state(prm.mat1, prm.mat2);
            const uint32_t lo = mine.temper(prm.tmat) >> 17;
            mine.next_state(prm.mat1, prm.mat2);
            const uint32_t hi = mine.temper(prm.tmat) >> 17;
            out[j] = lo | (hi << 16);
        }
        st.g0 = t.x; st.g1 = t.y; st.g2 = t.z; st.g3 = t.w;
    }
    if (live) st.gen_end = frontier;
    __syncwarp();
    return st;
}

// ---- bulk work, all 32 lanes for one group ------------------------------------------------------
// Drop the dead entries of elist, keeping the order.  Entry k naming point c is alive iff c is
// empty and its slot is k (a zeroed entry names the border point 0).  In place: a pass reads 128
// entries (the four loads are in flight together: the array lives in L2) and then writes at or
// below where it read.  Returns the new number of entries.
template <class C>
AGB_NOINLINE int compact(uint32_t* w, uint16_t* elist, int n_el, int lane) {
    int out = 0;
#pragma unroll 1
    for (int base = 0; base < n_el; base += 128) {
        uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = base + 32 * u + lane;
            c[u] = k < n_el ? (eld(elist + k) & 0x1ffu) : 0u;
        }
        __syncwarp();       // all of the pass is read before any of it is overwritten
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = base + 32 * u + lane;
            const uint32_t wc = w[c[u]];
            const bool alive = isEmpty(wc) && (int)fSlot(wc) == k;
            const unsigned m = __ballot_sync(kFull, alive);
            if (alive) {
                const int o = out + __popc(m & ((1u << lane) - 1u));
                est(elist + o, c[u]);
                w[c[u]] = gw::empty(c[u], (uint32_t)o);
            }
            out += __popc(m);
        }
    }
    __syncwarp();
    return out;
}

// Board::getRandomPieceComplex, Board.cpp:413-478, for one group.  Pass 1 (reserve flags, :419-433):
// one lane per entry of elist.  Pass 2 (:445-470, the rnd-th unreserved point in list order) counts
// 32 entries per round from the end of the array.  The first kKeep rounds of entries — all of them,
// as a rule: complex mode means few empty points — stay in registers with their flags between the two
// passes (64 entries), so the array (it lives in L2) is read once, all rounds in flight together, and not written;
// entries beyond that are flagged in place.  draw = the playout's next rand() value; *used says
// whether it was consumed.
template <class C>
AGB_NOINLINE int complex_pick(uint32_t* w, uint16_t* elist, int n_el, int spaces, uint32_t agent, int ko_key,
                              uint32_t draw, int lane, int* used) {
    constexpr int kKeep = 2;
    // The arguments are the same in all lanes, but the compiler cannot see that through the call: a redux
    // puts them in uniform registers, and every branch below becomes a uniform branch (convergence note).
    n_el = (int)__reduce_max_sync(kFull, (unsigned)n_el);
    spaces = (int)__reduce_max_sync(kFull, (unsigned)spaces);
    draw = __reduce_max_sync(kFull, draw);
    uint32_t ents[kKeep];
#pragma unroll
    for (int r = 0; r < kKeep; r++) {
        const int k = 32 * r + lane;
        ents[r] = k < n_el ? (eld(elist + k) & 0x1ffu) : 0u;                        // 0: a dead entry
    }
    int reserved_total = 0;
#pragma unroll
    for (int r = 0; r < kKeep; r++) {
        if (32 * r < n_el) {
            // dead entries look at some point of the board; their result is dropped
            const Ev e = eval_point<C::W, kEvRes>(w, agent, ents[r] ? (int)ents[r] : C::W + 1, ko_key);
            const bool res = e.bad && ents[r] != 0;
            if (res) ents[r] |= kResFlag;                                           // addReserve, :757-761
            reserved_total += __popc(__ballot_sync(kFull, res));
        }
    }
#pragma unroll 1
    for (int base = 32 * kKeep; base < n_el; base += 32) {
        const int k = base + lane;
        const int ent = k < n_el ? (int)(eld(elist + k) & 0x1ffu) : 0;
        const Ev e = eval_point<C::W, kEvRes>(w, agent, ent ? ent : C::W + 1, ko_key);
        const bool res = e.bad && ent != 0;
        if (ent) est(elist + k, (uint32_t)ent | (res ? kResFlag : 0u));
        reserved_total += __popc(__ballot_sync(kFull, res));
    }
    __syncwarp();
    const int range = spaces - reserved_total;
    int pick = -1;
    *used = 0;
    if (range != 0) {
        int rnd = (int)(draw % (uint32_t)range);
        *used = 1;
        int found = 0;
        // list order = the array from its end: first what is beyond the registers ...
#pragma unroll 1
        for (int top = n_el - 1; top >= 32 * kKeep && !found; top -= 32) {
            const int k = top - lane;                                               // lane order = list order
            const uint32_t ent = k >= 32 * kKeep ? eld(elist + k) : 0u;
            const bool ok = ent - 1u < 0x7fffu;                                     // alive and not reserved
            const unsigned m = __ballot_sync(kFull, ok);
            const int c = __popc(m);
            if (rnd < c) {
                const bool hit = ok && __popc(m & ((1u << lane) - 1u)) == rnd;
                pick = (int)__shfl_sync(kFull, ent, __ffs(__ballot_sync(kFull, hit)) - 1);
                found = 1;
            } else {
                rnd -= c;
            }
        }
        // ... then the rounds in registers, last round first, high lane first
#pragma unroll
        for (int r = kKeep - 1; r >= 0; r--) {
            if (!found && 32 * r < n_el) {
                const bool ok = ents[r] - 1u < 0x7fffu;
                const unsigned m = __ballot_sync(kFull, ok);
                const int c = __popc(m);
                if (rnd < c) {
                    // the rnd-th from the top is the (c-1-rnd)-th from the bottom
                    const bool hit = ok && __popc(m & ((1u << lane) - 1u)) == c - 1 - rnd;
                    pick = (int)(__shfl_sync(kFull, ents[r], __ffs(__ballot_sync(kFull, hit)) - 1) & 0x1ffu);
                    found = 1;
                } else {
                    rnd -= c;
                }
            }
        }
    }
    return pick;
}

// Board::countScore(agent) - countScore(enemy), Board.cpp:910-925: an empty point counts for the
// colour of its LEFT neighbour (right one in column 0)
template <class C>
AGB_DEV int score_diff(const uint32_t* w, uint32_t agent, int lane) {
    int d = 0;
#pragma unroll 1
    for (int idx = C::W + lane; idx < C::NP - C::W; idx += 32) {
        const uint32_t c = fCol(w[idx]);
        if (c == 3) continue;
        uint32_t owner = c;
        if (c == 0) {
            const uint32_t l = fCol(w[idx - 1]);
            owner = l != 3 ? l : fCol(w[idx + 1]);
        }
        d += (int)(owner == agent) - (int)(owner == 3u - agent);
    }
    return (int)__reduce_add_sync(kFull, (unsigned)d);
}

// ---- the kernel body ----------------------------------------------------------------------------
template <class C>
AGB_DEV void playout_group_body(const KernelArgs& a, uint32_t* smem_warp, int lane, int warp_slot, const uint4* jump_tab) {
    constexpr int G = C::G, W = C::W;
    constexpr uint32_t RM = (uint32_t)(C::R - 1);
    const int gl = lane & (G - 1);
    const int gbase = lane & ~(G - 1);
    const int grp = lane / G;
    uint32_t* const gsm = smem_warp + grp * C::words;
    uint16_t* const ring = reinterpret_cast<uint16_t*>(gsm + C::oRing);
    uint32_t* const w = gsm + C::oBoard;
    uint16_t* const elist_warp = a.elist_scratch + (size_t)warp_slot * (C::NG * C::E);   // this warp's playouts
    uint16_t* const elist = elist_warp + grp * C::E;
    uint16_t* const cbuf = reinterpret_cast<uint16_t*>(gsm + C::oCbuf);
    const uint4* __restrict__ jt = a.jump_bytes;

    // lane constants: direction of the stone update (lanes 4k..4k+3 = up, left, down, right) and
    // cell of the 3x3 block around the opponent's last move (centre skipped)
    const int dir = lane & 3;
    const int off_nb = dir == 0 ? -W : (dir == 1 ? -1 : (dir == 2 ? W : 1));
    const int c8 = gl & 7;
    const int cc = c8 + (c8 >= 4);
    const int off_cell = (cc / 3 - 1) * W + (cc % 3 - 1);
    const bool orth = (cc & 1) != 0;                // the cell touches the centre of the block
    const unsigned qmask = 15u << (lane & ~3);
    const int qbase = lane & ~3;

    // A group that never gets a playout (fewer units than groups) still runs along with safe
    // indices; what it reads must not send it out of the block's shared memory: all words zero
    // (an empty point whose root is point 0).
#pragma unroll 1
    for (int i = lane; i < C::warp_words; i += 32) smem_warp[i] = 0;
    __syncwarp();

    // state of this lane's playout, the same in all G lanes of the group
    enum { kFetch = 0, kPlay = 1, kEnded = 2, kDone = 3 };
    int phase = kFetch;
    uint32_t colour = 1;
    // (flags that live across a loop are ints: a bool is kept as a byte and costs a PRMT or two per test)
    int closes = 0;             // the move about to be made closes a pair (it is the second mover's)
    int prev = -1, stones = 0, ko_key = -1, ending = 0, n_el = 0, n_dead = -1;
    uint32_t pos = 0, mv = 0;
    GenState gs = {0, 0, 0, 0, 0};
    uint32_t unit = 0;          // which of this shard's units (the engine keeps a launch below 2^32 of them)
    int fault = 0;
    // AMAF marks: a log of the first plays with cstep <= 40 (key = point | 512 for the agent's), resolved
    // when the playout ends; cstep is 2 in the first pair (:293,:297)
    constexpr int kAmafRange = 40;      // AMAF_RANGE, GoContest/AgentGo.h:16
    uint16_t* const amaf_log = reinterpret_cast<uint16_t*>(gsm + C::oLog);
    int n_log = 0, cstep = 2;
    uint32_t amaf_agent = 0;

#pragma unroll 1
    for (;;) {
        // ---- service: score the playouts that ended, hand out new ones (all 32 lanes per group) ----
        unsigned need = __ballot_sync(kFull, phase == kFetch || phase == kEnded);
#pragma unroll 1
        while (need) {
            const int X = (__ffs(need) - 1) / G;
            const int src = X * G;
            need &= ~(C::GM << src);
            uint32_t* const xsm = smem_warp + X * C::words;
            uint32_t* const xw = xsm + C::oBoard;
            if (__shfl_sync(kFull, phase, src) == kEnded) {
                // (start and global id are worked out again from the unit: two registers less in the move loop)
                const long long xg = (long long)__shfl_sync(kFull, unit, src) * a.shard_count + a.shard_rank;
                const int xs = (int)(xg / a.P);
                const uint32_t xmv = __shfl_sync(kFull, mv, src);
                const uint32_t xpos = __shfl_sync(kFull, pos, src);
                const bool xfault = __shfl_sync(kFull, (int)fault, src) != 0;
                const StartDesc d = a.descs[xs];
                const int diff = xfault ? -32768 : score_diff<C>(xw, (uint32_t)d.agent, lane);
                if (C::AMAF && !xfault) {
                    // :329-338 — point q counts for the agent table when the agent's first play there has
                    // cstep <= 40, else for the enemy table when the enemy's has (the `else if`)
                    const uint16_t* xlog = reinterpret_cast<const uint16_t*>(xsm + C::oLog);
                    const int xn = __shfl_sync(kFull, n_log, src);
                    const int wv = diff - (a.avrg_dev ? *a.avrg_dev : d.avrg_win);
                    uint32_t agent_bits = 0;                // bit (q & 31) of lane (q >> 5): the agent marked q
#pragma unroll 1
                    for (int k = 0; k < xn; k++) {
                        const uint32_t e = xlog[k];
                        if ((e & 512u) && lane == (int)((e & 511u) >> 5)) agent_bits |= 1u << (e & 31u);
                    }
#pragma unroll 1
                    for (int b0 = 0; b0 < xn; b0 += 32) {
                        const bool valid = b0 + lane < xn;
                        const uint32_t e = valid ? xlog[b0 + lane] : 0u;
                        const uint32_t q = e & 511u;
                        const bool is_agent = (e & 512u) != 0;
                        const bool agent_marked = (__shfl_sync(kFull, agent_bits, (int)(q >> 5)) >> (q & 31u)) & 1u;
                        if (valid && wv != 0 && (is_agent || !agent_marked)) {
                            const uint32_t row = q / W - 1, col = q % W - 1;
                            const uint32_t slot = ((is_agent ? 0u : 2u) + (wv < 0 ? 1u : 0u)) * (kAmafRange + 1) + (e >> 10);
                            atomicAdd((unsigned long long*)&a.amaf[((size_t)d.amaf_class * (4 * (kAmafRange + 1)) + slot) * C::NN + row * C::N + col],
                                      (unsigned long long)(long long)wv);
                        }
                    }
                }
                if (lane == 0) {
                    atomicAdd(&a.stats[kStatMoves], (unsigned long long)xmv);
                    atomicAdd(&a.stats[kStatDraws], (unsigned long long)xpos);
                    atomicAdd(&a.stats[kStatPlayouts], 1ull);
                    if (xfault) atomicAdd(&a.stats[kStatFaults], 1ull);
                    if (a.diffs) a.diffs[xg] = (int16_t)diff;
                    if (!xfault) {
                        const int wr = diff - (a.avrg_dev ? *a.avrg_dev : d.avrg_win);       // AgentGo.cpp:327
                        unsigned long long* c = (unsigned long long*)a.counters;
                        if (wr >= 0) {                                           // sign test of AgentGo.cpp:328
                            atomicAdd(&c[xs], 1ull);
                            atomicAdd(&c[a.counter_stride + xs], (unsigned long long)wr);
                        } else {
                            atomicAdd(&c[2 * a.counter_stride + xs], (unsigned long long)(long long)wr);
                        }
                    }
                }
            }
            unsigned long long l = 0;
            if (lane == 0) l = atomicAdd(a.work_counter, 1ull);
            l = __shfl_sync(kFull, l, 0);
            const bool has = l < (unsigned long long)a.local_units;
            if (has) {
                const long long g = (long long)l * a.shard_count + a.shard_rank;
                const int s = (int)(g / a.P);
                const long long p = g - (long long)s * a.P;
                const GroupPosition& gp = a.gstarts[s];
                const StartDesc d = a.descs[s];
                __syncwarp();
#pragma unroll 1
                for (int i = lane; i < C::NP; i += 32) xw[i] = gp.w[i];          // Board::clone, Board.cpp:216-254
                {
                    uint32_t* xe = reinterpret_cast<uint32_t*>(elist_warp + X * C::E);
#pragma unroll 1
                    for (int i = lane; i < (C::NN + 1) / 2; i += 32) est32(xe + i, gp.elist_w[i]);
                }
                __syncwarp();
                int lm = gp.last[2 - d.first];                                   // last_move[enemy of the first mover]
                // Board.cpp:356 for the first move: the opponent's last stone may have been captured since
                if (lm >= 0 && fCol(xw[lm]) != 3u - (uint32_t)d.first) lm = -1;
                // tinymt32_init (RFC 8682)
                TinyMT t;
                t.init(playout_seed(a.seed_base, d.position_id, d.cand_idx, (uint64_t)p), a.rng.mat1, a.rng.mat2, a.rng.tmat);
                // this lane's segment of the first batch starts SEG * gl draws further on (matrix gl)
                const uint4 js = jump_lookup(jt + gl * (16 * 256), t.s0, t.s1, t.s2, t.s3);
                if (gl != 0) { t.s0 = js.x; t.s1 = js.y; t.s2 = js.z; t.s3 = js.w; }
                if (grp == X) {
                    phase = kPlay;
                    colour = (uint32_t)d.first;
                    closes = 0;
                    prev = lm;
                    stones = gp.stones;
                    ko_key = gp.ko_key_col > 0 ? ((gp.ko_key_col << 16) | gp.ko_idx) : -1;
                    ending = gp.ending;
                    n_el = gp.n_el;
                    n_dead = -1;
                    pos = 0; mv = 0;
                    gs.g0 = t.s0; gs.g1 = t.s1; gs.g2 = t.s2; gs.g3 = t.s3; gs.gen_end = 0;
                    unit = (uint32_t)l; fault = 0;
                    n_log = 0; cstep = 2; amaf_agent = (uint32_t)d.agent;
                }
            } else if (grp == X) {
                phase = kDone;
            }
        }
        if (!__any_sync(kFull, phase == kPlay)) break;
        const bool live = phase == kPlay;

#define AGB_ENSURE(needpred, k)                                                                     \
        if (__any_sync(kFull, (needpred) && gs.gen_end - pos < (uint32_t)(k))) {                     \
            gs = refill<C>(live, pos, ring, gs, a.rng, jump_tab, gl);                               \
        }

        // ---- Board::getRandomPiece, Board.cpp:341-411 ----
        int cplx = 0;
        if (ending) {                                                           // :348-351
            if (stones <= kEndingStones) ending = 0;
            else cplx = 1;
        }
        const bool search = live && !cplx;
        if (search) n_dead = -1;                                                // put stops maintaining elist
        int r = -1;
        bool found = false;
        // Both sampling loops of the reference (neighbourhood :354-380, rejection :382-406) consume two
        // rand() per iteration whatever happens inside, and their predicates are pure functions of a
        // state that does not change inside the loops.  So ONE evaluation pass serves both: its items
        // are the empty cells among the 8 around the opponent's last move and, in the lanes that
        // leaves, the first empty points of the rejection loop's pairs (which start 18 draws further on
        // when there is a neighbourhood phase).  Pairs are scanned G at a time — looking at a pair is
        // cheap, evaluating a point is not — until the pass holds kWant of them or kScan pairs were seen.  The
        // neighbourhood decides first; if it accepts nothing the rejection candidates are looked at in
        // draw order; if they were all bad the next pass continues after the pairs this one settled.
        // (:356 also wants the opponent's last stone to be still there; inside a playout it always is,
        // and for the first move of a playout that was settled when it was set up.)
        const bool nbact = search && prev >= 0;
        {
            constexpr int kScan = AGB_KSCAN;    // pairs looked at per pass at most (their index has 5 bits in an item)
            constexpr int kWant = AGB_KWANT;            // ... until the pass holds this many rejection candidates: one
                                                // more scan round costs a fifth of one more pass
            int lim = kEndingThreshold - 1;     // rejection pairs that may still be accepted: the 100th draw is
                                                // consumed but never accepted (:402-405)
            int rej = search ? 1 : 0;
            bool first_pass = true;
            const unsigned lt = (1u << gl) - 1u;
#pragma unroll 1
            while (__any_sync(kFull, rej)) {
                AGB_ENSURE(rej, 18 + 2 * kScan)
                const uint32_t off18 = (first_pass && nbact) ? 18u : 0u;
                unsigned nmask = 0;
                int n_items = 0, n_nb = 0;
                if (first_pass) {
                    const int cidx0 = nbact ? prev + off_cell : W + 1;
                    bool nb_empty = nbact && (G == 8 || gl < 8) && isEmpty(w[cidx0]);           // empty => on the board
#if AGB_NBFILTER
                    // A cell next to no stone at all cannot kill, save or chase (each needs a string beside it,
                    // :526-563, :565-626, :681-733); the same holds when the only stone beside it is the
                    // opponent's last one and that stone's string has three liberties or more (the move takes
                    // one: two are left, neither a capture nor an atari).  Such cells are not items: early in a
                    // game all eight cells are empty, and items they took were rejection candidates that had to
                    // wait for a second pass (69 % of the second passes).
                    {
                        const int fi = nb_empty ? cidx0 : W + 1;
                        const uint32_t n0 = w[fi - W], n1 = w[fi - 1], n2 = w[fi + W], n3 = w[fi + 1];
                        const uint32_t cb = __byte_perm(__byte_perm(n0, n1, 0x0040), __byte_perm(n2, n3, 0x0040), 0x5410) & 0x03030303u;
                        const int n_stone = __popc((cb + k01) & 0x02020202u);                  // colours 1, 2
                        const uint32_t hp_c = fHp(at(w, offA(w[nbact ? prev : W + 1])));
                        const bool inert = orth ? (n_stone == 1 && hp_c >= 3u) : n_stone == 0;
                        nb_empty = nb_empty && !inert;
                    }
#endif
                    nmask = (__ballot_sync(kFull, nb_empty) >> gbase) & C::GM;
                    if (nb_empty) cbuf[__popc(nmask & lt)] = (uint16_t)((uint32_t)cidx0 | 0x8000u);
                    n_items = n_nb = __popc(nmask);
                }
                // A round adds the empty points of G pairs to the items, in draw order.  The items are the first
                // G entries of the candidate buffer; entry G — the first empty point that found no room — says
                // where the pass ends (further ones are not stored).
                int b = 0, bend = 0;
                const int cap = min(G, n_items + kWant);                // scan until the pass holds this many items
                const int blim = min(kScan, lim);
                int scan = rej;
#pragma unroll 1
                do {
                    const int jp = b + gl;
                    const uint32_t row = mod_small<C::N>(ring[(pos + off18 + 2u * (uint32_t)jp) & RM]);
                    const uint32_t cl = mod_small<C::N>(ring[(pos + off18 + 2u * (uint32_t)jp + 1u) & RM]);
                    const int ridx = (int)(row * W + cl) + (W + 1);
                    const bool can = scan && jp < lim && isEmpty(w[ridx]);
                    const unsigned emask = (__ballot_sync(kFull, can) >> gbase) & C::GM;
                    const int rank = n_items + __popc(emask & lt);
                    if (can && rank <= G) cbuf[rank] = (uint16_t)((uint32_t)ridx | ((uint32_t)jp << 9));
                    n_items += __popc(emask);
                    b += G;
                    if (scan) bend = b;
                    scan = (scan && n_items < cap && b < blim) ? 1 : 0;
                } while (__any_sync(kFull, scan));
                __syncwarp();
                // pairs whose fate this pass settles: those scanned, or those before the first empty one
                // that found no room
                const int covered = n_items > G ? (int)(((uint32_t)cbuf[G] >> 9) & 31u) : bend;
                n_items = min(n_items, G);
                const bool has = gl < n_items;
                const uint32_t item = has ? (uint32_t)cbuf[gl] : (uint32_t)(W + 1);
                const Ev e = eval_point<W, kEvAll>(w, colour, (int)(item & 511u), ko_key);
                const bool ok = has && !e.bad;
                const bool is_nb = (item & 0x8000u) != 0;
                const unsigned goodr = (__ballot_sync(kFull, ok && !is_nb) >> gbase) & C::GM;
                if (first_pass) {
                    // by item: neighbourhood items come first, in cell order
                    const unsigned Kb = (__ballot_sync(kFull, ok && is_nb && e.kill) >> gbase) & C::GM;
                    const unsigned Sb = (__ballot_sync(kFull, ok && is_nb && e.survive) >> gbase) & C::GM;
                    const unsigned Cb = (__ballot_sync(kFull, ok && is_nb && e.chase) >> gbase) & C::GM;
                    if (__any_sync(kFull, (Kb | Sb | Cb) != 0)) {
                        // iteration k of `for (k = 0; k < 9; k++)` consumes draws 2k, 2k+1; lane k looks at
                        // it (G == 8: the ninth iteration is looked at by every lane of the group)
                        bool hit; int pt;
                        {
                            const uint32_t ra = mod_small<3>(ring[(pos + 2u * gl) & RM]);
                            const uint32_t rb = mod_small<3>(ring[(pos + 2u * gl + 1u) & RM]);
                            const uint32_t c9 = ra * 3u + rb;
                            const uint32_t cx = c9 - (c9 > 4u);
                            const unsigned m = Kb | (gl < 6 ? Sb : 0u) | (gl < 4 ? Cb : 0u);
                            const unsigned ip = (unsigned)__popc(nmask & ((1u << cx) - 1u));        // item of cell cx
                            hit = gl < 9 && c9 != 4u && ((nmask >> cx) & 1u) && ((m >> ip) & 1u);   // the centre is occupied
                            pt = prev + (int)(ra * W + rb) - (W + 1);
                        }
                        unsigned hm = (__ballot_sync(kFull, hit) >> gbase) & C::GM;
                        int pt8 = -1;
                        // (the ninth iteration only matters to a playout whose first eight accepted nothing
                        // and that has a kill cell: 23 % of the playouts that get here)
                        if (G == 8 && __any_sync(kFull, hm == 0 && Kb != 0)) {
                            const uint32_t ra = mod_small<3>(ring[(pos + 16u) & RM]);
                            const uint32_t rb = mod_small<3>(ring[(pos + 17u) & RM]);
                            const uint32_t c9 = ra * 3u + rb;
                            const uint32_t cx = c9 - (c9 > 4u);
                            const unsigned ip = (unsigned)__popc(nmask & ((1u << cx) - 1u));
                            if (c9 != 4u && ((nmask >> cx) & 1u) && ((Kb >> ip) & 1u)) hm |= 0x100u;
                            pt8 = prev + (int)(ra * W + rb) - (W + 1);
                        }
                        const int k = __ffs(hm) - 1;
                        const int ptk = __shfl_sync(kFull, pt, gbase + (k & (G - 1)));
                        if (nbact && hm) {
                            pos += 2u * (uint32_t)(k + 1);
                            r = (G == 8 && k == 8) ? pt8 : ptk;
                            found = true;
                            rej = 0;
                        }
                    }
                }
                const int gp = __ffs(goodr) - 1;
                const uint32_t win = __shfl_sync(kFull, item, gbase + (gp & (G - 1)));
                if (rej) {
                    pos += off18;       // nine iterations, two draws each, none accepted
                    if (goodr) {
                        r = (int)(win & 511u);
                        found = true;
                        rej = 0;
                        pos += 2u * (((win >> 9) & 31u) + 1u);
                    } else {
                        const int used = min(covered, lim + 1);                // pairs consumed by this pass
                        pos += 2u * (uint32_t)used;
                        lim -= used;
                        if (lim < 0) { ending = 1; cplx = 1; rej = 0; }
                    }
                }
                first_pass = false;
            }
        }
        // complex mode, :413-478: all 32 lanes for one group at a time
        {
            unsigned cm = __ballot_sync(kFull, live && cplx);
            if (cm) {
                AGB_ENSURE(cplx, 1)
#pragma unroll 1
                while (cm) {
                    const int X = (__ffs(cm) - 1) / G;
                    const int src = X * G;
                    cm &= ~(C::GM << src);
                    uint32_t* const xsm = smem_warp + X * C::words;
                    uint32_t* const xw = xsm + C::oBoard;
                    uint16_t* const xel = elist_warp + X * C::E;
                    int xn = __shfl_sync(kFull, n_el, src);
                    int xd = __shfl_sync(kFull, n_dead, src);
                    const int xstones = __shfl_sync(kFull, stones, src);
                    const int xko = __shfl_sync(kFull, ko_key, src);
                    const uint32_t xcol = __shfl_sync(kFull, colour, src);
                    const uint32_t xpos = __shfl_sync(kFull, pos, src);
                    const int spaces = C::NN - xstones;
                    int pick = -1, used = 0;
                    if (spaces != 0) {
                        // Complex mode is sticky (game_ending, :348-351): the array is made dense once
                        // and then kept current by put (zeroes the entry) and the captures (append).
                        if (xd < 0 || xd >= 8) { xn = compact<C>(xw, xel, xn, lane); xd = 0; }
                        const uint32_t draw = reinterpret_cast<const uint16_t*>(xsm + C::oRing)[xpos & RM];
                        pick = complex_pick<C>(xw, xel, xn, spaces, xcol, xko, draw, lane, &used);
                    }
                    if (grp == X) { r = pick; n_el = xn; n_dead = xd; pos += (uint32_t)used; }
                }
            }
        }

        // ---- Board::put, Board.cpp:68-177, or Board::pass, :200-203 ----
        const bool doput = live && r >= 0;
        if (live) ko_key = -1;
        if (__any_sync(kFull, doput)) {
            const int idx = doput ? r : W + 1;
            if (doput) { stones++; mv++; }
            // Nothing is unlinked from the empty list: the entry of the point dies by itself; only
            // while complex mode lasts it is zeroed, so that the array has no stale entries.
            {
                const bool maint = doput && n_dead >= 0;
                if (__any_sync(kFull, maint)) {
                    if (maint) {
                        if (gl == 0) est(elist + fSlot(w[idx]), 0u);
                        n_dead++;
                    }
                    __syncwarp();
                }
            }
            // lane d (mod 4) looks at neighbour d (up, left, down, right): colour, string, liberties
            const uint32_t v = w[idx + off_nb];
            const uint32_t root = fA(v);
            const uint32_t rw = w[root];
            const uint32_t col = fCol(v);
            const bool enemy = doput && col == 3u - colour, frnd = doput && col == colour;
            const unsigned same = __match_any_sync(kFull, root) & qmask;       // directions seeing the same string
            const int mult = __popc(same);
            const bool firstl = (same & ((1u << lane) - 1u)) == 0;
            const bool lastl = (same >> lane) == 1u;
            const int left = (int)fHp(rw) - mult;
            const int hp_self = __popc(__ballot_sync(kFull, col == 0) & qmask);               // own liberties, :123-127
            // captured strings, named by the LAST direction that sees them: that is when the reference
            // finds their liberties gone (:129-164), so this is the order in which they leave the board
            unsigned caps = (__ballot_sync(kFull, enemy && left == 0 && lastl) >> qbase) & 15u;
            const unsigned frq = (__ballot_sync(kFull, frnd && firstl) >> qbase) & 15u;        // distinct own strings
            const bool act = (G == 4) || gl < 4;                               // one quad of the group does the stores
            // enemy strings lose one liberty per adjacency (sn->hp -= 1, :130,:138,:146,:154)
            if (act && enemy && firstl) w[root] = rw - (uint32_t)mult * kHpOne;
            // Own strings: everything the four sequential neighbour steps of :129-167 do commutes except
            // the unions, whose net effect is known in closed form.  With F1..Fk the distinct own
            // strings in direction order the reference ends with root Fk, piece order Fk, ..., F1, new
            // stone (each union appends the mover's current string to the neighbour's,
            // SetNode.cpp:78-82), hp = sum(hp_i - mult_i) + own.  The liberties a capture gives back
            // are added afterwards (they commute with all of this).
            uint32_t fk = (uint32_t)idx;
            int hp_total = hp_self;
            if (__any_sync(kFull, frq != 0)) {
                const int top = frq ? 31 - __clz(frq) : 0;
                const uint32_t fk_ = __shfl_sync(kFull, root, qbase + top);
                if (frq) fk = fk_;
                int t = (frnd && firstl) ? left : 0;
                t += __shfl_xor_sync(kFull, t, 1);
                t += __shfl_xor_sync(kFull, t, 2);
                hp_total += t;
                // the last stone of this lane's string: the root itself, or what its second stone says
                // (only an own stone has a next stone: on an empty point that field is its slot in elist)
                const uint32_t nx = frnd ? fNext(rw) : kNilG;
                const uint32_t sw = w[nx != kNilG ? nx : root];
                const uint32_t tail = nx != kNilG ? fHp(sw) : root;
                // lane of F_i links its tail to the head of F_(i-1), the lane of F_1 to the new stone
                const unsigned lower = frq & ((1u << dir) - 1u);
                const uint32_t below = __shfl_sync(kFull, root, qbase + (lower ? 31 - __clz(lower) : 0));
                const uint32_t target = lower ? below : (uint32_t)idx;
                const bool own = act && frnd && firstl;
                if (own) {
                    const uint32_t tw = w[tail];
                    w[tail] = setNext(tw, target);
                }
                // strings F_1..F_(k-1) are relabelled to Fk (:284-289), each by its own lane
                // (8 % of the kernel's instructions were this loop: walks are 15 stones long on average, so
                // four steps per vote; the flag is an int — a bool that lives across the loop is kept as a
                // byte and costs three PRMT per test — and the cursor is the byte offset of the stone's word)
                int walk = (own && root != fk) ? 1 : 0;
                uint32_t cur = 4u * root;
                const uint32_t tail4 = 4u * tail, fkA = fk << gw::kAShift;
#pragma unroll 1
                while (__any_sync(kFull, walk != 0)) {
#pragma unroll
                    for (int u = 0; u < AGB_WALK_UNROLL; u++)
                        if (walk) {
                            const uint32_t cw = at(w, cur);
                            *reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(w) + cur) = (cw & ~kAMask) | fkA;
                            walk = cur != tail4 ? 1 : 0;
                            cur = (cw >> (gw::kNextShift - 2)) & (511u << 2);
                        }
                }
                __syncwarp();
                if (own && root == fk) {
                    const uint32_t rw2 = w[fk];
                    w[fk] = setF(rw2, (uint32_t)hp_total);
                    // the new stone is the last one; the second stone of Fk says so
                    const uint32_t sec = nx != kNilG ? nx : target;
                    if (sec != (uint32_t)idx) {
                        const uint32_t x = w[sec];
                        w[sec] = setF(x, (uint32_t)idx);
                    }
                }
            }
            // SetNode::init (SetNode.cpp:31-44) for a stone on its own, else the last stone of Fk
            if (doput && gl == 0)
                w[idx] = gw::stone(fk, kNilG, frq ? (uint32_t)idx : (uint32_t)hp_self, colour);
            __syncwarp();
            // Board::killSetNode, :293-339: one captured stone per group and step; lanes 0..3 of the
            // group give the liberty back to the strings around it (increments commute: atomics)
            if (__any_sync(kFull, caps != 0)) {
                uint32_t kcur = kNilG;
                int single = 0;
#pragma unroll 1
                for (;;) {
                    const bool want = kcur == kNilG && caps != 0;
                    if (__any_sync(kFull, want)) {
                        const uint32_t victim = __shfl_sync(kFull, root, qbase + (caps ? __ffs(caps) - 1 : 0));
                        if (want) {
                            caps &= caps - 1u;
                            kcur = victim;
                            single = fNext(w[victim]) == kNilG ? 1 : 0;
                        }
                    }
                    if (!__any_sync(kFull, kcur != kNilG)) break;
                    // the array is full: drop its stale entries (all 32 lanes per group)
                    unsigned full = __ballot_sync(kFull, kcur != kNilG && n_el == C::E);
#pragma unroll 1
                    while (full) {
                        __syncwarp();       // the stones this loop has taken off so far
                        const int X = (__ffs(full) - 1) / G;
                        const int src = X * G;
                        full &= ~(C::GM << src);
                        uint32_t* const xsm = smem_warp + X * C::words;
                        const int xn = compact<C>(xsm + C::oBoard, elist_warp + X * C::E,
                                                  __shfl_sync(kFull, n_el, src), lane);
                        if (grp == X) { n_el = xn; if (n_dead > 0) n_dead = 0; }
                    }
                    // (the group's first lane rewrites the stone's word below: the link is read by it
                    // alone, before that, and handed to the others.  Nothing else written here is
                    // read before the loop ends: the neighbours looked at are tested for the mover's
                    // colour, which neither a captured stone nor the empty point it becomes has.)
                    const uint32_t nxk = __shfl_sync(kFull, fNext(w[kcur != kNilG ? kcur : 0]), gbase);
                    if (kcur != kNilG) {
                        const uint32_t vn = w[(int)kcur + off_nb];              // one liberty back, :323-326
                        if (act && fCol(vn) == colour) atomicAdd(&w[fA(vn)], kHpOne);
                        stones--;
                        if (gl == 0) {
                            est(elist + n_el, kcur);                            // push at the list head, :311-321
                            w[kcur] = gw::empty(kcur, (uint32_t)n_el);          // empty, slot = n_el
                        }
                        n_el++;
                        if (single) ko_key = (int)(((3u - colour) << 16) | kcur);   // :328-333
                        kcur = nxk;
                    }
                    __syncwarp();   // one latch (convergence note)
                }
                __syncwarp();
            }
        }

        // ---- AMAF: mark[r] / mark2[r] = cstep of the FIRST play of that side at r (:304,:319) ----
        if (C::AMAF) {
            const bool logging = doput && cstep <= kAmafRange;      // only cstep <= AMAF_RANGE is ever used
            if (__any_sync(kFull, logging)) {
                const uint32_t key = (uint32_t)(r < 0 ? 0 : r) | (colour == amaf_agent ? 512u : 0u);
                bool dup = false;
#pragma unroll 1
                for (int b0 = 0; __any_sync(kFull, logging && b0 < n_log); b0 += G) {
                    const uint32_t e = b0 + gl < n_log ? amaf_log[b0 + gl] : 0xffffu;
                    const unsigned hitm = (__ballot_sync(kFull, (e & 1023u) == key) >> gbase) & C::GM;      // (no `||`: every lane votes)
                    dup = dup || hitm != 0;
                }
                if (logging && !dup) {
                    if (gl == 0) amaf_log[n_log] = (uint16_t)(key | ((uint32_t)cstep << 10));
                    n_log++;
                }
                __syncwarp();
            }
        }

        // ---- pair loop of MyWorker::work (AgentGo.cpp:295-325) / testAvrgWin (:422-443) ----
        // The move that closes a pair (the second mover's) ends the playout when neither side moved:
        // the first mover's move is what prev still holds, -1 for a pass.
        if (live) {
            const int before = prev;
            prev = r;                                                           // Board.cpp:171 / :202
            if (closes) {
                if ((before & r) < 0) phase = kEnded;
                else if ((int)mv > a.max_moves) { fault = 1; phase = kEnded; }   // fault detector, once per pair
                if (C::AMAF) cstep++;
            }
            colour = 3u - colour;
            closes ^= 1;
        }
        // (see the convergence note at the top of the file: one latch for the loop)
        __syncwarp();
#undef AGB_ENSURE
    }
}

}  // namespace grp
}  // namespace agb

#endif  // AGB_PLAYOUT_GROUP_CUH_
