/*
 * agentgo_b200.h — C ABI of the B200-native flat Monte-Carlo playout engine.
 *
 * This library is a drop-in for ONE path of Menooker/AgentGo: the random-playout
 * evaluation behind GTP `genmove`.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 *   reference                                           replaced by
 *   --------------------------------------------------  ---------------------------
 *   Scheduler<MyWorker>(true)      TinyMT.h:733         agb_create
 *   ~Scheduler                     TinyMT.h:739         agb_destroy
 *   GTPAdapter::bd.clear()         GTPAdapter.cpp:151   agb_clear_board
 *   onBoardSize(int)               GTPAdapter.h:16      agb_boardsize
 *   bd.put(colour,a,b)             GTPAdapter.cpp:78,92 agb_play
 *   Board::pass                    Board.cpp:200        agb_pass
 *   submit(MyJob)+go()+wait()      AgentGo.cpp:595-605  agb_playouts   (C >= 1)
 *   MyWorker::work playout loop    AgentGo.cpp:285-327  agb_playouts   (C >= 1)
 *   MyGame::testAvrgWin            AgentGo.cpp:412-448  agb_playouts   (C == 0, "root mode")
 *   MyGame::onMove (step>4 branch) AgentGo.cpp:573-688  agb_genmove
 *   MyGame::onMove (opening book)  AgentGo.cpp:468-572  agb_genmove   (cfg.genmove_book)
 *   MyWorker::work static marks    AgentGo.cpp:104-282  agb_static_units, agb_genmove (cfg.genmove_static)
 *   Board::check*                  Board.cpp:480-755    agb_check
 *   Board::countScore              Board.cpp:910        agb_count_score
 *   CalcResults (match scorer)     GeneHatchery.cpp:393 agb_calc_result
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions across the boundary;
 *   - every call returns AGB_OK (0) or a negative agb_status; agb_last_error()
 *     gives the text of the last failure on that engine;
 *   - an engine is NOT thread-safe: one caller thread, like the reference's
 *     single GTP thread;
 *   - the caller owns every buffer it passes in; results are complete when the
 *     call returns (the stream is synchronised inside);
 *   - colours: 0 empty, 1 black, 2 white (GoContest/AgentGo.h:11-13);
 *     a point is (row, col), idx = row*N + col (Board.cpp:100);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with AGB_ERR_CUDA.
 *
 * Randomness.  The reference draws from libc rand(); this
 * build injects one TinyMT32 generator (RFC 8682) per playout, in both the
 * oracle and the kernels:  rand() := tinymt32_uint32() >> 17   (15 bit, the
 * range of the reference platform's RAND_MAX = 0x7fff).  The per-playout seed
 * is the pure function agb_playout_seed() below, so the partition of playouts
 * over threads, blocks or GPUs can never change a result.
 */
#ifndef AGENTGO_B200_H_
#define AGENTGO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGB_VERSION_MAJOR 0
#define AGB_VERSION_MINOR 1

#define AGB_MAX_BOARD_SIZE 19
#define AGB_MAX_POINTS (AGB_MAX_BOARD_SIZE * AGB_MAX_BOARD_SIZE)

#define AGB_EMPTY 0
#define AGB_BLACK 1
#define AGB_WHITE 2

/* TinyMT32 parameter set of RFC 8682 */
#define AGB_TINYMT_MAT1 0x8f7011eeu
#define AGB_TINYMT_MAT2 0xfc78ff1fu
#define AGB_TINYMT_TMAT 0x3793fdffu

/* candidate index used in the seed tuple by root-mode playouts (C == 0) */
#define AGB_ROOT_CANDIDATE 0xffffffffu

typedef enum agb_status {
    AGB_OK = 0,
    AGB_ERR_INVALID = -1,   /* bad argument (size, colour, coordinates, NULL) */
    AGB_ERR_OCCUPIED = -2,  /* agb_play on an occupied point (reference: AG_PANIC, Board.cpp:75-78) */
    AGB_ERR_CUDA = -3,      /* CUDA runtime error or no device */
    AGB_ERR_NOMEM = -4,
    AGB_ERR_FAULT = -5,     /* a playout hit the per-playout move cap (fault detector only) */
    AGB_ERR_STATE = -6      /* call not valid in this state */
} agb_status;

/* kernel selector: the mapping of playouts onto the machine */
typedef enum agb_kernel {
    AGB_KERNEL_AUTO = 0,    /* per launch: GROUP8 from about one load of the machine up (110 playouts per SM,
                               boards larger than 9x9), else WARP (lower latency per playout) */
    AGB_KERNEL_THREAD = 1,  /* one thread per playout */
    AGB_KERNEL_WARP = 2,    /* one warp per playout   */
    AGB_KERNEL_GROUP8 = 3,  /* four playouts per warp, 8 lanes each  */
    AGB_KERNEL_GROUP16 = 4, /* two playouts per warp, 16 lanes each  */
    AGB_KERNEL_GROUP32 = 5  /* the group kernel's code with one playout per warp (comparison) */
} agb_kernel;

typedef struct agb_config {
    int board_size;      /* 9, 13 or 19 */
    int device;          /* CUDA device ordinal; -1 => position-only engine (compute calls fail) */
    int kernel;          /* agb_kernel */
    int shard_rank;      /* this engine evaluates global playout ids g with        */
    int shard_count;     /*   g % shard_count == shard_rank   (1 => everything)     */
    uint32_t mat1, mat2, tmat; /* TinyMT32 parameters; 0,0,0 => RFC 8682 set      */
    int max_moves;       /* per-playout move cap, 0 => default (fault detector)    */
    int genmove_amaf;    /* agb_genmove: 0 flat MC only (default); 1 adds the AMAF terms
                            amaf1 + amaf2 of AgentGo.cpp:329-338,654-655                 */
    int genmove_static;  /* agb_genmove: 1 adds the static pattern marks of MyWorker::work
                            (AgentGo.cpp:104-282, agb_static_units) to every candidate       */
    int genmove_book;    /* agb_genmove: 1 answers the first four calls after a clear_board from
                            the reference's 13x13 opening book (AgentGo.cpp:468-572); ignored on
                            other board sizes (the book is made of 13x13 literals)            */
    /* The eight tunables of the reference's command line (AgentGo.cpp:18-25 defaults, :709-719
     * argv[1..8]; the parameters its GA tuner searches).  agb_default_config sets the defaults.
     *   index_a, index_b      static pattern marks: bonus_a = int(P*index_a), bonus_b = int(P*index_b) (:95-96)
     *   index_c               scale of every playout term of the mark (:332,335,339,340)
     *   index_amaf_a1..b2     AMAF decay table: 1 - (k / (a1 - a2*stones)) ^ (b1 - b2*stones)   (:87-94)
     *   index_cnsv            cnsv = 2 ^ (avrg_win * index_cnsv), the weight of losing playouts (:99) */
    double index_a, index_b, index_c;
    double index_amaf_a1, index_amaf_a2, index_amaf_b1, index_amaf_b2;
    double index_cnsv;
} agb_config;
/* amaf + static + book = the complete MyGame::onMove of the reference */

typedef struct agb_engine agb_engine;

/* statistics of the last agb_playouts* call */
typedef struct agb_stats {
    float device_ms;        /* CUDA-event time of the kernel(s) on the engine's stream */
    int64_t playouts;       /* playouts this engine ran (its shard)                    */
    int64_t moves;          /* stones placed in those playouts                         */
    int64_t rng_draws;      /* rand() calls consumed                                   */
    int32_t kernel_launches;
    int32_t faults;         /* playouts stopped by the move cap                        */
} agb_stats;

/* one position of a batch: a move list replayed from the empty board.
 * moves[i] = colour<<16 | row<<8 | col ; row = col = 0xff encodes pass. */
typedef struct agb_position {
    const uint32_t* moves;
    int32_t n_moves;
    int32_t to_move;        /* colour whose root-mode playouts are evaluated */
} agb_position;

#define AGB_MOVE(colour, row, col) \
    (((uint32_t)(colour) << 16) | ((uint32_t)((row) & 0xff) << 8) | (uint32_t)((col) & 0xff))
#define AGB_MOVE_PASS(colour) AGB_MOVE(colour, 0xff, 0xff)

/* ---- life cycle ------------------------------------------------------- */
void agb_default_config(agb_config* cfg);
int agb_create(const agb_config* cfg, agb_engine** out);
void agb_destroy(agb_engine* h);
const char* agb_last_error(const agb_engine* h);
const char* agb_version(void);
int agb_device_count(void);

/* ---- position (host-side replay into the exact reference state) ------- */
int agb_boardsize(agb_engine* h, int n);             /* also clears the board */
int agb_clear_board(agb_engine* h);
int agb_play(agb_engine* h, int colour, int row, int col);
int agb_pass(agb_engine* h, int colour);
int agb_get_board(const agb_engine* h, uint8_t* colours /* N*N */, int* n);

/* which: 0 suicide, 1 kill, 2 survive, 3 dying, 4 chase, 5 true eye, 6 compete (ko),
 *        7 no-sense.   returns 0/1, or <0 on error.                          */
int agb_check(const agb_engine* h, int which, int colour, int row, int col);
int agb_count_score(const agb_engine* h, int colour);

/* Final score of a finished game the way the reference's match harness counts it (CalcResults,
 * GeneHatchery.cpp:379-436): on a copy of the board, sweep the points in row-major order; an empty
 * point takes the colour of the first stone among its up, down, left, right neighbours (values
 * written earlier in the same sweep count); repeat until a sweep changes nothing; then
 * *black_minus_white = #black - #(everything else).  Host only.                              */
int agb_calc_result(const agb_engine* h, int* black_minus_white);

/* Board::getRandomPiece(colour) (Board.cpp:341-411) on the engine's board, on the HOST, drawing
 * from a TinyMT32 stream seeded with `seed` (engine's parameter set; rand() := u32 >> 17).  This is
 * the move the reference's match referee falls back to when an engine passes
 * (GeneHatchery.cpp:474,520) and its commented-out "play random" mode (AgentGo.cpp:453-460).  The
 * move is NOT played; like the reference's call it may latch the board's game_ending flag
 * (Board.cpp:348-351,407).  *is_pass = 1 when the policy finds no move.                       */
int agb_random_piece(agb_engine* h, int colour, uint32_t seed, int* row, int* col, int* is_pass);

/* debugging / parity: dump the full compact state as int32 words;
 * returns the number of words written, or <0.                               */
int agb_export_state(const agb_engine* h, int32_t* out, int cap);

/* ---- the hot path ------------------------------------------------------
 * colour   : the side the evaluation is for ("agent" = isWh+1, AgentGo.cpp:78)
 * cand_rc  : C pairs (row, col).  C >= 1: for each candidate, play it for `colour`,
 *            then P playouts with the ENEMY moving first (AgentGo.cpp:300,310).
 *            C == 0 (cand_rc may be NULL): root mode, P playouts with `colour`
 *            moving first (AgentGo.cpp:425,434); outputs have one element.
 * P        : playouts per candidate.  Global playout id g = cand*P + p; this
 *            engine runs those with g % shard_count == shard_rank.
 * avrg_win : threshold subtracted from every diff (AgentGo.cpp:327).
 * seed_base: 64-bit base of the per-playout seed tuple (agb_playout_seed).
 * position_id : enters the seed tuple (0 for the engine's own board).
 * wins/sum_pos/sum_neg [max(C,1)] : with w = diff - avrg_win per playout,
 *            wins = #{w >= 0}, sum_pos = sum max(w,0), sum_neg = sum min(w,0)
 *            — this engine's SHARD only (sum them over shards; agb_counters_dev
 *            exposes the same numbers on the device for an NCCL allreduce).
 * per_playout_diff [max(C,1)*P] or NULL : diff = countScore(colour) -
 *            countScore(enemy) of every playout of this shard; entries of
 *            other shards are left untouched.
 */
int agb_playouts(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P,
                 int avrg_win, uint64_t seed_base, uint32_t position_id,
                 int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                 int16_t* per_playout_diff, agb_stats* stats);

/* Asynchronous halves of agb_playouts for callers that own the collective:
 * launch leaves the 3*max(C,1) int64 counters [wins | sum_pos | sum_neg] in
 * device memory (agb_counters_dev), finish synchronises and copies out.      */
int agb_playouts_launch(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P,
                        int avrg_win, uint64_t seed_base, uint32_t position_id,
                        int want_diffs);
int agb_playouts_finish(agb_engine* h, int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                        int16_t* per_playout_diff, agb_stats* stats);
/* device pointer (int64[3*max(C,1)]) and CUDA stream (cudaStream_t as void*) */
void* agb_counters_dev(agb_engine* h);
void* agb_stream(agb_engine* h);

/* Optional collective hook for sharded engines (shard_count > 1).  When set, every
 * agb_playouts*_finish / agb_genmove step calls fn(ctx, dev_int64, count, cuda_stream)
 * on the counters before they are copied out; fn must sum them over all shards in place
 * (one NCCL allreduce, ncclInt64/ncclSum, enqueued on that stream or ordered after it)
 * and return 0.  Without it a sharded engine returns its own shard's partial counts.    */
typedef int (*agb_allreduce_fn)(void* ctx, void* dev_int64, int count, void* cuda_stream);
int agb_set_allreduce(agb_engine* h, agb_allreduce_fn fn, void* ctx);
/* The collective of the batch path (agb_playouts_batch on sharded engines): positions are sharded,
 * every shard fills a table of count_per_shard int64 and fn must gather the tables of all shards,
 * in shard order, into recv_dev_int64 (one NCCL allgather, ncclInt64, on that stream) and return 0.
 * With it every shard's agb_playouts_batch returns the counters of ALL positions; without it a
 * shard returns its own positions and zeros elsewhere (last sentence).           */
typedef int (*agb_allgather_fn)(void* ctx, const void* send_dev_int64, void* recv_dev_int64,
                                int count_per_shard, void* cuda_stream);
int agb_set_allgather(agb_engine* h, agb_allgather_fn fn, void* ctx);

/* Single-process multi-GPU: engines[i] must be shard i of n on its own device.  Builds one NCCL
 * communicator (ncclCommInitAll; libnccl is dlopen'ed at this point) and installs the allreduce
 * and allgather hooks on every engine.  Afterwards issue the SAME agb_playouts* / agb_genmove call on every engine
 * concurrently, one host thread per engine: each returns the counters summed over all GPUs.     */
typedef struct agb_nccl_group agb_nccl_group;
int agb_attach_nccl(agb_engine** engines, int n, agb_nccl_group** out);
void agb_detach_nccl(agb_nccl_group* g);

/* AMAF first-play marks (AgentGo.cpp:289-290,304,319,329-338), candidate mode (C >= 1) only.
 * Integer contract: amaf[which][sign][step][q], int64, AGB_AMAF_WORDS(N) elements, accumulated
 * over ALL candidates like the reference's global amaf1/amaf2:
 *   which 0: q's first play by `colour` (mark),  1: by the enemy (mark2), counted only when the
 *            agent has no mark <= 40 at q (the `else if` of :334);
 *   sign  0: sum of w over playouts with w >= 0,  1: sum of w over playouts with w < 0
 *            (w = diff - avrg_win);
 *   step  = cstep of that first play, 2..AGB_AMAF_RANGE (0 and 1 stay zero);  q = row*N + col.
 * The reference's doubles follow on the host:
 *   amaf1[q] =  sum_s 100*amaf[s-1]*index_c*(A[0][0][s][q] + cnsv*A[0][1][s][q])
 *   amaf2[q] = -sum_s 100*amaf[s-1]*index_c*(A[1][0][s][q] + cnsv*A[1][1][s][q])          */
#define AGB_AMAF_RANGE 40   /* GoContest/AgentGo.h:16 */
#define AGB_AMAF_WORDS(n) (2 * 2 * (AGB_AMAF_RANGE + 1) * (n) * (n))
int agb_playouts_amaf(agb_engine* h, int colour, const int16_t* cand_rc, int C, int64_t P,
                      int avrg_win, uint64_t seed_base, uint32_t position_id,
                      int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                      int64_t* amaf /* AGB_AMAF_WORDS(N) */, agb_stats* stats);

/* batch of B independent positions (config 3): root-mode, P playouts each,
 * position b uses position_id = first_position_id + b and is evaluated by the
 * shard with (first_position_id + b) % shard_count == shard_rank.
 * outputs are [B]; per_playout_diff is [B*P] or NULL.  With an allgather hook
 * (agb_set_allgather / agb_attach_nccl) every shard returns the counters of all B
 * positions; per_playout_diff stays local to the shard that played them.      */
int agb_playouts_batch(agb_engine* h, const agb_position* pos, int B, int64_t P,
                       uint64_t seed_base, uint32_t first_position_id,
                       int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                       int16_t* per_playout_diff, agb_stats* stats);

/* GTP-compat genmove (MyGame::onMove, AgentGo.cpp:451-692): root playouts -> avrg_win,
 * candidate filter, P playouts per candidate, argmax of total_mark =
 * 2*100*(sum_pos + cnsv*sum_neg), plus amaf1 + amaf2 when cfg.genmove_amaf, plus the static
 * pattern marks when cfg.genmove_static; the first four calls come from the opening book when
 * cfg.genmove_book (13x13).  With all three set and P = 150 / 500 this is the reference's move
 * (tests/test_gpu_genmove_ref.py checks it against the reference's own onMove).  *is_pass = 1
 * when no candidate exists.  The move
 * is NOT played on the engine's board (the caller does, GTPAdapter.cpp:78).
 * P_per_candidate < 0: |P_per_candidate| is a fixed playout budget per genmove,
 * split evenly (rounded up) over the candidates.                              */
int agb_genmove(agb_engine* h, int colour, int64_t P_per_candidate, int64_t P_root,
                uint64_t seed_base, int* row, int* col, int* is_pass, agb_stats* stats);

/* Static pattern marks of MyWorker::work (AgentGo.cpp:104-282) for one candidate of `colour` on
 * the engine's board, as the two integers of
 *     total_mark[row][col] += units_a * bonus_a + units_b * bonus_b,
 *     bonus_a = int(numset * index_a), bonus_b = int(numset * index_b)       (AgentGo.cpp:95-96,
 *     index_a = 1000, index_b = 1 at :18-19; numset = playouts per candidate).
 * units_a: -1 for checkDying, -1 for checkNoSense (:105-107).  units_b: +1 checkSurvive (:109), the
 * contact / diagonal / knight's-move terms (:111-240) and 20 per cut pattern (:241-282).  The
 * reference's quirks are part of the contract: "friend" is tested as data == agent+1 (for black
 * that is a white stone, for white nothing), and two knight's-move tests read a different point
 * than they guard (:185-191, 217-223), past the end of Board::data on the last row.  Host only.  */
int agb_static_units(const agb_engine* h, int colour, int row, int col,
                     int64_t* units_a, int64_t* units_b);

/* total_mark of the last agb_genmove (N*N doubles, row-major; AgentGo.cpp:17,653-655): static
 * marks + 2 x flat-MC term + amaf1 + amaf2 as configured; 0 where nothing was added.  Returns the
 * number of doubles written, 0 when the last genmove was answered from the book, <0 on error.  */
int agb_last_marks(const agb_engine* h, double* total_mark, int cap);
/* number of agb_genmove calls since the last clear_board / boardsize (MyGame::step, AgentGo.cpp:353);
 * agb_set_genmove_step overrides it (a host that replays a game sets it to the move count). */
int agb_genmove_step(const agb_engine* h);
int agb_set_genmove_step(agb_engine* h, int step);

/* ---- diagnostics -------------------------------------------------------
 * Live micro-benchmarks of the two rates that bound the playout kernels:
 * INT32 lane-ops/s (mixed LOP3/IADD3/IMAD, and alu-pipe only) and conflict-free
 * shared-memory bytes/s, at the clock the device actually sustains.              */
int agb_measure_peaks(int device, double* int_ops_per_s, double* int_ops_per_s_alu_only,
                      double* smem_bytes_per_s, int* sm_count);

/* Shared-memory bounds violations counted since load by the AGB_CHECK_BOUNDS build of the
 * kernels (`python -m agentgo_b200.build --debug`); always 0 in the release build.          */
long long agb_debug_violations(void);
/* Coverage counters of the same build: 0 = compactions of the empty-list array forced by a full
 * array, 1 = compactions on entering / during complex mode, 2 = stones placed through the capture
 * path of put.  The checked build also counts, as a violation, any move that breaks the invariant
 * the release kernel relies on (a move of the policy is never a suicide).  0 in the release build. */
long long agb_debug_event(int which);

/* ---- the seed contract (shared verbatim by oracle and kernels) --------- */
static inline uint32_t agb_playout_seed(uint64_t seed_base, uint32_t position_id,
                                        uint32_t candidate_idx, uint64_t playout_idx) {
    uint64_t z = seed_base;
    z += 0x9E3779B97F4A7C15ull * (playout_idx + 1);
    z ^= 0xBF58476D1CE4E5B9ull * ((uint64_t)candidate_idx + 1);
    z += 0x94D049BB133111EBull * ((uint64_t)position_id + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z ^ (z >> 32));
}

#ifdef __cplusplus
}
#endif
#endif /* AGENTGO_B200_H_ */
