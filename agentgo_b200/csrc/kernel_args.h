// kernel_args.h — launch contract shared by the playout kernels and the engine.
#ifndef AGB_KERNEL_ARGS_H_
#define AGB_KERNEL_ARGS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "board_ops.h"

namespace agb {

// Device layout of a start position for the warp kernel: the board is padded to (N+2)x(N+2)
// with a border of colour 3, so no neighbour access needs a bounds check.  Every index stored
// in the state (links, roots, tails, last moves, ko point) is a padded index
// p = (row+1)*(N+2) + (col+1).
//
//   w0 = A | 4B<<16 | colour<<30      (A and 4B are the two 16-bit halves of the word, so links and
//                                      labels are rewritten with 16-bit stores, no read-modify-write;
//                                      B is kept times four = the byte offset of that stone's word)
//        stone:  A = root index of its string,  B = next stone in the string (kNilP = end)
//        empty:  A = slot of the point in `elist`, B = 0
//   w1 = hp | tail<<16                at root stones
//
// The reference's doubly linked list of empty points (Board::space_list; removal Board.cpp:105-118,
// head insertion per captured stone :311-321) is only ever OBSERVED by the ordered walk of
// getRandomPieceComplex (:445-470).  Its order is "captured stones, latest first, then the
// survivors of the initial list", so the device keeps it as an append-only array: list order =
// elist[n_el-1], ..., elist[0] without the dead entries.  An entry k naming cell c is alive iff
// c is empty and A(c) == k; filling a point costs nothing, a capture appends.
constexpr int kMaxW = kMaxN + 2;
constexpr int kMaxNP = kMaxW * kMaxW;
constexpr int kNilP = 511;
constexpr int kElistCap = 512;             // entries in shared memory; compacted when full
struct PaddedPosition {
    uint32_t w0[kMaxNP];
    uint32_t w1[kMaxNP];
    uint32_t elist_w[(kMaxNN + 1) / 2];    // n_el 16-bit entries
    int16_t n_el, stones, ko_key_col, ko_idx, last[2], ending, pad;
};

// Device layout of a start position for the group kernel (playout_group.cuh): ONE word per point of
// a padded board with ONE border column — rows -1 .. n of n + 1 points each, the point right of a row's
// last point being the point left of the next row's first; (n + 1)(n + 2) + 1 words (to_group in
// host_layout.cpp, GCfg in playout_group.cuh) —
//   bits 0-1 colour | bits 2-10 A | bits 11-19 next | bits 21-31 F
//        stone:  A = root of its string, next = next stone of the string (511 = end),
//                F = pseudo-liberties at the root stone / the string's last stone at its SECOND stone
//        empty:  A = the point itself, bits 11-20 = slot of the point in `elist`, F = 1
//        border: A = the point itself, colour 3, F = 1
// (A sits at bit 2 so that `w & 0x7fc` is the byte offset of the root's word; F is at the top so
// that the liberties of a root are one shift away; F = 1 on points without a stone makes "liberties
// minus multiplicity" zero there, which keeps the byte-packed arithmetic of eval_point borrow free)
// and the same append-only array of empty points as PaddedPosition.
namespace gw {
constexpr int kAShift = 2, kNextShift = 11, kFShift = 21;
constexpr uint32_t kOne = 1u << kFShift;
AGB_HD constexpr uint32_t stone(uint32_t root, uint32_t next, uint32_t f, uint32_t col) {
    return col | (root << kAShift) | (next << kNextShift) | (f << kFShift);
}
AGB_HD constexpr uint32_t empty(uint32_t self, uint32_t slot) { return (self << kAShift) | (slot << kNextShift) | kOne; }
AGB_HD constexpr uint32_t border(uint32_t self) { return 3u | (self << kAShift) | kOne; }
}  // namespace gw
// entries of a playout's empty-point array in the group kernel; it lives in GLOBAL memory (L2): it is
// read by complex mode and compaction only, and 752 B of shared memory per playout buy five more
// warps per SM.  1024 = what the 10-bit slot field of an empty point can name.
constexpr int kGroupElistCap = 1024;
constexpr int kGroupMaxPlayoutsPerSm = 128;
struct GroupPosition {
    uint32_t w[kMaxNP];
    uint32_t elist_w[(kMaxNN + 1) / 2];    // n_el 16-bit entries
    int16_t n_el, stones, ko_key_col, ko_idx, last[2], ending, pad;
};

// GF(2) jump-ahead table of TinyMT32: T^k split by state nibble (k = jump_steps(), the number of
// draws each lane generates per ring refill).  e[l*16 + v] is the image of state nibble l (bits
// 4l..4l+3 of s0|s1|s2|s3) having value v; the new state is the XOR of the 32 selected entries.
struct JumpTable {
    uint4 e[32 * 16];
};

// Jump tables of the group kernel.  Lane j of a playout's G lanes generates draws [j*SEG, (j+1)*SEG)
// of every batch of G*SEG draws from a generator state of its own.  Matrix 0 is T^(G*SEG), the step
// from one batch to the next (the same for every lane); matrix j = 1..G-1 is T^(SEG*j), the step from
// the seed state to lane j's first segment.  Each is split by state BYTE: e[(j*16 + b)*256 + v] is
// the image of byte b (bits 8b..8b+7 of s0|s1|s2|s3) having value v; a product is the XOR of the
// 16 selected entries.  No lane waits for another one (the cooperative product of the warp kernel
// needs a reduction over the lanes per step, and redux.sync with a partial member mask compiles to
// a loop).
struct JumpBytes {
    uint4 e[1];     // [G][16][256], allocated with the size for G
};
constexpr size_t jump_bytes_size(int lanes) { return (size_t)lanes * 16 * 256 * sizeof(uint4); }

// One launch evaluates `n_starts` start positions x P playouts each.  Unit of work: global
// playout id g = start*P + p; this engine runs the ids with g % shard_count == shard_rank.
// Units are handed out dynamically from `work_counter` (playout length varies 360..600 moves
// at 19x19), which cannot change results: the seed is a pure function of (start, p).
struct KernelArgs {
    const Position* starts;     // [n_starts] device (thread kernel)
    const PaddedPosition* pstarts;  // [n_starts] device (warp kernel)
    const GroupPosition* gstarts;   // [n_starts] device (group kernel)
    const JumpTable* jump;      // device (warp kernel)
    const uint4* jump_bytes;    // device (group kernel), JumpBytes for the launch's lanes per playout
    const JumpTable* jump_batch; // device (group kernel): T^(lanes*SEG) split by nibble, copied to shared memory
    uint16_t* elist_scratch;    // device (group kernel): kGroupElistCap entries per resident playout
    const StartDesc* descs;     // [n_starts] device
    int32_t n_starts;
    int64_t P;
    uint64_t seed_base;
    int32_t shard_rank, shard_count;
    int64_t local_units;        // number of ids this shard owns
    unsigned long long* work_counter;   // device, zeroed before launch
    long long* counters;        // device [3][counter_stride]: wins | sum_pos | sum_neg
    int32_t counter_stride;     // >= n_starts (the batch path pads every shard's table to the same size)
    int16_t* diffs;             // device [n_starts*P] or nullptr
    long long* amaf;            // device [2 which][2 sign][41 step][N*N] or nullptr (AMAF marks, warp kernel)
    unsigned long long* stats;  // device [4]: moves, rng draws, faults, playouts
    RngParams rng;
    int32_t max_moves;
    const int* avrg_dev;        // device, or nullptr: when set, avrg_win of every start is read from here (genmove:
                                // the root playouts that determine it are still in flight when this launch is built)
};

enum StatSlot { kStatMoves = 0, kStatDraws = 1, kStatFaults = 2, kStatPlayouts = 3, kStatCount = 4 };

struct LaunchInfo {
    int grid, block;
    size_t smem;
    const char* name;
};

// host helpers of the device layouts (host_layout.cpp)
void to_padded(const Position& p, PaddedPosition* out);
void to_group(const Position& p, GroupPosition* out);
constexpr int kJumpSteps = 16;             // draws each lane generates per ring refill, both kernels
int jump_steps();
void build_jump(const RngParams& prm, int steps, JumpTable* out);
void build_jump_bytes(const RngParams& prm, int lanes, int seg, uint4* out);   // jump_bytes_size(lanes) bytes
// bounds violations counted by the AGB_CHECK_BOUNDS build of the warp kernel (0 in release)
unsigned long long debug_violations();
// coverage counters of the same build (see playout_warp.cu); 0 in release
unsigned long long debug_event(int which);

// thread-per-playout kernel (playout_thread.cu)
cudaError_t launch_playout_thread(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info);
// warp-per-playout kernel (playout_warp.cu)
cudaError_t launch_playout_warp(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info);
// several playouts per warp, `lanes` (8, 16 or 32) lanes each (playout_group.cu)
cudaError_t launch_playout_group(int n, int lanes, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info);

}  // namespace agb

#endif
