// kernel_args.h — launch contract shared by the playout kernels and the engine.
#ifndef AGB_KERNEL_ARGS_H_
#define AGB_KERNEL_ARGS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "board_ops.h"

namespace agb {

// One launch evaluates `n_starts` start positions x P playouts each.  Unit of work: global
// playout id g = start*P + p; this engine runs the ids with g % shard_count == shard_rank.
// Units are handed out dynamically from `work_counter` (playout length varies 360..600 moves
// at 19x19), which cannot change results: the seed is a pure function of (start, p).
struct KernelArgs {
    const Position* starts;     // [n_starts] device
    const StartDesc* descs;     // [n_starts] device
    int32_t n_starts;
    int64_t P;
    uint64_t seed_base;
    int32_t shard_rank, shard_count;
    int64_t local_units;        // number of ids this shard owns
    unsigned long long* work_counter;   // device, zeroed before launch
    long long* counters;        // device [3][n_starts]: wins | sum_pos | sum_neg
    int16_t* diffs;             // device [n_starts*P] or nullptr
    unsigned long long* stats;  // device [4]: moves, rng draws, faults, playouts
    RngParams rng;
    int32_t max_moves;
};

enum StatSlot { kStatMoves = 0, kStatDraws = 1, kStatFaults = 2, kStatPlayouts = 3, kStatCount = 4 };

struct LaunchInfo {
    int grid, block;
    size_t smem;
    const char* name;
};

// thread-per-playout kernel (playout_thread.cu)
cudaError_t launch_playout_thread(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info);
// warp-per-playout kernel (playout_warp.cu)
cudaError_t launch_playout_warp(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info);

}  // namespace agb

#endif
