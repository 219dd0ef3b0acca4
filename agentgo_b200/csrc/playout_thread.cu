// playout_thread.cu — one THREAD per playout (sm_100a).
//
// The simplest correct mapping of MyWorker::work's playout loop (AgentGo.cpp:285-327) and
// testAvrgWin (AgentGo.cpp:412-448): every thread owns one board in shared memory and runs
// the sequential policy of board_ops.h.  Boards of the 32 lanes of a warp are interleaved
// word by word (stride 32) so lane l only ever touches bank l: no bank conflicts however the
// lanes diverge.  Cost: 2*N*N*128 B of shared memory per warp (92 416 B at 19x19), so only
// 2 warps fit per SM at 19x19 — this variant is occupancy- and divergence-bound and exists
// as the baseline the warp-per-playout kernel is measured against.
#include "kernel_args.h"

namespace agb {

template <int N, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) playout_thread_kernel(const KernelArgs a) {
    extern __shared__ uint32_t smem[];
    constexpr int NN = N * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    BoardRef<32, N> bd;
    bd.w0 = smem + (size_t)warp * (2 * NN * 32) + lane;
    bd.w1 = bd.w0 + NN * 32;
    bd.n_rt = N;

    unsigned long long moves = 0, draws = 0;
    unsigned int faults = 0, played = 0;

    for (;;) {
        const unsigned long long l = atomicAdd(a.work_counter, 1ull);
        if (l >= (unsigned long long)a.local_units) break;
        const int64_t g = (int64_t)l * a.shard_count + a.shard_rank;
        const int s = (int)(g / a.P);
        const int64_t p = g - (int64_t)s * a.P;
        const Position& pos = a.starts[s];
        const StartDesc d = a.descs[s];

        // Board::clone (Board.cpp:216-254): 2*NN words instead of 126 KB
        for (int i = 0; i < NN; i++) {
            bd.w0[i * 32] = pos.w0[i];
            bd.w1[i * 32] = pos.w1[i];
        }
        bd.load_scalars(pos);

        Rng rng;
        rng.seed(playout_seed(a.seed_base, d.position_id, d.cand_idx, (uint64_t)p), a.rng);
        uint32_t mv = 0;
        const int diff = bd.run_playout(rng, d.agent, d.first, a.max_moves, &mv);
        moves += mv;
        draws += rng.draws;
        played++;
        if (a.diffs) a.diffs[g] = (int16_t)diff;
        if (diff == -32768) {
            faults++;
        } else {
            const int w = diff - (a.avrg_dev ? __ldg(a.avrg_dev) : d.avrg_win);       // AgentGo.cpp:327
            unsigned long long* c = (unsigned long long*)a.counters;
            if (w >= 0) {                                    // sign test of AgentGo.cpp:328
                atomicAdd(&c[s], 1ull);
                atomicAdd(&c[a.counter_stride + s], (unsigned long long)w);
            } else {
                atomicAdd(&c[2 * a.counter_stride + s], (unsigned long long)(long long)w);
            }
        }
    }
    atomicAdd(&a.stats[kStatMoves], moves);
    atomicAdd(&a.stats[kStatDraws], draws);
    if (faults) atomicAdd(&a.stats[kStatFaults], (unsigned long long)faults);
    atomicAdd(&a.stats[kStatPlayouts], (unsigned long long)played);
}

template <int N, int WARPS>
static cudaError_t launch_t(const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    const size_t smem = (size_t)WARPS * 2 * N * N * 32 * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(playout_thread_kernel<N, WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t want = (a.local_units + WARPS * 32 - 1) / (WARPS * 32);
    int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    if (info) { info->grid = grid; info->block = WARPS * 32; info->smem = smem; info->name = "playout_thread_kernel"; }
    playout_thread_kernel<N, WARPS><<<grid, WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_playout_thread(int n, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    switch (n) {
        case 9: return launch_t<9, 8>(a, sm_count, s, info);
        case 13: return launch_t<13, 5>(a, sm_count, s, info);
        case 19: return launch_t<19, 2>(a, sm_count, s, info);
    }
    return cudaErrorInvalidValue;
}

}  // namespace agb
