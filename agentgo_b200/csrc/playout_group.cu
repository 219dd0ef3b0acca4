// playout_group.cu — kernel wrapper and launcher of the group kernel (playout_group.cuh): 32/G
// playouts per warp, G lanes each.
#include "playout_group.cuh"

namespace agb {

namespace {

template <class C, int WARPS, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(WARPS * 32, BLOCKS_PER_SM) playout_group_kernel(const KernelArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // behind the playouts: the jump matrix T^BATCH of the refills (8 KB, read by every warp of the block)
    uint4* tab = reinterpret_cast<uint4*>(smem + (size_t)WARPS * C::warp_words);
    for (int i = threadIdx.x; i < 32 * 16; i += WARPS * 32) tab[i] = a.jump_batch->e[i];
    __syncthreads();
    grp::playout_group_body<C>(a, smem + (size_t)warp * C::warp_words, lane, (int)blockIdx.x * WARPS + warp, tab);
}

template <int N, int G, int WARPS, int BLOCKS_PER_SM, bool AMAF = false>
cudaError_t launch_g(const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    typedef grp::GCfg<N, G, kJumpSteps, AMAF> C;
    static_assert(WARPS * BLOCKS_PER_SM * C::NG <= kGroupMaxPlayoutsPerSm, "elist_scratch is sized for 128 playouts per SM");
    static_assert(C::warp_words % 4 == 0, "the jump table behind the playouts is 16-byte aligned");
    const size_t smem = (size_t)WARPS * C::warp_words * sizeof(uint32_t) + sizeof(JumpTable);
    cudaError_t e = cudaFuncSetAttribute(playout_group_kernel<C, WARPS, BLOCKS_PER_SM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t per_block = (int64_t)WARPS * C::NG;
    const int64_t want = (a.local_units + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count * BLOCKS_PER_SM;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (info) { info->grid = grid; info->block = WARPS * 32; info->smem = smem; info->name = "playout_group_kernel"; }
    playout_group_kernel<C, WARPS, BLOCKS_PER_SM><<<grid, WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_playout_group(int n, int lanes, const KernelArgs& a, int sm_count, cudaStream_t s, LaunchInfo* info) {
    // 8 lanes per playout: 100 playouts per SM (2 224 B of shared memory each at 19x19) in one block of
    // 25 warps, 80 registers; 16 lanes: 64 playouts in 32 warps of 64 registers; 32 lanes (the mapping
    // of playout_warp.cu run through this code, for comparison): 32 playouts in 32 warps.
    if (lanes == 8 && a.amaf) {
        // with the AMAF log (160 B per playout): 84 playouts per SM, 21 warps of 96 registers
        switch (n) {
            case 9: return launch_g<9, 8, 21, 1, true>(a, sm_count, s, info);
            case 13: return launch_g<13, 8, 21, 1, true>(a, sm_count, s, info);
            case 19: return launch_g<19, 8, 21, 1, true>(a, sm_count, s, info);
        }
    } else if (a.amaf) {
        return cudaErrorInvalidValue;       // the marks are recorded with 8 lanes per playout (or by the warp kernel)
    } else if (lanes == 8) {
        switch (n) {
            case 9: return launch_g<9, 8, 25, 1>(a, sm_count, s, info);
            case 13: return launch_g<13, 8, 25, 1>(a, sm_count, s, info);
            case 19: return launch_g<19, 8, 25, 1>(a, sm_count, s, info);
        }
    } else if (lanes == 16) {
        switch (n) {
            case 9: return launch_g<9, 16, 8, 4>(a, sm_count, s, info);
            case 13: return launch_g<13, 16, 8, 4>(a, sm_count, s, info);
            case 19: return launch_g<19, 16, 8, 4>(a, sm_count, s, info);
        }
    } else if (lanes == 32) {
        switch (n) {
            case 9: return launch_g<9, 32, 8, 4>(a, sm_count, s, info);
            case 13: return launch_g<13, 32, 8, 4>(a, sm_count, s, info);
            case 19: return launch_g<19, 32, 8, 4>(a, sm_count, s, info);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace agb
