// diag.cu — live micro-benchmarks of the two machine rates that bound the playout kernels:
// INT32 issue (LOP3 on the alu pipe + IMAD on the fma pipe) and shared-memory bandwidth.
// the path is SM integer-issue / shared-memory bound, not HBM or tensor, and the
// peaks must be MEASURED at the clock the GPU actually runs at.
#include <cuda_runtime.h>
#include <stdint.h>

#include "agentgo_b200.h"

namespace {

constexpr int kIters = 4096;

// 8 independent chains per thread.  The instruction mix is pinned with inline PTX (the compiler moves
// plain adds between the two pipes as it likes):
//   mixed:    lop3 on the alu pipe and mad.lo on the fma pipe, one to one — each pipe takes a warp
//             instruction every second cycle, together they fill the issue slot of every cycle, which
//             is the integer ceiling of the SM (128 lane-ops per clock);
//   alu only: lop3 alone — the alu pipe's own ceiling (64 lane-ops per clock per SM).
// profiles/r02_peak_kernels.txt holds the ncu capture that shows which pipe each one saturates.
__global__ void __launch_bounds__(1024) int_peak_kernel(uint32_t* out, uint32_t seed, int mixed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 ^ 0x1234u, a3 = a0 + 77u;
    uint32_t b0 = a0 * 5u, b1 = a0 * 7u + 1u, b2 = a0 ^ 0xabcdefu, b3 = a0 + 99u;
    const uint32_t k = seed | 1u;
#define AGB_LOP(d, x, y) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(x), "r"(y))
#define AGB_MAD(d, x, y) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(x), "r"(y))
    if (mixed) {
#pragma unroll 8
        for (int i = 0; i < kIters; i++) {
            AGB_LOP(a0, k, b0); AGB_MAD(b0, k, a0); AGB_LOP(a1, k, b1); AGB_MAD(b1, k, a1);
            AGB_LOP(a2, k, b2); AGB_MAD(b2, k, a2); AGB_LOP(a3, k, b3); AGB_MAD(b3, k, a3);
            AGB_LOP(a0, k, b1); AGB_MAD(b0, k, a1); AGB_LOP(a1, k, b2); AGB_MAD(b1, k, a2);
        }
    } else {
#pragma unroll 8
        for (int i = 0; i < kIters; i++) {
            AGB_LOP(a0, k, b0); AGB_LOP(b0, k, a0); AGB_LOP(a1, k, b1); AGB_LOP(b1, k, a1);
            AGB_LOP(a2, k, b2); AGB_LOP(b2, k, a2); AGB_LOP(a3, k, b3); AGB_LOP(b3, k, a3);
            AGB_LOP(a0, k, b1); AGB_LOP(b0, k, a1); AGB_LOP(a1, k, b2); AGB_LOP(b1, k, a2);
            AGB_LOP(a2, k, b3); AGB_LOP(b2, k, a3); AGB_LOP(a3, k, b0); AGB_LOP(b3, k, a0);
        }
    }
#undef AGB_LOP
#undef AGB_MAD
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

// conflict-free LDS.32: lane l reads word (l + rotating offset) of its warp's 32-word rows
__global__ void __launch_bounds__(1024) smem_peak_kernel(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = (uint32_t)i * seed;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t idx = threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < kIters; i++) {
        acc += buf[idx & 8191u];
        idx += 1024u + 32u;             // stays on bank (lane) — conflict free, independent of acc
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace

extern "C" int agb_measure_peaks(int device, double* int_ops_per_s, double* int_ops_per_s_alu_only,
                                 double* smem_bytes_per_s, int* sm_count) {
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return AGB_ERR_CUDA;
    }
    const int blocks = prop.multiProcessorCount * 2, threads = 1024;
    uint32_t* out = nullptr;
    if (cudaMalloc(&out, sizeof(uint32_t) * blocks * threads) != cudaSuccess) return AGB_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double res[3] = {0, 0, 0};
    for (int which = 0; which < 3; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (which == 0) int_peak_kernel<<<blocks, threads>>>(out, 12345u + rep, 1);
            else if (which == 1) int_peak_kernel<<<blocks, threads>>>(out, 12345u + rep, 0);
            else smem_peak_kernel<<<blocks, threads>>>(out, 12345u + rep);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return AGB_ERR_CUDA; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double n_thr = (double)blocks * threads;
        if (which == 0) res[0] = n_thr * kIters * 12.0 / (best * 1e-3);       // 6 LOP3 + 6 IMAD
        else if (which == 1) res[1] = n_thr * kIters * 16.0 / (best * 1e-3);  // 16 LOP3
        else res[2] = n_thr * kIters * 4.0 / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (int_ops_per_s) *int_ops_per_s = res[0];
    if (int_ops_per_s_alu_only) *int_ops_per_s_alu_only = res[1];
    if (smem_bytes_per_s) *smem_bytes_per_s = res[2];
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return AGB_OK;
}
