// tests/csrc/simt_emu.h — TEST-ONLY warp emulator (never part of the product library).
//
// Runs the device code of a warp-synchronous CUDA kernel on the CPU so that its logic can be
// debugged and checked against the oracle on a machine without a GPU.  The 32 lanes of a warp are
// 32 fibers (own stacks, a 10-line x86-64 context switch); a lane runs until it reaches a warp
// collective (__ballot_sync, __shfl_sync, __match_any_sync, __reduce_*_sync, __syncwarp, ...),
// parks its operands and yields; when every lane has parked the scheduler resolves the collective
// and resumes the lanes.  Between two collectives the lanes of a warp run one after the other, in
// a fixed order or — EMU_SHUFFLE=1 — in a random order per interval, which makes a missing
// __syncwarp() between a shared-memory store and a load by another lane show up as a mismatch.
//
// The emulator also CHECKS the property the group kernel is built on: every collective is reached
// by all 32 lanes at the same call site (warp-uniform control flow).  A lane that arrives at a
// different site, or exits while others wait, aborts with both sites named.
//
// Only what the kernels of this repo use is provided.  The kernel source is included after this
// header with AGB_EMU defined; the CUDA spellings below are then ordinary functions / macros.
#ifndef AGB_SIMT_EMU_H_
#define AGB_SIMT_EMU_H_

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace emu {

enum Op { OP_NONE, OP_BALLOT, OP_ANY, OP_ALL, OP_SHFL, OP_SHFL_XOR, OP_MATCH, OP_RED_XOR, OP_RED_OR, OP_RED_ADD,
          OP_RED_MIN, OP_RED_MAX, OP_SYNC };

struct Warp {
    void* sched_sp;
    void* lane_sp[32];
    char* stack[32];
    bool done[32];
    int cur;
    // mailbox of the collective being gathered
    int op[32];
    int site[32];
    unsigned mask[32];
    uint64_t val[32];
    int arg[32];
    uint64_t res[32];
    void (*entry)(int lane, void* user);
    void* user;
    uint64_t n_collectives;
    uint64_t limit;             // EMU_MAX_COLLECTIVES: abort (with the last sites) when a warp runs longer
    int trace[64];
    bool shuffle;
    uint32_t rng;
};

extern thread_local Warp* g_warp;
extern "C" void emu_switch(void** save_sp, void* load_sp);
void run_warp(void (*entry)(int lane, void* user), void* user);
uint64_t collective(int op, int site, unsigned mask, uint64_t val, int arg);
uint64_t collectives_run();

inline int lane_id() { return g_warp->cur; }

}  // namespace emu

// ---- CUDA spellings -------------------------------------------------------------------------
#define __ballot_sync(m, p) ((unsigned)emu::collective(emu::OP_BALLOT, __LINE__, (m), (p) ? 1u : 0u, 0))
#define __any_sync(m, p) ((int)emu::collective(emu::OP_ANY, __LINE__, (m), (p) ? 1u : 0u, 0))
#define __all_sync(m, p) ((int)emu::collective(emu::OP_ALL, __LINE__, (m), (p) ? 1u : 0u, 0))
#define __match_any_sync(m, v) ((unsigned)emu::collective(emu::OP_MATCH, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __reduce_xor_sync(m, v) ((unsigned)emu::collective(emu::OP_RED_XOR, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __reduce_or_sync(m, v) ((unsigned)emu::collective(emu::OP_RED_OR, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __reduce_add_sync(m, v) ((unsigned)emu::collective(emu::OP_RED_ADD, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __reduce_min_sync(m, v) ((unsigned)emu::collective(emu::OP_RED_MIN, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __reduce_max_sync(m, v) ((unsigned)emu::collective(emu::OP_RED_MAX, __LINE__, (m), (uint64_t)(uint32_t)(v), 0))
#define __syncwarp() ((void)emu::collective(emu::OP_SYNC, __LINE__, 0xffffffffu, 0, 0))

namespace emu {
template <typename T>
inline T shfl_impl(int op, int site, unsigned m, T v, int arg) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = collective(op, site, m, raw, arg);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
}  // namespace emu
#define __shfl_sync(m, v, src) emu::shfl_impl(emu::OP_SHFL, __LINE__, (m), (v), (int)(src))
#define __shfl_xor_sync(m, v, x) emu::shfl_impl(emu::OP_SHFL_XOR, __LINE__, (m), (v), (int)(x))

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned b = (unsigned)(src >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) b = (b & 0x80u) ? 0xffu : 0u;     // sign replication
        r |= b << (8 * i);
    }
    return r;
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
// the lanes of the emulated warp run one at a time, so plain read-modify-write is atomic
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
using std::max;
using std::min;

#endif  // AGB_SIMT_EMU_H_
