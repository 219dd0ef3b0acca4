// tests/csrc/emu_group.cpp — TEST-ONLY harness (never part of the product library).
// Compiles the device code of the group kernel (agentgo_b200/csrc/playout_group.cuh) for the CPU on
// top of the warp emulator (simt_emu.h) and runs it on one emulated warp, so that the kernel's
// logic — warp-uniform control flow included — is checked against the oracle on a machine without
// a GPU.  Nothing here is reachable from the C ABI.
#define AGB_EMU 1
#include "simt_emu.h"

#include <vector>

#include "agentgo_b200.h"
#include "host_board.h"
#include "playout_group.cuh"

using namespace agb;

namespace {

template <class C>
struct Job {
    KernelArgs a;
    uint32_t* smem;
    const uint4* tab;
};

template <class C>
void lane_entry(int lane, void* user) {
    Job<C>* j = (Job<C>*)user;
    grp::playout_group_body<C>(j->a, j->smem, lane, 0, j->tab);
}

template <class C>
void run(const KernelArgs& a) {
    Job<C> j;
    j.a = a;
    // exactly the warp's share of the block's shared memory (a sanitizer build sees any access past it)
    uint32_t* smem = (uint32_t*)aligned_alloc(64, (((size_t)C::warp_words * 4 + 63) / 64) * 64);
    for (int i = 0; i < C::warp_words; i++) smem[i] = 0xdeadbeefu;
    j.smem = smem;
    std::vector<uint16_t> scratch((size_t)C::NG * C::E, 0xdeadu);      // the warp's empty-point arrays (global memory)
    j.a.elist_scratch = scratch.data();
    j.tab = a.jump_batch->e;        // (in shared memory on the GPU)
    emu::run_warp(&lane_entry<C>, &j);
    free(smem);
}

template <int G>
int dispatch(int n, const KernelArgs& a) {
    if (a.amaf) {
        if (G != 8) return -1;
        switch (n) {
            case 9: run<grp::GCfg<9, 8, kJumpSteps, true>>(a); return 0;
            case 13: run<grp::GCfg<13, 8, kJumpSteps, true>>(a); return 0;
            case 19: run<grp::GCfg<19, 8, kJumpSteps, true>>(a); return 0;
        }
        return -1;
    }
    switch (n) {
        case 9: run<grp::GCfg<9, G, kJumpSteps>>(a); return 0;
        case 13: run<grp::GCfg<13, G, kJumpSteps>>(a); return 0;
        case 19: run<grp::GCfg<19, G, kJumpSteps>>(a); return 0;
    }
    return -1;
}

}  // namespace

// P playouts of one start (root mode: cand_row < 0; candidate mode: the candidate is played and
// the other side moves first), like tests/csrc/host_sim.cpp.  out3 = wins, sum_pos, sum_neg;
// stat4 = moves, draws, faults, playouts; amaf (may be NULL, candidate mode, 8 lanes): the AMAF table the
// playouts accumulate into.  Returns the number of collectives run (>= 0) or -1.
extern "C" long long emu_group_playouts(int n, int lanes, const uint32_t* moves, int n_moves, int colour,
                                        int cand_row, int cand_col, uint32_t position_id, uint32_t cand_idx,
                                        int64_t P, uint64_t seed_base, int avrg_win, int max_moves,
                                        int16_t* diffs, int64_t* out3, int64_t* stat4, int64_t* amaf) {
    HostBoard hb(n);
    if (hb.replay(moves, n_moves)) return -1;
    Position start = hb.position();
    int first = colour;
    if (cand_row >= 0) {
        start = hb.with_move(colour, cand_row * n + cand_col);
        first = 3 - colour;
    } else {
        cand_idx = AGB_ROOT_CANDIDATE;
    }
    GroupPosition gp;
    to_group(start, &gp);
    StartDesc d = {position_id, cand_idx, (int16_t)colour, (int16_t)first, avrg_win, 0};
    RngParams prm = {AGB_TINYMT_MAT1, AGB_TINYMT_MAT2, AGB_TINYMT_TMAT};
    std::vector<uint4> jb(jump_bytes_size(lanes) / sizeof(uint4));
    build_jump_bytes(prm, lanes, kJumpSteps, jb.data());
    unsigned long long work = 0, stats[kStatCount] = {0, 0, 0, 0};
    long long counters[3] = {0, 0, 0};
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.gstarts = &gp;
    a.jump_bytes = jb.data();
    JumpTable jt;
    build_jump(prm, lanes * kJumpSteps, &jt);
    a.jump_batch = &jt;
    a.descs = &d;
    a.n_starts = 1;
    a.counter_stride = 1;
    a.P = P;
    a.seed_base = seed_base;
    a.shard_rank = 0; a.shard_count = 1;
    a.local_units = P;
    a.work_counter = &work;
    a.counters = counters;
    a.diffs = diffs;
    a.amaf = (long long*)amaf;      // [2 which][2 sign][41 step][n*n] or NULL; candidate mode
    a.stats = stats;
    a.rng = prm;
    a.max_moves = max_moves > 0 ? max_moves : kDefaultMaxMoves;
    const uint64_t before = emu::collectives_run();
    int st = lanes == 8 ? dispatch<8>(n, a) : (lanes == 16 ? dispatch<16>(n, a) : (lanes == 32 ? dispatch<32>(n, a) : -1));
    if (st) return -1;
    out3[0] = counters[0]; out3[1] = counters[1]; out3[2] = counters[2];
    for (int i = 0; i < kStatCount; i++) stat4[i] = (int64_t)stats[i];
    return (long long)(emu::collectives_run() - before);
}

// The start position in the group kernel's device layout (kernel_args.h, GroupPosition), for the layout
// test: words[(n + 1)(n + 2) + 1], elist[n * n] (16-bit entries widened), scal[8] = n_el, stones,
// ko_key_col, ko_idx, last[0], last[1], ending, words per playout of GCfg<n, 8>.  Returns the word count.
extern "C" int emu_group_layout(int n, const uint32_t* moves, int n_moves, uint32_t* words, uint32_t* elist, int32_t* scal) {
    HostBoard hb(n);
    if (hb.replay(moves, n_moves)) return -1;
    GroupPosition gp;
    to_group(hb.position(), &gp);
    const int np = (n + 1) * (n + 2) + 1;
    for (int i = 0; i < np; i++) words[i] = gp.w[i];
    const uint16_t* el = reinterpret_cast<const uint16_t*>(gp.elist_w);
    for (int i = 0; i < n * n; i++) elist[i] = el[i];
    scal[0] = gp.n_el; scal[1] = gp.stones; scal[2] = gp.ko_key_col; scal[3] = gp.ko_idx;
    scal[4] = gp.last[0]; scal[5] = gp.last[1]; scal[6] = gp.ending;
    scal[7] = n == 9 ? grp::GCfg<9, 8, kJumpSteps>::words : (n == 13 ? grp::GCfg<13, 8, kJumpSteps>::words : grp::GCfg<19, 8, kJumpSteps>::words);
    return np;
}
