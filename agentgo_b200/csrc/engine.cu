// engine.cu — see engine.h
#include "engine.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <thread>

namespace agb {

static const size_t kSmallWords = 8;   // [0] work counter, [1..4] stats

// MyGame::testAvrgWin's result, AgentGo.cpp:447: int(sum of the root playouts' scores / their number),
// C truncation, from the (all-reduced) root counters — on the device, so that the candidate launch can
// follow the root launch without a host round trip
__global__ void avrg_win_kernel(const long long* counters, int stride, long long playouts, int* out) {
    out[0] = (int)((counters[stride] + counters[2 * stride]) / playouts);
}

int Engine::cuda_fail(cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return fail(AGB_ERR_CUDA, buf);
}

#define AGB_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// a caller that zeroed the struct instead of calling agb_default_config gets the reference's values
static void defaults_for_zero_tunables(agb_config* c) {
    if (c->index_a == 0 && c->index_b == 0 && c->index_c == 0 && c->index_amaf_a1 == 0 && c->index_amaf_a2 == 0 &&
        c->index_amaf_b1 == 0 && c->index_amaf_b2 == 0 && c->index_cnsv == 0) {
        agb_config d;
        agb_default_config(&d);
        c->index_a = d.index_a; c->index_b = d.index_b; c->index_c = d.index_c;
        c->index_amaf_a1 = d.index_amaf_a1; c->index_amaf_a2 = d.index_amaf_a2;
        c->index_amaf_b1 = d.index_amaf_b1; c->index_amaf_b2 = d.index_amaf_b2; c->index_cnsv = d.index_cnsv;
    }
}

int Engine::create(const agb_config& cfg, Engine** out, std::string* err) {
    *out = nullptr;
    if (!HostBoard::size_supported(cfg.board_size)) {
        if (err) *err = "board_size must be 9, 13 or 19";
        return AGB_ERR_INVALID;
    }
    if (cfg.shard_count < 1 || cfg.shard_rank < 0 || cfg.shard_rank >= cfg.shard_count) {
        if (err) *err = "bad shard_rank/shard_count";
        return AGB_ERR_INVALID;
    }
    if (cfg.device == -1) {
        // position-only engine: replay / predicates / export work, every compute call fails
        Engine* h = new Engine();
        h->cfg = cfg;
        defaults_for_zero_tunables(&h->cfg);
        h->board.reset(cfg.board_size);
        *out = h;
        return AGB_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        // no CPU fallback by design: the product path fails loudly without a GPU
        if (err) *err = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return AGB_ERR_CUDA;
    }
    if (cfg.device < 0 || cfg.device >= ndev) {
        if (err) *err = "device ordinal out of range";
        return AGB_ERR_INVALID;
    }
    Engine* h = new Engine();
    h->cfg = cfg;
    if (h->cfg.mat1 == 0 && h->cfg.mat2 == 0 && h->cfg.tmat == 0) {
        h->cfg.mat1 = AGB_TINYMT_MAT1; h->cfg.mat2 = AGB_TINYMT_MAT2; h->cfg.tmat = AGB_TINYMT_TMAT;
    }
    if (h->cfg.max_moves <= 0) h->cfg.max_moves = kDefaultMaxMoves;
    defaults_for_zero_tunables(&h->cfg);
    h->board.reset(cfg.board_size);
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(cfg.device)) != cudaSuccess ||
        (e = cudaGetDeviceProperties(&prop, cfg.device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->stream_, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&h->slot_[0].ev0)) != cudaSuccess || (e = cudaEventCreate(&h->slot_[0].ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&h->slot_[1].ev0)) != cudaSuccess || (e = cudaEventCreate(&h->slot_[1].ev1)) != cudaSuccess ||
        (e = cudaMalloc(&h->slot_[0].d_small, kSmallWords * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&h->slot_[1].d_small, kSmallWords * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_avrg_, sizeof(int))) != cudaSuccess) {
        if (err) *err = std::string("CUDA init: ") + cudaGetErrorString(e);
        delete h;
        return AGB_ERR_CUDA;
    }
    h->sm_count_ = prop.multiProcessorCount;
    {
        JumpTable* jt = new JumpTable;
        RngParams prm = {h->cfg.mat1, h->cfg.mat2, h->cfg.tmat};
        build_jump(prm, jump_steps(), jt);
        e = cudaMalloc(&h->d_jump_, sizeof(JumpTable));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_jump_, jt, sizeof(JumpTable), cudaMemcpyHostToDevice);
        delete jt;
        if (e != cudaSuccess) {
            if (err) *err = std::string("CUDA init (jump table): ") + cudaGetErrorString(e);
            delete h;
            return AGB_ERR_CUDA;
        }
    }
    *out = h;
    return AGB_OK;
}

Engine::~Engine() {
    if (cfg.device < 0) return;
    cudaSetDevice(cfg.device);
    if (stream_) cudaStreamSynchronize(stream_);
    for (Slot& t : slot_) {
        cudaFree(t.d_starts); cudaFree(t.d_descs); cudaFree(t.d_counters); cudaFree(t.d_small);
        if (t.h_pinned) cudaFreeHost(t.h_pinned);
        if (t.h_stage) cudaFreeHost(t.h_stage);
        if (t.ev0) cudaEventDestroy(t.ev0);
        if (t.ev1) cudaEventDestroy(t.ev1);
    }
    cudaFree(d_diffs_); cudaFree(d_gather_); cudaFree(d_elist_); cudaFree(d_jump_); cudaFree(d_jump_bytes_); cudaFree(d_jump_batch_); cudaFree(d_amaf_);
    cudaFree(d_avrg_);
    if (stream_) cudaStreamDestroy(stream_);
}

int Engine::ensure(void** ptr, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return AGB_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    cudaError_t e = cudaMalloc(ptr, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(AGB_ERR_NOMEM, "cudaMalloc failed"); }
    *cap = want;
    return AGB_OK;
}

int Engine::launch(const std::vector<Position>& starts, const std::vector<StartDesc>& descs, int64_t P,
                   uint64_t seed_base, bool want_diffs, bool shard_units, bool want_amaf, int counter_stride) {
    if (cfg.device < 0) return fail(AGB_ERR_CUDA, "engine was created without a CUDA device (device = -1); no CPU fallback");
    if (sl().in_flight) return fail(AGB_ERR_STATE, "a launch is already in flight");
    if (starts.empty() || starts.size() != descs.size() || P <= 0)
        return fail(AGB_ERR_INVALID, "nothing to launch");
    AGB_CUDA(cudaSetDevice(cfg.device));
    const int S = (int)starts.size();
    const int stride = counter_stride > S ? counter_stride : S;
    int st;
    const bool thread_kernel = cfg.kernel == AGB_KERNEL_THREAD;
    if (want_amaf && thread_kernel)
        return fail(AGB_ERR_INVALID, "AMAF marks are not implemented by the thread kernel");
    // this shard's share of the launch
    const int sh_rank = shard_units ? cfg.shard_rank : 0, sh_count = shard_units ? cfg.shard_count : 1;
    const int64_t total = (int64_t)S * P;
    const int64_t local_units = total > sh_rank ? (total - sh_rank + sh_count - 1) / sh_count : 0;
    // Lanes per playout of the group kernel, 0 = another kernel.  AGB_KERNEL_AUTO: the group kernel (four playouts per warp) has the higher THROUGHPUT
    // (+23 % at 19x19), the warp kernel the lower LATENCY per playout (a warp serves one playout, not
    // four in lock step), so a launch of less than about one load of the machine (100 playouts per SM)
    // finishes earlier on the warp kernel: 0.7 ms against 2.3 ms for 2 000 playouts at 19x19, equal at
    // 16 000, 9.4 against 11.5 ms at 64 000 (tests/latency_probe.py, profiles/r02_latency_probe.txt).  The
    // reference's own genmove is 25 850 playouts.  On 9x9 a playout is 100 moves and the serial sections are
    // short: the warp kernel wins at every size (21.0 M against 20.4 M playouts/s, profiles/r02_quick_bench.txt).
    int group_lanes = 0;
    const bool big = local_units >= (int64_t)sm_count_ * 110 && board.n() > 9;
    if (cfg.kernel == AGB_KERNEL_GROUP8) group_lanes = 8;
    else if (cfg.kernel == AGB_KERNEL_AUTO) group_lanes = big ? 8 : 0;
    else if (cfg.kernel == AGB_KERNEL_GROUP16 && !want_amaf) group_lanes = 16;      // (the AMAF marks: 8 lanes per
    else if (cfg.kernel == AGB_KERNEL_GROUP32 && !want_amaf) group_lanes = 32;      //  playout or the warp kernel)
    const size_t amaf_bytes = sizeof(long long) * (size_t)AGB_AMAF_WORDS(board.n()) * (size_t)amaf_classes();
    if (want_amaf) {
        if ((st = ensure(&d_amaf_, &cap_amaf_, amaf_bytes))) return st;
        AGB_CUDA(cudaMemsetAsync(d_amaf_, 0, amaf_bytes, stream_));
    }
    const size_t start_bytes = (thread_kernel ? sizeof(Position) : (group_lanes ? sizeof(GroupPosition) : sizeof(PaddedPosition))) * (size_t)S;
    if ((st = ensure(&sl().d_starts, &sl().cap_starts, start_bytes))) return st;
    if ((st = ensure(&sl().d_descs, &sl().cap_descs, sizeof(StartDesc) * (size_t)S))) return st;
    if ((st = ensure(&sl().d_counters, &sl().cap_counters, sizeof(long long) * 3 * (size_t)stride))) return st;
    if (want_diffs && (st = ensure(&d_diffs_, &cap_diffs_, sizeof(int16_t) * (size_t)S * (size_t)P))) return st;

    // inputs are staged in PINNED host memory and copied asynchronously on the engine's stream
    const size_t desc_bytes = sizeof(StartDesc) * (size_t)S;
    if (sl().cap_stage < start_bytes + desc_bytes) {
        if (sl().h_stage) cudaFreeHost(sl().h_stage);
        sl().h_stage = nullptr; sl().cap_stage = 0;
        AGB_CUDA(cudaMallocHost(&sl().h_stage, start_bytes + desc_bytes));
        sl().cap_stage = start_bytes + desc_bytes;
    }
    if (thread_kernel) {
        memcpy(sl().h_stage, starts.data(), start_bytes);
    } else if (group_lanes) {
        GroupPosition* gp = (GroupPosition*)sl().h_stage;
        for (int i = 0; i < S; i++) to_group(starts[i], &gp[i]);
    } else {
        PaddedPosition* pp = (PaddedPosition*)sl().h_stage;
        for (int i = 0; i < S; i++) to_padded(starts[i], &pp[i]);
    }
    memcpy((char*)sl().h_stage + start_bytes, descs.data(), desc_bytes);
    AGB_CUDA(cudaMemcpyAsync(sl().d_starts, sl().h_stage, start_bytes, cudaMemcpyHostToDevice, stream_));
    AGB_CUDA(cudaMemcpyAsync(sl().d_descs, (char*)sl().h_stage + start_bytes, desc_bytes, cudaMemcpyHostToDevice, stream_));
    AGB_CUDA(cudaMemsetAsync(sl().d_counters, 0, sizeof(long long) * 3 * (size_t)stride, stream_));
    AGB_CUDA(cudaMemsetAsync(sl().d_small, 0, kSmallWords * sizeof(unsigned long long), stream_));

    KernelArgs a;
    memset(&a, 0, sizeof(a));
    a.starts = (const Position*)sl().d_starts;
    a.pstarts = (const PaddedPosition*)sl().d_starts;
    a.gstarts = (const GroupPosition*)sl().d_starts;
    a.jump = (const JumpTable*)d_jump_;
    if (group_lanes) {
        // the per-lane jump matrices of the group kernel, built on first use for this lane count
        if (jump_bytes_lanes_ != group_lanes) {
            std::vector<uint4> jb(jump_bytes_size(group_lanes) / sizeof(uint4));
            RngParams prm = {cfg.mat1, cfg.mat2, cfg.tmat};
            build_jump_bytes(prm, group_lanes, kJumpSteps, jb.data());
            if ((st = ensure(&d_jump_bytes_, &cap_jump_bytes_, jump_bytes_size(group_lanes)))) return st;
            AGB_CUDA(cudaMemcpy(d_jump_bytes_, jb.data(), jump_bytes_size(group_lanes), cudaMemcpyHostToDevice));
            JumpTable* jt = new JumpTable;
            build_jump(prm, group_lanes * kJumpSteps, jt);      // T^BATCH, the step of every lane from batch to batch
            st = ensure(&d_jump_batch_, &cap_jump_batch_, sizeof(JumpTable));
            cudaError_t ce = st ? cudaSuccess : cudaMemcpy(d_jump_batch_, jt, sizeof(JumpTable), cudaMemcpyHostToDevice);
            delete jt;
            if (st) return st;
            if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpy (jump table)");
            jump_bytes_lanes_ = group_lanes;
        }
        a.jump_batch = (const JumpTable*)d_jump_batch_;
        a.jump_bytes = (const uint4*)d_jump_bytes_;
        // the empty-point arrays of the resident playouts (global memory, kernel_args.h)
        if ((st = ensure(&d_elist_, &cap_elist_, (size_t)sm_count_ * kGroupMaxPlayoutsPerSm * kGroupElistCap * sizeof(uint16_t)))) return st;
        a.elist_scratch = (uint16_t*)d_elist_;
    }
    a.descs = (const StartDesc*)sl().d_descs;
    a.n_starts = S;
    a.P = P;
    a.seed_base = seed_base;
    a.shard_rank = sh_rank;
    a.shard_count = sh_count;
    a.local_units = local_units;
    if (a.local_units >= (1ll << 32)) return fail(AGB_ERR_INVALID, "more than 2^32 playouts in one launch of one shard");
    a.work_counter = (unsigned long long*)sl().d_small;
    a.stats = (unsigned long long*)sl().d_small + 1;
    a.counters = (long long*)sl().d_counters;
    a.counter_stride = stride;
    a.diffs = want_diffs ? (int16_t*)d_diffs_ : nullptr;
    a.amaf = want_amaf ? (long long*)d_amaf_ : nullptr;
    a.rng.mat1 = cfg.mat1; a.rng.mat2 = cfg.mat2; a.rng.tmat = cfg.tmat;
    a.max_moves = cfg.max_moves;
    a.avrg_dev = avrg_from_device_ ? d_avrg_ : nullptr;

    AGB_CUDA(cudaEventRecord(sl().ev0, stream_));
    if (a.local_units > 0) {
        cudaError_t e;
        const int n = board.n();
        if (thread_kernel) e = launch_playout_thread(n, a, sm_count_, stream_, &info_);
        else if (group_lanes) e = launch_playout_group(n, group_lanes, a, sm_count_, stream_, &info_);
        else e = launch_playout_warp(n, a, sm_count_, stream_, &info_);
        if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    }
    sl().in_flight = true;
    sl().n_starts = S; sl().stride = stride; sl().P = P; sl().want_diffs = want_diffs; sl().shard_units = shard_units; sl().want_amaf = want_amaf;
    return AGB_OK;
}

int Engine::finish(int64_t* wins, int64_t* sum_pos, int64_t* sum_neg, int16_t* diffs, agb_stats* stats,
                   int64_t* amaf) {
    if (!sl().in_flight) return fail(AGB_ERR_STATE, "no launch in flight");
    // everything that can be refused is refused before anything is enqueued behind the kernel
    if (amaf && !sl().want_amaf) return fail(AGB_ERR_STATE, "the launch in flight did not record AMAF marks");
    sl().in_flight = false;
    AGB_CUDA(cudaSetDevice(cfg.device));
    const int S = sl().n_starts;
    const int T = sl().stride;      // the counter table is [3][T], T >= S
    // an error from here on leaves work queued on the stream that may still write the caller's
    // buffers: wait for it before returning
    auto bail = [&](int status, const char* msg) {
        cudaStreamSynchronize(stream_);
        return fail(status, msg);
    };
    // snapshot the shard-local stats before any collective touches the counters
    const size_t need = sizeof(long long) * 3 * (size_t)T + kSmallWords * sizeof(unsigned long long);
    if (sl().cap_pinned < need) {
        if (sl().h_pinned) cudaFreeHost(sl().h_pinned);
        sl().h_pinned = nullptr; sl().cap_pinned = 0;
        if (cudaMallocHost(&sl().h_pinned, need) != cudaSuccess) return bail(AGB_ERR_NOMEM, "cudaMallocHost failed");
        sl().cap_pinned = need;
    }
    if (allreduce_) {
        int r = allreduce_(allreduce_ctx_, sl().d_counters, 3 * T, (void*)stream_);
        if (r) return bail(AGB_ERR_CUDA, "allreduce callback failed");
        if (sl().want_amaf) {
            r = allreduce_(allreduce_ctx_, d_amaf_, AGB_AMAF_WORDS(board.n()) * amaf_classes(), (void*)stream_);
            if (r) return bail(AGB_ERR_CUDA, "allreduce callback failed");
        }
    }
    // device time of the evaluation = kernel (+ collective), stream ordered
    AGB_CUDA(cudaEventRecord(sl().ev1, stream_));
    long long* hc = (long long*)sl().h_pinned;
    unsigned long long* hs = (unsigned long long*)(hc + 3 * (size_t)T);
    AGB_CUDA(cudaMemcpyAsync(hc, sl().d_counters, sizeof(long long) * 3 * (size_t)T, cudaMemcpyDeviceToHost, stream_));
    AGB_CUDA(cudaMemcpyAsync(hs, sl().d_small, kSmallWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
    if (diffs && sl().want_diffs && !(sl().shard_units && cfg.shard_count > 1))
        AGB_CUDA(cudaMemcpyAsync(diffs, d_diffs_, sizeof(int16_t) * (size_t)S * (size_t)sl().P, cudaMemcpyDeviceToHost, stream_));
    if (amaf) {
        AGB_CUDA(cudaMemcpyAsync(amaf, d_amaf_, sizeof(long long) * (size_t)AGB_AMAF_WORDS(board.n()) * (size_t)amaf_classes(),
                                 cudaMemcpyDeviceToHost, stream_));
    }
    AGB_CUDA(cudaStreamSynchronize(stream_));
    if (diffs && sl().want_diffs && sl().shard_units && cfg.shard_count > 1) {
        // only this shard's entries are defined on the device; leave the others untouched
        std::vector<int16_t> tmp((size_t)S * (size_t)sl().P);
        AGB_CUDA(cudaMemcpy(tmp.data(), d_diffs_, sizeof(int16_t) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int64_t g = cfg.shard_rank; g < (int64_t)tmp.size(); g += cfg.shard_count) diffs[g] = tmp[g];
    }
    for (int s = 0; s < S; s++) {
        if (wins) wins[s] = hc[s];
        if (sum_pos) sum_pos[s] = hc[T + s];
        if (sum_neg) sum_neg[s] = hc[2 * T + s];
    }
    float ms = 0.f;
    AGB_CUDA(cudaEventElapsedTime(&ms, sl().ev0, sl().ev1));
    const unsigned long long* stv = hs + 1;
    if (stats) {
        stats->device_ms = ms;
        stats->moves = (int64_t)stv[kStatMoves];
        stats->rng_draws = (int64_t)stv[kStatDraws];
        stats->faults = (int32_t)stv[kStatFaults];
        stats->playouts = (int64_t)stv[kStatPlayouts];
        stats->kernel_launches = 1;
    }
    if (stv[kStatFaults] != 0) return fail(AGB_ERR_FAULT, "a playout exceeded the move cap");
    return AGB_OK;
}

int Engine::playouts_launch(int colour, const int16_t* cand_rc, int C, int64_t P, int avrg_win,
                            uint64_t seed_base, uint32_t position_id, bool want_diffs, bool want_amaf,
                            bool split_amaf) {
    amaf_class_stones_.assign(1, 0);
    if (want_amaf && C == 0) return fail(AGB_ERR_INVALID, "AMAF marks exist in candidate mode only");
    if (colour != AGB_BLACK && colour != AGB_WHITE) return fail(AGB_ERR_INVALID, "bad colour");
    if (C < 0 || (C > 0 && !cand_rc) || P <= 0) return fail(AGB_ERR_INVALID, "bad candidates or P");
    const int n = board.n();
    std::vector<Position> starts;
    std::vector<StartDesc> descs;
    if (C == 0) {
        // root mode = MyGame::testAvrgWin (AgentGo.cpp:412-448): the mover plays first
        starts.push_back(board.position());
        StartDesc d = {position_id, AGB_ROOT_CANDIDATE, (int16_t)colour, (int16_t)colour, avrg_win, 0};
        descs.push_back(d);
    } else {
        if (split_amaf) amaf_class_stones_.clear();
        starts.reserve(C);
        for (int c = 0; c < C; c++) {
            const int row = cand_rc[2 * c], col = cand_rc[2 * c + 1];
            if (row < 0 || col < 0 || row >= n || col >= n) return fail(AGB_ERR_INVALID, "candidate off board");
            if (board.colour_at(row * n + col) != 0) return fail(AGB_ERR_OCCUPIED, "candidate point occupied");
            // MyWorker::work (AgentGo.cpp:83): candidate played, then the enemy moves first (:300,310)
            starts.push_back(board.with_move(colour, row * n + col));
            StartDesc d = {position_id, (uint32_t)c, (int16_t)colour, (int16_t)(3 - colour), avrg_win, 0};
            if (split_amaf) {
                const int stones = starts.back().num_black + starts.back().num_white;
                size_t k = 0;
                while (k < amaf_class_stones_.size() && amaf_class_stones_[k] != stones) k++;
                if (k == amaf_class_stones_.size()) amaf_class_stones_.push_back(stones);
                d.amaf_class = (int32_t)k;
            }
            descs.push_back(d);
        }
    }
    return launch(starts, descs, P, seed_base, want_diffs, true, want_amaf);
}

int Engine::playouts_batch(const agb_position* pos, int B, int64_t P, uint64_t seed_base,
                           uint32_t first_position_id, int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                           int16_t* diffs, agb_stats* stats) {
    if (!pos || B <= 0 || P <= 0) return fail(AGB_ERR_INVALID, "bad batch");
    std::vector<int> owner;     // batch index of each local start
    for (int b = 0; b < B; b++) {
        const uint32_t pid = first_position_id + (uint32_t)b;
        if ((int)(pid % (uint32_t)cfg.shard_count) != cfg.shard_rank) continue;
        const int colour = pos[b].to_move;
        if (colour != AGB_BLACK && colour != AGB_WHITE) return fail(AGB_ERR_INVALID, "bad to_move");
        owner.push_back(b);
    }
    // The move lists are replayed on host threads (a position is defined by its move list: the orders
    // the policy observes are history dependent), one HostBoard per thread, positions dealt round robin.
    const int S0 = (int)owner.size();
    std::vector<Position> starts((size_t)S0);
    std::vector<StartDesc> descs((size_t)S0);
    std::vector<int> bad((size_t)S0, 0);
    {
        const int n = board.n();
        int nthreads = (int)std::thread::hardware_concurrency();
        nthreads = std::max(1, std::min(std::min(nthreads, 16), S0 / 8));
        auto work = [&](int t) {
            HostBoard hb(n);
            for (int s = t; s < S0; s += nthreads) {
                const int b = owner[s];
                hb.clear();
                bad[s] = hb.replay(pos[b].moves, pos[b].n_moves);
                starts[s] = hb.position();
                const uint32_t pid = first_position_id + (uint32_t)b;
                StartDesc d = {pid, AGB_ROOT_CANDIDATE, (int16_t)pos[b].to_move, (int16_t)pos[b].to_move, 0, 0};
                descs[s] = d;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
        work(0);
        for (std::thread& t : th) t.join();
    }
    for (int s = 0; s < S0; s++)
        if (bad[s]) return fail(bad[s] == -2 ? AGB_ERR_OCCUPIED : AGB_ERR_INVALID, "bad move list in batch");
    for (int b = 0; b < B; b++) {
        if (wins) wins[b] = 0;
        if (sum_pos) sum_pos[b] = 0;
        if (sum_neg) sum_neg[b] = 0;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (starts.empty() && !(cfg.shard_count > 1 && allgather_)) return AGB_OK;
    const int S = (int)starts.size();
    // Several shards with a gather hook: every shard pads its table to the size of the largest one,
    // ONE allgather follows the kernel on the engine's stream and every shard returns the whole
    // [B] table ("shards by position, then one allgather of results").
    const int R = cfg.shard_count;
    const bool gather = R > 1 && allgather_ != nullptr;
    int stride = S;
    std::vector<int> n_of(R, 0);
    if (gather) {
        for (int b = 0; b < B; b++) n_of[(first_position_id + (uint32_t)b) % (uint32_t)R]++;
        stride = *std::max_element(n_of.begin(), n_of.end());
    }
    int st = 0;
    if (S > 0) {
        st = launch(starts, descs, P, seed_base, diffs != nullptr, false, false, stride);
        if (st) return st;
    } else if (gather) {
        // nothing to play here, but the collective needs this shard's (empty) table
        AGB_CUDA(cudaSetDevice(cfg.device));
        if ((st = ensure(&sl().d_counters, &sl().cap_counters, sizeof(long long) * 3 * (size_t)stride))) return st;
        AGB_CUDA(cudaMemsetAsync(sl().d_counters, 0, sizeof(long long) * 3 * (size_t)stride, stream_));
    }
    std::vector<long long> table;
    if (gather) {
        const size_t words = (size_t)R * 3 * (size_t)stride;
        if ((st = ensure(&d_gather_, &cap_gather_, sizeof(long long) * words))) return st;
        if (allgather_(allgather_ctx_, sl().d_counters, d_gather_, 3 * stride, (void*)stream_))
            return fail(AGB_ERR_CUDA, "allgather callback failed");
        table.resize(words);
        AGB_CUDA(cudaMemcpyAsync(table.data(), d_gather_, sizeof(long long) * words, cudaMemcpyDeviceToHost, stream_));
    }
    std::vector<int64_t> w(S), sp(S), sn(S);
    std::vector<int16_t> dl;
    if (diffs) dl.resize((size_t)S * (size_t)P);
    if (S > 0) {
        AllreduceFn saved = allreduce_;
        allreduce_ = nullptr;       // results of a batch are gathered per position, not summed
        st = finish(w.data(), sp.data(), sn.data(), diffs ? dl.data() : nullptr, stats);
        allreduce_ = saved;
        if (st) return st;
    } else if (gather) {
        AGB_CUDA(cudaStreamSynchronize(stream_));
    }
    if (gather) {
        // position b was slot k of shard r, k counting that shard's positions in batch order
        std::vector<int> next(R, 0);
        for (int b = 0; b < B; b++) {
            const int r = (int)((first_position_id + (uint32_t)b) % (uint32_t)R);
            const long long* t = table.data() + (size_t)r * 3 * (size_t)stride;
            const int k = next[r]++;
            if (wins) wins[b] = t[k];
            if (sum_pos) sum_pos[b] = t[stride + k];
            if (sum_neg) sum_neg[b] = t[2 * stride + k];
        }
    }
    for (int s = 0; s < S; s++) {
        const int b = owner[s];
        if (!gather) {
            if (wins) wins[b] = w[s];
            if (sum_pos) sum_pos[b] = sp[s];
            if (sum_neg) sum_neg[b] = sn[s];
        }
        if (diffs) memcpy(diffs + (size_t)b * (size_t)P, dl.data() + (size_t)s * (size_t)P, sizeof(int16_t) * (size_t)P);
    }
    return AGB_OK;
}

// MyGame::onMove (AgentGo.cpp:451-692):
//   step 1..4: opening book (:468-572, cfg.genmove_book, 13x13 only);
//   then avrg_win = testAvrgWin (:579), candidate filter (:586-592), playouts per candidate,
//   mc[i][j] = 100*index_c * sum(score), score = w >= 0 ? w : w*cnsv, cnsv = 2^(avrg_win*index_cnsv)
//   (:99,:328,:340), static pattern marks (:104-282, cfg.genmove_static), AMAF (:329-338,
//   cfg.genmove_amaf), argmax with first-wins ties in row-major order (:647-676).
int Engine::genmove(int colour, int64_t P_cand, int64_t P_root, uint64_t seed_base, int* row, int* col,
                    int* is_pass, agb_stats* stats) {
    if (colour != AGB_BLACK && colour != AGB_WHITE) return fail(AGB_ERR_INVALID, "bad colour");
    if (cfg.shard_count > 1 && !allreduce_)
        return fail(AGB_ERR_STATE, "genmove on a sharded engine needs an allreduce callback");
    const int n = board.n();
    agb_stats acc, st1;
    memset(&acc, 0, sizeof(acc));
    genmove_step++;                                           // :466
    last_marks.clear();
    if (cfg.genmove_book && genmove_step <= 4 && board.book_move(row, col)) {
        *is_pass = 0;
        if (stats) *stats = acc;
        return AGB_OK;
    }
    // The candidates do not depend on the root playouts (:586-592), only their threshold avrg_win does
    // (:579).  So both launches go out back to back on the engine's stream — root playouts (slot 1),
    // their allreduce, avrg_win worked out ON THE DEVICE, candidate playouts reading it from there
    // (slot 0) — and the host waits once, at the end.
    std::vector<int16_t> cand;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (board.is_candidate(colour, i, j)) { cand.push_back((int16_t)i); cand.push_back((int16_t)j); }
    const int C = (int)cand.size() / 2;
    int st;
    cur_ = 1;
    st = playouts_launch(colour, nullptr, 0, P_root, 0, seed_base, 0, false);
    cur_ = 0;
    if (st) return st;
    Slot& root = slot_[1];
    auto drop_root = [&](int status) {      // an error after the root launch: wait for it, forget it
        cudaStreamSynchronize(stream_);
        root.in_flight = false;
        avrg_from_device_ = false;
        return status;
    };
    if (allreduce_ && allreduce_(allreduce_ctx_, root.d_counters, 3 * root.stride, (void*)stream_))
        return drop_root(fail(AGB_ERR_CUDA, "allreduce callback failed"));
    avrg_win_kernel<<<1, 1, 0, stream_>>>((const long long*)root.d_counters, root.stride, (long long)P_root, d_avrg_);
    // root results for the host: counters, statistics, avrg_win (pinned; complete at the final wait)
    const size_t root_need = sizeof(long long) * 3 + kSmallWords * sizeof(unsigned long long) + sizeof(long long);
    if (root.cap_pinned < root_need) {
        if (root.h_pinned) cudaFreeHost(root.h_pinned);
        root.h_pinned = nullptr; root.cap_pinned = 0;
        if (cudaMallocHost(&root.h_pinned, root_need) != cudaSuccess) return drop_root(fail(AGB_ERR_NOMEM, "cudaMallocHost failed"));
        root.cap_pinned = root_need;
    }
    long long* rc_h = (long long*)root.h_pinned;
    unsigned long long* rs_h = (unsigned long long*)(rc_h + 3);
    int* avrg_h = (int*)(rs_h + kSmallWords);
    if (cudaMemcpyAsync(rc_h, root.d_counters, sizeof(long long) * 3, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaMemcpyAsync(rs_h, root.d_small, kSmallWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
        cudaMemcpyAsync(avrg_h, d_avrg_, sizeof(int), cudaMemcpyDeviceToHost, stream_) != cudaSuccess)
        return drop_root(fail(AGB_ERR_CUDA, "cudaMemcpyAsync failed"));
    AGB_CUDA(cudaEventRecord(root.ev1, stream_));
    *is_pass = 1; *row = -1; *col = -1;
    {
        // (no candidate survives the filter: nothing is played out, but the reference still takes the
        // first point that is empty and neither a true eye, a suicide nor the ko point — those the filter
        // dropped for checkNoSense compete with mark 0, :647-676)
        std::vector<int64_t> cw(C), csp(C), csn(C);
        if (P_cand < 0) P_cand = C > 0 ? (-P_cand + C - 1) / C : -P_cand;      // fixed budget per genmove, split over candidates
        const bool use_amaf = C > 0 && cfg.genmove_amaf != 0 && cfg.kernel != AGB_KERNEL_THREAD;
        std::vector<int64_t> amaf;
        if (C > 0) {
            avrg_from_device_ = true;       // the kernel takes avrg_win from d_avrg_, not from the descriptors
            st = playouts_launch(colour, cand.data(), C, P_cand, 0, seed_base + 1, 0, false, use_amaf, use_amaf);
            avrg_from_device_ = false;
            if (st) return drop_root(st);
            if (use_amaf) amaf.resize((size_t)AGB_AMAF_WORDS(n) * (size_t)amaf_classes());
            st = finish(cw.data(), csp.data(), csn.data(), nullptr, &st1, use_amaf ? amaf.data() : nullptr);   // waits for the stream
            if (st) return drop_root(st);
        } else {
            AGB_CUDA(cudaStreamSynchronize(stream_));
            memset(&st1, 0, sizeof(st1));
        }
        root.in_flight = false;
        // both launches are complete: the root's results are in pinned memory
        const unsigned long long* rst = rs_h + 1;
        if (rst[kStatFaults] != 0) return fail(AGB_ERR_FAULT, "a playout exceeded the move cap");
        const int avrg_win = *avrg_h;                             // int((sp + sn) / P_root), AgentGo.cpp:447
        float ms = 0.f;
        AGB_CUDA(cudaEventElapsedTime(&ms, root.ev0, C > 0 ? slot_[0].ev1 : root.ev1));
        acc.device_ms = ms;                                       // root launch to the end of the candidates' collective
        acc.playouts = (int64_t)rst[kStatPlayouts] + st1.playouts;
        acc.moves = (int64_t)rst[kStatMoves] + st1.moves;
        acc.rng_draws = (int64_t)rst[kStatDraws] + st1.rng_draws;
        acc.kernel_launches = 1 + st1.kernel_launches;
        last_avrg_win = avrg_win;
        const double index_c = cfg.index_c;
        const double cnsv = std::pow(2.0, avrg_win * cfg.index_cnsv);    // AgentGo.cpp:25,99
        // AMAF terms (:86-94,:329-338): decay table over the step of the first play.  The table
        // depends on the number of stones after the candidate is played, so the marks were
        // accumulated per distinct stone count; index_amaf_* defaults of :21-24, index_c = 1.
        std::vector<double> amaf12((size_t)n * n, 0.0);
        if (use_amaf) {
            const int NN = n * n, S = AGB_AMAF_RANGE + 1;
            for (int k = 0; k < amaf_classes(); k++) {
                const int64_t* A = amaf.data() + (size_t)k * AGB_AMAF_WORDS(n);
                const int num_piece = amaf_class_stones()[k];
                const double num_a = 1.0 / (cfg.index_amaf_a1 - cfg.index_amaf_a2 * num_piece);      // :87
                const double num_b = cfg.index_amaf_b1 - cfg.index_amaf_b2 * num_piece;              // :88
                double tab[AGB_AMAF_RANGE];
                for (int t = 0; t < AGB_AMAF_RANGE; t++) {
                    double v = 1 - std::pow(num_a * t, num_b);
                    if (!(v >= 0)) v = 0;           // also catches NaN (negative base on boards > 13x13)
                    else if (v > 1) v = 1;
                    tab[t] = v;
                }
                for (int q = 0; q < NN; q++) {
                    double a1 = 0, a2 = 0;
                    for (int sidx = 2; sidx <= AGB_AMAF_RANGE; sidx++) {
                        const double t = 100.0 * tab[sidx - 1] * index_c;
                        a1 += t * ((double)A[((size_t)(0 * 2 + 0) * S + sidx) * NN + q] + cnsv * (double)A[((size_t)(0 * 2 + 1) * S + sidx) * NN + q]);
                        a2 -= t * ((double)A[((size_t)(1 * 2 + 0) * S + sidx) * NN + q] + cnsv * (double)A[((size_t)(1 * 2 + 1) * S + sidx) * NN + q]);
                    }
                    amaf12[q] += a1 + a2;
                }
            }
        }
        // argmax over every eligible point in row-major order, first maximum wins (:647-676).
        // Points the candidate filter dropped only for checkNoSense still compete with mark 0.
        bool init = false;
        double best = 0;
        int c = 0;
        last_marks.assign((size_t)n * n, 0.0);
        // int(numset*index_a), int(numset*index_b) (:95-96); 64 bits: the reference's numset is 150, here
        // the playouts per candidate are the caller's
        const int64_t bonus_a = (int64_t)((double)P_cand * cfg.index_a);
        const int64_t bonus_b = (int64_t)((double)P_cand * cfg.index_b);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                const bool is_cand = c < C && cand[2 * c] == i && cand[2 * c + 1] == j;
                double mark = 0;
                if (is_cand) {
                    if (cfg.genmove_static) {                 // :104-282, added before any playout term
                        int64_t ua = 0, ub = 0;
                        board.static_units(colour, i, j, &ua, &ub);
                        mark = (double)(ua * bonus_a + ub * bonus_b);
                    }
                    // total_mark gets every score once (:339) and mc again (:653)
                    mark += 2.0 * 100.0 * index_c * ((double)csp[c] + cnsv * (double)csn[c]);
                    c++;
                }
                if (!board.is_eligible(colour, i, j)) { last_marks[(size_t)i * n + j] = mark; continue; }
                mark += amaf12[(size_t)i * n + j];                                      // :654-655
                last_marks[(size_t)i * n + j] = mark;
                if (!init || mark > best) { best = mark; *row = i; *col = j; init = true; }
            }
        *is_pass = init ? 0 : 1;                              // domove, :667,:688-691
    }
    if (stats) *stats = acc;
    return AGB_OK;
}

}  // namespace agb
