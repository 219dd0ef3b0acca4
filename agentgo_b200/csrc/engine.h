// engine.h — the object behind the C ABI (include/agentgo_b200.h).
//
// Replaces the CpuCount-driven thread pool the reference builds per genmove
// (Scheduler<MyWorker>, TinyMT.h:460-866; AgentGo.cpp:576-609): instead of 2*cores+2 worker
// threads pulling MyJobs, one persistent CUDA kernel pulls playout ids.
#ifndef AGB_ENGINE_H_
#define AGB_ENGINE_H_

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "agentgo_b200.h"
#include "host_board.h"
#include "kernel_args.h"

namespace agb {

typedef int (*AllreduceFn)(void* ctx, void* dev_int64, int count, void* cuda_stream);
typedef int (*AllgatherFn)(void* ctx, const void* send_dev_int64, void* recv_dev_int64, int count_per_shard, void* cuda_stream);

class Engine {
public:
    static int create(const agb_config& cfg, Engine** out, std::string* err);
    ~Engine();

    HostBoard board;
    std::string err;
    agb_config cfg;

    // general launch: starts x P playouts, ids sharded by g % shard_count (by_unit) or all local
    int launch(const std::vector<Position>& starts, const std::vector<StartDesc>& descs, int64_t P,
               uint64_t seed_base, bool want_diffs, bool shard_units, bool want_amaf = false, int counter_stride = 0);
    // waits, copies counters (and diffs) out.  Buffers are [n_starts] / [n_starts*P].
    // amaf: [amaf_classes() * AGB_AMAF_WORDS(n)] or nullptr; only valid after a launch with want_amaf
    int finish(int64_t* wins, int64_t* sum_pos, int64_t* sum_neg, int16_t* diffs, agb_stats* stats,
               int64_t* amaf = nullptr);

    int playouts_launch(int colour, const int16_t* cand_rc, int C, int64_t P, int avrg_win,
                        uint64_t seed_base, uint32_t position_id, bool want_diffs, bool want_amaf = false,
                        bool split_amaf = false);
    // AMAF tables of the launch in flight / last launch.  The reference recomputes its decay table
    // per candidate from the number of stones AFTER the candidate is played (AgentGo.cpp:86-94), so
    // candidates that capture use another table: with split_amaf the marks are accumulated per
    // distinct stone count (class k holds the candidates that leave amaf_class_stones()[k] stones).
    int amaf_classes() const { return (int)amaf_class_stones_.size(); }
    const std::vector<int>& amaf_class_stones() const { return amaf_class_stones_; }
    int playouts_batch(const agb_position* pos, int B, int64_t P, uint64_t seed_base,
                       uint32_t first_position_id, int64_t* wins, int64_t* sum_pos, int64_t* sum_neg,
                       int16_t* diffs, agb_stats* stats);
    int genmove(int colour, int64_t P_cand, int64_t P_root, uint64_t seed_base, int* row, int* col,
                int* is_pass, agb_stats* stats);

    int genmove_step = 0;                // MyGame::step, AgentGo.cpp:353
    int last_avrg_win = 0;               // avrg_win of the last genmove (testAvrgWin)
    std::vector<double> last_marks;      // total_mark of the last genmove (empty after a book move)

    void* counters_dev() const { return slot_[0].d_counters; }
    cudaStream_t stream() const { return stream_; }
    void set_allreduce(AllreduceFn fn, void* ctx) { allreduce_ = fn; allreduce_ctx_ = ctx; }
    void set_allgather(AllgatherFn fn, void* ctx) { allgather_ = fn; allgather_ctx_ = ctx; }
    int fail(int status, const std::string& msg) { err = msg; return status; }
    bool busy() const { return slot_[0].in_flight || slot_[1].in_flight; }   // between *_launch and *_finish: the position must not change

private:
    Engine() : board(13) {}
    int ensure(void** ptr, size_t* cap, size_t bytes);
    int cuda_fail(cudaError_t e, const char* what);

    int sm_count_ = 0;
    cudaStream_t stream_ = nullptr;
    // Buffers and state of one launch.  Two slots: genmove keeps the root launch (testAvrgWin) and the
    // candidate launch in flight together, stream-ordered, so that no host round trip separates them.
    struct Slot {
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        void* d_starts = nullptr; size_t cap_starts = 0;
        void* h_stage = nullptr; size_t cap_stage = 0;   // pinned staging of the start positions + descriptors
        void* d_descs = nullptr; size_t cap_descs = 0;
        void* d_counters = nullptr; size_t cap_counters = 0;
        void* d_small = nullptr;            // work counter + stats
        void* h_pinned = nullptr; size_t cap_pinned = 0;
        // state of the launch in flight
        bool in_flight = false;
        int n_starts = 0;
        int stride = 0;                     // counter stride of the launch in flight
        int64_t P = 0;
        bool want_diffs = false, shard_units = false, want_amaf = false;
    };
    Slot slot_[2];
    int cur_ = 0;
    Slot& sl() { return slot_[cur_]; }
    void* d_jump_ = nullptr;             // JumpTable (TinyMT32 T^16), built at create time
    void* d_jump_bytes_ = nullptr; size_t cap_jump_bytes_ = 0;   // JumpBytes of the group kernel
    void* d_jump_batch_ = nullptr; size_t cap_jump_batch_ = 0;   // its T^BATCH nibble table (goes to shared memory)
    void* d_elist_ = nullptr; size_t cap_elist_ = 0;   // group kernel: empty-point arrays of the resident playouts
    int jump_bytes_lanes_ = 0;           // lanes per playout d_jump_bytes_ was built for
    void* d_diffs_ = nullptr; size_t cap_diffs_ = 0;
    void* d_amaf_ = nullptr; size_t cap_amaf_ = 0;   // AMAF table [2][2][41][NN] int64
    int* d_avrg_ = nullptr;              // avrg_win of the genmove in flight, computed on the device
    bool avrg_from_device_ = false;      // the next launch reads avrg_win from d_avrg_
    std::vector<int> amaf_class_stones_ = std::vector<int>(1, 0);
    LaunchInfo info_ = {0, 0, 0, ""};
    AllreduceFn allreduce_ = nullptr;
    void* allreduce_ctx_ = nullptr;
    AllgatherFn allgather_ = nullptr;    // batch path: gathers every shard's [3][stride] table
    void* allgather_ctx_ = nullptr;
    void* d_gather_ = nullptr; size_t cap_gather_ = 0;
};

}  // namespace agb

// the opaque handle of the C ABI
struct agb_engine {
    agb::Engine* e;
};

#endif
