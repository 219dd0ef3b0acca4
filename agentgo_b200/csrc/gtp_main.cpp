// gtp_main.cpp — `agentgo_gtp`: the GTP engine (reference: _tmain, AgentGo.cpp:707-741) and a
// self-play driver for BASELINE.json configs[3] (device-timed genmove latency).
//
//   agentgo_gtp [--boardsize N] [--playouts P] [--total-playouts M] [--root-playouts R]
//               [--device D] [--gpus N] [--seed S] [--kernel auto|thread|warp]
//               [--amaf] [--static] [--book] [--reference-policy]
//               [--selfplay [--max-moves K]]
//
//   --playouts P        playouts per candidate (reference: numset = 150, AgentGo.cpp:85)
//   --total-playouts M  fixed budget per genmove, split evenly over the candidates
//   --root-playouts R   testAvrgWin playouts (reference: 500, AgentGo.cpp:413)
//   --gpus N            N GPUs of this box in one process: playouts sharded over N engines, counters
//                       combined by one NCCL allreduce per evaluation (agb_attach_nccl)
//   --amaf              add the AMAF terms amaf1 + amaf2 (AgentGo.cpp:329-338,654-655) to the marks
//   --static            add the static pattern marks of MyWorker::work (AgentGo.cpp:104-282)
//   --book              answer the first four genmoves from the 13x13 opening book (AgentGo.cpp:468-572)
//   --reference-policy  all three: the complete MyGame::onMove of the reference
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <iostream>
#include <sstream>
#include <vector>

#include "gtp_adapter.h"

int main(int argc, char** argv) {
    agb_config cfg;
    agb_default_config(&cfg);
    long long P = 150, R = 500, total = 0;
    unsigned long long seed = 20151001ull;
    bool selfplay = false;
    int max_moves = 800, gpus = 1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "--boardsize") cfg.board_size = atoi(next("--boardsize"));
        else if (a == "--playouts") P = atoll(next("--playouts"));
        else if (a == "--total-playouts") total = atoll(next("--total-playouts"));
        else if (a == "--root-playouts") R = atoll(next("--root-playouts"));
        else if (a == "--device") cfg.device = atoi(next("--device"));
        else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
        else if (a == "--max-moves") max_moves = atoi(next("--max-moves"));
        else if (a == "--selfplay") selfplay = true;
        else if (a == "--amaf") cfg.genmove_amaf = 1;
        else if (a == "--static") cfg.genmove_static = 1;
        else if (a == "--book") cfg.genmove_book = 1;
        else if (a == "--reference-policy") cfg.genmove_amaf = cfg.genmove_static = cfg.genmove_book = 1;
        else if (a == "--gpus") gpus = atoi(next("--gpus"));
        else if (a == "--kernel") {
            const std::string k = next("--kernel");
            cfg.kernel = k == "thread" ? AGB_KERNEL_THREAD : (k == "warp" ? AGB_KERNEL_WARP : AGB_KERNEL_AUTO);
        } else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    // --gpus N: one engine per GPU (devices device..device+N-1), shard i of N, one NCCL communicator
    std::vector<agb_engine*> engines;
    const int first_dev = cfg.device;
    for (int i = 0; i < gpus; i++) {
        agb_config c = cfg;
        if (gpus > 1) { c.device = first_dev + i; c.shard_rank = i; c.shard_count = gpus; }
        agb_engine* e = nullptr;
        if (agb_create(&c, &e) != AGB_OK) {
            fprintf(stderr, "agb_create failed: %s\n", agb_last_error(nullptr));
            return 1;
        }
        engines.push_back(e);
    }
    agb_nccl_group* group = nullptr;
    if (gpus > 1 && agb_attach_nccl(engines.data(), gpus, &group) != AGB_OK) {
        fprintf(stderr, "agb_attach_nccl failed: %s\n", agb_last_error(engines[0]));
        return 1;
    }
    agb_engine* eng = engines[0];
    MyGame gm(eng, total > 0 ? -total : P, R, seed);
    gm.peers.assign(engines.begin() + 1, engines.end());
    int rc = 0;
    if (!selfplay) {
        gm.MainLoop(std::cin, std::cout);
    } else {
        // both sides are this engine; the game ends on two consecutive passes
        std::vector<double> dev_ms, wall_ms;
        std::ostringstream sink;
        int passes = 0, moves = 0;
        long long playouts = 0;
        for (int t = 0; t < max_moves && passes < 2; t++) {
            std::ostringstream reply;
            const auto t0 = std::chrono::steady_clock::now();
            gm.HandleLine(t % 2 == 0 ? "genmove b" : "genmove w", reply);
            wall_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            dev_ms.push_back(gm.last.device_ms);
            playouts += gm.last.playouts;
            if (reply.str().find("pass") != std::string::npos) passes++; else { passes = 0; moves++; }
            fprintf(stderr, "%3d %s %s", t + 1, t % 2 == 0 ? "b" : "w", reply.str().substr(2, 5).c_str());
        }
        auto pct = [](std::vector<double> v, double q) {
            if (v.empty()) return 0.0;
            std::sort(v.begin(), v.end());
            return v[std::min(v.size() - 1, (size_t)(q * (v.size() - 1) + 0.5))];
        };
        printf("{\"selfplay\": {\"gpus\": %d, \"board_size\": %d, \"genmoves\": %zu, \"stones_played\": %d, \"playouts\": %lld, "
               "\"device_ms_median\": %.3f, \"device_ms_p95\": %.3f, \"wall_ms_median\": %.3f, \"wall_ms_p95\": %.3f, "
               "\"score_black\": %d, \"score_white\": %d}}\n",
               gpus, cfg.board_size, dev_ms.size(), moves, playouts, pct(dev_ms, 0.5), pct(dev_ms, 0.95), pct(wall_ms, 0.5),
               pct(wall_ms, 0.95), agb_count_score(eng, AGB_BLACK), agb_count_score(eng, AGB_WHITE));
    }
    agb_detach_nccl(group);
    for (agb_engine* e : engines) agb_destroy(e);
    return rc;
}
