// gtp_main.cpp — `agentgo_gtp`: the GTP engine (reference: _tmain, AgentGo.cpp:707-741) and a
// self-play driver for BASELINE.json configs[3] (device-timed genmove latency).
//
//   agentgo_gtp [--boardsize N] [--playouts P] [--total-playouts M] [--root-playouts R]
//               [--device D] [--gpus N] [--seed S] [--kernel auto|thread|warp|group8|group16|group32]
//               [--amaf] [--static] [--book] [--reference-policy]
//               [--selfplay [--max-moves K]]
//               [--match G [--white-playouts P] [--white-policy flat|reference] [--black-cmd CMD] [--white-cmd CMD]]
//               [index_a index_b index_c amaf_a1 amaf_a2 amaf_b1 amaf_b2 cnsv]
//
//   eight numbers       the reference's own command line (AgentGo.cpp:709-719): the tunables of the static marks,
//                       the playout term, the AMAF decay table and cnsv (agb_config::index_*)
//   --playouts P        playouts per candidate (reference: numset = 150, AgentGo.cpp:85)
//   --total-playouts M  fixed budget per genmove, split evenly over the candidates
//   --root-playouts R   testAvrgWin playouts (reference: 500, AgentGo.cpp:413)
//   --gpus N            N GPUs of this box in one process: playouts sharded over N engines, counters
//                       combined by one NCCL allreduce per evaluation (agb_attach_nccl)
//   --amaf              add the AMAF terms amaf1 + amaf2 (AgentGo.cpp:329-338,654-655) to the marks
//   --static            add the static pattern marks of MyWorker::work (AgentGo.cpp:104-282)
//   --book              answer the first four genmoves from the 13x13 opening book (AgentGo.cpp:468-572)
//   --reference-policy  all three: the complete MyGame::onMove of the reference
//   --match G           G games under the referee of the reference's GA tuner (SimulateOneGame,
//                       GeneHatchery.cpp:438-568) between two GTP engines run as CHILD PROCESSES over
//                       pipes: black is this program with the options above, white this program with
//                       --white-*, or any GTP engine given as a shell command (--black-cmd, --white-cmd);
//                       the referee keeps its own board, gives a side that passes a random policy
//                       move instead (getRandomPiece, :474,:520), ends the game when a side has no
//                       move after the other had none either, and scores with CalcResults (:393-436)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <iostream>
#include <sstream>
#include <vector>

#include <signal.h>
#include <strings.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "gtp_adapter.h"

// A GTP engine as a child process on two pipes (reference: CreatePipes / CreateRedirectedProcess /
// MyPrintf / WaitForResponse, GeneHatchery.cpp:221-335).
struct GtpChild {
    pid_t pid = -1;
    FILE* to = nullptr;
    FILE* from = nullptr;
    bool spawn(const std::string& cmd) {
        int in[2], out[2];
        if (pipe(in) != 0 || pipe(out) != 0) return false;
        pid = fork();
        if (pid < 0) return false;
        if (pid == 0) {
            dup2(in[0], 0);
            dup2(out[1], 1);
            close(in[0]); close(in[1]); close(out[0]); close(out[1]);
            execl("/bin/sh", "sh", "-c", cmd.c_str(), (char*)nullptr);
            _exit(127);
        }
        close(in[0]);
        close(out[1]);
        to = fdopen(in[1], "w");
        from = fdopen(out[0], "r");
        return to && from;
    }
    // one command, one reply: everything up to the empty line that ends a GTP response
    std::string ask(const std::string& line) {
        if (!to || !from) return "";
        fprintf(to, "%s\n", line.c_str());
        fflush(to);
        std::string rep;
        char buf[512];
        while (fgets(buf, sizeof(buf), from)) {
            if (buf[0] == '\n' || (buf[0] == '\r' && buf[1] == '\n')) {
                if (!rep.empty()) break;
                continue;
            }
            rep += buf;
        }
        while (!rep.empty() && (rep.back() == '\n' || rep.back() == '\r')) rep.pop_back();
        return rep;
    }
    void stop() {
        if (to) { fprintf(to, "quit\n"); fflush(to); fclose(to); to = nullptr; }
        if (from) { fclose(from); from = nullptr; }
        if (pid > 0) {
            int st = 0;
            for (int i = 0; i < 50 && waitpid(pid, &st, WNOHANG) == 0; i++) usleep(20000);
            if (waitpid(pid, &st, WNOHANG) == 0) { kill(pid, SIGKILL); waitpid(pid, &st, 0); }
            pid = -1;
        }
    }
};

int main(int argc, char** argv) {
    agb_config cfg;
    agb_default_config(&cfg);
    long long P = 150, R = 500, total = 0;
    unsigned long long seed = 20151001ull;
    bool selfplay = false;
    int max_moves = 800, gpus = 1, match_games = 0, white_policy = -1;
    long long white_P = -1;
    std::vector<double> positional;
    std::string black_cmd, white_cmd;
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
        else if (a == "--match") match_games = atoi(next("--match"));
        else if (a == "--white-playouts") white_P = atoll(next("--white-playouts"));
        else if (a == "--black-cmd") black_cmd = next("--black-cmd");
        else if (a == "--white-cmd") white_cmd = next("--white-cmd");
        else if (a == "--white-policy") white_policy = std::string(next("--white-policy")) == "reference" ? 1 : 0;
        else if (a == "--kernel") {
            const std::string k = next("--kernel");
            cfg.kernel = k == "thread" ? AGB_KERNEL_THREAD : (k == "warp" ? AGB_KERNEL_WARP : (k == "group8" ? AGB_KERNEL_GROUP8 :
                         (k == "group16" ? AGB_KERNEL_GROUP16 : (k == "group32" ? AGB_KERNEL_GROUP32 : AGB_KERNEL_AUTO))));
        } else if (a.size() >= 1 && (isdigit((unsigned char)a[0]) || ((a[0] == '-' || a[0] == '.') && a.size() > 1 && a[1] != '-'))) {
            positional.push_back(atof(a.c_str()));
        } else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    // the reference's own command line: exactly eight doubles (AgentGo.cpp:709-719, `if (argc == 9)`)
    if (positional.size() == 8) {
        cfg.index_a = positional[0]; cfg.index_b = positional[1]; cfg.index_c = positional[2];
        cfg.index_amaf_a1 = positional[3]; cfg.index_amaf_a2 = positional[4];
        cfg.index_amaf_b1 = positional[5]; cfg.index_amaf_b2 = positional[6]; cfg.index_cnsv = positional[7];
    } else if (!positional.empty()) {
        fprintf(stderr, "expected the reference's eight tunables (index_a index_b index_c amaf_a1 amaf_a2 amaf_b1 amaf_b2 cnsv), got %zu numbers\n",
                positional.size());
        return 2;
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
    if (match_games > 0) {
        // Row f4 of SURVEY.md §8: the referee of the reference's GA tuner (SimulateOneGame,
        // GeneHatchery.cpp:438-568).  Like there, the two players are CHILD PROCESSES that speak GTP
        // over pipes (CreatePipes / CreateRedirectedProcess, :221-281; here pipe/fork/exec), started
        // anew for every game; the referee keeps its own board, asks `genmove`, relays `play`, gives a
        // side that passes a random policy move instead (:474,:520), ends the game when a side has no
        // move after the other had none, and scores with CalcResults (:393-436).  Any GTP engine can
        // be a player (--black-cmd / --white-cmd); by default both are this program.
        agb_detach_nccl(group);
        for (agb_engine* e : engines) agb_destroy(e);       // the players own the GPUs, not the referee
        engines.clear();
        agb_config rc = cfg;
        rc.device = -1;                                    // the referee only keeps a board
        agb_engine* ref = nullptr;
        if (agb_create(&rc, &ref) != AGB_OK) {
            fprintf(stderr, "agb_create failed: %s\n", agb_last_error(nullptr));
            return 1;
        }
        auto self_cmd = [&](bool white, unsigned long long sd) {
            std::ostringstream c;
            const long long pl = white && white_P > 0 ? white_P : P;
            const int pol = white && white_policy >= 0 ? white_policy : (cfg.genmove_static ? 1 : 0);
            c << argv[0] << " --boardsize " << cfg.board_size << " --device " << first_dev << " --gpus " << gpus
              << " --root-playouts " << R << " --seed " << sd;
            if (!white && total > 0) c << " --total-playouts " << total; else c << " --playouts " << pl;
            if (pol) c << " --reference-policy";
            else if (!white && cfg.genmove_amaf) c << " --amaf";
            char num[64];
            const double tun[8] = {cfg.index_a, cfg.index_b, cfg.index_c, cfg.index_amaf_a1, cfg.index_amaf_a2,
                                   cfg.index_amaf_b1, cfg.index_amaf_b2, cfg.index_cnsv};
            if (!white)     // black is the engine under test: it gets the tunables, like the tuner's debugee (:443-449)
                for (double v : tun) { snprintf(num, sizeof(num), " %.17g", v); c << num; }
            return c.str();
        };
        const char cname[2] = {'b', 'w'};
        auto vertex = [&](int row, int col) {
            return std::string(1, GaIntToChar(row)) + std::to_string(col + 1);
        };
        std::vector<int> dscores;
        int black_wins = 0;
        long long stones = 0;
        for (int g = 0; g < match_games; g++) {
            GtpChild side[2];
            const std::string cmd[2] = {
                black_cmd.empty() ? self_cmd(false, seed + 1000003ull * (unsigned long long)g) : black_cmd,
                white_cmd.empty() ? self_cmd(true, seed + 0x9e3779b9ull + 1000003ull * (unsigned long long)g) : white_cmd};
            for (int w = 0; w < 2; w++)
                if (!side[w].spawn(cmd[w])) { fprintf(stderr, "cannot start: %s\n", cmd[w].c_str()); return 1; }
            bool broken = false;
            // IsReplyOK, :369-375: a reply that does not start with '=' ends the match
            auto tell = [&](int who, const std::string& c) {
                const std::string rep = side[who].ask(c);
                if (rep.empty() || rep[0] != '=') {
                    fprintf(stderr, "%s engine: bad reply to '%s': '%s'\n", who ? "white" : "black", c.c_str(), rep.c_str());
                    broken = true;
                }
                return rep;
            };
            for (int w = 0; w < 2; w++) {
                tell(w, "clear_board");                     // :459-466
                tell(w, "boardsize " + std::to_string(cfg.board_size));
                tell(w, "clear_board");                     // (a boardsize this engine honours resets the board)
            }
            agb_clear_board(ref);
            // played[c]: did side c put a stone in its last turn (bplay / lastwplay, :461,:511)
            bool played[2] = {true, true};
            int dscore = 0;
            for (int t = 0; t < 2 * max_moves && !broken; t++) {
                const int c = t & 1, o = 1 - c;
                // Genmove, :337-367: "= <letter><number>" is a move, "= pass" is none, anything else an error
                const std::string rep = side[c].ask(std::string("genmove ") + cname[c]);
                int row = -1, col = -1, num = 0;
                char letter = 0;
                bool moved = false;
                if (sscanf(rep.c_str(), "= %c%d", &letter, &num) == 2 && strncasecmp(rep.c_str() + 2, "pass", 4) != 0) {
                    row = GaCharToInt(tolower((int)letter));
                    col = num - 1;
                    moved = true;
                } else if (!(rep.size() >= 6 && rep[0] == '=' && strncasecmp(rep.c_str() + rep.find_first_not_of(" \t", 1), "pass", 4) == 0)) {
                    fprintf(stderr, "%s engine: bad reply to genmove: '%s'\n", c ? "white" : "black", rep.c_str());
                    broken = true;
                    break;
                }
                if (moved) {                                            // letter = row, number-1 = col
                    if (agb_play(ref, c + 1, row, col) != AGB_OK) {
                        fprintf(stderr, "%s engine played an illegal move %s\n", c ? "white" : "black", rep.c_str());
                        broken = true;
                        break;
                    }
                    tell(o, std::string("play ") + cname[c] + " " + vertex(row, col));
                } else {
                    int is_pass = 1;
                    agb_random_piece(ref, c + 1, (uint32_t)(seed + 977ull * (unsigned long long)g + (unsigned long long)t), &row, &col, &is_pass);
                    if (is_pass) {
                        if (!played[o]) { agb_calc_result(ref, &dscore); break; }   // :476-481,:522-527
                        agb_pass(ref, c + 1);
                        tell(o, std::string("play ") + cname[c] + " pass");
                    } else {
                        tell(c, std::string("play ") + cname[c] + " " + vertex(row, col));
                        tell(o, std::string("play ") + cname[c] + " " + vertex(row, col));
                        agb_play(ref, c + 1, row, col);
                        moved = true;
                    }
                }
                played[c] = moved;
                if (moved) stones++;
                if (moved) fprintf(stderr, "%3d %c %s\n", t + 1, cname[c], vertex(row, col).c_str());
                else fprintf(stderr, "%3d %c pass\n", t + 1, cname[c]);
                if (t == 2 * max_moves - 1) agb_calc_result(ref, &dscore);
            }
            for (int w = 0; w < 2; w++) side[w].stop();     // TerminateProcess, :562-563
            if (broken) { agb_destroy(ref); return 1; }
            dscores.push_back(dscore);
            if (dscore > 0) black_wins++;
            fprintf(stderr, "game %d: black - white = %d\n", g + 1, dscore);
        }
        printf("{\"match\": {\"games\": %d, \"board_size\": %d, \"black_playouts\": %lld, \"white_playouts\": %lld, "
               "\"black_policy\": \"%s\", \"white_policy\": \"%s\", \"players\": \"child processes over pipes\", "
               "\"black_wins\": %d, \"stones_played\": %lld, \"dscore\": [",
               match_games, cfg.board_size, P, white_P > 0 ? white_P : P,
               !black_cmd.empty() ? "external" : (cfg.genmove_static ? "reference" : "flat"),
               !white_cmd.empty() ? "external" : ((white_policy >= 0 ? white_policy : cfg.genmove_static) ? "reference" : "flat"),
               black_wins, stones);
        for (size_t i = 0; i < dscores.size(); i++) printf("%s%d", i ? ", " : "", dscores[i]);
        printf("]}}\n");
        agb_destroy(ref);
        return 0;
    }
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
