// gtp_adapter.cpp — see gtp_adapter.h.  Wire format follows AgentGo/GTPAdapter.cpp line by line:
//   * no command ids; success is "= <payload>\n\n" for genmove/play/komi/version/name/
//     protocol_version/list_commands (:79,83,105,112,123,127,131,135,139) but "=\n\n" (no space)
//     for boardsize/clear_board/quit (:147,153,157);
//   * failure is "? unknown command : <line>\n\n" (:57,102,116,163);
//   * keywords are matched case-insensitively except `play` and `komi`, which go through sscanf
//     (:86,107,120); lines are cut at 198 characters (:24,31); empty lines are ignored (:33);
//   * `play <c> pass` — in fact any `play <c> ...` that is not a vertex — is acknowledged and
//     leaves the board untouched, it does not call Board::pass (:107-113).
#include "gtp_adapter.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include <chrono>
#include <iostream>
#include <thread>

#define STR_VERSION "0.0.1"
#define STR_NAME "AgentGo"
#define STR_PRO_VERSION "2"

char GaIntToChar(int i) { return (char)((i <= 'H' - 'A') ? i + 'A' : i + 'A' + 1); }
int GaCharToInt(int c) { return (c <= 'h') ? c - 'a' : c - 'a' - 1; }

static bool seqi(const char* a, const char* b) { return strcasecmp(a, b) == 0; }
static bool seqn(const char* a, const char* b, size_t n) { return strncasecmp(a, b, n) == 0; }

void GTPAdapter::MainLoop(std::istream& in, std::ostream& out) {
    std::string line;
    while (std::getline(in, line)) {
        if (!HandleLine(line, out)) return;
    }
}

bool GTPAdapter::HandleLine(const std::string& line, std::ostream& out) {
    char buf[200];
    size_t n = line.size() < 198 ? line.size() : 198;
    memcpy(buf, line.data(), n);
    buf[n] = 0;
    if (n > 0 && buf[n - 1] == '\r') buf[--n] = 0;
    int c = 0, a = 0, b = 0;
    float k = 0.0f;
    char cc = 0, ca = 0;
    if (buf[0] == 0) return true;
    if (seqn(buf, "genmove ", 8)) {
        bool played;
        const auto t0 = std::chrono::steady_clock::now();
        const int color = tolower((int)buf[8]);
        int isW;
        if (color == 'w') isW = 1;
        else if (color == 'b') isW = 0;
        else {
            out << "? " << "unknown command : " << buf << "\n\n" << std::flush;
            return true;
        }
        move_error.clear();
        played = onMove(isW, a, b);
        if (!move_error.empty()) {
            // the engine could not produce a move (CUDA error, move-cap fault, no device): that is a
            // failure, not a pass — the reply is GTP's failure line and the board stays as it was
            out << "? " << move_error << "\n\n" << std::flush;
            return true;
        }
        if (played) {
            if (agb_play(eng, isW ? AGB_WHITE : AGB_BLACK, a, b) != AGB_OK) {
                out << "? " << "genmove produced " << GaIntToChar(a) << b + 1 << " which cannot be played: "
                    << agb_last_error(eng) << "\n\n" << std::flush;
                return true;
            }
            onMoved(isW, a, b);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            avgtime = (avgtime * times + ms) / (times + 1);
            times++;
            out << "= " << GaIntToChar(a) << b + 1 << "\n\n" << std::flush;
        } else {
            out << "= " << "pass" << "\n\n" << std::flush;
        }
    } else if (sscanf(buf, "play %c %c%d", &cc, &ca, &b) == 3) {
        a = tolower(ca);
        c = tolower(cc);
        if (c == 'w' || c == 'b') {
            const int isW = c == 'w';
            // the reference would AG_PANIC on an occupied or off-board point (Board.cpp:74-83);
            // here the move is refused with the reference's failure line
            if (agb_play(eng, isW ? AGB_WHITE : AGB_BLACK, GaCharToInt(a), b - 1) != AGB_OK) {
                out << "? " << "unknown command : " << buf << "\n\n" << std::flush;
                return true;
            }
            onPlay(isW, GaCharToInt(a), b - 1);
        } else {
            out << "? " << "unknown command : " << buf << "\n\n" << std::flush;
            return true;
        }
        out << "= " << "\n\n" << std::flush;
    } else if (sscanf(buf, "play %c pass", &cc) == 1) {
        c = tolower(cc);
        if (c == 'w' || c == 'b') out << "= " << "\n\n" << std::flush;
        else out << "? " << "unknown command : " << buf << "\n\n" << std::flush;
    } else if (sscanf(buf, "komi %f", &k) == 1) {
        komi = k;
        out << "= " << "\n\n" << std::flush;
    } else if (seqi(buf, "version")) {
        out << "= " << STR_VERSION << "\n\n" << std::flush;
    } else if (seqi(buf, "list_commands")) {
        out << "= " << "protocol_version\nname\nversion\nlist_commands\nquit\nboardsize\nclear_board\nkomi\nplay\ngenmove"
            << "\n\n" << std::flush;
    } else if (seqi(buf, "name")) {
        out << "= " << STR_NAME << "\n\n" << std::flush;
    } else if (seqi(buf, "protocol_version")) {
        out << "= " << STR_PRO_VERSION << "\n\n" << std::flush;
    } else if (seqn(buf, "boardsize ", 10)) {
        if (strlen(buf) > 10) onBoardSize(atoi(&buf[10]));
        out << "=" << "\n\n" << std::flush;
    } else if (seqi(buf, "clear_board")) {
        agb_clear_board(eng);
        onClear();
        out << "=" << "\n\n" << std::flush;
    } else if (seqi(buf, "quit")) {
        out << "=" << "\n\n" << std::flush;
        return false;
    } else {
        out << "? " << "unknown command : " << buf << "\n\n" << std::flush;
    }
    return true;
}

// The reference's onBoardSize is an empty stub (AgentGo.cpp:387-390): `boardsize 19` is
// acknowledged and ignored because BOARD_SIZE is a compile-time 13.  Here a supported size
// (9, 13, 19) is honoured; anything else is acknowledged and ignored like the reference does.
void MyGame::onBoardSize(int sz) {
    if (agb_boardsize(eng, sz) == AGB_OK) {
        step = 0;
        for (agb_engine* p : peers) agb_boardsize(p, sz);
    }
}

void MyGame::onClear() {
    step = 0;
    for (agb_engine* p : peers) agb_clear_board(p);
}

// the peers mirror the first engine's board; one that cannot follow would play another game
static void mirror(const std::vector<agb_engine*>& peers, int colour, int a, int b) {
    for (agb_engine* p : peers)
        if (agb_play(p, colour, a, b) != AGB_OK) {
            fprintf(stderr, "a peer engine refused %d %d: %s\n", a, b, agb_last_error(p));
            exit(1);
        }
}

void MyGame::onPlay(int isW, int a, int b) { mirror(peers, isW + 1, a, b); }

void MyGame::onMoved(int isW, int a, int b) { mirror(peers, isW + 1, a, b); }

// MyGame::onMove (AgentGo.cpp:449-693): avrg_win from root playouts, the candidate filter,
// playouts per candidate, argmax — all inside agb_genmove, which also holds the 13x13 opening
// book of steps 1-4 (:468-572), the static pattern marks (:104-282) and the AMAF terms
// (:329-338) behind agb_config::genmove_book / genmove_static / genmove_amaf.
bool MyGame::onMove(int isW, int& a, int& b) {
    step += 1;
    int row = -1, col = -1, is_pass = 1;
    const unsigned long long sd = seed + 2ull * (unsigned long long)step;
    // every engine of a multi-GPU group evaluates its shard of the same genmove; the NCCL hook
    // inside sums the counters, so all of them return the same move
    struct Res { int st, row, col, pass; };
    std::vector<Res> pres(peers.size(), Res{0, -1, -1, 1});
    std::vector<std::thread> th;
    for (size_t i = 0; i < peers.size(); i++)
        th.emplace_back([=, &pres]() {
            Res& r = pres[i];
            r.st = agb_genmove(peers[i], isW + 1, P_cand, P_root, sd, &r.row, &r.col, &r.pass, nullptr);
        });
    const int st = agb_genmove(eng, isW + 1, P_cand, P_root, sd, &row, &col, &is_pass, &last);
    for (std::thread& t : th) t.join();
    if (st != AGB_OK) {
        move_error = std::string("genmove failed: ") + agb_last_error(eng);
        fprintf(stderr, "agb_genmove failed: %s\n", agb_last_error(eng));
        return false;
    }
    for (size_t i = 0; i < peers.size(); i++) {
        if (pres[i].st != AGB_OK) {
            move_error = std::string("genmove failed on a peer engine: ") + agb_last_error(peers[i]);
            return false;
        }
        if (pres[i].pass != is_pass || (!is_pass && (pres[i].row != row || pres[i].col != col))) {
            move_error = "the engines of the group disagree on the move";
            return false;
        }
    }
    if (is_pass) return false;
    a = row;
    b = col;
    return true;
}
