// gtp_adapter.h — GTP command surface of the reference, kept verbatim for this path.
//
// Mirrors AgentGo/GTPAdapter.{h,cpp}: same class name, same virtual hooks (onClear,
// onBoardSize, onPlay, onMove, onMoved — GTPAdapter.h:15-19), same command set and the same
// reply quirks (GTPAdapter.cpp:22-166).  The reference's `Board bd` member is replaced by the
// engine's replay board behind the C ABI (agb_play / agb_clear_board / agb_boardsize).
// Host code only: there is no CUDA in this file.
#ifndef AGB_GTP_ADAPTER_H_
#define AGB_GTP_ADAPTER_H_

#include <iosfwd>
#include <string>
#include <vector>

#include "agentgo_b200.h"

class GTPAdapter {
protected:
    agb_engine* eng;      // owns the board (reference: `Board bd`, GTPAdapter.h:8)
    float komi;           // stored, never read (GTPAdapter.cpp:120-124), as in the reference
    double avgtime;       // running mean of genmove wall time in ms (GO_TIME_STAT, GTPAdapter.cpp:62-73)
    long times;
    std::string move_error;   // set by onMove when no move could be produced: the reply is "? <text>", not "= pass"

private:
    virtual void onClear() = 0;
    virtual void onBoardSize(int sz) { (void)sz; }
    virtual void onPlay(int isW, int a, int b) = 0;
    virtual bool onMove(int isW, int& a, int& b) = 0;   // false => pass
    virtual void onMoved(int isW, int a, int b) = 0;

public:
    explicit GTPAdapter(agb_engine* e) : eng(e), komi(0), avgtime(0), times(0) {}
    virtual ~GTPAdapter() {}
    // reads commands until `quit` or end of input; one reply per command
    void MainLoop(std::istream& in, std::ostream& out);
    // handles one line; returns false after `quit`
    bool HandleLine(const std::string& line, std::ostream& out);
    double AvgTimeMs() const { return avgtime; }
};

// letter <-> row index, skipping 'I' (GaIntToChar / GaCharToInt, GTPAdapter.cpp:19-20)
char GaIntToChar(int i);
int GaCharToInt(int c);

// The engine-backed player (reference: class MyGame, AgentGo.cpp:351-705).
class MyGame : public GTPAdapter {
public:
    MyGame(agb_engine* e, long long playouts_per_candidate, long long root_playouts, unsigned long long seed_base)
        : GTPAdapter(e), step(0), P_cand(playouts_per_candidate), P_root(root_playouts), seed(seed_base) {
        last.device_ms = 0; last.playouts = 0;
    }
    agb_stats last;       // statistics of the last genmove
    void SetSeed(unsigned long long s) { seed = s; }   // the reference reseeds per process (srand(time(0)), AgentGo.cpp:724)
    // further engines of a multi-GPU group (agb_attach_nccl): they mirror every board change and
    // run the same genmove concurrently, one host thread each
    std::vector<agb_engine*> peers;

private:
    int step;
    long long P_cand, P_root;
    unsigned long long seed;
    void onClear() override;
    void onBoardSize(int sz) override;
    void onPlay(int isW, int a, int b) override;
    bool onMove(int isW, int& a, int& b) override;
    void onMoved(int isW, int a, int b) override;
};

#endif
