/*
 * oracle/ref_genmove_shim.h — TEST INFRASTRUCTURE ONLY.
 *
 * Second shim of oracle/build_ref.sh, used for oracle/_ref/liboracle_ref_genmove_13.so: the
 * reference's own move selection (AgentGo.cpp:15-705 — the globals, MyJob, MyWorker::work with its
 * static pattern heuristics, MyGame::testAvrgWin and MyGame::onMove with the opening book) compiled
 * where it lies, after the GoContest sources and GTPAdapter.h.  AgentGo.cpp as a whole cannot be
 * compiled on Linux (<Windows.h>, _tmain/_TCHAR/_wtof, the Win32 thread library TinyMT.*), so the
 * four types it takes from TinyMT.h / TinyOP.h are supplied here with the same interface and a
 * SERIAL schedule:
 *   TObject, TJob, SWorker                    TinyMT.h (bases only)
 *   Scheduler<W>(bool) submit/go/wait          TinyMT.h:733,860,751,574 — go() runs the jobs one
 *                                              after the other in submission order, so the
 *                                              reference's racy `double +=` on its global tables
 *                                              become a defined left-to-right sum
 *   TinyOP<T>(n) tnew/tdelete                  TinyOP.h (object pool) -> new / delete
 * oracle_job_index is the submission index of the job being run (= the candidate index of the seed
 * contract); oracle_seed_candidate / oracle_seed_root are called by the two lines the build inserts
 * at the top of the reference's playout loops (after AgentGo.cpp:288 and :419) to start the injected
 * per-playout TinyMT32 stream — the reference seeds libc rand() once per process instead.
 */
#include <math.h>
#include <vector>
#define protected public   /* GTPAdapter::bd (GTPAdapter.h:8) is driven by the harness; no behaviour change */

extern "C" void oracle_seed_candidate(int job, int playout);
extern "C" void oracle_seed_root(int playout);
static int oracle_job_index = 0;

namespace TinyMT {
struct TObject {
    virtual ~TObject() {}
};
struct TJob {
    virtual ~TJob() {}
};
struct SWorker {
    virtual void work(TJob* j) = 0;
    virtual ~SWorker() {}
};
template <typename W>
struct Scheduler {
    std::vector<TJob*> jobs;
    explicit Scheduler(bool) {}
    void submit(TJob* j, unsigned long) { jobs.push_back(j); }
    void go() {
        W w;
        for (size_t i = 0; i < jobs.size(); i++) {
            oracle_job_index = (int)i;
            w.work(jobs[i]);
        }
    }
    bool wait() { return true; }
};
}  // namespace TinyMT
template <typename T>
struct TinyOP {
    explicit TinyOP(int) {}
    T* tnew(T* old) { return new T(old); }
    void tdelete(T* p) { delete p; }
};
using namespace TinyMT;
