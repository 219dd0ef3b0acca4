// host_board.h — host-side replay board of the engine.
//
// Plays the role of `Board bd` owned by the reference's GTPAdapter (GTPAdapter.h:8): it is
// mutated by play / clear_board / boardsize and is the root every playout starts from.
// Because the playout policy observes history-dependent orders (string piece order, empty
// list order), a position is defined by its move list; this class replays moves into the
// exact compact state (board_ops.h) that is uploaded to the GPU.
#ifndef AGB_HOST_BOARD_H_
#define AGB_HOST_BOARD_H_

#include <stdint.h>

#include "board_ops.h"

namespace agb {

class HostBoard {
public:
    explicit HostBoard(int n = 13);

    static bool size_supported(int n) { return n == 9 || n == 13 || n == 19; }

    void reset(int n);           // boardsize + clear
    void clear();                // Board::clear
    int n() const { return pos_.n; }
    int nn() const { return pos_.n * pos_.n; }

    // 0 ok, -1 invalid colour/coordinates, -2 occupied
    int play(int colour, int row, int col);
    void pass(int colour);
    // replays AGB_MOVE-encoded moves; returns 0 or the failing status
    int replay(const uint32_t* moves, int n_moves);

    int colour_at(int idx) const;
    // which: 0 suicide 1 kill 2 survive 3 dying 4 chase 5 true eye 6 compete 7 no-sense
    int check(int which, int colour, int row, int col) const;
    int count_score(int colour) const;
    // CalcResults of the reference's match harness, GeneHatchery.cpp:393-436: black - white
    int calc_result() const;
    // Board::getRandomPiece on this board with a TinyMT32 stream seeded by `seed`; idx or -1
    int random_piece(int colour, uint32_t seed, const RngParams& prm);
    // the candidate filter of MyGame::onMove, AgentGo.cpp:586-592
    bool is_candidate(int colour, int row, int col) const;
    // the argmax filter of MyGame::onMove, AgentGo.cpp:651 (no checkNoSense there)
    bool is_eligible(int colour, int row, int col) const;

    // canonical dump shared with the oracle (16 + 7*NN int32 words)
    int export_state(int32_t* out, int cap) const;

    // Static pattern marks of MyWorker::work for the candidate (row, col) of `colour`
    // (AgentGo.cpp:104-282), as integers: total_mark[row][col] receives
    //     units_a * bonus_a + units_b * bonus_b,   bonus_x = int(numset * index_x)   (:95-96).
    // Every quirk of the reference is kept (see the definition).  The point must be empty.
    void static_units(int colour, int row, int col, int64_t* units_a, int64_t* units_b) const;
    // The 13x13 opening book of MyGame::onMove for its first four calls (AgentGo.cpp:468-572).
    // false when the board is not 13x13 (the book is made of 13x13 literals).
    bool book_move(int* row, int* col) const;

    const Position& position() const { return pos_; }
    // copy of the position after `colour` plays idx (AgentGo.cpp:83)
    Position with_move(int colour, int idx) const;

private:
    typedef BoardRef<1, 0> Ref;
    Ref open() const;
    void close(const Ref& r);
    bool no_sense(int colour, int row, int col) const;
    void sync_space_shadow();
    int data_at(int row, int col) const;
    Position pos_;
    // Board::space_list as the reference leaves it in memory: the links of the empty points are the
    // live list (= pos_), those of occupied points keep the values they had when the point was
    // taken (Board.cpp:105-118 does not clear them).  Only static_units looks at it, through the
    // reference's out-of-range reads of Board::data (see data_at).
    int16_t space_shadow_[kMaxNN][2];
};

}  // namespace agb

#endif
