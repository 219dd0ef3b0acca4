// tests/csrc/emu_main.cpp — TEST-ONLY: command-line driver of the emulated group kernel (tests/csrc/emu_group.cpp),
// built with -fsanitize=address by tests/test_emu_group.py so that any access of the kernel past its warp's shared
// memory or its empty-point arrays is reported.  usage: emu_main <board size> <lanes per playout> <playouts>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
extern "C" long long emu_group_playouts(int n, int lanes, const uint32_t* moves, int n_moves, int colour,
                                        int cand_row, int cand_col, uint32_t position_id, uint32_t cand_idx,
                                        int64_t P, uint64_t seed_base, int avrg_win, int max_moves,
                                        int16_t* diffs, int64_t* out3, int64_t* stat4, int64_t* amaf);
int main(int argc, char** argv) {
    int n = atoi(argv[1]), lanes = atoi(argv[2]), P = atoi(argv[3]);
    int16_t diffs[4096]; int64_t o[3], s[4];
    long long r = emu_group_playouts(n, lanes, nullptr, 0, 1, -1, -1, 0, 0, P, 20151001, 0, 0, diffs, o, s, nullptr);
    printf("r=%lld wins %lld moves %lld draws %lld playouts %lld\n", r, (long long)o[0], (long long)s[0], (long long)s[1], (long long)s[3]);
    for (int i = 0; i < P && i < 16; i++) printf("%d ", diffs[i]);
    printf("\n");
}
