/*
 * oracle/tinymt32.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * TinyMT32 (Saito & Matsumoto; RFC 8682 "TinyMT32 Pseudorandom Number Generator")
 * written from the published algorithm.  The reference (Menooker/AgentGo) draws
 * playout randomness from libc rand() (GoContest/Board.cpp:360,361,386,387,447),
 * seeded once with srand(time(0)) (AgentGo.cpp:724) — i.e. from the MSVC 2010 C
 * runtime, which is not in the reference checkout and is not reproducible.
 * SURVEY.md §0 P1 fixes the substitute: one TinyMT32 stream per playout and
 *     rand() := tinymt32_generate_uint32() >> 17     (15 bit = MSVC RAND_MAX 0x7fff)
 * Known answer (RFC 8682 parameter set, seed 1): 2545341989 981918433 3715302833
 * 2387538352 3591001365 — checked in tests/test_oracle.py.
 */
#ifndef ORACLE_TINYMT32_H_
#define ORACLE_TINYMT32_H_
#include <stdint.h>

typedef struct {
    uint32_t status[4];
    uint32_t mat1, mat2, tmat;
} oracle_tinymt32_t;

#define ORACLE_TINYMT32_MASK 0x7fffffffu
#define ORACLE_TINYMT32_SH0 1
#define ORACLE_TINYMT32_SH1 10
#define ORACLE_TINYMT32_SH8 8

static inline void oracle_tinymt32_next_state(oracle_tinymt32_t* r) {
    uint32_t x, y;
    y = r->status[3];
    x = (r->status[0] & ORACLE_TINYMT32_MASK) ^ r->status[1] ^ r->status[2];
    x ^= (x << ORACLE_TINYMT32_SH0);
    y ^= (y >> ORACLE_TINYMT32_SH0) ^ x;
    r->status[0] = r->status[1];
    r->status[1] = r->status[2];
    r->status[2] = x ^ (y << ORACLE_TINYMT32_SH1);
    r->status[3] = y;
    if (y & 1) {
        r->status[1] ^= r->mat1;
        r->status[2] ^= r->mat2;
    }
}

static inline uint32_t oracle_tinymt32_temper(const oracle_tinymt32_t* r) {
    uint32_t t0, t1;
    t0 = r->status[3];
    t1 = r->status[0] + (r->status[2] >> ORACLE_TINYMT32_SH8);
    t0 ^= t1;
    if (t1 & 1) t0 ^= r->tmat;
    return t0;
}

static inline uint32_t oracle_tinymt32_generate_uint32(oracle_tinymt32_t* r) {
    oracle_tinymt32_next_state(r);
    return oracle_tinymt32_temper(r);
}

static inline void oracle_tinymt32_init(oracle_tinymt32_t* r, uint32_t seed) {
    int i;
    r->status[0] = seed;
    r->status[1] = r->mat1;
    r->status[2] = r->mat2;
    r->status[3] = r->tmat;
    for (i = 1; i < 8; i++) {
        r->status[i & 3] ^= (uint32_t)i + 1812433253u *
            (r->status[(i - 1) & 3] ^ (r->status[(i - 1) & 3] >> 30));
    }
    /* period certification */
    if ((r->status[0] & ORACLE_TINYMT32_MASK) == 0 && r->status[1] == 0 &&
        r->status[2] == 0 && r->status[3] == 0) {
        r->status[0] = 'T'; r->status[1] = 'I'; r->status[2] = 'N'; r->status[3] = 'Y';
    }
    for (i = 0; i < 8; i++) oracle_tinymt32_next_state(r);
}

#endif
