/*
 * oracle/ref_shim.h — TEST INFRASTRUCTURE ONLY.
 *
 * Prepended (by oracle/build_ref.sh) to the reference's own GoContest sources when
 * they are compiled, where they lie under /root/reference, into oracle/_ref/.
 * It supplies what the MSVC build got from elsewhere:
 *   - the libc headers Board.cpp relies on transitively;
 *   - AG_PANIC (Config.h:3 is MSVC inline asm `__asm int 3`)  -> abort();
 *   - dprintf/dinitdbg (DbgPipe.h:8-9 map to a Win32 pipe)     -> no-ops;
 *   - rand  -> oracle_rand, the injection point of the per-playout TinyMT32;
 *   - `private` -> `public` so the harness can export set_nodes / game_ending
 *     for state-parity tests (no behaviour change).
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <memory.h>
#include <time.h>
#include <stack>
#include <thread>
#include <vector>
#include <atomic>

extern "C" int oracle_rand(void);

#define AG_DBG
#define AG_PANIC(a) abort()
#define dinitdbg(...) ((void)0)
#define dprintf(...) ((void)0)
#define rand oracle_rand
#define private public
