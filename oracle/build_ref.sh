#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE ONLY.
#
# Compiles the REFERENCE's own hot-path sources (GoContest/{Piece,SetNode,Board}.{h,cpp})
# with g++, from where they lie under $AGB_REFERENCE (default /root/reference), into
# oracle/_ref/liboracle_ref_<N>.so for N = 9, 13, 19.  The sources are streamed through
# a pipe into the compiler; no reference source is written into the repo.
#
# The .sln/.vcxproj are MSVC/Win32-only, so the reference's build system is not used.
# Semantics-preserving edits applied in the stream (each forced by a g++ failure,
# SURVEY.md App. A):
#   1. CRLF -> LF, GBK -> UTF-8 (comments only)
#   2. local #include "..." lines dropped (files are concatenated in dependency order;
#      AgentGo.h:4-5 use backslash paths to Config.h / DbgPipe.h, whose Win32 content
#      is replaced by ref_shim.h)
#   3. `#pragma once` dropped (single translation unit)
#   4. AgentGo.h:15  BOARD_SIZE 13 -> N   (unconditional #define; one build per size;
#      the literals 120 (Board.cpp:349) and ENDING_THRESHOLD 100 (Board.h:13) are kept)
#   5. Board.cpp:907 checkNoSense falls off the end on the false path (UB; g++ -O2 turns
#      the function into `return true`): rewritten as `return (at_islbd || near_islbd);`
#   rand -> oracle_rand and AG_PANIC/dprintf come from ref_shim.h.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${AGB_REFERENCE:-/root/reference}"
src="$ref/AgentGo/GoContest"
out="$here/_ref"
if [ ! -f "$src/Board.cpp" ]; then
    echo "build_ref: reference not found at $ref (prebuilt oracle/_ref/*.so are used if present)" >&2
    exit 3
fi
mkdir -p "$out"
CXX="${CXX:-g++}"
stream() {
    n="$1"
    cat "$here/ref_shim.h"
    for f in AgentGo.h Piece.h SetNode.h Board.h Piece.cpp SetNode.cpp Board.cpp; do
        echo "#line 1 \"reference:GoContest/$f\""
        tr -d '\r' < "$src/$f" | iconv -f GBK -t UTF-8 -c \
          | sed -e 's/^#include ".*$//' \
                -e 's/^#pragma once.*$//' \
                -e "s/^#define BOARD_SIZE[[:space:]]\+13/#define BOARD_SIZE $n/" \
                -e 's/if( at_islbd || near_islbd ) return true;/return ( at_islbd || near_islbd );/'
        echo
    done
    echo "#line 1 \"oracle/ref_harness.cpp\""
    cat "$here/ref_harness.cpp"
}
# The reference's move selection (AgentGo.cpp:15-705: globals, MyJob, MyWorker::work with the static
# pattern heuristics, MyGame::testAvrgWin / onMove with the opening book), 13x13 only (its literals).
# Additional stream edits, each forced by the build or by the harness:
#   6. lines 1-14 (Win32 includes, `using namespace TinyMT`) and 706-end (_tmain) are not streamed;
#      TinyMT.h / TinyOP.h types come from ref_genmove_shim.h with a serial schedule
#   7. `class X` -> `struct X` at line starts (the harness reads MyGame::step and calls onMove)
#   (GeneHatchery.cpp:379-436, the scorer CalcResults of the match harness, is streamed unedited)
#   8. one line inserted after AgentGo.cpp:288 and :419 (top of the two playout loops) that starts the
#      injected per-playout TinyMT32 stream (the reference seeds libc rand() once per process)
stream_genmove() {
    stream 13 | sed -e '/^#line 1 "oracle\/ref_harness.cpp"/,$d'
    cat "$here/ref_genmove_shim.h"
    echo "#line 1 \"reference:GTPAdapter.h\""
    tr -d '\r' < "$ref/AgentGo/GTPAdapter.h" | iconv -f GBK -t UTF-8 -c | sed -e 's/^#include ".*$//'
    echo
    echo "#line 15 \"reference:AgentGo.cpp\""
    tr -d '\r' < "$ref/AgentGo/AgentGo.cpp" | iconv -f GBK -t UTF-8 -c | sed -n '15,705p' \
      | sed -e 's/^class /struct /' \
            -e 's/^\([[:space:]]*\)mj->bd->clone(bdnew);/&oracle_seed_candidate(oracle_job_index, ii);/' \
            -e 's/^\([[:space:]]*\)bdtest.clone(bd);/&oracle_seed_root(ii);/'
    echo
    # the match scorer of the GA tuner: SET_SCORE and CalcResults, GeneHatchery.cpp:379-436
    echo "#line 379 \"reference:GeneHatchery.cpp\""
    tr -d '\r' < "$ref/GeneHatchery/GeneHatchery.cpp" | iconv -f GBK -t UTF-8 -c | sed -n '379,436p'
    echo
    echo "#line 1 \"oracle/ref_harness.cpp\""
    cat "$here/ref_harness.cpp"
    echo "#line 1 \"oracle/ref_genmove_harness.cpp\""
    cat "$here/ref_genmove_harness.cpp"
}
stream_genmove | "$CXX" -x c++ -std=c++11 -O2 -fPIC -shared -pthread -w \
    -I "$here" -I "$here/../include" -o "$out/liboracle_ref_genmove_13.so" -
echo "built $out/liboracle_ref_genmove_13.so"
for n in 9 13 19; do
    stream "$n" | "$CXX" -x c++ -std=c++11 -O2 -fPIC -shared -pthread -w \
        -I "$here" -I "$here/../include" -o "$out/liboracle_ref_$n.so" -
    echo "built $out/liboracle_ref_$n.so"
done
