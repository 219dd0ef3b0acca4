#!/bin/bash
# Build agentgo_b200/libagentgo_b200_<name>.so: the release objects with playout_group.cu recompiled
# under extra flags (A/B runs of kernel tunables: AGB_LIB=<that file> python tests/quick_bench.py group8).
# usage: tools/build_variant.sh <name> [nvcc flags ...]
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
obj=$here/agentgo_b200/build/rel
mkdir -p $obj/var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I $here/agentgo_b200/csrc -I $here/include \
  "$@" -x cu -c $here/agentgo_b200/csrc/playout_group.cu -o $obj/var/playout_group_$name.o
objs=$(ls $obj/*.o | grep -v playout_group.cu.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $here/agentgo_b200/libagentgo_b200_$name.so $objs $obj/var/playout_group_$name.o -lcudart -ldl
echo $here/agentgo_b200/libagentgo_b200_$name.so
