#!/usr/bin/env bash
# quick GPU iteration: parity subset, timing, and the warp-instruction count per playout
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python tests/prof_run.py warp 200000
python tests/prof_run.py warp 400000 9
python tests/prof_run.py warp 50000 > /dev/null 2>&1 && \
ncu --metrics smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:playout_warp -s 1 -c 1 python tests/prof_run.py warp 50000 2>&1 | grep -E "inst_executed|alu_cycles" | awk '{print $0} /smsp__inst_executed.sum/{gsub(",","",$3); printf("warp-instr per playout: %.0f\n", $3/50000)}'
