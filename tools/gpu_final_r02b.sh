# refresh of the round-2 headline measurements after a group8-only kernel change (run under gpurun):
# tests, the bench line of both arms, the ncu launch list of the bench command, the ncu capture of group8
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r02.json 2>> gpurun_out/bench_r02.err; echo ref rc=$?
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_plain.json 2>> gpurun_out/bench_r02.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
python tests/prof_run.py group8 100000 > gpurun_out/plain_group8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:playout_ -s 1 -c 1 -o gpurun_out/r02_group8 -f python tests/prof_run.py group8 100000 > gpurun_out/ncu_group8.log 2>&1
tail -n 1 gpurun_out/plain_group8.log
timeout 200 python tests/quick_bench.py warp group8 group16 group32 thread 2>&1 | grep "^kernel" > gpurun_out/quick_bench_r02.txt; cat gpurun_out/quick_bench_r02.txt
tail -3 gpurun_out/bench_r02.err
