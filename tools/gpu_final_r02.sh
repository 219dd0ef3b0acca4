# round-2 measurements on one GPU (run under gpurun): tests, the bench line of both arms, the ncu launch
# list of the bench command, ncu captures of the mappings and of the peak micro-kernels, configs 1, 3 and 4
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r02.json 2>> gpurun_out/bench_r02.err; echo ref rc=$?
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_plain.json 2>> gpurun_out/bench_r02.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
for k in group8 group16 group32 warp; do
  python tests/prof_run.py $k 100000 > gpurun_out/plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:playout_ -s 1 -c 1 -o gpurun_out/r02_$k -f python tests/prof_run.py $k 100000 > gpurun_out/ncu_$k.log 2>&1
  tail -n 1 gpurun_out/plain_$k.log
done
python -c "import agentgo_b200 as ag; print(ag.measure_peaks(0))" > gpurun_out/plain_peaks.log 2>&1 && \
ncu --set full --clock-control none -k regex:peak_kernel -s 3 -c 9 -o gpurun_out/r02_peaks -f python -c "import agentgo_b200 as ag; print(ag.measure_peaks(0))" > gpurun_out/ncu_peaks.log 2>&1
timeout 300 python bench.py --config 1 --steps 5 > gpurun_out/config1_r02.json 2>> gpurun_out/bench_r02.err; echo c1 rc=$?
timeout 600 python bench.py --config 3 --steps 1 > gpurun_out/config3_r02.json 2>> gpurun_out/bench_r02.err; echo c3 rc=$?
timeout 600 python bench.py --config 4 > gpurun_out/selfplay_r02.json 2>> gpurun_out/bench_r02.err; echo c4 rc=$?
timeout 200 python tests/quick_bench.py warp group8 group16 group32 thread 2>&1 | grep "^kernel" > gpurun_out/quick_bench_r02.txt; cat gpurun_out/quick_bench_r02.txt
tail -5 gpurun_out/bench_r02.err
