# ncu captures of round 2 (run under gpurun on one GPU): the three mappings of the group kernel, the
# warp kernel, and the peak micro-kernels of agb_measure_peaks
set -x
for k in group8 group16 group32 warp; do
  python tests/prof_run.py $k 100000 > gpurun_out/plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:playout_ -s 1 -c 1 -o gpurun_out/r02_$k -f python tests/prof_run.py $k 100000 > gpurun_out/ncu_$k.log 2>&1
  tail -n 1 gpurun_out/plain_$k.log
done
python -c "import agentgo_b200 as ag; print(ag.measure_peaks(0))" > gpurun_out/plain_peaks.log 2>&1 && \
ncu --set full --clock-control none -k regex:peak_kernel -s 3 -c 9 -o gpurun_out/r02_peaks -f python -c "import agentgo_b200 as ag; print(ag.measure_peaks(0))" > gpurun_out/ncu_peaks.log 2>&1
cat gpurun_out/plain_peaks.log
