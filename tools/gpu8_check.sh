# 8-GPU checks (run under gpurun --gpus 8): the 2-GPU tests, the bench line at N = 8 with shard_check,
# strong_16M and genmove_1M, config 3 on 8 GPUs
timeout 300 python -m pytest tests/test_gpu_api.py::test_batch_gather_on_two_gpus tests/test_gtp.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench8_err.log; echo rc=$?
tail -2 gpurun_out/bench8_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 3 --steps 1 > gpurun_out/config3_r02_8gpu.json 2>> gpurun_out/bench8_err.log; echo rc=$?
timeout 200 python bench.py --impl reference --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_ref_r02_8gpu.json 2>> gpurun_out/bench8_err.log; echo rc=$?
timeout 100 python -c "import __graft_entry__ as g; g.smoke()"
