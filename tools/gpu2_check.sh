# 2-GPU checks of the multi-GPU paths (run under gpurun --gpus 2): gather test, GTP multi-GPU, torchrun bench configs 2 and 3
timeout 600 python -m pytest tests/test_gpu_api.py::test_batch_gather_on_two_gpus tests/test_gtp.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_try.json 2> gpurun_out/bench2_err.log; echo rc=$?
tail -3 gpurun_out/bench2_err.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --positions 64 --batch-playouts 4000 --steps 1 > gpurun_out/bench_c3_2gpu_try.json 2> gpurun_out/bench3_err.log; echo rc=$?
tail -3 gpurun_out/bench3_err.log
timeout 300 python bench.py --config 3 --positions 64 --batch-playouts 4000 --steps 1 --no-cpu-baseline > gpurun_out/bench_c3_1gpu_try.json 2>> gpurun_out/bench3_err.log; echo rc=$?
