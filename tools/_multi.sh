N=$1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rs 2>&1 | tail -4 > gpurun_out/r02_gputests_multi_n${N}_d.log; cat gpurun_out/r02_gputests_multi_n${N}_d.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|warn\|^\*\*\*" | tail -8 > gpurun_out/r02_multi_gpu_check_n${N}_d.log; cat gpurun_out/r02_multi_gpu_check_n${N}_d.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n${N}_d.json 2> gpurun_out/r02_bench_n${N}_d.err; tail -c 300 gpurun_out/r02_bench_n${N}_d.err; head -c 300 gpurun_out/r02_bench_n${N}_d.json
