python -m pytest tests/test_gpu_extract.py tests/test_gpu_workload.py -x -q 2>&1 | tail -3
python tools/fast_experiment.py 2>&1 | tail -2
EAOF_FAST_GENERIC=1 python tools/fast_experiment.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 2 -c 1 -o gpurun_out/r02_fast_rows4 -f python tools/fast_experiment.py > gpurun_out/ncu_fast_rows2.log 2>&1; tail -1 gpurun_out/ncu_fast_rows2.log
