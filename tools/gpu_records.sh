#!/bin/bash
# The commands behind profiles/rNN_*: run on a GPU box through gpurun, e.g.
#   gpurun --timeout 2400 -- 'bash tools/gpu_records.sh single r02 f'
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_records.sh multi r02 e 8'
#   gpurun --timeout 600 -- 'bash tools/gpu_records.sh ncu k_fast r02_fast_rows4 "python tools/fast_experiment.py"'
# Outputs land in gpurun_out/ (scratch); what is meant to be judged is copied to profiles/ by hand.
set -u
mode=$1
case "$mode" in
single)  # GPU test suite, bench, reference arm, launch list, sanitizer passes over the kernels written this round
    R=$2; T=$3
    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_gputests_${T}.log; cat gpurun_out/${R}_gputests_${T}.log
    timeout 600 python bench.py > gpurun_out/${R}_bench_n1_${T}.json 2> gpurun_out/${R}_bench_n1_${T}.err; head -c 200 gpurun_out/${R}_bench_n1_${T}.json
    timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref_${T}.json 2>/dev/null
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_${T}.csv python bench.py --steps 2 --warmup 3 \
        --no-cpu-baseline --no-sweep --no-determinism --no-other-configs --no-next-rows --sustained-seconds 0 > /dev/null 2>&1
    K="stages_match_oracle or adversarial or batch_equals_single or other_pyramid_parameters or (kernel_variants and 333) or thresholds_outside or bow or tensor_core"
    timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_extract.py tests/test_gpu_match.py tests/test_gpu_voc.py -x -q -k "$K" \
        > gpurun_out/${R}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
    timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_extract.py tests/test_gpu_match.py -x -q \
        -k "stages_match_oracle or batch_equals_single or (kernel_variants and 333 and (BULK or TMA or OCT_WIDTH)) or bow or tensor_core" \
        > gpurun_out/${R}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
    timeout 400 compute-sanitizer --tool initcheck --print-limit 400 --error-exitcode 9 python -m pytest tests/test_gpu_extract.py tests/test_gpu_match.py tests/test_gpu_ingest.py -x -q \
        -k "stages_match_oracle or batch_equals_single or adversarial or bow or tensor_core or ring or projection" \
        > gpurun_out/${R}_sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"
    ;;
multi)   # N GPUs of one box: multi-GPU pytest, sharded extraction + NCCL all-gather sweep against one GPU, host-feed ceiling, bench
    R=$2; T=$3; N=$4
    timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rs 2>&1 | tail -4 > gpurun_out/${R}_gputests_multi_n${N}_${T}.log
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    timeout 400 $TR --master-port 29551 tools/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK > gpurun_out/${R}_multi_gpu_check_n${N}_${T}.log; cat gpurun_out/${R}_multi_gpu_check_n${N}_${T}.log
    timeout 200 $TR --master-port 29553 tools/pcie_bw_concurrent.py 2>&1 | grep PCIE_CONCURRENT > gpurun_out/${R}_pcie_concurrent_n${N}.log
    timeout 700 $TR --master-port 29552 bench.py --gpus $N --steps 20 --warmup 3 --no-next-rows > gpurun_out/${R}_bench_n${N}_${T}.json 2> gpurun_out/${R}_bench_n${N}_${T}.err
    head -c 250 gpurun_out/${R}_bench_n${N}_${T}.json
    ;;
ncu)     # one `ncu --set full` capture of a kernel: read it with tools/ncu_summary.py, ncu_ops.py, ncu_lines.py, ncu_sass.py
    K=$2; OUT=$3; CMD=$4
    ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -o gpurun_out/$OUT -f $CMD > gpurun_out/$OUT.log 2>&1; tail -1 gpurun_out/$OUT.log
    ;;
*) echo "usage: gpu_records.sh single|multi|ncu ..."; exit 2;;
esac
