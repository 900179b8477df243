#!/bin/bash
# Final evidence pass (1 GPU): launch list of the bench command, full ncu captures of the backward kernels, bench lines.  usage: bash tools/gpu_round4.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler_bwd_tc_kernel -c 1 -s 64 -f -o gpurun_out/prof_bwd_tc_$tag \
  python tools/train_prof.py 1024 1 > gpurun_out/ncu_bwd_$tag.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.txt; cut -c1-300 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_${tag}_err.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$tag.json
