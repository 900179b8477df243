#!/bin/bash
# GPU visit: heads profile + full tests + bench.  usage: bash tools/gpu_round2.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.txt
timeout 120 python tools/bench_heads.py 204800 5 | tee gpurun_out/heads_$tag.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:heads_fwd_kernel -c 1 -s 2 -f -o gpurun_out/prof_heads_$tag \
  python tools/bench_heads.py 204800 1 > gpurun_out/ncu_heads_$tag.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.txt; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_${tag}_err.txt
