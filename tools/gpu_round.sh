#!/bin/bash
# One GPU-box visit: parity tests, kernel timings, full ncu captures of the forward kernels, launch list, bench line.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.txt
python tools/bench_dec.py 204800 5 > gpurun_out/timing_$tag.txt 2>&1
python tools/bench_enc.py 1024 10 >> gpurun_out/timing_$tag.txt 2>&1
cat gpurun_out/timing_$tag.txt
if [ -z "$NO_NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:euler_fwd_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_fwd_tc_$tag \
  python tools/bench_dec.py 204800 1 dw > gpurun_out/ncu_fwd_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:euler_fwd_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_fwd_tc_philox_$tag \
  python tools/bench_dec.py 204800 1 philox > gpurun_out/ncu_fwd_philox_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:enc_fwd_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_enc_fwd_$tag \
  python tools/bench_enc.py 1024 1 > gpurun_out/ncu_enc_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$tag.log 2>&1
fi
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.txt; cat gpurun_out/bench_$tag.json
