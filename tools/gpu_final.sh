#!/bin/bash
# last visit of a session: smoke, full GPU tests, bench line of both arms with the final binary.  usage: bash tools/gpu_final.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.txt
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.txt; cut -c1-200 gpurun_out/bench_$tag.json; tail -2 gpurun_out/bench_${tag}_err.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; cut -c1-120 gpurun_out/bench_ref_$tag.json
