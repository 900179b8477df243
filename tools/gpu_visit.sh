#!/bin/bash
# One GPU visit, run under gpurun from the repo root:   gpurun --timeout 1500 -- 'bash tools/gpu_visit.sh <tag> [tests] [bench] [launches] [ncu] [sweep]'
# Every artefact lands in gpurun_out/ with the tag in its name; copy what should be judged into profiles/ afterwards (tools/collect_profiles.sh).
tag=${1:-x}; shift
what=${@:-tests bench}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests)    timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.txt
              timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ;;
    bench)    timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.txt; cut -c1-240 gpurun_out/bench_$tag.json; tail -2 gpurun_out/bench_${tag}_err.txt
              timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>/dev/null; cut -c1-160 gpurun_out/bench_ref_$tag.json ;;
    launches) # every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
              timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
                python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-scenes 0 > gpurun_out/bench_under_ncu_$tag.log 2>&1; wc -l gpurun_out/launches_$tag.csv ;;
    ncu)      # full captures of the dominant kernels (one launch each, after warm-up launches of the same kernel)
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_fwd_dw \
                python tools/bench_dec.py 204800 2 dw > /dev/null 2>&1
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_fwd_philox \
                python tools/bench_dec.py 204800 2 philox > /dev/null 2>&1
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_enc_fwd \
                python tools/bench_enc.py 1024 2 > /dev/null 2>&1
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:euler_bwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_bwd_tc \
                python tools/train_prof.py 1024 1 > /dev/null 2>&1
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_bwd_sweep_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_enc_bwd_sweep \
                python tools/train_prof.py 1024 1 > /dev/null 2>&1
              timeout 300 ncu --set full --clock-control none --import-source on -k regex:heads_fwd_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_heads_fwd \
                python tools/bench_heads.py 204800 2 > /dev/null 2>&1
              ls -la gpurun_out/${tag}_*.ncu-rep ;;
    sweep)    timeout 900 python bench_sweep.py --out gpurun_out/sweep_$tag.json > /dev/null 2> gpurun_out/sweep_${tag}_err.txt; tail -2 gpurun_out/sweep_${tag}_err.txt; ls -la gpurun_out/sweep_$tag.json ;;
  esac
done
