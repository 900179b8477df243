for v in base var base var; do
  if [ $v = base ]; then unset TRAJSDE_LIB_PATH; else export TRAJSDE_LIB_PATH=$PWD/bench_micro/libtrajsde_b200_$v.so; fi
  echo "== $v"; timeout 100 python tools/bench_dec.py 204800 20; 
done
export TRAJSDE_LIB_PATH=$PWD/bench_micro/libtrajsde_b200_var.so
timeout 200 python -m pytest tests/test_euler_fwd_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/train_prof.py 1024 5 | tail -1
