for v in base var; do
  if [ $v = base ]; then unset TRAJSDE_LIB_PATH; else export TRAJSDE_LIB_PATH=$PWD/bench_micro/libtrajsde_b200_$v.so; fi
  echo "== $v"; timeout 100 python tools/bench_dec.py 204800 10 philox; timeout 100 python tools/bench_enc.py 1024 10 | head -3; timeout 100 python tools/train_prof.py 1024 5
done
