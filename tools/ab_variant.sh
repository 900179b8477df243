# A/B of the shipped library against a variant build (first: EXTRA="-D..." bash bench_micro/build_variant_lib.sh); run under gpurun from the repo root
for v in base var; do
  if [ $v = var ]; then export TRAJSDE_LIB_PATH=$PWD/bench_micro/libtrajsde_b200_var.so; else unset TRAJSDE_LIB_PATH; fi
  echo "== $v"
  timeout 100 python tools/bench_heads.py 204800 5 | head -1
  timeout 100 python tools/bench_dec.py 204800 5
  timeout 100 python tools/bench_enc.py 1024 10 | head -2
  timeout 100 python tools/train_prof.py 1024 5
done
