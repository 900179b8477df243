#!/bin/bash
# experimental build of the library with extra -D flags (A/B kernel experiments; never shipped):
#   EXTRA="-DTRAJSDE_MBAR_SUSPEND_ALL" OUT=bench_micro/libtrajsde_b200_var.so bash bench_micro/build_variant_lib.sh
set -e
cd "$(dirname "$0")/.."
SRC="abi.cu euler_exact.cu euler_bwd_exact.cu euler_bwd_tc.cu gru_bwd.cu gru_bwd_tc.cu enc_bwd.cu enc_bwd_sweep.cu euler_tc.cu enc_tc.cu heads.cu heads_bwd.cu stage_ops.cu"
OBJ=""
for f in $SRC; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $EXTRA -c trajsde_b200/csrc/$f -o /tmp/var_${f%.cu}.o &
  OBJ="$OBJ /tmp/var_${f%.cu}.o"
done
wait
nvcc -shared -o ${OUT:-bench_micro/libtrajsde_b200_var.so} $OBJ -gencode arch=compute_100a,code=sm_100a
