#!/bin/bash
# instrumented debug build of the library (phase clocks of the tensor-core backward); never shipped
set -e
cd "$(dirname "$0")/.."
SRC="abi.cu euler_exact.cu euler_bwd_exact.cu euler_bwd_tc.cu gru_bwd.cu gru_bwd_tc.cu enc_bwd.cu euler_tc.cu enc_tc.cu heads.cu"
OBJ=""
for f in $SRC; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DTRAJSDE_BWD_TIMELINE $EXTRA -c trajsde_b200/csrc/$f -o /tmp/tl_${f%.cu}.o &
  OBJ="$OBJ /tmp/tl_${f%.cu}.o"
done
wait
nvcc -shared -o ${OUT:-bench_micro/libtrajsde_b200_tl.so} $OBJ -gencode arch=compute_100a,code=sm_100a
