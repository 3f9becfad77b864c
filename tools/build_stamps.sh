#!/bin/bash
# debug variant of the library with in-kernel %globaltimer stamps (tools/stamp_gemm.py); never loaded by the product path
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/dm_stamps
for f in dm_api dm_gemm dm_elem; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -DDM_STAMPS \
    -c disentangle_mlp_b200/csrc/$f.cu -o /tmp/dm_stamps/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o disentangle_mlp_b200/lib/libdm_b200_stamps.so /tmp/dm_stamps/*.o
