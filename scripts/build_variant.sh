#!/bin/bash
# Tuning builds: scripts/build_variant.sh NAME "-DNB_SEL_CAPTURE=0 ..."  ->  ems-decoder-of-nb-ldpc-codes_b200/variants/NAME.so
# (selected at run time with NBLDPC_B200_LIB; never used by the tests or the default bench)
set -e
cd "$(dirname "$0")/../ems-decoder-of-nb-ldpc-codes_b200"
mkdir -p variants build/var_$1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I../include -Icsrc $2 -c csrc/nbldpc_cuda.cu -o build/var_$1/nbldpc_cuda.o
gcc -O2 -fPIC -ffp-contract=off -std=gnu11 -I../include -Icsrc -c csrc/nbldpc_host.c -o build/var_$1/nbldpc_host.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$1.so build/var_$1/nbldpc_host.o build/var_$1/nbldpc_cuda.o -lm
echo built variants/$1.so
