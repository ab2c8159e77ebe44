#!/bin/sh
# Builds egg_b200/libeggphot_<name>.so with extra nvcc flags for the kernels (e.g. timing: -DEGG_PHASE_TIMING).
# usage: tools/build_variant.sh <name> [flags...];  run with EGGPHOT_LIB=egg_b200/libeggphot_<name>.so python tools/prof_run.py
set -e
name=$1; shift
cd "$(dirname "$0")/../egg_b200/csrc"
make -s lib
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 20054,177"
$NV "$@" -c phot_kernels.cu -o /tmp/pk_$name.o
$NV -shared -o ../libeggphot_$name.so /tmp/pk_$name.o eggphot_capi.o phot_tables.o -lcudart_static -lpthread -ldl -lrt
