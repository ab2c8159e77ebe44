#!/bin/sh
# Builds egg_b200/libeggphot_<name>.so with extra nvcc flags for the kernels (e.g. -DEGG_WALK_THREADS=608).
# usage: tools/build_variant.sh <name> [flags...];  run with EGGPHOT_LIB=egg_b200/libeggphot_<name>.so python bench.py ...
set -e
name=$1; shift
cd "$(dirname "$0")/../egg_b200/csrc"
make -s lib
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -diag-suppress 20054,177"
$NV "$@" -c phot_kernels.cu -o /tmp/pk_$name.o &
$NV "$@" -c phot_lane_kernels.cu -o /tmp/pl_$name.o &
$NV "$@" -c eggphot_capi.cu -o /tmp/pc_$name.o &
DEFS=""; for f in "$@"; do case "$f" in -D*) DEFS="$DEFS $f";; esac; done
g++ -std=c++17 -O2 -fPIC $DEFS -c phot_tables.cpp -o /tmp/pt_$name.o   # (the shared-memory budget of the tables depends on EGG_WALK_THREADS)
wait
$NV -shared -o ../libeggphot_$name.so /tmp/pk_$name.o /tmp/pl_$name.o /tmp/pc_$name.o /tmp/pt_$name.o props_kernels.o -lcudart_static -lpthread -ldl -lrt
