#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...]: an experimental build of the library with extra compiler flags, as build/libdkgv_NAME.so
# (select it with DKGV_LIB=build/libdkgv_NAME.so).  Units that do not see the flags' macros are reused from build/.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/var_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for u in dkgv share_fd final pairing; do
  nvcc $FLAGS "$@" -c -o build/var_$name/$u.o dvt_circuits_b200/csrc/$u.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libdkgv_$name.so build/var_$name/dkgv.o build/var_$name/share_fd.o build/var_$name/final.o \
  build/var_$name/pairing.o build/flows.o build/comm.o build/dkg_host.o -ldl
echo built build/libdkgv_$name.so
