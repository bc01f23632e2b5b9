#!/bin/bash
# Builds the kernel-logic emulator (CPU, tests only): the product sources compiled with
# g++ -DLESGO_EMUL.  -ffp-contract=off keeps g++ from fusing multiply-adds so the
# un-contracted expressions behave like the __dmul_rn/__dadd_rn device versions.
set -e
cd "$(dirname "$0")"
mkdir -p _build
SRC=../../lesgo_b200/csrc
FLAGS="-O2 -std=c++17 -fPIC -DLESGO_EMUL ${LESGO_EMUL_DEFS} -ffp-contract=off -I. -Wno-unused-function"
pids=()
for f in lesgo_gpu xfwd_scale xfwd_vort xfwd_convec xinv ypass prodfwd fftw_shim comm extras; do
  [ -f $SRC/$f.cu ] || continue
  if [ ! -f _build/$f.o ] || [ -n "$(find $SRC ../emul -newer _build/$f.o \( -name '*.h' -o -name "$f.cu" \) | head -1)" ]; then
    g++ $FLAGS -x c++ -c $SRC/$f.cu -o _build/$f.o &
    pids+=($!)
  fi
done
g++ $FLAGS -c emul_rt.cpp -o _build/emul_rt.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -shared -Wl,-Bsymbolic -o _build/liblesgo_emul.so _build/*.o -lpthread
echo built _build/liblesgo_emul.so
