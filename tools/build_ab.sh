#!/bin/bash
# usage: tools/build_ab.sh NAME [extra nvcc flags...]  -> ab/lib_NAME.so (solver.cu rebuilt with the flags, other objects reused)
set -e
cd "$(dirname "$0")/../landing_controller_b200/csrc"
name=$1; shift
make -s all >/dev/null
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v -c solver.cu -o /tmp/ab/solver_$name.o 2> /tmp/ab/solver_$name.log
grep -A2 "k_solve" /tmp/ab/solver_$name.log | grep -E "spill|Used" 
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/lib_$name.so build/eval.o /tmp/ab/solver_$name.o build/tvlqr.o build/kino.o build/capi.o build/casadi_abi.o -lcudart_static -lpthread -ldl -lrt
