#!/bin/bash
# registers / spills of the full-operator atomic instantiation of k_apply2d_thread per order, with extra -D flags
cd "$(dirname "$0")/../continuum-mechanics-mfem_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. -Xptxas -v "$@" -c kernels_apply_2d.cu -o /tmp/k2d.o 2>&1 | grep -E "error|Compiling entry|Used|spill" | paste - - - | grep "error\|ELb1ELb1ELb1ELb1E" | sed -E 's/.*threadILi([0-9])ELi([0-9])ELi([0-9])ELi([0-9]).*sm_100a.\s*(.*)ptxas info.*Used ([0-9]+) registers.*/P=\1 NW=\2 NBUF=\3 MINB=\4 regs=\6 \5/'
