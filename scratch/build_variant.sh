#!/bin/bash
# Developer experiment helper: build scratch/exp/lib_<name>.so with extra nvcc flags on rg_edge.cu only
# (the other objects come from the normal build).  Run a variant with REDGNN_B200_LIB=<path>.
set -e
name=$1; shift
cd "$(dirname "$0")/../redgnn_b200/csrc"
make -s
mkdir -p ../../scratch/exp
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    -I../../include --expt-relaxed-constexpr "$@" -c rg_edge.cu -o /tmp/rg_edge_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/exp/lib_$name.so \
    rg_abi.o rg_expand.o /tmp/rg_edge_$name.o rg_node.o rg_node_tc.o rg_node_bwd.o rg_attn.o rg_graph.o
echo built scratch/exp/lib_$name.so
