#!/bin/bash
# The host-only text reader (csrc/rg_text.cpp) under AddressSanitizer + UBSan: links an instrumented
# rg_text.o with the regular device objects into a scratch library and runs tests/test_text_cpu.py on it.
# (compute-sanitizer is closed on the GPU pool; this covers the one translation unit that runs on the host.)
set -e
cd "$(dirname "$0")/../redgnn_b200/csrc"
make -s
g++ -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -O1 -g -fPIC -pthread -I../../include \
    -c rg_text.cpp -o /tmp/rg_text_asan.o
g++ -shared -o /tmp/libredgnn_asan.so rg_abi.o rg_expand.o rg_edge.o rg_node.o rg_node_tc.o rg_node_bwd.o rg_attn.o \
    rg_graph.o /tmp/rg_text_asan.o -L/usr/local/cuda/lib64 -lcudart_static -lrt -ldl -pthread -fsanitize=address,undefined
cd ../..
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
REDGNN_B200_LIB=/tmp/libredgnn_asan.so python -m pytest tests/test_text_cpu.py -q -x "$@"
