#!/bin/bash
# Diagnostics build of the CUDA library with launch knobs (FPC_EXPERIMENT), into tools/libfpc_x.so.
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -cudart shared -DFPC_EXPERIMENT -o tools/libfpc_x.so alphazero_4_player_chess_b200/csrc/fpc_kernels.cu alphazero_4_player_chess_b200/csrc/fpc_puct.cu
