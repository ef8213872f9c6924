#!/bin/bash
# Sweep of the chunk size of the step kernels with frozen coordinates (tuning build: variants/libblp_tuning.so).
mkdir -p gpurun_out
export BLP_LIB=$PWD/variants/libblp_tuning.so
run() { echo "== $*"; env "$@" timeout 200 python tools/gpu_freeze_ab.py c5 512 0 0.05 2>&1 | tail -n 1; }
timeout 300 python tools/gpu_freeze_ab.py c5 512 0 0.0 0.05 0.02 0.1 2>&1 | tail -n 4
run BLP_ROWS_PER_WARP2P=24
run BLP_ROWS_PER_WARP2P=32
run BLP_ROWS_PER_WARP2P=48
run BLP_ROWS_PER_WARP2P=32 BLP_ROWS_PER_WARP2=24
run BLP_ROWS_PER_WARP2P=32 BLP_FREEZE_RELEASE=0.1
run BLP_ROWS_PER_WARP2P=32 BLP_GRAPH_LANES=1
run BLP_ROWS_PER_WARP2P=32 BLP_GRAPH_LANES=4
