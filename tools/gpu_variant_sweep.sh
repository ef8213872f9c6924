#!/bin/bash
# Chunk sizes and occupancy variants of the wide step kernels with frozen coordinates (C5, 1024 nodes through 512 slots).
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/gpu_freeze_ab.py c5 1024 512 0.05:0.98 2>&1 | tail -n 1 | python -c "
import sys, json; j = json.loads(sys.stdin.read()); print({k: j[k] for k in ('mean_it', 'total_ms', 'step_ms', 'us_per_node_iter', 'skipped_cols', 'step_resets')})"; }
T=$PWD/variants/libblp_tuning.so
run BLP_LIB=$T
run BLP_LIB=$T BLP_ROWS_PER_WARP2P=32
run BLP_LIB=$T BLP_ROWS_PER_WARP2P=48
run BLP_LIB=$T BLP_ROWS_PER_WARP2P=32 BLP_ROWS_PER_WARP2=24
run BLP_LIB=$T BLP_ROWS_PER_WARP2P=8 BLP_ROWS_PER_WARP2=8
run BLP_LIB=$T BLP_GRAPH_LANES=4
run BLP_LIB=$PWD/variants/libblp_minb6.so
run BLP_LIB=$PWD/variants/libblp_minb4.so
run BLP_LIB=$PWD/variants/libblp_u8.so
run BLP_LIB=$T BLP_FREEZE_MARGIN=0.02
