export BLP_LIB=$PWD/variants/libblp_tuning.so
run() { echo "== $*"; env "$@" timeout 200 python tools/gpu_freeze_ab.py c4 512 0 0.05:0.98 2>&1 | tail -n 1 | cut -c1-400; }
run BLP_POW_PASSES=3
run BLP_POW_PASSES=8 BLP_POW_FIRST=80
run BLP_POW_PASSES=20 BLP_POW_FIRST=200
run BLP_POW_PASSES=8 BLP_POW_FIRST=80 BLP_STEP_SAFETY=0.9
echo "== c5"; timeout 300 python tools/gpu_freeze_ab.py c5 512 0 0.05:0.98 2>&1 | tail -n 1 | cut -c1-400
