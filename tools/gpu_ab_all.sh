#!/bin/bash
# A/B of the variant builds under variants/ (made with -D switches of blp_kernels.cuh): one C5 slice each.
mkdir -p gpurun_out
for lib in variants/libblp_*.so; do
  BLP_LIB=$PWD/$lib timeout 300 python tools/gpu_ab.py ${1:-c5} ${2:-512} 2>&1 | tail -1
done
python - <<'PY'
import glob, numpy as np
fs = sorted(glob.glob('gpurun_out/ab_obj_libblp_*.npy'))
ref = np.load(fs[0])
for f in fs[1:]:
    o = np.load(f)
    print(f, 'max |obj diff| vs', fs[0], float(np.max(np.abs(o - ref))))
PY
