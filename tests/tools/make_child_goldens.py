"""Golden answers for frontier children of the large bench workloads (C4, C5), made OFFLINE with the
HiGHS oracle: node ids GOLD_FIRST .. GOLD_FIRST+63 of ``instances.frontier_nodes`` (seed 0), each
solved warm from the root basis as the reference hands the parent's basis to a child
(base_node.py:589, 608). Stored under bench_data/<workload>_children.npz.

Who reads them: tests/test_gpu_full_size.py (objective within 1e-6, equal status, through the C ABI)
and bench.py, which puts 8 of these nodes into every timed step and validates them AFTER the timed
region ('validated': n/n in its JSON line).

    python tests/tools/make_child_goldens.py c5 [cores]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from simple_mip_solver_b200.instances import GOLD_COUNT, GOLD_FIRST, frontier_nodes  # noqa: E402


def main():
    name = sys.argv[1]
    cores = int(sys.argv[2]) if len(sys.argv) > 2 else bench.host_cores()
    d, depth, root = bench.load_instance(name)
    _, _, deltas = frontier_nodes(d, root['x'], GOLD_FIRST, GOLD_COUNT, depth, seed=0)
    arm = bench.CpuArm(d, root, cores)
    t = time.time()
    out, dt = arm.run(deltas)
    arm.close()
    status = np.array([o[0] for o in out], dtype=np.int32)
    obj = np.array([o[1] for o in out])
    secs = np.array([o[2] for o in out])
    print(name, 'children', len(out), 'status counts', np.bincount(status, minlength=4).tolist(),
          f'wall {time.time() - t:.1f} s, mean {secs.mean():.2f} s per LP per core')
    np.savez_compressed(os.path.join(ROOT, 'bench_data', f'{name}_children.npz'), node_ids=np.arange(GOLD_FIRST, GOLD_FIRST + GOLD_COUNT),
                        status=status, objective=obj, seconds=secs, seed=0, depth=depth,
                        oracle='HiGHS 1.12 dual simplex, presolve off, tolerances 1e-9, warm start from the root basis')


if __name__ == '__main__':
    mp.set_start_method('fork')
    main()
