import os, time, sys, numpy as np
sys.path.insert(0, '.')
import bench
from threadpoolctl import threadpool_limits, threadpool_info
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), [(d['user_api'], d['num_threads']) for d in threadpool_info()])
try:
    print("cgroup cpu.max:", open('/sys/fs/cgroup/cpu.max').read().strip())
except Exception as e:
    print("no cgroup v2 cpu.max", e)
frames, truth, init = bench.sample_frames_for_cpu(64, 2000)
for limit in (None, 1):
    for workers in (1, 2, 4, 8, 16):
        n = max(8, workers * 4)
        if limit:
            with threadpool_limits(limit):
                w, _ = bench.cpu_refine(frames[:n], init[:n], workers)
        else:
            w, _ = bench.cpu_refine(frames[:n], init[:n], workers)
        print(f"blas limit {limit} workers {workers:2d}: {n / w[0]:7.1f} poses/s")
