import sys, numpy as np, json
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu
g = bench.workload("c2", 5000, 720)
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
for fine in (True, False):
    gpu.debug_set_fine_occupancy(fine)
    for _ in range(3):
        r = gpu.find_stf(poses, fetch=False)
    print(fine, {k: (int(v) if not isinstance(v, float) else round(v, 3)) for k, v in r.items() if k.startswith("n_") or k.startswith("ms_")})
