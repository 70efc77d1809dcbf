import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from hitl_slam_b200 import capi
g = bench.workload("c2", 5000, 720)
for mb in (2, 3, 4):
    capi.lib_path = lambda name="libhitl_gpu.so", mb=mb: os.path.join(capi.LIB_DIR, "libhitl_gpu_mb%d.so" % mb if name == "libhitl_gpu.so" else name)
    gpu = capi.HitlGpu(0)
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
    poses = g["poses"].astype(np.float64)
    gpu.find_stf(poses, fetch=False)
    gpu.set_odometry_blocks(bench.odometry_consts_host(g["poses"])); gpu.set_stf_blocks_from_search(0.05, 0.025)
    for prec in (0, 1):
        ev = [gpu.eval(poses, precision=prec, fetch=False)["ms"] for _ in range(8)]
    ne = [gpu.normal_eq(poses, fetch=False)["ms"] for _ in range(8)]
    ev0 = [gpu.eval(poses, precision=0, fetch=False)["ms"] for _ in range(8)]
    ev1 = [gpu.eval(poses, precision=1, fetch=False)["ms"] for _ in range(8)]
    print("minblocks", mb, "eval f64 %.3f ms  f32 %.3f ms  normal_eq %.3f ms" % (min(ev0), min(ev1), min(ne)))
    gpu.close()
