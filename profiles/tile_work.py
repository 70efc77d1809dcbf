"""Distribution of per-tile cost of the search kernel (cycles measured in-kernel): how long the heaviest
32-point tile runs bounds the kernel from below, whatever the number of SMs or GPUs."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from hitl_slam_b200 import HitlGpu
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
from hitl_slam_b200 import synth
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
for _ in range(3):
    info = gpu.find_stf(poses, fetch=False)
w = gpu.debug_tile_work().astype(np.float64) / 1.965e6   # ms at 1965 MHz
print("tiles %d  kernel %.2f ms  sum of tile times %.1f ms (= %.1f warp-ms per SM-warp-slot of %d)" % (len(w), info["ms_search"], w.sum(), w.sum() / (148 * 64), 148 * 64))
print("tile ms: mean %.4f  p50 %.4f  p90 %.4f  p99 %.4f  p99.9 %.4f  max %.4f" % (w.mean(), *np.percentile(w, [50, 90, 99, 99.9]), w.max()))
top = np.argsort(-w)[:10]
print("heaviest tiles:", [(int(t), round(float(w[t]), 3)) for t in top])
gpu.close()
