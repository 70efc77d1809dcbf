"""Checks that the adaptive re-tiling is a setup cost only: repeated searches over one shard range; host time per call, device time,
tile counts before / after each call.  usage: python profiles/diag_retile.py [workload] [lo] [hi] [calls]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 4613
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 12
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = gpu.pinned_copy(g["poses"].astype(np.float64))
for c in range(calls):
    t0 = time.perf_counter()
    r = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
    ms = (time.perf_counter() - t0) * 1e3
    print("call %2d: host %.3f ms  device %.3f ms (search %.3f)  tiles %d -> %d  max_tile_ms %.3f" % (c, ms, r["ms_total"], r["ms_search"], r["n_tiles"], r["n_tiles_next"], r["max_tile_cycles"] / 1.965e6))
gpu.close()
