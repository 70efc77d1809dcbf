"""Shard view of the search kernel on ONE GPU: what each of 8 ranks runs (1/8 of the source poses against all targets),
critical path (heaviest tile) vs packed work, and the descriptors of the heaviest tiles after the adaptive re-tiling settled.
The split policy is read from the environment (HITL_SPLIT_LIMIT_DIV / HITL_MIN_TARGET_SPAN / HITL_SPLIT_ROUNDS).
usage: python profiles/diag_shard.py [workload] [calls]"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
n = len(poses)
print("policy", {k: os.environ.get(k) for k in ("HITL_SPLIT_LIMIT_DIV", "HITL_MIN_TARGET_SPAN", "HITL_SPLIT_ROUNDS")})
slots = 148 * 64
for lo, hi in ((0, n // 8), (3 * n // 8, n // 2), (7 * n // 8, n), (0, n)):
    hist = []
    for _ in range(calls):
        r = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
        hist.append(round(r["ms_search"], 3))
    print("range", lo, hi, "ms_search per call", hist, "ms_total", round(r["ms_total"], 3), "tiles", r["n_tiles"],
          "packed_ms", round(r["sum_tile_cycles"] / slots / 1.965e6, 3), "max_tile_ms", round(r["max_tile_cycles"] / 1.965e6, 3))
    w = gpu.debug_tile_work().astype(np.float64) / 1.965e6
    d = gpu.debug_tile_desc()
    top = np.argsort(-w)[:6]
    for t in top:
        print("   tile %d: %.3f ms  scan %d  k0 %d len %d  targets [%d, %s]  open %d" % (t, w[t], d["scan"][t], d["k0"][t], d["len"][t], d["jlo"][t],
              "end" if d["jhi"][t] == 0xFFFFFFFF else str(d["jhi"][t]), d["open"][t]))
gpu.close()
