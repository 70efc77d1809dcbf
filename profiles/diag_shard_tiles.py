"""Heaviest work units of ONE source shard after the adaptive tiling settled (one GPU): elapsed time, points, target range, open points,
sweep end.  usage: python profiles/diag_shard_tiles.py [workload] [lo] [hi] [calls]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
n = len(g["poses"])
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 7 * n // 8
hi = int(sys.argv[3]) if len(sys.argv) > 3 else n
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 8
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
for _ in range(calls):
    r = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
print("range", lo, hi, "ms_search", round(r["ms_search"], 3), "tiles", r["n_tiles"], "packed_ms", round(r["sum_tile_cycles"] / (148 * 64) / 1.965e6, 3),
      "max_tile_ms", round(r["max_tile_cycles"] / 1.965e6, 3))
w = gpu.debug_tile_work().astype(np.float64) / 1.965e6
d = gpu.debug_tile_desc()
mine = np.flatnonzero((d["scan"] >= lo) & (d["scan"] < hi))
order = mine[np.argsort(-w[mine])]
print("units in shard", len(mine), "sum ms", round(w[mine].sum(), 1), "units > 0.3 ms:", int((w[mine] > 0.3).sum()), "> 0.6 ms:", int((w[mine] > 0.6).sum()))
for t in order[:16]:
    print("   unit %6d: %.3f ms  scan %4d  k0 %3d len %2d  targets [%d, %s]  open %d" % (t, w[t], d["scan"][t], d["k0"][t], d["len"][t], d["jlo"][t],
          "end" if d["jhi"][t] == 0xFFFFFFFF else str(d["jhi"][t]), d["open"][t]))
h, e = np.histogram(w[mine], bins=[0, 0.01, 0.03, 0.1, 0.3, 0.6, 1.0, 3.0])
print("histogram of unit ms:", list(zip([round(x, 2) for x in e[1:]], h.tolist())))
gpu.close()
