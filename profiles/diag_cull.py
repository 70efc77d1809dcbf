"""Cull / walk diagnostics of the search kernel on one workload (default c2): counters and kernel time per setting.
usage: python profiles/diag_cull.py [workload] [poses] [beams]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else synth.CONFIGS[name]["n_poses"]
beams = int(sys.argv[3]) if len(sys.argv) > 3 else synth.CONFIGS[name]["beams"]
g = bench.workload(name, n_poses, beams)
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
for variant, carve, fine in ((0, -1, True), (0, 50, True), (0, 30, True), (1, -1, True), (0, -1, False)):
    gpu.debug_set_search_variant(variant, carve)
    gpu.debug_set_fine_occupancy(fine)
    for _ in range(4):
        r = gpu.find_stf(poses, fetch=False)
    print("variant", variant, "carveout", carve, "fine", fine, {k: (int(v) if not isinstance(v, float) else round(v, 3)) for k, v in r.items() if k.startswith("n_") or k.startswith("ms_")})

# shard view: one quarter of the source poses (what one of 4 ranks runs), critical path vs packed work
gpu.debug_set_search_variant(0, -1); gpu.debug_set_fine_occupancy(True)
n = len(poses)
for lo, hi in ((0, n), (0, n // 4), (3 * n // 4, n), (0, n // 8), (7 * n // 8, n)):
    for _ in range(5):
        r = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
    slots = 148 * 64
    print("range", lo, hi, "ms_search", round(r["ms_search"], 3), "ms_total", round(r["ms_total"], 3), "tiles", r["n_tiles"],
          "packed_ms@1.965GHz", round(r["sum_tile_cycles"] / slots / 1.965e6, 3), "max_tile_ms", round(r["max_tile_cycles"] / 1.965e6, 3))
