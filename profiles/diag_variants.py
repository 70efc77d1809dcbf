"""Shard view (1/8 of the source poses, what each of 8 ranks runs) under the three residency variants of the search kernel
(hitl_debug_set_search_variant: 0 = 16 CTAs/SM = 64 warps, 1 = 12 CTAs/SM, 2 = 10 CTAs/SM): a shard is bounded by its heaviest tile,
and a tile's latency grows with the number of warps sharing the SM's issue slots.
usage: python profiles/diag_variants.py [workload]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
n = len(poses)
for lo, hi in ((0, n // 8), (3 * n // 8, n // 2), (7 * n // 8, n), (0, n // 2)):
    gpu.debug_set_search_variant(0, -1)
    for _ in range(6):
        gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)           # lets the adaptive re-tiling of this range settle
    for variant in (0, 1, 2, 0):
        gpu.debug_set_search_variant(variant, -1)
        ms = []
        for _ in range(4):
            r = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
            ms.append(round(r["ms_search"], 3))
        print("range", lo, hi, "variant", variant, "ms_search", ms, "max_tile_ms", round(r["max_tile_cycles"] / 1.965e6, 3),
              "sum_tile_ms/slots(64w)", round(r["sum_tile_cycles"] / (148 * 64) / 1.965e6, 3), "tiles", r["n_tiles"])
gpu.close()
