"""Where the tree walks of one search go (default c2): counters of the last warmed call, and the share of walks / in-radius / gate failures /
speculative drops, plus the distribution of open (below-cap) points per tile.  usage: python profiles/diag_walks.py [workload]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
gpu = HitlGpu(0)
gpu.set_scans(g["offsets"], g["pts"], g["nrm"]); gpu.build_kdtrees()
poses = g["poses"].astype(np.float64)
for _ in range(5):
    r = gpu.find_stf(poses, fetch=False)
print({k: (int(v) if not isinstance(v, float) else round(v, 3)) for k, v in r.items()})
m = int(g["offsets"][-1])
print("per point: walks %.2f in_radius %.2f gate_fail %.2f over_cap %.2f raw matches %.2f coarse passes %.2f" % (
    r["n_traversals"] / m, r["n_in_radius"] / m, r["n_gate_fail"] / m, r["n_over_cap"] / m, r["n_raw_matches"] / m, r["n_coarse_pass"] / m), "dir_culled per point %.2f" % (r["n_dir_culled"] / m))
opn = gpu.debug_tile_desc()["open"]
work = gpu.debug_tile_work()
opn = np.asarray(opn); work = np.asarray(work, np.float64)
print("tiles", len(opn), "tiles with open points", int((opn > 0).sum()), "open points total", int(opn.sum()))
order = np.argsort(-work)
cum = np.cumsum(work[order]) / work.sum()
for frac in (0.01, 0.05, 0.1, 0.25, 0.5):
    k = int(frac * len(work))
    print("top %4.0f%% tiles by work: %.1f%% of cycles, mean open points %.2f" % (100 * frac, 100 * cum[k - 1], opn[order[:k]].mean()))
print("work share of tiles with open points: %.1f%%" % (100 * work[opn > 0].sum() / work.sum()))
