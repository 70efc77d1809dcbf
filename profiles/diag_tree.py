"""Tree build timing: device builder vs host builder (threads), c2 and c3."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from hitl_slam_b200 import HitlGpu, synth
gpu = HitlGpu(0)
for name in sys.argv[1:] or ["c2"]:
    g = bench.workload(name, synth.CONFIGS[name]["n_poses"], synth.CONFIGS[name]["beams"])
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    for host in (False, True, False):
        gpu.debug_set_tree_builder(host)
        t0 = time.perf_counter(); gpu.build_kdtrees(); t1 = time.perf_counter()
        nodes = gpu.get_kdtrees()
        print(name, "host" if host else "device", "build %.1f ms" % ((t1 - t0) * 1e3), "exact segments", gpu.debug_tree_stats(), "checksum", hex(int(np.bitwise_xor.reduce(nodes.view(np.uint32).ravel().astype(np.uint64) * np.arange(1, nodes.view(np.uint32).size + 1, dtype=np.uint64) % np.uint64(0xFFFFFFFB)))))
