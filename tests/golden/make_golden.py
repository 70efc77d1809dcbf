"""Regenerates the committed golden fixtures under tests/golden/.

The reference ships no golden vectors for this path, so the fixtures are outputs of the
CPU oracle (oracle/hitl_oracle.hpp) on small seeded synthetic maps; the reference's own JointOpt /
EMInput, compiled where they lie (oracle/_ref/libhitl_ref.so), reproduce them from the fixture inputs
(tests/test_oracle_ref_backend.py::test_golden_fixtures_equal_the_references_output).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hitl_slam_b200 import build, synth  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

build.build_all()
HERE = os.path.dirname(os.path.abspath(__file__))
o = Oracle()

g = synth.generate("tiny", normals="compensated")
np.savez_compressed(os.path.join(HERE, "tiny_compensated.npz"), poses=g["poses"], offsets=g["offsets"], pts=g["pts"], nrm=g["nrm"])
S = o.scans(g["offsets"], g["pts"], g["nrm"])
r = S.find_stf(g["poses"].astype(np.float64))
np.savez_compressed(os.path.join(HERE, "tiny_stf.npz"), **{k: r[k] for k in ("pair_i", "pair_j", "pair_off", "k", "idx")}, n_queries=r["n_queries"])
rng = np.random.default_rng(0)
x = g["poses"].astype(np.float64) + rng.normal(size=g["poses"].shape) * 0.01
res, J = S.eval_stf(x, r)
consts = o.odometry_consts(g["poses"])
ro, Jo = o.eval_odometry(consts, x)
np.savez_compressed(os.path.join(HERE, "tiny_eval.npz"), x=x, r_stf=res, J_stf=J, odo_consts=consts, r_odo=ro, J_odo=Jo)

g = synth.generate("small")
S = o.scans(g["offsets"], g["pts"], g["nrm"], build_trees=False)
world = S.world_transform(g["poses"])
strokes = synth.make_strokes(g)
op, oi = o.em_inliers(g["offsets"], world, strokes[:2].reshape(-1))
sets = o.em_assign(g["offsets"], world, strokes)
run = o.em_run(g["offsets"], world, strokes)
np.savez_compressed(os.path.join(HERE, "small_em.npz"), strokes=strokes, inl_pose=op, inl_idx=oi,
                    set0_pose=sets[0][0], set0_off=sets[0][1], set0_obs=sets[0][2], set1_pose=sets[1][0], set1_off=sets[1][1], set1_obs=sets[1][2],
                    run_segs=run["segs"], run_corrected=run["corrected"], run_anchor=run["anchor"], run_backprop=np.array(run["backprop"]))
print("tiny:", len(r["pair_i"]), "pairs", len(r["k"]), "matches", r["n_queries"], "queries")
print("small EM:", len(op), "inliers; sets", len(sets[0][0]), len(sets[1][0]), "run", run)
