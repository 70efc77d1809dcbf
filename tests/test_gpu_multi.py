"""The multi-GPU exchange inside the C ABI (comm.cu: hitl_comm_*, hitl_normal_eq_allreduce, hitl_gather_stf_blocks).

One rank per GPU.  The 1-rank tests run on any GPU box; the 2-rank test needs two visible GPUs (`gpurun --gpus 2`) and is skipped
otherwise.  The CPU (gloo) coverage of the sharding logic is tests/test_host_mirror_cpu.py::test_two_rank_sharding_over_gloo."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
STD, CORR = 0.05, 1.0 / 40.0


def _load(gpu, g):
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.build_kdtrees()


def _jitter(g, seed):
    return g["poses"].astype(np.float64) + np.random.default_rng(seed).normal(size=g["poses"].shape) * 0.01


def test_single_rank_communicator_is_the_identity(gpu, oracle, maps):
    """world = 1: the collective calls work without peers — hitl_normal_eq_allreduce returns hitl_normal_eq's blocks and
    hitl_gather_stf_blocks the r / J of hitl_eval, in block order."""
    from hitl_slam_b200 import HitlGpu
    g = maps("small")
    _load(gpu, g)
    poses, x = g["poses"].astype(np.float64), _jitter(g, 5)
    gpu.comm_init(HitlGpu.comm_unique_id(), 0, 1)
    try:
        info = gpu.comm_info()
        assert info["world"] == 1 and info["rank"] == 0 and info["nccl_version"] > 20000
        found = gpu.find_stf(poses)
        gpu.set_odometry_blocks(oracle.odometry_consts(g["poses"]))
        gpu.set_stf_blocks_from_search(STD, CORR)
        want = gpu.normal_eq(x)
        got = gpu.normal_eq_allreduce(x)
        for k in ("H_diag", "g"):
            assert np.abs(got[k] - want[k]).max() <= 1e-12 * np.abs(want[k]).max()
        assert abs(got["cost"] - want["cost"]) <= 1e-12 * want["cost"]
        again = gpu.normal_eq_allreduce(None)                       # resident buffer, no re-evaluation
        assert np.array_equal(again["H_diag"], got["H_diag"]) and again["cost"] == got["cost"]
        ev = gpu.eval(x)
        ga = gpu.gather_stf_blocks(0)
        assert int(ga["counts"][0]) == len(found["pair_i"])
        assert np.array_equal(ga["pair_i"], found["pair_i"]) and np.array_equal(ga["pair_j"], found["pair_j"])
        assert np.array_equal(ga["r"], ev["r_stf"]) and np.array_equal(ga["J"].reshape(ev["J_stf"].shape), ev["J_stf"])
        from hitl_slam_b200 import HitlError
        with pytest.raises(HitlError):
            gpu.comm_init(HitlGpu.comm_unique_id(), 0, 1)           # one communicator per context
        with pytest.raises(HitlError):
            gpu.gather_stf_blocks(0, cap_blocks=1)                   # HITL_ERR_OVERFLOW, after the collective completed
    finally:
        gpu.comm_destroy()
    assert gpu.comm_info()["world"] == 1
    gpu.set_stf_blocks_from_search(STD, CORR)
    from hitl_slam_b200 import HitlError
    with pytest.raises(HitlError):
        gpu.gather_stf_blocks(0)                                     # blocks changed since the last hitl_eval


def _rank_main(rank, world, uid, conn, name):
    sys.path.insert(0, ROOT)
    try:
        from hitl_slam_b200 import HitlGpu, synth
        from hitl_slam_b200.sharding import shard_ranges
        g = synth.generate(name)
        gpu = HitlGpu(rank)
        gpu.comm_init(uid, rank, world)
        # sharded uploads: every rank passes the same host arrays, each moves 1/world over PCIe, the slices are all-gathered over NVLink
        gpu.set_scans_sharded(g["offsets"], g["pts"], g["nrm"])
        gpu.build_kdtrees()
        comp, nodes = gpu.get_kdtrees_compact().copy(), gpu.get_kdtrees().copy()
        gpu.set_scans_sharded(g["offsets"], g["pts"], g["nrm"])
        gpu.set_kdtrees_compact_sharded(comp)
        assert np.array_equal(gpu.get_kdtrees(), nodes), "sharded compact tree upload"
        gpu.set_kdtrees_sharded(nodes)
        assert np.array_equal(gpu.get_kdtrees(), nodes), "sharded tree upload"
        poses, x = g["poses"].astype(np.float64), _jitter(g, 5)
        lo, hi = shard_ranges(g["offsets"], world)[rank]
        part = gpu.find_stf(poses, src_lo=lo, src_hi=hi)
        if rank == 0:                                                # odometry blocks live on one rank only
            from hitl_slam_b200 import HostLib
            gpu.set_odometry_blocks(HostLib().odometry_consts(g["poses"]))
        gpu.set_stf_blocks_from_search(STD, CORR)
        ne = gpu.normal_eq_allreduce(x)
        gpu.eval(x)
        ga = gpu.gather_stf_blocks(0, cap_blocks=400000)
        out = {"H_diag": ne["H_diag"], "g": ne["g"], "cost": ne["cost"], "n_local": len(part["pair_i"]), "counts": ga["counts"]}
        if rank == 0:
            out.update(pair_i=ga["pair_i"], pair_j=ga["pair_j"], r=ga["r"], J=ga["J"])
        gpu.comm_destroy()
        gpu.close()
        conn.send(out)
    except Exception as e:                                           # the parent must not wait for ever
        conn.send({"error": repr(e)})


def test_two_rank_exchange(gpu, host, maps):
    """Two ranks, one per GPU, through the C ABI only: sharded searches, in-library ncclAllReduce of the packed normal equations
    (every rank ends with the WHOLE problem's blocks), NCCL gather of 14 doubles per STF block to rank 0 in the reference's block
    order — compared with the single-GPU evaluation of the whole problem."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from hitl_slam_b200 import HitlGpu
    name = "small"
    g = maps(name)
    _load(gpu, g)
    poses, x = g["poses"].astype(np.float64), _jitter(g, 5)
    found = gpu.find_stf(poses)
    gpu.set_odometry_blocks(host.odometry_consts(g["poses"]))
    gpu.set_stf_blocks_from_search(STD, CORR)
    want, ev = gpu.normal_eq(x), gpu.eval(x)
    uid = HitlGpu.comm_unique_id()
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(2):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_rank_main, args=(r, 2, uid, b, name))
        p.start()
        pipes.append(a); procs.append(p)
    outs = []
    for a, p in zip(pipes, procs):
        assert a.poll(300), "rank did not answer"
        outs.append(a.recv())
        p.join(60)
    for o in outs:
        assert "error" not in o, o
    for o in outs:                                                   # every rank holds the whole problem's H_diag / g / cost
        for k in ("H_diag", "g"):
            assert np.abs(o[k] - want[k]).max() <= 1e-9 * np.abs(want[k]).max()
        assert abs(o["cost"] - want["cost"]) <= 1e-9 * want["cost"]
        assert [int(c) for c in o["counts"]] == [outs[0]["n_local"], outs[1]["n_local"]]
    assert outs[0]["n_local"] + outs[1]["n_local"] == len(found["pair_i"]) and min(outs[0]["n_local"], outs[1]["n_local"]) > 0
    assert np.array_equal(outs[0]["pair_i"], found["pair_i"]) and np.array_equal(outs[0]["pair_j"], found["pair_j"])
    assert np.array_equal(outs[0]["r"], ev["r_stf"]) and np.array_equal(outs[0]["J"].reshape(ev["J_stf"].shape), ev["J_stf"])


def test_sharded_jointopt_in_one_process(gpu, host, maps):
    """The C++ drop-in on two GPUs, one process: hitl::JointOpt with a second context (JointOpt::UseShardContexts).  Scans + trees are
    replicated, each context searches its own source range, the concatenated list and every GPU-backed cost block (evaluated through
    CostFunction::Evaluate, blocks of both shards) equal the single-context stage's; the post-HitL solve ends at the same poses."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from hitl_slam_b200 import HitlGpu, HostSession
    g = maps("small")
    x = _jitter(g, 11)
    one = HostSession(gpu, host)
    one.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    want = one.find_stf()
    nb = want["n_pairs"]
    n_odo = len(g["poses"]) - 1
    picks = [n_odo + b for b in (0, 1, nb // 2 - 1, nb // 2, nb // 2 + 1, nb - 2, nb - 1)]
    want_blocks = [one.evaluate_block(b, with_stf=True, pose_array=x) for b in picks]
    one.solver_options(1, max_iterations=8)
    one.set_poses(g["poses"])
    one.find_stf()
    s1 = one.solve(1)
    p1 = one.poses()[1].copy()
    one.close()
    second = HitlGpu(1)
    two = HostSession(gpu, host)
    two.use_shard_contexts([second])
    two.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    got = two.find_stf()
    info = two.shard_info()
    assert len(info) == 2 and info[0][0] == 0 and info[0][1] == info[1][0] and info[1][1] == len(g["poses"]) and info[0][2] + info[1][2] == nb and min(info[0][2], info[1][2]) > 0
    for k in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(got[k], want[k]), k
    assert got["n_queries"] == want["n_queries"]
    for b, (r0, j0, j1, tot) in zip(picks, want_blocks):
        r, a, c, tot2 = two.evaluate_block(b, with_stf=True, pose_array=x)
        assert tot2 == tot and np.array_equal(r, r0) and np.array_equal(a, j0) and np.array_equal(c, j1), b
    two.solver_options(1, max_iterations=8)
    two.set_poses(g["poses"])
    two.find_stf()
    s2 = two.solve(1)
    assert s2["termination"] == s1["termination"] and abs(s2["final_cost"] - s1["final_cost"]) <= 1e-9 * max(s1["final_cost"], 1e-30)
    assert np.abs(two.poses()[1] - p1).max() <= 1e-9
    two.close()
    second.close()
