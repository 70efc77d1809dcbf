#!/usr/bin/env python
"""bench.py — headline measurement of the hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

One "step" = one pass of the hot path over the synthetic map: FindSTFCorrespondences over all
ordered pose pairs + one residual/Jacobian/normal-equation evaluation of the resulting blocks
(+ odometry blocks).  Metric: M correspondence+Jacobian evals/s = (KD queries the reference
semantics execute + matched correspondences pushed through residual+Jacobian) / time.

 * value      whole-job throughput, scans/trees resident in HBM, CUDA events on the library's stream
 * e2e        same metric through the C ABI with HOST buffers: scans + trees + poses H2D, CSR +
              residuals + Jacobians D2H inside the timed region
 * roofline   stf_search_kernel: algorithmic bytes (40 B/query + 8 B/match) / kernel time
 * cpu_baseline  the oracle port (OpenMP over source poses, reference flags) on a bounded sample
 * --impl reference   the same oracle on the host cores as the reference arm
Multi-GPU: source poses are sharded by point count (no data-path collective in the search);
the per-iteration exchange is one NCCL all-reduce of the packed normal-equation blocks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "correspondence+Jacobian evals/s"
UNIT = "M evals/s"
STD_DEV, CORR = 0.05, 1.0 / 40.0


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries print there too (NCCL's version banner comes from whichever communicator
    a process creates first), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of this rank's GPU during the timed region.  In-process NVML (the library nvidia-smi reads; one
    cheap query per sample, no child process per rank — eight nvidia-smi start-ups inside an 8-rank timed region perturb the launches
    they are meant to observe); falls back to `nvidia-smi --query-gpu ... -lms` when the NVML binding is unavailable.  Started before
    the warm-up; only samples taken between mark_begin() and stop() are reported."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_s=0.010):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self.period = index, [], None, period_s
        self.t_begin, self.halt, self.source = None, threading.Event(), None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            self.source = "nvml"
            while not self.halt.is_set():
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.perf_counter(), sm, mx, [k for k, b in bits.items() if r & b]))
                self.halt.wait(self.period)
            return
        except Exception:
            pass
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                r = [x.strip() for x in line.split(",")]
                if len(r) >= 8 and r[1].replace(".", "").isdigit() and r[2].replace(".", "").isdigit():
                    self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                      [k for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]) if v.lower().startswith("active")]))
        except Exception:
            pass

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        self.halt.set()
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.rows if self.t_begin is None or self.t_begin <= r[0] <= t_end] or self.rows[-1:]
        sm = [r[1] for r in rows]
        mx = [r[2] for r in rows]
        reasons = set(k for r in rows for k in r[3])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def workload(name, n_poses=None, beams=None):
    from hitl_slam_b200 import synth
    cache = os.path.join(os.environ.get("HITL_SYNTH_DIR", "/tmp/hitl_synth"), "%s_%s_%s.npz" % (name, n_poses, beams))
    if os.path.exists(cache):
        z = np.load(cache)
        out = {k: z[k] for k in z.files if k != "config_json"}
        if "config_json" in z.files:
            out["config"] = json.loads(str(z["config_json"]))
            return out
    g = synth.generate(name, n_poses=n_poses, beams=beams)
    out = {k: g[k] for k in ("poses", "offsets", "pts", "nrm", "cov")}
    os.makedirs(os.path.dirname(cache), exist_ok=True)
    np.savez(cache, config_json=np.array(json.dumps(g["config"])), **out)
    out["config"] = g["config"]
    return out


def cpu_sample(g, seconds=15.0, threads=None):
    """Oracle port timed on the host cores on a bounded sample of the same workload: chunks of consecutive
    source poses, visited in a strided order so that a partial sample is spread over the whole trajectory,
    each searched against ALL target poses (OpenMP over the chunk's source poses, as JointOptimization.cpp:575)
    and followed by the evaluation of the blocks it produced; stops at the time budget."""
    from oracle.pyoracle import Oracle
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    orc = Oracle(fast=True)
    S = orc.scans(g["offsets"], g["pts"], g["nrm"])
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    cores = orc.num_threads()
    chunk = max(cores, 8)
    n_chunks = (n + chunk - 1) // chunk
    stride = next(s for s in (61, 37, 17, 7, 3, 1) if n_chunks % s != 0 or s == 1)   # coprime stride -> a permutation of the chunks
    t_search = t_eval = 0.0
    queries = matches = n_src = done = 0
    while done < n_chunks and t_search + t_eval < seconds:
        c = (done * stride) % n_chunks
        lo, hi = c * chunk, min((c + 1) * chunk, n)
        t0 = time.perf_counter()
        r = S.find_stf(poses, src_lo=lo, src_hi=hi)
        t1 = time.perf_counter()
        S.eval_stf(poses, r, STD_DEV, CORR, want_jac=True, parallel=True)
        t2 = time.perf_counter()
        t_search += t1 - t0
        t_eval += t2 - t1
        queries += r["n_queries"]
        matches += len(r["k"])
        n_src += hi - lo
        done += 1
    t_used = t_search + t_eval
    return {"value": (queries + matches) / t_used / 1e6 if t_used else 0.0, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d source poses (%d-pose chunks spread over the trajectory) vs all %d targets: %d queries + %d Jacobian evals in %.1f s "
                      "(search %.1f s, eval %.1f s); oracle -O3 -march=native -fopenmp" % (n_src, n, chunk, n, queries, matches, t_used, t_search, t_eval),
            "search_Mq_per_s": queries / t_search / 1e6 if t_search else 0.0, "eval_Mm_per_s": matches / t_eval / 1e6 if t_eval else 0.0, "seconds": t_used}


REF_SAMPLE_UNIT = 1.1e9      # poses x points per unit of source stride: c2 (5 000 x 3.58 M) -> every 16th source pose, ~2 s of reference search on 16 threads


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU legs (reference arm, cpu_baseline) run on every host
    thread this process may use, and say how many.  Must run before the first OpenMP library is loaded."""
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    return host_threads()


class RefSampler:
    """The REFERENCE's own CPU path on the FULL map (oracle/_ref/libhitl_ref_fast.so = JointOptimization.cpp + kdtree.cpp compiled where
    they lie with the reference's Release flags; prebuilt, it travels to the GPU box).  Set-up, untimed: JointOpt over all poses,
    JointOpt::BuildKDTrees.  One step, timed: JointOpt::FindSTFCorrespondences over the whole pose range — OpenMP over source poses as the
    reference runs it (JointOptimization.cpp:575) — for a FIXED sample of source poses (every s-th pose, s from the map size alone, so
    the sample is the same in every run and at every N) against ALL target poses, then the evaluation of the blocks AddSTFConstraints
    builds through AutoDiffCostFunction on ONE thread (the reference leaves Ceres at one thread, :154-155).  The unmodified loop is
    restricted to the sample through its own data (RefJointOpt.restrict_sources; every source pose is independent of the others).
    Executed-query counts (cap-skipped queries excluded, as the metric defines them) come from the oracle port on the same sample
    (the reference does not count them; both produce identical lists, tests/test_oracle_ref_backend.py)."""

    def __init__(self, g):
        from oracle.pyoracle import Oracle, RefBackend
        self.ref = RefBackend(fast=True)
        orc = Oracle(fast=True)
        self.cores = orc.num_threads()
        n = len(g["poses"])
        self.n, self.n_points = n, int(g["offsets"][-1])
        self.stride = max(1, int(round(n * float(self.n_points) / REF_SAMPLE_UNIT)))
        self.phase = self.stride // 2
        self.ids = np.arange(self.phase, n, self.stride)
        self.poses = g["poses"].astype(np.float64)
        self.J = self.ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
        self.J.restrict_sources(self.ids)
        port = orc.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(self.poses, src_lo=self.phase, src_stride=self.stride)
        self.queries, self.port_matches = int(port["n_queries"]), int(len(port["k"]))

    def step(self):
        t0 = time.perf_counter()
        corr = self.J.find_stf(self.poses)
        t1 = time.perf_counter()
        self.J.eval_blocks(2, self.poses, max(len(corr["pair_i"]), 1))
        t2 = time.perf_counter()
        matches, t_search, t_eval = int(len(corr["k"])), t1 - t0, t2 - t1
        t_used = t_search + t_eval
        return {"value": (self.queries + matches) / t_used / 1e6, "unit": UNIT, "cores": self.cores, "kind": "reference",
                "sample": "every %d-th source pose of the full map (%d of %d poses, fixed) against all %d target poses: JointOpt::FindSTFCorrespondences "
                          "(%d executed queries, %d matches, %d blocks) in %.2f s on %d OpenMP threads + AutoDiffCostFunction evaluation of the blocks in %.2f s on 1 thread; "
                          "reference sources compiled with -O3 -march=x86-64-v3 -fopenmp -DNDEBUG (FMA contraction on, as -march=native gives the reference); "
                          "the port's timing build finds %d matches on the same sample"
                          % (self.stride, len(self.ids), self.n, self.n, self.queries, matches, len(corr["pair_i"]), t_search, self.cores, t_eval, self.port_matches),
                "search_Mq_per_s": self.queries / t_search / 1e6 if t_search else 0.0, "eval_Mm_per_s": matches / t_eval / 1e6 if t_eval else 0.0, "seconds": t_used,
                "source_stride": self.stride, "source_poses": int(len(self.ids))}


def reference_footprint_gb(g):
    """Host memory the reference's JointOpt needs for a map: info_mat_ (n_poses^2 floats, JointOptimization.cpp:1313) + ~120 B per KD node."""
    n, m = float(len(g["poses"])), float(g["offsets"][-1])
    return (4.0 * n * n + 120.0 * m) / 1e9


def ref_sample(g, steps=3):
    """cpu_baseline of the own arm: the mean of `steps` RefSampler steps (None when oracle/_ref was not built)."""
    from oracle.pyoracle import RefBackend
    if not RefBackend.available(fast=True):
        return None
    if reference_footprint_gb(g) > 24.0:
        return None                                  # the reference allocates an N x N float image (info_mat_) and one heap node per point: not runnable at this size
    smp = RefSampler(g)
    rs = [smp.step() for _ in range(steps + 1)][1:]
    out = dict(rs[-1])
    out["value"] = float(np.mean([r["value"] for r in rs]))
    out["seconds"] = float(np.sum([r["seconds"] for r in rs]))
    return out


def rebuild_fast_oracle_native():
    """liboracle_fast.so is compiled with -march=native: rebuild it on THIS box's CPU."""
    try:
        subprocess.run(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "_build", "liboracle_fast.so")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass


def load_workload_detached(name, n_poses, beams):
    """The reference arm maps none of the repo's product libraries: the synthetic map comes from the .npz cache, and when the cache is
    missing a CHILD process generates it (the generator runs through libhitl_host.so's .stfs.covars writer / loader)."""
    cache = os.path.join(os.environ.get("HITL_SYNTH_DIR", "/tmp/hitl_synth"), "%s_%s_%s.npz" % (name, n_poses, beams))
    if not os.path.exists(cache):
        subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import bench; bench.workload(%r, %r, %r)" % (ROOT, name, n_poses, beams)],
                       check=True, stdout=subprocess.DEVNULL)
    z = np.load(cache)
    out = {k: z[k] for k in z.files if k != "config_json"}
    out["config"] = json.loads(str(z["config_json"]))
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    rebuild_fast_oracle_native()
    g = load_workload_detached(args.workload, args.poses, args.beams)
    from oracle.pyoracle import RefBackend
    if RefBackend.available(fast=True) and reference_footprint_gb(g) <= 24.0:
        smp = RefSampler(g)
        step = smp.step
    else:                                            # oracle/_ref absent (built only where /root/reference exists): the oracle port
        per_step = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
        step = lambda: cpu_sample(g, seconds=per_step)
    vals = []
    for it in range(args.warmup + args.steps):
        r = step()
        if it >= args.warmup:
            vals.append(r)
    v = float(np.mean([x["value"] for x in vals]))
    last = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([x["seconds"] for x in vals])) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 search / f64 residuals",
            "data": "synthetic", "config": config_of(args, g, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                             "value_min": float(np.min([x["value"] for x in vals])), "value_max": float(np.max([x["value"] for x in vals]))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


def config_of(args, g, world=1):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "%s: synthetic corridor-lattice map, %d poses x %d beams (%d points) in .stfs.covars format, thr 0.15 m, 25 deg, cap 6, skip 1"
                        % (args.workload, len(g["poses"]), args.beams or 0, int(g["offsets"][-1])),
            "n_poses": int(len(g["poses"])), "n_points": int(g["offsets"][-1]),
            "l2": "scans + trees + record buffers exceed the 126 MB L2; a 512 MB buffer is also written between timed steps",
            "parallelism": ("source-pose shards x%d cut at equal measured work (scans+trees replicated), no collective in the search, 1 in-library ncclAllReduce (hitl_normal_eq_allreduce) of packed J^TJ/J^Tr per step" % world
                            if world > 1 else "single GPU")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--poses", type=int, default=None)
    ap.add_argument("--beams", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--balance-passes", type=int, default=8, help="multi-GPU setup: searches used to cut the source ranges at equal measured work")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--no-correction", action="store_true", help="skip the correction-latency leg")
    ap.add_argument("--no-largest-map", action="store_true", help="N > 1: skip the config-3 (20k x 1080) leg that measures the N-GPU speed-up on the largest map")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing parity check of the benchmarked search against the oracle")
    ap.add_argument("--replay", type=int, default=-1, help="BASELINE config 4: replay this many sequential corrections on --replay-workload and report per-correction latency "
                                                          "(default: 10 at N = 1 on the default workload, else 0)")
    ap.add_argument("--replay-workload", default="c4")
    ap.add_argument("--replay-seconds", type=float, default=150.0, help="wall-clock budget of the replay leg (stroke picking included)")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" or world == 1:
        use_all_host_threads()                         # the CPU legs say how many threads they used (`cores`)
    from hitl_slam_b200 import synth
    if args.beams is None:
        args.beams = synth.CONFIGS[args.workload]["beams"]
    if args.poses is None:
        args.poses = synth.CONFIGS[args.workload]["n_poses"]
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.replay < 0:
        args.replay = 10 if (world == 1 and args.workload == "c2" and args.poses == synth.CONFIGS["c2"]["n_poses"] and not args.no_correction) else 0

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / warnings go to stderr: stdout carries exactly one JSON line
    import torch
    import torch.distributed as dist
    from hitl_slam_b200 import HitlGpu, capi
    from hitl_slam_b200.sharding import shard_ranges, shard_ranges_by_work
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # rank 0 generates (and caches) the map; the others read the cache
    if rank == 0:
        g = workload(args.workload, args.poses, args.beams)
    if world > 1:
        dist.barrier()
    if rank != 0:
        g = workload(args.workload, args.poses, args.beams)
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    lo, hi = shard_ranges(g["offsets"], world)[rank]

    gpu = HitlGpu(local_rank)
    stream = torch.cuda.ExternalStream(gpu.lib.hitl_stream(gpu.ctx), device=torch.device("cuda", local_rank))
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    t0 = time.perf_counter()
    gpu.build_kdtrees()
    t_build = time.perf_counter() - t0
    nodes = gpu.get_kdtrees()
    # odometry blocks (all ranks evaluate their share: rank 0 takes them; trivial cost)
    odo = odometry_consts_host(g["poses"]) if rank == 0 else np.zeros((0, 9), np.float32)
    gpu.set_odometry_blocks(odo)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    poses_pinned = gpu.pinned_copy(poses)          # the per-call pose upload (120 KB at c2) comes from page-locked memory: no staging copy in the driver

    debug = bool(os.environ.get("HITL_BENCH_DEBUG"))
    trace = []

    if world > 1:
        # The product's own communicator (comm.cu: NCCL bound inside libhitl_gpu.so): rank 0 draws the id, torch.distributed only carries
        # the 128 bytes to the other ranks (out-of-band plumbing, like MPI would); every collective of a step is a C-ABI call.
        uid = [HitlGpu.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.comm_init(uid[0], rank, world)

    def step():
        t0 = time.perf_counter()
        info = gpu.find_stf(poses_pinned, src_lo=lo, src_hi=hi, fetch=False)
        t1 = time.perf_counter()
        gpu.set_stf_blocks_from_search(STD_DEV, CORR)
        # this rank's J^T J / J^T r blocks and, on the same stream without a host synchronisation in between, the in-library
        # ncclAllReduce(sum, f64) of the packed [H_diag | g | cost] buffer (hitl_normal_eq_allreduce; at N = 1 it is hitl_normal_eq)
        ne = gpu.normal_eq_allreduce(poses_pinned, fetch=False)
        if debug:
            trace.append((t1 - t0, time.perf_counter() - t1, info["ms_total"], ne["ms"], gpu.last_kernel_ms("allreduce") if world > 1 else 0.0))
        return info, ne

    if world > 1:
        # Setup (not a step): cut the source ranges at equal MEASURED work.  Every search reports the SM cycles it spent on
        # each source pose; summed over the ranks this is a partition-independent cost per pose (all GPUs are alike), so
        # the ranges are re-cut at equal cycle sums and the estimate is refined over a few passes.
        est = None
        for p in range(args.balance_passes):
            info_b = gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
            work = torch.from_numpy(gpu.stf_work().astype(np.float64)).cuda()
            dist.all_reduce(work)
            wk = work.cpu().numpy()
            est = wk if est is None else 0.5 * (est + wk)
            if os.environ.get("HITL_BENCH_DEBUG"):
                sys.stderr.write("[rank %d] balance %d: [%d, %d) ms_search %.3f ms_total %.3f tiles %d -> %d\n" % (rank, p, lo, hi, info_b["ms_search"], info_b["ms_total"], info_b["n_tiles"], info_b["n_tiles_next"]))
            lo, hi = shard_ranges_by_work(est, world)[rank]
    sampler = ClockSampler(local_rank)
    sampler.start()                                # before the warm-up: NVML start-up is not inside the timed region
    for w in range(args.warmup):
        info_w, _ = step()
        if os.environ.get("HITL_BENCH_DEBUG"):
            sys.stderr.write("[rank %d] warmup %d: ms_search %.3f ms_total %.3f tiles %d -> %d, range [%d, %d)\n" % (rank, w, info_w["ms_search"], info_w["ms_total"], info_w["n_tiles"], info_w["n_tiles_next"], lo, hi))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark_begin()
    launches0 = gpu.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    infos = []
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                      # evict L2 between timed steps (not timed)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            starts[k].record()
        infos.append(step())
        with torch.cuda.stream(stream):
            ends[k].record()
        # The step ends with an asynchronous all-reduce when N > 1: wait for it before the (untimed) L2 flush of the next iteration is
        # launched, otherwise the 512 MB fill runs concurrently with the collective it is not part of and delays it (measured at N = 2:
        # 6.18 -> 5.63 ms per step).
        ends[k].synchronize()
    torch.cuda.synchronize()
    launches = gpu.launch_count() - launches0
    if debug and trace:
        for k, tr in enumerate(trace[-args.steps:]):
            sys.stderr.write("[rank %d] step %d: host find_stf %.3f ms (device %.3f) | host normal_eq+all-reduce %.3f ms (device %.3f, all-reduce on stream %.3f) | events %.3f ms\n"
                             % (rank, k, tr[0] * 1e3, tr[2], tr[1] * 1e3, tr[3], tr[4], starts[k].elapsed_time(ends[k])))
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(infos[-1][0]["n_queries"]), float(infos[-1][0]["n_matches"]), float(infos[-1][0]["n_pairs"]),
                        float(infos[-1][0]["n_traversals"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    total_ms = float(t.item())
    queries, matches, pairs, trav = [int(x) for x in cnt.tolist()]
    # per-rank view of the last timed step (diagnosis of shard balance): search kernel ms, whole find_stf ms, tiles
    mine = torch.tensor([infos[-1][0]["ms_search"], infos[-1][0]["ms_total"], float(infos[-1][0]["n_tiles"]), float(hi - lo)], dtype=torch.float64, device="cuda")
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank = [[round(float(x), 3) for x in r.tolist()] for r in per_rank]
    ms_per_step = total_ms / args.steps
    evals = queries + matches
    value = evals / (ms_per_step * 1e-3) / 1e6

    # ---- parity of the benchmarked workload itself (outside every timed region; the oracle is the checker, never the thing measured) ----
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(gpu, g, poses, lo, hi, world)
            if world > 1:
                pc = torch.tensor([float(parity["chunks"]), float(parity["ok"])], dtype=torch.float64, device="cuda")
                dist.all_reduce(pc)
                parity["chunks"], parity["ok"] = int(pc[0].item()), int(pc[1].item())
            parity["all_ok"] = parity["chunks"] == parity["ok"]
        except Exception as e:
            parity = {"error": str(e)[:200]}
            if world > 1:
                dist.all_reduce(torch.zeros(2, dtype=torch.float64, device="cuda"))

    # ---- roofline of the dominant kernel (this rank's launch) ----
    peak, peak_src = load_peaks()
    ms_search = float(np.mean([i[0]["ms_search"] for i in infos]))
    my_q, my_m = infos[-1][0]["n_queries"], infos[-1][0]["n_raw_matches"]
    roofline = search_roofline(args, world, gpu, ms_search, ms_per_step, my_q, my_m, clocks, peak, peak_src)
    try:
        ms_ev = gpu.last_kernel_ms("eval_stf_kernel")
        ev_bytes = 32.0 * infos[-1][0]["n_matches"] + 160.0 * infos[-1][0]["n_pairs"]
        roofline["kernels"]["eval_stf_kernel"] = {
            "bound": "hbm", "achieved": ev_bytes / (ms_ev * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": ev_bytes / (ms_ev * 1e-3) / 1e9 / peak, "ms_kernel": ms_ev,
            "algorithmic_bytes": ev_bytes, "traffic": counter_of("eval_stf_kernel", "dram_bytes", args, world),
            "note": "32 B gathered per correspondence + 160 B per block (SURVEY.md 8d); normal-equation mode of this rank's blocks, last timed step"}
    except Exception:
        pass

    # ---- e2e through the C ABI with HOST buffers (page-locked, from hitl_host_alloc) ----
    # One step = what a caller holding the map on the host pays: scans + trees + poses H2D, the search,
    # the correspondence CSR D2H, block registration, residual + Jacobian evaluation, r + J D2H.
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))
    h_off, h_pts, h_nrm = gpu.pinned_copy(np.ascontiguousarray(g["offsets"], np.uint32)), gpu.pinned_copy(g["pts"].astype(np.float32)), gpu.pinned_copy(g["nrm"].astype(np.float32))
    h_nodes, h_poses, h_odo = gpu.pinned_copy(nodes), gpu.pinned_copy(poses), gpu.pinned_copy(odo)
    last = infos[-1][0]
    npair, nmatch = int(last["n_pairs"]), int(last["n_matches"])
    o_stf = (gpu.pinned(npair + 1, np.uint32), gpu.pinned(npair + 1, np.uint32), gpu.pinned(npair + 2, np.uint64), gpu.pinned(nmatch + 1, np.uint32), gpu.pinned(nmatch + 1, np.uint32))
    n_res, n_jac = 3 * len(odo) + 2 * npair, 18 * len(odo) + 12 * npair
    o_ev = (gpu.pinned(n_res + 1, np.float64), gpu.pinned(n_jac + 1, np.float64))
    # Opt-in (HITL_E2E_COMPACT=1, not the default until it has been validated on a GPU): the compact boundary formats — trees as one u32
    # per node (the points are already uploaded with the scans) and 16-bit point indices in the correspondence lists.
    compact = os.environ.get("HITL_E2E_COMPACT", "1") != "0"       # validated on a B200 (tests/test_gpu_parity.py::test_compact_tree_and_index_formats): the default
    if compact:
        h_compact = gpu.pinned_copy((nodes["index"].astype(np.uint32) & 0x7FFFFFFF) | (nodes["dim"].astype(np.uint32) << 31))
        o_stf16 = (o_stf[0], o_stf[1], o_stf[2], gpu.pinned(nmatch + 1, np.uint16), gpu.pinned(nmatch + 1, np.uint16))
    map_bytes = h_pts.nbytes + h_nrm.nbytes + (h_compact.nbytes if compact else h_nodes.nbytes)
    h2d = (map_bytes + world - 1) // world + h_off.nbytes + h_poses.nbytes * 2 + h_odo.nbytes      # per rank: 1 / world of the map crosses ITS PCIe link
    d2h = 0

    e2e_parts = []

    def e2e_step(upload_map=True):
        t = [time.perf_counter()]
        if upload_map:
            # N > 1: every rank holds the map on its host; each uploads 1/N of it and the slices are all-gathered over NVLink (hitl_set_*_sharded)
            (gpu.set_scans_sharded if world > 1 else gpu.set_scans)(h_off, h_pts, h_nrm)
            t.append(time.perf_counter())
            if compact:
                (gpu.set_kdtrees_compact_sharded if world > 1 else gpu.set_kdtrees_compact)(h_compact)
            else:
                (gpu.set_kdtrees_sharded if world > 1 else gpu.set_kdtrees)(h_nodes)
        else:
            t.append(t[0])
        t.append(time.perf_counter())
        info_c = gpu.find_stf(h_poses, src_lo=lo, src_hi=hi, fetch=False)
        t.append(time.perf_counter())
        out = gpu.get_stf16(info_c["n_pairs"], info_c["n_matches"], out=o_stf16) if compact else gpu.get_stf(info_c["n_pairs"], info_c["n_matches"], out=o_stf)
        t.append(time.perf_counter())
        gpu.set_odometry_blocks(h_odo)
        gpu.set_stf_blocks_from_search(STD_DEV, CORR)
        ev = gpu.eval(h_poses, fetch=True, out=o_ev)
        t.append(time.perf_counter())
        e2e_parts.append([(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])])
        return out["pair_i"].nbytes * 2 + out["pair_off"].nbytes + out["k"].nbytes * 2 + ev["r_stf"].nbytes + ev["J_stf"].nbytes + ev["r_odometry"].nbytes + ev["J_odometry"].nbytes

    e2e_warmup = max(1, min(args.warmup, 5)) if e2e_steps else 0
    for _ in range(e2e_warmup):
        e2e_step()                                 # warm-up: first touch of the result buffers; the first calls on a freshly uploaded map may re-tile
    e2e_parts = e2e_parts[-1:]                     # parts[0] = last warm-up step, parts[1:] = the timed steps
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d2h = e2e_step()
    t_e2e = (time.perf_counter() - t0) / max(e2e_steps, 1) if e2e_steps else float("inf")
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # the same step for a caller whose map is already resident (a session uploads scans and trees once, JointOptimization.cpp:1307,
    # and per correction / iteration only poses go up and results come down): reported beside the headline, never instead of it
    t_res = float("inf")
    if e2e_steps:
        e2e_step(False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step(False)
        t_res = (time.perf_counter() - t0) / e2e_steps
    tr = torch.tensor([t_res], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    pm = np.median(np.array(e2e_parts[1:1 + e2e_steps]), axis=0) if e2e_steps else np.zeros(5)
    e2e = {"value": evals / float(te.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": float(te.item()) * 1e3, "steps": e2e_steps, "warmup": e2e_warmup,
           "steps_ms": [float(sum(p)) for p in e2e_parts[1:1 + e2e_steps]],
           "parts_ms": {"set_scans": float(pm[0]), "set_kdtrees": float(pm[1]), "find_stf": float(pm[2]), "get_stf": float(pm[3]), "blocks_eval_fetch": float(pm[4])},
           "map_resident": {"value": evals / float(tr.item()) / 1e6, "ms_per_step": float(tr.item()) * 1e3, "h2d_bytes_per_step": int(h_poses.nbytes * 2 + h_odo.nbytes), "d2h_bytes_per_step": int(d2h),
                            "what": "the same step when scans + trees are already resident (uploaded once per session): poses up, correspondences + residuals + Jacobians down"},
           "timing": "host wall clock around the synchronous C-ABI calls (every call ends with a stream sync), pinned host buffers, max over ranks",
           "upload": ("sharded: each rank uploads 1/%d of scans + trees, ncclAllGather over NVLink in place (hitl_set_scans_sharded / hitl_set_kdtrees_compact_sharded); "
                      "h2d_bytes_per_step is per rank" % world) if world > 1 else "whole map over one PCIe link",
           "formats": "compact (u32 tree nodes, u16 point indices)" if compact else "hitl_kdnode trees (24 B/node), u32 point indices"}

    # ---- correction latency (second half of the BASELINE metric): one human correction on this map ----
    correction = None
    if rank == 0 and not args.no_correction:
        try:
            correction = correction_latency(gpu, g, cpu=(world == 1 and not args.no_cpu))
        except Exception as e:            # the headline line must still print
            correction = {"error": str(e)[:200]}

    if rank == 0 and correction and "error" not in correction:
        try:
            # The E-step as an HBM stream: one pass that reads every chunk (the first E-step after the world clouds changed; here the
            # chunk cull is switched off for it), L2 evicted before it.  The later E-steps of a correction skip, unread, the chunks out
            # of the stroke's reach: their duration is reported beside it and is not a bandwidth figure.
            seg = synth.pick_strokes(g)[:2].reshape(-1)
            gpu.debug_set_em_cull(False)
            ms_full = []
            for k in range(3):
                flush.fill_(k)
                torch.cuda.synchronize()
                gpu.em_inliers(seg, fetch=False)
                ms_full.append(gpu.last_kernel_ms("em_inliers_kernel"))
            gpu.debug_set_em_cull(True)
            gpu.em_inliers(seg, fetch=False)             # records the chunk boxes
            n_in = gpu.em_inliers(seg, fetch=False)
            ms_culled = gpu.last_kernel_ms("em_inliers_kernel")
            ms_em = float(np.median(ms_full))
            em_bytes = 8.0 * float(g["offsets"][-1])
            roofline["kernels"]["em_inliers_kernel"] = {
                "bound": "hbm", "achieved": em_bytes / (ms_em * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": em_bytes / (ms_em * 1e-3) / 1e9 / peak, "ms_kernel": ms_em,
                "algorithmic_bytes": em_bytes, "traffic": counter_of("em_inliers_kernel", "dram_bytes", args, world),
                "ms_kernel_culled": ms_culled, "inliers": int(n_in),
                "note": "8 B per world-frame point per E-step (SURVEY.md 8d): a full pass over the resident world clouds, L2 evicted first, median of 3; "
                        "launch-latency sized at this map (28.6 MB).  ms_kernel_culled: a later E-step of the same correction, which skips the "
                        "2048-point chunks whose bounding box is out of the stroke's reach (not a bandwidth figure)"}
        except Exception as e:
            roofline["kernels"]["em_inliers_kernel"] = {"error": str(e)[:200]}

    replay = None
    if rank == 0 and args.replay > 0:
        try:
            gr = g if args.replay_workload == args.workload else workload(args.replay_workload, synth.CONFIGS[args.replay_workload]["n_poses"], synth.CONFIGS[args.replay_workload]["beams"])
            replay = correction_replay(gpu, gr, args.replay, budget_s=args.replay_seconds if args.replay > 10 else min(args.replay_seconds, 80.0))
            replay["workload"] = args.replay_workload
        except Exception as e:
            replay = {"error": str(e)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rebuild_fast_oracle_native()
        port = cpu_sample(g, seconds=args.cpu_seconds)
        cpu = ref_sample(g)
        if cpu is None:
            cpu = port
        else:
            cpu["port"] = port            # the oracle port on a full-map sample, timed beside the reference's own code

    largest = None
    if world > 1 and not args.no_largest_map and args.workload == "c2":
        try:
            gpu.comm_destroy()
            gpu.close()                                  # frees the c2 context before the 21.4 M-point map is loaded
            largest = largest_map_leg(args, rank, world, local_rank)
        except Exception as e:
            largest = {"error": str(e)[:300]}

    if rank == 0:
        cfg = config_of(args, g, world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 search / f64 residuals", "data": "synthetic", "config": cfg,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "parity_checked": parity, "cpu_baseline": cpu, "correction_latency": correction, "correction_replay": replay, "largest_map": largest,
                "detail": {"queries_per_step": queries, "jacobian_evals_per_step": matches, "residual_blocks": pairs, "tree_walks_per_step": trav, "tile_pairs_per_step": int(infos[-1][0]["n_tile_pairs"]),
                           "kdtree_build_device_s": t_build, "ms_find_stf": float(np.mean([i[0]["ms_total"] for i in infos])), "ms_normal_eq": float(np.mean([i[1]["ms"] for i in infos])),
                           "per_rank_[ms_search,ms_find_stf,tiles,source_poses]": per_rank}}
        emit_line(line)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


def counter_of(kernel, key, args, world):
    """Per-launch hardware counters of the committed `ncu --set full` capture (profiles/ncu_counters.json: workload c2, one GPU, this
    code).  They describe ONE launch on c2 at N = 1; other workloads / shard sizes get None rather than a number that is not theirs."""
    if args.workload != "c2" or world != 1 or args.poses != 5000 or args.beams != 720:
        return None
    try:
        c = json.load(open(os.path.join(ROOT, "profiles", "ncu_counters.json")))[kernel]
        if key == "dram_bytes":
            return c["dram_bytes_read"] + c["dram_bytes_write"]
        return c.get(key)
    except Exception:
        return None


def search_roofline(args, world, gpu, ms_search, ms_per_step, n_queries, n_raw_matches, clocks, hbm_peak, peak_src):
    """stf_search_kernel is ISSUE-bound (divergent tree walks): its physical roofline is the warp-instruction issue rate
    of the chip, 4 schedulers x SMs x SM clock.  achieved = warp-instructions of one launch (smsp__inst_executed.sum of the committed
    ncu capture of this code on this workload) / the kernel's duration measured live with CUDA events.  The HBM figures stay beside it:
    `traffic` / `frac_measured_traffic` are the DRAM bytes the launch really moved, and `hbm_stream_model` is SURVEY.md 8d's pair-tile
    stream model (40 B per reference-semantics query + 8 B per match), which is NOT a physical fraction: the kernel proves > 98 % of
    those queries empty with exact bitmap tests and streams nothing for them, so the model exceeds 1."""
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    issue_peak = gpu.sm_count() * 4 * sm_mhz * 1e6 / 1e9                     # G warp-instructions / s
    inst = counter_of("stf_search_kernel", "inst_executed", args, world)
    traffic = counter_of("stf_search_kernel", "dram_bytes", args, world)
    achieved = inst / (ms_search * 1e-3) / 1e9 if inst else None
    alg_bytes = 40.0 * n_queries + 8.0 * n_raw_matches
    model = alg_bytes / (ms_search * 1e-3) / 1e9
    return {"kernel": "stf_search_kernel", "bound": "issue", "achieved": achieved, "peak": issue_peak, "unit": "G warp-inst/s",
            "frac": achieved / issue_peak if achieved else None,
            "traffic": traffic, "frac_measured_traffic": traffic / (ms_search * 1e-3) / 1e9 / hbm_peak if traffic else None,
            "issue_active": counter_of("stf_search_kernel", "issue_active_pct", args, world), "threads_per_inst": counter_of("stf_search_kernel", "threads_per_inst", args, world),
            "warp_instructions": inst, "ms_kernel": ms_search, "share_of_step": ms_search / ms_per_step,
            "peak_source": "issue: %d SMs x 4 schedulers x %.0f MHz (SM clock sampled during the timed region); HBM: %s" % (gpu.sm_count(), sm_mhz, peak_src),
            "hbm_stream_model": {"achieved": model, "peak": hbm_peak, "unit": "GB/s", "frac": model / hbm_peak, "algorithmic_bytes": alg_bytes, "physical": False,
                                 "note": "40 B per reference-semantics query + 8 B per match (SURVEY.md 8d pair-tile model); exceeds 1 because exact culling "
                                         "proves most counted queries empty without streaming their scans - not evidence of bandwidth"},
            "kernels": {}}


def parity_check(gpu, g, poses, lo, hi, world):
    """The search the bench just timed (same context, same adaptive tiling, same source shard) fetched once more and compared bit for bit
    with the CPU oracle's FindSTFCorrespondences on a few source-pose chunks of this rank's shard against ALL targets — the last poses of
    the shard included, where tiles are heaviest.  One rank: 4 chunks of 16 poses on all host threads; N ranks: one 4-pose chunk per rank
    (torchrun leaves each rank one OpenMP thread)."""
    from oracle.pyoracle import Oracle
    S = Oracle().scans(g["offsets"], g["pts"], g["nrm"])
    res = gpu.find_stf(poses, src_lo=lo, src_hi=hi)
    width = 16 if world == 1 else 4
    width = min(width, hi - lo)
    starts = sorted(set([hi - width] if world > 1 else [int(x) for x in np.linspace(lo, hi - width, 4)]))
    t0 = time.perf_counter()
    checks = S.check_chunks(res, poses, [(a, a + width) for a in starts])
    return {"chunks": len(checks), "ok": int(sum(1 for c in checks if c[2])), "source_chunks": [[c[0], c[1]] for c in checks], "matches_compared": int(sum(c[4] for c in checks)),
            "against": "oracle FindSTFCorrespondences (parity build) on these source poses vs all targets: pair_i, pair_j, pair_off, k, idx bit for bit", "seconds": time.perf_counter() - t0}


def correction_latency(gpu, g, cpu=True, reps=5):
    """One human correction (colinear, two strokes picked on a revisited wall) through the C++ host mirror:
    world-frame clouds -> EMInput::Run (E-steps and M-steps on the GPU, observation sets on the GPU,
    ordering on the host) -> constraint targets -> problem build + one batched evaluation of every block
    (residuals + Jacobians back on the host).  Wall clock around synchronous calls; the host LM solve is
    reported separately (latency excludes the host linear solve)."""
    from hitl_slam_b200 import HostSession, synth
    strokes = synth.pick_strokes(g)
    sess = HostSession(gpu)
    sess.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    lat, solve, parts, em = [], [], [], None
    for it in range(reps + 1):
        sess.clear_constraints()
        sess.set_poses(g["poses"])
        t0 = time.perf_counter()
        sess.world_transform(keep_host_copy=False)
        ta = time.perf_counter()
        em = sess.em_run(4, strokes)
        tb = time.perf_counter()
        nc = sess.add_constraints_from_em()
        tc = time.perf_counter()
        sess.evaluate_block(0, with_stf=False)
        t1 = time.perf_counter()
        summ = sess.joint_opt_run(post=False)
        t2 = time.perf_counter()
        if it:                              # first pass warms allocations
            lat.append((t1 - t0) * 1e3)
            solve.append((t2 - t1) * 1e3)
            parts.append([(ta - t0) * 1e3, (tb - ta) * 1e3, (tc - tb) * 1e3, (t1 - tc) * 1e3])
    pm = np.median(np.array(parts), axis=0)
    out = {"ms": float(np.median(lat)), "ms_min": float(np.min(lat)), "ms_with_host_solve": float(np.median(lat) + np.median(solve)), "unit": "ms per correction",
           "parts_ms": {"world_transform": float(pm[0]), "em_run": float(pm[1]), "constraint_targets": float(pm[2]), "build_and_evaluate_blocks": float(pm[3])},
           "what": "world transform + EM (E-step AND M-step on the GPU, the rounds of both strokes chained with one host wait: hitl_em_refit_chain; observation sets on the GPU, ordering on the host) + constraint targets + build & one batched evaluation of all odometry+human blocks",
           "em_rounds": list(em["rounds"]), "corrected_poses": int(len(em["corrected"])), "anchor_poses": int(len(em["anchor"])), "human_blocks": int(nc),
           "solver_steps": int(summ["successful_steps"] + summ["unsuccessful_steps"]), "n_points": int(g["offsets"][-1])}
    sess.close()
    if cpu:
        from oracle.pyoracle import Oracle
        orc = Oracle(fast=True)
        S = orc.scans(g["offsets"], g["pts"], g["nrm"], build_trees=False)
        t0 = time.perf_counter()
        world = S.world_transform(g["poses"])
        ref = orc.em_run(g["offsets"], world, strokes)
        consts = orc.odometry_consts(g["poses"])
        orc.eval_odometry(consts, g["poses"].astype(np.float64))
        out["cpu_ms"] = (time.perf_counter() - t0) * 1e3
        out["cpu_what"] = "oracle port, 1 thread as in the reference: world transform + EM (same strokes, %d rounds) + odometry block evaluation" % ref["rounds"]
        try:
            out["cpu_reference"] = ref_correction_cpu(g, strokes)
        except Exception as e:               # a baseline leg must never take the headline line down
            out["cpu_reference"] = {"error": str(e)[:200]}
    return out


def ref_correction_cpu(g, strokes):
    """The same correction on the REFERENCE's own code (oracle/_ref/libhitl_ref_fast.so), one thread as the reference runs it, solve excluded
    like the GPU figure: world clouds (JointOpt::CopyTempLaserScans) + EMInput::Run (E-steps, M-steps, observation sets, ordering; the call
    includes one whole-cloud copy, as HitLSLAM.cpp:400 makes one) + AppExpCorrect::Run + Backprop::Run + the odometry and human blocks that
    AddOdometryConstraints / AddHumanConstraints build, evaluated once through AutoDiffCostFunction.  Session set-up (cloud copies into
    JointOpt, BuildKDTrees) is not timed.  Returns None when the library is absent."""
    from oracle.pyoracle import RefBackend
    if not RefBackend.available(fast=True):
        return None
    ref = RefBackend(fast=True)
    n = len(g["poses"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    cov = np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (n, 1))
    t0 = time.perf_counter()
    world = J.world_clouds()
    t1 = time.perf_counter()
    em = ref.em_run(g["offsets"], world, strokes)
    t2 = time.perf_counter()
    applied = em["backprop"][0] >= 0 and em["backprop"][1] >= 1
    n_hc = 0
    if applied:
        p1, c3, hc_i, hc_f = ref.app_exp_run(4, em["segs"], g["poses"], em["corrected"], em["anchor"])
        p2, _ = ref.backprop(p1, cov, em["backprop"][0], em["backprop"][1], c3)
        J.set_poses(p2)
        J.set_human_constraints([(hc_i, hc_f)])
        n_hc = len(hc_i)
    x = J.pose_array()
    J.eval_blocks(0, x, n)
    if n_hc:
        J.eval_blocks(1, x, n_hc)
    t3 = time.perf_counter()
    return {"ms": (t3 - t0) * 1e3, "ms_world": (t1 - t0) * 1e3, "ms_em": (t2 - t1) * 1e3, "ms_correct_backprop_blocks": (t3 - t2) * 1e3, "human_blocks": int(n_hc),
            "what": "reference's own code, 1 thread: CopyTempLaserScans + EMInput::Run + AppExpCorrect::Run + Backprop::Run + odometry/human blocks built and evaluated once (no solve)"}


def correction_replay(gpu, g, n_corrections, cpu_every=10, budget_s=150.0):
    """BASELINE config 4: a replay of sequential human corrections on one map, each drawn on the map AS IT IS after the
    previous ones (strokes picked from the current poses, untimed).  Per correction, as HitLSLAM::Run wires it: world clouds ->
    EM -> explicit correction -> COP-SLAM back-propagation -> constraint targets -> joint optimisation.  The latency of the
    accelerated path excludes the host LM solve, which is reported beside it; every cpu_every-th correction
    is also run through the single-threaded oracle chain on the same inputs."""
    from hitl_slam_b200 import HostSession, synth
    from oracle.pyoracle import Oracle
    sess = HostSession(gpu)
    sess.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    n = len(g["poses"])
    cov = (np.ascontiguousarray(g["cov"], np.float32).reshape(n, 9).copy() if "cov" in g
           else np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (n, 1)))
    orc = Oracle(fast=True)
    S = orc.scans(g["offsets"], g["pts"], g["nrm"], build_trees=False)
    cur = dict(g)
    lat, solve, parts, cpu_ms, spans, blocks, applied = [], [], [], [], [], [], []
    start, t_begin = 0, time.perf_counter()
    dry_runs = 0
    for c in range(n_corrections):
        if time.perf_counter() - t_begin > budget_s:
            break
        cur["poses"] = sess.poses()[0]
        # Untimed: the "human" draws two strokes on a revisited wall of the map AS IT IS NOW.  A real user draws strokes the tool accepts;
        # here every candidate pair is first put through the same EM (dry run, it changes no session state) and is drawn again elsewhere
        # when EM finds the observers of the two strokes interleaved in time (HitLSLAM::Run would stop after EM, HitLSLAM.cpp:413).
        # The first corrections close real loop-closure gaps; later ones also accept walls with little separation left.
        strokes = None
        sess.world_transform(keep_host_copy=False)
        for attempt in range(8):
            try:
                cand, start = synth.pick_strokes(cur, min_sep=(0.04 if c < 4 else 0.015) if attempt < 5 else 0.0, start=start, return_next=True)
            except RuntimeError:
                if attempt >= 5:
                    break
                continue
            dry = sess.em_run(4, cand)
            dry_runs += 1
            strokes = cand
            if dry["backprop"][0] >= 0 and dry["backprop"][1] >= 1:
                break
        if strokes is None:
            break
        if cpu_every and c % cpu_every == 0:
            cov_c = cov.copy()
            t0 = time.perf_counter()
            world = S.world_transform(cur["poses"])
            em = orc.em_run(g["offsets"], world, strokes)
            if em["backprop"][0] >= 0 and em["backprop"][1] >= 1:
                p1, c3 = orc.app_exp_corrections(4, em["segs"], cur["poses"], em["corrected"])
                if c3 is not None:
                    p1, cov_c = orc.backprop(p1, cov_c, em["backprop"][0], em["backprop"][1], c3)
                orc.eval_odometry(orc.odometry_consts(p1), p1.astype(np.float64))
            cpu_ms.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        sess.world_transform(keep_host_copy=False)
        out = sess.correct(4, strokes, cov=cov, solve=True)
        t1 = time.perf_counter()
        total = (t1 - t0) * 1e3
        lat.append(total - out["ms"]["joint_opt"])
        solve.append(out["ms"]["joint_opt"])
        parts.append([out["ms"]["em"], out["ms"]["explicit"], out["ms"]["backprop"], out["ms"]["backprop_device"]])
        spans.append(out["backprop"][1] - out["backprop"][0] if out["applied"] else 0)
        blocks.append(out["n_constraints"])
        applied.append(bool(out["applied"]))
    sess.close()
    if not lat:
        return {"error": "no usable stroke pair on this map"}
    attempted = len(lat)
    keep = np.array(applied, bool) if any(applied) else np.ones(len(lat), bool)
    # statistics over the corrections that went all the way (EM found cleanly ordered observers of both strokes); the others stop
    # after EM, as HitLSLAM::Run does when the back-propagation bounds are invalid (HitLSLAM.cpp:413)
    lat, solve, parts = np.array(lat)[keep], np.array(solve)[keep], np.array(parts)[keep]
    spans = [sp for sp, k in zip(spans, keep) if k]
    return {"corrections": int(len(lat)), "attempted": int(attempted), "applied_fraction": float(len(lat)) / max(attempted, 1), "stroke_dry_runs": int(dry_runs), "unit": "ms per correction",
            "latency_ms": {"median": float(np.median(lat)), "p90": float(np.percentile(lat, 90)), "max": float(lat.max()), "min": float(lat.min())},
            "host_solve_ms": {"median": float(np.median(solve)), "max": float(solve.max())},
            "parts_ms_median": {"em": float(np.median(parts[:, 0])), "explicit_correction": float(np.median(parts[:, 1])),
                                "backprop": float(np.median(parts[:, 2])), "backprop_device": float(np.median(parts[:, 3]))},
            "backprop_span_poses": {"median": float(np.median(spans)), "max": int(max(spans))}, "human_blocks_total": int(sum(blocks)),
            "cpu_ms": {"median": float(np.median(cpu_ms)) if cpu_ms else None, "max": float(max(cpu_ms)) if cpu_ms else None, "samples": len(cpu_ms),
                       "what": "oracle port, 1 thread as in the reference: world transform + EM + explicit correction + back-propagation + odometry block evaluation"},
            "what": "world transform + EM + explicit correction + back-propagation (pose update on the GPU) + constraint targets; host LM solve listed separately",
            "n_poses": int(n), "n_points": int(g["offsets"][-1])}


def largest_map_leg(args, rank, world, local_rank, name="c3", steps=3, warmup=4):
    """north_star: ">= 5x at 8 GPUs on the largest synthetic map".  BASELINE config 3 (20 000 poses x 1080 beams, 21.4 M points) sharded over
    the N ranks exactly like the headline step (source ranges cut at equal measured work, replicated scans + trees, one in-library
    all-reduce per step), timed as the max over ranks of CUDA-event time per step; then rank 0 ALONE runs the same step over the whole
    map on its one GPU, in the same process and run, and `speedup_vs_1` is the ratio of the two."""
    import torch
    import torch.distributed as dist
    from hitl_slam_b200 import HitlGpu, synth
    from hitl_slam_b200.sharding import shard_ranges, shard_ranges_by_work
    cfg = synth.CONFIGS[name]
    if rank == 0:
        g = workload(name, cfg["n_poses"], cfg["beams"])
    dist.barrier()
    if rank != 0:
        g = workload(name, cfg["n_poses"], cfg["beams"])
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    gpu = HitlGpu(local_rank)
    stream = torch.cuda.ExternalStream(gpu.lib.hitl_stream(gpu.ctx), device=torch.device("cuda", local_rank))
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.build_kdtrees()
    uid = [HitlGpu.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    gpu.comm_init(uid[0], rank, world)
    gpu.set_odometry_blocks(odometry_consts_host(g["poses"]) if rank == 0 else np.zeros((0, 9), np.float32))
    pp = gpu.pinned_copy(poses)
    lo, hi = shard_ranges(g["offsets"], world)[rank]
    est = None
    for _ in range(3):                                   # setup: cut the source ranges at equal measured work
        gpu.find_stf(pp, src_lo=lo, src_hi=hi, fetch=False)
        work = torch.from_numpy(gpu.stf_work().astype(np.float64)).cuda()
        dist.all_reduce(work)
        wk = work.cpu().numpy()
        est = wk if est is None else 0.5 * (est + wk)
        lo, hi = shard_ranges_by_work(est, world)[rank]

    def one(lo_, hi_, collective):
        info = gpu.find_stf(pp, src_lo=lo_, src_hi=hi_, fetch=False)
        gpu.set_stf_blocks_from_search(STD_DEV, CORR)
        ne = gpu.normal_eq_allreduce(pp, fetch=False) if collective else gpu.normal_eq(pp, fetch=False)
        return info, ne

    def timed(lo_, hi_, collective, sync_ranks):
        for _ in range(warmup):
            one(lo_, hi_, collective)
        tot, last = 0.0, None
        for _ in range(steps):
            torch.cuda.synchronize()
            if sync_ranks:
                dist.barrier()
            with torch.cuda.stream(stream):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            last = one(lo_, hi_, collective)
            with torch.cuda.stream(stream):
                b.record()
            b.synchronize()
            tot += a.elapsed_time(b)
        return tot / steps, last

    ms_n, last = timed(lo, hi, True, True)
    t = torch.tensor([ms_n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([float(last[0]["n_queries"]), float(last[0]["n_matches"]), float(last[0]["n_pairs"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(cnt)
    per_rank = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(per_rank, torch.tensor([last[0]["ms_search"], last[0]["ms_total"], float(hi - lo)], dtype=torch.float64, device="cuda"))
    ms_n = float(t.item())
    out = None
    gpu.comm_destroy()
    if rank == 0:
        ms_1, last1 = timed(0, n, False, False)          # the whole map on ONE GPU (this rank's), same code, same run
        evals = float(cnt[0].item() + cnt[1].item())
        assert int(last1[0]["n_queries"]) == int(cnt[0].item()) and int(last1[0]["n_matches"]) == int(cnt[1].item()), "sharded and single-GPU searches disagree"
        out = {"workload": "%s: %d poses x %d beams (%d points)" % (name, n, cfg["beams"], int(g["offsets"][-1])), "n_gpus": world,
               "ms_per_step": ms_n, "ms_per_step_1gpu": ms_1, "speedup_vs_1": ms_1 / ms_n, "value": evals / (ms_n * 1e-3) / 1e6, "value_1gpu": evals / (ms_1 * 1e-3) / 1e6, "unit": UNIT,
               "queries_per_step": int(cnt[0].item()), "jacobian_evals_per_step": int(cnt[1].item()), "residual_blocks": int(cnt[2].item()),
               "counts_equal_single_gpu": True, "steps": steps, "warmup": warmup,
               "per_rank_[ms_search,ms_find_stf,source_poses]": [[round(float(x), 3) for x in r.tolist()] for r in per_rank]}
    gpu.close()
    dist.barrier()
    return out


def odometry_consts_host(poses_f32):
    """PoseConstraint constants of AddOdometryConstraints (JointOptimization.cpp:736-825) from the float poses
    (host-side problem building; the non-degenerate branch, float arithmetic)."""
    p = np.asarray(poses_f32, np.float32)
    t = (p[1:, :2] - p[:-1, :2]).astype(np.float32)
    a = (-p[:-1, 2]).astype(np.float32)
    c, s = np.cos(a).astype(np.float32), np.sin(a).astype(np.float32)
    rx, ry = (c * t[:, 0] - s * t[:, 1]).astype(np.float32), (s * t[:, 0] + c * t[:, 1]).astype(np.float32)
    nr = np.sqrt(rx * rx + ry * ry).astype(np.float32)
    nr[nr == 0] = 1
    rx, ry = rx / nr, ry / nr
    rot = (p[1:, 2] - p[:-1, 2]).astype(np.float64)
    rot = (rot - 2 * np.pi * np.rint(rot / (2 * np.pi))).astype(np.float32)
    out = np.zeros((len(p) - 1, 9), np.float32)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = rx, ry, -ry, rx
    out[:, 4], out[:, 5], out[:, 6] = 0.03, 0.03, 0.01
    out[:, 7] = np.sqrt(t[:, 0] ** 2 + t[:, 1] ** 2)
    out[:, 8] = rot
    return out


if __name__ == "__main__":
    main()
