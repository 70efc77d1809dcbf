#!/usr/bin/env python
"""Per-launch hardware counters of the kernels bench.py's `roofline` cites, from `ncu --set full` captures -> profiles/ncu_counters.json.

  python profiles/ncu_counters.py stf_search_kernel=gpurun_out/prof_search_X.ncu-rep eval_stf_kernel=... em_inliers_kernel=...

Each capture holds ONE warmed launch of the kernel on workload c2 (5 000 x 720) on one B200 (profiles/run_gpu_round.sh)."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = {"dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write", "smsp__inst_executed.sum": "inst_executed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
        "gpu__time_duration.sum": "ncu_duration", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1.0, "": 1.0, "%": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    path = os.path.join(HERE, "ncu_counters.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    srcs = []
    for arg in sys.argv[1:]:
        kernel, rep = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                rec[KEYS[h]] = float(v.replace(",", "")) * SCALE.get(u, 1.0)
        if "ncu_duration" in rec:
            rec["ncu_duration_ms"] = rec.pop("ncu_duration")
        out[kernel] = rec
        srcs.append("%s: %s" % (kernel, os.path.basename(rep)))
    out["source"] = "ncu --set full --clock-control none --import-source on, workload c2 (5 000 x 720), one B200, one warmed launch each; " + "; ".join(srcs) + \
                    (" | " + out["source"] if out.get("source") and not srcs else "")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
