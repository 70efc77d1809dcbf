#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the numbers DESIGN.md / bench.py cite.

  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for ki, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== launch %d: %s" % (ki, name[:100]))
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                print("  %-72s %-12s %s" % (h, u, v))
        print("  -- warp stall reasons (pct of samples, issue-stalled) --")
        for h, u, v in zip(hdr, units, vals):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                print("  %-72s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) < 3:
        return
    h = src[1]
    ia, isrc, ismp, iex, ith = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in src[2:] if len(r) > iex and r[iex].isdigit()]
    tot_s = sum(int(r[ismp] or 0) for r in body) or 1
    tot_i = sum(int(r[iex] or 0) for r in body) or 1
    print("== SASS hot spots: %d instructions, %d samples, %d warp-instructions executed" % (len(body), tot_s, tot_i))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][ismp] or 0))[:top]
    for k in sorted(order):
        r = body[k]
        st = sorted(((int(r[h.index(c)] or 0), c) for c in stalls), reverse=True)[:2]
        print("  #%-4d %-58s smp %5.2f%%  exec %5.2f%%  thr %5s  %s" % (k, r[isrc][:58], 100 * int(r[ismp] or 0) / tot_s, 100 * int(r[iex]) / tot_i, r[ith][:5],
                                                                     " ".join("%s=%d" % (c.replace("stall_", ""), v) for v, c in st if v)))
    agg = {}
    for r in body:
        for c in stalls:
            agg[c] = agg.get(c, 0) + int(r[h.index(c)] or 0)
    print("== stall samples by reason: " + ", ".join("%s %.1f%%" % (c.replace("stall_", ""), 100 * v / tot_s) for c, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))


if __name__ == "__main__":
    main()
