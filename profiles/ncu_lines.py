#!/usr/bin/env python
"""Per-source-line roll-up of an .ncu-rep captured with --import-source on and -lineinfo:
warp-instructions executed, samples and average active threads per CUDA source line.

  python profiles/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = "?", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 10 and r[0].isdigit():
            d = dict(zip(hdr, r))

            def num(k):
                v = d.get(k, "0")
                return int(v) if v.isdigit() else 0
            lines.append((cur_file, int(r[0]), r[1].strip(), num("Instructions Executed"), num("# Samples"), num("Thread Instructions Executed")))
    tot_i = sum(x[3] for x in lines) or 1
    tot_s = sum(x[4] for x in lines) or 1
    print("== %d source lines, %d warp-instructions, %d samples" % (len(lines), tot_i, tot_s))
    for f, ln, src, ins, smp, tins in sorted(lines, key=lambda x: -x[3])[:top]:
        print("  %-14s:%-4d inst %5.2f%%  smp %5.2f%%  thr %4.1f  %s" % (f, ln, 100 * ins / tot_i, 100 * smp / tot_s, tins / max(ins, 1), src[:90]))


if __name__ == "__main__":
    main()
