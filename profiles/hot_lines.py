#!/usr/bin/env python
"""Top source lines of one kernel in an .ncu-rep captured with --import-source on (needs -lineinfo), with their file, plus
totals per file and per user-given line range.
usage: python profiles/hot_lines.py x.ncu-rep kernel_regex [top] [file:lo-hi=label ...]"""
import csv
import io
import os
import subprocess
import sys


def main(path, kernel, top=40, ranges=()):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    lines, cur = [], "?"
    for r in rows:
        if r and r[0] == "File Path":
            cur = os.path.basename(r[1])
        elif len(r) > ct and r[0].isdigit():
            try:
                lines.append((int(r[ci]), int(r[ct]), int(r[cs]), cur, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot_i, tot_t, tot_s = sum(x[0] for x in lines), sum(x[1] for x in lines), sum(x[2] for x in lines)
    print("# %s  kernel ~ %s : %d warp instructions, %.2f threads/instruction, %d samples" % (path, kernel, tot_i, tot_t / max(tot_i, 1), tot_s))

    def agg(label, sel):
        i, t, s = sum(x[0] for x in sel), sum(x[1] for x in sel), sum(x[2] for x in sel)
        print("%6.2f%% %6.2f%% %6.2f  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), t / max(i, 1), label))

    print("%7s %7s %6s  per file" % ("inst%", "smpl%", "thr/i"))
    for f in sorted({x[3] for x in lines}):
        agg(f, [x for x in lines if x[3] == f])
    if ranges:
        print("%7s %7s %6s  per range" % ("inst%", "smpl%", "thr/i"))
        for spec in ranges:
            loc, label = spec.split("=")
            f, lh = loc.split(":")
            lo, hi = [int(v) for v in lh.split("-")]
            agg("%s (%s)" % (label, loc), [x for x in lines if x[3] == f and lo <= x[4] <= hi])
    print("%7s %7s %6s  %-22s %s" % ("inst%", "smpl%", "thr/i", "file:line", "source"))
    for i, t, s, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%6.2f%% %6.2f%% %6.2f  %-22s %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), t / max(i, 1), "%s:%d" % (f, ln), src[:140]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4:])
