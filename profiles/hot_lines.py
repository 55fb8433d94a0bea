#!/usr/bin/env python
"""Top source lines of one kernel in an .ncu-rep captured with --import-source on (needs -lineinfo).
usage: python profiles/hot_lines.py x.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys


def main(path, kernel, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    lines = []
    for r in rows:
        if len(r) > ct and r[0].isdigit():
            try:
                lines.append((int(r[ci]), int(r[ct]), int(r[cs]), r[0], r[1].strip()))
            except ValueError:
                pass
    tot_i, tot_t, tot_s = sum(x[0] for x in lines), sum(x[1] for x in lines), sum(x[2] for x in lines)
    print("# %s  kernel ~ %s : %d warp instructions, %.2f threads/instruction, %d samples" % (path, kernel, tot_i, tot_t / max(tot_i, 1), tot_s))
    print("%7s %7s %6s  %5s  %s" % ("inst%", "smpl%", "thr/i", "line", "source"))
    for i, t, s, ln, src in sorted(lines, reverse=True)[:top]:
        print("%6.2f%% %6.2f%% %6.2f  %5s  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), t / max(i, 1), ln, src[:150]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
