"""Extracts the reference's own Triangle_3 x Triangle_3 known answers into tests/golden/tri_tri_kats.json.

Source: Intersections_3/test/Intersections_3/triangle_3_triangle_3_intersection.cpp (run in THIS container, where /root/reference exists; the GPU box only
sees the committed JSON).  Every `t1=Triangle(...); t2=Triangle(...)` pair of that file is followed by assertions on CGAL::intersection(t1, t2) and (t2, t1):
`obj.empty()` -> do_intersect is false, any object_cast -> true.  Coordinates are double literals or NT(p, q) rationals that are exact doubles (1/2).
"""
import json
import os
import re
import sys
from fractions import Fraction

REF = "/root/reference/Intersections_3/test/Intersections_3/triangle_3_triangle_3_intersection.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tri_tri_kats.json")


def number(tok):
    tok = tok.strip()
    m = re.fullmatch(r"NT\(\s*(-?\d+)\s*,\s*(-?\d+)\s*\)", tok)
    if m:
        fr = Fraction(int(m.group(1)), int(m.group(2)))
        assert Fraction(float(fr)) == fr, tok
        return float(fr)
    return float(tok)


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def triangle(text):
    pts = re.findall(r"Point\(((?:[^()]|\([^()]*\))*)\)", text)
    assert len(pts) == 3, text
    return [number(c) for p in pts for c in split_top(p)]


def main():
    lines = open(REF).read().split("\n")
    kats, t1, t2, line1 = [], None, None, 0
    for no, ln in enumerate(lines, 1):
        s = ln.strip()
        if s.startswith("t1=Triangle"):
            t1, line1, t2 = triangle(s), no, None
        elif s.startswith("t2=Triangle"):
            t2 = triangle(s)
            # the first assertion after the pair tells whether the intersection is empty
            for k in range(no, min(no + 4, len(lines))):
                if "assert(" in lines[k]:
                    kats.append({"line": line1, "t1": t1, "t2": t2, "intersect": "empty()" not in lines[k]})
                    break
            else:
                raise SystemExit("no assertion after line %d" % no)
    with open(OUT, "w") as f:          # one pair per line
        f.write('{"source": "Intersections_3/test/Intersections_3/triangle_3_triangle_3_intersection.cpp",\n "kats": [\n')
        f.write(",\n".join("  " + json.dumps(k) for k in kats))
        f.write("\n ]}\n")
    print(len(kats), "pairs,", sum(not k["intersect"] for k in kats), "disjoint ->", OUT)


if __name__ == "__main__":
    sys.exit(main())
