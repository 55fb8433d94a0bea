"""OFF ingest on the device (csrc/ingest.cuh): the decimal-literal parser (Eisel-Lemire) compiled for the host against Python's float(),
which is correctly rounded like the `std::istream >> double` of the reference's reader (Stream_support/include/CGAL/IO/OFF/File_scanner_OFF.h)."""
import struct

import numpy as np

import hc


def check(literals):
    vals, status = hc.parse_decimals(literals)
    for s, v, st in zip(literals, vals, status):
        if st == 2:                                                         # more than 19 digits and they matter: the library hands it to strtod
            assert sum(c.isdigit() for c in s.split("e")[0].split("E")[0].lstrip("+-0.")) > 19, s
            continue
        assert st == 0, (s, st)
        want = float(s)
        assert struct.pack("<d", v) == struct.pack("<d", want), (s, v, want)
    return status


def test_literals_round_like_strtod():
    rng = np.random.default_rng(61)
    lits = ["0", "-0", "0.0", "-0.0", "+1", "1.", ".5", "-.25", "1e0", "1E+2", "1e-2", "123456789012345678", "1234567890123456789", "0.1", "0.2", "0.3",
            "9007199254740993", "9007199254740992", "9007199254740991", "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "2.2250738585072014e-308",
            "2.2250738585072011e-308", "1.7976931348623157e308", "1.7976931348623159e308", "1e309", "1e-400", "8.98846567431158e307", "0.000001", "100000000000000000000",
            "1e23", "8.5e-5", "3.14159265358979323", "299792458", "6.02214076e23", "1.602176634e-19", "0.30000000000000004", "5e-324", "1e22", "1e-22"]
    # doubles printed with 17 significant digits (what mesh writers emit) and with fewer digits; random digit strings; halfway cases between adjacent doubles
    x = np.concatenate([rng.uniform(-1, 1, 4000), rng.uniform(-1e6, 1e6, 2000), rng.normal(size=2000) * 10.0 ** rng.integers(-300, 300, 2000)])
    lits += ["%.17g" % v for v in x] + ["%.9g" % v for v in x[:3000]] + ["%.6f" % v for v in x[:3000]] + [repr(float(v)) for v in x[:3000]]
    for _ in range(4000):
        nd = int(rng.integers(1, 20))
        digits = "".join(str(d) for d in rng.integers(0, 10, nd))
        p = int(rng.integers(0, nd + 1))
        lits.append(("-" if rng.integers(0, 2) else "") + (digits[:p] or "0") + "." + digits[p:] + ("e%d" % rng.integers(-330, 310) if rng.integers(0, 2) else ""))
    for v in np.abs(x[:1500]):                                              # the exact midpoint of two neighbouring doubles has a finite decimal expansion: round to even
        from fractions import Fraction
        from decimal import Decimal, getcontext
        getcontext().prec = 1200
        lo = float(v)
        hi = float(np.nextafter(lo, np.inf))
        mid = (Fraction(lo) + Fraction(hi)) / 2
        lits.append(format(Decimal(mid.numerator) / Decimal(mid.denominator), "f") if 1e-5 < lo < 1e15 else "%.17g" % lo)
    status = check(lits)
    assert (status == 0).mean() > 0.9


def test_long_literals_and_garbage():
    rng = np.random.default_rng(62)
    lits = []
    for _ in range(3000):                                                   # 20 .. 40 digits: decided when truncation cannot matter, else status 2 — never wrong
        nd = int(rng.integers(20, 41))
        digits = str(rng.integers(1, 10)) + "".join(str(d) for d in rng.integers(0, 10, nd - 1))
        p = int(rng.integers(0, nd + 1))
        lits.append((digits[:p] or "0") + "." + digits[p:] + ("e%d" % rng.integers(-300, 280) if rng.integers(0, 2) else ""))
    status = check(lits)
    assert (status == 0).mean() > 0.5
    vals, st = hc.parse_decimals(["", "abc", "1.2.3", "1e", "1e+", "--1", "1x", "0x10", "nan", "inf", "1,5", "e5", "."])
    assert (st == 1).all()
