"""The C-ABI library loads and exports every symbol include/b200aabb.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200aabb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200aabb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import cgal_b200
    from cgal_b200 import _build
    _build.build()
    lib = cgal_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libb200aabb.so does not export %s" % s
    assert set(syms) == set(cgal_b200.aabb_tree.SYMBOLS)
    assert b"sm_100a" in lib.b200aabb_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cgal_b200
    with pytest.raises(cgal_b200.B200AABBError) as e:
        cgal_b200.AABB_tree()
    assert e.value.code == -5


def test_product_does_not_reference_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "cgal_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle/" not in txt and "liboracle" not in txt and "import orc" not in txt, f
    so = open(os.path.join(ROOT, "cgal_b200", "libb200aabb.so"), "rb").read()
    assert b"liboracle" not in so and b"orc_" not in so


def test_every_bound_symbol_declares_its_argument_types():
    """ctypes passes an undeclared Python int as a 32-bit C int: a pointer handed to such a function is truncated (found on the GPU box with
    b200aabb_side_of_mesh); every symbol the mirror binds must declare argtypes"""
    import cgal_b200
    from cgal_b200 import aabb_tree
    L = cgal_b200.load_library()
    missing = [s for s in aabb_tree.SYMBOLS if getattr(L, s).argtypes is None and s not in ("b200aabb_version",)]
    assert not missing, missing
