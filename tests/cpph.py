"""builds tests/cpp/test_aabb_tree (the C++ host-side class over the C ABI; test infrastructure)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_aabb_tree")


def build():
    from cgal_b200 import _build
    lib = _build.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_aabb_tree.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "mini_cgal.h"), os.path.join(ROOT, "include", "CGAL_B200", "AABB_tree.h"),
            os.path.join(ROOT, "include", "b200aabb.h"), lib]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(p) for p in deps):
        return EXE
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", "-o", EXE, src, "-L" + os.path.dirname(lib), "-lb200aabb",
                           "-Wl,-rpath,$ORIGIN/../../../cgal_b200"])
    return EXE
