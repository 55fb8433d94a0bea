"""Device-side generators of the benchmark workloads + result checksums (include/b200work.h), and their numpy counterparts.

The generators write straight into torch CUDA tensors (no PCIe traffic): config 5 of BASELINE.json generates its 1 B queries this way.
They reproduce cgal_b200.workloads (numpy) bit for bit (tests/test_gpu_workload.py)."""
import ctypes as C

import numpy as np

from . import workloads as W
from .aabb_tree import B200AABBError, load_library

_K1, _K2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03)
_BOUND = False


def _lib():
    global _BOUND
    L = load_library()
    if not _BOUND:
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.b200work_bbox_points.argtypes = [C.c_int, vp, C.POINTER(C.c_double), u32, u64, u64, vp]
        L.b200work_pinhole_rays.argtypes = [C.c_int, vp, C.POINTER(C.c_double), u32, u32, u32, u32, vp]
        L.b200work_checksum128.argtypes = [C.c_int, vp, vp, u64, u64, C.POINTER(u64)]
        L.b200work_dfma_tflops.argtypes = [C.c_int, C.POINTER(C.c_double)]
        _BOUND = True
    return L


def _stream(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def bbox_points(bbox, n, out, seed=W.SEED_POINTS, offset=0):
    """workloads.bbox_points(bbox, n, seed, offset=offset) written into the float64 CUDA tensor `out` (n x 3)"""
    b = (C.c_double * 6)(*[float(x) for x in bbox])
    assert out.is_cuda and out.is_contiguous() and out.numel() >= 3 * n
    rc = _lib().b200work_bbox_points(out.device.index or 0, _stream(out), b, int(seed), int(offset), int(n), C.c_void_p(out.data_ptr()))
    if rc:
        raise B200AABBError(rc)
    return out


def pinhole_camera(nx=4096, ny=2048, eye_height=600.0, vfov_deg=60.0, eye_xy=(0.0, 0.0)):
    """the 13 camera numbers of workloads.pinhole_rays (eye, fwd, right, up, tan(vfov / 2)), computed as numpy computes them"""
    eye = np.array([eye_xy[0], eye_xy[1], eye_height])
    target = np.array([nx / 2.0, ny / 2.0, 0.0])
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    th = np.tan(np.radians(vfov_deg) / 2)
    return np.concatenate([eye, fwd, right, up, [th]])


def pinhole_rays_host(camera, width, height, rows):
    """numpy evaluation of the same rays from the 13 camera numbers (workloads.pinhole_rays with a movable eye)"""
    eye, fwd, right, up, th = camera[0:3], camera[3:6], camera[6:9], camera[9:12], camera[12]
    r0, r1 = rows
    px = (np.arange(width, dtype=np.float64) + 0.5) / width * 2 - 1
    py = 1 - (np.arange(r0, r1, dtype=np.float64) + 0.5) / height * 2
    sx, sy = np.meshgrid(px * th * width / height, py * th)
    d = fwd + sx[..., None] * right + sy[..., None] * up
    out = np.empty(((r1 - r0) * width, 6), dtype=np.float64)
    out[:, :3] = eye
    out[:, 3:] = (eye + d).reshape(-1, 3)
    return out


def pinhole_rays(camera, width, height, row0, nrows, out):
    """rows [row0, row0 + nrows) of the image written into the float64 CUDA tensor `out` (nrows * width x 6)"""
    c = (C.c_double * 13)(*[float(x) for x in camera])
    assert out.is_cuda and out.is_contiguous() and out.numel() >= 6 * nrows * width
    rc = _lib().b200work_pinhole_rays(out.device.index or 0, _stream(out), c, int(width), int(height), int(row0), int(nrows), C.c_void_p(out.data_ptr()))
    if rc:
        raise B200AABBError(rc)
    return out


def checksum128(t, index0=0):
    """(sum0, sum1) of a contiguous CUDA tensor's bytes; shard checksums with consecutive word positions add up modulo 2^64"""
    s = (C.c_uint64 * 2)()
    nbytes = t.numel() * t.element_size()
    rc = _lib().b200work_checksum128(t.device.index or 0, _stream(t), C.c_void_p(t.data_ptr()), nbytes, int(index0), s)
    if rc:
        raise B200AABBError(rc)
    return int(s[0]), int(s[1])


def dfma_tflops(device=0):
    """measured FP64 FMA peak of the device, TFLOP/s"""
    v = C.c_double()
    rc = _lib().b200work_dfma_tflops(int(device), C.byref(v))
    if rc:
        raise B200AABBError(rc)
    return v.value


def _mix64(x):
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def checksum128_host(a, index0=0):
    """the same checksum of a numpy array's bytes"""
    raw = np.ascontiguousarray(a).view(np.uint8).ravel()
    pad = (-len(raw)) % 8
    if pad:
        raw = np.concatenate([raw, np.zeros(pad, np.uint8)])
    w = raw.view(np.uint64)
    with np.errstate(over="ignore"):
        pos = np.arange(len(w), dtype=np.uint64) + np.uint64(index0 + 1)
        s0 = int(np.sum(_mix64(w ^ (_K1 * pos)), dtype=np.uint64))
        s1 = int(np.sum(_mix64(w + _K2 * pos), dtype=np.uint64))
    return s0, s1


def add128(a, b):
    return ((a[0] + b[0]) & 0xFFFFFFFFFFFFFFFF, (a[1] + b[1]) & 0xFFFFFFFFFFFFFFFF)
