"""Synthetic meshes and query sets of the BASELINE.json configs, generated on the host with numpy.

Everything is a pure function of (seed, stream, index) through a counter-based SplitMix64, so the GPU run, the CPU
oracle and the tests see identical inputs.  Seeds: mesh 23, rays 101, points 202, segments 303.
"""
import numpy as np

SEED_MESH, SEED_RAYS, SEED_POINTS, SEED_SEGMENTS = 23, 101, 202, 303
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, stream, index):
    """uint64 hash of (seed, stream, index); index is an array of uint64"""
    with np.errstate(over="ignore"):
        x = (np.asarray(index, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(seed * 1315423911 + stream * 2654435761 + 1)) & _M64
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def uniform(seed, stream, n, offset=0):
    """n doubles in [0,1): u = (x >> 11) * 2^-53"""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    return (splitmix64(seed, stream, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


# ---------------------------------------------------------------------------------------------------- meshes
def soup_of(verts, faces):
    return np.ascontiguousarray(verts[faces.astype(np.int64)].reshape(-1, 9))


def icosphere(level):
    """icosahedron subdivided `level` times, every new vertex normalised in double; level 8 -> 1,310,720 triangles"""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        es = np.sort(e, axis=1)
        key = es[:, 0] * np.int64(len(v)) + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // len(v), uniq % len(v)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = inv.reshape(3, -1) + len(v)      # midpoint vertex of edges (01), (12), (20) per face
        v = np.concatenate([v, mid])
        f = np.concatenate([np.stack([f[:, 0], m[0], m[2]], 1), np.stack([f[:, 1], m[1], m[0]], 1),
                            np.stack([f[:, 2], m[2], m[1]], 1), np.stack([m[0], m[1], m[2]], 1)])
    return np.ascontiguousarray(v), np.ascontiguousarray(f.astype(np.uint32))


def _perlin_table(seed):
    p = np.arange(256, dtype=np.int64)
    keys = splitmix64(seed, 7, np.arange(256, dtype=np.uint64))
    p = p[np.argsort(keys, kind="stable")]
    return np.concatenate([p, p])


def _perlin2(x, y, perm):
    """Perlin improved noise (2002) restricted to z = 0"""
    xi, yi = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    xf, yf = x - xi, y - yi
    xi &= 255
    yi &= 255
    u = xf * xf * xf * (xf * (xf * 6 - 15) + 10)
    v = yf * yf * yf * (yf * (yf * 6 - 15) + 10)

    def grad(h, gx, gy):
        h = h & 15
        a = np.where(h < 8, gx, gy)
        b = np.where(h < 4, gy, np.where((h == 12) | (h == 14), gx, 0.0))
        return np.where(h & 1, -a, a) + np.where(h & 2, -b, b)

    aa = perm[perm[perm[xi] + yi]]
    ab = perm[perm[perm[xi] + yi + 1]]
    ba = perm[perm[perm[xi + 1] + yi]]
    bb = perm[perm[perm[xi + 1] + yi + 1]]
    x1 = grad(aa, xf, yf) + u * (grad(ba, xf - 1, yf) - grad(aa, xf, yf))
    x2 = grad(ab, xf, yf - 1) + u * (grad(bb, xf - 1, yf - 1) - grad(ab, xf, yf - 1))
    return x1 + v * (x2 - x1)


def terrain(nx=4096, ny=2048, amplitude=48.0, octaves=6, base_period=512.0, seed=SEED_MESH):
    """(nx+1) x (ny+1) vertex grid, unit spacing, z = amplitude * fBm(Perlin); every quad split along the same diagonal.
    nx=4096, ny=2048 -> 16,777,216 triangles."""
    perm = _perlin_table(seed)
    xs, ys = np.meshgrid(np.arange(nx + 1, dtype=np.float64), np.arange(ny + 1, dtype=np.float64))
    z = np.zeros_like(xs)
    amp, freq = 1.0, 1.0 / base_period
    for _ in range(octaves):
        z += amp * _perlin2(xs * freq + 0.5, ys * freq + 0.5, perm)
        amp *= 0.5
        freq *= 2.0
    verts = np.stack([xs.ravel(), ys.ravel(), (amplitude * z).ravel()], 1)
    i, j = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64))
    v00 = (j * (nx + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    faces = np.empty((2 * nx * ny, 3), dtype=np.uint32)
    faces[0::2] = np.stack([v00, v10, v11], 1)
    faces[1::2] = np.stack([v00, v11, v01], 1)
    return np.ascontiguousarray(verts), faces


def triangle_soup(n, seed=SEED_MESH, chunk=1 << 22):
    """n triangles: centroid uniform in [0,1]^3, two edge vectors uniform in [-s,s]^3, s = 0.5 n^(-1/3); (n, 9) array"""
    s = 0.5 * n ** (-1.0 / 3.0)
    out = np.empty((n, 9), dtype=np.float64)
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        u = uniform(seed, 1, 9 * m, 9 * b).reshape(m, 9)
        c = u[:, 0:3]
        e1 = (2 * u[:, 3:6] - 1) * s
        e2 = (2 * u[:, 6:9] - 1) * s
        flat = ~np.any(np.cross(e1, e2), axis=1)
        e2[flat, 0] += s      # never an exactly degenerate triangle
        out[b:b + m, 0:3] = c - (e1 + e2) / 3
        out[b:b + m, 3:6] = out[b:b + m, 0:3] + e1
        out[b:b + m, 6:9] = out[b:b + m, 0:3] + e2
    return out


# ---------------------------------------------------------------------------------------------------- queries
def sphere_rays(bbox, n, seed=SEED_RAYS, offset=0):
    """source and second point uniform on the sphere of radius max_extent/2 about the bbox centre
    (AABB_tree/test/AABB_tree/aabb_test_ray_intersection.cpp:87-125); (n, 6)"""
    bbox = np.asarray(bbox, dtype=np.float64)
    c, r = 0.5 * (bbox[:3] + bbox[3:]), 0.5 * (bbox[3:] - bbox[:3]).max()
    u = uniform(seed, 1, 4 * n, 4 * offset).reshape(n, 2, 2)
    z = 2 * u[..., 0] - 1
    phi = 2 * np.pi * u[..., 1]
    s = np.sqrt(np.maximum(0.0, 1 - z * z))
    p = np.stack([s * np.cos(phi), s * np.sin(phi), z], -1)
    return np.ascontiguousarray((c + r * p).reshape(n, 6))


def bbox_points(bbox, n, seed=SEED_POINTS, margin=0.0, offset=0):
    """points uniform in the bbox grown by `margin` x extent (AABB_test_util.h:41-48); (n, 3)"""
    bbox = np.asarray(bbox, dtype=np.float64)
    ext = bbox[3:] - bbox[:3]
    lo, hi = bbox[:3] - margin * ext, bbox[3:] + margin * ext
    return np.ascontiguousarray(lo + uniform(seed, 1, 3 * n, 3 * offset).reshape(n, 3) * (hi - lo))


def bbox_segments(bbox, n, seed=SEED_SEGMENTS, offset=0):
    """two independent uniform points of the bbox per query (AABB_test_util.h:41-48); (n, 6)"""
    bbox = np.asarray(bbox, dtype=np.float64)
    u = uniform(seed, 1, 6 * n, 6 * offset).reshape(n, 2, 3)
    return np.ascontiguousarray((bbox[:3] + u * (bbox[3:] - bbox[:3])).reshape(n, 6))


def grid_points(res=256, lo=-1.5, hi=1.5, z_slices=None):
    """cell centres of a res^3 grid over [lo,hi]^3, x fastest; (res^3, 3) or only the given z slices"""
    h = (hi - lo) / res
    c = lo + (np.arange(res, dtype=np.float64) + 0.5) * h
    zs = c if z_slices is None else c[np.asarray(z_slices)]
    z, y, x = np.meshgrid(zs, c, c, indexing="ij")
    return np.ascontiguousarray(np.stack([x.ravel(), y.ravel(), z.ravel()], 1))


def pinhole_rays(width, height, nx=4096, ny=2048, eye_height=600.0, vfov_deg=60.0, rows=None):
    """primary rays of a pinhole camera above the terrain corner looking at its centre; record = (eye, fl(eye + dir)).
    rows = (r0, r1) restricts to image rows [r0, r1); (n, 6)"""
    eye = np.array([0.0, 0.0, eye_height])
    target = np.array([nx / 2.0, ny / 2.0, 0.0])
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    th = np.tan(np.radians(vfov_deg) / 2)
    r0, r1 = (0, height) if rows is None else rows
    px = (np.arange(width, dtype=np.float64) + 0.5) / width * 2 - 1
    py = 1 - (np.arange(r0, r1, dtype=np.float64) + 0.5) / height * 2
    sx, sy = np.meshgrid(px * th * width / height, py * th)
    d = fwd + sx[..., None] * right + sy[..., None] * up
    out = np.empty(((r1 - r0) * width, 6), dtype=np.float64)
    out[:, :3] = eye
    out[:, 3:] = (eye + d).reshape(-1, 3)
    return out


def shadow_segments(verts, n, seed=SEED_SEGMENTS, sun=(0.8, 0.5, 0.2), length=4096.0, offset=0):
    """segments from terrain surface points (lifted by 1e-3 |z|+1e-3) towards a fixed, low sun (about a quarter of the terrain is in shadow); (n, 6)"""
    sun = np.asarray(sun, dtype=np.float64)
    sun = sun / np.linalg.norm(sun)
    idx = (splitmix64(seed, 2, np.arange(offset, offset + n, dtype=np.uint64)) % np.uint64(len(verts))).astype(np.int64)
    p = verts[idx].copy()
    p[:, 2] += 1e-3 * np.abs(p[:, 2]) + 1e-3
    return np.ascontiguousarray(np.concatenate([p, p + length * sun], 1))


def short_segments(n, seed=SEED_SEGMENTS, length=0.05, offset=0):
    """segments of fixed length with uniform start in [0,1]^3 and uniform direction (config 4); (n, 6)"""
    u = uniform(seed, 3, 5 * n, 5 * offset).reshape(n, 5)
    z = 2 * u[:, 3] - 1
    phi = 2 * np.pi * u[:, 4]
    s = np.sqrt(np.maximum(0.0, 1 - z * z))
    d = np.stack([s * np.cos(phi), s * np.sin(phi), z], 1)
    p = u[:, :3]
    return np.ascontiguousarray(np.concatenate([p, p + length * d], 1))
