#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its 16M-triangle configuration (configs[2]/[4]), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA library (libb200aabb.so) through its C ABI
  python bench.py --impl reference [--steps K] [--warmup W]       the reference's CPU algorithm (the oracle port) on host cores

Workload: Perlin terrain 4096 x 2048 quads = 16,777,216 triangles; the tree is built on rank 0 and
replicated with NCCL through the C ABI (b200aabb_comm_*); every rank then owns a fixed shard of queries ("weak" scaling): image rows
r, r+8, r+16, ... of the 8192 x 8192 pinhole camera (8192 x 1024 = 8,388,608 Ray_3 first_intersection queries) and 8,388,608
closest_point_and_primitive queries uniform in the tree's bbox.  One STEP = one first_intersection batch + one closest_point batch
over the rank's shard.  `value` = M queries/s with the shard resident in HBM; `e2e` = the same steps through the C ABI with pinned HOST
buffers (H2D + kernels + D2H inside the timed region; B200AABB_OPT_ASYNC_HOST lets consecutive batches overlap, one synchronize at the
end; two legs of `steps` steps are timed and the faster one reported, both are in `e2e.legs_ms_per_step`).  Input sets (403 MB rays + 201 MB points per rank) exceed the 126 MB L2, so no explicit flush is needed between steps.
`detail.config5`: BASELINE.json's config 5 — 2^30 mixed queries GENERATED ON THE DEVICE in 2^26-query batches, results reduced to a
128-bit checksum per batch, the batches split over the ranks (a fixed total: the strong-scaling line).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY = 4096, 2048                 # terrain quads -> 16,777,216 triangles
IMG = 8192                          # primary-ray image is IMG x IMG
SHARDS = 8                          # a rank's shard = every 8th image row
RAYS_PER_RANK = IMG * IMG // SHARDS
POINTS_PER_RANK = RAYS_PER_RANK
BYTES_PER_RAY, BYTES_PER_POINT = 85, 61      # algorithmic bytes, SURVEY.md §8(d)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel on this very workload, from the committed `ncu --set full`
# capture of the final state of the round (the file is named in the line; the tree itself — nodes + records — is read about once)
NCU_DRAM = {"closest_point_warp_kernel": (10.899676e9 + 4.182020e9, "profiles/r2_08_closest_point_warp_kernel_final.txt"),
            "first_intersection_kernel": (1.113175e9 + 0.945382e9, "profiles/r2_08_first_intersection_kernel_final.txt")}
WORKLOAD = "terrain16M(4096x2048 quads, Perlin fBm seed 23) + per GPU 8,388,608 pinhole Ray_3 first_intersection (rows r::8 of 8192^2) + 8,388,608 bbox closest_point_and_primitive"
C5_BATCH, C5_BATCHES = 1 << 26, 8            # config 5: 8 ray batches (one 8192^2 image each) + 8 point batches of 2^26 = 2^30 queries


def shard_rays(rank):
    from cgal_b200 import workloads as W
    out = np.empty((RAYS_PER_RANK, 6), dtype=np.float64)
    rows = np.arange(rank % SHARDS, IMG, SHARDS)
    for k in range(0, len(rows), 64):      # generate in strips of 64 rows to bound temporaries
        for j, r in enumerate(rows[k:k + 64]):
            out[(k + j) * IMG:(k + j + 1) * IMG] = W.pinhole_rays(IMG, IMG, NX, NY, rows=(int(r), int(r) + 1))
    return out


def shard_points(bbox, rank, n=POINTS_PER_RANK):
    from cgal_b200 import workloads as W
    return W.bbox_points(bbox, n, offset=rank * POINTS_PER_RANK)


class ClockSampler:
    """nvidia-smi samples taken DURING the timed region (B200_PROFILING.md)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------- reference arm
def oracle_tree(verts, faces):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    o = orc.OracleTree(verts, faces)       # median-split build + kd-tree hint (accelerate_distance_queries)
    return o


def cpu_rates(o, rays, pts, threads, rays_1t=1 << 17, pts_1t=1 << 14):
    """the oracle on `threads` host threads and on ONE thread (north star: both), plus its traversal work per query"""
    import orc
    orc.traversal_totals()
    t0 = time.perf_counter()
    rr = o.first_intersection(rays, threads=threads)
    t1 = time.perf_counter()
    rb, rp = orc.traversal_totals()
    rc = o.closest_point(pts, threads=threads)
    t2 = time.perf_counter()
    pb, pp = orc.traversal_totals()
    r1 = np.ascontiguousarray(rays[::max(1, len(rays) // rays_1t)][:rays_1t])     # (strided: the same mix of hits and misses as the whole sample)
    p1 = pts[:pts_1t]
    s0 = time.perf_counter()
    o.first_intersection(r1, threads=1)
    s1 = time.perf_counter()
    o.closest_point(p1, threads=1)
    s2 = time.perf_counter()
    orc.traversal_totals()
    rates = {"first_intersection_Mrays_s": len(rays) / (t1 - t0) / 1e6, "closest_point_Mqueries_s": len(pts) / (t2 - t1) / 1e6,
             "single_thread": {"first_intersection_Mrays_s": len(r1) / (s1 - s0) / 1e6, "closest_point_Mqueries_s": len(p1) / (s2 - s1) / 1e6,
                               "sample": "%d rays (strided) + %d points" % (len(r1), len(p1))},
             "traversal": {"ray_box_tests_per_query": rb / len(rays), "ray_primitives_per_query": rp / len(rays),
                           "point_box_tests_per_query": pb / len(pts), "point_primitives_per_query": pp / len(pts)}}
    return rates, rr, rc


def run_reference(args):
    """the reference's own CPU algorithm for the path (oracle port: CGAL itself needs Boost, absent here), all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cgal_b200 import workloads as W
    threads = os.cpu_count() or 1
    verts, faces = W.terrain(NX, NY)
    o = oracle_tree(verts, faces)
    n_rays, n_pts = 1 << 20, 1 << 17
    rays_all = shard_rays(0)
    pts_all = W.bbox_points(o.bbox(), n_pts * (args.steps + args.warmup))
    stride = RAYS_PER_RANK // n_rays      # every stride-th ray of the shard: same distribution as the full shard

    def step(i):
        rays = np.ascontiguousarray(rays_all[(i % stride)::stride][:n_rays])
        pts = pts_all[i * n_pts:(i + 1) * n_pts]
        t0 = time.perf_counter()
        o.first_intersection(rays, threads=threads)
        t1 = time.perf_counter()
        o.closest_point(pts, threads=threads)
        return t1 - t0, time.perf_counter() - t1

    for i in range(args.warmup):
        step(i)
    tr = tp = 0.0
    for i in range(args.steps):
        a, b = step(args.warmup + i)
        tr += a
        tp += b
    rays_ps, pts_ps = n_rays * args.steps / tr, n_pts * args.steps / tp
    # the GPU step is 50 % rays / 50 % points: harmonic combination gives the mixed rate of the same step on the CPU
    mixed = 2.0 / (1.0 / rays_ps + 1.0 / pts_ps) / 1e6
    one, _, _ = cpu_rates(o, np.ascontiguousarray(rays_all[::stride][:1 << 17]), pts_all[:1 << 14], 1)
    sample = "%d steps x (%d rays + %d points) against the full 16,777,216-triangle tree; mixed rate for the 50/50 step" % (args.steps, n_rays, n_pts)
    line = {"impl": "reference", "metric": "Mqueries/s (50% first_intersection Ray_3 + 50% closest_point_and_primitive)", "value": mixed, "unit": "Mqueries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tr + tp) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "reference_sample": sample},
            "detail": {"first_intersection_Mrays_s": rays_ps / 1e6, "closest_point_Mqueries_s": pts_ps / 1e6, "build_Mtri_s": len(faces) / o.build_seconds() / 1e6,
                       "build_s": o.build_seconds(), "kd_hint_build_s": o.kd_seconds(),
                       "single_thread": {"first_intersection_Mrays_s": one["first_intersection_Mrays_s"], "closest_point_Mqueries_s": one["closest_point_Mqueries_s"]},
                       "traversal": one["traversal"]},
            "cpu_baseline": {"value": mixed, "unit": "Mqueries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mixed, "unit": "Mqueries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def replicate_tree(tree, rank, world, local, dist, torch, B):
    """rank 0 built the tree; the C ABI replicates it with NCCL (b200aabb_comm_broadcast_tree).  Both variants SURVEY.md §8e names are
    timed: the built tree (64-byte nodes + 96-byte records; binary32 copies derived by the receiver) and the primitives + a build on every rank."""
    uid = [B.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = B.Comm(local, rank, world, uid[0])        # (warms the communicator up: the first collective sets up NCCL's channels)
    out = {}
    for mode, name in ((1, "primitives_then_build_on_every_rank"), (0, "built_tree")):   # (the replica kept is the built tree)
        times = []
        for _ in range(3):
            dist.barrier()
            ms, nbytes = comm.broadcast_tree(tree, root=0, mode=mode)
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        out[name] = {"bytes": nbytes, "ms_max_over_ranks": min(times), "ms_all": times, "GB_s": nbytes / min(times) / 1e6}
    return comm, out


def pcie_ceiling(torch, h2d_bytes, d2h_bytes, reps=3):
    """bare pinned cudaMemcpyAsync of one step's bytes, H2D and D2H concurrently on two streams: what the host link allows"""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    return best


def other_queries(tree, torch, B, pts_d):
    """NOT part of `value`: the other batched queries of the same tree, timed once each for the line's `detail` (SURVEY.md §8f rows built in round 2):
    Triangle_3 id queries (2^22 half-unit triangles at the benchmark's points) and tree-vs-tree do_intersect of the terrain against itself under a translation."""
    import time

    def timed(fn, reps=3):
        best, out = None, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None else min(best, ms)
        return best, out
    n = min(1 << 22, pts_d.shape[0])
    p = pts_d[:n]
    e1 = torch.tensor([0.5, 0.0, 0.25], dtype=torch.float64, device="cuda")
    e2 = torch.tensor([0.0, 0.5, 0.25], dtype=torch.float64, device="cuda")
    tri = torch.cat([p, p + e1, p + e2], dim=1).contiguous()
    ms_any, hit = timed(lambda: tree.do_intersect(tri, B.TRIANGLE))
    ms_cnt, cnt = timed(lambda: tree.number_of_intersected_primitives(tri, B.TRIANGLE))
    shift = lambda x, y, z: [[1.0, 0, 0, x], [0, 1.0, 0, y], [0, 0, 1.0, z]]      # noqa: E731
    tt = {}
    for name, m in (("moved_by_(0.5, 0.5, 0)", shift(0.5, 0.5, 0.0)), ("lifted_by_200", shift(0.0, 0.0, 200.0))):
        ms, r = timed(lambda: tree.do_intersect_tree(tree, other_xf=np.array(m)))
        tt[name] = {"ms": ms, "intersect": bool(r)}
    return {"triangle_queries": n, "triangle_do_intersect_Mq_s": n / ms_any / 1e3, "triangle_count_Mq_s": n / ms_cnt / 1e3, "triangle_hit_fraction": float((hit != 0).float().mean()),
            "triangle_count_total": int(cnt.sum()), "terrain_against_itself_do_intersect_tree": tt, "timing": "wall clock around the call, best of 3, not in the timed region of `value`"}


def config5(tree, rank, world, local, dist, torch, B, bbox, oracle_check=None):
    """BASELINE.json config 5: 2^30 mixed queries generated on the device (b200work_*) in 2^26 batches, batches split over the ranks;
    128-bit checksum of every result array; the first 2^20 rays / points of batch 0 are kept for parity with the oracle (rank 0)."""
    from cgal_b200 import device_workloads as D
    per = C5_BATCH // world                     # queries of a batch this rank answers (a block of image rows / of point indices)
    rows = IMG // world
    rays = torch.empty((per, 6), dtype=torch.float64, device="cuda")
    pts = torch.empty((per, 3), dtype=torch.float64, device="cuda")
    o_r = {"prim": torch.empty(per, dtype=torch.int32, device="cuda"), "t": torch.empty(per, dtype=torch.float64, device="cuda"),
           "point": torch.empty((per, 3), dtype=torch.float64, device="cuda"), "flags": torch.empty(per, dtype=torch.uint8, device="cuda")}
    o_p = {"point": torch.empty((per, 3), dtype=torch.float64, device="cuda"), "prim": torch.empty(per, dtype=torch.int32, device="cuda"),
           "sqdist": torch.empty(per, dtype=torch.float64, device="cuda"), "flags": torch.empty(per, dtype=torch.uint8, device="cuda")}
    want_r, want_p = ("prim", "t", "point", "flags"), ("point", "prim", "sqdist", "flags")
    cams = [D.pinhole_camera(NX, NY, eye_xy=(256.0 * k, 128.0 * k)) for k in range(C5_BATCHES)]
    sums = {}
    kept = {}

    def add(name, t, index0):
        sums[name] = D.add128(sums.get(name, (0, 0)), D.checksum128(t, index0=index0))

    def ev():
        return torch.cuda.Event(enable_timing=True)
    q_ms = {"rays": 0.0, "points": 0.0}
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(C5_BATCHES):
        a, b, c = ev(), ev(), ev()
        D.pinhole_rays(cams[k], IMG, IMG, rank * rows, rows, rays)
        a.record()
        tree.first_intersection(rays, want=want_r, out=o_r)
        b.record()
        for name in want_r:                       # word positions are global: the shards' checksums add up to the batch's
            words_per_q = {"prim": 0.5, "t": 1, "point": 3, "flags": 0.125}[name]
            add("ray_" + name, o_r[name], int((k * C5_BATCH + rank * per) * words_per_q))
        if k == 0 and rank == 0:
            kept["rays"] = rays[:1 << 20].cpu().numpy()
            kept["ray_out"] = {n: o_r[n][:1 << 20].cpu().numpy() for n in want_r}
        D.bbox_points(bbox, per, pts, offset=k * C5_BATCH + rank * per)
        c.record()
        d = ev()
        tree.closest_point_and_primitive(pts, want=want_p, out=o_p)
        d.record()
        for name in want_p:
            words_per_q = {"prim": 0.5, "sqdist": 1, "point": 3, "flags": 0.125}[name]
            add("point_" + name, o_p[name], int((k * C5_BATCH + rank * per) * words_per_q))
        if k == 0 and rank == 0:
            kept["pts"] = pts[:1 << 20].cpu().numpy()
            kept["pt_out"] = {n: o_p[n][:1 << 20].cpu().numpy() for n in want_p}
        torch.cuda.synchronize()
        q_ms["rays"] += a.elapsed_time(b)
        q_ms["points"] += c.elapsed_time(d)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    names = sorted(sums)
    vec = torch.tensor([[sums[n][0] - (1 << 64) if sums[n][0] >= (1 << 63) else sums[n][0], sums[n][1] - (1 << 64) if sums[n][1] >= (1 << 63) else sums[n][1]] for n in names],
                       dtype=torch.int64, device="cuda")
    tm = torch.tensor([total_ms, q_ms["rays"], q_ms["points"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vec)                       # (int64 addition wraps: sums modulo 2^64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, ray_ms, pt_ms = [float(x) for x in tm.tolist()]
    digest = "".join("%016x%016x" % (int(a) & 0xFFFFFFFFFFFFFFFF, int(b) & 0xFFFFFFFFFFFFFFFF) for a, b in vec.tolist())
    import hashlib
    nq = 2 * C5_BATCHES * C5_BATCH
    res = {"queries": nq, "batches": "%d x 2^26 rays (one 8192^2 image each, eye moved per batch) + %d x 2^26 bbox points, generated on the device" % (C5_BATCHES, C5_BATCHES),
           "ms_total_incl_generation_and_checksums": total_ms, "Mqueries_s_total": nq / total_ms / 1e3,
           "first_intersection_Mrays_s": C5_BATCHES * C5_BATCH / ray_ms / 1e3, "closest_point_Mqueries_s": C5_BATCHES * C5_BATCH / pt_ms / 1e3,
           "Mqueries_s_queries_only": nq / (ray_ms + pt_ms) / 1e3, "scaling": "strong (2^30 queries split over %d GPU(s))" % world,
           "checksum128_per_array": {n: "%016x%016x" % (int(a) & 0xFFFFFFFFFFFFFFFF, int(b) & 0xFFFFFFFFFFFFFFFF) for n, (a, b) in zip(names, vec.tolist())},
           "checksum_of_checksums": hashlib.sha256(digest.encode()).hexdigest()[:32]}
    return res, kept


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 2^30-query config 5 leg (profiling runs)")
    ap.add_argument("--small", action="store_true", help="1M-triangle terrain, for quick profiling only (not a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" and not args.small else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import cgal_b200 as B
    from cgal_b200 import device_workloads as D
    from cgal_b200 import workloads as W
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA library is the product and there is no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny = (1024, 512) if args.small else (NX, NY)

    tree = B.AABB_tree(device=local)
    tree.set_stream(torch.cuda.current_stream().cuda_stream)
    n_prims = 2 * nx * ny
    build_ms, build_wall_ms, bcast = None, None, None
    verts = faces = None
    if rank == 0:
        verts, faces = W.terrain(nx, ny)
        tree.insert(verts, faces)
        tree.build()                       # warm-up build (allocations, module load)
        times, walls = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            tree.build()
            walls.append((time.perf_counter() - t0) * 1e3)   # host wall clock of the call: the build's buffers are kept by the handle, so a re-build allocates nothing
            times.append(tree.stats()["last_build_ms"])
        build_ms, build_wall_ms = min(times), min(walls)
    comm = None
    if world > 1:
        comm, bcast = replicate_tree(tree, rank, world, local, dist, torch, B)
    bbox = tree.bbox()

    # the rank's shard, generated on the device (bit-equal to cgal_b200.workloads, tests/test_gpu_workload.py)
    img = IMG // 4 if args.small else IMG
    n_rays = img * img // SHARDS
    n_pts = n_rays
    rays_d = torch.empty((n_rays, 6), dtype=torch.float64, device="cuda")
    cam = D.pinhole_camera(nx, ny)
    for j in range(img // SHARDS):          # rows rank, rank + 8, ...
        D.pinhole_rays(cam, img, img, (rank % SHARDS) + j * SHARDS, 1, rays_d[j * img:(j + 1) * img])
    pts_d = torch.empty((n_pts, 3), dtype=torch.float64, device="cuda")
    D.bbox_points(bbox, n_pts, pts_d, offset=rank * POINTS_PER_RANK)
    out_r = {"prim": torch.empty(n_rays, dtype=torch.int32, device="cuda"), "t": torch.empty(n_rays, dtype=torch.float64, device="cuda"),
             "point": torch.empty((n_rays, 3), dtype=torch.float64, device="cuda"), "flags": torch.empty(n_rays, dtype=torch.uint8, device="cuda")}
    out_p = {"point": torch.empty((n_pts, 3), dtype=torch.float64, device="cuda"), "prim": torch.empty(n_pts, dtype=torch.int32, device="cuda"),
             "sqdist": torch.empty(n_pts, dtype=torch.float64, device="cuda"), "flags": torch.empty(n_pts, dtype=torch.uint8, device="cuda")}
    want_r, want_p = ("prim", "t", "point", "flags"), ("point", "prim", "sqdist", "flags")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def device_step(marks=None):
        if marks is not None:
            marks[0].record()
        tree.first_intersection(rays_d, want=want_r, out=out_r)
        if marks is not None:
            marks[1].record()
        tree.closest_point_and_primitive(pts_d, want=want_p, out=out_p)
        if marks is not None:
            marks[2].record()

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = tree.stats()["kernel_launches"]
    marks = [[ev(), ev(), ev()] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for i in range(args.steps):
        device_step(marks[i])
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    stats = tree.stats()
    launches = stats["kernel_launches"] - l0
    ray_ms = sum(m[0].elapsed_time(m[1]) for m in marks) / args.steps
    pt_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / args.steps
    hit_rate = float((out_r["prim"] != -1).float().mean().item())
    # traversal-traffic model: nodes loaded / primitives tested per query, from the kernels' own counters
    tree.first_intersection(rays_d, want=want_r, out=out_r)
    sr = tree.stats()
    tree.closest_point_and_primitive(pts_d, want=want_p, out=out_p)
    sp = tree.stats()
    ties, exact, near = sr["ties"] + sp["ties"], sr["exact_fallbacks"] + sp["exact_fallbacks"], sr["near_ties"] + sp["near_ties"]
    traffic_model = {"ray_nodes_per_query": sr["nodes_visited"] / n_rays, "ray_leaves_per_query": sr["leaves_tested"] / n_rays,
                     "point_nodes_per_query": sp["nodes_visited"] / n_pts, "point_leaves_per_query": sp["leaves_tested"] / n_pts, "point_leaves_culled_fp32_per_query": sp["leaves_culled"] / n_pts,
                     "ray_ms_incl_sort": sr["last_query_ms"], "point_ms_incl_sort": sp["last_query_ms"], "node_bytes": 64, "leaf_bytes": 96}
    # a camera that fills the frame (every ray hits): looking almost straight down on the terrain's centre
    allhit = None
    if not args.small:
        cam2 = D.pinhole_camera(nx, ny, eye_height=1400.0, eye_xy=(nx / 2.0, ny / 2.0 - 1.0))
        rays2 = torch.empty_like(rays_d)
        for j in range(img // SHARDS):
            D.pinhole_rays(cam2, img, img, (rank % SHARDS) + j * SHARDS, 1, rays2[j * img:(j + 1) * img])
        prim2 = torch.empty(n_rays, dtype=torch.int32, device="cuda")
        tree.first_intersection(rays2, want=("prim",), out={"prim": prim2})
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            tree.first_intersection(rays2, want=want_r, out=out_r)
        b.record()
        torch.cuda.synchronize()
        allhit = {"Mrays_s": world * n_rays * 3 / a.elapsed_time(b) / 1e3, "hit_rate": float((prim2 != -1).float().mean().item()),
                  "camera": "eye 1400 above the terrain centre looking down, 60 deg, rows r::8 of 8192^2"}
        tree.first_intersection(rays_d, want=want_r, out=out_r)    # (restore the benchmark batch's results for the parity sample)
        del rays2, prim2

    # e2e: the same steps through the C ABI with pinned HOST buffers (copies inside the timed region)
    e2e_ms, ceiling_ms, e2e_legs = None, None, []
    h2d = n_rays * 48 + n_pts * 24
    d2h = n_rays * (4 + 8 + 24 + 1) + n_pts * (24 + 4 + 8 + 1)
    if not args.no_e2e:
        def pinned(shape, dtype):
            return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
        rays_p, pts_p = pinned((n_rays, 6), torch.float64), pinned((n_pts, 3), torch.float64)
        rays_p[:], pts_p[:] = rays_d.cpu().numpy(), pts_d.cpu().numpy()
        hout_r = {"prim": pinned(n_rays, torch.int32).view(np.uint32), "t": pinned(n_rays, torch.float64), "point": pinned((n_rays, 3), torch.float64), "flags": pinned(n_rays, torch.uint8)}
        hout_p = {"point": pinned((n_pts, 3), torch.float64), "prim": pinned(n_pts, torch.int32).view(np.uint32), "sqdist": pinned(n_pts, torch.float64), "flags": pinned(n_pts, torch.uint8)}

        def host_step():
            tree.first_intersection(rays_p, want=want_r, out=hout_r)
            tree.closest_point_and_primitive(pts_p, want=want_p, out=hout_p)
        tree.set_option(B.OPT_ASYNC_HOST, 1)     # batches return once enqueued; one synchronize after the last step
        for _ in range(max(1, args.warmup)):
            host_step()
        tree.synchronize()
        e2e_legs = []
        for _ in range(2):                       # two timed legs of `steps` steps each, the faster one is reported (both are in the line): a leg shares the
            barrier()                            # host's PCIe and DRAM with whatever else runs on the box (one leg in ten came out 25 % slow)
            f0, f1 = ev(), ev()
            f0.record()
            for _ in range(args.steps):
                host_step()
            tree.synchronize()
            f1.record()
            barrier()
            e2e_legs.append(f0.elapsed_time(f1))
        e2e_ms = min(e2e_legs)
        tree.set_option(B.OPT_ASYNC_HOST, 0)
        # the HOST-pointer path must return what the DEVICE-pointer path returned
        assert np.array_equal(hout_r["prim"], out_r["prim"].cpu().numpy().view(np.uint32)) and np.array_equal(hout_p["sqdist"], out_p["sqdist"].cpu().numpy())
        barrier()
        ceiling_ms = pcie_ceiling(torch, h2d, d2h)

    c5, kept = (None, {})
    if not args.no_config5 and not args.small:
        c5, kept = config5(tree, rank, world, local, dist, torch, B, bbox)
    extras = None
    if rank == 0 and world == 1 and not args.small:
        try:
            extras = other_queries(tree, torch, B, pts_d)
        except Exception as exc:          # (informational leg: never takes the line down)
            extras = {"error": repr(exc)}

    if dist is not None:
        legs = e2e_legs if e2e_ms is not None else [0.0, 0.0]
        tm = torch.tensor([total_ms, legs[0], ray_ms, pt_ms, ceiling_ms or 0.0, legs[1]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms, leg0, ray_ms, pt_ms, ceil_max, leg1 = [float(x) for x in tm.tolist()]
        if e2e_ms is not None:
            e2e_legs = [leg0, leg1]              # each leg: the slowest rank; the faster leg is reported
            e2e_ms = min(e2e_legs)
        ceiling_ms = ceil_max if ceiling_ms is not None else None
        cnt = torch.tensor([launches, ties, exact, near], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        launches, ties, exact, near = [int(x) for x in cnt.tolist()]

    if rank == 0:
        q_per_step = world * (n_rays + n_pts)
        value = q_per_step * args.steps / (total_ms * 1e-3) / 1e6
        peak, peak_src = measured_peak_gbs()
        # dominant kernel = whichever of the two traversal kernels takes longer per step
        if ray_ms >= pt_ms:
            dom, dom_bytes, dom_ms = "first_intersection_kernel", n_rays * BYTES_PER_RAY, ray_ms
        else:
            dom, dom_bytes, dom_ms = "closest_point_warp_kernel", n_pts * BYTES_PER_POINT, pt_ms
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic, traffic_src = NCU_DRAM.get(dom, (None, None))
        try:
            dfma = D.dfma_tflops(local)
        except Exception:
            dfma = None
        line = {"metric": "Mqueries/s (50% first_intersection Ray_3 + 50% closest_point_and_primitive)", "value": value, "unit": "Mqueries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD if not args.small else "SMALL profiling variant (1M-triangle terrain): not a bench value", "triangles": n_prims,
                           "rays_per_gpu": n_rays, "points_per_gpu": n_pts, "l2": "inputs (%.0f MB per GPU) larger than L2" % ((n_rays * 48 + n_pts * 24) / 1e6),
                           "parallelism": "tree replicated by NCCL through the C ABI (b200aabb_comm_broadcast_tree), queries sharded x%d" % world},
                "detail": {"first_intersection_Mrays_s": world * n_rays / ray_ms / 1e3, "first_intersection_hits_Mrays_s": hit_rate * world * n_rays / ray_ms / 1e3,
                           "first_intersection_all_hit_camera": allhit, "closest_point_Mqueries_s": world * n_pts / pt_ms / 1e3,
                           "build_ms": build_ms, "build_call_wall_ms": build_wall_ms, "build_Mtri_s": n_prims / build_ms / 1e3, "build_GB_s_algorithmic(492B/tri)": 492.0 * n_prims / build_ms / 1e6,
                           "build_frac_of_hbm_roofline(492B/tri)": 492.0 * n_prims / build_ms / 1e6 / peak,
                           "ray_hit_rate_rank0": hit_rate, "ties": ties, "near_ties": near, "exact_stage_queries": exact, "traversal_rank0": traffic_model,
                           "fp64_dfma_peak_TFLOP_s": dfma, "tree_broadcast": bcast, "config5": c5, "other_queries": extras},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic if not args.small and n_pts == POINTS_PER_RANK else None, "traffic_source": traffic_src,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                             "note": "traversal is bound by the L1 data stage (one wavefront per scattered 32-byte sector of a node / triangle fetch) and by instruction issue at 20 warps/SM, "
                                     "not by HBM (DESIGN.md §5); traffic = dram__bytes_read+write of one launch from the ncu --set full capture named in traffic_source"},
                "clocks": clocks, "gpu_launches": launches}
        if e2e_ms is not None:
            e2e_v = q_per_step * args.steps / (e2e_ms * 1e-3) / 1e6
            line["e2e"] = {"value": e2e_v, "unit": "Mqueries/s", "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h,
                           "frac_of_value": e2e_v / value, "mode": "B200AABB_OPT_ASYNC_HOST: %d steps enqueued, one b200aabb_synchronize" % args.steps,
                           "legs_ms_per_step": [x / args.steps for x in e2e_legs], "bare_memcpy_ms_per_step": ceiling_ms, "frac_of_bare_memcpy_ceiling": (ceiling_ms / (e2e_ms / args.steps)) if ceiling_ms else None}
        if world == 1 and not args.no_cpu_baseline:
            o = oracle_tree(verts, faces)
            threads = os.cpu_count() or 1
            nr, npnt = 1 << 20, 1 << 17
            stride = max(1, n_rays // nr)
            rays_h = rays_d.cpu().numpy()
            cr = np.ascontiguousarray(rays_h[::stride][:nr])
            cp = pts_d[:npnt].cpu().numpy()
            rates, rr, rc = cpu_rates(o, cr, cp, threads)
            rps, pps = rates["first_intersection_Mrays_s"] * 1e6, rates["closest_point_Mqueries_s"] * 1e6
            # the sample doubles as a parity check of the benchmarked batch against the oracle
            g_prim = out_r["prim"].cpu().numpy().view(np.uint32)[::stride][:nr]
            g_t = out_r["t"].cpu().numpy()[::stride][:nr]
            g_fl = out_r["flags"].cpu().numpy()[::stride][:nr]
            g_tie = (g_fl & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
            p_prim = out_p["prim"].cpu().numpy().view(np.uint32)[:npnt]
            p_tie = (out_p["flags"].cpu().numpy()[:npnt] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
            parity = {"rays_checked": len(cr), "ray_t_bit_equal_outside_near_ties": bool(np.array_equal(g_t[~g_tie], rr["t"][~g_tie])),
                      "ray_t_max_rel_diff": float(np.max(np.abs(g_t - rr["t"]) / np.maximum(np.abs(rr["t"]), 1e-300))), "ray_id_mismatch_not_tie": int(((g_prim != rr["prim"]) & ~g_tie).sum()),
                      "ray_near_ties_in_sample": int(((g_fl & B.FLAG_NEAR_TIE) != 0).sum()),
                      "points_checked": len(cp), "sqdist_bit_equal": bool(np.array_equal(out_p["sqdist"].cpu().numpy()[:npnt], rc["sqd"])),
                      "point_id_mismatch_not_tie": int(((p_prim != rc["prim"]) & ~p_tie).sum())}
            if kept:      # config 5: the first 2^20 rays / 2^17 points of batch 0 against the oracle
                k5r = o.first_intersection(kept["rays"], threads=threads)
                k5p = o.closest_point(kept["pts"][:npnt], threads=threads)
                fl = kept["ray_out"]["flags"]
                tie5 = (fl & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
                ptie5 = (kept["pt_out"]["flags"][:npnt] & B.FLAG_TIE) != 0
                parity["config5_first_batch"] = {"rays_checked": len(kept["rays"]), "ray_t_bit_equal_outside_near_ties": bool(np.array_equal(kept["ray_out"]["t"][~tie5], k5r["t"][~tie5])),
                                                 "ray_id_mismatch_not_tie": int(((kept["ray_out"]["prim"].view(np.uint32) != k5r["prim"]) & ~tie5).sum()),
                                                 "points_checked": npnt, "sqdist_bit_equal": bool(np.array_equal(kept["pt_out"]["sqdist"][:npnt], k5p["sqd"])),
                                                 "point_id_mismatch_not_tie": int(((kept["pt_out"]["prim"][:npnt].view(np.uint32) != k5p["prim"]) & ~ptie5).sum())}
            line["cpu_baseline"] = {"value": 2.0 / (1.0 / rps + 1.0 / pps) / 1e6, "unit": "Mqueries/s", "cores": threads, "kind": "port",
                                    "sample": "%d rays (every %dth of the shard) + %d points against the full tree; 50/50 mixed rate" % (len(cr), stride, len(cp)),
                                    "first_intersection_Mrays_s": rps / 1e6, "closest_point_Mqueries_s": pps / 1e6, "single_thread": rates["single_thread"],
                                    "oracle_traversal": rates["traversal"], "build_s": o.build_seconds(),
                                    "build_Mtri_s": n_prims / o.build_seconds() / 1e6, "kd_hint_build_s": o.kd_seconds(), "parity_on_sample": parity}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        del comm
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
