#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its 16M-triangle configuration (configs[2]/[4]), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA library (libb200aabb.so) through its C ABI
  python bench.py --impl reference [--steps K] [--warmup W]       the reference's CPU algorithm (the oracle port) on host cores

Workload: Perlin terrain 4096 x 2048 quads = 16,777,216 triangles; the tree is built on
rank 0 and broadcast over NCCL; every rank then owns a fixed shard of queries ("weak" scaling): image rows
r, r+8, r+16, ... of the 8192 x 8192 pinhole camera (8192 x 1024 = 8,388,608 Ray_3 first_intersection queries) and
8,388,608 closest_point_and_primitive queries uniform in the tree's bbox.  One STEP = one first_intersection batch +
one closest_point batch over the rank's shard.  `value` = M queries/s with the shard resident in HBM; `e2e` = the same
step through the C ABI with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
Input sets (403 MB rays + 201 MB points per rank) exceed the 126 MB L2, so no explicit flush is needed between steps.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel on this very workload, from the committed ncu --set full
# captures (profiles/r1_15_warp_pool_final.txt, profiles/r1_03_while_while.txt): the tree itself (nodes + records) is read about once
NCU_DRAM_BYTES_PER_LAUNCH = {"closest_point_warp_kernel": 10.892577e9 + 4.193736e9, "first_intersection_kernel": 1.074644e9 + 0.303553e9}
WORKLOAD = "terrain16M(4096x2048 quads, Perlin fBm seed 23) + per GPU 8,388,608 pinhole Ray_3 first_intersection (rows r::8 of 8192^2) + 8,388,608 bbox closest_point_and_primitive"


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
    t0 = time.perf_counter()
    o = orc.OracleTree(verts, faces)       # median-split build + kd-tree hint (accelerate_distance_queries)
    return o, time.perf_counter() - t0


def cpu_sample(o, n_rays, n_pts, threads, step=0):
    """one bounded step of the CPU path: n_rays first_intersection + n_pts closest_point; returns seconds (rays, points)"""
    rays = shard_rays(0)[step * n_rays % RAYS_PER_RANK:][:n_rays] if n_rays else None
    t_r = t_p = 0.0
    if n_rays:
        t0 = time.perf_counter()
        o.first_intersection(rays, threads=threads)
        t_r = time.perf_counter() - t0
    if n_pts:
        from cgal_b200 import workloads as W
        pts = W.bbox_points(o.bbox(), n_pts, offset=step * n_pts)
        t0 = time.perf_counter()
        o.closest_point(pts, threads=threads)
        t_p = time.perf_counter() - t0
    return t_r, t_p


def run_reference(args):
    """the reference's own CPU algorithm for the path (oracle port: CGAL itself needs Boost, absent here), all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cgal_b200 import workloads as W
    threads = os.cpu_count() or 1
    verts, faces = W.terrain(NX, NY)
    o, build_s = oracle_tree(verts, faces)
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
    sample = "%d steps x (%d rays + %d points) against the full 16,777,216-triangle tree; mixed rate for the 50/50 step" % (args.steps, n_rays, n_pts)
    line = {"impl": "reference", "metric": "Mqueries/s (50% first_intersection Ray_3 + 50% closest_point_and_primitive)", "value": mixed, "unit": "Mqueries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (tr + tp) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "reference_sample": sample},
            "detail": {"first_intersection_Mrays_s": rays_ps / 1e6, "closest_point_Mqueries_s": pts_ps / 1e6, "build_Mtri_s": len(faces) / o.build_seconds() / 1e6,
                       "build_s": o.build_seconds(), "kd_hint_build_s": o.kd_seconds()},
            "cpu_baseline": {"value": mixed, "unit": "Mqueries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mixed, "unit": "Mqueries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- our arm
def broadcast_tree(tree, rank, world, dist, torch, n_prims):
    """rank 0 built the tree; the others allocate same-sized buffers and receive them with NCCL (NVLink/NVSwitch)"""
    from cgal_b200 import sharding
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    _, total = sharding.broadcast_tree(tree, n_prims, src=0)
    e1.record()
    torch.cuda.synchronize()
    return total, e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--small", action="store_true", help="1M-triangle terrain, for quick profiling only (not a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" and not args.small else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import cgal_b200 as B
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
    launches0 = tree.stats()["kernel_launches"]
    n_prims = 2 * nx * ny
    build_ms, bcast = None, None
    verts = faces = None
    if rank == 0:
        verts, faces = W.terrain(nx, ny)
        tree.insert(verts, faces)
        tree.build()                       # warm-up build (allocations, module load)
        times = []
        for _ in range(3):
            tree.build()
            times.append(tree.stats()["last_build_ms"])
        build_ms = min(times)
    if world > 1:
        nbytes, ms = broadcast_tree(tree, rank, world, dist, torch, n_prims)
        bcast = {"bytes": nbytes, "ms": ms, "GB_s": nbytes / ms / 1e6}
    bbox = tree.bbox()

    if args.small:   # same camera over the smaller terrain; profiling only
        rays_h = np.ascontiguousarray(W.pinhole_rays(IMG // 4, IMG // 4, nx, ny)[rank % 4::4])
    else:
        rays_h = shard_rays(rank)
    n_rays = len(rays_h)
    pts_h = shard_points(bbox, rank, n_rays)
    n_pts = len(pts_h)
    rays_d, pts_d = torch.from_numpy(rays_h).cuda(), torch.from_numpy(pts_h).cuda()
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

    # e2e: the same step through the C ABI with pinned HOST buffers (copies inside the timed region)
    e2e_ms = None
    h2d = n_rays * 48 + n_pts * 24
    d2h = n_rays * (4 + 8 + 24 + 1) + n_pts * (24 + 4 + 8 + 1)
    if not args.no_e2e:
        def pinned(shape, dtype):
            return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
        rays_p, pts_p = pinned((n_rays, 6), torch.float64), pinned((n_pts, 3), torch.float64)
        rays_p[:], pts_p[:] = rays_h, pts_h
        hout_r = {"prim": pinned(n_rays, torch.int32).view(np.uint32), "t": pinned(n_rays, torch.float64), "point": pinned((n_rays, 3), torch.float64), "flags": pinned(n_rays, torch.uint8)}
        hout_p = {"point": pinned((n_pts, 3), torch.float64), "prim": pinned(n_pts, torch.int32).view(np.uint32), "sqdist": pinned(n_pts, torch.float64), "flags": pinned(n_pts, torch.uint8)}

        def host_step():
            tree.first_intersection(rays_p, want=want_r, out=hout_r)
            tree.closest_point_and_primitive(pts_p, want=want_p, out=hout_p)
        host_step()
        barrier()
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(args.steps):
            host_step()
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)
        # the HOST-pointer path must return what the DEVICE-pointer path returned
        assert np.array_equal(hout_r["prim"], out_r["prim"].cpu().numpy().view(np.uint32)) and np.array_equal(hout_p["sqdist"], out_p["sqdist"].cpu().numpy())

    if dist is not None:
        tm = torch.tensor([total_ms, e2e_ms or 0.0, ray_ms, pt_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms, e2e_max, ray_ms, pt_ms = [float(x) for x in tm.tolist()]
        e2e_ms = e2e_max if e2e_ms is not None else None
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
        line = {"metric": "Mqueries/s (50% first_intersection Ray_3 + 50% closest_point_and_primitive)", "value": value, "unit": "Mqueries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD if not args.small else "SMALL profiling variant (1M-triangle terrain): not a bench value", "triangles": n_prims,
                           "rays_per_gpu": n_rays, "points_per_gpu": n_pts, "l2": "inputs (%.0f MB per GPU) larger than L2" % ((n_rays * 48 + n_pts * 24) / 1e6),
                           "parallelism": "tree replicated by NCCL broadcast, queries sharded x%d" % world},
                "detail": {"first_intersection_Mrays_s": world * n_rays / ray_ms / 1e3, "closest_point_Mqueries_s": world * n_pts / pt_ms / 1e3,
                           "build_ms": build_ms, "build_Mtri_s": n_prims / build_ms / 1e3, "build_GB_s_algorithmic(492B/tri)": 492.0 * n_prims / build_ms / 1e6,
                           "ray_hit_rate_rank0": hit_rate, "ties": ties, "near_ties": near, "exact_stage_queries": exact, "traversal_rank0": traffic_model,
                           "tree_broadcast": bcast},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(dom) if not args.small and n_pts == POINTS_PER_RANK else None,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                             "note": "traversal is bound by the L1 data stage (one wavefront per scattered node / triangle fetch) and by instruction issue (51 %) at 20 warps/SM, not by HBM (DESIGN.md §5); "
                                     "traffic = dram__bytes_read+write of one launch from the ncu --set full capture named in profiles/README.md"},
                "clocks": clocks, "gpu_launches": launches}
        if e2e_ms is not None:
            line["e2e"] = {"value": q_per_step * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mqueries/s", "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h}
        if world == 1 and not args.no_cpu_baseline:
            o, _ = oracle_tree(verts, faces)
            threads = os.cpu_count() or 1
            nr, npnt = 1 << 20, 1 << 17
            stride = max(1, n_rays // nr)
            cr = np.ascontiguousarray(rays_h[::stride][:nr])
            cp = pts_h[:npnt]
            t0 = time.perf_counter()
            rr = o.first_intersection(cr, threads=threads)
            t1 = time.perf_counter()
            rc = o.closest_point(cp, threads=threads)
            t2 = time.perf_counter()
            rps, pps = len(cr) / (t1 - t0), len(cp) / (t2 - t1)
            # the sample doubles as a parity check of the benchmarked batch against the oracle
            g_prim = out_r["prim"].cpu().numpy().view(np.uint32)[::stride][:nr]
            g_t = out_r["t"].cpu().numpy()[::stride][:nr]
            g_tie = (out_r["flags"].cpu().numpy()[::stride][:nr] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
            p_prim = out_p["prim"].cpu().numpy().view(np.uint32)[:npnt]
            p_tie = (out_p["flags"].cpu().numpy()[:npnt] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
            parity = {"rays_checked": len(cr), "ray_t_bit_equal_outside_near_ties": bool(np.array_equal(g_t[~g_tie], rr["t"][~g_tie])),
                      "ray_t_max_rel_diff": float(np.max(np.abs(g_t - rr["t"]) / np.maximum(np.abs(rr["t"]), 1e-300))), "ray_id_mismatch_not_tie": int(((g_prim != rr["prim"]) & ~g_tie).sum()),
                      "points_checked": len(cp), "sqdist_bit_equal": bool(np.array_equal(out_p["sqdist"].cpu().numpy()[:npnt], rc["sqd"])),
                      "point_id_mismatch_not_tie": int(((p_prim != rc["prim"]) & ~p_tie).sum())}
            line["cpu_baseline"] = {"value": 2.0 / (1.0 / rps + 1.0 / pps) / 1e6, "unit": "Mqueries/s", "cores": threads, "kind": "port",
                                    "sample": "%d rays (every %dth of the shard) + %d points against the full tree; 50/50 mixed rate" % (len(cr), stride, len(cp)),
                                    "first_intersection_Mrays_s": rps / 1e6, "closest_point_Mqueries_s": pps / 1e6, "build_s": o.build_seconds(),
                                    "build_Mtri_s": n_prims / o.build_seconds() / 1e6, "kd_hint_build_s": o.kd_seconds(), "parity_on_sample": parity}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
