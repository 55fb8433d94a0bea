"""Host-side multi-GPU logic (cgal_b200/sharding.py) with world_size 2 over gloo on CPU: tree broadcast as flat buffers,
contiguous query blocks, all-gather of results.  The "device" tree is the host emulation of the device headers (tests/hc.py)."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class FlatTree:
    """stands in for cgal_b200.AABB_tree: a built tree is three flat buffers (nodes 64 B, records 96 B, root box 48 B)"""

    def __init__(self, nodes=None, tris=None, bbox=None):
        self.bufs = [nodes, tris, bbox]
        self.finished = False

    def import_alloc(self, n):
        self.bufs = [np.zeros(max(n - 1, 1) * 64, np.uint8), np.zeros(n * 96, np.uint8), np.zeros(6)]

    def export_layout(self):
        return [(b.ctypes.data, b.nbytes) for b in self.bufs]

    def import_finish(self):
        self.finished = True


def host_view(ptr, nbytes):
    import torch
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr)))


def worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import hc
    from cgal_b200 import sharding, workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(os.path.join(HERE, "golden", "meshes.npz"))
        soup = W.soup_of(z["coverrear_v"], z["coverrear_f"])
        host = hc.HostTree(soup)                      # every rank can emulate the device build; only rank 0's is broadcast
        if rank == 0:
            nodes, tris, bbox = host.export()
            tree = FlatTree(nodes, tris, bbox)
        else:
            tree = FlatTree()
        n, nbytes = sharding.broadcast_tree(tree, len(soup) if rank == 0 else -1, src=0, as_tensor=host_view)
        assert n == len(soup) and nbytes == (n - 1) * 64 + n * 96 + 48
        assert rank == 0 or tree.finished
        ref_nodes, ref_tris, ref_bbox = host.export()
        assert np.array_equal(tree.bufs[0], ref_nodes) and np.array_equal(tree.bufs[1], ref_tris) and np.array_equal(tree.bufs[2], ref_bbox)
        assert hc.validate(tree.bufs[0], tree.bufs[1], n) == 0

        pts = W.bbox_points(ref_bbox, 1001)           # odd size: unequal blocks
        b, e, part = sharding.sharded_query(lambda q: host.closest_point(q), pts)
        assert (b, e) == sharding.block_range(len(pts), rank, world) and len(part["prim"]) == e - b
        _, _, full = sharding.sharded_query(lambda q: {k: v for k, v in host.closest_point(q).items() if k in ("cp", "prim", "sqd")}, pts, gather=True)
        whole = host.closest_point(pts)
        for k in ("cp", "prim", "sqd"):
            assert full[k].dtype == whole[k].dtype and np.array_equal(full[k], whole[k]), k
        segs = W.bbox_segments(ref_bbox, 777)
        _, _, counts = sharding.sharded_query(lambda q: host.intersect(q, False, True), segs, gather=True)
        assert np.array_equal(counts, host.intersect(segs, False, True))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_block_ranges_partition_the_batch():
    from cgal_b200 import sharding
    for n in (0, 1, 7, 8, 1000, 8388608):
        for world in (1, 2, 3, 8):
            r = [sharding.block_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            s = sharding.block_sizes(n, world)
            assert sum(s) == n and max(s) - min(s) <= 1
    with pytest.raises(ValueError):
        sharding.block_range(10, 2, 2)


def test_broadcast_and_sharded_queries_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
