"""-m gpu: the multi-GPU entry points of the C ABI (b200aabb_comm_*): NCCL broadcast of a built tree / of the primitives + rebuild.
With one GPU the communicator has one rank (the plumbing still runs); with two or more, one process drives them (ncclCommInitAll)."""
import numpy as np
import pytest

import orc
from gpu_util import check_closest_point, check_first_intersection

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ngpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.cuda.device_count()


def test_single_rank_communicator(meshes, ngpu):
    import cgal_b200 as B
    v, f = meshes["elephant"]
    c = B.Comm(0, 0, 1, B.Comm.unique_id())
    t = B.AABB_tree(v, f)
    for mode in (0, 1):
        ms, nb = c.broadcast_tree(t, root=0, mode=mode)
        assert ms >= 0 and nb > 0
    assert t.size() == len(f)


@pytest.mark.parametrize("mode", [0, 1])
def test_replicas_answer_like_the_root(meshes, ngpu, mode):
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    if ngpu < 2:
        pytest.skip("needs two GPUs")
    n = min(ngpu, 4)
    v, f = meshes["elephant"]
    comms = B.Comm.init_all(list(range(n)))
    trees = [B.AABB_tree(device=d) for d in range(n)]
    trees[0].insert(v, f)
    ms, nb = B.broadcast_tree_all(comms, trees, root=0, mode=mode)
    assert nb >= len(f) * 72
    o = orc.OracleTree(v, f)
    bb = trees[0].bbox()
    rays, pts = W.sphere_rays(bb, 3000), W.bbox_points(bb, 3000)
    for t in trees:
        assert t.size() == len(f) and np.array_equal(t.bbox(), bb)
        check_first_intersection(t.first_intersection(rays), o.first_intersection(rays), o, rays)
        check_closest_point(t.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)
    # a replica received as a built tree has no primitives to rebuild from: build() is a no-op, never a crash
    trees[1].build()
    check_closest_point(trees[1].closest_point_and_primitive(pts), o.closest_point(pts), o, pts)


@pytest.mark.parametrize("mode", [0, 1])
def test_replicas_of_a_segment_tree(meshes, ngpu, mode):
    """the kind of primitive travels with the broadcast: a replica of a tree of Segment_3 primitives projects on segments too, and refuses what the root refuses"""
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    from test_segment_primitives import mesh_edges
    if ngpu < 2:
        pytest.skip("needs two GPUs")
    v, f = meshes["elephant"]
    segs = mesh_edges(v, f)
    comms = B.Comm.init_all([0, 1])
    trees = [B.AABB_tree(device=d) for d in range(2)]
    trees[0].insert(segments=segs)
    B.broadcast_tree_all(comms, trees, root=0, mode=mode)
    pts = W.bbox_points(trees[0].bbox(), 3000)
    want = orc.OracleTree(segments=segs).closest_point(pts)
    for t in trees:
        assert t.size() == len(segs)
        assert np.array_equal(t.closest_point_and_primitive(pts)["sqdist"], want["sqd"])
        with pytest.raises(B.B200AABBError):
            t.first_intersection(np.zeros((2, 6)))
    trees[0].insert(v, f)                                                   # triangles again: the replicas follow
    B.broadcast_tree_all(comms, trees, root=0, mode=mode)
    assert trees[1].first_intersection(W.sphere_rays(trees[1].bbox(), 16))["prim"].shape == (16,)
