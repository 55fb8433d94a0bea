"""Multi-GPU plumbing of the query path: one process per GPU over torch.distributed.

The tree is built on ONE rank, broadcast as flat device buffers (NCCL over NVLink 5 / NVSwitch), and every rank answers a
contiguous block of the query batch against its replica.  There is no collective in the steady-state query loop; results
stay on the rank that computed them unless the caller asks for an all-gather.

This replaces the caller-side `tbb::parallel_for` over queries of the reference
(Polygon_mesh_processing/include/CGAL/Polygon_mesh_processing/distance.h:127-169).
The functions only need an object with export_layout() / import_alloc(n) / import_finish() (cgal_b200.AABB_tree), so the
host-side logic is testable with the gloo backend on CPU (tests/test_sharding_gloo.py).
"""
import numpy as np

BROADCAST_CHUNK = 1 << 30      # bytes per broadcast call (keeps NCCL's int element counts far from overflow)


def block_range(n, rank, world):
    """[begin, end) of rank's block when n queries are split into `world` contiguous blocks (sizes differ by at most 1)"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank %r outside world %r" % (rank, world))
    return (rank * n) // world, ((rank + 1) * n) // world


def block_sizes(n, world):
    return [block_range(n, r, world)[1] - block_range(n, r, world)[0] for r in range(world)]


class _CudaView:
    """exposes a raw device buffer to torch without copying"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def cuda_view(ptr, nbytes):
    import torch
    return torch.as_tensor(_CudaView(ptr, nbytes), device="cuda")


def broadcast_tree(tree, n_primitives, src=0, group=None, as_tensor=cuda_view):
    """Replicates the tree built on rank `src` to every rank.  n_primitives is only read on `src`.
    Returns (primitives, bytes received/sent).  Collective: every rank must call it."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    count = torch.tensor([int(n_primitives) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(count, src=src, group=group)
    n = int(count.item())
    if rank != src:
        tree.import_alloc(n)
    total = 0
    for ptr, nbytes in tree.export_layout():
        if nbytes == 0:
            continue
        view = as_tensor(ptr, nbytes)
        for off in range(0, nbytes, BROADCAST_CHUNK):
            dist.broadcast(view[off:off + BROADCAST_CHUNK], src=src, group=group)
        total += nbytes
    if rank != src:
        tree.import_finish()
    return n, total


def sharded_query(fn, queries, group=None, gather=False):
    """Runs fn(block) on this rank's block of `queries` (an array-like indexed along axis 0, identical on every rank).
    fn returns an array or a dict of arrays with one row per query of the block.
    gather=False: returns (begin, end, result_of_block).  gather=True: every rank gets the results of the WHOLE batch
    (all_gather of blocks padded to the largest block)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(queries)
    b, e = block_range(n, rank, world)
    res = fn(queries[b:e])
    if not gather:
        return b, e, res
    sizes = block_sizes(n, world)
    pad = max(sizes) if sizes else 0

    def gather_one(x):
        is_np = isinstance(x, np.ndarray)
        np_dtype = x.dtype if is_np else None
        if is_np and np_dtype == np.uint32:      # primitive ids: collectives lack uint32, the bit pattern travels as int32
            x = x.view(np.int32)
        t = torch.from_numpy(np.ascontiguousarray(x)) if is_np else x
        if is_np and dist.get_backend(group) == "nccl":
            t = t.cuda()
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[:t.shape[0]] = t
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        full = torch.cat([p[:s] for p, s in zip(parts, sizes)])
        return full.cpu().numpy().view(np_dtype) if is_np else full

    if isinstance(res, dict):
        return 0, n, {k: gather_one(v) for k, v in res.items()}
    return 0, n, gather_one(res)
