"""The device-side workload generators and checksums (include/b200work.h) against their numpy originals: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


def test_bbox_points_match_numpy(torch):
    from cgal_b200 import device_workloads as D
    from cgal_b200 import workloads as W
    bbox = np.array([0.0, -3.5, -41.25, 4096.0, 2048.0, 39.0625])
    for n, off in ((1, 0), (1000, 0), (77777, 123456789), (1 << 20, (1 << 33) + 5)):
        out = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        D.bbox_points(bbox, n, out, offset=off)
        np.testing.assert_array_equal(out.cpu().numpy(), W.bbox_points(bbox, n, offset=off))


def test_pinhole_rays_match_numpy(torch):
    from cgal_b200 import device_workloads as D
    from cgal_b200 import workloads as W
    cam = D.pinhole_camera(4096, 2048)
    np.testing.assert_array_equal(D.pinhole_rays_host(cam, 512, 512, (100, 108)), W.pinhole_rays(512, 512, 4096, 2048, rows=(100, 108)))
    for width, height, r0, nr in ((8192, 8192, 4000, 16), (512, 384, 0, 384), (8192, 8192, 8191, 1)):
        out = torch.empty((nr * width, 6), dtype=torch.float64, device="cuda")
        D.pinhole_rays(cam, width, height, r0, nr, out)
        np.testing.assert_array_equal(out.cpu().numpy(), W.pinhole_rays(width, height, 4096, 2048, rows=(r0, r0 + nr)))
    cam2 = D.pinhole_camera(4096, 2048, eye_xy=(512.0, 256.0))
    out = torch.empty((4 * 1024, 6), dtype=torch.float64, device="cuda")
    D.pinhole_rays(cam2, 1024, 1024, 500, 4, out)
    np.testing.assert_array_equal(out.cpu().numpy(), D.pinhole_rays_host(cam2, 1024, 1024, (500, 504)))


def test_checksum_matches_numpy_and_adds_over_shards(torch):
    from cgal_b200 import device_workloads as D
    rng = np.random.default_rng(1)
    for dtype, n in ((np.float64, 100003), (np.uint32, 77777), (np.uint8, 1000001)):
        a = rng.integers(0, 255, size=n * np.dtype(dtype).itemsize, dtype=np.uint8).view(dtype)
        t = torch.from_numpy(a.view(np.uint8)).cuda()
        assert D.checksum128(t) == D.checksum128_host(a)
        assert D.checksum128(t, index0=12345) == D.checksum128_host(a, index0=12345)
    a = rng.standard_normal(4096 * 3)
    t = torch.from_numpy(a).cuda()
    whole = D.checksum128(t)
    parts = (0, 0)
    for k in range(4):
        parts = D.add128(parts, D.checksum128(t[k * 3072:(k + 1) * 3072], index0=k * 3072))
    assert parts == whole
