"""Host-side pieces of the export / projection callers ("next" rows f1, f3 of SURVEY.md section 8): file formats the reference's
downstream stages read (segmentation/segmenter.py, segmentation/merger.py), ray sharding, the crop OBB, and the oracle's
restatement of nerfstudio's pinhole ray generation / AABB slab test on hand-computable cases.  No GPU."""
import os

import numpy as np
import torch

import helpers  # noqa: F401  (sys.path)
from cropnerf_b200 import export
from oracle import nerfstudio_torch as ns


def test_ply_layout_is_what_the_segmenter_reads(tmp_path):
    pts = torch.tensor([[0.5, -1.25, 2.0], [1e-3, 7.0, -3.5]])
    rgb = torch.tensor([[0.0, 0.5, 1.0], [1.2, -0.1, 0.25]])
    path = str(tmp_path / "semantics_pc.ply")
    export.write_ply(path, pts, rgb)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode("ascii").splitlines()
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex 2"
    assert [ln.split()[1:] for ln in lines[3:9]] == [["float", "x"], ["float", "y"], ["float", "z"], ["uchar", "red"], ["uchar", "green"], ["uchar", "blue"]]
    rec = np.frombuffer(body, dtype=[("p", "<f4", 3), ("c", "u1", 3)])
    assert rec.shape == (2,)
    np.testing.assert_array_equal(rec["p"], pts.numpy())
    np.testing.assert_array_equal(rec["c"], np.array([[0, 127, 255], [255, 0, 63]], dtype=np.uint8))


def test_cluster_info_round_trip(tmp_path):
    clusters = [
        {"aabb": np.array([[[0, 0, 0], [1, 1, 1]], [[-1, -1, -1], [0, 0, 0]]], dtype=np.float64), "pcd": {0: np.zeros((4, 3)), 1: np.ones((2, 3))}},
        {"aabb": np.array([[[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]]]), "pcd": {0: np.zeros((1, 3))}},
    ]
    path = str(tmp_path / "all_super_cluster_info_nsub_2.npy")
    export.save_cluster_info(path, clusters)
    back = export.load_cluster_info(path)
    assert len(back) == 2 and back[0]["aabb"].shape == (2, 2, 3) and back[1]["aabb"].dtype == np.float32
    np.testing.assert_allclose(back[1]["aabb"][0, 1], [0.4, 0.5, 0.6], rtol=1e-6)
    assert set(back[0]["pcd"]) == {0, 1}
    # the reference reads it with np.load(..., allow_pickle=True) and indexes [i]['aabb'] (fruit_nerf.py:265-268)
    raw = np.load(path, allow_pickle=True)
    assert raw[0]["aabb"].shape[0] == 2


def test_png_quantisation_and_round_trip(tmp_path):
    import cv2

    img = torch.tensor([[[0.0, 0.5, 1.0], [0.25, 2.0, -1.0]]])  # [1,2,3]
    q = export.to_png_bytes(img)
    np.testing.assert_array_equal(q, np.array([[[0, 128, 255], [64, 255, 0]]], dtype=np.uint8))  # floor(x*255 + 0.5), clamped
    path = str(tmp_path / "wo_occ_cluster_0.png")
    export.write_png(path, img)
    back = cv2.imread(path, cv2.IMREAD_COLOR)[..., ::-1]
    np.testing.assert_array_equal(back, q)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10_000_000):
        for w in (1, 2, 3, 8):
            spans = [export.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_oriented_box_within():
    box = export.OrientedBox.from_params((1.0, 0.0, 0.0), (0.0, 0.0, np.pi / 2), (2.0, 1.0, 1.0))  # long axis rotated onto world y
    pts = torch.tensor([[1.0, 0.9, 0.0], [1.9, 0.0, 0.0], [1.0, 0.0, 0.49], [1.0, 0.0, 0.51]])
    assert box.within(pts).tolist() == [True, False, True, False]


def test_oracle_pinhole_rays_hand_cases():
    c2w = torch.tensor([[1.0, 0, 0, 0.5], [0, 1.0, 0, -0.25], [0, 0, 1.0, 2.0]])
    fx = fy = 100.0
    cx, cy = 32.0, 24.0
    # the pixel whose centre is the principal point looks straight down -z; one pixel to the right tilts towards +x, one pixel
    # down tilts towards -y (image rows grow downwards, OpenGL camera +y is up)
    coords = torch.tensor([[23.5 + 0.5, 31.5 + 0.5], [24.0, 33.0], [25.0, 32.0]])
    o, d, area = ns.generate_pinhole_rays(c2w, fx, fy, cx, cy, coords)
    assert torch.equal(o, c2w[:, 3].expand(3, 3))
    torch.testing.assert_close(d[0], torch.tensor([0.0, 0.0, -1.0]))
    n1 = (1 + 0.01**2) ** 0.5
    torch.testing.assert_close(d[1], torch.tensor([0.01 / n1, 0.0, -1 / n1]))
    torch.testing.assert_close(d[2], torch.tensor([0.0, -0.01 / n1, -1 / n1]))
    torch.testing.assert_close(area[0, 0], torch.tensor(1e-4), rtol=1e-3, atol=0)
    assert torch.allclose(d.norm(dim=-1), torch.ones(3), atol=1e-6)


def test_oracle_intersect_aabb_hand_cases():
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    o = torch.tensor([[0.0, 0.0, 3.0], [0.0, 0.0, 3.0], [0.0, 0.0, 0.0], [5.0, 5.0, 3.0]])
    d = torch.tensor([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]])
    tmin, tmax = ns.intersect_aabb(o, d, aabb)
    assert tmin.tolist() == [2.0, 1e10, 0.0, 1e10]  # hit, parallel miss, origin inside (near clamps to 0), offset miss
    assert tmax.tolist() == [4.0, 1e10, 1.0, 1e10]


def test_volume_surface_rays_match_reference_executed_fixture():
    """tests/golden/ref_volume_rays.npz = the reference's get_corners_of_aabb + sample_surface_points +
    OrthographicRayGenerator executed verbatim (oracle/ref_shim.py:run_volume_rays_case): bit for bit."""
    ref = np.load(os.path.join(helpers.GOLDEN, "ref_volume_rays.npz"))
    n = int(ref["n"])
    for name in ("exporter_default", "positive_z", "anisotropic"):
        o, d, far = export.volume_surface_rays(ref[name + "_aabb"], n)
        np.testing.assert_array_equal(o.numpy(), ref[name + "_origins"])
        np.testing.assert_array_equal(d.numpy()[None].repeat(o.shape[0], 0), ref[name + "_directions"])
        assert np.all(ref[name + "_fars"] == np.float32(far))
    assert np.allclose(export.rescale_to_dataparser(torch.ones(2, 3), 0.5).numpy(), 4.0)


def test_aabb_near_far_equals_oracle_intersect_aabb():
    g = torch.Generator().manual_seed(4)
    o = torch.randn((4000, 3), generator=g) * 1.5
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn((4000, 3), generator=g), dim=-1)  # roughly towards the box
    d[:50, 0] = 0.0  # axis-parallel rays: +-inf slab distances
    aabb = torch.tensor([[-0.4, -0.3, -0.5], [0.3, 0.5, 0.2]])
    n, f = export.aabb_near_far(o, d, aabb)
    tmin, tmax = ns.intersect_aabb(o, d, aabb.reshape(-1))
    hit = tmin < 1e10
    assert 200 < int(hit.sum()) < 3800
    assert torch.equal(n[:, 0] < 1e10, hit)
    assert torch.equal(n[:, 0], tmin) and torch.equal(f[:, 0], tmax)


def test_cluster_info_loads_without_running_pickled_code(tmp_path):
    """the reference's cluster-info file is a pickled object array; the loader only reconstructs numpy arrays / plain containers
    and refuses anything else unless the caller opts into the reference's unrestricted np.load."""
    import os
    import pickle

    import numpy as np
    import pytest

    import cropnerf_b200.export as E

    good = [{"aabb": np.random.rand(3, 2, 3).astype(np.float32), "pcd": {0: np.random.rand(5, 3)}}, {"aabb": np.random.rand(1, 2, 3), "pcd": {}}]
    p = str(tmp_path / "good.npy")
    E.save_cluster_info(p, good)
    out = E.load_cluster_info(p)
    assert len(out) == 2 and out[0]["aabb"].shape == (3, 2, 3) and np.allclose(out[0]["pcd"][0], good[0]["pcd"][0])
    assert np.allclose(E.load_cluster_info(p, allow_pickle=True)[1]["aabb"], out[1]["aabb"])

    class Evil:
        def __reduce__(self):
            return (os.getcwd, ())

    q = str(tmp_path / "evil.npy")
    np.save(q, np.asarray([{"aabb": Evil()}], dtype=object), allow_pickle=True)
    with pytest.raises(pickle.UnpicklingError):
        E.load_cluster_info(q)
