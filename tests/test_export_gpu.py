"""Projection export ("next" rows f1 + f3): ``project_clusters`` (device ray generation + AABB clip -> render of the hit rays ->
PNG tree for segmentation/merger.py) against the oracle running the reference's loop (fruit_nerf.py:254-318) ray by ray."""
import os

import cv2
import numpy as np
import pytest
import torch

from helpers import product_model

from cropnerf_b200 import export, synthetic
from oracle import cases
from oracle import nerfstudio_torch as ns

pytestmark = pytest.mark.gpu


def _oracle_images(oracle, cam, box):
    H, W = cam.height, cam.width
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords = torch.stack([yy, xx], -1).reshape(-1, 2).float() + 0.5
    o, d, area = ns.generate_pinhole_rays(cam.c2w, cam.fx, cam.fy, cam.cx, cam.cy, coords)
    tmin, tmax = ns.intersect_aabb(o, d, torch.as_tensor(box, dtype=torch.float32).reshape(-1))
    valid = tmin < 1e10
    wo = torch.zeros((H * W, 3))
    vis = torch.zeros((H * W, 3))
    if int(valid.sum()) >= 10:
        rb = ns.RayBundle(origins=o[valid].contiguous(), directions=d[valid], pixel_area=area[valid], camera_indices=torch.zeros((int(valid.sum()), 1), dtype=torch.long),
                          nears=tmin[valid][:, None], fars=tmax[valid][:, None])
        with torch.no_grad():
            wo[valid] = oracle(rb)["semantics"].expand(-1, 3)
            rb.fars = rb.nears
            rb.nears = torch.zeros_like(rb.nears)
            front = torch.zeros(H * W)
            front[valid] = oracle.get_density_for_ray_bundle(rb)
        vis = wo.clone()
        vis[front >= 0.5] = 0.0
    return export.to_png_bytes(wo.view(H, W, 3)), export.to_png_bytes(vis.view(H, W, 3)), valid.view(H, W)


@pytest.mark.parametrize("world_size", [1, 2])
def test_project_clusters_png_tree_matches_oracle(dev, tmp_path, world_size):
    cfg = cases.make_config(dict(log2_hashmap_size=14))
    oracle, state = cases.build_oracle(cfg, 6, 0, 0.5)
    oracle.eval()
    model = product_model(cfg, state, 6, dev, False)
    c2ws = synthetic.make_cameras(3, seed=7)
    cams = [export.PinholeCamera(c2w=c2ws[i], fx=60.0, fy=60.0, cx=24.0, cy=18.0, width=48, height=36) for i in range(3)]
    clusters = [
        {"aabb": np.array([[[-0.25, -0.2, -0.2], [0.15, 0.2, 0.25]], [[0.1, 0.1, 0.1], [0.4, 0.35, 0.3]]], dtype=np.float32), "pcd": {}},
        {"aabb": np.array([[[5.0, 5.0, 5.0], [5.1, 5.1, 5.1]]], dtype=np.float32), "pcd": {}},  # off-screen box: black images
    ]
    info = str(tmp_path / "all_super_cluster_info_nsub_2.npy")
    export.save_cluster_info(info, clusters)
    clusters = export.load_cluster_info(info)
    seg = []
    for i in range(3):
        p = str(tmp_path / f"seg_{i}.png")
        cv2.imwrite(p, np.full((4, 4), i, np.uint8))
        seg.append(p)
    out_dir = str(tmp_path / "projection")
    total = {"pairs": 0, "images": 0, "rays": 0}
    for rank in range(world_size):  # ranks write disjoint (super-cluster, camera) directories; no communication
        st = export.project_clusters(model, cams, clusters, out_dir, rank=rank, world_size=world_size, segmentation_files=seg)
        for k in total:
            total[k] += st[k]
    assert total["pairs"] == 6 and total["images"] == 2 * (2 * 3 + 1 * 3) and total["rays"] > 0
    n_checked = 0
    for k, cluster in enumerate(clusters):
        for j, cam in enumerate(cams):
            cam_dir = os.path.join(out_dir, f"super_cluster_{k}", f"cam_{j}")
            assert os.path.exists(os.path.join(cam_dir, f"seg_{j}.png"))
            for i in range(cluster["aabb"].shape[0]):
                wo_ref, vis_ref, valid = _oracle_images(oracle, cam, cluster["aabb"][i])
                wo = cv2.imread(os.path.join(cam_dir, f"wo_occ_cluster_{i}.png"), cv2.IMREAD_COLOR)[..., ::-1]
                vis = cv2.imread(os.path.join(cam_dir, f"visible_cluster_{i}.png"), cv2.IMREAD_COLOR)[..., ::-1]
                assert wo.shape == (36, 48, 3)
                if k == 1:
                    assert not wo.any() and not vis.any() and not valid.any()
                    continue
                assert valid.sum() > 50, "test boxes should be on screen"
                # 8-bit images: allow one code of rounding, on all but 0.5 % of the pixels (slab-test ties on the box silhouette)
                assert (np.abs(wo.astype(int) - wo_ref.astype(int)) <= 1).mean() >= 0.995
                assert (np.abs(vis.astype(int) - vis_ref.astype(int)) <= 1).mean() >= 0.995
                assert not wo[~valid.numpy()].any()
                n_checked += 1
    assert n_checked == 6


@pytest.mark.parametrize("world_size", [1, 2])
def test_sample_volume_matches_oracle(dev, world_size):
    """Volumetric export (row f4): orthographic rays from the box face, per-sample thresholds, device-side compaction."""
    cfg = cases.make_config(dict(log2_hashmap_size=14))
    oracle, state = cases.build_oracle(cfg, 6, 0, 0.5)
    oracle.test_mode = oracle.field.test_mode = "export"
    oracle.setup_inference(True, 40)
    oracle.eval()
    model = product_model(cfg, state, 6, dev, False, test_mode="export")
    model.field.test_mode = "export"
    model.setup_inference(True, 40)
    model.eval()
    box = ((-0.6, -0.5, -0.4), (0.5, 0.6, 0.7))
    n_side = 12
    with pytest.raises(RuntimeError, match="export"):
        export.sample_volume(product_model(cfg, state, 6, dev, False), box, n_side)
    o, d, far = export.volume_surface_rays(box, n_side)
    rb = ns.RayBundle(origins=o, directions=d.repeat(o.shape[0], 1), pixel_area=torch.zeros(o.shape[0], 1), camera_indices=None,
                      nears=torch.zeros(o.shape[0], 1), fars=torch.full((o.shape[0], 1), far))
    with torch.no_grad():
        ref = oracle(rb)
    den, sem = ref["density"].reshape(-1), ref["semantics"].reshape(-1)
    thr_d, thr_s = float(den.median()), float(sem[den >= den.median()].median())
    parts = [export.sample_volume(model, box, n_side, num_rays_per_batch=50, rank=r, world_size=world_size, semantic_threshold=thr_s, density_threshold=thr_d)
             for r in range(world_size)]
    assert sum(p["stats"]["rays"] for p in parts) == o.shape[0] and sum(p["stats"]["samples"] for p in parts) == den.numel()
    masks = {"density": den >= thr_d, "semantic": (sem >= thr_s) & (den >= thr_d), "semantic_colormap": (ref["semantics_colormap"].reshape(-1) >= 0.999) & (den >= thr_d)}
    for key, mask in masks.items():
        pts = torch.cat([p[key]["points"] for p in parts])
        col = torch.cat([p[key]["colors"] for p in parts])
        n_ref = int(mask.sum())
        assert abs(pts.shape[0] - n_ref) <= max(2, int(2e-3 * n_ref)), (key, pts.shape[0], n_ref)  # threshold ties at fp32 round-off
        assert col.shape == (pts.shape[0], 4)
        if key == "density":
            assert n_ref > 100
        if pts.shape[0] == n_ref and n_ref:
            ref_pts = ref["point_location"].reshape(-1, 3)[mask]
            if torch.allclose(pts, ref_pts, atol=1e-6):  # same samples kept: colours must agree too
                ref_rgb = ref["rgb"].reshape(-1, 3)[mask]
                assert (col[:, :3] - ref_rgb).abs().max() < 2e-4


@pytest.mark.parametrize("world_size", [1, 2])
def test_generate_point_cloud_matches_oracle_loop(dev, world_size):
    """``ns-export pointcloud`` render loop (BASELINE config 3; export/exporter_utils_nerfacto.py:126-183): render a ray batch, back-project the
    median depth, keep rays labelled fruit (sigmoid(sem) > 0.9) inside the crop OBB, until num_points are collected; ranks draw disjoint ray
    streams and collect num_points / world_size each, with no communication."""
    cfg = cases.make_config(dict(log2_hashmap_size=14))
    oracle, state = cases.build_oracle(cfg, 20, 0, 0.5)
    state = dict(state)
    state["field.field_head_semantics.net.bias"] = torch.full_like(state["field.field_head_semantics.net.bias"], 3.0)  # a scene with fruit in it
    oracle.load_state_dict(state)
    oracle.eval()
    model = product_model(cfg, state, 20, dev, False)
    obb = export.OrientedBox.from_params((0.05, -0.05, 0.0), (0.0, 0.0, 0.3), (1.6, 1.7, 1.8))
    B, num_points = 512, 900

    def batch(rank, i):
        return synthetic.make_rays(B, seed=1000 * rank + i, num_cameras=20)

    def next_rays(rank):
        def fn(i):
            r = batch(rank, i)
            return export.RayBundle(r["origins"].to(dev), r["directions"].to(dev), r["pixel_area"].to(dev), r["camera_indices"].to(dev))
        return fn

    total = 0
    for rank in range(world_size):
        got = export.generate_point_cloud(model, next_rays(rank), num_points, crop_obb=obb, rank=rank, world_size=world_size)
        lo, hi = export.shard_range(num_points, rank, world_size)
        want = hi - lo
        assert got["points"].shape == (want, 3) and got["rgbs"].shape == (want, 3) and got["view_directions"].shape == (want, 3)
        total += want
        # the oracle running the reference's loop on the same ray stream
        pts, cols, used = [], [], 0
        while sum(p.shape[0] for p in pts) < want:
            r = batch(rank, used)
            with torch.no_grad():
                out = oracle(cases.oracle_bundle(r))
            p = r["origins"] + r["directions"] * out["depth"]
            mask = (out["semantics_colormap"][:, 0] > 0) & obb.within(p)
            pts.append(p[mask]); cols.append(out["rgb"][mask]); used += 1
        # the completing batch is the reference loop's last one; the pipelined loop may have launched up to `lag` = 2 batches more, whose
        # points fall past `want` and are dropped
        assert got["rays_needed"] == used * B, (got["rays_needed"], used * B)
        assert used * B <= got["rays_rendered"] <= (used + 2) * B
        ref_p, ref_c = torch.cat(pts)[:want], torch.cat(cols)[:want]
        gp, gc = got["points"].cpu(), got["rgbs"].cpu()
        # identical rays kept in identical order, up to label ties at the 0.9 threshold: compare the leading run that matches
        same = (gp - ref_p).abs().max(dim=1).values < 2e-4
        assert same.float().mean().item() >= 0.98, same.float().mean().item()
        assert (gc[same] - ref_c[same]).abs().max().item() < 2e-4
        assert bool(obb.within(gp).all())
    assert total == num_points


def test_render_cluster_projection_on_arbitrary_rays(dev):
    """The per-(camera, cluster) step of fruit_nerf.py:283-315 for a flat bundle of arbitrary rays (torch slab test + the two renders)."""
    cfg = cases.make_config(dict(log2_hashmap_size=14))
    oracle, state = cases.build_oracle(cfg, 6, 0, 0.5)
    oracle.eval()
    model = product_model(cfg, state, 6, dev, False)
    rays = synthetic.make_rays(1500, seed=9, num_cameras=6)
    aabb = torch.tensor([[-0.35, -0.3, -0.3], [0.3, 0.35, 0.4]])
    rb = export.RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"])
    got = export.render_cluster_projection(model, rb, aabb)
    tmin, tmax = ns.intersect_aabb(rays["origins"], rays["directions"], aabb.reshape(-1))
    valid = tmin < 1e10
    assert torch.equal(got["valid"].cpu(), valid) and int(valid.sum()) > 100
    ob = ns.RayBundle(origins=rays["origins"][valid], directions=rays["directions"][valid], pixel_area=rays["pixel_area"][valid],
                      camera_indices=rays["camera_indices"][valid], nears=tmin[valid][:, None], fars=tmax[valid][:, None])
    with torch.no_grad():
        sem = oracle(ob)["semantics"]
        ob.fars = ob.nears
        ob.nears = torch.zeros_like(ob.nears)
        front = oracle.get_density_for_ray_bundle(ob)
    g_sem, g_front = got["semantics"].cpu()[valid][:, :1], got["front_opacity"].cpu()[valid]
    assert ((g_sem - sem).abs() / (sem.abs() + 1e-2)).max().item() < 2e-3
    assert (g_front - front).abs().max().item() < 2e-4
    assert float(got["semantics"].cpu()[~valid].abs().max()) == 0.0
    occluded = got["front_opacity"] >= 0.5
    assert float(got["visible"][occluded].abs().max() if bool(occluded.any()) else 0.0) == 0.0
    assert torch.equal(got["visible"][~occluded], got["semantics"][~occluded])


def test_extract_points_kernel_against_torch(dev):
    """cnb_extract_points = exporter_utils_nerfacto.py:153-176 on one batch: same kept set as the torch expressions, in ray order, counts
    chained through the two device counters, entries past the capacity counted but not written."""
    import ctypes as C

    from cropnerf_b200 import _lib as L

    g = torch.Generator().manual_seed(5)
    obb = export.OrientedBox.from_params((0.1, -0.2, 0.05), (0.2, -0.1, 0.4), (1.2, 1.0, 0.9))
    arr = export._obb_array(obb)
    cap = 100
    pts = torch.empty((cap, 3), device=dev); cols = torch.empty((cap, 3), device=dev); dirs = torch.empty((cap, 3), device=dev)
    counters = torch.zeros((2,), device=dev, dtype=torch.int32)
    ref_p, ref_c, ref_d = [], [], []
    for i, n in enumerate((1000, 1, 777, 256)):
        o = torch.rand((n, 3), generator=g) - 0.5
        d = torch.nn.functional.normalize(torch.randn((n, 3), generator=g), dim=-1)
        depth = torch.rand((n, 1), generator=g) * 1.5
        sem = torch.randn((n, 1), generator=g) * 3.0
        rgb = torch.rand((n, 3), generator=g)
        scratch = torch.empty((int(L.lib().cnb_extract_points_scratch_ints(n)),), device=dev, dtype=torch.int32)
        od, dd, de, se, rg = (t.to(dev).contiguous() for t in (o, d, depth, sem, rgb))
        L.check(L.lib().cnb_extract_points(od.data_ptr(), dd.data_ptr(), de.data_ptr(), se.data_ptr(), rg.data_ptr(), n, arr, 1, 0.9, scratch.data_ptr(),
                                           counters[i & 1:].data_ptr(), counters[(i + 1) & 1:].data_ptr(), cap, pts.data_ptr(), cols.data_ptr(),
                                           dirs.data_ptr(), L.stream_ptr(dev)), "extract_points")
        p = o + d * depth
        mask = (torch.sigmoid(sem[:, 0]) - 0.9 > 0) & obb.within(p)
        ref_p.append(p[mask]); ref_c.append(rgb[mask]); ref_d.append(d[mask])
        assert int(counters[(i + 1) & 1]) == sum(x.shape[0] for x in ref_p), i
    ref_p, ref_c, ref_d = torch.cat(ref_p), torch.cat(ref_c), torch.cat(ref_d)
    total = ref_p.shape[0]
    assert total > cap, "the case must overflow the capacity"
    assert torch.equal(pts.cpu(), ref_p[:cap]) and torch.equal(cols.cpu(), ref_c[:cap]) and torch.equal(dirs.cpu(), ref_d[:cap])


def test_generate_rays_boxes_equals_one_box_at_a_time(dev):
    """cnb_generate_rays_boxes (all sub-cluster boxes of a super-cluster in one pass) = cnb_generate_rays(aabb=box) per box: the same
    hit pixels with bit-identical rays, near and far."""
    import ctypes as C

    from cropnerf_b200 import _lib as L

    c2w = synthetic.make_cameras(2, seed=3)[1]
    cam = export.PinholeCamera(c2w=c2w, fx=70.0, fy=65.0, cx=40.0, cy=30.0, width=80, height=60)
    boxes = torch.tensor([[-0.25, -0.2, -0.2, 0.15, 0.2, 0.25], [0.1, 0.1, 0.1, 0.4, 0.35, 0.3], [5.0, 5.0, 5.0, 5.1, 5.1, 5.1], [-0.25, -0.2, -0.2, 0.15, 0.2, 0.25]])
    k, npix = boxes.shape[0], 80 * 60
    cap = k * npix
    o = torch.empty((cap, 3), device=dev); d = torch.empty((cap, 3), device=dev); area = torch.empty((cap,), device=dev)
    nears = torch.empty((cap,), device=dev); fars = torch.empty((cap,), device=dev)
    tags = torch.empty((cap,), device=dev, dtype=torch.int32); count = torch.zeros((1,), device=dev, dtype=torch.int32)
    cs = export._camera_struct(cam)
    bd = boxes.to(dev)
    L.check(L.lib().cnb_generate_rays_boxes(C.byref(cs), bd.data_ptr(), k, cap, o.data_ptr(), d.data_ptr(), area.data_ptr(), nears.data_ptr(),
                                            fars.data_ptr(), tags.data_ptr(), count.data_ptr(), L.stream_ptr(dev)), "generate_rays_boxes")
    n = int(count)
    order = torch.argsort(tags[:n])
    tg = tags[:n][order].cpu()
    got = {name: t[:n][order].cpu() for name, t in (("o", o), ("d", d), ("area", area), ("near", nears), ("far", fars))}
    expect_tags = []
    for b in range(k):
        rb = export.generate_rays(cam.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, dev, aabb=boxes[b])
        valid = (rb.nears[:, 0] < 1e10).cpu()
        pix = torch.nonzero(valid)[:, 0]
        sel = (tg >= b * npix) & (tg < (b + 1) * npix)
        assert torch.equal(tg[sel] - b * npix, pix.to(torch.int32)), b
        assert torch.equal(got["d"][sel], rb.directions.cpu()[valid]) and torch.equal(got["o"][sel], rb.origins.cpu()[valid])
        assert torch.equal(got["near"][sel], rb.nears.cpu()[valid][:, 0]) and torch.equal(got["far"][sel], rb.fars.cpu()[valid][:, 0])
        assert torch.equal(got["area"][sel], rb.pixel_area.cpu()[valid][:, 0])
        expect_tags.append(pix.numel())
    assert n == sum(expect_tags) and expect_tags[2] == 0 and expect_tags[0] == expect_tags[3] > 50
    # a capacity that is too small: every hit is still counted
    count.zero_()
    L.check(L.lib().cnb_generate_rays_boxes(C.byref(cs), bd.data_ptr(), k, 16, o.data_ptr(), d.data_ptr(), area.data_ptr(), nears.data_ptr(),
                                            fars.data_ptr(), tags.data_ptr(), count.data_ptr(), L.stream_ptr(dev)), "generate_rays_boxes")
    assert int(count) == n


def test_volume_face_rays_device_bit_equal_to_the_host_construction(dev):
    """cnb_volume_face_rays = sample_surface_points + OrthographicRayGenerator (volume_surface_rays is pinned bit for bit to the
    reference-executed fixture ref_volume_rays.npz by test_export_formats_cpu): same origins, direction and ray length on the device."""
    for box, n_side in ((((-1.0, -1.0, -0.682), (1.0, 1.0, 1.318)), 37), (((-0.5, -0.25, 0.125), (0.5, 0.75, 0.625)), 13), (((-0.3, -0.1, -0.4), (0.6, 0.2, 0.1)), 50)):
        o, direction, far = export.volume_surface_rays(box, n_side)
        grid = export.volume_face_grid(box, n_side)
        assert grid["nx"] * grid["ny"] == o.shape[0]
        first, n = 5, o.shape[0] - 9
        rb = export.volume_face_rays_device(grid, first, n, dev)
        assert torch.equal(rb.origins.cpu(), o[first : first + n])
        assert torch.equal(rb.directions.cpu(), direction.repeat(n, 1))
        assert torch.equal(rb.fars.cpu(), torch.full((n, 1), far)) and not rb.nears.any()
