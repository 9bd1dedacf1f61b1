"""Row f4 -- `MAP` (N/3DM/tsdf.py:56-108), the tensor-pipeline reconstruction: CUDA VoxelBlockGrid-style integration
(16^3 blocks, depth-touch activation, projective sdf, depth_max) + extraction with the tensor pipeline's defaults,
against the oracle's restatement (orc_vbg_integrate; parity vs real Open3D unpinned) -- `-m gpu`."""
import inspect

import numpy as np
import pytest
import torch

import oracle
from bodyslam_b200.tsdf import MAP
from util import canon_mesh, small_scene

pytestmark = pytest.mark.gpu


def k3x3(K):
    return np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], dtype=np.float64)


class FakeRGBD:
    """the attributes MAP.integrate reads from the reference's RGBD (slam_utils.py:172-207)"""

    def __init__(self, depth_u16, color, depth_scale):
        self.o3d_t_depth, self.o3d_t_color = depth_u16, color
        d = depth_u16.astype(np.float32) / depth_scale
        self.depth_min, self.depth_max = float(d.min()), float(d.max())


def test_map_signature_matches_the_reference():
    sig = inspect.signature(MAP.__init__)
    names = list(sig.parameters)[1:9]
    assert names == ["width", "height", "intrinsic", "device", "depth_scale", "voxel_size", "block_count", "trunc_voxel_multiplier"]
    assert sig.parameters["voxel_size"].default == 0.0058 and sig.parameters["block_count"].default == 40000
    assert sig.parameters["trunc_voxel_multiplier"].default == 8.0
    assert list(inspect.signature(MAP.integrate).parameters)[1:] == ["curr_rgbd", "i", "curr_global_pose"]
    for m in ("extract_pcd", "extract_mesh", "save_pcd", "save_mesh"):
        assert callable(getattr(MAP, m))


@pytest.mark.parametrize("scene,res,vs", [("laparoscopy512", 96, 0.0058), ("colonoscopy256", 64, 0.008)])
def test_map_integrate_matches_oracle(cuda, scene, res, vs, tmp_path):
    sc = small_scene(scene, res=64, frame_ids=np.arange(0, 10, 2))     # neighbouring views: voxels reach the weight threshold 3
    bs = vs * 16
    centre = sc["origin"] + 0.5 * 64 * sc["voxel_length"]
    origin = np.floor((centre - 0.5 * res * vs) / bs + 0.5) * bs
    m = MAP(640, 480, k3x3(sc["K"]), "CUDA:0", 1000.0, voxel_size=vs, trunc_voxel_multiplier=4.0, resolution=res, origin=origin)
    V = oracle.o3d.Volume(res, vs, vs * 4.0, origin, with_color=True)
    poses = [np.linalg.inv(E) for E in sc["E"]]
    co, blocks = [], 0
    for i in range(5):
        fr = FakeRGBD(sc["depth_u16"][i], sc["color"][i], 1000.0)
        m.integrate(fr, i, poses[i])
        n, touched = V.integrate_vbg(sc["depth_u16"][i], sc["K"], poses[i], rgb=sc["color"][i], depth_scale=1000.0, depth_max=fr.depth_max,
                                     trunc_voxel_multiplier=4.0, return_touched=True)
        co.append(n)
        blocks += int(touched.sum())
    assert sum(co) > 1000 and blocks > 0, (co, blocks)
    t, w, c = (x.cpu().numpy() for x in m.model.export_dense(with_color=True))
    assert np.array_equal(w, V.grid("weight")), "weights / occupancy differ"
    assert np.array_equal(t, V.grid("tsdf")), f"tsdf differs by {np.abs(t - V.grid('tsdf')).max()}"
    assert np.array_equal(c.reshape(-1), V.color)
    # the batch form gives the same map and per-frame update counts
    m2 = MAP(640, 480, k3x3(sc["K"]), "CUDA:0", 1000.0, voxel_size=vs, trunc_voxel_multiplier=4.0, resolution=res, origin=origin)
    counts = torch.zeros(5, dtype=torch.int64, device=cuda)
    dmax = [float((sc["depth_u16"][i].astype(np.float32) / 1000.0).max()) for i in range(5)]
    m2.integrate_batch(sc["depth_u16"], sc["color"], np.stack(poses), dmax, update_counts=counts)
    assert counts.cpu().tolist() == co
    for x, y in zip(m2.model.export_dense(True), m.model.export_dense(True)):
        assert torch.equal(x, y)
    # extraction with the tensor pipeline's defaults: weight >= 3, vertices on voxel corners
    oracle.o3d.set_extract_flavour(3.0, 0.0)
    try:
        ref, refp = V.extract_mesh(), V.extract_points()
    finally:
        oracle.o3d.set_extract_flavour(0.0, 0.5)
    mesh, pcd = m.extract_mesh(), m.extract_pcd()
    a = canon_mesh(mesh.vertices.cpu().numpy(), mesh.vertex_keys.cpu().numpy(), mesh.triangles.cpu().numpy(), (res,) * 3)
    b = canon_mesh(ref["vertices"], ref["keys"], ref["triangles"], (res,) * 3)
    assert len(ref["triangles"]) > (100 if scene == "laparoscopy512" else 0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.abs(a[1] - b[1]).max() <= 1e-4 * vs
    assert int(pcd.points.shape[0]) == len(refp["points"])
    # a lower threshold keeps more surface
    legacy_like = MAP(640, 480, k3x3(sc["K"]), "CUDA:0", 1000.0, voxel_size=vs, trunc_voxel_multiplier=4.0, resolution=res, origin=origin,
                      weight_threshold=0.0)
    legacy_like.integrate_batch(sc["depth_u16"], sc["color"], np.stack(poses), dmax)
    assert legacy_like.extract_mesh().triangles.shape[0] >= mesh.triangles.shape[0]
    m.save_mesh(str(tmp_path / "m.ply")); m.save_pcd(str(tmp_path / "p.ply"))
    st = m.clip_stats()
    assert st["points"] > 0 and 0 <= st["outside"] <= st["points"]


def test_map_rejects_bad_input(cuda):
    sc = small_scene("laparoscopy512", res=32, frames=1)
    m = MAP(640, 480, k3x3(sc["K"]), "CUDA:0", 1000.0, resolution=64)
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        m.integrate_batch(sc["depth_u16"].astype(np.float32), sc["color"], np.eye(4)[None], 1.0)
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        m.integrate_batch(sc["depth_u16"][:, :100], sc["color"], np.eye(4)[None], 1.0)
    with pytest.raises(RuntimeError, match="whole 16"):
        MAP(640, 480, k3x3(sc["K"]), "CUDA:0", 1000.0, resolution=72)
