"""K4 parity: CUDA marching cubes / surface points vs the CPU oracle -- `-m gpu`.

Bar: vertex, triangle and point COUNTS exact; the vertex set identical as a set of voxel-edge
keys; triangles identical after canonical ordering; coordinates within 1e-4 * voxel scale."""
import numpy as np
import pytest
import torch

import oracle
from bodyslam_b200.tsdf import DenseTSDFVolume
from util import canon_mesh, mesh_is_closed_and_oriented, small_scene

pytestmark = pytest.mark.gpu


def analytic_field(n, kind, seed=0):
    """(tsdf, weight) [n,n,n] f32 test volumes"""
    g = (np.arange(n) + 0.5) / n - 0.5
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    w = np.ones((n, n, n), np.float32)
    if kind == "sphere":
        t = np.sqrt(x * x + y * y + z * z) - 0.3
    elif kind == "noise":  # smooth random field: exercises most of the 256 cube cases
        rng = np.random.default_rng(seed)
        t = np.zeros_like(x)
        for _ in range(12):
            k = rng.normal(size=3) * 14
            t += rng.normal() * np.sin(k[0] * x + k[1] * y + k[2] * z + rng.uniform(0, 6.28))
        t /= 4
    elif kind == "holes":  # random zero-weight voxels + clipped range
        rng = np.random.default_rng(seed)
        t = np.sqrt(x * x + y * y + z * z) - 0.3 + 0.03 * np.sin(40 * x) * np.sin(33 * y)
        w = (rng.uniform(size=t.shape) > 0.05).astype(np.float32) * rng.integers(1, 5, size=t.shape).astype(np.float32)
    t = np.clip(t * 8, -1, 1).astype(np.float32)
    return t, w


def oracle_volume(n, t, w, vl=0.01, origin=(0.1, -0.2, 0.3), color=None):
    V = oracle.o3d.Volume(t.shape, vl, 0.04, origin, with_color=color is not None)
    V.tsdf[:] = t.reshape(-1)
    V.weight[:] = w.reshape(-1)
    if color is not None:
        V.color[:] = color.reshape(-1)
    return V


def gpu_volume(cuda, t, w, vl=0.01, origin=(0.1, -0.2, 0.3), color=None, gz0=0, z_total=None):
    vol = DenseTSDFVolume(vl, 0.04, t.shape, origin, color=color is not None, device=cuda, gz0=gz0, z_total=z_total)
    vol.import_dense(t, w, color)
    return vol


def compare_mesh(mesh, ref, dims, vl, color=False):
    V, T = ref["vertices"], ref["triangles"]
    assert mesh.vertices.shape[0] == len(V), f"vertex count {mesh.vertices.shape[0]} vs {len(V)}"
    assert mesh.triangles.shape[0] == len(T), f"triangle count {mesh.triangles.shape[0]} vs {len(T)}"
    kc, vc, tc = canon_mesh(mesh.vertices.cpu().numpy(), mesh.vertex_keys.cpu().numpy(), mesh.triangles.cpu().numpy(), dims)
    ko, vo, to = canon_mesh(V, ref["keys"], T, dims)
    assert np.array_equal(kc, ko), "vertex edge keys differ"
    assert np.array_equal(tc, to), "triangles differ after canonical ordering"
    assert np.abs(vc - vo).max() <= 1e-4 * vl + 1e-6 * np.abs(vo).max()
    if color:
        order_g = np.argsort(((mesh.vertex_keys.cpu().numpy().astype(np.int64) * [4 * 10**8, 4 * 10**4, 4, 1]).sum(1)))
        order_o = np.argsort(((ref["keys"].astype(np.int64) * [4 * 10**8, 4 * 10**4, 4, 1]).sum(1)))
        assert np.abs(mesh.vertex_colors.cpu().numpy()[order_g] - ref["colors"][order_o]).max() <= 1e-5


@pytest.mark.parametrize("kind,n", [("sphere", 64), ("noise", 48), ("holes", 40), ("noise", 37)])
def test_marching_cubes_matches_oracle(cuda, kind, n):
    t, w = analytic_field(n, kind)
    ref = oracle_volume(n, t, w).extract_mesh()
    mesh = gpu_volume(cuda, t, w).extract_triangle_mesh()
    assert len(ref["triangles"]) > 500
    compare_mesh(mesh, ref, t.shape, 0.01)
    if kind == "sphere":
        assert mesh_is_closed_and_oriented(mesh.triangles.cpu().numpy())
        r = np.linalg.norm(mesh.vertices.cpu().numpy() - (np.array([0.1, -0.2, 0.3]) + 0.32), axis=1)
        assert np.abs(r - 0.3 * 0.64).max() < 0.01 * 0.6


def test_marching_cubes_ragged_box_with_color(cuda):
    t, w = analytic_field(48, "noise", seed=3)
    t, w = t[:45, :38, :41].copy(), w[:45, :38, :41].copy()
    col = np.random.default_rng(1).uniform(0, 255, size=t.shape + (3,)).astype(np.float32)
    ref = oracle_volume(0, t, w, color=col).extract_mesh()
    mesh = gpu_volume(cuda, t, w, color=col).extract_triangle_mesh()
    compare_mesh(mesh, ref, t.shape, 0.01, color=True)


def test_marching_cubes_empty_and_full(cuda):
    n = 16
    z = np.zeros((n, n, n), np.float32)
    m = gpu_volume(cuda, z, z).extract_triangle_mesh()
    assert m.vertices.shape[0] == 0 and m.triangles.shape[0] == 0
    m = gpu_volume(cuda, z + 0.5, z + 1).extract_triangle_mesh()   # all outside: no sign change
    assert m.vertices.shape[0] == 0 and m.triangles.shape[0] == 0


def test_mesh_after_integration_matches_oracle(cuda):
    sc = small_scene("laparoscopy512", res=128, frames=8)
    from test_gpu_tsdf import run_gpu, run_oracle
    V, _ = run_oracle(sc)
    vol, _ = run_gpu(sc, cuda)
    ref = V.extract_mesh()
    mesh = vol.extract_triangle_mesh()
    assert len(ref["triangles"]) > 5000
    compare_mesh(mesh, ref, (128,) * 3, sc["voxel_length"])


def test_slab_meshes_with_halos_concatenate_to_full_mesh(cuda):
    from bodyslam_b200.sharding import merge_slab_meshes
    t, w = analytic_field(48, "noise", seed=5)
    ref = oracle_volume(48, t, w).extract_mesh()
    parts = []
    bounds = [(0, 16), (16, 32), (32, 48)]
    vols = [gpu_volume(cuda, t[:, :, a:b].copy(), w[:, :, a:b].copy(), gz0=a, z_total=48, origin=(0.1, -0.2, 0.3)) for a, b in bounds]
    for i, vol in enumerate(vols):
        lo = vols[i - 1].export_plane(vols[i - 1].nz - 1) if i > 0 else None
        hi = vols[i + 1].export_plane(0) if i + 1 < len(vols) else None
        parts.append((vol.extract_triangle_mesh(halo_lo=lo, halo_hi=hi), bounds[i][0]))
    mesh = merge_slab_meshes([p for p, _ in parts], [z for _, z in parts], ny=48)
    compare_mesh(mesh, ref, t.shape, 0.01)


@pytest.mark.parametrize("kind,n", [("sphere", 48), ("noise", 40), ("holes", 40)])
def test_surface_points_match_oracle(cuda, kind, n):
    t, w = analytic_field(n, kind, seed=2)
    ref = oracle_volume(n, t, w).extract_points()
    pcd = gpu_volume(cuda, t, w).extract_point_cloud()
    assert pcd.points.shape[0] == len(ref["points"]) > 300
    kg = pcd.point_keys.cpu().numpy().astype(np.int64)
    ko = ref["keys"].astype(np.int64)
    cg = ((kg[:, 0] * 1000 + kg[:, 1]) * 1000 + kg[:, 2]) * 4 + kg[:, 3]
    co = ((ko[:, 0] * 1000 + ko[:, 1]) * 1000 + ko[:, 2]) * 4 + ko[:, 3]
    og, oo = np.argsort(cg), np.argsort(co)
    assert np.array_equal(cg[og], co[oo])
    assert np.abs(pcd.points.cpu().numpy()[og] - ref["points"][oo]).max() <= 1e-4 * 0.01 + 1e-6
    assert np.abs(pcd.normals.cpu().numpy()[og] - ref["normals"][oo]).max() <= 1e-4


@pytest.mark.parametrize("unit", [False, True])
def test_incremental_point_extraction_equals_full_extraction(cuda, unit):
    """row f1: `extract_pcd` after every frame (N/3DM/slam.py:126,195) in incremental mode -- only the bricks whose 3x3x3
    neighbourhood the integration changed are re-extracted, the rest comes from the per-brick cache -- gives exactly the
    arrays of a full extraction (points, normals, colours, keys, same order), frame after frame"""
    import copy
    from bodyslam_b200 import ops
    from bodyslam_b200.tsdf import DenseTSDFVolume
    res = 128 if unit else 96                     # 128: the scene's box already sits on the 32-voxel unit grid
    sc = small_scene("laparoscopy512", res=res, frame_ids=np.arange(0, 120, 10))
    origin = sc["origin"]
    vol = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], res, origin, color=True, device=cuda, unit_activation=unit)
    vol.set_incremental_points(True, normals=True)
    depth = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, cuda)
    col = torch.from_numpy(sc["color"]).to(cuda)
    recomputed = []
    for i in range(len(sc["E"])):
        vol.integrate_batch(depth[i:i + 1], col[i:i + 1], sc["intrinsic"], sc["E"][i:i + 1])
        inc = vol.extract_point_cloud()
        cand, rec = vol.points_last_stats()
        recomputed.append(rec / max(cand, 1))
        ref = copy.deepcopy(vol).extract_point_cloud()          # a fresh volume: full extraction
        assert inc.points.shape[0] == ref.points.shape[0] > 100, (i, inc.points.shape[0], ref.points.shape[0])
        for a, b in ((inc.points, ref.points), (inc.normals, ref.normals), (inc.colors, ref.colors), (inc.point_keys, ref.point_keys)):
            assert torch.equal(a, b)
    assert recomputed[0] == 1.0
    # nothing integrated since: (almost) nothing is recomputed -- only bricks too large for a cache slot
    again = vol.extract_point_cloud()
    cand, rec = vol.points_last_stats()
    assert rec <= 0.05 * cand and torch.equal(again.points, inc.points) and torch.equal(again.normals, inc.normals)
    # a one-off request without normals falls back to a full extraction and the cache starts over
    nn = vol.extract_point_cloud(normals=False)
    assert nn.normals is None and torch.equal(nn.points, inc.points)
    full_again = vol.extract_point_cloud()
    assert vol.points_last_stats()[1] == vol.points_last_stats()[0] and torch.equal(full_again.normals, inc.normals)
    # reset invalidates the cache
    vol.reset()
    assert vol.extract_point_cloud().points.shape[0] == 0
    vol.integrate_batch(depth[:1], col[:1], sc["intrinsic"], sc["E"][:1])
    first = vol.extract_point_cloud()
    ref = copy.deepcopy(vol).extract_point_cloud()
    assert torch.equal(first.points, ref.points) and torch.equal(first.normals, ref.normals)
