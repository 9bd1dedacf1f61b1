"""CUDA path against the COMMITTED oracle pins (tests/golden/oracle_pins.json) -- `-m gpu`.
Same seeded scenes as tests/golden/make_oracle_pins.py; sha256 of the exported tsdf / weight grids,
per-frame update counts and mesh sizes must equal the pinned ones bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from bodyslam_b200 import ops
from bodyslam_b200.tsdf import DenseTSDFVolume
from util import small_scene

pytestmark = pytest.mark.gpu


def digest(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()


@pytest.mark.parametrize("scene", ["laparoscopy512", "colonoscopy256"])
@pytest.mark.parametrize("mode", ["dense_z8", "scalable"])
def test_cuda_volume_equals_committed_pins(cuda, golden_dir, scene, mode):
    pins = json.load(open(os.path.join(golden_dir, "oracle_pins.json")))[f"{scene}/64/{mode}"]
    frames = len(pins["counts"])
    sc = small_scene(scene, res=64, frames=frames)
    origin = sc["origin"]
    if mode == "scalable":
        ul = sc["voxel_length"] * 32
        origin = np.floor(origin / ul + 0.5) * ul
    vol = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 64, origin, color=True, device=cuda, unit_activation=(mode == "scalable"))
    depth = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, cuda)
    assert digest(depth[0]) == pins["depth0"]
    counts = torch.zeros(frames, dtype=torch.int64, device=cuda)
    vol.integrate_batch(depth, torch.from_numpy(sc["color"]).to(cuda), sc["intrinsic"], sc["E"], update_counts=counts)
    t, w = vol.export_dense()
    assert counts.cpu().tolist() == pins["counts"]
    assert digest(w) == pins["weight"] and digest(t) == pins["tsdf"]
    mesh = vol.extract_triangle_mesh()
    assert int(mesh.vertices.shape[0]) == pins["vertices"] and int(mesh.triangles.shape[0]) == pins["triangles"]
