#!/bin/bash
# compute-sanitizer over the hot path (run on a GPU box):  bash tools/sanitize.sh > gpurun_out/sanitizer.log
# memcheck + racecheck + synccheck on smoke() (a4, K3 with named barriers / shared flag words / persistent
# claims, K4 marching cubes, K1, K2) and on a small sharded-layout / unit-activation / colour integration.
set -u
CS=/usr/local/cuda/bin/compute-sanitizer
PY='
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
g.smoke()
from util import small_scene
from bodyslam_b200 import ops, mdem
from bodyslam_b200.tsdf import DenseTSDFVolume
dev = torch.device("cuda", 0)
sc = small_scene("laparoscopy512", res=64, frames=3, W=320, H=240)
d = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, dev)
col = torch.from_numpy(sc["color"]).to(dev)
ul = sc["voxel_length"] * 32
org = np.floor(sc["origin"] / ul + 0.5) * ul
for kw in (dict(unit_activation=True), dict(gz0=8, z_total=64, z_interleave=2), dict()):
    for zpw in (8, 4, 2):
        res = (64, 64, 32) if "gz0" in kw else 64
        v = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], res, org, color=True, device=dev, **kw)
        v.set_z_split(zpw)
        v.integrate_batch(d, col, sc["intrinsic"], sc["E"])
        if "gz0" not in kw:
            v.extract_triangle_mesh(); v.extract_point_cloud()
mdem.colorize(np.random.default_rng(0).uniform(0.1, 3, (120, 160)).astype(np.float32), cmap="viridis", invalid_val=0)
torch.cuda.synchronize()
print("SANITIZE_WORKLOAD_OK")
'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 $CS --tool $tool --print-limit 20 python -c "$PY" 2>&1 | grep -E "SANITIZE_WORKLOAD_OK|smoke ok|ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard|error" | head -40
done
