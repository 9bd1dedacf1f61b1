"""Pins of the CPU oracle's own output on seeded synthetic scenes (NOT reference vectors: Open3D is not
available here).  They guard the oracle -- the thing every GPU parity test compares
against -- from drifting unnoticed: sha256 of the float32 tsdf / weight grids, update counts and mesh sizes.

    python tests/golden/make_oracle_pins.py        # rewrites tests/golden/oracle_pins.json
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute():
    import oracle
    from util import small_scene

    out = {}
    for scene, res, frames in (("laparoscopy512", 64, 4), ("colonoscopy256", 64, 3)):
        sc = small_scene(scene, res=res, frames=frames)
        for mode in ("dense_z8", "literal", "scalable"):
            if mode == "scalable":
                ul = sc["voxel_length"] * 32
                origin = np.floor(sc["origin"] / ul + 0.5) * ul
            else:
                origin = sc["origin"]
            V = oracle.o3d.Volume(res, sc["voxel_length"], sc["sdf_trunc"], origin, with_color=True)
            counts = []
            for i in range(frames):
                d = oracle.o3d.depth_from_u16(sc["depth_u16"][i])
                if mode == "scalable":
                    counts.append(int(V.integrate_scalable(d, sc["K"], sc["E"][i], rgb=sc["color"][i])))
                else:
                    counts.append(int(V.integrate(d, sc["K"], sc["E"][i], rgb=sc["color"][i], z_restart=8 if mode == "dense_z8" else 0)))
            m = V.extract_mesh()
            out[f"{scene}/{res}/{mode}"] = {"counts": counts, "tsdf": digest(V.tsdf), "weight": digest(V.weight), "color": digest(V.color),
                                           "vertices": int(len(m["vertices"])), "triangles": int(len(m["triangles"])),
                                           "depth0": digest(oracle.o3d.depth_from_u16(sc["depth_u16"][0]))}
    return out


if __name__ == "__main__":
    pins = compute()
    with open(os.path.join(HERE, "oracle_pins.json"), "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)
    print(json.dumps({k: v["counts"] for k, v in pins.items()}, indent=1))
