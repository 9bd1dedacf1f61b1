"""PLY writers for the 3DM outputs (`o3d.io.write_point_cloud` / `write_triangle_mesh` of
N/3DM/tsdf.py:37,52).

Byte layout restated from Open3D's PLY writers (FilePLY.cpp: WritePointCloudToPLY / WriteTriangleMeshToPLY
through rply, default `write_ascii=False`), so that a file written here equals the one the reference would
have written for the same geometry:
  header   "ply" / "format binary_little_endian 1.0" / "comment Created by Open3D" / "element vertex N" /
           "property double x|y|z" [/ "property double nx|ny|nz" when normals exist]
           [/ "property uchar red|green|blue" when colours exist]
           [/ "element face M" / "property list uchar uint vertex_indices"] / "end_header", '\n' line ends;
  vertices packed little-endian records in that property order; colours = utility::ColorToUint8 =
           round(clamp(c, 0, 1) * 255);
  faces    uchar 3 + three uint32 indices.
Open3D is not installable here, so the layout is pinned by tests/test_capi_and_host.py against this
restatement (header bytes and record offsets), not against a file Open3D wrote."""
from __future__ import annotations

import numpy as np

from .geometry import to_numpy


def _ply_header(n_vert, props, n_face=None):
    lines = ["ply", "format binary_little_endian 1.0", "comment Created by Open3D", f"element vertex {n_vert}"]
    lines += [f"property {t} {n}" for t, n in props]
    if n_face is not None:
        lines += [f"element face {n_face}", "property list uchar uint vertex_indices"]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def _vertex_block(xyz, normals=None, colors=None):
    cols = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    props = [("double", "x"), ("double", "y"), ("double", "z")]
    if normals is not None:
        cols += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
        props += [("double", "nx"), ("double", "ny"), ("double", "nz")]
    if colors is not None:
        cols += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        props += [("uchar", "red"), ("uchar", "green"), ("uchar", "blue")]
    rec = np.empty(len(xyz), dtype=cols)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if normals is not None:
        rec["nx"], rec["ny"], rec["nz"] = normals[:, 0], normals[:, 1], normals[:, 2]
    if colors is not None:
        c8 = np.floor(np.clip(np.asarray(colors, dtype=np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)   # utility::ColorToUint8 (std::round)
        rec["red"], rec["green"], rec["blue"] = c8[:, 0], c8[:, 1], c8[:, 2]
    return rec, props


def write_point_cloud(path: str, pcd) -> bool:
    xyz = to_numpy(pcd.points).astype(np.float64)
    nrm = None if getattr(pcd, "normals", None) is None else to_numpy(pcd.normals).astype(np.float64)
    col = None if getattr(pcd, "colors", None) is None else to_numpy(pcd.colors)
    rec, props = _vertex_block(xyz, nrm, col)
    with open(path, "wb") as f:
        f.write(_ply_header(len(xyz), props))
        f.write(rec.tobytes())
    return True


def write_triangle_mesh(path: str, mesh) -> bool:
    xyz = to_numpy(mesh.vertices).astype(np.float64)
    col = None if getattr(mesh, "vertex_colors", None) is None else to_numpy(mesh.vertex_colors)
    tri = to_numpy(mesh.triangles).astype("<u4")
    rec, props = _vertex_block(xyz, None, col)
    faces = np.empty(len(tri), dtype=[("n", "u1"), ("v", "<u4", (3,))])
    faces["n"] = 3
    faces["v"] = tri
    with open(path, "wb") as f:
        f.write(_ply_header(len(xyz), props, len(tri)))
        f.write(rec.tobytes())
        f.write(faces.tobytes())
    return True


def read_ply(path: str):
    """Minimal reader for the files written above (tests / round trips)."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply"
        fmt = f.readline().split()[1]
        assert fmt == b"binary_little_endian"
        nv = nf = 0
        props = []
        in_vertex = False
        while True:
            line = f.readline().strip()
            if line == b"end_header":
                break
            tok = line.split()
            if tok[0] == b"element":
                in_vertex = tok[1] == b"vertex"
                if in_vertex:
                    nv = int(tok[2])
                elif tok[1] == b"face":
                    nf = int(tok[2])
            elif tok[0] == b"property" and in_vertex:
                props.append((tok[2].decode(), {"double": "<f8", "float": "<f4", "uchar": "u1"}[tok[1].decode()]))
        verts = np.frombuffer(f.read(nv * np.dtype(props).itemsize), dtype=props, count=nv)
        faces = None
        if nf:
            fd = np.dtype([("n", "u1"), ("v", "<u4", (3,))])
            faces = np.frombuffer(f.read(nf * fd.itemsize), dtype=fd, count=nf)["v"]
    return verts, faces
