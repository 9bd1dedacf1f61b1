"""Structural validation of the marching-cubes tables (oracle copy and CUDA copy), CPU only."""
import os
import re

import numpy as np
import pytest

import oracle
from util import mesh_is_closed_and_oriented

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E2V = [(0, 1), (1, 2), (3, 2), (0, 3), (4, 5), (5, 6), (7, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
SHIFT = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]


def parse(path, start, end):
    src = open(path).read()
    body = src[src.index(start):src.index(end, src.index(start))]
    body = body[body.index("{") + 1:]
    nums = [int(t) for t in re.findall(r"-?\d+", body.replace("ORC_X", "-1"))]
    return np.array(nums[:4096]).reshape(256, 16)


@pytest.fixture(scope="module")
def tables():
    a = parse(os.path.join(ROOT, "oracle", "mc_tables.h"), "orc_tri_table[256][16]", "#undef ORC_X")
    b = parse(os.path.join(ROOT, "bodyslam_b200", "csrc", "mc_tables.cuh"), "kTriTable[256 * 16]", "kNumTris")
    return a, b


def faces(e):
    pa, pb = SHIFT[E2V[e][0]], SHIFT[E2V[e][1]]
    return {(ax, pa[ax]) for ax in range(3) if pa[ax] == pb[ax]}


def test_cuda_and_oracle_tables_are_identical(tables):
    a, b = tables
    assert np.array_equal(a, b)
    src = open(os.path.join(ROOT, "bodyslam_b200", "csrc", "mc_tables.cuh")).read()
    nt = [int(t) for t in re.findall(r"\d+", src[src.index("kNumTris[256]"):src.index("kEdgeShift")])[1:257]]
    assert nt == [int((row != -1).sum()) // 3 for row in a]
    assert sum(nt) == 820


def test_every_case_uses_exactly_the_sign_change_edges_and_is_a_face_bounded_manifold(tables):
    T, _ = tables
    for ci, row in enumerate(T):
        n = int((row != -1).sum())
        assert n % 3 == 0 and (row[n:] == -1).all()
        used = set(row[:n].tolist())
        sign_change = {e for e, (a, b) in enumerate(E2V) if ((ci >> a) & 1) != ((ci >> b) & 1)}
        assert used == sign_change, ci
        edges = {}
        for t in range(0, n, 3):
            a, b, c = row[t:t + 3]
            assert len({a, b, c}) == 3
            for u, v in ((a, b), (b, c), (c, a)):
                edges.setdefault((min(u, v), max(u, v)), []).append((u, v))
        for k, lst in edges.items():
            if len(lst) == 1:
                assert faces(k[0]) & faces(k[1]), (ci, k)       # open edges lie on a cube face
            else:
                assert len(lst) == 2 and lst[0] != lst[1], (ci, k)  # interior edges: twice, opposite directions


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_meshes_of_closed_surfaces_are_watertight(seed):
    """random smooth closed level sets: every mesh edge is shared by exactly two triangles with
    opposite orientation -- fails for almost any typo in the 256x16 table"""
    n = 40
    g = (np.arange(n) + 0.5) / n - 0.5
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    rng = np.random.default_rng(seed)
    t = np.sqrt(x * x + y * y + z * z) - 0.28
    for _ in range(10):
        k = rng.normal(size=3) * 18
        t += 0.035 * rng.normal() * np.sin(k[0] * x + k[1] * y + k[2] * z + rng.uniform(0, 6.28))
    t[[0, -1], :, :] = t[:, [0, -1], :] = t[:, :, [0, -1]] = 1.0   # keep the level set off the box boundary
    V = oracle.o3d.Volume(n, 0.01, 0.04)
    V.tsdf[:] = np.clip(t * 10, -1, 1).astype(np.float32).reshape(-1)
    V.weight[:] = 1.0
    m = V.extract_mesh()
    assert len(m["triangles"]) > 3000
    assert mesh_is_closed_and_oriented(m["triangles"])
    # every vertex lies on its voxel edge
    k = m["keys"]
    lo = (k[:, :3] + 0.5) * 0.01
    d = m["vertices"] - lo
    ax = k[:, 3]
    assert np.all(d[np.arange(len(d)), ax] >= -1e-12) and np.all(d[np.arange(len(d)), ax] <= 0.01 + 1e-12)
    d[np.arange(len(d)), ax] = 0
    assert np.abs(d).max() < 1e-12
