"""Mesh-extraction lattice driver (mli_nerf_b200/mesh.py) against the block lattice of the reference
(/root/reference/projects/neuralangelo/utils/mesh.py:68-104; that module imports mcubes/trimesh, which this image does
not have, so the expected values are restated here from its arithmetic: blocks of block_res+1 points along each axis of
torch.arange(min, max, intv), block index = (bx * nby + by) * nbz + bz)."""
import numpy as np
import pytest
import torch

from mli_nerf_b200 import mesh as M


def _full_lattice(bounds, intv):
    ax = [torch.arange(float(a), float(b), intv) for a, b in bounds]
    return torch.stack(torch.meshgrid(*ax, indexing="ij"), dim=-1)


@pytest.mark.parametrize("bounds,intv,block_res", [
    ([[-1.0, 1.0]] * 3, 2.0 / 16, 8), ([[-1.0, 1.0]] * 3, 2.0 / 20, 8),
    ([[-0.3, 0.3], [-0.21, 0.21], [-0.18, 0.15]], 0.05, 4), ([[-1.0, 1.0]] * 3, 2.0 / 7, 64)])
def test_blocks_tile_the_lattice(bounds, intv, block_res):
    full = _full_lattice(bounds, intv)
    lat = M.LatticeBlocks(bounds, intv, block_res)
    assert lat.res == tuple(full.shape[:3])
    assert len(lat) == int(np.prod([int(np.ceil(r / block_res)) for r in full.shape[:3]]))
    seen = torch.zeros(full.shape[:3], dtype=torch.int32)
    for idx in range(len(lat)):
        s = lat.block_start(idx)
        xyz = lat.xyz(idx)
        sl = tuple(slice(s[d], s[d] + block_res + 1) for d in range(3))
        assert torch.equal(xyz, full[sl])  # bit-identical coordinates, neighbouring blocks share one layer of points
        assert all(1 <= n <= block_res + 1 for n in xyz.shape[:3])
        seen[sl] += 1
    assert int(seen.min()) >= 1
    with pytest.raises(IndexError):
        lat.xyz(len(lat))


def test_rank_blocks_partition():
    for n, w in ((27, 1), (27, 2), (27, 8), (5, 8)):
        parts = [list(M.rank_blocks(n, r, w)) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))


class _StubMesh:
    """The handful of trimesh.Trimesh members the driver touches."""

    class _Visual:
        def __init__(self, c):
            self.vertex_colors = c

    def __init__(self, vertices=None, faces=None, vertex_colors=None):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.asarray(vertices, dtype=np.float64)
        self.faces = np.zeros((0, 3), dtype=int) if faces is None else np.asarray(faces)
        n = len(self.vertices)
        self.visual = self._Visual(np.full((n, 4), 255, np.uint8) if vertex_colors is None else np.asarray(vertex_colors))


class _StubTrimesh:
    Trimesh = _StubMesh

    class util:
        @staticmethod
        def concatenate(meshes):
            if not meshes:
                return _StubMesh()
            off, V, F = 0, [], []
            for m in meshes:
                V.append(m.vertices)
                F.append(m.faces + off)
                off += len(m.vertices)
            return _StubMesh(np.concatenate(V), np.concatenate(F))


def _edge_crossings(sdf, level):
    """Stand-in for mcubes.marching_cubes: one vertex per x-edge that crosses `level` (linear interpolation, lattice
    units) and a degenerate face per vertex -- enough to check offsets, scaling, filtering and concatenation."""
    a, b = sdf[:-1], sdf[1:]
    i, j, k = np.nonzero((a < level) != (b < level))
    t = (level - a[i, j, k]) / (b[i, j, k] - a[i, j, k])
    V = np.stack([i + t, j.astype(np.float64), k.astype(np.float64)], axis=-1)
    return V, np.repeat(np.arange(len(V))[:, None], 3, axis=1)


def test_extract_mesh_driver_on_analytic_sphere():
    radius = 0.6
    sdf_func = lambda x: (x.norm(dim=-1, keepdim=True) - radius)  # noqa: E731
    res = 24
    mesh = M.extract_mesh(sdf_func, [[-1.0, 1.0]] * 3, 2.0 / res, block_res=8, device="cpu",
                          marching_cubes_fn=_edge_crossings, mesh_cls=_StubTrimesh)
    V = mesh.vertices
    assert len(V) > 100 and mesh.faces.max() == len(V) - 1
    # every vertex sits on the sphere to within the linear-interpolation error of one lattice edge
    assert np.abs(np.linalg.norm(V, axis=-1) - radius).max() < (2.0 / res) ** 2
    # a surface outside the unit sphere is filtered away block by block (mesh.py:136-149)
    far = M.extract_mesh(lambda x: (x.norm(dim=-1, keepdim=True) - 1.2), [[-1.5, 1.5]] * 3, 3.0 / res, block_res=8,
                         device="cpu", marching_cubes_fn=_edge_crossings, mesh_cls=_StubTrimesh)
    assert far.vertices.shape[0] == 0


def test_sdf_blocks_reassemble_to_one_shot_query():
    f = lambda x: torch.sin(3 * x[..., :1]) + x[..., 1:2] * x[..., 2:3]  # noqa: E731
    bounds, intv, br = [[-1.0, 1.0]] * 3, 2.0 / 10, 4
    full = _full_lattice(bounds, intv)
    want = f(full)[..., 0].numpy()
    got = np.full(want.shape, np.nan, dtype=np.float32)
    lat = M.LatticeBlocks(bounds, intv, br)
    for idx, xyz0, sdf in M.sdf_blocks(f, bounds, intv, br, device="cpu"):
        s = lat.block_start(idx)
        assert np.array_equal(xyz0, full[s].numpy())
        got[s[0]:s[0] + sdf.shape[0], s[1]:s[1] + sdf.shape[1], s[2]:s[2] + sdf.shape[2]] = sdf
    assert np.array_equal(got, want)
