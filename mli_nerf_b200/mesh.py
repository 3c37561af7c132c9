"""Mesh-extraction lattice sweep: the block lattice of
/root/reference/projects/neuralangelo/utils/mesh.py:25-117 evaluated with the encode + SDF-trunk kernels.

The reference walks a `bounds / intv` lattice in blocks of (block_res+1)^3 points (neighbouring blocks share one
layer of points), builds every block on the host in a DataLoader worker, uploads it, calls `neural_sdf.sdf`, downloads
the values and hands them to `mcubes.marching_cubes`.  Here the lattice coordinates of a block are formed on the
device from the three axis vectors (same `torch.arange` values, so the points are bit-identical), the SDF comes from
`Model.sdf` (one encode + trunk launch pair per chunk) and only the [bx, by, bz] value block travels to the host.

`extract_mesh` keeps the reference's signature.  The iso-surface step itself stays what it is in the reference -- the
third-party `mcubes` / `trimesh` host libraries (not part of this image: they are imported on first use, or passed in
through `marching_cubes_fn` / `mesh_cls`); blocks are dealt to ranks round-robin and gathered with
`all_gather_object` like the reference does.
"""
import numpy as np
import torch
import torch.distributed as dist


class LatticeBlocks:
    """Block decomposition of the lattice (`LatticeGrid`, mesh.py:68-104): same axes, block order and block extents."""

    def __init__(self, bounds, intv, block_res=64):
        self.block_res = int(block_res)
        self.intv = float(intv)
        (x_min, x_max), (y_min, y_max), (z_min, z_max) = [(float(a), float(b)) for a, b in bounds]
        self.axes = [torch.arange(x_min, x_max, intv), torch.arange(y_min, y_max, intv), torch.arange(z_min, z_max, intv)]
        self.res = tuple(len(a) for a in self.axes)
        self.num_blocks = tuple(int(np.ceil(r / self.block_res)) for r in self.res)
        self._dev_axes = {}

    def __len__(self):
        return self.num_blocks[0] * self.num_blocks[1] * self.num_blocks[2]

    def block_start(self, idx):
        nby, nbz = self.num_blocks[1], self.num_blocks[2]
        return ((idx // (nby * nbz)) * self.block_res, ((idx // nbz) % nby) * self.block_res, (idx % nbz) * self.block_res)

    def xyz(self, idx, device="cpu"):
        """Points of block `idx`, [bx, by, bz, 3] with bx, by, bz <= block_res + 1 (mesh.py:82-98)."""
        if idx < 0 or idx >= len(self):
            raise IndexError(idx)
        key = str(device)
        if key not in self._dev_axes:
            self._dev_axes[key] = [a.to(device) for a in self.axes]
        ax = self._dev_axes[key]
        s = self.block_start(idx)
        x, y, z = torch.meshgrid(*[ax[d][s[d]:s[d] + self.block_res + 1] for d in range(3)], indexing="ij")
        return torch.stack([x, y, z], dim=-1)


def rank_blocks(n_blocks, rank, world):
    """Blocks of one rank: round-robin like `DistributedSampler(shuffle=False)` (mesh.py:107-110), without the wrapped
    padding duplicates that sampler appends when `world` does not divide `n_blocks` (they only add repeated blocks)."""
    return range(rank, n_blocks, world)


@torch.no_grad()
def sdf_blocks(sdf_func, bounds, intv, block_res=64, device="cuda", rank=0, world=1):
    """Yields (block index, xyz of the block's first point as numpy [3], sdf values as numpy [bx, by, bz]) for this
    rank's blocks (the loop body of extract_mesh, mesh.py:31-36, up to the marching-cubes call)."""
    lattice = LatticeBlocks(bounds, intv, block_res)
    for idx in rank_blocks(len(lattice), rank, world):
        xyz = lattice.xyz(idx, device)
        sdf = sdf_func(xyz)[..., 0]
        yield idx, xyz[0, 0, 0].cpu().numpy(), sdf.cpu().numpy()


def _default_marching_cubes():
    try:
        import mcubes
    except ImportError as e:  # pragma: no cover
        raise ImportError("extract_mesh needs the `mcubes` package for the iso-surface step (as the reference does), "
                          "or pass marching_cubes_fn=") from e
    return mcubes.marching_cubes


def _default_mesh_cls():
    try:
        import trimesh
    except ImportError as e:  # pragma: no cover
        raise ImportError("extract_mesh needs the `trimesh` package to assemble the mesh (as the reference does), "
                          "or pass mesh_cls= / use sdf_blocks()") from e
    return trimesh


def _block_mesh(V, F, xyz0, intv, texture_func, filter_lcc, tm):
    """mesh.py:119-133: lattice units -> world, optional vertex colours, unit-sphere filter, largest component."""
    if V.shape[0] == 0:
        return tm.Trimesh()
    V = V * intv + xyz0
    mesh = tm.Trimesh(V, F, vertex_colors=texture_func(V)) if texture_func is not None else tm.Trimesh(V, F)
    mask = np.linalg.norm(mesh.vertices, axis=-1) < 1.0  # mesh.py:136-149
    if not np.any(mask):
        return tm.Trimesh()
    indices = np.full(len(mesh.vertices), -1, dtype=int)
    indices[mask] = np.arange(mask.sum())
    faces = np.asarray(mesh.faces)
    keep = mask[faces[:, 0]] & mask[faces[:, 1]] & mask[faces[:, 2]]
    mesh = tm.Trimesh(np.asarray(mesh.vertices)[mask], indices[faces[keep]],
                      vertex_colors=np.asarray(mesh.visual.vertex_colors)[mask])
    if filter_lcc:  # mesh.py:152-159
        comps = mesh.split(only_watertight=False)
        areas = np.array([c.area for c in comps], dtype=float)
        mesh = comps[areas.argmax()] if len(areas) > 0 and mesh.vertices.shape[0] > 0 else tm.Trimesh()
    return mesh


@torch.no_grad()
def extract_mesh(sdf_func, bounds, intv, block_res=64, texture_func=None, filter_lcc=False, device="cuda",
                 marching_cubes_fn=None, mesh_cls=None):
    """`extract_mesh` of mesh.py:25-49.  `sdf_func` is what the script builds (`lambda x: -model.sdf(x)`,
    scripts/extract_mesh.py:101).  Returns the concatenated mesh on rank 0 and None elsewhere."""
    mc = marching_cubes_fn or _default_marching_cubes()
    tm = mesh_cls or _default_mesh_cls()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    blocks = []
    for _, xyz0, sdf in sdf_blocks(sdf_func, bounds, intv, block_res, device, rank, world):
        V, F = mc(sdf, 0.)
        blocks.append(_block_mesh(np.asarray(V), np.asarray(F), xyz0, intv, texture_func, filter_lcc, tm))
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, blocks)
    else:
        gathered = [blocks]
    if rank != 0:
        return None
    every = [m for bl in gathered for m in bl if m.vertices.shape[0] > 0]
    return tm.util.concatenate(every)
