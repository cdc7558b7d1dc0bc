"""Host side of `from_amr`: the integer tables (reference fava/mesh/FLASH/_flash.py:963-1022 and
:1157-1199), restated with the reference's arithmetic so that every index is bit-identical — int32
truncation of the block corner ids, `0.5 + ...` rounding of the sub-domain corners, the
`any(0 not in row)` sub-domain rule, inclusive intersection test — and their translation into the
`fava_prolong_leaf` table the gather kernel consumes."""

from __future__ import annotations

import numpy as np

MESH_MDIM = 3


class AmrPlan:
    def __init__(self):
        self.subdomain_flag = False
        self.ref_lev_max = 0
        self.grid_delta = None  # float64 (3,1)
        self.local_BCIDs = None  # int32 (nblocks,3,2)
        self.subdomain_BCIDs = None  # int32 (3,2)
        self.leaf_IDs = np.zeros(0, dtype=np.int64)
        self.scales = np.zeros(0, dtype=np.int64)
        self.total_cells = None  # int32 (3,)  NX,NY,NZ
        self.refdom_bound_box = None

    def slab_leaves(self, z0: int, z1: int, nzb: int):
        """(block ids, fine-cell corners relative to the output, scales) of the selected leaves that touch
        output planes [z0, z1), in list order (later entries win, like the reference's dict)."""
        ids = self.leaf_IDs
        off = self.local_BCIDs[ids, :, 0].astype(np.int64)
        if self.subdomain_flag:
            off = off - self.subdomain_BCIDs[None, :, 0]
        keep = (off[:, 2] < z1) & (off[:, 2] + nzb * self.scales > z0)
        return ids[keep], off[keep], self.scales[keep]


def build_plan(mesh, subdomain_coords: np.ndarray, refine_level: int) -> AmrPlan | None:
    p = AmrPlan()
    ndim = int(mesh.ndim)
    p.subdomain_flag = any(0 not in row for row in subdomain_coords)  # :965
    if p.subdomain_flag:  # :967-977
        b = mesh.domain_bounds
        for a in range(ndim):
            if subdomain_coords[a, 0] < b[a, 0] or b[a, 1] < subdomain_coords[a, 1]:
                return None
    ref_lev_max = int(mesh.refine_level_max)  # allreduce(MAX) of the same global table (:985)
    ref_lev = min(refine_level, ref_lev_max)  # :995
    if ref_lev > 0:
        ref_lev_max = ref_lev
    p.ref_lev_max = ref_lev_max

    bb = mesh.block_bounds
    gbb = np.zeros_like(bb[0])  # keeps the file dtype (f32 for plt), :1000-1002
    gbb[:, 0] = np.min(bb[..., 0], axis=0)
    gbb[:, 1] = np.max(bb[..., 1], axis=0)
    cellfac = 2 ** (ref_lev_max - 1)
    grid_delta = (np.diff(gbb, axis=1).flatten() / (mesh.nCellsVec * mesh.nBlksVec * cellfac))[:, None]  # :1005-1007
    half = grid_delta * 0.5
    # int32 assignment truncates toward zero (:1010-1015); vectorised over blocks
    local = ((bb - gbb[None, :, 0, None] + half[None, ...]) / grid_delta[None, ...]).astype(np.int32)
    sub = np.zeros((MESH_MDIM, 2), dtype=np.int32)
    if p.subdomain_flag:  # :1017-1022
        sub[:, :] = (0.5 + (subdomain_coords[:MESH_MDIM, :] - gbb[:MESH_MDIM, :1]) / grid_delta[:MESH_MDIM, :]).astype(np.int32)
    max_scale = int(2 ** (ref_lev_max - 1))
    fine_blks = max_scale * np.array([mesh.nblockx, mesh.nblocky, mesh.nblockz], dtype=np.int32)
    subd_cells = np.ones_like(fine_blks)
    if p.subdomain_flag:
        subd_cells[:ndim] = np.diff(sub[:ndim, :]).flatten()  # :1033-1034
    local[:, ndim:MESH_MDIM, 1] = 0  # :1159 / :1175

    level = mesh.refine_level
    leaf = mesh.node_type == 1
    if ref_lev > -1:  # :1160-1172
        maybe = (leaf & (level < ref_lev)) | (level == ref_lev)
    else:  # :1176-1182
        maybe = leaf
    if p.subdomain_flag:  # _intersects_subdomain :1386-1393 (inclusive on both ends)
        hit = np.all((sub[None, :, 0] <= local[:, :, 1]) & (local[:, :, 0] <= sub[None, :, 1]), axis=1)
        maybe = maybe & hit
    p.leaf_IDs = np.flatnonzero(maybe).astype(np.int64)
    p.scales = (2 ** (ref_lev_max - level[p.leaf_IDs])).astype(np.int64)  # :1270-1271
    if p.scales.size and p.scales.min() < 1:
        raise RuntimeError("from_amr: a selected block is finer than the target level")
    if p.subdomain_flag:
        p.refdom_bound_box = gbb[:, :1] + sub * grid_delta  # :1185
        p.total_cells = np.copy(subd_cells)  # :1191
    else:
        p.refdom_bound_box = np.copy(gbb)  # :1188
        p.total_cells = np.ones_like(fine_blks)
        p.total_cells[:ndim] = fine_blks[:ndim] * mesh.nCellsVec[:ndim]  # :1193-1194
    if np.any(p.total_cells < 1):
        raise RuntimeError("BAD!")  # the reference's only diagnostic for an empty target (:1044, :1082, :1128)
    p.grid_delta = grid_delta
    p.local_BCIDs = local
    p.subdomain_BCIDs = sub
    return p
