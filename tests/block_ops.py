"""Index-space transformations of a block (tests only): the same physical grid, state and wall distance seen through rotated
(i, j, k) axes, so that two blocks of one duct meet with reversed index ranges / swapped transverse axes -- the orientations
the unpack maps Pilo..PkDir and dir_switch describe (reference: src/mapping.f90:185-258, src/interface1.f90:144-168)."""
import importlib

import numpy as np


def _apply(a, op):
    """a[..., k, j, i] cell / node array (leading axes free).  op: 'rot180' (j, k) -> (J-1-j, K-1-k);
    'rot90': new j' = old k, new k' = reversed old j (a right-handed rotation about the i axis)."""
    if op == "rot180":
        return np.ascontiguousarray(a[..., ::-1, ::-1, :])
    if op == "rot90":
        t = np.swapaxes(a, -3, -2)          # t[..., a, b, i] = old[..., k=b, j=a, i]
        return np.ascontiguousarray(t[..., ::-1, :, :])   # new[k', j', i] = old[k=j', j=J-1-k', i]
    raise ValueError(op)


def rotate_block(blk, op):
    """In place.  Physical vectors (velocity, node coordinates) keep their components: only indices move."""
    nodes = np.moveaxis(blk.nodes, -1, 0)                # [3, k, j, i]
    blk.nodes = np.ascontiguousarray(np.moveaxis(_apply(nodes, op), 0, -1))
    blk.qp = _apply(blk.qp, op)
    if blk.dist is not None:
        blk.dist = _apply(blk.dist, op)
    b = list(blk.bc_id)
    if op == "rot180":
        blk.bc_id = [b[0], b[1], b[3], b[2], b[5], b[4]]
    else:   # jmin' = old kmin, jmax' = old kmax, kmin' = old jmax, kmax' = old jmin
        blk.jmx, blk.kmx = blk.kmx, blk.jmx
        blk.bc_id = [b[0], b[1], b[4], b[5], b[3], b[2]]
    blk.default_maps()
    blk.build_geometry()
    return blk


def unrotate_cells(a, op):
    """Inverse of the index transformation for an interior cell array a[..., k', j', i]."""
    if op == "rot180":
        return np.ascontiguousarray(a[..., ::-1, ::-1, :])
    if op == "rot90":     # new[k', j', i] = old[k=j', j=J-1-k', i]  ->  old[k, j, i] = new[k'=J-1-j, j'=k, i]
        return np.ascontiguousarray(np.swapaxes(a[..., ::-1, :, :], -3, -2))
    raise ValueError(op)


def two_block_duct_with_rotated_neighbour(blocks, op):
    """blocks: the two blocks of a (2,1,1) duct (block 0 imax <-> block 1 imin).  Block 1 is re-stored through rotated axes and
    both unpack maps are set to what mapping.f90:185-258 derives for that orientation."""
    a, b = blocks
    nj, nk = a.jmx - 1, a.kmx - 1            # cells of block 0 across the interface
    rotate_block(b, op)
    if op == "rot180":                        # both transverse ranges run backwards on both sides
        a.plo[1], a.phi[1], a.pdir[1] = [nj, nk], [1, 1], [-1, -1]
        b.plo[0], b.phi[0], b.pdir[0] = [nj, nk], [1, 1], [-1, -1]
    else:                                     # j' = k, k' = reversed j: outer / inner loops swap (dir_switch = 1)
        a.plo[1], a.phi[1], a.pdir[1] = [nj, 1], [1, nk], [-1, 1]
        b.plo[0], b.phi[0], b.pdir[0] = [1, nj], [nk, 1], [1, -1]
        a.dir_switch[1] = 1
        b.dir_switch[0] = 1
    return blocks
