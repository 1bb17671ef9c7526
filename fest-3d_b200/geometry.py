"""Host-side grid and metric set-up that feeds the drop-in boundary.

This is set-up code (run once, on the host) -- it is NOT on the accelerated path.  It produces the
arrays the reference's driver owns and passes by reference into ``get_next_solution``
(``cells``, ``Ifaces``, ``Jfaces``, ``Kfaces``; reference: src/solver.f90:32-63) in the reference's own
Fortran memory layout, so the same buffers can be handed to the C ABI (``fest3d_gpu_set_geometry``).

Reference behaviour followed:
  * ghost nodes by linear extrapolation, i then j then k      -- src/grid.f90:137-236
  * face area vectors, areas, unit normals                    -- src/geometry.f90:236-314, 65-99, 216-223
  * 5-tetrahedra hexahedron volumes on cells 0..imx, else 1.0 -- src/geometry.f90:316-496
  * cell centres as the mean of the 8 nodes                   -- src/geometry.f90:500-545

Array convention used throughout the python host: numpy arrays indexed ``[k, j, i(, c)]`` in C order,
which is byte-identical to the Fortran ``(i, j, k)`` arrays with an AoS record of ``c`` doubles.
Index ``0`` of every axis is the Fortran lower bound ``-2``.
"""
from __future__ import annotations

import numpy as np

G = 3  # ghost layers (interface1.f90:23); cell (i) lives at python index i+2, node (i) at i+2


def read_grid(path):
    """ASCII grid: first line ``imx jmx kmx`` then ``x y z`` per node, i fastest (grid.f90:78-133)."""
    with open(path) as f:
        imx, jmx, kmx = (int(t) for t in f.readline().split()[:3])
        data = np.loadtxt(f, dtype=np.float64)
    data = data.reshape(kmx, jmx, imx, 3)
    return data


def ghost_grid(nodes_int):
    """nodes_int[kmx, jmx, imx, 3] -> nodes[-2:kmx+3, -2:jmx+3, -2:imx+3, 3] (grid.f90:137-236)."""
    kmx, jmx, imx, _ = nodes_int.shape
    n = np.zeros((kmx + 6, jmx + 6, imx + 6, 3))
    n[3:3 + kmx, 3:3 + jmx, 3:3 + imx] = nodes_int
    o = 2  # python index of Fortran index 0

    def ext(a, axis, mx):
        s = [slice(None)] * 4

        def at(idx):
            t = list(s)
            t[axis] = idx + o
            return tuple(t)
        a[at(0)] = 2 * a[at(1)] - a[at(2)]
        a[at(-1)] = 2 * a[at(0)] - a[at(1)]
        a[at(-2)] = 2 * a[at(-1)] - a[at(0)]
        a[at(mx + 1)] = 2 * a[at(mx)] - a[at(mx - 1)]
        a[at(mx + 2)] = 2 * a[at(mx + 1)] - a[at(mx)]
        a[at(mx + 3)] = 2 * a[at(mx + 2)] - a[at(mx + 1)]

    ext(n, 2, imx)
    ext(n, 1, jmx)
    ext(n, 0, kmx)
    return n


def _cross_half(d1, d2):
    out = np.empty_like(d1)
    out[..., 0] = 0.5 * (d1[..., 1] * d2[..., 2] - d1[..., 2] * d2[..., 1])
    out[..., 1] = 0.5 * (d1[..., 2] * d2[..., 0] - d1[..., 0] * d2[..., 2])
    out[..., 2] = 0.5 * (d1[..., 0] * d2[..., 1] - d1[..., 1] * d2[..., 0])
    return out


def _finish_faces(vec, bc_ids, axis, mx):
    """area = |vec|, n = vec/area where area != 0; pole (-7) faces get A = 0 and copied normals."""
    A = np.sqrt(vec[..., 0] ** 2 + vec[..., 1] ** 2 + vec[..., 2] ** 2)
    o = 2
    lo_id, hi_id = bc_ids

    def sl(a, b):
        s = [slice(None)] * 3
        s[axis] = slice(a + o, b + o + 1)
        return tuple(s)
    if lo_id == -7:
        A[sl(-2, 1)] = 0.0
    if hi_id == -7:
        A[sl(mx, mx + 3)] = 0.0
    nz = A != 0.0
    nrm = vec.copy()
    nrm[nz] = vec[nz] / A[nz][:, None]
    if lo_id == -7:
        src = [slice(None)] * 3
        src[axis] = slice(2 + o, 2 + o + 1)
        for idx in (1, 0, -1, -2):
            nrm[sl(idx, idx)] = nrm[tuple(src)]
    if hi_id == -7:
        src = [slice(None)] * 3
        src[axis] = slice(mx - 1 + o, mx - 1 + o + 1)
        for idx in (mx, mx + 1, mx + 2, mx + 3):
            nrm[sl(idx, idx)] = nrm[tuple(src)]
    out = np.empty(vec.shape[:3] + (4,))
    out[..., 0] = A
    out[..., 1:] = nrm
    return out


def _tet(p1, p2, p3, p4):
    v = ((p4[..., 0] - p1[..., 0]) * ((p2[..., 1] - p1[..., 1]) * (p3[..., 2] - p1[..., 2]) - (p2[..., 2] - p1[..., 2]) * (p3[..., 1] - p1[..., 1]))
         + (p4[..., 1] - p1[..., 1]) * ((p2[..., 2] - p1[..., 2]) * (p3[..., 0] - p1[..., 0]) - (p2[..., 0] - p1[..., 0]) * (p3[..., 2] - p1[..., 2]))
         + (p4[..., 2] - p1[..., 2]) * ((p2[..., 0] - p1[..., 0]) * (p3[..., 1] - p1[..., 1]) - (p2[..., 1] - p1[..., 1]) * (p3[..., 0] - p1[..., 0])))
    return -v / 6.0


def compute_geometry(nodes, bc_id):
    """nodes[-2:kmx+3, -2:jmx+3, -2:imx+3, 3] -> (cells, Ifaces, Jfaces, Kfaces) in reference layout."""
    nk, nj, ni, _ = nodes.shape
    imx, jmx, kmx = ni - 6, nj - 6, nk - 6
    N = nodes
    # I faces (-2:imx+3, -2:jmx+2, -2:kmx+2): d1 = n(i,j+1,k+1)-n(i,j,k); d2 = n(i,j,k+1)-n(i,j+1,k)
    d1 = N[1:, 1:, :] - N[:-1, :-1, :]
    d2 = N[1:, :-1, :] - N[:-1, 1:, :]
    If = _finish_faces(_cross_half(d1, d2), (bc_id[0], bc_id[1]), 2, imx)
    # J faces (-2:imx+2, -2:jmx+3, -2:kmx+2): d1 = n(i+1,j,k+1)-n(i,j,k); d2 = n(i+1,j,k)-n(i,j,k+1)
    d1 = N[1:, :, 1:] - N[:-1, :, :-1]
    d2 = N[:-1, :, 1:] - N[1:, :, :-1]
    Jf = _finish_faces(_cross_half(d1, d2), (bc_id[2], bc_id[3]), 1, jmx)
    # K faces (-2:imx+2, -2:jmx+2, -2:kmx+3): d1 = n(i+1,j+1,k)-n(i,j,k); d2 = n(i,j+1,k)-n(i+1,j,k)
    d1 = N[:, 1:, 1:] - N[:, :-1, :-1]
    d2 = N[:, 1:, :-1] - N[:, :-1, 1:]
    Kf = _finish_faces(_cross_half(d1, d2), (bc_id[4], bc_id[5]), 0, kmx)

    cells = np.empty((kmx + 5, jmx + 5, imx + 5, 4))
    cells[..., 0] = 1.0
    # centroid: mean of the 8 nodes, summed in the reference's order (geometry.f90:509-517)
    p = {
        (0, 0, 0): N[:-1, :-1, :-1], (1, 0, 0): N[:-1, :-1, 1:], (1, 1, 0): N[:-1, 1:, 1:], (1, 1, 1): N[1:, 1:, 1:],
        (1, 0, 1): N[1:, :-1, 1:], (0, 1, 0): N[:-1, 1:, :-1], (0, 1, 1): N[1:, 1:, :-1], (0, 0, 1): N[1:, :-1, :-1],
    }
    order = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1), (0, 1, 0), (0, 1, 1), (0, 0, 1)]
    acc = p[order[0]].copy()
    for key in order[1:]:
        acc = acc + p[key]
    cells[..., 1:] = 0.125 * acc
    # volumes on cells 0..imx (python 2..imx+2) (geometry.f90:448-496)
    o = 2
    sk, sj, si = slice(o, kmx + o + 1), slice(o, jmx + o + 1), slice(o, imx + o + 1)
    P = [None,
         p[(0, 0, 0)][sk, sj, si], p[(1, 0, 0)][sk, sj, si], p[(1, 1, 0)][sk, sj, si], p[(0, 1, 0)][sk, sj, si],
         p[(0, 0, 1)][sk, sj, si], p[(1, 0, 1)][sk, sj, si], p[(1, 1, 1)][sk, sj, si], p[(0, 1, 1)][sk, sj, si]]
    v1 = _tet(P[1], P[5], P[8], P[6])
    v1 = v1 + _tet(P[7], P[8], P[6], P[3])
    v1 = v1 + _tet(P[8], P[4], P[1], P[3])
    v1 = v1 + _tet(P[6], P[1], P[3], P[8])
    v1 = v1 + _tet(P[1], P[2], P[6], P[3])
    v2 = _tet(P[2], P[6], P[5], P[7])
    v2 = v2 + _tet(P[8], P[5], P[7], P[4])
    v2 = v2 + _tet(P[5], P[1], P[2], P[4])
    v2 = v2 + _tet(P[7], P[2], P[4], P[5])
    v2 = v2 + _tet(P[2], P[3], P[7], P[4])
    cells[sk, sj, si, 0] = np.maximum(v2, v1)
    if np.any(cells[..., 0] <= 0.0):
        raise ValueError("non-positive cell volume (reference: Fatal_error, geometry.f90:476-494)")
    return (np.ascontiguousarray(cells), np.ascontiguousarray(If), np.ascontiguousarray(Jf), np.ascontiguousarray(Kf))


def wall_distance(nodes, wall_nodes):
    """Brute-force nearest wall node at every node, averaged to cells (wall_dist.f90:84-131).

    ``wall_nodes`` is the global (all blocks) list of no-slip surface nodes, already rounded through the
    ``ES18.10E3`` text file the reference writes and re-reads (wall.f90:93-96)."""
    nk, nj, ni, _ = nodes.shape
    if len(wall_nodes) == 0:
        nd = np.full((nk, nj, ni), 1.0e20)
    else:
        from scipy.spatial import cKDTree
        tree = cKDTree(wall_nodes)
        nd, _ = tree.query(nodes.reshape(-1, 3), k=1)
        nd = nd.reshape(nk, nj, ni)
    d = 0.125 * (nd[:-1, :-1, :-1] + nd[:-1, 1:, :-1] + nd[1:, 1:, :-1] + nd[1:, :-1, :-1]
                 + nd[1:, :-1, 1:] + nd[:-1, :-1, 1:] + nd[:-1, 1:, 1:] + nd[1:, 1:, 1:])
    return np.ascontiguousarray(d)


def surface_nodes(nodes, bc_id):
    """No-slip (-5) surface nodes of one block in the reference's face order (wall.f90:186-260), rounded the way
    the shared text file rounds them ('(3(ES18.10E3,4x))')."""
    nk, nj, ni, _ = nodes.shape
    imx, jmx, kmx = ni - 6, nj - 6, nk - 6
    o = 2
    I = slice(1 + o, imx + o + 1)
    J = slice(1 + o, jmx + o + 1)
    K = slice(1 + o, kmx + o + 1)
    pick = [nodes[K, J, 1 + o], nodes[K, J, imx + o], nodes[K, 1 + o, I], nodes[K, jmx + o, I], nodes[1 + o, J, I], nodes[kmx + o, J, I]]
    out = [pick[f].reshape(-1, 3) for f in range(6) if bc_id[f] == -5]
    if not out:
        return np.zeros((0, 3))
    pts = np.concatenate(out, axis=0)
    return np.array([[float("%.10E" % v) for v in row] for row in pts])
