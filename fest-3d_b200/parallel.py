"""Host-side logic of the multi-rank path (one rank per GPU, one or more blocks per rank).

The data path itself is in the library (csrc/api.cu:exchange): pack kernels -> ncclSend/ncclRecv of every interface
face in one group -> unpack kernels, and one ncclAllReduce for the residual norms.  This module holds what the host must
agree on across ranks and what the CPU tests can check without a GPU:

  * block -> rank ownership (the reference's rule is rank == block, src/layout.f90:23-25, 101-105),
  * the ordered list of halo messages each rank posts.  NCCL pairs the sends and receives between two ranks in posting
    order, so both sides must enumerate their shared interface faces in the same order: sends are sorted by
    (peer rank, own block, own face), receives by (peer rank, neighbour block, neighbour face) -- the same key seen
    from the other side (replaces the tag-1 blocking MPI_SENDRECVs of src/interface1.f90:140-462),
  * the hand-over of the NCCL unique id and the max-over-ranks reduction of the timed region.
"""
from __future__ import annotations

from dataclasses import dataclass


def block_grid(n_ranks):
    """Block lattice of the weak-scaling duct: 1, 2, 4, 8 ranks -> 1x1x1, 2x1x1, 2x2x1, 2x2x2 blocks."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n_ranks, (n_ranks, 1, 1))


def block_to_rank(n_blocks, n_ranks):
    """Owner rank of every block: contiguous groups of n_blocks / n_ranks (rank == block when they are equal)."""
    if n_blocks % n_ranks:
        raise ValueError("n_blocks (%d) must be a multiple of n_ranks (%d)" % (n_blocks, n_ranks))
    per = n_blocks // n_ranks
    return [b // per for b in range(n_blocks)]


def rank_blocks(n_blocks, n_ranks, rank):
    return [b for b, r in enumerate(block_to_rank(n_blocks, n_ranks)) if r == rank]


@dataclass(frozen=True)
class HaloMsg:
    peer_rank: int
    block: int        # the block whose interior layers travel
    face: int         # its face (1..6)
    my_block: int     # the local block the message belongs to
    my_face: int
    n_doubles: int    # 3 layers x n_var x face cells (interface1.f90:77-79)


def _face_cells(blk, face):
    mx = (blk.imx, blk.jmx, blk.kmx)
    ax = (face - 1) // 2
    a, b = (1 if ax == 0 else 0), (1 if ax == 2 else 2)
    return (mx[a] - 1) * (mx[b] - 1)


def halo_plan(blocks, owners, rank):
    """(sends, recvs) of `rank`, each in posting order.  `blocks`: the BlockSetup objects this rank owns (any order)."""
    sends, recvs = [], []
    for blk in blocks:
        for f in range(6):
            nb = blk.bc_id[f] if blk.bc_id[f] >= 0 else blk.pbc_id[f]
            if nb < 0 or owners[nb] == rank:
                continue    # physical boundary, or a neighbour in this process (device-to-device link)
            n = 3 * blk.n_var * _face_cells(blk, f + 1)
            sends.append(HaloMsg(owners[nb], blk.block_id, f + 1, blk.block_id, f + 1, n))
            recvs.append(HaloMsg(owners[nb], nb, blk.otherface[f], blk.block_id, f + 1, n))
    key = lambda m: (m.peer_rank, m.block, m.face)
    return sorted(sends, key=key), sorted(recvs, key=key)


def broadcast_unique_id(dist, make_id, rank, device="cpu"):
    """Rank 0 creates the 128-byte NCCL unique id, everybody receives it (the one host-side collective of set-up)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def max_over_ranks(dist, value, device="cpu"):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
