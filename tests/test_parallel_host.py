"""World-size-2 (gloo, CPU) tests of the host-side multi-rank logic: block ownership, the halo message plan whose posting
order pairs sends with receives (NCCL semantics), the unique-id hand-over and the max-over-ranks reduction.  The data path
itself (pack / NCCL / unpack kernels) is covered on the GPU box by tests/test_gpu_parity.py::test_duct_multiblock_*."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, nb, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    par = importlib.import_module("fest3d_b200.parallel")
    syn = importlib.import_module("fest3d_b200.synthetic")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_blocks = nb[0] * nb[1] * nb[2]
        owners = par.block_to_rank(n_blocks, world)
        mine = par.rank_blocks(n_blocks, world, rank)
        blocks = syn.make_duct_blocks(None, n3=(6, 5, 4), nb=nb, only_blocks=mine)
        assert [b.block_id for b in blocks] == mine
        uid = par.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
        assert uid == bytes(range(128))
        sends, recvs = par.halo_plan(blocks, owners, rank)
        assert len(sends) == len(recvs)
        # post everything the way the library does (all sends, all receives, one group) and check that message m of the
        # receive list is the one its key says: payload = [block, face] of the sender
        reqs, bufs = [], []
        for m in sends:
            t = torch.full((m.n_doubles,), float(100 * m.block + m.face), dtype=torch.float64)
            reqs.append(dist.isend(t, m.peer_rank)); bufs.append(t)
        got = []
        for m in recvs:
            t = torch.empty(m.n_doubles, dtype=torch.float64)
            reqs.append(dist.irecv(t, m.peer_rank)); got.append((m, t))
        for r in reqs:
            r.wait()
        for m, t in got:
            assert float(t[0]) == 100 * m.block + m.face and float(t[-1]) == 100 * m.block + m.face, (rank, m)
        tmax = par.max_over_ranks(dist, 1.0 + rank)
        assert tmax == float(world)
        q.put((rank, len(sends), sum(m.n_doubles for m in sends)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nb", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_halo_plan_pairs_up_world2(nb):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nb, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = sorted(q.get() for _ in range(2))
    # the two ranks exchange the same number of faces and bytes
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2] and out[0][1] > 0


def test_block_ownership():
    par = importlib.import_module("fest3d_b200.parallel")
    assert par.block_to_rank(8, 8) == list(range(8))          # the reference's rank == block
    assert par.block_to_rank(8, 2) == [0, 0, 0, 0, 1, 1, 1, 1]
    assert par.rank_blocks(8, 4, 3) == [6, 7]
    assert par.block_grid(4) == (2, 2, 1)
    with pytest.raises(ValueError):
        par.block_to_rank(8, 3)


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under torchrun: ranks != 0 exit without work (contract, tier 4)."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
