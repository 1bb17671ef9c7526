"""Run under torchrun with N ranks (one GPU each): every rank owns F3D_BLOCKS_PER_RANK (default 1) blocks of an SST duct, the
halo swap goes through ncclSend/ncclRecv between ranks and device-to-device between the blocks of one rank, the norms through
ncclAllReduce inside the library; every rank compares the residual-norm history (Res_abs(0) included) and the final state of
its blocks with the CPU oracle stepping all blocks in lock step.
Exit code 0 = parity (history and state within 1e-10)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import helpers
    import oracle_py
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    par = importlib.import_module("fest3d_b200.parallel")
    syn = importlib.import_module("fest3d_b200.synthetic")
    solver = importlib.import_module("fest3d_b200.solver")
    per = int(os.environ.get("F3D_BLOCKS_PER_RANK", "1"))
    n_blocks = world * per
    nb = par.block_grid(n_blocks)
    # F3D_MP_CASE: the model / integrator the ranks run (the halo messages carry n_var = 7, 8, 7, 7 fields per cell)
    case = os.environ.get("F3D_MP_CASE", "sst_rk4")
    extra = {"sst_rk4": dict(turbulence="sst", time_step_accuracy="RK4", CFL=0.5),
             "lctm_rk2": dict(turbulence="sst", transition="lctm2015", time_step_accuracy="RK2", CFL=0.5),
             "kkl_none": dict(turbulence="kkl", time_step_accuracy="none", CFL=0.5),
             "sst_implicit": dict(turbulence="sst", time_step_accuracy="implicit", CFL=20.0)}[case]
    kw = dict(n3=(12, 10, 8), nb=nb, **extra)
    all_blocks = syn.make_duct_blocks(None, **kw)
    owners = par.block_to_rank(n_blocks, world)
    mine = [b for b in all_blocks if owners[b.block_id] == rank]
    s = solver.Solver(mine, devices=[local] * len(mine))
    uid = par.broadcast_unique_id(dist, solver.Solver.unique_id, rank, device="cuda")
    s.init_comm(world, rank, uid, owners)
    n_it = 6
    hist = s.iterate(n_it)
    w = oracle_py.OracleWorld(all_blocks)
    ho, ms = [], []
    for it in range(1, n_it + 1):
        ho.append(w.step(it)[1]); ms.append(helpers.boundary_mass_flux_scale(w, all_blocks))
    ho = np.array(ho)
    floor = np.abs(ho[:, 1:]).max(axis=0) * 1e-3
    rel = (np.abs(hist[:, 1:] - ho[:, 1:]) / np.maximum(np.abs(ho[:, 1:]), floor)).max()
    rel = max(rel, (np.abs(hist[:, 0] - ho[:, 0]) / np.array(ms)).max())   # Res_abs(0) on the scale of sum |boundary mass flux|
    d = max(max(helpers.state_rel_diff(gb.get_state(), w.get_state(gb.blk.block_id))) for gb in s.blocks)
    ok = rel < 1e-10 and d < 1e-10
    print("%s rank %d (%d blocks): history %.2e state %.2e %s" % (case, rank, len(mine), rel, d, "ok" if ok else "FAIL"), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    s.close()
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
