#!/usr/bin/env python
"""bench.py -- FP64 cell-updates/s of the explicit residual + update hot path on synthetic structured grids.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          one rank per GPU (torchrun for N>1), weak scaling
  python bench.py --impl reference ...                   the CPU arm: the oracle port on the host cores
  python bench.py --interpolant weno --scheme ausmP      BASELINE.json's second synthetic configuration (not the headline line)

A *step* is one iteration of get_next_solution + find_resnorm (src/solver.f90:184-185) on a 256^3-cell block per GPU,
MUSCL + AUSM + SST, single-stage explicit update (time_step_accuracy 'none'): one residual evaluation + one update per
cell = one cell-update.  `value` is timed with the state resident in HBM; `e2e` goes through the C ABI with HOST
buffers (set_state H2D + step + get_state D2H + norms) inside the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "FP64 cell-updates/s (explicit residual evaluation + time update)"
UNIT = "cell-updates/s"


def algorithmic_bytes_per_cell_update(nv, viscous, sst, stages_rk):
    """SURVEY 8(d) / BASELINE.md 3, contract geometry representation (the kernel reads A,nx,ny,nz per face):
    nv read + nv write + volume + 3 faces x 4 [+ 3 centre] [+ dist] ; multi-stage adds dt + U_store + 2 R_store."""
    n_geom = 13 + (3 if viscous else 0)
    n = nv + nv + n_geom + (1 if sst else 0)
    if stages_rk:
        n += 1 + nv + 2 * nv
    return 8 * n


def block_grid(n_ranks):
    return importlib.import_module("fest3d_b200.parallel").block_grid(n_ranks)


def cpu_block_lattice(cores):
    """One oracle thread per block (the reference runs one MPI rank per block): the largest lattice that fits the host cores."""
    best = (1, 1, 1)
    for nb in [(2, 1, 1), (2, 2, 1), (2, 2, 2), (4, 2, 2), (4, 4, 2), (4, 4, 4), (8, 4, 4)]:
        if nb[0] * nb[1] * nb[2] <= cores:
            best = nb
    return best


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nme, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(args, n, rank, world):
    """CPU arm: the reference's own implementation cannot be built here (Fortran + MPI, no compiler in the image), so this
    times the oracle port (kind 'port', timing build -O3) on the host cores: one thread per block, 8 blocks of a bounded
    size, same scheme / flow / BCs as the GPU arm."""
    if rank != 0:
        return
    import oracle_py
    syn = importlib.import_module("fest3d_b200.synthetic")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nb = cpu_block_lattice(cores)
    nblk = nb[0] * nb[1] * nb[2]
    m = args.cpu_cells
    blocks = syn.make_duct_blocks(m, nb=nb, scheme_name=args.scheme, interpolant=args.interpolant, turbulence="sst", time_step_accuracy="none", CFL=0.5)
    w = oracle_py.OracleWorld(blocks, fast=True)
    it = 1
    for _ in range(args.warmup):
        w.step(it); it += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        err, _r = w.step(it); it += 1
    dt = time.perf_counter() - t0
    cells = nblk * m ** 3
    val = cells * args.steps / dt
    sample = "%d blocks of %d^3 cells (one thread each), %d steps; same %s duct as the GPU arm at reduced size" % (nblk, m, args.steps, scheme_label(args))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(n, args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(cores, nblk), "host_cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


SCHEME_LABEL = {"van_leer": "van Leer", "ldfss0": "LDFSS(0)", "ausm": "AUSM", "ausmP": "AUSM+", "ausmUP": "AUSM+-up", "slau": "SLAU"}
INTERP_LABEL = {"none": "first order", "muscl": "MUSCL", "ppm": "PPM", "weno": "WENO", "weno_NM": "WENO-NM"}


def scheme_label(args):
    return "%s + %s + SST" % (INTERP_LABEL[args.interpolant], SCHEME_LABEL[args.scheme])


def workload_config(n_gpus, args):
    return {"workload": "synthetic duct, %d^3 cells per GPU (%s blocks), %s k-omega, explicit single-stage update, local time step" % (args.cells, "x".join(map(str, block_grid(n_gpus))), scheme_label(args)),
            "cells_per_gpu": args.cells ** 3, "n_var": 7, "time_integration": "none (1 stage)",
            "l2_policy": "inputs larger than L2 (state+geometry %.1f GB per GPU vs 126 MB L2)" % (args.cells ** 3 * 8 * 31 / 1e9)}


def cpu_baseline_sample(args):
    """Oracle port timed on a bounded sample on this box's host cores (reported baseline, not the target)."""
    import oracle_py
    syn = importlib.import_module("fest3d_b200.synthetic")
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nb = cpu_block_lattice(cores)
    nblk = nb[0] * nb[1] * nb[2]
    m = args.cpu_cells
    blocks = syn.make_duct_blocks(m, nb=nb, scheme_name=args.scheme, interpolant=args.interpolant, turbulence="sst", time_step_accuracy="none", CFL=0.5)
    w = oracle_py.OracleWorld(blocks, fast=True)
    w.step(1)
    steps = 3
    t0 = time.perf_counter()
    for it in range(2, 2 + steps):
        w.step(it)
    dt = time.perf_counter() - t0
    return {"value": nblk * m ** 3 * steps / dt, "unit": UNIT, "cores": min(cores, nblk), "host_cores": cores, "kind": "port",
            "sample": "%d blocks of %d^3 cells, one thread per block, %d steps of the same %s duct" % (nblk, m, steps, scheme_label(args))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=256, help="cells per block edge per GPU")
    ap.add_argument("--cpu-cells", type=int, default=48, help="cells per block edge of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    # the headline configuration is the default; BASELINE.json's other synthetic config (WENO + AUSM+ + SST) is --interpolant weno --scheme ausmP
    ap.add_argument("--scheme", default="ausm", choices=sorted(SCHEME_LABEL))
    ap.add_argument("--interpolant", default="muscl", choices=sorted(INTERP_LABEL))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus

    if args.impl == "reference":
        run_reference(args, n, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    syn = importlib.import_module("fest3d_b200.synthetic")
    solver_mod = importlib.import_module("fest3d_b200.solver")

    nb = block_grid(world)
    blocks = syn.make_duct_blocks(args.cells, nb=nb, scheme_name=args.scheme, interpolant=args.interpolant, turbulence="sst",
                                  time_step_accuracy="none", CFL=0.5, only_blocks=[rank])
    blk = blocks[0]
    s = solver_mod.Solver(blocks, devices=[local_rank])
    gb = s.blocks[0]
    if world > 1:
        par = importlib.import_module("fest3d_b200.parallel")
        uid = par.broadcast_unique_id(dist, solver_mod.Solver.unique_id, rank, device="cuda")
        s.init_comm(world, rank, uid, par.block_to_rank(world, world))
    stream = torch.cuda.Stream()          # a real (non-default) stream, so CUDA events bracket exactly our launches
    torch.cuda.set_stream(stream)
    gb.set_stream(stream.cuda_stream)
    nvp1 = blk.n_var + 1
    cells = (blk.imx - 1) * (blk.jmx - 1) * (blk.kmx - 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    s.iterate(args.warmup, want_norms=False)
    launches0 = gb.launch_count()
    gb.kernel_timing(True)
    gb.kernel_time_ms(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    norms = s.iterate(args.steps, want_norms=True)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = gb.launch_count() - launches0
    k_ms, k_n = gb.kernel_time_ms(reset=True)
    gb.kernel_timing(False)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = cells * world * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ----
    q_host = torch.from_numpy(np.ascontiguousarray(blk.qp)).pin_memory()
    q_back = torch.empty_like(q_host).pin_memory()
    q_np, qb_np = q_host.numpy(), q_back.numpy()
    e2e_steps = max(3, min(args.steps, 5))
    gb.set_state(q_np); s.iterate(1); gb.get_state(qb_np)   # warm-up of the path
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        gb.set_state(q_np)            # H2D of the step's input state
        r = s.iterate(1)              # one cell-update everywhere + norms (D2H of n_var+1 doubles)
        gb.get_state(qb_np)           # D2H of the step's result
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.stop_flag = True
    tm2 = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm2, op=dist.ReduceOp.MAX)
    e2e_value = cells * world * e2e_steps / (float(tm2.item()) * 1e-3)
    state_bytes = int(q_host.numel() * 8)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        bpc = algorithmic_bytes_per_cell_update(blk.n_var, True, True, False)
        k_avg_ms = k_ms / max(k_n, 1)
        achieved = bpc * cells / (k_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_residual_dram_bytes_per_launch")
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "g3::k_sweep3<7,%s,%s,viscous> (fused reconstruction + flux + source + dt + update, generation %s)" % (INTERP_LABEL[args.interpolant], SCHEME_LABEL[args.scheme], os.environ.get("F3D_SWEEP_GEN", "3")), "kernel_ms": k_avg_ms,
                         "kernel_share_of_step": k_ms / ms, "algorithmic_bytes_per_cell_update": bpc, "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes + nvp1 * 8,
                    "steps": e2e_steps},
            "gpu_launches": launches, "clocks": sampler.summary(),
            "res_abs_last": [float(x) for x in norms[-1]],
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample(args)
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
