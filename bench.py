#!/usr/bin/env python
"""bench.py -- FP64 cell-updates/s of the explicit residual + update hot path on synthetic structured grids.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          one rank per GPU (torchrun for N>1), weak scaling
  python bench.py --impl reference ...                   the CPU arm: the oracle port on the host cores
  python bench.py --interpolant weno --scheme ausmP --mode residual
                                                         BASELINE.json's second synthetic configuration: residual evaluations/s
  python bench.py --scaling strong --gpus N              BASELINE.json's multi-block configuration: 512^3 as 8 blocks on N GPUs
  python bench.py --gradients fused                      the fused form of the viscous path (A/B against the default staged form)

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
    steps = 8
    t0 = time.perf_counter()
    for it in range(2, 2 + steps):
        w.step(it)
    dt = time.perf_counter() - t0
    return {"value": nblk * m ** 3 * steps / dt, "unit": UNIT, "cores": min(cores, nblk), "host_cores": cores, "kind": "port",
            "sample": "%d blocks of %d^3 cells, one thread per block, %d steps of the same %s duct" % (nblk, m, steps, scheme_label(args))}


def bind_to_gpu_numa_node(torch, local_rank):
    """Host side of the e2e leg: run (and first-touch the pinned buffers) on the CPUs of the NUMA node the GPU hangs off, as a deployment
    with numactl would.  Returns (description, previous affinity) -- the CPU baseline leg gets its full affinity back afterwards."""
    prev = os.sched_getaffinity(0)
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return {"gpu_numa_node": None}, prev
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        use = cpus & prev
        if use:
            os.sched_setaffinity(0, use)
        return {"gpu_numa_node": node, "cpus_allowed": len(prev), "cpus_on_node": len(use), "bound": bool(use)}, prev
    except Exception as e:          # no sysfs / no permission: run where we are
        return {"gpu_numa_node": "unknown (%s)" % type(e).__name__}, prev


def kernel_profile(kernel_key):
    """Per-kernel constants taken from committed ncu captures (profiles/kernels.json): DRAM bytes per launch at the bench workload
    and FP64-pipe warp instructions per cell-warp.  Keyed by the kernel that actually ran; a kernel without a capture gets null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernels.json"))).get(kernel_key, {})
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=256, help="cells per block edge")
    ap.add_argument("--cpu-cells", type=int, default=96, help="cells per block edge of the CPU sample (96^3 x ~60 arrays per block: far beyond the host caches)")
    ap.add_argument("--no-cpu", action="store_true")
    # the headline configuration is the default; BASELINE.json's other synthetic config (WENO + AUSM+ + SST, residual evaluations) is
    # --interpolant weno --scheme ausmP --mode residual; its multi-block config is --scaling strong (512^3 as 8 blocks on 1/2/4/8 GPUs)
    ap.add_argument("--scheme", default="ausm", choices=sorted(SCHEME_LABEL))
    ap.add_argument("--interpolant", default="muscl", choices=sorted(INTERP_LABEL))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--mode", default="step", choices=["step", "residual"], help="residual: one get_total_conservative_Residue per step (no update)")
    ap.add_argument("--gradients", default=None, choices=["staged", "fused"], help="form of the viscous path (default: the library's)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    if args.gradients:
        os.environ["F3D_GRADIENTS"] = args.gradients

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
    par = importlib.import_module("fest3d_b200.parallel")

    # weak: one block per GPU (1, 2, 4, 8 blocks); strong: always the 2 x 2 x 2 blocks of the 512^3 case, 8 / world of them per GPU
    if args.scaling == "strong":
        nb, n_blocks = (2, 2, 2), 8
        if 8 % world:
            raise SystemExit("--scaling strong needs 1, 2, 4 or 8 GPUs")
    else:
        nb, n_blocks = block_grid(world), world
    owners = par.block_to_rank(n_blocks, world)
    mine = [b for b in range(n_blocks) if owners[b] == rank]
    # blocks are built and uploaded one at a time; the host copies of the big arrays are dropped afterwards (8 blocks of 256^3 are
    # ~30 GB of numpy arrays otherwise)
    gblocks, blocks = [], []
    q_keep = None
    for b in mine:
        blk = syn.make_duct_blocks(args.cells, nb=nb, scheme_name=args.scheme, interpolant=args.interpolant, turbulence="sst",
                                   time_step_accuracy="none", CFL=0.5, only_blocks=[b])[0]
        gblocks.append(solver_mod.GpuBlock(blk, local_rank))
        if q_keep is None:
            q_keep = blk.qp
        blk.cells = blk.Ifaces = blk.Jfaces = blk.Kfaces = blk.nodes = blk.dist = None
        if len(blocks):
            blk.qp = None
        blocks.append(blk)
    s = solver_mod.Solver.from_gpu_blocks(gblocks)
    blk = blocks[0]
    gb = s.blocks[0]
    if world > 1:
        uid = par.broadcast_unique_id(dist, solver_mod.Solver.unique_id, rank, device="cuda")
        s.init_comm(world, rank, uid, owners)
    stream = torch.cuda.Stream()          # a real (non-default) stream, so CUDA events bracket exactly our launches
    torch.cuda.set_stream(stream)
    gb.set_stream(stream.cuda_stream)     # block 0 of the rank runs on the timed stream; the others on their own, joined by events
    nvp1 = blk.n_var + 1
    cells_blk = (blk.imx - 1) * (blk.jmx - 1) * (blk.kmx - 1)
    cells_rank = cells_blk * len(mine)
    cells_all = cells_blk * n_blocks
    path = gb.gradient_path()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_call(k, want):
        if args.mode == "residual":
            for _ in range(k):
                s.residual_only()
            return None
        return s.iterate(k, want_norms=want)

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    one_call(args.warmup, False)
    launches0 = sum(g.launch_count() for g in s.blocks)
    for g in s.blocks:
        g.kernel_timing(True); g.kernel_time_ms(reset=True); g.gradient_time_ms(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    norms = one_call(args.steps, True)
    e1.record(stream)
    barrier()                              # (every block's stream is synchronised here: the elapsed time below is taken up to e1 on
    ms = e0.elapsed_time(e1)               # block 0's stream, which the norm assembly of the call makes wait for all blocks)
    launches = sum(g.launch_count() for g in s.blocks) - launches0
    k_ms = k_n = g_ms = g_n = 0
    for g in s.blocks:
        t, c = g.kernel_time_ms(reset=True); k_ms += t; k_n += c
        t, c = g.gradient_time_ms(reset=True); g_ms += t; g_n += c
        g.kernel_timing(False)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = cells_all * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ("e2e") ----
    e2e_value = None
    state_bytes = 0
    e2e_steps = max(3, min(args.steps, 12))   # the pipeline needs a few periods to reach its steady rate (one upload / download of 1 GB each per step)
    numa_info, prev_affinity = {}, None
    if args.mode == "step":
        numa_info, prev_affinity = bind_to_gpu_numa_node(torch, local_rank)
        q_host = torch.from_numpy(np.ascontiguousarray(q_keep)).pin_memory()
        q_back = torch.empty_like(q_host).pin_memory()
        q_np, qb_np = q_host.numpy(), q_back.numpy()
        for g in s.blocks:
            g.set_state(q_np)
        s.iterate(1)
        for g in s.blocks:
            g.get_state(qb_np)   # warm-up of the path
        # ... and of the asynchronous one: its staging buffers (3 GB of cudaMalloc per block), copy streams and events are created by the
        # first asynchronous call -- tens to hundreds of milliseconds on a fresh box, which read as 0.44 - 0.75 G cell-updates/s from run to
        # run while they sat inside the timed region
        for _ in range(max(2, min(args.warmup, 3))):
            s.set_states_async([q_np] * len(s.blocks))
            s.iterate_begin(1)
            s.iterate_end()
            s.get_states_async([qb_np] * len(s.blocks))
        s.state_wait()
        barrier()
        e0.record(stream)
        s.set_states_async([q_np] * len(s.blocks))             # H2D of step 0's input state (every block of the rank)
        for step in range(e2e_steps):
            s.iterate_begin(1)                                 # takes the uploaded state, one cell-update everywhere (queued, returns)
            if step + 1 < e2e_steps:
                s.set_states_async([q_np] * len(s.blocks))     # H2D of the NEXT step's input: overlaps this step and the previous D2H
            r = s.iterate_end()                                # this step's norms (D2H of n_var+1 doubles)
            s.get_states_async([qb_np] * len(s.blocks))        # D2H of this step's result (PCIe is full duplex)
        s.state_wait()                                         # the last result has arrived on the host
        e1.record(stream)
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        tm2 = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm2, op=dist.ReduceOp.MAX)
        e2e_value = cells_all * e2e_steps / (float(tm2.item()) * 1e-3)
        state_bytes = int(q_host.numel() * 8) * len(s.blocks)
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)
    sampler.stop_flag = True

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
        g_avg_ms = g_ms / max(g_n, 1) if g_n else 0.0
        achieved = bpc * cells_blk / (k_avg_ms * 1e-3) / 1e9
        step_gbs = bpc * cells_rank / (ms / args.steps * 1e-3) / 1e9
        kname = ("g4::k_fused" if path == "fused" else "g3::k_sweep3") + "<7,%s,%s,viscous>" % (INTERP_LABEL[args.interpolant], SCHEME_LABEL[args.scheme])
        kp = kernel_profile(kname)
        fp64_frac = None
        if kp.get("fp64_warp_instr_per_cell_warp"):   # FP64 pipe: one warp instruction per 2 cycles per SM sub-partition (4 per SM, 148 SMs)
            sm_mhz = (sampler.summary().get("sm_mhz") or 1965.0)
            fp64_frac = kp["fp64_warp_instr_per_cell_warp"] * (cells_blk / 32.0) * 2.0 / (148 * 4) / (k_avg_ms * 1e-3 * sm_mhz * 1e6)
        halo = 0
        for b in blocks:
            for f in range(6):
                if b.bc_id[f] >= 0:
                    mxs = (b.imx, b.jmx, b.kmx); ax = f // 2
                    a_, b_ = (1 if ax == 0 else 0), (1 if ax == 2 else 2)
                    halo += 3 * b.n_var * (mxs[a_] - 1) * (mxs[b_] - 1) * 8
        cfg = workload_config(world, args)
        cfg.update({"scaling_mode": args.scaling, "blocks_total": n_blocks, "blocks_per_gpu": len(mine), "gradients": path, "mode": args.mode,
                    "halo_bytes_sent_per_stage_per_gpu": halo})
        if args.scaling == "strong":
            cfg["workload"] = "synthetic duct, %d^3 cells as 2x2x2 blocks of %d^3 (%d per GPU), %s k-omega, explicit single-stage update, local time step" % (
                2 * args.cells, args.cells, len(mine), scheme_label(args))
        if args.mode == "residual":
            cfg["workload"] = cfg["workload"].replace("explicit single-stage update", "residual evaluation only (get_total_conservative_Residue)")
        line = {
            "metric": METRIC if args.mode == "step" else "FP64 residual evaluations/s (cells x get_total_conservative_Residue calls per second)",
            "value": value, "unit": UNIT if args.mode == "step" else "cell-residuals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": kp.get("dram_bytes_per_launch"),
                         "kernel": kname + (" (reconstruction + flux + source + dt + update; gradients staged from k_gradients)" if path == "staged" else
                                            " (gradients + reconstruction + flux + source + dt + update in one tile pass)"),
                         "kernel_ms": k_avg_ms, "kernel_share_of_step": k_ms / max(len(s.blocks), 1) / ms * 1.0 if len(s.blocks) == 1 else k_ms / ms,
                         "gradient_kernels_ms": g_avg_ms, "step_achieved": step_gbs, "step_frac": step_gbs / peak,
                         "fp64_pipe_frac": fp64_frac, "fp64_pipe_note": "FP64 warp instructions per cell-warp of this kernel (ncu, profiles/kernels.json) x cells / measured kernel time, against 1 per 2 cycles per SM sub-partition at the sampled SM clock",
                         "algorithmic_bytes_per_cell_update": bpc, "peak_source": peak_src},
            "gpu_launches": launches, "clocks": sampler.summary(),
        }
        if e2e_value is not None:
            line["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes + nvp1 * 8, "steps": e2e_steps, "host_numa": numa_info}
        else:
            line["e2e"] = {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "residual-only mode keeps no host-visible result per call; the headline step mode carries the end-to-end figure"}
        if norms is not None:
            line["res_abs_last"] = [float(x) for x in norms[-1]]
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample(args)
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
