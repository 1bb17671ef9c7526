"""Host-side mirror of the reference interface for the accelerated path.

``GpuBlock`` is one block (= one MPI rank of the reference) living on one GPU.  ``Solver`` owns the blocks of this
process and exposes the two calls the reference's driver makes every iteration (src/solver.f90:184-185):

    get_next_solution()   <- src/update.f90:129   (all RK variants, halo exchange, BCs, residual, dt, update)
    find_resnorm()        <- src/resnorm.f90:62   (Res_abs(0:n_var), summed over all blocks)

fused into ``iterate(n)`` because the device path reduces the norms inside the last stage's kernel.  Everything
goes through the C ABI (capi.py); nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from . import case as case_mod


class Fest3dError(RuntimeError):
    """The library's replacement of the reference's Fatal_error (message + STOP, src/error.h:1)."""

    def __init__(self, rc, info=None):
        self.rc = rc
        self.info = info
        names = {1: "NaN in flux", 2: "NaN in gradient", 4: "NaN in viscosity", 8: "negative density/pressure or NaN after update",
                 16: "non-positive cell volume", 32: "checkpoint I/O", 64: "configuration not supported by the device path", 128: "CUDA error",
                 256: "bad argument (or an interface face nobody is attached to)", 512: "another rank reported an error"}
        msg = ", ".join(v for k, v in names.items() if rc & k) or "error %d" % rc
        if info is not None and (info.i or info.j or info.k):
            msg += " at block %d cell (%d,%d,%d)" % (info.block_id, info.i, info.j, info.k)
        super().__init__(msg)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def fill_config(cfg, blk):
    s, f, c = blk.scheme, blk.flow, blk.control
    cfg.imx, cfg.jmx, cfg.kmx, cfg.n_var = blk.imx, blk.jmx, blk.kmx, blk.n_var
    cfg.scheme = case_mod.SCHEMES[s.scheme_name]; cfg.interpolant = case_mod.INTERPOLANTS[s.interpolant]
    cfg.turbulence = case_mod.TURBULENCE[s.turbulence]; cfg.transition = case_mod.TRANSITION[s.transition]
    cfg.time_accuracy = case_mod.TIME_ACCURACY[s.time_step_accuracy]
    cfg.time_stepping = 1 if s.time_stepping_method == "g" else 0
    for d in range(3):
        cfg.limiter[d] = s.limiter[d]; cfg.tlimiter[d] = s.tlimiter[d]; cfg.pb_switch[d] = s.pb_switch[d]
    cfg.accur = s.accur
    cfg.mu_variation = 1 if f.mu_variation == "sutherland_law" else 0
    for i in range(6):
        cfg.bc_id[i] = blk.bc_id[i]; cfg.pbc_id[i] = blk.pbc_id[i]
        cfg.dir_switch[i] = blk.dir_switch[i]; cfg.otherface[i] = blk.otherface[i]
        for t in range(2):
            cfg.plo[i][t] = blk.plo[i][t]; cfg.phi[i][t] = blk.phi[i][t]; cfg.pdir[i][t] = blk.pdir[i][t]
    cfg.block_id, cfg.n_blocks = blk.block_id, blk.n_blocks
    cfg.CFL = c.CFL; cfg.global_time_step = s.global_time_step
    for k in ("gm", "R_gas", "mu_ref", "T_ref", "Sutherland_temp", "Pr", "tPr", "density_inf", "x_speed_inf",
              "y_speed_inf", "z_speed_inf", "pressure_inf", "tk_inf", "tw_inf", "vel_mag", "MInf", "tv_inf", "tu_inf", "tkl_inf", "tgm_inf"):
        setattr(cfg, k, getattr(f, k))
    for sl in range(capi.NFIX):
        for i in range(6):
            cfg.fixed[sl][i] = float(blk.fixed[sl, i])
    return cfg


class GpuBlock:
    def __init__(self, blk, device=0, device_geometry=False):
        """device_geometry: build ghost nodes and metrics on the device from the block's grid nodes (fest3d_gpu_setup_geometry)
        instead of uploading the host's cells / Ifaces / Jfaces / Kfaces arrays."""
        self.L = capi.lib()
        self.blk = blk
        self.device = device
        self.cfg = fill_config(capi.Fest3dGpuConfig(), blk)
        self.h = C.c_void_p()
        rc = self.L.fest3d_gpu_create(C.byref(self.h), C.byref(self.cfg), device)
        if rc:
            raise Fest3dError(rc)
        dist = np.ascontiguousarray(blk.dist) if blk.dist is not None else None
        if device_geometry:
            self.setup_geometry(blk.nodes[3:3 + blk.kmx, 3:3 + blk.jmx, 3:3 + blk.imx], dist)
        else:
            self._check(self.L.fest3d_gpu_set_geometry(self.h, _dp(blk.cells), _dp(blk.Ifaces), _dp(blk.Jfaces), _dp(blk.Kfaces), _dp(dist)))
        self.set_state(blk.qp)

    def _check(self, rc):
        if rc:
            info = capi.Fest3dGpuError()
            self.L.fest3d_gpu_error(self.h, C.byref(info))
            raise Fest3dError(rc, info)

    def close(self):
        if self.h:
            self.L.fest3d_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data in / out (host arrays in the reference layout) --
    def set_state(self, qp):
        q = np.ascontiguousarray(qp, dtype=np.float64)
        self._check(self.L.fest3d_gpu_set_state(self.h, _dp(q)))

    def get_state(self, out=None):
        b = self.blk
        q = out if out is not None else np.empty((b.n_var, b.kmx + 5, b.jmx + 5, b.imx + 5))
        self._check(self.L.fest3d_gpu_get_state(self.h, _dp(q)))
        return q

    def set_state_async(self, qp):
        """Start the upload (qp: C-contiguous float64, ideally pinned) and return; takes effect at the next iterate / residual."""
        assert qp.flags["C_CONTIGUOUS"] and qp.dtype == np.float64
        self._check(self.L.fest3d_gpu_set_state_async(self.h, _dp(qp)))

    def get_state_async(self, out):
        """Snapshot qp in stream order and start its download into `out`; valid after state_wait()."""
        assert out.flags["C_CONTIGUOUS"] and out.dtype == np.float64
        self._check(self.L.fest3d_gpu_get_state_async(self.h, _dp(out)))

    def state_wait(self):
        self._check(self.L.fest3d_gpu_state_wait(self.h))

    def setup_geometry(self, grid_nodes, dist=None, want_nodes=False):
        """ghost_grid (grid.f90:137-236) + the metric set-up of geometry.f90:43-545 on the device from the interior nodes
        grid_nodes[kmx, jmx, imx, 3] (the body of the grid file); returns the ghosted node array if asked for."""
        b = self.blk
        g = np.ascontiguousarray(grid_nodes, dtype=np.float64)
        assert g.shape == (b.kmx, b.jmx, b.imx, 3), g.shape
        nodes = np.empty((b.kmx + 6, b.jmx + 6, b.imx + 6, 3)) if want_nodes else None
        d = np.ascontiguousarray(dist, dtype=np.float64) if dist is not None else None
        self._check(self.L.fest3d_gpu_setup_geometry(self.h, _dp(g), _dp(d), _dp(nodes)))
        return nodes

    def get_geometry(self):
        """(cells, Ifaces, Jfaces, Kfaces) as the device holds them, in the reference's layouts."""
        b = self.blk
        cells = np.empty((b.kmx + 5, b.jmx + 5, b.imx + 5, 4))
        If = np.empty((b.kmx + 5, b.jmx + 5, b.imx + 6, 4))
        Jf = np.empty((b.kmx + 5, b.jmx + 6, b.imx + 5, 4))
        Kf = np.empty((b.kmx + 6, b.jmx + 5, b.imx + 5, 4))
        self._check(self.L.fest3d_gpu_get_geometry(self.h, _dp(cells), _dp(If), _dp(Jf), _dp(Kf)))
        return cells, If, Jf, Kf

    def checkpoint_begin(self, path, it):
        """Snapshot the state now (stream order) and write it to `path` in the background while the solver keeps stepping."""
        self._check(self.L.fest3d_gpu_checkpoint_begin(self.h, os.fsencode(path), int(it)))

    def checkpoint_wait(self):
        self._check(self.L.fest3d_gpu_checkpoint_wait(self.h))

    def restart(self, path):
        """Upload the state of a checkpoint file; returns the iteration number stored in it."""
        it = C.c_int(0)
        self._check(self.L.fest3d_gpu_restart(self.h, os.fsencode(path), C.byref(it)))
        return it.value

    def find_wall_dist(self, wall_nodes, want_time=False):
        """find_wall_dist (wall_dist.f90:84-131) on the device from the block's node array and the global list of wall surface
        nodes; fills the context's wall-distance field and returns dist(-2:kmx+2, -2:jmx+2, -2:imx+2)."""
        b = self.blk
        nodes = np.ascontiguousarray(b.nodes, dtype=np.float64)
        wall = np.ascontiguousarray(wall_nodes, dtype=np.float64).reshape(-1, 3)
        out = np.empty((b.kmx + 5, b.jmx + 5, b.imx + 5))
        ms = C.c_double(0.0)
        self._check(self.L.fest3d_gpu_find_wall_dist(self.h, _dp(nodes), _dp(wall) if len(wall) else None, len(wall), _dp(out), C.byref(ms)))
        return (out, ms.value) if want_time else out

    def get_residue(self):
        b = self.blk
        r = np.empty((b.n_var, b.kmx - 1, b.jmx - 1, b.imx - 1))
        self._check(self.L.fest3d_gpu_get_residue(self.h, _dp(r)))
        return r

    def aux(self, which, shape):
        a = np.empty(shape)
        self._check(self.L.fest3d_gpu_get_aux(self.h, which, _dp(a)))
        return a

    def set_stream(self, cuda_stream_handle):
        self._check(self.L.fest3d_gpu_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._check(self.L.fest3d_gpu_sync(self.h))

    def launch_count(self):
        return int(self.L.fest3d_gpu_launch_count(self.h))

    def kernel_timing(self, on=True):
        self.L.fest3d_gpu_kernel_timing(self.h, 1 if on else 0)

    def kernel_time_ms(self, reset=True):
        n = C.c_longlong(0)
        t = self.L.fest3d_gpu_kernel_time_ms(self.h, C.byref(n), 1 if reset else 0)
        return float(t), int(n.value)

    def gradient_time_ms(self, reset=True):
        n = C.c_longlong(0)
        t = self.L.fest3d_gpu_gradient_time_ms(self.h, C.byref(n), 1 if reset else 0)
        return float(t), int(n.value)

    def gradient_path(self):
        return "fused" if self.L.fest3d_gpu_gradient_path(self.h) == 1 else "staged"


class Solver:
    """The blocks of this process, stepped in lock step (drop-in for the reference's per-iteration calls)."""

    def __init__(self, blocks, devices=None, device_geometry=False, gpu_blocks=None):
        self.L = capi.lib()
        if gpu_blocks is not None:
            self.blocks = list(gpu_blocks)
            blocks = [g.blk for g in self.blocks]
        else:
            devices = devices or [0] * len(blocks)
            self.blocks = [GpuBlock(b, d, device_geometry) for b, d in zip(blocks, devices)]
        for i, a in enumerate(self.blocks):
            for b in self.blocks[i:]:      # b is a: a block that is its own (periodic) neighbour
                ids_a = set(a.blk.bc_id) | set(a.blk.pbc_id)
                if b.blk.block_id in ids_a:
                    self.L.fest3d_gpu_link_local(a.h, b.h)
        self.n_var = blocks[0].n_var
        self.current_iter = 1   # control%current_iter after setup (solver.f90:140)
        self._handles = (C.c_void_p * len(self.blocks))(*[b.h for b in self.blocks])

    @classmethod
    def from_gpu_blocks(cls, gpu_blocks):
        """Blocks that were created (and uploaded) one at a time, e.g. to drop the host copies of their big arrays in between."""
        return cls(None, gpu_blocks=gpu_blocks)

    def close(self):
        for b in self.blocks:
            b.close()

    def residual_only(self):
        """get_total_conservative_Residue on every block, nothing copied back (throughput runs of the residual path)."""
        rc = self.L.fest3d_gpu_residual_group(self._handles, len(self.blocks), self.current_iter)
        if rc:
            self.blocks[0]._check(rc)

    def set_states_async(self, qps):
        for b, q in zip(self.blocks, qps):
            b.set_state_async(q)

    def get_states_async(self, outs):
        for b, q in zip(self.blocks, outs):
            b.get_state_async(q)

    def state_wait(self):
        for b in self.blocks:
            b.state_wait()

    def init_comm(self, n_ranks, rank, unique_id, block_to_rank):
        arr = (C.c_int * len(block_to_rank))(*block_to_rank)
        for b in self.blocks:
            b._check(self.L.fest3d_gpu_comm_init(b.h, n_ranks, rank, unique_id, arr))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = capi.lib().fest3d_gpu_comm_unique_id(buf)
        if rc:
            raise Fest3dError(rc)
        return buf.raw

    def residual(self):
        """get_total_conservative_Residue on every block (update.f90:495); returns the per-block residue arrays."""
        rc = self.L.fest3d_gpu_residual_group(self._handles, len(self.blocks), self.current_iter)
        if rc:
            self.blocks[0]._check(rc)
        return [b.get_residue() for b in self.blocks]

    def iterate(self, n_iters=1, want_norms=True):
        """n iterations of get_next_solution + find_resnorm.  Returns Res_abs[n_iters, n_var+1] (or None)."""
        res = np.zeros((n_iters, self.n_var + 1)) if want_norms else None
        rc = self.L.fest3d_gpu_step_group(self._handles, len(self.blocks), self.current_iter, n_iters, _dp(res))
        if rc:
            for b in self.blocks:
                b._check(rc)
        self.current_iter += n_iters
        return res

    def iterate_begin(self, n_iters=1):
        """Queue n iterations and return; iterate_end() waits for them and returns Res_abs[n_iters, n_var+1]."""
        rc = self.L.fest3d_gpu_step_group_begin(self._handles, len(self.blocks), self.current_iter, n_iters)
        if rc:
            for b in self.blocks:
                b._check(rc)
        self._in_flight = n_iters

    def iterate_end(self, want_norms=True):
        n_iters = self._in_flight
        res = np.zeros((n_iters, self.n_var + 1)) if want_norms else None
        rc = self.L.fest3d_gpu_step_group_end(self._handles, len(self.blocks), _dp(res))
        if rc:
            for b in self.blocks:
                b._check(rc)
        self.current_iter += n_iters
        return res

    def checkpoint_begin(self, prefix):
        """Asynchronous checkpoint of every block (file prefix + '_<block id>.f3dckpt'); stepping may continue at once."""
        for b in self.blocks:
            b.checkpoint_begin("%s_%02d.f3dckpt" % (prefix, b.blk.block_id), self.current_iter)

    def checkpoint_wait(self):
        for b in self.blocks:
            b.checkpoint_wait()

    def restart(self, prefix):
        its = {b.restart("%s_%02d.f3dckpt" % (prefix, b.blk.block_id)) for b in self.blocks}
        assert len(its) == 1, its
        self.current_iter = its.pop()
        return self.current_iter

    # names of the reference, for hosts written against them
    def get_next_solution(self):
        self._last = self.iterate(1)
        return self

    def find_resnorm(self):
        return self._last[0]
