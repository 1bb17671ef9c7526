"""ctypes binding of the C ABI (include/fest3d_gpu.h -> libfest3d_gpu.so).

No CPU fallback: importing works anywhere (so that the symbol table can be checked on a CPU box), but every compute
entry point needs the CUDA library AND a device, and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("F3D_LIB") or os.path.join(HERE, "libfest3d_gpu.so")   # F3D_LIB: an alternative build (A/B measurements)
NFIX = 13

SYMBOLS = [
    "fest3d_gpu_create", "fest3d_gpu_destroy", "fest3d_gpu_set_stream", "fest3d_gpu_sync", "fest3d_gpu_set_geometry",
    "fest3d_gpu_set_state", "fest3d_gpu_get_state", "fest3d_gpu_step", "fest3d_gpu_step_group", "fest3d_gpu_residual",
    "fest3d_gpu_residual_group", "fest3d_gpu_get_residue", "fest3d_gpu_get_aux", "fest3d_gpu_error",
    "fest3d_gpu_comm_unique_id", "fest3d_gpu_comm_init", "fest3d_gpu_link_local", "fest3d_gpu_launch_count",
    "fest3d_gpu_kernel_timing", "fest3d_gpu_kernel_time_ms", "fest3d_gpu_version", "fest3d_gpu_find_wall_dist", "fest3d_gpu_setup_geometry", "fest3d_gpu_get_geometry",
    "fest3d_gpu_checkpoint_begin", "fest3d_gpu_checkpoint_wait", "fest3d_gpu_restart", "fest3d_gpu_gradient_time_ms", "fest3d_gpu_gradient_path",
    "fest3d_gpu_set_state_async", "fest3d_gpu_get_state_async", "fest3d_gpu_state_wait", "fest3d_gpu_step_group_begin", "fest3d_gpu_step_group_end",
]


class Fest3dGpuConfig(C.Structure):
    _fields_ = [
        ("imx", C.c_int), ("jmx", C.c_int), ("kmx", C.c_int), ("n_var", C.c_int),
        ("scheme", C.c_int), ("interpolant", C.c_int), ("turbulence", C.c_int), ("transition", C.c_int),
        ("time_accuracy", C.c_int), ("time_stepping", C.c_int),
        ("limiter", C.c_int * 3), ("tlimiter", C.c_int * 3), ("pb_switch", C.c_int * 3),
        ("accur", C.c_int), ("mu_variation", C.c_int),
        ("bc_id", C.c_int * 6), ("pbc_id", C.c_int * 6), ("dir_switch", C.c_int * 6), ("otherface", C.c_int * 6),
        ("plo", (C.c_int * 2) * 6), ("phi", (C.c_int * 2) * 6), ("pdir", (C.c_int * 2) * 6),
        ("block_id", C.c_int), ("n_blocks", C.c_int),
        ("CFL", C.c_double), ("global_time_step", C.c_double),
        ("gm", C.c_double), ("R_gas", C.c_double), ("mu_ref", C.c_double), ("T_ref", C.c_double),
        ("Sutherland_temp", C.c_double), ("Pr", C.c_double), ("tPr", C.c_double),
        ("density_inf", C.c_double), ("x_speed_inf", C.c_double), ("y_speed_inf", C.c_double),
        ("z_speed_inf", C.c_double), ("pressure_inf", C.c_double),
        ("tk_inf", C.c_double), ("tw_inf", C.c_double), ("vel_mag", C.c_double), ("MInf", C.c_double), ("tv_inf", C.c_double), ("tu_inf", C.c_double), ("tkl_inf", C.c_double), ("tgm_inf", C.c_double),
        ("fixed", (C.c_double * 6) * NFIX),
    ]


class Fest3dGpuError(C.Structure):
    _fields_ = [("flags", C.c_int), ("block_id", C.c_int), ("i", C.c_int), ("j", C.c_int), ("k", C.c_int), ("cuda_error", C.c_int)]


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".hpp"))]
    srcs.append(os.path.join(HERE, "..", "include", "fest3d_gpu.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", csrc, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load libfest3d_gpu.so (raises if it has not been built -- there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libfest3d_gpu.so is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, vp, ip = C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int)
    L.fest3d_gpu_create.argtypes = [C.POINTER(vp), C.POINTER(Fest3dGpuConfig), C.c_int]
    L.fest3d_gpu_destroy.argtypes = [vp]
    L.fest3d_gpu_set_stream.argtypes = [vp, vp]
    L.fest3d_gpu_sync.argtypes = [vp]
    L.fest3d_gpu_set_geometry.argtypes = [vp, dp, dp, dp, dp, dp]
    L.fest3d_gpu_set_state.argtypes = [vp, dp]
    L.fest3d_gpu_get_state.argtypes = [vp, dp]
    L.fest3d_gpu_step.argtypes = [vp, C.c_int, C.c_int, dp]
    L.fest3d_gpu_step_group.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, dp]
    L.fest3d_gpu_residual.argtypes = [vp, C.c_int, dp]
    L.fest3d_gpu_residual_group.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    L.fest3d_gpu_get_residue.argtypes = [vp, dp]
    L.fest3d_gpu_get_aux.argtypes = [vp, C.c_int, dp]
    L.fest3d_gpu_error.argtypes = [vp, C.POINTER(Fest3dGpuError)]
    L.fest3d_gpu_comm_unique_id.argtypes = [C.c_char_p]
    L.fest3d_gpu_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, ip]
    L.fest3d_gpu_link_local.argtypes = [vp, vp]
    L.fest3d_gpu_launch_count.argtypes = [vp]
    L.fest3d_gpu_launch_count.restype = C.c_longlong
    L.fest3d_gpu_kernel_timing.argtypes = [vp, C.c_int]
    L.fest3d_gpu_kernel_time_ms.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
    L.fest3d_gpu_kernel_time_ms.restype = C.c_double
    L.fest3d_gpu_gradient_time_ms.argtypes = [vp, C.POINTER(C.c_longlong), C.c_int]
    L.fest3d_gpu_gradient_time_ms.restype = C.c_double
    L.fest3d_gpu_gradient_path.argtypes = [vp]
    L.fest3d_gpu_set_state_async.argtypes = [vp, dp]
    L.fest3d_gpu_get_state_async.argtypes = [vp, dp]
    L.fest3d_gpu_state_wait.argtypes = [vp]
    L.fest3d_gpu_step_group_begin.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
    L.fest3d_gpu_step_group_end.argtypes = [C.POINTER(vp), C.c_int, dp]
    L.fest3d_gpu_version.restype = C.c_char_p
    L.fest3d_gpu_find_wall_dist.argtypes = [vp, dp, dp, C.c_longlong, dp, dp]
    L.fest3d_gpu_setup_geometry.argtypes = [vp, dp, dp, dp]
    L.fest3d_gpu_get_geometry.argtypes = [vp, dp, dp, dp, dp]
    L.fest3d_gpu_checkpoint_begin.argtypes = [vp, C.c_char_p, C.c_int]
    L.fest3d_gpu_checkpoint_wait.argtypes = [vp]
    L.fest3d_gpu_restart.argtypes = [vp, C.c_char_p, ip]
    _lib = L
    return L
