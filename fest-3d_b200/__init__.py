"""fest3d-b200: B200-native explicit residual-evaluation + time-update path of FEST-3D.

The package holds the CUDA kernels + C ABI (csrc/, libfest3d_gpu.so) and the host-side mirror of the
reference interface for this path (solver.py: get_next_solution / find_resnorm), plus the stand-in host
that reads the reference's own case files (case.py, geometry.py).
"""
__all__ = ["case", "geometry"]
