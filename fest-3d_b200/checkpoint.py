"""Host-side reader / writer of the binary checkpoint side format the library writes (include/fest3d_gpu.h,
fest3d_gpu_checkpoint_begin; csrc/checkpoint.cu): a 64-byte header

    "F3DCKPT1" | int32 imx, jmx, kmx, n_var, iter | 3 x int32 0 | uint64 n_doubles | 16 bytes 0

followed by qp(-2:imx+2, -2:jmx+2, -2:kmx+2, 1:n_var) as little-endian float64 -- in this package's array convention
qp[n_var, kmx+5, jmx+5, imx+5].  Post-processing (the reference's writers, src/read_write/write/*.f90, take the same array) and tests
read checkpoints with this; a host that has qp in memory can also write one for fest3d_gpu_restart."""
from __future__ import annotations

import numpy as np

MAGIC = b"F3DCKPT1"
HEADER = np.dtype([("magic", "S8"), ("imx", "<i4"), ("jmx", "<i4"), ("kmx", "<i4"), ("n_var", "<i4"), ("iter", "<i4"),
                   ("zero", "<i4", (3,)), ("n_doubles", "<u8"), ("pad", "S16")])
assert HEADER.itemsize == 64


def read_checkpoint(path):
    """-> (header dict, qp[n_var, kmx+5, jmx+5, imx+5]).  Raises ValueError on a foreign or truncated file."""
    with open(path, "rb") as f:
        raw = f.read(HEADER.itemsize)
        if len(raw) != HEADER.itemsize:
            raise ValueError("%s: shorter than a checkpoint header" % path)
        h = np.frombuffer(raw, dtype=HEADER)[0]
        if h["magic"] != MAGIC:
            raise ValueError("%s: not a fest3d_gpu checkpoint" % path)
        imx, jmx, kmx, nv = int(h["imx"]), int(h["jmx"]), int(h["kmx"]), int(h["n_var"])
        n = nv * (kmx + 5) * (jmx + 5) * (imx + 5)
        if int(h["n_doubles"]) != n:
            raise ValueError("%s: header says %d doubles, the extents give %d" % (path, int(h["n_doubles"]), n))
        q = np.fromfile(f, dtype="<f8", count=n)
        if q.size != n:
            raise ValueError("%s: truncated (%d of %d doubles)" % (path, q.size, n))
    hdr = dict(imx=imx, jmx=jmx, kmx=kmx, n_var=nv, iter=int(h["iter"]))
    return hdr, q.reshape(nv, kmx + 5, jmx + 5, imx + 5)


def write_checkpoint(path, qp, it):
    """qp[n_var, kmx+5, jmx+5, imx+5] (ghost cells included) -> a file fest3d_gpu_restart accepts."""
    q = np.ascontiguousarray(qp, dtype="<f8")
    nv, nk, nj, ni = q.shape
    h = np.zeros(1, dtype=HEADER)
    h["magic"] = MAGIC
    h["imx"], h["jmx"], h["kmx"], h["n_var"], h["iter"] = ni - 5, nj - 5, nk - 5, nv, it
    h["n_doubles"] = q.size
    with open(path, "wb") as f:
        f.write(h.tobytes())
        f.write(q.tobytes())
