"""Readers for the reference's own case files and assembly of per-block set-ups.

The Fortran host keeps reading ``system/control.md``, ``fvscheme.md``, ``flow.md``, ``mesh/layout/layout.md``,
``mapping.txt``, ``periodic.txt``, ``mesh/bc/bc_NN.md`` and the grid files unchanged; this module is the stand-in
for that host in an image without a Fortran compiler.  It turns the reference's run-time *strings* into the
enums of the C ABI (include/fest3d_gpu.h) once, on the host.

Reference behaviour followed:
  * token reader: skip 3 header lines, skip blank / '#' lines     -- src/read_write/read/read.f90:40-75
  * control / scheme / flow token order                           -- read.f90:99-161, 187-242, 269-346
  * free-stream derived values, n_var                             -- src/state.f90:75-100, 280-322
  * layout.md rows                                                -- src/layout.f90:71-108
  * mapping.txt -> unpack ranges Pxlo/Pxhi/PxDir, dir_switch      -- src/mapping.f90:64-258
  * bc_NN.md fixed values                                         -- src/boundary/read_bc.f90:17-147
"""
from __future__ import annotations

import copy
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import geometry as geo

SCHEMES = {"van_leer": 0, "ldfss0": 1, "ausm": 2, "ausmP": 3, "ausmUP": 4, "slau": 5}
INTERPOLANTS = {"none": 0, "muscl": 1, "ppm": 2, "weno": 3, "weno_NM": 4}
TURBULENCE = {"none": 0, "sa": 1, "saBC": 2, "sst": 3, "sst2003": 4, "kkl": 5}
TRANSITION = {"none": 0, "bc": 1, "lctm2015": 2}
TIME_ACCURACY = {"none": 0, "RK2": 1, "RK4": 2, "TVDRK2": 3, "TVDRK3": 4, "implicit": 5, "plusgs": 6}
FIX_SLOTS = ["density", "pressure", "x_speed", "y_speed", "z_speed", "tk", "tw", "wall_temperature", "Tpressure", "Ttemperature", "tv", "tkl", "tgm"]
FIX_KEYS = {"FIX_DENSITY": 0, "FIX_PRESSURE": 1, "FIX_X_SPEED": 2, "FIX_Y_SPEED": 3, "FIX_Z_SPEED": 4,
            "FIX_tk": 5, "FIX_tw": 6, "WALL_TEMPERATURE": 7, "TOTAL_PRESSURE": 8, "TOTAL_TEMPERATURE": 9, "FIX_tv": 10, "FIX_tkl": 11, "FIX_tgm": 12}


def _tokens(path):
    """get_next_token (read.f90:40-75): drop the 3 header lines, then every non-blank, non-'#' line."""
    with open(path) as f:
        lines = f.read().splitlines()[3:]
    return [ln.strip() for ln in lines if ln.strip() and not ln.lstrip().startswith("#")]


@dataclass
class Scheme:
    scheme_name: str = "ausm"
    interpolant: str = "muscl"
    limiter: tuple = (1, 1, 1)
    pb_switch: tuple = (0, 0, 0)
    tlimiter: tuple = (1, 1, 1)
    turbulence: str = "none"
    transition: str = "none"
    time_stepping_method: str = "l"
    global_time_step: float = 1e-5
    time_step_accuracy: str = "none"
    accur: int = 0


@dataclass
class Flow:
    density_inf: float = 1.2
    x_speed_inf: float = 100.0
    y_speed_inf: float = 0.0
    z_speed_inf: float = 0.0
    pressure_inf: float = 101325.0
    tu_inf: float = 1.0
    mu_ratio_inf: float = 1.0
    tgm_inf: float = 1.0
    mu_ref: float = 0.0
    mu_variation: str = "constant"
    T_ref: float = 300.0
    Sutherland_temp: float = 110.0
    Pr: float = 0.7
    tPr: float = 0.9
    gm: float = 1.4
    R_gas: float = 287.0
    # derived (state.f90:75-100)
    vel_mag: float = 0.0
    MInf: float = 0.0
    tk_inf: float = 0.0
    tw_inf: float = 0.0
    tv_inf: float = 0.0
    tkl_inf: float = 0.0

    def derive(self, turbulence):
        self.vel_mag = math.sqrt(self.x_speed_inf ** 2 + self.y_speed_inf ** 2 + self.z_speed_inf ** 2)
        self.MInf = self.vel_mag / math.sqrt(self.gm * self.pressure_inf / self.density_inf)
        if turbulence in ("sst", "sst2003"):
            ti = self.tu_inf / 100
            self.tk_inf = 1.5 * ((self.vel_mag * ti) ** 2)
            self.tw_inf = self.density_inf * self.tk_inf / (self.mu_ref * self.mu_ratio_inf)
        if turbulence == "sa":      # state.f90:105-106
            self.tv_inf = self.mu_ratio_inf * self.mu_ref / self.density_inf
        if turbulence == "kkl":     # state.f90:101-103
            c_inf = math.sqrt(self.gm * self.pressure_inf / self.density_inf)
            self.tk_inf = 9 * (1e-9) * (c_inf ** 2)
            self.tkl_inf = 1.5589 * (1e-6) * (self.mu_ref * c_inf) / self.density_inf
        return self


@dataclass
class Control:
    CFL: float = 1.0
    start_from: int = 0
    max_iters: int = 1
    checkpoint_iter: int = 0
    res_write_interval: int = 10
    tolerance: float = 1e-14
    tolerance_type: str = "abs"


def read_control(path):
    t = _tokens(path)
    c = Control()
    c.CFL = float(t[0]); c.start_from = int(t[1]); c.max_iters = int(t[2]); c.checkpoint_iter = int(t[3])
    c.res_write_interval = int(t[10])
    tol = t[11].split()
    c.tolerance = float(tol[0]); c.tolerance_type = tol[1] if len(tol) > 1 else "abs"
    return c


def read_scheme(path):
    t = _tokens(path)
    s = Scheme()
    s.scheme_name = t[0]; s.interpolant = t[1]
    sw = [int(v) for v in t[2].split()]
    s.limiter = tuple(sw[0:3]); s.pb_switch = tuple(sw[3:6])
    s.tlimiter = tuple(int(v) for v in t[3].split())
    s.turbulence = t[4]; s.transition = t[5]
    ts = t[6].split()
    s.time_stepping_method = ts[0]
    if ts[0] == "g" and len(ts) > 1:
        s.global_time_step = float(ts[1])
    s.time_step_accuracy = t[7]
    s.accur = int(t[8])
    return s


def read_flow(path):
    t = _tokens(path)
    f = Flow()
    (f.density_inf, f.x_speed_inf, f.y_speed_inf, f.z_speed_inf, f.pressure_inf, f.tu_inf, f.mu_ratio_inf,
     f.tgm_inf, f.mu_ref) = (float(v) for v in t[1:10])
    f.mu_variation = t[10]
    f.T_ref = float(t[11]); f.Sutherland_temp = float(t[12])
    pr = t[13].split(); f.Pr = float(pr[0]); f.tPr = float(pr[1])
    f.gm = float(t[14]); f.R_gas = float(t[15])
    return f


def read_layout(path):
    """layout.f90:71-108 -> list of (gridfile, bcfile, [imin..kmax ids])."""
    rows = [ln.split() for ln in open(path).read().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    nproc = int(rows[0][0])
    out = []
    for r in rows[2:2 + nproc]:
        out.append((r[1], r[2], [int(v) for v in r[3:9]]))
    return out


def n_var_of(turbulence, transition="none"):
    nv = {"none": 5, "sa": 6, "saBC": 6}.get(turbulence, 7)   # state.f90:291-306
    return nv + (1 if transition == "lctm2015" else 0)


@dataclass
class BlockSetup:
    """Everything one rank of the reference owns before the first call of get_next_solution."""
    imx: int
    jmx: int
    kmx: int
    bc_id: list
    scheme: Scheme
    flow: Flow
    control: Control
    block_id: int = 0
    n_blocks: int = 1
    pbc_id: list = field(default_factory=lambda: [-1] * 6)
    dir_switch: list = field(default_factory=lambda: [0] * 6)
    otherface: list = field(default_factory=lambda: [2, 1, 4, 3, 6, 5])
    plo: list = None      # [6][2]
    phi: list = None
    pdir: list = None
    fixed: np.ndarray = None   # [len(FIX_SLOTS), 6]
    nodes: np.ndarray = None
    cells: np.ndarray = None
    Ifaces: np.ndarray = None
    Jfaces: np.ndarray = None
    Kfaces: np.ndarray = None
    dist: np.ndarray = None
    qp: np.ndarray = None      # [n_var, kmx+5, jmx+5, imx+5]

    @property
    def n_var(self):
        return n_var_of(self.scheme.turbulence, self.scheme.transition)

    def default_maps(self):
        """mapping.f90:85-103 defaults + change_map_to_particular_range for an identity-oriented neighbour."""
        tr = [(self.jmx, self.kmx), (self.jmx, self.kmx), (self.imx, self.kmx), (self.imx, self.kmx),
              (self.imx, self.jmx), (self.imx, self.jmx)]
        self.plo = [[1, 1] for _ in range(6)]
        self.phi = [[a - 1, b - 1] for a, b in tr]
        self.pdir = [[1, 1] for _ in range(6)]

    def fill_fixed_defaults(self):
        """read_bc.f90:121-160 fill_fixed_values."""
        f = self.flow
        self.fixed = np.zeros((len(FIX_SLOTS), 6))
        self.fixed[0, :] = f.density_inf; self.fixed[1, :] = f.pressure_inf
        self.fixed[2, :] = f.x_speed_inf; self.fixed[3, :] = f.y_speed_inf; self.fixed[4, :] = f.z_speed_inf
        self.fixed[5, :] = f.tk_inf; self.fixed[6, :] = f.tw_inf; self.fixed[10, :] = f.tv_inf; self.fixed[11, :] = f.tkl_inf; self.fixed[12, :] = f.tgm_inf

    def init_state(self):
        """state.f90:193-247 init_state_with_infinity_values (ghosts included)."""
        f = self.flow
        nv = self.n_var
        q = np.empty((nv, self.kmx + 5, self.jmx + 5, self.imx + 5))
        q[0] = f.density_inf; q[1] = f.x_speed_inf; q[2] = f.y_speed_inf; q[3] = f.z_speed_inf; q[4] = f.pressure_inf
        if nv >= 7:
            q[5] = f.tk_inf; q[6] = f.tkl_inf if self.scheme.turbulence == "kkl" else f.tw_inf     # state.f90:228-238
        elif nv == 6:               # state.f90:240-242
            q[5] = f.tv_inf
        if self.scheme.transition == "lctm2015":    # state.f90:260-266: the intermittency is the last variable
            q[nv - 1] = f.tgm_inf
        self.qp = q

    def build_geometry(self):
        self.cells, self.Ifaces, self.Jfaces, self.Kfaces = geo.compute_geometry(self.nodes, self.bc_id)


def _map_range(lo, hi):
    """mapping.f90:185-258 change_map_to_particular_range for one transverse axis: node range -> cell loop."""
    plo, phi, pdir = lo, hi, 1
    if lo == 1:
        plo = 1
    if hi == 1:
        phi = 1; pdir = -1
    if lo > 1:
        plo = lo - 1; pdir = -1
    if hi > 1:
        phi = hi - 1
    return plo, phi, pdir


def read_bc_file(path, blk):
    """read_bc.f90:28-120: six '# face' sections, '- NAME [value]' lines."""
    blk.fill_fixed_defaults()
    lines = open(path).read().splitlines()[3:]
    face = 0
    for ln in lines:
        if ln.startswith("#"):
            face += 1
            if face > 6:
                break
            continue
        if ln.startswith("- ") and face >= 1:
            parts = ln[2:].split()
            if len(parts) >= 2 and parts[0] in FIX_KEYS:
                try:
                    blk.fixed[FIX_KEYS[parts[0]], face - 1] = float(parts[1])
                except ValueError:
                    pass


def load_case(case_dir, scheme_override=None, control_override=None, flow_override=None):
    """Read a reference case directory (e.g. tests/SmoothBump) into a list of BlockSetup (one per block)."""
    sysd = os.path.join(case_dir, "system")
    scheme = read_scheme(os.path.join(sysd, "fvscheme.md"))
    control = read_control(os.path.join(sysd, "control.md"))
    flow = read_flow(os.path.join(sysd, "flow.md"))
    for obj, ov in ((scheme, scheme_override), (control, control_override), (flow, flow_override)):
        for k, v in (ov or {}).items():
            setattr(obj, k, v)
    flow.derive(scheme.turbulence)
    layout = read_layout(os.path.join(sysd, "mesh", "layout", "layout.md"))
    blocks = []
    for b, (gridfile, bcfile, ids) in enumerate(layout):
        nodes_int = geo.read_grid(os.path.join(sysd, "mesh", "gridfiles", gridfile))
        kmx, jmx, imx, _ = nodes_int.shape
        blk = BlockSetup(imx=imx, jmx=jmx, kmx=kmx, bc_id=list(ids), scheme=copy.copy(scheme), flow=copy.copy(flow),
                         control=copy.copy(control), block_id=b, n_blocks=len(layout))
        blk.nodes = geo.ghost_grid(nodes_int)
        blk.default_maps()
        read_bc_file(os.path.join(sysd, "mesh", "bc", bcfile), blk)
        blocks.append(blk)
    # mapping.txt (mapping.f90:118-177): rows b1 f1 s11 e11 s12 e12 b2 f2 s21 e21 s22 e22 dir_switch class
    mp = os.path.join(sysd, "mesh", "layout", "mapping.txt")
    if os.path.exists(mp):
        for ln in open(mp).read().splitlines()[1:]:
            t = ln.split()
            if len(t) < 13:
                continue
            b1, f1 = int(t[0]), int(t[1])
            f2, s21, e21, s22, e22, sw = int(t[7]), int(t[8]), int(t[9]), int(t[10]), int(t[11]), int(t[12])
            blk = blocks[b1]
            blk.otherface[f1 - 1] = f2
            blk.dir_switch[f1 - 1] = sw
            a = _map_range(s21, e21); bb = _map_range(s22, e22)
            blk.plo[f1 - 1] = [a[0], bb[0]]; blk.phi[f1 - 1] = [a[1], bb[1]]; blk.pdir[f1 - 1] = [a[2], bb[2]]
    pp = os.path.join(sysd, "mesh", "layout", "periodic.txt")
    if os.path.exists(pp):
        for ln in open(pp).read().splitlines()[1:]:
            t = ln.split()
            if len(t) >= 4:
                blocks[int(t[0])].pbc_id[int(t[2]) - 1] = int(t[1])
    # geometry, wall distance, state
    for blk in blocks:
        blk.build_geometry()
    if scheme.turbulence != "none":
        wall = np.concatenate([geo.surface_nodes(blk.nodes, blk.bc_id) for blk in blocks], axis=0)
        for blk in blocks:
            blk.dist = geo.wall_distance(blk.nodes, wall)
    for blk in blocks:
        blk.init_state()
    return blocks


def merge_blocks_i(blocks):
    """Concatenate blocks that abut in i into one block (BASELINE config 1 'single block' variant): drops the
    duplicated interface node plane; outer BC ids are kept."""
    b0 = blocks[0]
    ints = []
    for n, blk in enumerate(blocks):
        ni = blk.nodes[3:3 + blk.kmx, 3:3 + blk.jmx, 3:3 + blk.imx]
        ints.append(ni if n == 0 else ni[:, :, 1:])
    nodes_int = np.concatenate(ints, axis=2)
    kmx, jmx, imx, _ = nodes_int.shape
    ids = list(b0.bc_id)
    ids[1] = blocks[-1].bc_id[1]
    out = BlockSetup(imx=imx, jmx=jmx, kmx=kmx, bc_id=ids, scheme=copy.copy(b0.scheme), flow=copy.copy(b0.flow),
                     control=copy.copy(b0.control), block_id=0, n_blocks=1)
    out.nodes = geo.ghost_grid(nodes_int)
    out.default_maps()
    out.fill_fixed_defaults()
    out.fixed[:, 1] = blocks[-1].fixed[:, 1]
    out.build_geometry()
    if b0.scheme.turbulence != "none":
        out.dist = geo.wall_distance(out.nodes, geo.surface_nodes(out.nodes, out.bc_id))
    out.init_state()
    return out


def read_tecplot_state(path, blk, var_order=("u", "v", "w", "Density", "Pressure")):
    """Restart / result reader (read_output_tec.f90:43-190): ASCII block format, three nodal blocks then one
    cell-centred block per variable over interior cells, i fastest.  Only u,v,w,Density,Pressure are taken
    (what the shipped cases list in output_control.md); the interior of ``blk.qp`` is overwritten."""
    with open(path) as f:
        lines = f.read().splitlines()
    start = next(n for n, ln in enumerate(lines) if ln.strip().upper().startswith("SOLUTIONTIME")) + 1
    vals = np.array(" ".join(lines[start:]).split(), dtype=np.float64)
    nn = blk.imx * blk.jmx * blk.kmx
    nc = (blk.imx - 1) * (blk.jmx - 1) * (blk.kmx - 1)
    slot = {"Density": 0, "u": 1, "v": 2, "w": 3, "Pressure": 4}
    off = 3 * nn
    for name in var_order:
        a = vals[off:off + nc].reshape(blk.kmx - 1, blk.jmx - 1, blk.imx - 1)
        blk.qp[slot[name], 3:3 + blk.kmx - 1, 3:3 + blk.jmx - 1, 3:3 + blk.imx - 1] = a
        off += nc
    return blk
