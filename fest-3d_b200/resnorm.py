"""Host side of find_resnorm that stays on the host in a drop-in: the library delivers Res_abs(0:n_var) per
iteration; this module keeps the reference's bookkeeping on top of it.

  * get_relative_resnorm   src/resnorm.f90:227-239   Res_save = Res_abs while current_iter <= Res_itr (= 3), Res_rel = Res_abs / Res_save
  * the named norms         src/resnorm.f90:256-360   what a `Res_list` entry of res_control.md writes
  * converged               src/convergence.f90:10-100  tolerance test on one named norm, only after iteration 10

Pure numpy; no device code.  (start_from /= 0 takes Res_save from the restart file's `previous_Res`, passed in as `previous_res`.)
"""
import numpy as np

RES_ITR = 3   # resnorm.f90:29


def _rss(v):
    return float(np.sqrt(np.sum(np.asarray(v, dtype=np.float64) ** 2)))


class ResnormHistory:
    def __init__(self, turbulence="none", previous_res=None):
        self.turbulence = turbulence
        self.res_abs = None
        self.res_rel = None
        self.res_save = None if previous_res is None else np.array(previous_res, dtype=np.float64)
        self._restart = previous_res is not None
        self.current_iter = 0

    def update(self, current_iter, res_abs):
        """One iteration's Res_abs(0:n_var) as fest3d_gpu_step delivers it."""
        self.current_iter = int(current_iter)
        self.res_abs = np.array(res_abs, dtype=np.float64)
        if not self._restart and self.current_iter <= RES_ITR:      # resnorm.f90:232
            self.res_save = self.res_abs.copy()
        if self.res_save is None:                                    # first call after iteration Res_itr without a restart record
            self.res_save = self.res_abs.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            self.res_rel = self.res_abs / self.res_save              # resnorm.f90:238
        return self.res_rel

    # ---- named norms: resnorm.f90:259-360 (write_resnorm) ----
    def named(self, name):
        a, r, t = self.res_abs, self.res_rel, self.turbulence
        two_eq = t in ("sst", "sst2003", "kkl")
        table = {
            "Mass_abs": lambda: a[0], "Resnorm_abs": lambda: _rss(a[1:]), "Viscous_abs": lambda: _rss(a[1:6]),
            "Turbulent_abs": lambda: _rss(a[6:]) if t != "none" else None,
            "Continuity_abs": lambda: a[1], "X_mom_abs": lambda: a[2], "Y_mom_abs": lambda: a[3], "Z_mom_abs": lambda: a[4], "Energy_abs": lambda: a[5],
            "Mass_rel": lambda: r[0], "Resnorm_rel": lambda: _rss(r[1:]), "Viscous_rel": lambda: _rss(r[1:6]),
            "Turbulent_rel": lambda: _rss(r[6:]) if t != "none" else None,
            "Continuity_rel": lambda: r[1], "X-mom_rel": lambda: r[2], "Y-mom_rel": lambda: r[3], "Z-mom_rel": lambda: r[4], "Energy_rel": lambda: r[5],
            "TKE_abs": lambda: a[6] if two_eq else None, "Tv_abs": lambda: a[6] if t in ("sa", "saBC") else None,
            "Omega_abs": lambda: a[7] if t in ("sst", "sst2003") else None, "Kl_abs": lambda: a[7] if t == "kkl" else None,
            "TKE_rel": lambda: r[6] if two_eq else None, "Tv_rel": lambda: r[6] if t in ("sa", "saBC") else None,
            "Omega_rel": lambda: r[7] if t in ("sst", "sst2003") else None, "Kl_rel": lambda: r[7] if t == "kkl" else None,
        }
        if name not in table:
            raise KeyError(name)
        v = table[name]()
        return None if v is None else float(v)

    def line(self, res_list, last_iter=0):
        """The values of one line of the residual file (resnorm.f90:256-258: iteration number, then the listed norms)."""
        return [self.current_iter + last_iter] + [self.named(n) for n in res_list if self.named(n) is not None]

    # ---- convergence.f90:10-100.  KEPT as in the reference: 'Z-mom_abs' tests Res_abs(3) and 'Y-mom_abs' Res_abs(4) (:31-34), likewise the
    # _rel pair (:50-53); an unknown tolerance type falls back to Resnorm_abs (:85-88).
    def converged(self, tolerance, tolerance_type):
        a, r = self.res_abs, self.res_rel
        table = {
            "Mass_abs": lambda: a[0], "Resnorm_abs": lambda: _rss(a[1:]), "Viscous_abs": lambda: _rss(a[1:6]), "Turbulent_abs": lambda: _rss(a[6:]),
            "Continuity_abs": lambda: a[1], "X-mom_abs": lambda: a[2], "Z-mom_abs": lambda: a[3], "Y-mom_abs": lambda: a[4], "Energy_abs": lambda: a[5],
            "Mass_rel": lambda: r[0], "Resnorm_rel": lambda: _rss(r[1:]), "Viscous_rel": lambda: _rss(r[1:6]), "Turbulent_rel": lambda: _rss(r[6:]),
            "Continuity_rel": lambda: r[1], "X-mom_rel": lambda: r[2], "Z-mom_rel": lambda: r[3], "Y-mom_rel": lambda: r[4], "Energy_rel": lambda: r[5],
            "TKE_abs": lambda: a[6], "tv_abs": lambda: a[6], "Dissipation_abs": lambda: a[7], "Omega_abs": lambda: a[7], "Kl_abs": lambda: a[7],
            "TKE_rel": lambda: r[6], "tv_rel": lambda: r[6], "Dissipation_rel": lambda: r[7], "Omega_rel": lambda: r[7], "Kl_rel": lambda: r[7],
        }
        check = float(table.get(tolerance_type, table["Resnorm_abs"])())
        return bool(check < tolerance and self.current_iter > 10)
