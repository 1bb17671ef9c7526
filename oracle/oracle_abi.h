/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * C ABI of the CPU oracle: a literal C++ re-statement of FEST-3D's explicit
 * residual-evaluation + time-update path (reference: src/update.f90 and the
 * modules it calls).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED against a runnable reference: the container has no Fortran
 * compiler and no MPI, so the reference cannot be executed here.  The oracle is
 * pinned against every known-answer value the reference's own unit tests hold
 * for this path (tests/test_*.f90, see tests/test_oracle_kat.py).
 */
#ifndef FEST3D_ORACLE_ABI_H
#define FEST3D_ORACLE_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

/* enums <- strings (reference: scheme.f90:92-104, face_interpolant.f90:91-101,
 * update.f90:171-225, time.f90:323-326) */
enum { ORC_VAN_LEER = 0, ORC_LDFSS0 = 1, ORC_AUSM = 2, ORC_AUSMP = 3, ORC_AUSMUP = 4, ORC_SLAU = 5 };
enum { ORC_NONE = 0, ORC_MUSCL = 1, ORC_PPM = 2, ORC_WENO = 3, ORC_WENO_NM = 4 };
enum { ORC_TURB_NONE = 0, ORC_TURB_SA = 1, ORC_TURB_SST = 3, ORC_TURB_SST2003 = 4, ORC_TURB_KKL = 5 };
enum { ORC_T_NONE = 0, ORC_T_RK2 = 1, ORC_T_RK4 = 2, ORC_T_TVDRK2 = 3, ORC_T_TVDRK3 = 4, ORC_T_IMPLICIT = 5 };

/* fixed-value slots (reference: vartypes.f90:307-334) */
enum {
  ORC_FIX_DENSITY = 0, ORC_FIX_PRESSURE, ORC_FIX_X_SPEED, ORC_FIX_Y_SPEED, ORC_FIX_Z_SPEED,
  ORC_FIX_TK, ORC_FIX_TW, ORC_FIX_WALL_TEMP, ORC_FIX_TPRESSURE, ORC_FIX_TTEMPERATURE, ORC_FIX_TV, ORC_FIX_TKL, ORC_FIX_TGM,
  ORC_NFIX
};

typedef struct {
  int imx, jmx, kmx, n_var;          /* node counts; n_var 5, 6 (sa) or 7, +1 with transition = lctm2015 */
  int scheme, interpolant, turbulence, transition;
  int time_accuracy;                 /* ORC_T_* */
  int time_stepping;                 /* 0 = 'l' local, 1 = 'g' global */
  int limiter[3];                    /* i,j,k limiter_switch */
  int tlimiter[3];                   /* i,j,k tlimiter_switch */
  int pb_switch[3];                  /* i,j,k pressure-based switching (muscl, ppm) */
  int accur;                         /* higher-order BC switch */
  int mu_variation;                  /* 0 constant, 1 sutherland_law */
  int bc_id[6];                      /* imin,imax,jmin,jmax,kmin,kmax; <0 physical, >=0 neighbour block */
  int pbc_id[6];                     /* periodic partner block or -1 */
  int dir_switch[6];
  int otherface[6];                  /* 1..6 face of the neighbour that is attached */
  int plo[6][2], phi[6][2], pdir[6][2]; /* unpack ranges of the two transverse axes (mapping.f90:185-258),
                                            axis order: faces 1,2 -> (j,k); 3,4 -> (i,k); 5,6 -> (i,j) */
  int block_id, n_blocks;
  double CFL, global_time_step;
  double gm, R_gas, mu_ref, T_ref, Sutherland_temp, Pr, tPr;
  double density_inf, x_speed_inf, y_speed_inf, z_speed_inf, pressure_inf;
  double tk_inf, tw_inf, vel_mag, MInf;
  double tv_inf;
  double tu_inf;                     /* percent */
  double tkl_inf;                    /* free-stream kL of the k-kL model (state.f90:101-103) */
  double tgm_inf;                    /* free-stream intermittency of transition = lctm2015 (vartypes.f90:258, default 1) */
  double fixed[ORC_NFIX][6];
} OracleConfig;

typedef struct OracleWorld OracleWorld;

/* wall_dist.f90:84-131 (test infrastructure for the device wall-distance kernel) */
void oracle_find_wall_dist(int imx, int jmx, int kmx, const double* nodes, const double* wall, long long n_wall, double* dist_out);

/* A world is the set of blocks (= MPI ranks of the reference) stepped in lock step. */
OracleWorld* oracle_create(int n_blocks, const OracleConfig* cfgs);
void oracle_destroy(OracleWorld* w);

/* Host arrays use the reference's Fortran layouts:
 *   cells  (-2:imx+2,-2:jmx+2,-2:kmx+2) of {volume,cx,cy,cz}
 *   Ifaces (-2:imx+3,-2:jmx+2,-2:kmx+2) of {A,nx,ny,nz}; Jfaces, Kfaces alike
 *   dist   (-2:imx+2,-2:jmx+2,-2:kmx+2)          (may be NULL when turbulence == none)
 *   qp     (-2:imx+2,-2:jmx+2,-2:kmx+2,1:n_var)                                   */
int oracle_set_geometry(OracleWorld* w, int b, const double* cells, const double* Ifaces,
                        const double* Jfaces, const double* Kfaces, const double* dist);
int oracle_set_state(OracleWorld* w, int b, const double* qp);
int oracle_get_state(OracleWorld* w, int b, double* qp);

/* One get_total_conservative_Residue on every block (update.f90:495-547), Temp refreshed
 * first as get_next_solution does (update.f90:170).  Residue is (1:imx-1,1:jmx-1,1:kmx-1,1:n_var). */
int oracle_residual(OracleWorld* w, int current_iter);
int oracle_get_residue(OracleWorld* w, int b, double* residue);

/* One iteration: get_next_solution + find_resnorm (solver.f90:184-185).
 * res_abs receives n_var+1 assembled norms (resnorm.f90:211-225). Returns 0, or an error class. */
int oracle_step(OracleWorld* w, int current_iter, double* res_abs);

/* which: 0 delta_t (interior), 1 mu, 2 mu_t, 3 sst_F1 (-2:imx+2 ...), 4 Temp,
 * 10..15 x_qp_left, x_qp_right, y_.., z_.. ; 20,21,22 F,G,H ; 30+3*c+d gradqp_d(:,:,:,c) */
int oracle_get_aux(OracleWorld* w, int b, int which, double* out);

/* Stand-alone kernels for the reference's unit-test known answers (tests/test_*.f90). */
void oracle_kat_flux(int scheme, int n_var, double gm, double MInf, const double* left, const double* right,
                     const double* face /*A,nx,ny,nz*/, int mask, double* flux);
void oracle_kat_states(int interpolant, int n, const double* q /*cells -2..n-3*/, const double* vol,
                       int limiter, double* left, double* right);
/* compute_residue (scheme.f90:111-141) + the mass imbalance of get_absolute_resnorm (resnorm.f90:190-198) on given F, G, H */
void oracle_kat_residue(int imx, int jmx, int kmx, int n_var, const double* F, const double* G, const double* H,
                        double* residue, double* merror);

#ifdef __cplusplus
}
#endif
#endif
