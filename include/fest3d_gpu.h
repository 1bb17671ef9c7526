/* fest3d_gpu.h -- C ABI of the B200 (sm_100a) explicit residual-evaluation + time-update path of FEST-3D.
 *
 * Drop-in boundary (reference file:line each entry point replaces):
 *   fest3d_gpu_create        <- setup_update / setup_interface / setup_scheme / setup_time / setup_gradients /
 *                               setup_viscosity / setup_bc (flux-zero masks, c1,c2,c3)   src/solver.f90:115-128,
 *                               src/update.f90:85-125, src/boundary/bc.f90:48-66
 *   fest3d_gpu_set_geometry  <- the cells / Ifaces / Jfaces / Kfaces / dist arguments of get_next_solution
 *                               src/update.f90:129-147, src/wall/wall_dist.f90:22
 *   fest3d_gpu_set_state / fest3d_gpu_get_state
 *                            <- the qp argument (in/out) of get_next_solution; checkpoint() reads it back
 *                               src/solver.f90:139,184,186
 *   fest3d_gpu_step          <- call get_next_solution(...) ; call find_resnorm(...)      src/solver.f90:184-185
 *                               (src/update.f90:129-226, src/resnorm.f90:62-91 incl. the MPI_ALLGATHER :207)
 *   fest3d_gpu_residual      <- get_total_conservative_Residue                            src/update.f90:495-547
 *   fest3d_gpu_comm_*        <- MPI_SENDRECV halo swap of apply_interface                 src/interface1.f90:96-493
 *   fest3d_gpu_error         <- Fatal_error (message + STOP)                              src/error.h:1
 *
 * All host arrays use the reference's Fortran layouts (column-major, lower bound -2, AoS records of 4 doubles):
 *   cells  (-2:imx+2,-2:jmx+2,-2:kmx+2) of {volume,centerx,centery,centerz}   src/vartypes.f90:29-36
 *   Ifaces (-2:imx+3,-2:jmx+2,-2:kmx+2) of {A,nx,ny,nz}; Jfaces/Kfaces alike  src/vartypes.f90:39-46
 *   dist   (-2:imx+2,-2:jmx+2,-2:kmx+2)
 *   qp     (-2:imx+2,-2:jmx+2,-2:kmx+2,1:n_var)
 *   residue(1:imx-1,1:jmx-1,1:kmx-1,1:n_var)
 * The host keeps ownership of its arrays; the library owns device mirrors.  qp on the host is stale between
 * fest3d_gpu_get_state calls.  Every function returns 0 on success; there is no CPU fallback: a missing device or a
 * kernel failure is an error return, never a silent host path.  Calls on one context must be serialised by the caller.
 */
#ifndef FEST3D_GPU_H
#define FEST3D_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

/* enums <- run-time strings of the reference (src/face/flux/convective/scheme.f90:92-104,
 * src/face/state/face_interpolant.f90:91-101, src/update.f90:171-225, src/time.f90:323-326) */
enum { F3D_VAN_LEER = 0, F3D_LDFSS0 = 1, F3D_AUSM = 2, F3D_AUSMP = 3, F3D_AUSMUP = 4, F3D_SLAU = 5 };
enum { F3D_INTERP_NONE = 0, F3D_MUSCL = 1, F3D_PPM = 2, F3D_WENO = 3, F3D_WENO_NM = 4 };
enum { F3D_TURB_NONE = 0, F3D_TURB_SA = 1, F3D_TURB_SABC = 2, F3D_TURB_SST = 3, F3D_TURB_SST2003 = 4, F3D_TURB_KKL = 5 };
enum { F3D_TRANS_NONE = 0, F3D_TRANS_BC = 1, F3D_TRANS_LCTM2015 = 2 };
enum { F3D_T_NONE = 0, F3D_T_RK2 = 1, F3D_T_RK4 = 2, F3D_T_TVDRK2 = 3, F3D_T_TVDRK3 = 4, F3D_T_IMPLICIT = 5, F3D_T_PLUSGS = 6 };

/* slots of the per-face fixed values (src/vartypes.f90:307-334, src/boundary/read_bc.f90:28-147) */
enum {
  F3D_FIX_DENSITY = 0, F3D_FIX_PRESSURE, F3D_FIX_X_SPEED, F3D_FIX_Y_SPEED, F3D_FIX_Z_SPEED,
  F3D_FIX_TK, F3D_FIX_TW, F3D_FIX_WALL_TEMP, F3D_FIX_TPRESSURE, F3D_FIX_TTEMPERATURE, F3D_FIX_TV, F3D_FIX_TKL, F3D_FIX_TGM,
  F3D_NFIX
};

/* error classes (bit mask), the Fatal_error sites of the path */
enum {
  F3D_ERR_NAN_FLUX = 1,      /* any(isnan(F/G/H))        ausm.f90:195-213, viscous.f90:125-133 */
  F3D_ERR_NAN_GRADIENT = 2,  /* any(isnan(grad))         gradients.f90:478 */
  F3D_ERR_NAN_VISCOSITY = 4, /* any(isnan(mu))           viscosity.f90:542 */
  F3D_ERR_NEGATIVE_STATE = 8,/* rho<0, p<0 or NaN after update   update.f90:448-452 */
  F3D_ERR_GEOMETRY = 16,      /* non-positive cell volume  geometry.f90:476-494 (fest3d_gpu_setup_geometry only) */
  F3D_ERR_IO = 32,            /* checkpoint file cannot be written / read (fest3d_gpu_checkpoint_*, fest3d_gpu_restart) */
  F3D_ERR_UNSUPPORTED = 64,  /* plusgs / saBC (and kkl, lctm2015 or a viscous implicit run with F3D_GRADIENTS=fused) */
  F3D_ERR_CUDA = 128,
  F3D_ERR_ARGUMENT = 256,    /* also: an interface / periodic face that was neither linked locally nor given a communicator */
  F3D_ERR_PEER = 512         /* another rank reported an error in this call: every rank returns (Fatal_error stops the whole job) */
};

typedef struct {
  int imx, jmx, kmx, n_var;            /* node counts of the block; n_var 5 (none), 6 (sa) or 7 (sst, sst2003, kkl); 8 = sst/sst2003 with
                                          transition = lctm2015                                   state.f90:291-320 */
  int scheme, interpolant, turbulence, transition;
  int time_accuracy;                   /* F3D_T_*                                                update.f90:171 */
  int time_stepping;                   /* 0 = 'l' local, 1 = 'g' global                          time.f90:323-326 */
  int limiter[3];                      /* i,j,k limiter_switch                                   vartypes.f90:206-211 */
  int tlimiter[3];                     /* i,j,k turbulent limiter switch */
  int pb_switch[3];                    /* i,j,k pressure-based switching of muscl / ppm (ignored by the other interpolants)
                                                                                                 muscl.f90:37-112, ppm.f90:108-170 */
  int accur;                           /* higher-order boundary switch -> c1,c2,c3               bc.f90:48-50 */
  int mu_variation;                    /* 0 constant, 1 sutherland_law                           viscosity.f90:109-138 */
  int bc_id[6];                        /* imin,imax,jmin,jmax,kmin,kmax: <0 physical BC id, >=0 neighbour block */
  int pbc_id[6];                       /* periodic partner block or -1                           mapping.f90:261-288 */
  int dir_switch[6];                   /*                                                        mapping.f90:118-177 */
  int otherface[6];                    /* face (1..6) of the neighbour attached to each face */
  int plo[6][2], phi[6][2], pdir[6][2];/* unpack loop ranges of the two transverse axes          mapping.f90:185-258;
                                          axis order: faces 1,2 -> (j,k); 3,4 -> (i,k); 5,6 -> (i,j) */
  int block_id, n_blocks;              /* == process_id, total_process of the reference */
  double CFL, global_time_step;
  double gm, R_gas, mu_ref, T_ref, Sutherland_temp, Pr, tPr;
  double density_inf, x_speed_inf, y_speed_inf, z_speed_inf, pressure_inf;
  double tk_inf, tw_inf, vel_mag, MInf;
  double tv_inf;                       /* free-stream nu-tilde of the SA model                   state.f90:105-106 */
  double tu_inf;                       /* free-stream turbulence intensity in percent (transition = bc)   source.f90:579,1164 */
  double tkl_inf;                      /* free-stream kL of the k-kL model                       state.f90:101-103 */
  double tgm_inf;                      /* free-stream intermittency (transition = lctm2015)       vartypes.f90:258, state.f90:265 */
  double fixed[F3D_NFIX][6];           /* fixed_density(6), fixed_pressure(6) ...                read_bc.f90 */
} Fest3dGpuConfig;

typedef struct {
  int flags;        /* OR of F3D_ERR_* */
  int block_id;
  int i, j, k;      /* first offending cell (Fortran indices) when known, else 0 */
  int cuda_error;   /* cudaError_t of the last failing runtime call, else 0 */
} Fest3dGpuError;

typedef struct Fest3dGpuCtx Fest3dGpuCtx;

/* life cycle -------------------------------------------------------------------------------------------------- */
int fest3d_gpu_create(Fest3dGpuCtx** ctx, const Fest3dGpuConfig* cfg, int device);
int fest3d_gpu_destroy(Fest3dGpuCtx* ctx);
/* run the context on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL -> the context's own stream */
int fest3d_gpu_set_stream(Fest3dGpuCtx* ctx, void* cuda_stream);
int fest3d_gpu_sync(Fest3dGpuCtx* ctx);

/* data in / out (host pointers, reference layouts) ------------------------------------------------------------- */
int fest3d_gpu_set_geometry(Fest3dGpuCtx* ctx, const double* cells, const double* Ifaces, const double* Jfaces,
                            const double* Kfaces, const double* dist /* may be NULL without turbulence */);
int fest3d_gpu_set_state(Fest3dGpuCtx* ctx, const double* qp);
int fest3d_gpu_get_state(Fest3dGpuCtx* ctx, double* qp);
/* Asynchronous, full-duplex forms (pinned host memory gives the overlap): set_state_async starts the upload on a copy stream and
 * returns -- the state takes effect at the next fest3d_gpu_step / fest3d_gpu_residual call, which is stream-ordered behind it, so the
 * upload of the next state overlaps the iterations still running and the download of the previous result; get_state_async snapshots
 * qp in stream order (after every step issued so far; an upload still pending for the next step is NOT applied by it) and downloads
 * it on a second copy stream; the host buffers may be reused /
 * read after fest3d_gpu_state_wait.  The checkpoint host of src/solver.f90:139,186 only needs the outbound one. */
int fest3d_gpu_set_state_async(Fest3dGpuCtx* ctx, const double* qp);
int fest3d_gpu_get_state_async(Fest3dGpuCtx* ctx, double* qp);
int fest3d_gpu_state_wait(Fest3dGpuCtx* ctx);
/* SURVEY 8(f) rank 1 -- find_wall_dist (src/wall/wall_dist.f90:84-131) on the device: minimum distance of every node
 * nodes(-2:imx+3,-2:jmx+3,-2:kmx+3) (nodetype records x,y,z) to the n_wall surface nodes wall_xyz[n_wall][3] (the contents of
 * the surface-node file, wall_dist.f90:74-82), averaged over the eight nodes of each cell.  Fills the context's wall-distance
 * field (so it replaces the `dist` argument of fest3d_gpu_set_geometry; call it after set_geometry) and, if dist_out != NULL,
 * returns dist(-2:imx+2,-2:jmx+2,-2:kmx+2).  kernel_ms (may be NULL) receives the device time of the node kernel. */
int fest3d_gpu_find_wall_dist(Fest3dGpuCtx* ctx, const double* nodes, const double* wall_xyz, long long n_wall, double* dist_out,
                              double* kernel_ms);

/* SURVEY 8(f) rank 2 -- grid ghost layers and metrics on the device instead of fest3d_gpu_set_geometry's upload of the 16 metric
 * arrays: ghost_grid (src/grid.f90:137-236), compute_face_area_vectors / compute_face_areas / normalize_face_normals /
 * compute_volumes / compute_cell_centre (src/geometry.f90:43-545), pole (-7) faces included; bc ids from the context's config.
 * grid_xyz = the body of the block's grid file, nodes (1:imx,1:jmx,1:kmx) of {x,y,z}, i fastest (grid.f90:78-133).  dist as in
 * set_geometry, or NULL when fest3d_gpu_find_wall_dist follows (it needs the ghosted nodes: pass nodes_out).  nodes_out (may be
 * NULL) receives nodes(-2:imx+3,-2:jmx+3,-2:kmx+3).  A non-positive volume returns F3D_ERR_GEOMETRY with the cell in
 * fest3d_gpu_error, where the reference stops with Fatal_error (geometry.f90:476-494).  Results equal the host computation bit
 * for bit (IEEE operations, no contraction). */
int fest3d_gpu_setup_geometry(Fest3dGpuCtx* ctx, const double* grid_xyz, const double* dist, double* nodes_out);
/* the metric arrays back in the reference layouts (any pointer may be NULL): what the host's writers / post-processing use */
int fest3d_gpu_get_geometry(Fest3dGpuCtx* ctx, double* cells, double* Ifaces, double* Jfaces, double* Kfaces);

/* SURVEY 8(f) rank 3 -- checkpoint without stalling the solver (replaces the synchronous get_state + text dump of
 * src/solver.f90:139,186 -> src/read_write/write/dump_solution.f90).  fest3d_gpu_checkpoint_begin snapshots qp in stream order and
 * returns at once; the device->host copy (copy stream, pinned memory) and the file write (writer thread) overlap the following
 * fest3d_gpu_step calls.  One checkpoint per context is in flight: a second begin, fest3d_gpu_checkpoint_wait and
 * fest3d_gpu_destroy wait for it.  File = 64-byte header { "F3DCKPT1", int32 imx, jmx, kmx, n_var, iter, 3 x int32 0, uint64
 * n_doubles, 16 bytes 0 } followed by qp(-2:imx+2,-2:jmx+2,-2:kmx+2,1:n_var) as float64 (ghost cells included, so that a
 * restarted run continues bit for bit).  fest3d_gpu_restart checks the header against the context (F3D_ERR_ARGUMENT on a
 * mismatch, F3D_ERR_IO on a short or foreign file), uploads the state and returns the stored iteration number. */
int fest3d_gpu_checkpoint_begin(Fest3dGpuCtx* ctx, const char* path, int iter);
int fest3d_gpu_checkpoint_wait(Fest3dGpuCtx* ctx);
int fest3d_gpu_restart(Fest3dGpuCtx* ctx, const char* path, int* iter);

/* the hot path -------------------------------------------------------------------------------------------------- */
/* n_iters iterations of { get_next_solution ; find_resnorm }.  current_iter is control%current_iter of the first
 * one (1-based; fixed-value BCs are applied while current_iter <= 2, bc_primitive.f90:236).  res_abs_out receives
 * (n_var+1) doubles per iteration: Res_abs(0:n_var) already summed over all blocks (resnorm.f90:211-225).
 * May be NULL: then no norm is copied back (the reductions still run on the device). */
int fest3d_gpu_step(Fest3dGpuCtx* ctx, int current_iter, int n_iters, double* res_abs_out);
/* several blocks that live in this process (any devices) stepped in lock step; interfaces between them are
 * exchanged device-to-device, interfaces to blocks of other processes through the communicator. */
int fest3d_gpu_step_group(Fest3dGpuCtx** ctxs, int n_ctx, int current_iter, int n_iters, double* res_abs_out);
/* The same in two halves (n_iters <= 127): _begin queues the iterations and returns at once, _end waits for them and delivers the
 * norms / the error state.  In between the host is free, e.g. to start the upload of the next state with fest3d_gpu_set_state_async
 * while these iterations run (one begin in flight per group). */
int fest3d_gpu_step_group_begin(Fest3dGpuCtx** ctxs, int n_ctx, int current_iter, int n_iters);
int fest3d_gpu_step_group_end(Fest3dGpuCtx** ctxs, int n_ctx, double* res_abs_out);
/* one residual evaluation (Temp refresh, halo exchange, ghost fill, ... , source); residue_out may be NULL */
int fest3d_gpu_residual(Fest3dGpuCtx* ctx, int current_iter, double* residue_out);
int fest3d_gpu_residual_group(Fest3dGpuCtx** ctxs, int n_ctx, int current_iter);
int fest3d_gpu_get_residue(Fest3dGpuCtx* ctx, double* residue_out);

/* debugging / parity views: which = 0 delta_t(1:imx-1,..) ; 1 mu ; 2 mu_t ; 3 sst_F1 ; 4 Temp ; 5 the CC.f90 field of
 * transition = lctm2015, DCCVn . CCnormal (all -2:imx+2,..) ;
 * 30,31,32 gradqp_x,y,z (0:imx,0:jmx,0:kmx,n_grad) */
int fest3d_gpu_get_aux(Fest3dGpuCtx* ctx, int which, double* out);
int fest3d_gpu_error(Fest3dGpuCtx* ctx, Fest3dGpuError* info);

/* multi-process (one rank per GPU) halo exchange + norm all-reduce over NCCL ------------------------------------ */
int fest3d_gpu_comm_unique_id(char id_out[128]);                    /* rank 0 calls, host broadcasts the 128 bytes */
/* Call once per context.  All contexts of a process share ONE communicator (the first call creates it, the others attach to
 * it), so a rank may own several blocks; blocks of the same rank are linked with fest3d_gpu_link_local.  An error on any rank
 * makes fest3d_gpu_step_group return on every rank (F3D_ERR_PEER where the error is not local). */
int fest3d_gpu_comm_init(Fest3dGpuCtx* ctx, int n_ranks, int rank, const char id[128],
                         const int* block_to_rank /* [n_blocks] owner rank of every block */);
/* link two contexts of the same process so that their shared interface is exchanged device-to-device */
int fest3d_gpu_link_local(Fest3dGpuCtx* a, Fest3dGpuCtx* b);

/* instrumentation ------------------------------------------------------------------------------------------------ */
/* number of kernel launches issued by this context since creation (the bench's gpu_launches claim) */
long long fest3d_gpu_launch_count(Fest3dGpuCtx* ctx);
/* device time of the dominant kernel (fused residual/update) accumulated by CUDA events on the launching stream
 * since the last reset: returns total milliseconds, *n_launches gets the launch count.  reset != 0 clears.
 * Timing is off until fest3d_gpu_kernel_timing(ctx, 1). */
int fest3d_gpu_kernel_timing(Fest3dGpuCtx* ctx, int enable);
double fest3d_gpu_kernel_time_ms(Fest3dGpuCtx* ctx, long long* n_launches, int reset);
/* the same for the Green-Gauss gradient + viscosity kernels of the staged form of the viscous path (0 launches when fused) */
double fest3d_gpu_gradient_time_ms(Fest3dGpuCtx* ctx, long long* n_launches, int reset);
/* which form of the viscous path this context runs (environment F3D_GRADIENTS at fest3d_gpu_create): 0 = staged -- gradients and
 * viscosities by their own kernel into HBM arrays that the sweep stages with TMA (default: the faster one on B200 as measured);
 * 1 = fused -- computed inside the sweep's tile pass, no gradient / viscosity array in HBM */
int fest3d_gpu_gradient_path(Fest3dGpuCtx* ctx);
const char* fest3d_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FEST3D_GPU_H */
