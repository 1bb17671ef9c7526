// Internal device context of one block.  HBM layout: every field (primitive variable, metric, work array) is one
// padded scalar array of FS doubles with the SAME strides, so cell (i,j,k) and face (i,j,k) of any direction share
// one linear index:  idx = base + i + sj*j + sk*k   (Fortran indices, i fastest -- the reference's qp(i,j,k,n) is
// already variable-major, src/vartypes.f90:21-26).  Row pitch sj is a multiple of 16 doubles and the allocation is
// shifted by 13 doubles so that interior cell i=1 of every row starts a 128-byte line.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <vector>
#include "../../include/fest3d_gpu.h"

namespace f3d {

constexpr int kG3TX = 32, kG3TY = 4;   // tile of the fused sweep (fused_kernel.cuh), also the box of its tensor maps

struct Layout {
  int imx, jmx, kmx, nv, ng;        // ng = number of gradient components (4, 5 with sa, 6 with sst, 7 with sst + lctm2015)
  int pi, pj, pk;                   // padded extents
  long long sj, sk, base, fs;       // strides, index of cell (0,0,0), doubles per field
  __host__ __device__ inline long long idx(int i, int j, int k) const { return base + i + sj * j + sk * (long long)k; }
};

// geometry field order inside Ctx::geom
enum { G_VOL = 0, G_CX, G_CY, G_CZ, G_IA, G_INX, G_INY, G_INZ, G_JA, G_JNX, G_JNY, G_JNZ, G_KA, G_KNX, G_KNY, G_KNZ, G_DIST, G_NFIELDS };

// parameters every kernel needs, passed by value (fits the 4 KB kernel-parameter space comfortably)
struct Params {
  Layout L;
  int scheme, interpolant, turbulence, time_stepping, mu_variation;
  int limiter[3], tlimiter[3], pb_switch[3];
  int bc_id[6];
  int phys[6];            // 1 when the face is a physical boundary that gets the boundary-state override (id<0 && id!=-10)
  int farlike[6];         // 1 when id is -8 or -9 (face state = ghost value)
  int ppm_flag;           // boundary re-reconstruction active (ppm / weno / weno_NM or a pole face)
  int current_iter;
  int viscous, sst, sa;   // sst: a two-equation model with n_var 7 (sst, sst2003, kkl: same field layout); sa: Spalart-Allmaras (n_var 6)
  int kkl;                // k-kL model: variable 7 is kL; its own mu_t, viscous-flux constants, source and point-implicit terms
  int lctm;               // transition = lctm2015: n_var 8 (intermittency), n_grad 7; modified F1, its own SST source, gamma diffusion flux
  int trans_bc;           // transition = bc: algebraic gamma_BC factor on the production term (source.f90:467-604, 985-1194)
  double zlo[3], zhi[3];  // make_{F,G,H}_flux_zero at the first / last face of each direction (bc.f90:53-66)
  double c1, c2, c3;
  double CFL, global_time_step;
  double gm, R_gas, mu_ref, T_ref, Sutherland_temp, Pr, tPr;
  double inv_Pr, inv_tPr, inv_gm1;   // reciprocals of Pr, tPr, gm-1
  double density_inf, x_speed_inf, y_speed_inf, z_speed_inf, pressure_inf, tk_inf, tw_inf, tv_inf, tkl_inf, tgm_inf, MInf;
  double gama1, gama2, cd_floor, mut_floor, pk_limiter;
  double gama1_default, gama2_default;   // global_sst.f90:15-16 as they stand when add_sst_source never runs (transition = bc)
  double tu_inf, nu_cr, re_theta_t;   // transition = bc: free-stream turbulence intensity (percent), chi_2 / Reynolds_number, 803.73 (Tu + 0.6067)^-1.027
  double fixed[F3D_NFIX][6];
  double res_scale[9];    // Res_scale(1:n_var) (resnorm.f90:136-150)
};

struct Link {             // what sits behind an interface face
  int kind = 0;           // 0 none, 1 local context (same process), 2 remote rank (NCCL)
  struct Ctx* peer = nullptr;
  int rank = -1;
  int neighbour_block = -1;
};

struct Ctx {
  Fest3dGpuConfig cfg;
  Params P;
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;   // halo swap hand-overs: send buffers packed / ghost layers unpacked (api.cu:exchange)
  // device memory
  double* qp = nullptr;       // nv fields (current state)
  double* qp2 = nullptr;      // nv fields (next state of the fused update; swapped with qp)
  double* ustore = nullptr;   // nv fields
  double* rstore = nullptr;   // nv fields
  double* residue = nullptr;  // nv fields
  double* temp = nullptr;     // 1 field
  double* dt = nullptr;       // 1 field
  double* geom = nullptr;     // G_NFIELDS fields
  // Two forms of the viscous path (F3D_GRADIENTS = staged | fused, api.cu):
  //   staged  Green-Gauss gradients + viscosities by their own kernel (grad.cu) into these arrays, staged by the sweep with TMA
  //           (sweep3_kernel.cuh) -- the faster form on B200 as measured (profiles/r02_g4_summary.md), default
  //   fused   computed inside the tile pass, nothing in HBM (fused_kernel.cuh); the arrays then exist only for the debugging views
  //           of fest3d_gpu_get_aux and are allocated on its first call
  int fused = 0;
  double* grad = nullptr;     // 3*ng fields: component c, direction d -> field 3*c+d
  double* mu = nullptr;       // mu [, mu_t [, F1]] as the model has them, then (staged path) a copy of the cell centre x,y,z
  int n_mu = 0;               // 1 laminar, 2 sa, 3 sst
  double* src = nullptr;      // lctm2015: the three source terms x volume, made by k_gradients<7>
  // 4-D tensor maps [field][k][j][i]: q (36 x 8 cells x nv fields); fused path: Temp (36 x 8 x 1), geometry fields volume + centre
  // (36 x 6 x 4); staged path: gradients and the aux array (36 x 6 x fields)
  CUtensorMap tm_q[2], tm_temp, tm_geo, tm_grad, tm_aux;
  double* tm_q_ptr[2] = {nullptr, nullptr};
  bool tmaps_ok = false;
  double* gbc = nullptr;      // per-face (A,nx,ny,nz) records the ghost-gradient rule reads (mis-indexed for J/K faces)
  long long gbc_off[6];
  long long* gbc_off_dev = nullptr;   // the same offsets on the device
  // implicit LU-SGS (time_accuracy = implicit, lusgs.cu): delQstar and delQ (nv fields each; ghost cells stay zero) and the three face
  // fields of spectral radius x area
  double* lusgs_dqs = nullptr;
  double* lusgs_dq = nullptr;
  double* lusgs_lam = nullptr;
  double* red = nullptr;      // reduction partials
  int red_blocks = 0;
  double* norms_dev = nullptr;   // (nv+1) per iteration slot
  double* norms_host = nullptr;  // pinned
  int* err_dev = nullptr;        // [0..3] sticky error word + first offending cell; [4] iteration slot of the norm buffer (launch_norms)
  int* err_host = nullptr;
  double* sendbuf[6] = {nullptr};
  double* recvbuf[6] = {nullptr};
  size_t buf_elems[6] = {0};
  double* staging = nullptr;     // host<->device staging for AoS geometry upload
  double* state_staging = nullptr;   // contiguous copy of qp in the reference layout (set_state / get_state, and the inbound side of the
                                     // asynchronous transfers)
  // asynchronous transfers: TWO inbound and TWO outbound staging buffers, used alternately, so that the next upload does not wait for the
  // re-layout of the previous one (nor a download for the previous download): the PCIe streams then run back to back (api.cu)
  double* state_staging_in2 = nullptr;   // inbound buffer 1 (buffer 0 is state_staging)
  double* state_staging_out = nullptr;   // outbound buffers 0, 1
  double* state_staging_out2 = nullptr;
  int in_next = 0, in_pending = 0, out_next = 0;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;       // H2D / D2H run on their own streams: full duplex beside the compute stream
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr}, ev_relaid = nullptr, ev_d2h[2] = {nullptr, nullptr};
  bool state_pending = false;        // an uploaded state waits in state_staging to be laid out into qp (done at the next step / residual)
  Link link[6];
  int steps_in_flight = 0;       // iterations queued by fest3d_gpu_step_group_begin and not yet collected (kept by the group's first context)
  bool halo_posted = false;      // the send buffers are packed and the messages of the coming stage are on their way (api.cu:exchange_post)
  struct Checkpoint* ckpt = nullptr;   // asynchronous checkpoint state (checkpoint.cu), created on first use
  void* nccl = nullptr;          // the process-wide communicator record (api.cu:CommShared)
  int n_ranks = 1, rank = 0;
  std::vector<int> block_to_rank;
  // instrumentation
  long long launches = 0;
  int timing = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  double ktime_ms = 0.0;
  long long ktime_n = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool2;   // the same for the gradient kernels of the staged path
  size_t ev_used2 = 0;
  double ktime2_ms = 0.0;
  long long ktime2_n = 0;
  // CUDA graphs of one whole iteration (every stage of every context of a step group), keyed by the group and the buffer parity;
  // kept by the first context of the group (api.cu:step_group)
  struct IterGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; bool swaps = false; };
  std::map<unsigned long long, IterGraph> graphs;
  Fest3dGpuError last_error{};
  bool geometry_set = false, state_set = false;
};

#define F3D_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      ctx->last_error.flags |= F3D_ERR_CUDA;                                                 \
      ctx->last_error.cuda_error = (int)e_;                                                  \
      fprintf(stderr, "fest3d_gpu: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return F3D_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// kernel launchers (each returns a CUDA error code through the context)
int launch_temp(Ctx* ctx);
int launch_bc(Ctx* ctx);
int launch_gradients(Ctx* ctx);   // staged form: gradients, viscosities, F1 into ctx->grad / ctx->mu (grad.cu)
int launch_lusgs(Ctx* ctx);       // implicit update of qp in place from residue and dt (lusgs.cu)
int launch_dvdy(Ctx* ctx);        // lctm2015: the CC.f90 field into aux field 3 (grad.cu)
int launch_residual(Ctx* ctx, int mode, double TF, double SF, int use_store_sum, int first_stage, int last_stage);
int launch_blend(Ctx* ctx, double a, double b);
int launch_copy_fields(Ctx* ctx, double* dst, const double* src, int nfields);
int launch_zero_fields(Ctx* ctx, double* dst, int nfields);
int launch_norms(Ctx* ctx);   // into the slot counted by the device word err_dev[4], which it advances
int launch_pack(Ctx* ctx, int face);
int launch_unpack(Ctx* ctx, int face, const double* buf);
int launch_global_dt(Ctx* ctx);
int launch_ghost_shell_copy(Ctx* ctx, double* dst, const double* src);
int launch_state_relayout(Ctx* ctx, double* fields, double* flat, int to_fields);
void checkpoint_free(Ctx* ctx);
int launch_wall_distance(Ctx* ctx, const double* nodes_host, const double* wall_host, long long n_wall, double* dist_out, double* kernel_ms);

// residual modes
enum { MODE_RESIDUE_ONLY = 0, MODE_UPDATE = 1 };

}  // namespace f3d

struct Fest3dGpuCtx : f3d::Ctx {};
