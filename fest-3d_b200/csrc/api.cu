// C ABI of the device path (include/fest3d_gpu.h): context life cycle, host <-> device layout conversion, the stage
// sequencing of get_next_solution (src/update.f90:129-226), the halo exchange that replaces apply_interface's
// MPI_SENDRECVs (src/interface1.f90:96-493) and the norm assembly of find_resnorm (src/resnorm.f90:171-225).
#include <dlfcn.h>
#include <algorithm>
#include <tuple>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "ctx.hpp"

using namespace f3d;

namespace f3d {
int upload_records(Ctx* ctx, const double* host, int n0, int n1, int n2, double* field0);
int residual_grid_ctas(const Layout& L);
int launch_setup_geometry(Ctx* ctx, const double* grid_host, double* nodes_out);
int download_records(Ctx* ctx, const double* field0, int n0, int n1, int n2, double* host);
}

// ---- NCCL through dlopen (the library must load on hosts without NCCL; multi-rank entry points fail loudly there) ----
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
#define SYM(f, s) f = (decltype(f))dlsym(h, s); if (!f) return false;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    return true;
  }
} g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
}  // namespace

static int fail(Fest3dGpuCtx* ctx, int cls) { if (ctx) ctx->last_error.flags |= cls; return cls; }
#define F3D_CUDA_RC(call) do { if ((call) != cudaSuccess) return F3D_ERR_CUDA; } while (0)
static int apply_pending_state(Fest3dGpuCtx* ctx);

extern "C" const char* fest3d_gpu_version(void) { return "fest3d-b200 0.1 (sm_100a)"; }

static int create_impl(Fest3dGpuCtx* ctx, const Fest3dGpuConfig* cfg, int device, bool sst, bool sa);

extern "C" int fest3d_gpu_create(Fest3dGpuCtx** out, const Fest3dGpuConfig* cfg, int device) {
  if (!out || !cfg) return F3D_ERR_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "fest3d_gpu: no CUDA device available -- this library has no CPU fallback\n");
    return F3D_ERR_CUDA;
  }
  // scope check: only what this path implements; everything else is an explicit error, never a silent fallback
  const bool kkl = cfg->turbulence == F3D_TURB_KKL;
  const bool sst = cfg->turbulence == F3D_TURB_SST || cfg->turbulence == F3D_TURB_SST2003 || kkl;   // the two-equation layout (n_var 7, n_grad 6)
  const bool sa = cfg->turbulence == F3D_TURB_SA;   // 'saBC' has no case in the reference's source dispatcher (source.f90:119-153)
  if (cfg->turbulence != F3D_TURB_NONE && !sst && !sa) return F3D_ERR_UNSUPPORTED;
  // transition = bc with sa / sst / sst2003; lctm2015 with sst / sst2003 (the only cases of the reference's source dispatcher, source.f90:119-153)
  const bool lctm = cfg->transition == F3D_TRANS_LCTM2015;
  if (cfg->transition == F3D_TRANS_BC && !((sst || sa) && !kkl)) return F3D_ERR_UNSUPPORTED;
  if (lctm && !(sst && !kkl)) return F3D_ERR_UNSUPPORTED;
  if (cfg->transition != F3D_TRANS_NONE && cfg->transition != F3D_TRANS_BC && !lctm) return F3D_ERR_ARGUMENT;
  if (kkl || lctm) { const char* e = getenv("F3D_GRADIENTS"); if (e && strcmp(e, "fused") == 0) return F3D_ERR_UNSUPPORTED; }   // staged form only
  if (cfg->time_accuracy > F3D_T_IMPLICIT || cfg->time_accuracy < 0) return F3D_ERR_UNSUPPORTED;   // plusgs
  if (cfg->time_accuracy == F3D_T_IMPLICIT) {
    // LU-SGS: every routine of lusgs.f90's dispatcher (:186 laminar / inviscid, :686 SST, :1198 k-kL, :1680 SA, :2262 lctm2015).  The sweeps
    // read mu / mu_t / F1 (and, for sa and lctm2015, the velocity gradients) as arrays, which only the staged form of the viscous path keeps.
    const char* e = getenv("F3D_GRADIENTS");
    if (cfg->mu_ref != 0.0 && e && strcmp(e, "fused") == 0) return F3D_ERR_UNSUPPORTED;
  }
  if (cfg->scheme < 0 || cfg->scheme > F3D_SLAU || cfg->interpolant < 0 || cfg->interpolant > F3D_WENO_NM) return F3D_ERR_ARGUMENT;
  if (cfg->n_var != (sst ? 7 : (sa ? 6 : 5)) + (lctm ? 1 : 0)) return F3D_ERR_ARGUMENT;   // state.f90:291-320
  if ((sst || sa) && cfg->mu_ref == 0.0) return F3D_ERR_UNSUPPORTED;
  if (cfg->imx < 2 || cfg->jmx < 2 || cfg->kmx < 2) return F3D_ERR_ARGUMENT;

  Fest3dGpuCtx* ctx = new Fest3dGpuCtx();
  ctx->cfg = *cfg;
  ctx->device = device;
  const int rc = create_impl(ctx, cfg, device, sst, sa);
  if (rc) { fest3d_gpu_destroy(ctx); return rc; }   // every allocation made so far is released
  *out = ctx;
  return 0;
}

static int create_impl(Fest3dGpuCtx* ctx, const Fest3dGpuConfig* cfg, int device, bool sst, bool sa) {
  F3D_CUDA(cudaSetDevice(device));
  Params& P = ctx->P;
  memset(&P, 0, sizeof(P));
  Layout& L = P.L;
  L.imx = cfg->imx; L.jmx = cfg->jmx; L.kmx = cfg->kmx; L.nv = cfg->n_var; L.ng = sst ? 6 : (sa ? 5 : 4);
  if (cfg->transition == F3D_TRANS_LCTM2015) L.ng += 1;   // gradients.f90:259-264
  // a row holds i = -2 .. imx+2 behind its 15-element lead-in (interior cell i = 1 starts a 128-byte line), so that the tensor maps
  // of the sweep can address it as one dimension
  L.pi = ((cfg->imx + 18 + 15) / 16) * 16; L.pj = cfg->jmx + 6; L.pk = cfg->kmx + 6;
  L.sj = L.pi; L.sk = (long long)L.pi * L.pj;
  L.base = 13 + 2 + 2 * L.sj + 2 * L.sk;
  L.fs = ((13 + L.sk * L.pk + 31) / 32) * 32;
  P.scheme = cfg->scheme; P.interpolant = cfg->interpolant; P.turbulence = cfg->turbulence;
  P.time_stepping = cfg->time_stepping; P.mu_variation = cfg->mu_variation;
  const bool pb_interp = cfg->interpolant == F3D_MUSCL || cfg->interpolant == F3D_PPM;   // the only callers (muscl.f90:231, ppm.f90:220)
  for (int d = 0; d < 3; ++d) { P.limiter[d] = cfg->limiter[d]; P.tlimiter[d] = cfg->tlimiter[d]; P.pb_switch[d] = (pb_interp && cfg->pb_switch[d] == 1) ? 1 : 0; }
  P.ppm_flag = (cfg->interpolant == F3D_PPM || cfg->interpolant == F3D_WENO || cfg->interpolant == F3D_WENO_NM) ? 1 : 0;
  for (int f = 0; f < 6; ++f) {
    int id = cfg->bc_id[f];
    if (cfg->pbc_id[f] >= 0) id = -10;                      // bc.f90:26-31
    P.bc_id[f] = id;
    P.phys[f] = (id < 0 && id != -10) ? 1 : 0;
    P.farlike[f] = (id == -8 || id == -9) ? 1 : 0;
    if (id == -7) P.ppm_flag = 1;                           // boundary_state_reconstruction.f90:46-47
  }
  auto wallish = [](int id) { return id == -5 || id == -6 || id == -7; };
  for (int d = 0; d < 3; ++d) { P.zlo[d] = wallish(P.bc_id[2 * d]) ? 0.0 : 1.0; P.zhi[d] = wallish(P.bc_id[2 * d + 1]) ? 0.0 : 1.0; }
  P.c2 = 1 + cfg->accur; P.c3 = 0.5 * cfg->accur; P.c1 = P.c2 - P.c3;
  P.current_iter = 1;
  P.viscous = cfg->mu_ref != 0.0; P.sst = sst ? 1 : 0; P.sa = sa ? 1 : 0; P.kkl = cfg->turbulence == F3D_TURB_KKL ? 1 : 0;
  P.tkl_inf = cfg->tkl_inf;
  P.lctm = cfg->transition == F3D_TRANS_LCTM2015 ? 1 : 0; P.tgm_inf = cfg->tgm_inf;
  P.trans_bc = cfg->transition == F3D_TRANS_BC ? 1 : 0; P.tu_inf = cfg->tu_inf;
  P.re_theta_t = (803.73 * (pow(cfg->tu_inf + 0.6067, -1.027)));   // source.f90:579, 1164
  P.nu_cr = cfg->mu_ref != 0.0 ? 5.0 / (cfg->density_inf * cfg->vel_mag * 1.0 / cfg->mu_ref) : 0.0;   // chi_2 / Reynolds_number (state.f90:89)
  P.CFL = cfg->CFL; P.global_time_step = cfg->global_time_step;
  P.gm = cfg->gm; P.R_gas = cfg->R_gas; P.mu_ref = cfg->mu_ref; P.T_ref = cfg->T_ref; P.Sutherland_temp = cfg->Sutherland_temp;
  P.Pr = cfg->Pr; P.tPr = cfg->tPr;
  P.inv_Pr = 1.0 / cfg->Pr; P.inv_tPr = 1.0 / cfg->tPr; P.inv_gm1 = 1.0 / (cfg->gm - 1.0);
  P.density_inf = cfg->density_inf; P.x_speed_inf = cfg->x_speed_inf; P.y_speed_inf = cfg->y_speed_inf; P.z_speed_inf = cfg->z_speed_inf;
  P.pressure_inf = cfg->pressure_inf; P.tk_inf = cfg->tk_inf; P.tw_inf = cfg->tw_inf; P.tv_inf = cfg->tv_inf; P.MInf = cfg->MInf;
  const double kappa = 0.41;
  if (cfg->turbulence == F3D_TURB_SST2003) {   // source.f90:205-211, viscosity.f90:243,251
    P.gama1 = 5.0 / 9.0; P.gama2 = 0.44; P.cd_floor = 1.0e-10; P.mut_floor = 1.0e-10; P.pk_limiter = 10;
  } else {                                      // global_sst.f90:15-16
    P.gama1 = (0.075 / 0.09) - ((0.5 * (kappa * kappa)) / sqrt(0.09));
    P.gama2 = (0.0828 / 0.09) - ((0.856 * (kappa * kappa)) / sqrt(0.09));
    P.cd_floor = 1.0e-20; P.mut_floor = 1.e-20; P.pk_limiter = 20;
  }
  P.gama1_default = (0.075 / 0.09) - ((0.5 * (kappa * kappa)) / sqrt(0.09));
  P.gama2_default = (0.0828 / 0.09) - ((0.856 * (kappa * kappa)) / sqrt(0.09));
  memcpy(P.fixed, cfg->fixed, sizeof(P.fixed));
  // Res_scale (resnorm.f90:136-150); slot 0 is the mass imbalance scale (1)
  double sc[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
  sc[1] = cfg->density_inf * cfg->vel_mag;
  sc[2] = sc[3] = sc[4] = cfg->density_inf * cfg->vel_mag * cfg->vel_mag;
  sc[5] = (0.5 * cfg->density_inf * (cfg->vel_mag * cfg->vel_mag * cfg->vel_mag) + ((cfg->gm / (cfg->gm - 1.)) * cfg->pressure_inf));
  if (sst) { sc[6] = cfg->density_inf * cfg->vel_mag * cfg->tk_inf; sc[7] = cfg->density_inf * cfg->vel_mag * (P.kkl ? cfg->tkl_inf : cfg->tw_inf); }   // resnorm.f90:148-153
  if (sa) sc[6] = cfg->density_inf * cfg->vel_mag * cfg->tv_inf;   // resnorm.f90:157-158
  // lctm2015: setup_scale never assigns Res_scale(8) (resnorm.f90:136-167), the reference divides by whatever its allocation holds: no
  // reference value exists for that norm; it is reported here with scale 1 (sc[8] above)

  F3D_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming));
  F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
  const size_t fb = (size_t)L.fs * sizeof(double);
  const int nv = L.nv;
  auto dalloc = [&](double** p, size_t nfields) -> cudaError_t {
    cudaError_t e = cudaMalloc((void**)p, nfields * fb);
    if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, nfields * fb, ctx->stream);
    return e;
  };
  F3D_CUDA(dalloc(&ctx->qp, nv)); F3D_CUDA(dalloc(&ctx->qp2, nv)); F3D_CUDA(dalloc(&ctx->residue, nv));
  F3D_CUDA(dalloc(&ctx->temp, 1)); F3D_CUDA(dalloc(&ctx->dt, 1)); F3D_CUDA(dalloc(&ctx->geom, G_NFIELDS));
  if (cfg->time_accuracy != F3D_T_NONE && cfg->time_accuracy != F3D_T_IMPLICIT) F3D_CUDA(dalloc(&ctx->ustore, nv));
  if (cfg->time_accuracy == F3D_T_RK2 || cfg->time_accuracy == F3D_T_RK4) F3D_CUDA(dalloc(&ctx->rstore, nv));
  if (cfg->time_accuracy == F3D_T_IMPLICIT) { F3D_CUDA(dalloc(&ctx->lusgs_dqs, nv)); F3D_CUDA(dalloc(&ctx->lusgs_dq, nv)); F3D_CUDA(dalloc(&ctx->lusgs_lam, 3)); }
  ctx->n_mu = sst ? 3 : (sa ? 2 : 1);
  {   // form of the viscous path: "staged" (default: measured faster on B200) or "fused" (no gradient / viscosity array in HBM)
    const char* e = getenv("F3D_GRADIENTS");
    ctx->fused = (e && strcmp(e, "fused") == 0) ? 1 : 0;
  }
  // (sa: the 15 gradient fields are followed by one more, the cross-diffusion scalar grad(rho) . grad(nu-tilde) of its source term -- the
  // slot the tensor-map box of the sweep would otherwise pad: grad.cu:k_gradients)
  if (P.viscous && !ctx->fused) { F3D_CUDA(dalloc(&ctx->grad, sa ? 16 : 3 * L.ng)); F3D_CUDA(dalloc(&ctx->mu, ctx->n_mu + 3)); }
  if (P.lctm) F3D_CUDA(dalloc(&ctx->src, 3));
  // staging for the AoS records: the largest face array
  const size_t rec_max = (size_t)4 * (L.imx + 6) * (L.jmx + 6) * (L.kmx + 6) * sizeof(double);
  F3D_CUDA(cudaMalloc((void**)&ctx->staging, rec_max));
  ctx->red_blocks = residual_grid_ctas(L);
  F3D_CUDA(cudaMalloc((void**)&ctx->red, sizeof(double) * (size_t)std::max(ctx->red_blocks * (nv + 1), 256)));
  F3D_CUDA(cudaMalloc((void**)&ctx->norms_dev, sizeof(double) * (1024 + 64)));
  F3D_CUDA(cudaMemcpyAsync(ctx->norms_dev + 1024, sc, sizeof(double) * 9, cudaMemcpyHostToDevice, ctx->stream));
  F3D_CUDA(cudaMallocHost((void**)&ctx->norms_host, sizeof(double) * 1024));
  F3D_CUDA(cudaMalloc((void**)&ctx->err_dev, sizeof(int) * 8));
  F3D_CUDA(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int) * 8, ctx->stream));
  F3D_CUDA(cudaMallocHost((void**)&ctx->err_host, sizeof(int) * 4));
  {   // 4-D tensor maps [field][k][j][i] (pitches fs, sk, sj) of the arrays the sweep stages; box = one tile plane of all fields
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return fail(ctx, F3D_ERR_CUDA);
    auto encode = [&](CUtensorMap* tm, double* base, int nfields, int rows, int box_fields) -> bool {   // box: 36 x rows x 1 x box_fields
      const cuuint64_t dims[4] = {(cuuint64_t)L.sj, (cuuint64_t)L.pj, (cuuint64_t)L.pk, (cuuint64_t)nfields};
      const cuuint64_t strides[3] = {(cuuint64_t)L.sj * 8, (cuuint64_t)L.sk * 8, (cuuint64_t)L.fs * 8};
      const cuuint32_t box[4] = {(cuuint32_t)(kG3TX + 4), (cuuint32_t)rows, 1u, (cuuint32_t)box_fields};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      return ((EncodeFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = encode(&ctx->tm_q[0], ctx->qp, nv, kG3TY + 4, nv) && encode(&ctx->tm_q[1], ctx->qp2, nv, kG3TY + 4, nv);
    ctx->tm_q_ptr[0] = ctx->qp; ctx->tm_q_ptr[1] = ctx->qp2;
    ok = ok && encode(&ctx->tm_temp, ctx->temp, 1, kG3TY + 4, 1);
    ok = ok && encode(&ctx->tm_geo, ctx->geom, G_NFIELDS, kG3TY + 2, 4);   // volume + centre x, y, z = geometry fields 0..3
    if (P.viscous && !ctx->fused) {
      // lctm2015 (eight variables): the sweep stages the first 18 gradient fields and mu, mu_t, F1 + the CC.f90 field; the intermittency
      // gradient and the cell centres stay in global memory (sweep_common.cuh:RecF)
      const int ngf = P.sa ? 16 : 3 * L.ng, naux = P.lctm ? ctx->n_mu + 1 : ctx->n_mu + 3;
      ok = ok && encode(&ctx->tm_grad, ctx->grad, ngf, kG3TY + 2, P.lctm ? 18 : ((ngf + 1) & ~1)) && encode(&ctx->tm_aux, ctx->mu, naux, kG3TY + 2, (naux + 1) & ~1);
    }
    if (!ok) { fprintf(stderr, "fest3d_gpu: cuTensorMapEncodeTiled failed\n"); return fail(ctx, F3D_ERR_CUDA); }
    ctx->tmaps_ok = true;
  }
  // ghost-gradient face records
  if (P.viscous) {
    size_t tot = 0;
    const int mx[3] = {L.imx, L.jmx, L.kmx};
    for (int f = 0; f < 6; ++f) {
      const int ax = f / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
      ctx->gbc_off[f] = (long long)tot;
      tot += (size_t)4 * (mx[a_ax] - 1) * (mx[b_ax] - 1);
    }
    F3D_CUDA(cudaMalloc((void**)&ctx->gbc, tot * sizeof(double)));
    F3D_CUDA(cudaMalloc((void**)&ctx->gbc_off_dev, 6 * sizeof(long long)));
    F3D_CUDA(cudaMemcpy(ctx->gbc_off_dev, ctx->gbc_off, 6 * sizeof(long long), cudaMemcpyHostToDevice));
  }
  // halo buffers for interface faces
  {
    const int mx[3] = {L.imx, L.jmx, L.kmx};
    for (int f = 0; f < 6; ++f) {
      if (!(cfg->bc_id[f] >= 0 || cfg->pbc_id[f] >= 0)) continue;
      const int ax = f / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
      ctx->buf_elems[f] = (size_t)(mx[a_ax] - 1) * (mx[b_ax] - 1) * 3 * nv;
      F3D_CUDA(cudaMalloc((void**)&ctx->sendbuf[f], ctx->buf_elems[f] * sizeof(double)));
      F3D_CUDA(cudaMalloc((void**)&ctx->recvbuf[f], ctx->buf_elems[f] * sizeof(double)));
      ctx->link[f].neighbour_block = cfg->bc_id[f] >= 0 ? cfg->bc_id[f] : cfg->pbc_id[f];
    }
  }
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->last_error.block_id = cfg->block_id;
  return 0;
}

static void comm_release(Fest3dGpuCtx* ctx);

extern "C" int fest3d_gpu_destroy(Fest3dGpuCtx* ctx) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  cudaSetDevice(ctx->device);
  checkpoint_free(ctx);   // joins a writer thread that may still be copying
  cudaDeviceSynchronize();
  double* bufs[] = {ctx->qp, ctx->qp2, ctx->ustore, ctx->rstore, ctx->residue, ctx->temp, ctx->dt, ctx->geom, ctx->grad, ctx->mu, ctx->gbc,
                    ctx->red, ctx->norms_dev, ctx->staging, ctx->state_staging, ctx->lusgs_dqs, ctx->lusgs_dq, ctx->lusgs_lam, ctx->src};
  for (double* b : bufs) if (b) cudaFree(b);
  for (int f = 0; f < 6; ++f) { if (ctx->sendbuf[f]) cudaFree(ctx->sendbuf[f]); if (ctx->recvbuf[f]) cudaFree(ctx->recvbuf[f]); }
  if (ctx->err_dev) cudaFree(ctx->err_dev);
  for (auto& g : ctx->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  if (ctx->gbc_off_dev) cudaFree(ctx->gbc_off_dev);
  if (ctx->norms_host) cudaFreeHost(ctx->norms_host);
  if (ctx->err_host) cudaFreeHost(ctx->err_host);
  for (auto& e : ctx->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  for (auto& e : ctx->ev_pool2) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  comm_release(ctx);
  for (double* b : {ctx->state_staging_out, ctx->state_staging_out2, ctx->state_staging_in2}) if (b) cudaFree(b);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  for (cudaEvent_t e : {ctx->ev_h2d[0], ctx->ev_h2d[1], ctx->ev_in_free[0], ctx->ev_in_free[1], ctx->ev_relaid, ctx->ev_d2h[0], ctx->ev_d2h[1]}) if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->ev_pack) cudaEventDestroy(ctx->ev_pack);
  if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
  delete ctx;
  return 0;
}

extern "C" int fest3d_gpu_set_stream(Fest3dGpuCtx* ctx, void* s) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return 0;
}

extern "C" int fest3d_gpu_sync(Fest3dGpuCtx* ctx) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  F3D_CUDA(cudaSetDevice(ctx->device));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// host (reference layout) <-> device (padded SoA) copies of cell-shaped arrays: one strided 3-D copy per variable
static int copy_cells(Fest3dGpuCtx* ctx, double* dev_field, double* host, int nfields, bool to_device, int lo, int e0, int e1, int e2) {
  const Layout& L = ctx->P.L;
  for (int v = 0; v < nfields; ++v) {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    double* d = dev_field + (long long)v * L.fs + L.idx(lo, lo, lo);
    double* h = host + (size_t)v * e0 * e1 * e2;
    cudaPitchedPtr dp = make_cudaPitchedPtr(d, (size_t)L.sj * sizeof(double), L.sj, L.pj);
    cudaPitchedPtr hp = make_cudaPitchedPtr(h, (size_t)e0 * sizeof(double), e0, e1);
    p.srcPtr = to_device ? hp : dp;
    p.dstPtr = to_device ? dp : hp;
    p.extent = make_cudaExtent((size_t)e0 * sizeof(double), e1, e2);
    p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    F3D_CUDA(cudaMemcpy3DAsync(&p, ctx->stream));
  }
  return 0;
}

// staged path: the cell-centre fields copied behind the viscosity fields: one "aux" array for the tensor-map staging of the sweep
// (lctm2015: the static CC.f90 field instead, which depends on the wall distance)
static int init_aux_fields(Fest3dGpuCtx* ctx) {
  if (!ctx->P.viscous || ctx->fused) return 0;
  const long long fs = ctx->P.L.fs;
  if (ctx->P.lctm) {
    const int rc = launch_dvdy(ctx);
    if (rc) return rc;
    F3D_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  F3D_CUDA(cudaMemcpyAsync(ctx->mu + (long long)ctx->n_mu * fs, ctx->geom + (long long)G_CX * fs, 3 * fs * sizeof(double), cudaMemcpyDeviceToDevice,
                           ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fest3d_gpu_set_geometry(Fest3dGpuCtx* ctx, const double* cells, const double* Ifaces, const double* Jfaces,
                                       const double* Kfaces, const double* dist) {
  if (!ctx || !cells || !Ifaces || !Jfaces || !Kfaces) return fail(ctx, F3D_ERR_ARGUMENT);
  if ((ctx->P.sst || ctx->P.sa) && !dist) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  const long long fs = L.fs;
  int rc;
  if ((rc = upload_records(ctx, cells, L.imx + 5, L.jmx + 5, L.kmx + 5, ctx->geom + (long long)G_VOL * fs))) return rc;
  if ((rc = upload_records(ctx, Ifaces, L.imx + 6, L.jmx + 5, L.kmx + 5, ctx->geom + (long long)G_IA * fs))) return rc;
  if ((rc = upload_records(ctx, Jfaces, L.imx + 5, L.jmx + 6, L.kmx + 5, ctx->geom + (long long)G_JA * fs))) return rc;
  if ((rc = upload_records(ctx, Kfaces, L.imx + 5, L.jmx + 5, L.kmx + 6, ctx->geom + (long long)G_KA * fs))) return rc;
  if (dist) {
    if ((rc = copy_cells(ctx, ctx->geom + (long long)G_DIST * fs, const_cast<double*>(dist), 1, true, -2, L.imx + 5, L.jmx + 5, L.kmx + 5))) return rc;
  }
  if (ctx->P.viscous) {
    // face records for the ghost-gradient rule.  The reference passes Jfaces / Kfaces to a dummy declared with the Ifaces
    // shape (gradients.f90:549-612): element (i,j,k) is then read at record offset (i+2) + (imx+6)*((j+2) + (jmx+5)*(k+2)) of
    // the actual array.  Reproduced here on the host, once.
    if ((rc = init_aux_fields(ctx))) return rc;
    const int mx[3] = {L.imx, L.jmx, L.kmx};
    const double* arrs[3] = {Ifaces, Jfaces, Kfaces};
    std::vector<double> rec;
    for (int f = 0; f < 6; ++f) {
      const int ax = f / 2, a_ax = (ax == 0) ? 1 : 0, b_ax = (ax == 2) ? 1 : 2;
      const int na = mx[a_ax] - 1, nb = mx[b_ax] - 1;
      rec.resize((size_t)4 * na * nb);
      for (int b = 0; b < nb; ++b)
        for (int a = 0; a < na; ++a) {
          int idx[3]; idx[a_ax] = a + 1; idx[b_ax] = b + 1; idx[ax] = (f % 2 == 0) ? 1 : mx[ax];
          const size_t off = (size_t)(idx[0] + 2) + (size_t)(L.imx + 6) * ((size_t)(idx[1] + 2) + (size_t)(L.jmx + 5) * (size_t)(idx[2] + 2));
          memcpy(&rec[4 * ((size_t)b * na + a)], arrs[ax] + 4 * off, 4 * sizeof(double));
        }
      F3D_CUDA(cudaMemcpy(ctx->gbc + ctx->gbc_off[f], rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->geometry_set = true;
  return 0;
}

// find_wall_dist (src/wall/wall_dist.f90:84-131) on the device; replaces the `dist` argument of fest3d_gpu_set_geometry
extern "C" int fest3d_gpu_find_wall_dist(Fest3dGpuCtx* ctx, const double* nodes, const double* wall_xyz, long long n_wall, double* dist_out,
                                         double* kernel_ms) {
  if (!ctx || !nodes || n_wall < 0 || (n_wall > 0 && !wall_xyz)) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  return launch_wall_distance(ctx, nodes, wall_xyz, n_wall, dist_out, kernel_ms);
}

// Full-state transfers: one contiguous DMA between the host array (reference layout) and a device staging buffer, plus a
// re-layout kernel (pitched per-row DMA of the padded fields reaches only about half of the PCIe rate).
static int ensure_state_staging(Fest3dGpuCtx* ctx) {
  if (ctx->state_staging) return 0;
  const Layout& L = ctx->P.L;
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  F3D_CUDA(cudaMalloc((void**)&ctx->state_staging, n * sizeof(double)));
  return 0;
}

extern "C" int fest3d_gpu_set_state(Fest3dGpuCtx* ctx, const double* qp) {
  if (!ctx || !qp) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  int rc = ensure_state_staging(ctx);
  if (rc) return rc;
  if ((rc = apply_pending_state(ctx))) return rc;   // an asynchronous upload still waiting: lay it out first, this state then replaces it
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  F3D_CUDA(cudaMemcpyAsync(ctx->state_staging, qp, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = launch_state_relayout(ctx, ctx->qp, ctx->state_staging, 1))) return rc;
  // a fresh state (initial field, restart from a good checkpoint) clears the sticky device error word of an earlier failure
  F3D_CUDA(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int) * 4, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  { const int bid = ctx->last_error.block_id; ctx->last_error = Fest3dGpuError{}; ctx->last_error.block_id = bid; }
  ctx->state_set = true;
  return 0;
}

extern "C" int fest3d_gpu_get_state(Fest3dGpuCtx* ctx, double* qp) {
  if (!ctx || !qp) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  int rc = ensure_state_staging(ctx);
  if (rc) return rc;
  if ((rc = apply_pending_state(ctx))) return rc;
  if (ctx->ev_h2d[0]) F3D_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[0], 0));   // (buffer 0 is the one the synchronous calls share)
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  if ((rc = launch_state_relayout(ctx, ctx->qp, ctx->state_staging, 0))) return rc;
  F3D_CUDA(cudaMemcpyAsync(qp, ctx->state_staging, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Asynchronous, full-duplex form of the two transfers (a host that wants qp back every few iterations, the e2e leg of bench.py):
//   fest3d_gpu_set_state_async  H2D on the inbound copy stream into the staging buffer and return; the re-layout into the padded
//                               fields is queued by the next fest3d_gpu_step / fest3d_gpu_residual call (stream-ordered behind the copy)
//   fest3d_gpu_get_state_async  re-layout on the compute stream (stream order: after every step issued so far), D2H on the outbound copy
//                               stream; the host buffer is valid after fest3d_gpu_state_wait
// so the upload of the next state overlaps the running iterations AND the download of the previous result (PCIe is full duplex).
static int ensure_async_state(Fest3dGpuCtx* ctx) {
  int rc = ensure_state_staging(ctx);
  if (rc) return rc;
  if (ctx->copy_in) return 0;
  const Layout& L = ctx->P.L;
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  F3D_CUDA(cudaMalloc((void**)&ctx->state_staging_out, n * sizeof(double)));
  F3D_CUDA(cudaMalloc((void**)&ctx->state_staging_out2, n * sizeof(double)));
  F3D_CUDA(cudaMalloc((void**)&ctx->state_staging_in2, n * sizeof(double)));
  F3D_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
  F3D_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[b], cudaEventDisableTiming));
    F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_in_free[b], cudaEventDisableTiming));
    F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[b], cudaEventDisableTiming));
  }
  F3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_relaid, cudaEventDisableTiming));
  return 0;
}

// queue the re-layout of an uploaded state (called at the head of every step / residual call)
static int apply_pending_state(Fest3dGpuCtx* ctx) {
  if (!ctx->state_pending) return 0;
  const int b = ctx->in_pending;
  F3D_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
  int rc = launch_state_relayout(ctx, ctx->qp, b ? ctx->state_staging_in2 : ctx->state_staging, 1);
  if (rc) return rc;
  F3D_CUDA(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int) * 4, ctx->stream));
  F3D_CUDA(cudaEventRecord(ctx->ev_in_free[b], ctx->stream));   // this staging buffer may take the upload after the next
  ctx->state_pending = false;
  ctx->state_set = true;
  return 0;
}

extern "C" int fest3d_gpu_set_state_async(Fest3dGpuCtx* ctx, const double* qp) {
  if (!ctx || !qp) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_async_state(ctx);
  if (rc) return rc;
  if (ctx->state_pending && (rc = apply_pending_state(ctx))) return rc;   // a second upload without a step in between: keep the order
  const Layout& L = ctx->P.L;
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  const int b = ctx->in_next;
  F3D_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_in_free[b], 0));   // the re-layout that last read this buffer (two uploads ago) is done
  F3D_CUDA(cudaMemcpyAsync(b ? ctx->state_staging_in2 : ctx->state_staging, qp, n * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_in));
  F3D_CUDA(cudaEventRecord(ctx->ev_h2d[b], ctx->copy_in));
  ctx->in_pending = b; ctx->in_next = b ^ 1;
  ctx->state_pending = true;
  { const int bid = ctx->last_error.block_id; ctx->last_error = Fest3dGpuError{}; ctx->last_error.block_id = bid; }
  return 0;
}

extern "C" int fest3d_gpu_get_state_async(Fest3dGpuCtx* ctx, double* qp) {
  if (!ctx || !qp) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_async_state(ctx);
  if (rc) return rc;
  // (an upload that is still pending -- fest3d_gpu_set_state_async of the NEXT state -- is not applied here: it takes effect at the
  // next step call; what is downloaded is the state behind the steps issued so far)
  const Layout& L = ctx->P.L;
  const size_t n = (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5);
  const int b = ctx->out_next;
  double* const out = b ? ctx->state_staging_out2 : ctx->state_staging_out;
  F3D_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));   // the download before the previous one has left this outbound buffer
  if ((rc = launch_state_relayout(ctx, ctx->qp, out, 0))) return rc;
  F3D_CUDA(cudaEventRecord(ctx->ev_relaid, ctx->stream));
  F3D_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_relaid, 0));
  F3D_CUDA(cudaMemcpyAsync(qp, out, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_out));
  F3D_CUDA(cudaEventRecord(ctx->ev_d2h[b], ctx->copy_out));
  ctx->out_next = b ^ 1;
  return 0;
}

extern "C" int fest3d_gpu_state_wait(Fest3dGpuCtx* ctx) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  F3D_CUDA(cudaSetDevice(ctx->device));
  if (ctx->copy_in) { F3D_CUDA(cudaStreamSynchronize(ctx->copy_in)); F3D_CUDA(cudaStreamSynchronize(ctx->copy_out)); }
  return 0;
}

extern "C" int fest3d_gpu_get_residue(Fest3dGpuCtx* ctx, double* out) {
  if (!ctx || !out) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  int rc = copy_cells(ctx, ctx->residue, out, L.nv, false, 1, L.imx - 1, L.jmx - 1, L.kmx - 1);
  if (rc) return rc;
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fest3d_gpu_get_aux(Fest3dGpuCtx* ctx, int which, double* out) {
  if (!ctx || !out) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  int rc = F3D_ERR_ARGUMENT;
  const bool view = (which >= 1 && which <= 3) || (which >= 30 && which <= 32);
  if (view && ctx->P.viscous && ctx->fused) {
    // fused path: mu / mu_t / F1 / gradients never reach HBM (the sweep keeps them in shared memory): for these views the stand-alone
    // kernels of grad.cu recompute them from the current qp / Temp (ghost cells as the last stage left them)
    const size_t fb = (size_t)L.fs * sizeof(double);
    if (!ctx->grad) {
      F3D_CUDA(cudaMalloc((void**)&ctx->grad, (size_t)(ctx->P.sa ? 16 : 3 * L.ng) * fb));
      F3D_CUDA(cudaMalloc((void**)&ctx->mu, (size_t)ctx->n_mu * fb));
    }
    F3D_CUDA(cudaMemsetAsync(ctx->grad, 0, (size_t)(ctx->P.sa ? 16 : 3 * L.ng) * fb, ctx->stream));
    F3D_CUDA(cudaMemsetAsync(ctx->mu, 0, (size_t)ctx->n_mu * fb, ctx->stream));
    if ((rc = launch_gradients(ctx))) return fail(ctx, rc);
    rc = F3D_ERR_ARGUMENT;
  }
  if (which == 0) rc = copy_cells(ctx, ctx->dt, out, 1, false, 1, L.imx - 1, L.jmx - 1, L.kmx - 1);
  else if (which >= 1 && which <= 3 && ctx->mu && which <= ctx->n_mu) rc = copy_cells(ctx, ctx->mu + (long long)(which - 1) * L.fs, out, 1, false, -2, L.imx + 5, L.jmx + 5, L.kmx + 5);
  else if (which == 4) rc = copy_cells(ctx, ctx->temp, out, 1, false, -2, L.imx + 5, L.jmx + 5, L.kmx + 5);
  else if (which == 5 && ctx->P.lctm && ctx->mu) rc = copy_cells(ctx, ctx->mu + 3 * L.fs, out, 1, false, -2, L.imx + 5, L.jmx + 5, L.kmx + 5);
  else if (which >= 30 && which <= 32 && ctx->grad) {
    // gradqp_d(0:imx,0:jmx,0:kmx,1:n_grad): component c of direction d lives in field 3*c+d
    const int d = which - 30;
    rc = 0;
    for (int c = 0; c < L.ng && !rc; ++c)
      rc = copy_cells(ctx, ctx->grad + (long long)(3 * c + d) * L.fs, out + (size_t)c * (L.imx + 1) * (L.jmx + 1) * (L.kmx + 1), 1, false, 0, L.imx + 1, L.jmx + 1, L.kmx + 1);
  }
  if (rc) return fail(ctx, rc);
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fest3d_gpu_error(Fest3dGpuCtx* ctx, Fest3dGpuError* info) {
  if (!ctx || !info) return F3D_ERR_ARGUMENT;
  *info = ctx->last_error;
  return 0;
}

extern "C" long long fest3d_gpu_launch_count(Fest3dGpuCtx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int fest3d_gpu_kernel_timing(Fest3dGpuCtx* ctx, int enable) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  ctx->timing = enable;
  return 0;
}

extern "C" double fest3d_gpu_kernel_time_ms(Fest3dGpuCtx* ctx, long long* n, int reset) {
  if (!ctx) return -1.0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (size_t e = 0; e < ctx->ev_used; ++e) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool[e].first, ctx->ev_pool[e].second) == cudaSuccess) { ctx->ktime_ms += ms; ctx->ktime_n++; }
  }
  ctx->ev_used = 0;
  const double t = ctx->ktime_ms;
  if (n) *n = ctx->ktime_n;
  if (reset) { ctx->ktime_ms = 0.0; ctx->ktime_n = 0; }
  return t;
}

extern "C" double fest3d_gpu_gradient_time_ms(Fest3dGpuCtx* ctx, long long* n, int reset) {
  if (!ctx) return -1.0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (size_t e = 0; e < ctx->ev_used2; ++e) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool2[e].first, ctx->ev_pool2[e].second) == cudaSuccess) { ctx->ktime2_ms += ms; ctx->ktime2_n++; }
  }
  ctx->ev_used2 = 0;
  const double t = ctx->ktime2_ms;
  if (n) *n = ctx->ktime2_n;
  if (reset) { ctx->ktime2_ms = 0.0; ctx->ktime2_n = 0; }
  return t;
}

extern "C" int fest3d_gpu_gradient_path(Fest3dGpuCtx* ctx) { return ctx ? ctx->fused : -1; }

// ---- multi-block plumbing -------------------------------------------------------------------------------------------
extern "C" int fest3d_gpu_comm_unique_id(char id_out[128]) {
  if (!g_nccl.load()) { fprintf(stderr, "fest3d_gpu: libnccl.so.2 not found\n"); return F3D_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return F3D_ERR_CUDA;
  memcpy(id_out, id.internal, 128);
  return 0;
}

// One communicator per process: every context of a rank shares it (a second ncclCommInitRank with the same id and rank would
// hang), and every NCCL call of the rank goes to ONE communication stream, so that the send / receive lists of two ranks pair
// up whatever the number of blocks each of them owns.
namespace {
struct CommShared {
  ncclUniqueId uid;
  ncclComm_t comm = nullptr;
  int n_ranks = 0, rank = -1, device = -1, refs = 0;
  cudaStream_t stream = nullptr;     // all sends / receives / all-reduces of this process
  cudaEvent_t ev_done = nullptr;     // recorded behind the last NCCL call of an exchange
  int* err_word = nullptr;           // device int: OR of the error words of this rank's contexts, max-reduced over the ranks
  int* err_host = nullptr;           // pinned
};
std::vector<CommShared*> g_comms;
}  // namespace

extern "C" int fest3d_gpu_comm_init(Fest3dGpuCtx* ctx, int n_ranks, int rank, const char id[128], const int* block_to_rank) {
  if (!ctx || !id || !block_to_rank || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, F3D_ERR_ARGUMENT);
  if (!g_nccl.load()) { fprintf(stderr, "fest3d_gpu: libnccl.so.2 not found\n"); return fail(ctx, F3D_ERR_UNSUPPORTED); }
  if (ctx->nccl) return fail(ctx, F3D_ERR_ARGUMENT);   // already initialised
  F3D_CUDA(cudaSetDevice(ctx->device));
  CommShared* cs = nullptr;
  for (CommShared* c : g_comms)
    if (memcmp(c->uid.internal, id, 128) == 0) { cs = c; break; }
  if (cs) {
    // the same rank of the same communicator lives on one device: a process that drives several GPUs needs one rank per device
    if (cs->rank != rank || cs->n_ranks != n_ranks || cs->device != ctx->device) return fail(ctx, F3D_ERR_UNSUPPORTED);
  } else {
    cs = new CommShared();
    memcpy(cs->uid.internal, id, 128);
    cs->n_ranks = n_ranks; cs->rank = rank; cs->device = ctx->device;
    if (g_nccl.CommInitRank(&cs->comm, n_ranks, cs->uid, rank) != 0) { delete cs; return fail(ctx, F3D_ERR_CUDA); }
    bool ok = cudaStreamCreateWithFlags(&cs->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&cs->ev_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&cs->err_word, sizeof(int)) == cudaSuccess && cudaMallocHost((void**)&cs->err_host, sizeof(int)) == cudaSuccess;
    if (!ok) return fail(ctx, F3D_ERR_CUDA);
    g_comms.push_back(cs);
  }
  cs->refs++;
  ctx->nccl = cs; ctx->n_ranks = n_ranks; ctx->rank = rank;
  ctx->block_to_rank.assign(block_to_rank, block_to_rank + ctx->cfg.n_blocks);
  for (int f = 0; f < 6; ++f) {
    Link& lk = ctx->link[f];
    if (lk.neighbour_block < 0 || lk.neighbour_block >= ctx->cfg.n_blocks) continue;
    if (ctx->block_to_rank[lk.neighbour_block] == rank) continue;   // a block of this process: fest3d_gpu_link_local
    lk.rank = ctx->block_to_rank[lk.neighbour_block];
    lk.kind = 2;
  }
  return 0;
}

static void comm_release(Fest3dGpuCtx* ctx) {
  CommShared* cs = (CommShared*)ctx->nccl;
  if (!cs) return;
  ctx->nccl = nullptr;
  if (--cs->refs > 0) return;
  cudaSetDevice(cs->device);
  if (cs->stream) cudaStreamSynchronize(cs->stream);
  if (g_nccl.CommDestroy) g_nccl.CommDestroy(cs->comm);
  if (cs->stream) cudaStreamDestroy(cs->stream);
  if (cs->ev_done) cudaEventDestroy(cs->ev_done);
  if (cs->err_word) cudaFree(cs->err_word);
  if (cs->err_host) cudaFreeHost(cs->err_host);
  g_comms.erase(std::remove(g_comms.begin(), g_comms.end(), cs), g_comms.end());
  delete cs;
}

extern "C" int fest3d_gpu_link_local(Fest3dGpuCtx* a, Fest3dGpuCtx* b) {
  if (!a || !b) return F3D_ERR_ARGUMENT;
  int n = 0;   // a == b: a block that is its own neighbour (periodic with itself through PbcId)
  for (int f = 0; f < 6; ++f) {
    if (a->link[f].neighbour_block == b->cfg.block_id) { a->link[f].kind = 1; a->link[f].peer = b; ++n; }
    if (b->link[f].neighbour_block == a->cfg.block_id) { b->link[f].kind = 1; b->link[f].peer = a; ++n; }
  }
  return n ? 0 : F3D_ERR_ARGUMENT;
}

namespace {

struct Msg { int peer_rank, block, face; Fest3dGpuCtx* ctx; int my_face; };

// apply_interface for every context of this process (interface1.f90:96-493).  No host synchronisation: every context keeps
// its own stream, the hand-overs are events --
//   pack (own stream) -> ev_pack -> { the process's communication stream: all ncclSend / ncclRecv of the rank in one group ;
//   a local neighbour's stream: reads this context's send buffer } -> unpack (own stream) -> ev_halo
// and a context re-packs a send buffer only after the local neighbour that reads it has recorded its ev_halo.
// The swap comes in two halves so that it can be POSTED AHEAD: exchange_post (packs + the NCCL group on the communication stream)
// right behind the sweep that produced the layers, exchange_finish (waits + unpacks) at the head of the next stage -- the
// carry-over of the ghost shell and the Temp refresh of the next iteration then run on the compute stream while the messages are
// in flight (they read or write ghost cells only before the unpack).
int exchange_post(Fest3dGpuCtx** cs, int n) {
  bool any = false, remote = false;
  for (int c = 0; c < n; ++c)
    for (int f = 0; f < 6; ++f) {
      if (!cs[c]->sendbuf[f]) continue;
      // an interface or periodic face nobody is attached to would leave its ghost layers stale: never a silent degraded path
      if (cs[c]->link[f].kind == 0) return fail(cs[c], F3D_ERR_ARGUMENT);
      any = true;
      remote |= cs[c]->link[f].kind == 2;
    }
  if (!any) return 0;
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    for (int f = 0; f < 6; ++f)   // the readers of last stage's send buffers are done
      if (ctx->link[f].kind == 1 && ctx->link[f].peer != ctx) F3D_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->link[f].peer->ev_halo, 0));
    for (int f = 0; f < 6; ++f)
      if (ctx->sendbuf[f]) { int rc = launch_pack(ctx, f + 1); if (rc) return rc; }
    F3D_CUDA(cudaEventRecord(ctx->ev_pack, ctx->stream));
  }
  // remote messages over NCCL, posted in a canonical order per peer so that sends and receives pair up
  CommShared* sh = nullptr;
  if (remote) {
    std::vector<Msg> sends, recvs;
    for (int c = 0; c < n; ++c)
      for (int f = 0; f < 6; ++f) {
        const Link& lk = cs[c]->link[f];
        if (lk.kind != 2) continue;
        if (!cs[c]->nccl) return fail(cs[c], F3D_ERR_ARGUMENT);
        if (sh && sh != (CommShared*)cs[c]->nccl) return fail(cs[c], F3D_ERR_UNSUPPORTED);   // one communicator per process
        sh = (CommShared*)cs[c]->nccl;
        sends.push_back({lk.rank, cs[c]->cfg.block_id, f + 1, cs[c], f + 1});
        recvs.push_back({lk.rank, lk.neighbour_block, cs[c]->cfg.otherface[f], cs[c], f + 1});
      }
    auto key = [](const Msg& a, const Msg& b) { return std::tie(a.peer_rank, a.block, a.face) < std::tie(b.peer_rank, b.block, b.face); };
    std::sort(sends.begin(), sends.end(), key);
    std::sort(recvs.begin(), recvs.end(), key);
    cudaSetDevice(sh->device);
    for (int c = 0; c < n; ++c) {
      bool has = false;
      for (int f = 0; f < 6; ++f) has |= cs[c]->link[f].kind == 2;
      if (has) F3D_CUDA_RC(cudaStreamWaitEvent(sh->stream, cs[c]->ev_pack, 0));
    }
    g_nccl.GroupStart();
    for (const Msg& m : sends) g_nccl.Send(m.ctx->sendbuf[m.my_face - 1], m.ctx->buf_elems[m.my_face - 1], kNcclFloat64, m.peer_rank, sh->comm, sh->stream);
    for (const Msg& m : recvs) g_nccl.Recv(m.ctx->recvbuf[m.my_face - 1], m.ctx->buf_elems[m.my_face - 1], kNcclFloat64, m.peer_rank, sh->comm, sh->stream);
    if (g_nccl.GroupEnd() != 0) return F3D_ERR_CUDA;
    F3D_CUDA_RC(cudaEventRecord(sh->ev_done, sh->stream));
  }
  for (int c = 0; c < n; ++c) cs[c]->halo_posted = true;
  return 0;
}

int exchange_finish(Fest3dGpuCtx** cs, int n) {
  CommShared* sh = nullptr;
  for (int c = 0; c < n; ++c) {
    if (!cs[c]->halo_posted) return 0;   // nothing was posted (no interface at all)
    for (int f = 0; f < 6; ++f) if (cs[c]->link[f].kind == 2) sh = (CommShared*)cs[c]->nccl;
  }
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    ctx->halo_posted = false;
    bool waited_comm = false;
    for (int f = 0; f < 6; ++f) {
      const Link& lk = ctx->link[f];
      if (lk.kind == 0) continue;
      const double* src = ctx->recvbuf[f];
      if (lk.kind == 1) {
        const int of = ctx->cfg.otherface[f];
        const double* peer_buf = lk.peer->sendbuf[of - 1];
        if (lk.peer != ctx) F3D_CUDA(cudaStreamWaitEvent(ctx->stream, lk.peer->ev_pack, 0));
        if (lk.peer->device != ctx->device) {
          F3D_CUDA(cudaMemcpyPeerAsync(ctx->recvbuf[f], ctx->device, peer_buf, lk.peer->device, ctx->buf_elems[f] * sizeof(double), ctx->stream));
        } else {
          src = peer_buf;
        }
      } else if (!waited_comm) {
        F3D_CUDA(cudaStreamWaitEvent(ctx->stream, sh->ev_done, 0));
        waited_comm = true;
      }
      int rc = launch_unpack(ctx, f + 1, src);
      if (rc) return rc;
    }
    F3D_CUDA(cudaEventRecord(ctx->ev_halo, ctx->stream));
  }
  return 0;
}

int exchange(Fest3dGpuCtx** cs, int n) {
  bool posted = true;
  for (int c = 0; c < n; ++c) posted &= cs[c]->halo_posted;
  if (!posted) { int rc = exchange_post(cs, n); if (rc) return rc; }
  return exchange_finish(cs, n);
}

// one get_total_conservative_Residue (+ update) on every context
int stage(Fest3dGpuCtx** cs, int n, bool update, double TF, double SF, int use_sum, int first, int last, bool post_ahead = false, bool implicit = false) {
  int rc = exchange(cs, n);
  if (rc) return rc;
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    if ((rc = launch_bc(ctx))) return rc;
    if (ctx->P.viscous && !ctx->fused && (rc = launch_gradients(ctx))) return rc;   // staged path: gradients + viscosities into HBM
    if (implicit) {   // update.f90:216-219: residual, time step, then the LU-SGS sweeps update qp in place (ghost layers stay as filled)
      if ((rc = launch_residual(ctx, MODE_RESIDUE_ONLY, 1.0, 1.0, 0, 1, 1))) return rc;
      if (ctx->P.time_stepping == 1 && !(ctx->P.global_time_step > 0) && (rc = launch_global_dt(ctx))) return rc;
      if ((rc = launch_lusgs(ctx))) return rc;
      continue;
    }
    if (!update) {
      if ((rc = launch_residual(ctx, MODE_RESIDUE_ONLY, 1.0, 1.0, 0, 1, 0))) return rc;
      continue;
    }
    const bool global_min = first && ctx->P.time_stepping == 1 && !(ctx->P.global_time_step > 0);
    int fst = first;
    if (global_min) {   // delta_t = minval(delta_t) needs the whole field before any cell is updated (time.f90:286)
      if ((rc = launch_residual(ctx, MODE_RESIDUE_ONLY, 1.0, 1.0, 0, 1, 0))) return rc;
      if ((rc = launch_global_dt(ctx))) return rc;
      fst = 0;
    }
    if ((rc = launch_residual(ctx, MODE_UPDATE, TF, SF, use_sum, fst, last))) return rc;
    std::swap(ctx->qp, ctx->qp2);
  }
  if (implicit) return post_ahead ? exchange_post(cs, n) : 0;
  if (!update) return 0;
  // the interior layers of the new state exist: the next stage's swap can leave now, beside the ghost-shell carry-over
  if (post_ahead && (rc = exchange_post(cs, n))) return rc;
  for (int c = 0; c < n; ++c) {
    F3D_CUDA_RC(cudaSetDevice(cs[c]->device));
    if ((rc = launch_ghost_shell_copy(cs[c], cs[c]->qp, cs[c]->qp2))) return rc;
  }
  return 0;
}

int check_errors(Fest3dGpuCtx* ctx) {
  F3D_CUDA(cudaSetDevice(ctx->device));
  F3D_CUDA(cudaMemcpyAsync(ctx->err_host, ctx->err_dev, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  F3D_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->err_host[0]) {
    ctx->last_error.flags |= ctx->err_host[0];
    ctx->last_error.i = ctx->err_host[1]; ctx->last_error.j = ctx->err_host[2]; ctx->last_error.k = ctx->err_host[3];
    return ctx->err_host[0];
  }
  return 0;
}

}  // namespace

extern "C" int fest3d_gpu_residual_group(Fest3dGpuCtx** cs, int n, int current_iter) {
  if (!cs || n < 1) return F3D_ERR_ARGUMENT;
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    { int rp = apply_pending_state(ctx); if (rp) return rp; }
    if (!ctx->geometry_set || !ctx->state_set) return fail(ctx, F3D_ERR_ARGUMENT);
    ctx->P.current_iter = current_iter;
    int rc = launch_temp(ctx);
    if (rc) return rc;
  }
  int rc = stage(cs, n, false, 1.0, 1.0, 0, 1, 0);
  if (rc) return rc;
  int e = 0;
  for (int c = 0; c < n; ++c) e |= check_errors(cs[c]);
  return e;
}

extern "C" int fest3d_gpu_residual(Fest3dGpuCtx* ctx, int current_iter, double* residue_out) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  Fest3dGpuCtx* one[1] = {ctx};
  int rc = fest3d_gpu_residual_group(one, 1, current_iter);
  if (residue_out) { int r2 = fest3d_gpu_get_residue(ctx, residue_out); if (r2) return r2; }
  return rc;
}

namespace {

// one iteration of get_next_solution on every context (update.f90:129-226) + the norm assembly launch of find_resnorm
int issue_iteration(Fest3dGpuCtx** cs, int n, int ta, int iter, bool more_follow = false) {
  // post_ahead: the stage is followed, inside this call, by another stage with nothing in between that changes the interior state
  // (the whole-array blends of the TVD schemes do): A = after an inner stage, Z = after the last stage of the iteration
  const bool A = true, Z = more_follow;
  int rc = 0;
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    ctx->P.current_iter = iter;
    if ((rc = launch_temp(ctx))) return rc;   // update.f90:170
    if (ta != F3D_T_NONE && ta != F3D_T_IMPLICIT && (rc = launch_copy_fields(ctx, ctx->ustore, ctx->qp, ctx->P.L.nv))) return rc;       // U_store = qp
    if ((ta == F3D_T_RK2 || ta == F3D_T_RK4) && (rc = launch_zero_fields(ctx, ctx->rstore, ctx->P.L.nv))) return rc;  // R_store = 0
  }
  auto blend_all = [&](double a, double b) { for (int c = 0; c < n && !rc; ++c) { cudaSetDevice(cs[c]->device); rc = launch_blend(cs[c], a, b); } return rc; };
  switch (ta) {   // update.f90:171-215
    case F3D_T_NONE:
      rc = stage(cs, n, true, 1., 1., 0, 1, 1, Z); break;
    case F3D_T_RK4:
      if ((rc = stage(cs, n, true, 0.5, 1., 0, 1, 0, A))) break;
      if ((rc = stage(cs, n, true, 0.5, 2., 0, 0, 0, A))) break;
      if ((rc = stage(cs, n, true, 1.0, 2., 0, 0, 0, A))) break;
      rc = stage(cs, n, true, 1. / 6., 1., 1, 0, 1, Z); break;
    case F3D_T_RK2:
      if ((rc = stage(cs, n, true, 0.5, 1., 0, 1, 0, A))) break;
      rc = stage(cs, n, true, 0.5, 1., 1, 0, 1, Z); break;
    case F3D_T_TVDRK3:
      if ((rc = stage(cs, n, true, 1.0, 1., 0, 1, 0, A))) break;
      if ((rc = stage(cs, n, true, 1.0, 1., 0, 0, 0))) break;
      if ((rc = blend_all(0.75, 0.25))) break;
      if ((rc = stage(cs, n, true, 1.0, 1., 0, 0, 1))) break;
      rc = blend_all((1. / 3.), (2. / 3.)); break;
    case F3D_T_TVDRK2:
      if ((rc = stage(cs, n, true, 1.0, 1., 0, 1, 0, A))) break;
      if ((rc = stage(cs, n, true, 1.0, 1., 0, 0, 1))) break;
      rc = blend_all(0.5, 0.5); break;
    case F3D_T_IMPLICIT:
      rc = stage(cs, n, false, 1., 1., 0, 1, 1, Z, true); break;
    default: rc = F3D_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  for (int c = 0; c < n; ++c) { cudaSetDevice(cs[c]->device); if ((rc = launch_norms(cs[c]))) return rc; }
  return 0;
}

// CUDA graph of one iteration.  Small blocks are launch-bound (the SmoothBump case: ~45 launches of a few microseconds of work per
// RK4 iteration), so from the third iteration on (the fixed-value boundary conditions stop changing with current_iter then,
// bc_primitive.f90:236) the whole iteration -- every stage of every context, their streams forked from and joined to the first
// context's stream -- is captured once per buffer parity and replayed.  Not used with NCCL links (their host-side grouping),
// with kernel timing, or across devices; F3D_GRAPHS=0 turns it off.
bool graphs_allowed(Fest3dGpuCtx** cs, int n) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("F3D_GRAPHS"); on = (e && e[0] == '0') ? 0 : 1; }
  if (!on) return false;
  for (int c = 0; c < n; ++c) {
    if (cs[c]->timing || cs[c]->device != cs[0]->device) return false;
    for (int f = 0; f < 6; ++f) {
      const Link& lk = cs[c]->link[f];
      if (lk.kind == 2) return false;
      if (lk.kind == 1) {   // every local neighbour must be part of this group (its stream is forked into the capture)
        bool in = false;
        for (int d = 0; d < n; ++d) in |= cs[d] == lk.peer;
        if (!in) return false;
      }
    }
  }
  return true;
}

// the graph runs on the first context's stream: everything queued on the other contexts' streams goes before it ...
int graph_join_before(Fest3dGpuCtx** cs, int n) {
  for (int c = 1; c < n; ++c) {
    F3D_CUDA_RC(cudaEventRecord(cs[c]->ev_halo, cs[c]->stream));
    F3D_CUDA_RC(cudaStreamWaitEvent(cs[0]->stream, cs[c]->ev_halo, 0));
  }
  return 0;
}
// ... and whatever is queued on them afterwards (norm download, the next direct iteration) comes after it
int graph_fork_after(Fest3dGpuCtx** cs, int n) {
  if (n > 1) F3D_CUDA_RC(cudaEventRecord(cs[0]->ev_pack, cs[0]->stream));
  for (int c = 1; c < n; ++c) F3D_CUDA_RC(cudaStreamWaitEvent(cs[c]->stream, cs[0]->ev_pack, 0));
  return 0;
}

int run_iteration_graph(Fest3dGpuCtx** cs, int n, int ta, int iter, bool* used) {
  *used = false;
  Fest3dGpuCtx* c0 = cs[0];
  unsigned long long key = 1469598103934665603ULL;
  auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ULL; };
  mix((unsigned long long)n); mix((unsigned long long)ta);
  for (int c = 0; c < n; ++c) { mix((unsigned long long)(uintptr_t)cs[c]); mix((unsigned long long)(uintptr_t)cs[c]->qp); mix((unsigned long long)(uintptr_t)cs[c]->stream); }
  auto it = c0->graphs.find(key);
  if (it == c0->graphs.end()) {
    // capture.  Host-side state the issue path advances (buffer swaps, launch counters) is advanced by the capture pass itself.
    std::vector<double*> q0(n), q1(n);
    std::vector<long long> l0(n);
    for (int c = 0; c < n; ++c) { q0[c] = cs[c]->qp; q1[c] = cs[c]->qp2; l0[c] = cs[c]->launches; }
    cudaSetDevice(c0->device);
    if (cudaStreamBeginCapture(c0->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return 0; }
    bool ok = true;
    for (int c = 1; c < n && ok; ++c) {
      ok = cudaEventRecord(c0->ev_pack, c0->stream) == cudaSuccess && cudaStreamWaitEvent(cs[c]->stream, c0->ev_pack, 0) == cudaSuccess;
    }
    // every event the exchange waits on must have been recorded inside the capture (the first stage would otherwise wait on the
    // previous iteration's, across the capture boundary)
    for (int c = 0; c < n && ok; ++c)
      ok = cudaEventRecord(cs[c]->ev_halo, cs[c]->stream) == cudaSuccess && (c == 0 || cudaEventRecord(cs[c]->ev_pack, cs[c]->stream) == cudaSuccess);
    int rc = ok ? issue_iteration(cs, n, ta, iter) : F3D_ERR_CUDA;
    for (int c = 1; c < n; ++c) {
      if (cudaEventRecord(cs[c]->ev_halo, cs[c]->stream) != cudaSuccess || cudaStreamWaitEvent(c0->stream, cs[c]->ev_halo, 0) != cudaSuccess) ok = false;
    }
    cudaGraph_t g = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(c0->stream, &g);
    Ctx::IterGraph ig;
    if (rc == 0 && ok && ee == cudaSuccess && g && cudaGraphInstantiate(&ig.exec, g, 0) == cudaSuccess) {
      ig.launches = c0->launches - l0[0];
      ig.swaps = c0->qp != q0[0];
      cudaGraphDestroy(g);
      it = c0->graphs.emplace(key, ig).first;
      // the capture pass has advanced the host-side state of this iteration; now run it
      if (graph_join_before(cs, n)) return F3D_ERR_CUDA;
      if (cudaGraphLaunch(it->second.exec, c0->stream) != cudaSuccess) return F3D_ERR_CUDA;
      if (graph_fork_after(cs, n)) return F3D_ERR_CUDA;
      *used = true;
      return 0;
    }
    // not capturable here: undo the host-side state, remember not to try again for this key, issue directly
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    for (int c = 0; c < n; ++c) { cs[c]->qp = q0[c]; cs[c]->qp2 = q1[c]; cs[c]->launches = l0[c]; }
    c0->graphs.emplace(key, Ctx::IterGraph{});
    return 0;
  }
  if (!it->second.exec) return 0;
  cudaSetDevice(c0->device);
  if (graph_join_before(cs, n)) return F3D_ERR_CUDA;
  if (cudaGraphLaunch(it->second.exec, c0->stream) != cudaSuccess) return F3D_ERR_CUDA;
  if (graph_fork_after(cs, n)) return F3D_ERR_CUDA;
  for (int c = 0; c < n; ++c) {
    cs[c]->launches += it->second.launches;
    cs[c]->P.current_iter = iter;
    if (it->second.swaps) std::swap(cs[c]->qp, cs[c]->qp2);
  }
  *used = true;
  return 0;
}

}  // namespace

namespace {

int step_prepare(Fest3dGpuCtx** cs, int n, int n_iters, int* ta_out) {
  if (!cs || n < 1 || n_iters < 1) return F3D_ERR_ARGUMENT;
  const int ta = cs[0]->cfg.time_accuracy;
  for (int c = 0; c < n; ++c) {
    cudaSetDevice(cs[c]->device);
    int rp = apply_pending_state(cs[c]);
    if (rp) return rp;
    if (!cs[c]->geometry_set || !cs[c]->state_set || cs[c]->cfg.time_accuracy != ta) return fail(cs[c], F3D_ERR_ARGUMENT);
  }
  *ta_out = ta;
  return 0;
}

// queue the iterations it0 .. it0+nit-1 of a call and the download of their norms and of the error word; no host synchronisation
int chunk_issue(Fest3dGpuCtx** cs, int n, int ta, int current_iter, int it0, int nit, int n_iters, bool graphs) {
  const int nvp1 = cs[0]->P.L.nv + 1;
  int rc = 0;
  for (int c = 0; c < n; ++c) {   // the norm slot counter of the chunk starts at 0
    F3D_CUDA_RC(cudaSetDevice(cs[c]->device));
    F3D_CUDA_RC(cudaMemsetAsync(cs[c]->err_dev + 4, 0, sizeof(int), cs[c]->stream));
  }
  for (int it = 0; it < nit; ++it) {
    const int iter = current_iter + it0 + it;
    bool used = false;
    if (graphs && iter > 2 && (rc = run_iteration_graph(cs, n, ta, iter, &used))) return rc;
    // outside a graph the last stage may post the next iteration's swap ahead, as long as that iteration belongs to this call
    if (!used && (rc = issue_iteration(cs, n, ta, iter, !graphs && (it0 + it + 1 < n_iters)))) return rc;
  }
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    F3D_CUDA(cudaMemcpyAsync(ctx->norms_host, ctx->norms_dev, sizeof(double) * nit * nvp1, cudaMemcpyDeviceToHost, ctx->stream));
    F3D_CUDA(cudaMemcpyAsync(ctx->err_host, ctx->err_dev, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  return 0;
}

// find_resnorm: sum over all blocks (MPI_ALLGATHER + sum, resnorm.f90:201-225), then sqrt / abs.  The blocks of this process are added
// on the host; across ranks ONE ncclAllReduce per chunk carries the sums and, in an extra slot, the error state, so that every rank
// leaves with an error as soon as any rank has one (the reference's Fatal_error stops the whole job) instead of posting the next
// chunk's sends and receives towards a rank that has already returned.
int chunk_collect(Fest3dGpuCtx** cs, int n, int it0, int nit, double* res_abs_out) {
  const int nvp1 = cs[0]->P.L.nv + 1;
  CommShared* sh = nullptr;
  int e = 0;
  for (int c = 0; c < n; ++c) {
    Fest3dGpuCtx* ctx = cs[c];
    F3D_CUDA(cudaSetDevice(ctx->device));
    if (ctx->nccl && ctx->n_ranks > 1) sh = (CommShared*)ctx->nccl;
    F3D_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->err_host[0]) {
      ctx->last_error.flags |= ctx->err_host[0];
      ctx->last_error.i = ctx->err_host[1]; ctx->last_error.j = ctx->err_host[2]; ctx->last_error.k = ctx->err_host[3];
      e |= ctx->err_host[0];
    }
  }
  double* tot = cs[0]->norms_host;
  for (int x = 0; x < nit * nvp1; ++x) {
    double sum = 0.0;
    for (int c = 0; c < n; ++c) sum = sum + cs[c]->norms_host[x];
    tot[x] = sum;
  }
  if (sh) {
    Fest3dGpuCtx* ctx = cs[0];
    const int cnt = nit * nvp1 + 1;
    tot[cnt - 1] = e ? 1.0 : 0.0;
    F3D_CUDA(cudaSetDevice(sh->device));
    F3D_CUDA(cudaMemcpyAsync(ctx->norms_dev, tot, sizeof(double) * cnt, cudaMemcpyHostToDevice, sh->stream));
    if (g_nccl.AllReduce(ctx->norms_dev, ctx->norms_dev, (size_t)cnt, kNcclFloat64, kNcclSum, sh->comm, sh->stream) != 0) return fail(ctx, F3D_ERR_CUDA);
    F3D_CUDA(cudaMemcpyAsync(tot, ctx->norms_dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, sh->stream));
    F3D_CUDA(cudaStreamSynchronize(sh->stream));
    if (tot[cnt - 1] != 0.0 && !e) { e = F3D_ERR_PEER; for (int c = 0; c < n; ++c) cs[c]->last_error.flags |= F3D_ERR_PEER; }
  }
  if (res_abs_out) {
    for (int it = 0; it < nit; ++it)
      for (int l = 0; l < nvp1; ++l) res_abs_out[(size_t)(it0 + it) * nvp1 + l] = (l == 0) ? fabs(tot[it * nvp1 + l]) : sqrt(tot[it * nvp1 + l]);
  }
  return e;
}

}  // namespace

extern "C" int fest3d_gpu_step_group(Fest3dGpuCtx** cs, int n, int current_iter, int n_iters, double* res_abs_out) {
  int ta = 0;
  int rc = step_prepare(cs, n, n_iters, &ta);
  if (rc) return rc;
  const int chunk = 1016 / (cs[0]->P.L.nv + 1);   // norms of a chunk + the error slot fit the 1024-double norm buffers
  const bool graphs = graphs_allowed(cs, n);
  for (int it0 = 0; it0 < n_iters; it0 += chunk) {
    const int nit = std::min(chunk, n_iters - it0);
    if ((rc = chunk_issue(cs, n, ta, current_iter, it0, nit, n_iters, graphs))) return rc;
    if ((rc = chunk_collect(cs, n, it0, nit, res_abs_out))) return rc;
  }
  return 0;
}

// The same in two halves: _begin queues the iterations and returns, _end waits for them and delivers the norms.  Between the two the
// host is free -- e.g. to start the upload of the next state (fest3d_gpu_set_state_async) while these iterations run.
extern "C" int fest3d_gpu_step_group_begin(Fest3dGpuCtx** cs, int n, int current_iter, int n_iters) {
  int ta = 0;
  int rc = step_prepare(cs, n, n_iters, &ta);
  if (rc) return rc;
  if (n_iters > 1016 / (cs[0]->P.L.nv + 1) || cs[0]->steps_in_flight) return fail(cs[0], F3D_ERR_ARGUMENT);
  if ((rc = chunk_issue(cs, n, ta, current_iter, 0, n_iters, n_iters, graphs_allowed(cs, n)))) return rc;
  cs[0]->steps_in_flight = n_iters;
  return 0;
}

extern "C" int fest3d_gpu_step_group_end(Fest3dGpuCtx** cs, int n, double* res_abs_out) {
  if (!cs || n < 1 || !cs[0]->steps_in_flight) return F3D_ERR_ARGUMENT;
  const int nit = cs[0]->steps_in_flight;
  cs[0]->steps_in_flight = 0;
  return chunk_collect(cs, n, 0, nit, res_abs_out);
}

extern "C" int fest3d_gpu_step(Fest3dGpuCtx* ctx, int current_iter, int n_iters, double* res_abs_out) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  Fest3dGpuCtx* one[1] = {ctx};
  return fest3d_gpu_step_group(one, 1, current_iter, n_iters, res_abs_out);
}

// ghost_grid + the metric set-up of geometry.f90 on the device (geometry.cu) instead of the AoS upload
extern "C" int fest3d_gpu_setup_geometry(Fest3dGpuCtx* ctx, const double* grid_xyz, const double* dist, double* nodes_out) {
  if (!ctx || !grid_xyz) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  int rc = launch_setup_geometry(ctx, grid_xyz, nodes_out);
  if (rc) return rc;
  if (dist) {
    if ((rc = copy_cells(ctx, ctx->geom + (long long)G_DIST * L.fs, const_cast<double*>(dist), 1, true, -2, L.imx + 5, L.jmx + 5, L.kmx + 5))) return rc;
  }
  if ((rc = init_aux_fields(ctx))) return rc;
  if ((rc = check_errors(ctx))) return rc;   // non-positive volume -> F3D_ERR_GEOMETRY with the cell (geometry.f90:476-494)
  ctx->geometry_set = true;
  return 0;
}

extern "C" int fest3d_gpu_get_geometry(Fest3dGpuCtx* ctx, double* cells, double* Ifaces, double* Jfaces, double* Kfaces) {
  if (!ctx || !ctx->geometry_set) return fail(ctx, F3D_ERR_ARGUMENT);
  F3D_CUDA(cudaSetDevice(ctx->device));
  const Layout& L = ctx->P.L;
  const long long fs = L.fs;
  int rc = 0;
  if (cells && (rc = download_records(ctx, ctx->geom + (long long)G_VOL * fs, L.imx + 5, L.jmx + 5, L.kmx + 5, cells))) return rc;
  if (Ifaces && (rc = download_records(ctx, ctx->geom + (long long)G_IA * fs, L.imx + 6, L.jmx + 5, L.kmx + 5, Ifaces))) return rc;
  if (Jfaces && (rc = download_records(ctx, ctx->geom + (long long)G_JA * fs, L.imx + 5, L.jmx + 6, L.kmx + 5, Jfaces))) return rc;
  if (Kfaces && (rc = download_records(ctx, ctx->geom + (long long)G_KA * fs, L.imx + 5, L.jmx + 5, L.kmx + 6, Kfaces))) return rc;
  return 0;
}
