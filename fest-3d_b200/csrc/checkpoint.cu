// Checkpoint / restart without stalling the solver -- SURVEY 8(f) rank 3.  The reference's checkpoint (src/solver.f90:139-186 ->
// src/read_write/write/dump_solution.f90, write_output*.f90) formats the whole qp array as text on the rank that steps the block;
// at 256^3 that costs more than thousands of device iterations, and a drop-in host would first have to pull qp back
// synchronously (fest3d_gpu_get_state).  Here the state is SNAPSHOT in stream order (one re-layout kernel into a private device
// buffer in the reference's qp(-2:imx+2,-2:jmx+2,-2:kmx+2,1:n_var) layout), the device->host copy runs on a copy stream into
// pinned memory while the next iterations execute, and a writer thread puts it on disk in a binary side format:
//   64-byte header { "F3DCKPT1", int32 imx, jmx, kmx, n_var, iter, 3 x int32 0, uint64 n_doubles, 16 bytes 0 } + n_doubles x float64.
// fest3d_gpu_restart reads it back (header checked against the context) and uploads it: since every other device field is
// re-derived from qp each iteration (Temp, delta_t, gradients, mu / mu_t / F1, RK stores), a restarted run continues bit for bit.
#include "ctx.hpp"
#include <unistd.h>
#include <cstring>
#include <string>
#include <thread>

extern "C" int fest3d_gpu_set_state(Fest3dGpuCtx* ctx, const double* qp);

namespace f3d {

struct CkptHeader {
  char magic[8];
  int32_t imx, jmx, kmx, n_var, iter, zero[3];
  uint64_t n_doubles;
  char pad[16];
};
static_assert(sizeof(CkptHeader) == 64, "checkpoint header is 64 bytes");

struct Checkpoint {
  double* dev = nullptr;    // the snapshot, reference layout
  double* host = nullptr;   // pinned
  size_t n = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_snap = nullptr, ev_done = nullptr;
  std::thread writer;
  bool pending = false;
  int write_rc = 0;
};

static size_t state_doubles(const Layout& L) { return (size_t)L.nv * (L.imx + 5) * (L.jmx + 5) * (L.kmx + 5); }

static int checkpoint_join(Ctx* ctx) {
  Checkpoint* ck = ctx->ckpt;
  if (!ck || !ck->pending) return 0;
  if (ck->writer.joinable()) ck->writer.join();
  ck->pending = false;
  if (ck->write_rc) { ctx->last_error.flags |= F3D_ERR_IO; return F3D_ERR_IO; }
  return 0;
}

void checkpoint_free(Ctx* ctx) {
  Checkpoint* ck = ctx->ckpt;
  if (!ck) return;
  checkpoint_join(ctx);
  if (ck->dev) cudaFree(ck->dev);
  if (ck->host) cudaFreeHost(ck->host);
  if (ck->copy_stream) cudaStreamDestroy(ck->copy_stream);
  if (ck->ev_snap) cudaEventDestroy(ck->ev_snap);
  if (ck->ev_done) cudaEventDestroy(ck->ev_done);
  delete ck;
  ctx->ckpt = nullptr;
}

static int checkpoint_begin(Ctx* ctx, const char* path, int iter) {
  int rc = checkpoint_join(ctx);   // one checkpoint in flight per context: the previous one must be on disk before its buffers are reused
  if (rc) return rc;
  const Layout& L = ctx->P.L;
  if (!ctx->ckpt) {
    Checkpoint* ck = new Checkpoint;
    ctx->ckpt = ck;
    ck->n = state_doubles(L);
    cudaError_t e = cudaMalloc((void**)&ck->dev, ck->n * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ck->host, ck->n * sizeof(double));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ck->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ck->ev_snap, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ck->ev_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {   // never leave a half-built checkpoint state behind: the next begin starts from scratch
      checkpoint_free(ctx);
      F3D_CUDA(e);
    }
  }
  Checkpoint* ck = ctx->ckpt;
  if ((rc = launch_state_relayout(ctx, ctx->qp, ck->dev, 0))) return rc;   // the snapshot: stream order, later iterations do not touch it
  F3D_CUDA(cudaEventRecord(ck->ev_snap, ctx->stream));
  F3D_CUDA(cudaStreamWaitEvent(ck->copy_stream, ck->ev_snap, 0));
  F3D_CUDA(cudaMemcpyAsync(ck->host, ck->dev, ck->n * sizeof(double), cudaMemcpyDeviceToHost, ck->copy_stream));
  F3D_CUDA(cudaEventRecord(ck->ev_done, ck->copy_stream));
  CkptHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "F3DCKPT1", 8);
  h.imx = L.imx; h.jmx = L.jmx; h.kmx = L.kmx; h.n_var = L.nv; h.iter = iter; h.n_doubles = ck->n;
  ck->write_rc = 0;
  ck->pending = true;
  const std::string file(path);
  const int device = ctx->device;
  ck->writer = std::thread([ck, h, file, device]() {
    cudaSetDevice(device);
    if (cudaEventSynchronize(ck->ev_done) != cudaSuccess) { ck->write_rc = 1; return; }
    // write beside the target, flush to the device, then rename: a crash, kill or full disk mid-write leaves the previous good
    // checkpoint of the same name untouched (the reference purges old time directories only after the new one is complete)
    const std::string tmp = file + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { ck->write_rc = 2; return; }
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(ck->host, sizeof(double), ck->n, f) == ck->n;
    ok = ok && fflush(f) == 0 && fsync(fileno(f)) == 0;
    if (fclose(f) != 0 || !ok) { ck->write_rc = 3; remove(tmp.c_str()); return; }
    if (rename(tmp.c_str(), file.c_str()) != 0) { ck->write_rc = 4; remove(tmp.c_str()); }
  });
  return 0;
}

static int restart(Ctx* ctx, const char* path, int* iter) {
  const Layout& L = ctx->P.L;
  FILE* f = fopen(path, "rb");
  if (!f) { ctx->last_error.flags |= F3D_ERR_IO; return F3D_ERR_IO; }
  CkptHeader h;
  int rc = 0;
  double* buf = nullptr;
  const size_t n = state_doubles(L);
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "F3DCKPT1", 8) != 0) rc = F3D_ERR_IO;
  else if (h.imx != L.imx || h.jmx != L.jmx || h.kmx != L.kmx || h.n_var != L.nv || h.n_doubles != n) rc = F3D_ERR_ARGUMENT;   // another block / model
  if (!rc && cudaMallocHost((void**)&buf, n * sizeof(double)) != cudaSuccess) rc = F3D_ERR_CUDA;
  if (!rc && fread(buf, sizeof(double), n, f) != n) rc = F3D_ERR_IO;
  fclose(f);
  if (!rc) rc = fest3d_gpu_set_state(static_cast<Fest3dGpuCtx*>(ctx), buf);
  if (buf) cudaFreeHost(buf);
  if (rc) { ctx->last_error.flags |= rc; return rc; }
  if (iter) *iter = h.iter;
  return 0;
}

}  // namespace f3d

extern "C" int fest3d_gpu_checkpoint_begin(Fest3dGpuCtx* ctx, const char* path, int iter) {
  if (!ctx || !path || !ctx->state_set) { if (ctx) ctx->last_error.flags |= F3D_ERR_ARGUMENT; return F3D_ERR_ARGUMENT; }
  F3D_CUDA(cudaSetDevice(ctx->device));
  return f3d::checkpoint_begin(ctx, path, iter);
}

extern "C" int fest3d_gpu_checkpoint_wait(Fest3dGpuCtx* ctx) {
  if (!ctx) return F3D_ERR_ARGUMENT;
  return f3d::checkpoint_join(ctx);
}

extern "C" int fest3d_gpu_restart(Fest3dGpuCtx* ctx, const char* path, int* iter) {
  if (!ctx || !path) { if (ctx) ctx->last_error.flags |= F3D_ERR_ARGUMENT; return F3D_ERR_ARGUMENT; }
  F3D_CUDA(cudaSetDevice(ctx->device));
  return f3d::restart(ctx, path, iter);
}
