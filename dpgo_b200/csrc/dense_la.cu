// In-tree FP64 dense linear algebra on the device (sm_100a): kernels around the device functions of
// dense_la.cuh and the CUDA back end of the launch sequences in dense_la_seq.h.  Replaces the
// cuSOLVER potrf/potri + cuBLAS symm/gemm calls of the round-1 preconditioner set-up
// (ref: src/PoseGraph.cpp:598-613).
#include "dense_la.h"

#include <cstdarg>
#include <cstdio>

#include "dense_la_seq.h"

namespace dpgo {
namespace dla {

static thread_local char g_err[256] = "";
const char *last_error() { return g_err; }
static void set_err(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void __launch_bounds__(kThreads) k_dla_gemm(const GemmDesc *descs, GemmFlags f) {
  __shared__ __align__(32) double sA[2 * BK * LDS];
  __shared__ __align__(32) double sB[2 * BK * LDS];
  const GemmDesc g = descs[blockIdx.z];
  dla_gemm_tile(g, f, (int)blockIdx.x, (int)blockIdx.y, sA, sB);
}

constexpr int kDiagSmem = 2 * TS * (TS + 1) * (int)sizeof(double);

__global__ void __launch_bounds__(kThreads) k_dla_diag(const SpdDesc *descs, int panel, int *info) {
  extern __shared__ __align__(16) unsigned char dla_dsm[];
  double *sL = reinterpret_cast<double *>(dla_dsm);
  double *sW = sL + TS * (TS + 1);
  const SpdDesc d = descs[blockIdx.z];
  dla_diag_block(d, panel, (int)blockIdx.z, sL, sW, info);
}

__global__ void __launch_bounds__(kThreads) k_dla_copy(const SpdDesc *descs, int what) {
  const SpdDesc d = descs[blockIdx.z];
  const int bi = (int)blockIdx.x, bj = what == 0 ? (int)blockIdx.x : (int)blockIdx.y;
  dla_copy_tile(d, what, bi, bj);
}

namespace {

struct CudaBackend {
  cudaStream_t st;
  bool ok = true;
  bool check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
      if (ok) set_err("dense_la: %s: %s", what, cudaGetErrorString(e));
      ok = false;
      return false;
    }
    return true;
  }
  // stream-ordered scratch (the device's default memory pool keeps freed blocks: dpgo_create raises its release
  // threshold), so a rebuild of the preconditioner at every GNC weight update does not pay for cudaMalloc / cudaFree
  void *alloc(size_t bytes) {
    void *p = nullptr;
    if (!check(cudaMallocAsync(&p, bytes, st), "cudaMallocAsync")) return nullptr;
    if (!check(cudaMemsetAsync(p, 0, bytes, st), "cudaMemsetAsync")) return nullptr;
    return p;
  }
  void release(void *p) {
    if (p) cudaFreeAsync(p, st);
  }
  bool upload(void *dst, const void *src, size_t bytes) {
    // pageable source: the copy has left the host buffer when the call returns
    return check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "upload");
  }
  bool download(void *dst, const void *src, size_t bytes) {
    if (!check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st), "download")) return false;
    return check(cudaStreamSynchronize(st), "synchronize") && ok;
  }
  void diag(const SpdDesc *d, int panel, int count, int *info) {
    static bool configured = false;
    if (!configured) {
      check(cudaFuncSetAttribute(k_dla_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem), "diag smem");
      configured = true;
    }
    k_dla_diag<<<dim3(1, 1, count), kThreads, kDiagSmem, st>>>(d, panel, info);
    check(cudaPeekAtLastError(), "k_dla_diag");
  }
  void gemm(const GemmDesc *g, int tiles_m, int tiles_n, int count, const GemmFlags &f) {
    // blockIdx.z carries the batch (<= 65535 per launch)
    for (int z0 = 0; z0 < count; z0 += 65535) {
      const int nz = count - z0 < 65535 ? count - z0 : 65535;
      k_dla_gemm<<<dim3(tiles_m, tiles_n, nz), kThreads, 0, st>>>(g + z0, f);
    }
    check(cudaPeekAtLastError(), "k_dla_gemm");
  }
  void copy(const SpdDesc *d, int what, int tiles, int count) {
    k_dla_copy<<<dim3(tiles, what == 0 ? 1 : tiles, count), kThreads, 0, st>>>(d, what);
    check(cudaPeekAtLastError(), "k_dla_copy");
  }
};

}  // namespace

int spd_inverse_batched(cudaStream_t st, const SpdItem *items, int count, bool symmetrize) {
  CudaBackend be{st};
  std::vector<SeqItem> v(count > 0 ? count : 0);
  for (int b = 0; b < count; ++b) v[b] = SeqItem{items[b].A, items[b].n, items[b].lda};
  const int rc = spd_inverse_seq(be, v.data(), count, symmetrize);
  if (!be.ok) return -1;
  if (rc < 0) set_err("dense_la: out of device memory in the SPD inverse");
  return rc;
}

int gemm_batched(cudaStream_t st, const GemmDesc *descs, int count, const GemmFlags &flags) {
  if (count <= 0) return 0;
  CudaBackend be{st};
  int tm = 0, tn = 0;
  for (int b = 0; b < count; ++b) {
    tm = tiles_of(descs[b].M) > tm ? tiles_of(descs[b].M) : tm;
    tn = tiles_of(descs[b].N) > tn ? tiles_of(descs[b].N) : tn;
  }
  if (tm == 0 || tn == 0) return 0;
  GemmDesc *d = nullptr;   // stream-ordered: allocated, read by the queued kernels and released on st
  if (!be.check(cudaMallocAsync((void **)&d, (size_t)count * sizeof(GemmDesc), st), "cudaMallocAsync")) return -1;
  be.upload(d, descs, (size_t)count * sizeof(GemmDesc));
  be.gemm(d, tm, tn, count, flags);
  be.check(cudaFreeAsync(d, st), "cudaFreeAsync");
  return be.ok ? 0 : -1;
}

}  // namespace dla
}  // namespace dpgo
