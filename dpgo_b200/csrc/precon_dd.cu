// Two-level (domain decomposition) exact preconditioner: set-up and stand-alone application.
//
// Same operator as the reference's CHOLMOD solve with Q + 0.1 I (ref: src/PoseGraph.cpp:598-613,
// src/QuadraticProblem.cpp:56-69), computed by exact block elimination:
//   1. nested dissection of the pose graph on the host (BFS level-set vertex separators) ->
//      interior domains D_1..D_K with no edges between them and a separator S;
//   2. dense inverses A_k^-1 of every interior block and Sigma^-1 of the Schur complement
//      Sigma = A_SS - sum_k A_Sk A_k^-1 A_kS (in-tree batched Cholesky / inverse / tile GEMM of
//      dense_la.cu, all domains in lockstep, couplings restricted to the separator columns S_k
//      each domain touches; set-up only);
//   3. application = 3 streamed dense block products + 2 tiny sparse products (kernels.cuh).
// Bytes per application: sum_k 2 (n_k(d+1))^2 8 + (|S|(d+1))^2 8  (58 MB instead of 800 MB on
// sphere2500, L2-resident).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <vector>

#include "dense_la.h"
#include "device_state.h"
#include "dissect.h"
#include "kernels.cuh"

namespace dpgo {

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dpgo::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,              \
                      cudaGetErrorString(_e));                                           \
      return DPGO_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)
#define DPGO_TRY(expr)                                                                   \
  do {                                                                                   \
    int _rc = (expr);                                                                    \
    if (_rc != DPGO_OK) return _rc;                                                      \
  } while (0)

struct DdState {
  int nS = 0, nI = 0, nB = 0, K = 0, sep_col0 = 0, pcols = 0, nsplit1 = 1, nsplit3 = 1, nstrips1 = 0, nstrips3 = 0;
  int V = 1;                       // virtual CTAs the strips are balanced over (= SMs)
  double bytes_per_apply = 0;
  double *M1 = nullptr, *M3 = nullptr;
  DdStrip *strips1 = nullptr, *strips3 = nullptr;
  int *cta1 = nullptr, *cta3 = nullptr, *chunks1 = nullptr, *chunks3 = nullptr;
  int *pcol = nullptr, *srow = nullptr, *bcol = nullptr, *icol = nullptr;
  int *si_rowptr = nullptr, *si_colidx = nullptr, *bs_rowptr = nullptr, *bs_colidx = nullptr;
  double *si_blocks = nullptr, *bs_blocks = nullptr;
  int *d_si_src = nullptr, *d_bs_src = nullptr;   // block of Q each coupling block is a copy of (weight-only refresh)
  double *y = nullptr, *t = nullptr, *zs = nullptr, *u = nullptr, *w = nullptr, *rp = nullptr;
  bool configured = false;
  // symbolic part kept for weight-only rebuilds (GNC): the dissection, the strip tables and the index lists depend on
  // the sparsity pattern only, so dpgo_update_weights redoes the numeric part alone (dd_numeric)
  bool symbolic = false;
  int sym_n = 0, sym_nnzb = 0, sym_thr = 0, sym_split1 = 0, sym_split3 = 0, mS = 0, padS = 0;
  std::vector<int> h_group, h_lpos, h_dom_m, h_dom_pad, h_sk_ptr, h_sk, si_src, bs_src;
  std::vector<long long> h_dom_base;
};

namespace {

template <typename T>
int upload_vec(T **dptr, const std::vector<T> &v) {
  const size_t ne = std::max<size_t>(v.size(), 1);
  CUDA_TRY(cudaMalloc((void **)dptr, ne * sizeof(T)));
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return DPGO_OK;
}

}  // namespace

// ---- set-up kernels ---------------------------------------------------------------------------------
// stage-major strips of a symmetric m x m block (lower triangle of col-major A valid), zero padded to
// pad x pad: dst[((cb * (pad/32) + chunk) * 32 + kk) * 64 + jj] = A(cb*64 + jj, chunk*32 + kk)
__global__ void k_dd_layout(const double *A, int m, int lda, int pad, double *dst) {
  const size_t total = (size_t)pad * pad;
  const int nchunks = pad / kStageK;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int jj = (int)(t % kGemvCols);
    const size_t u = t / kGemvCols;
    const int kk = (int)(u % kStageK);
    const size_t v = u / kStageK;
    const int chunk = (int)(v % nchunks);
    const int cb = (int)(v / nchunks);
    const size_t j = (size_t)cb * kGemvCols + jj, k = (size_t)chunk * kStageK + kk;
    double val = 0.0;
    if (j < (size_t)m && k < (size_t)m) val = (j >= k) ? A[j + k * (size_t)lda] : A[k + j * (size_t)lda];
    dst[t] = val;
  }
}

// Sigma(colmap[i], colmap[j]) -= W(i, j) : the Schur update of one domain touches only S_k x S_k
__global__ void k_dd_scatter_sub(double *Sg, int ldS, const int *colmap, const double *W, int ncomp) {
  const size_t total = (size_t)ncomp * ncomp;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % ncomp), j = (int)(t / ncomp);
    Sg[(size_t)colmap[i] + (size_t)colmap[j] * ldS] -= W[t];
  }
}

// All dense blocks of the block elimination in one pass over the block-CSR entries of Q:
//   interior x interior (same domain k)  -> A_k   (m_k x m_k at a_off[k], + shift on the diagonal)
//   interior k x separator               -> B_k   (m_k x tm_k at b_off[k]; column = position of the separator
//                                                  pose in the ascending list S_k, found by bisection)
//   separator x separator                -> Sigma (mS x mS, + shift on the diagonal)
struct ElimView {
  const int *group, *lpos;       // per pose: domain (-1 = separator), position inside it
  const int *dom_m;              // scalars per domain
  const long long *a_off, *b_off;
  const int *sk_ptr, *sk;        // S_k as separator positions, CSR over the domains
  double *A, *B, *Sg;
  int mS;
};

__global__ void k_dd_scatter_all(const int *browidx, const int *colidx, const double *blocks, int nnzb, int dh,
                                 double shift, ElimView v) {
  const size_t total = (size_t)nnzb * dh * dh;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / (dh * dh));
    const int ab = (int)(t % (dh * dh));
    const int a = ab / dh, b = ab % dh;
    const int i = browidx[e], j = colidx[e];
    const int gi = v.group[i], gj = v.group[j];
    const double val = blocks[t];
    if (gi < 0) {
      if (gj >= 0) continue;                       // separator x interior: the transpose of a B_k entry
      const size_t row = (size_t)v.lpos[i] * dh + a, col = (size_t)v.lpos[j] * dh + b;
      v.Sg[row + col * (size_t)v.mS] = val + (row == col ? shift : 0.0);
    } else if (gj == gi) {
      const size_t m = (size_t)v.dom_m[gi];
      const size_t row = (size_t)v.lpos[i] * dh + a, col = (size_t)v.lpos[j] * dh + b;
      v.A[v.a_off[gi] + row + col * m] = val + (row == col ? shift : 0.0);
    } else if (gj < 0) {
      int lo = v.sk_ptr[gi], hi = v.sk_ptr[gi + 1] - 1;
      const int want = v.lpos[j];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (v.sk[mid] < want) lo = mid + 1; else hi = mid;
      }
      const size_t m = (size_t)v.dom_m[gi];
      const size_t row = (size_t)v.lpos[i] * dh + a, col = (size_t)(lo - v.sk_ptr[gi]) * dh + b;
      v.B[v.b_off[gi] + row + col * m] = val;
    }
  }
}

namespace {

// Device results of the elimination of every interior domain (released by the destructor)
struct Elimination {
  int K = 0;
  std::vector<int> m, tm;                    // scalars per domain, scalars of S_k
  std::vector<long long> a_off, b_off, w_off;
  std::vector<int> cmap_off;                 // offsets of the per-domain separator scalar lists in d_cmap
  double *Ainv = nullptr;                    // A_k^-1, full symmetric m_k x m_k
  double *B = nullptr, *C = nullptr;         // A_kS[:, S_k] and C_k = A_k^-1 A_kS[:, S_k]  (m_k x tm_k)
  double *W = nullptr;                       // B_k^T C_k (tm_k x tm_k)
  int *d_cmap = nullptr;
  std::vector<void *> aux;
  cudaStream_t stream = nullptr;             // everything here is stream-ordered scratch
  ~Elimination() {
    void *p[] = {Ainv, B, C, W, d_cmap};
    for (void *q : p) if (q) cudaFreeAsync(q, stream);
    for (void *q : aux) if (q) cudaFreeAsync(q, stream);
  }
};

template <typename T>
int upload_scratch(T **dptr, const std::vector<T> &v, cudaStream_t st) {
  const size_t ne = std::max<size_t>(v.size(), 1);
  CUDA_TRY(cudaMallocAsync((void **)dptr, ne * sizeof(T), st));
  if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return DPGO_OK;
}

// A_k^-1 for every domain, C_k on the columns of S_k, and Sigma = A_SS - sum_k B_k^T C_k accumulated into Sg
// (mS x mS, zeroed here).  sk lists separator POSITIONS (ascending per domain).
int eliminate_domains(dpgo_dev *h, const std::vector<int> &group, const std::vector<int> &lpos,
                      const std::vector<int> &dom_m, const std::vector<int> &sk_ptr, const std::vector<int> &sk,
                      double *Sg, int mS, Elimination &E) {
  const int K = (int)dom_m.size(), dh = h->d + 1;
  E.K = K;
  E.stream = h->stream;
  E.m = dom_m;
  E.tm.resize(K); E.a_off.resize(K); E.b_off.resize(K); E.w_off.resize(K); E.cmap_off.assign(K + 1, 0);
  long long na = 0, nb = 0, nw = 0;
  std::vector<int> cmap;
  for (int k = 0; k < K; ++k) {
    E.tm[k] = (sk_ptr[k + 1] - sk_ptr[k]) * dh;
    E.a_off[k] = na; E.b_off[k] = nb; E.w_off[k] = nw;
    na += (long long)dom_m[k] * dom_m[k];
    nb += (long long)dom_m[k] * E.tm[k];
    nw += (long long)E.tm[k] * E.tm[k];
    for (int a = sk_ptr[k]; a < sk_ptr[k + 1]; ++a)
      for (int c = 0; c < dh; ++c) cmap.push_back(sk[a] * dh + c);
    E.cmap_off[k + 1] = (int)cmap.size();
  }
  auto dalloc = [&](double **p, long long count) -> int {
    const size_t bytes = (size_t)std::max<long long>(count, 1) * sizeof(double);
    CUDA_TRY(cudaMallocAsync((void **)p, bytes, h->stream));
    CUDA_TRY(cudaMemsetAsync(*p, 0, bytes, h->stream));
    return DPGO_OK;
  };
  DPGO_TRY(dalloc(&E.Ainv, na));
  DPGO_TRY(dalloc(&E.B, nb));
  DPGO_TRY(dalloc(&E.C, nb));
  DPGO_TRY(dalloc(&E.W, nw));
  DPGO_TRY(upload_scratch(&E.d_cmap, cmap, h->stream));
  int *d_group = nullptr, *d_lpos = nullptr, *d_m = nullptr, *d_skptr = nullptr, *d_sk = nullptr;
  long long *d_aoff = nullptr, *d_boff = nullptr;
  DPGO_TRY(upload_scratch(&d_group, group, h->stream)); E.aux.push_back(d_group);
  DPGO_TRY(upload_scratch(&d_lpos, lpos, h->stream)); E.aux.push_back(d_lpos);
  DPGO_TRY(upload_scratch(&d_m, dom_m, h->stream)); E.aux.push_back(d_m);
  DPGO_TRY(upload_scratch(&d_skptr, sk_ptr, h->stream)); E.aux.push_back(d_skptr);
  DPGO_TRY(upload_scratch(&d_sk, sk, h->stream)); E.aux.push_back(d_sk);
  DPGO_TRY(upload_scratch(&d_aoff, E.a_off, h->stream)); E.aux.push_back(d_aoff);
  DPGO_TRY(upload_scratch(&d_boff, E.b_off, h->stream)); E.aux.push_back(d_boff);
  if (mS > 0) CUDA_TRY(cudaMemsetAsync(Sg, 0, (size_t)mS * mS * sizeof(double), h->stream));
  {
    const size_t total = (size_t)h->nnzb * dh * dh;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8));
    const ElimView v{d_group, d_lpos, d_m, d_aoff, d_boff, d_skptr, d_sk, E.Ainv, E.B, Sg, mS};
    k_dd_scatter_all<<<grid, 256, 0, h->stream>>>(h->d_browidx, h->d_colidx, h->d_blocks, h->nnzb, dh,
                                                  0.1 /* ref: src/PoseGraph.cpp:603 */, v);
    CUDA_TRY(cudaPeekAtLastError());
  }
  // ---- A_k^-1, all domains in lockstep
  {
    std::vector<dla::SpdItem> items(K);
    for (int k = 0; k < K; ++k) items[k] = dla::SpdItem{E.Ainv + E.a_off[k], dom_m[k], std::max(dom_m[k], 1)};
    const int rc = dla::spd_inverse_batched(h->stream, items.data(), K, /*symmetrize=*/true);
    if (rc > 0) {
      set_error("interior block %d of Q + 0.1 I is not positive definite", rc - 1);
      return DPGO_ENUMERIC;
    }
    if (rc < 0) {
      set_error("%s", dla::last_error());
      return DPGO_ECUDA;
    }
  }
  // ---- C_k = A_k^-1 B_k and the Schur updates W_k = B_k^T C_k
  std::vector<dla::GemmDesc> gc, gw;
  for (int k = 0; k < K; ++k) {
    if (E.tm[k] <= 0) continue;
    const int m = dom_m[k], tm = E.tm[k];
    gc.push_back(dla::GemmDesc{E.Ainv + E.a_off[k], E.B + E.b_off[k], E.C + E.b_off[k], m, tm, m, m, m, m});
    gw.push_back(dla::GemmDesc{E.B + E.b_off[k], E.C + E.b_off[k], E.W + E.w_off[k], tm, tm, m, m, m, tm});
  }
  if (!gc.empty()) {
    if (dla::gemm_batched(h->stream, gc.data(), (int)gc.size(), dla::GemmFlags{0, 0, 0, dla::K_FULL, 1.0, 0.0}) != 0 ||
        dla::gemm_batched(h->stream, gw.data(), (int)gw.size(), dla::GemmFlags{1, 0, 0, dla::K_FULL, 1.0, 0.0}) != 0) {
      set_error("%s", dla::last_error());
      return DPGO_ECUDA;
    }
  }
  // domains share separator poses: the updates of Sigma are applied one domain after the other (fixed order)
  for (int k = 0; k < K; ++k) {
    const int tm = E.tm[k];
    if (tm <= 0) continue;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(((size_t)tm * tm + 255) / 256, (size_t)h->num_sms * 8));
    k_dd_scatter_sub<<<grid, 256, 0, h->stream>>>(Sg, mS, E.d_cmap + E.cmap_off[k], E.W + E.w_off[k], tm);
  }
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

int invert_schur(dpgo_dev *h, double *Sg, int mS) {
  const dla::SpdItem item{Sg, mS, mS};
  const int rc = dla::spd_inverse_batched(h->stream, &item, 1, false);
  if (rc > 0) {
    set_error("Schur complement of Q + 0.1 I is not positive definite");
    return DPGO_ENUMERIC;
  }
  if (rc < 0) {
    set_error("%s", dla::last_error());
    return DPGO_ECUDA;
  }
  return DPGO_OK;
}

}  // namespace

// ---- application kernels -----------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kBlock, 1) k_strip_gemv(DdStripSet S, int V, const double *vec, const int *icol,
                                                          double *out, size_t outstride) {
  extern __shared__ __align__(128) unsigned char dsm[];
  GemvPipe pp = gemv_pipe_init<kDdStages, kDdVecChunks>(dsm);
  __shared__ StripPlanStore plan;
  strip_plan_fill(&plan, S, V);
  phase_strip_gemv<R, kDdStages>(pp, S, V, &plan, vec, icol, out, outstride);
}
template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_dd_sep_rhs(DdView dd, const double *rvec) {
  phase_dd_sep_rhs<R, D>(make_ctx(), dd, rvec);
}
template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_dd_back_rhs(DdView dd) {
  phase_dd_back_rhs<R, D>(make_ctx(), dd);
}
template <int R, int D>
__global__ void __launch_bounds__(kBlock) k_dd_finish(DdView dd, const double *Y, const double *rvec, double *z,
                                                      double *neg_out, int n, double *partials) {
  double acc[1] = {0.0};
  phase_dd_finish<R, D>(make_ctx(), dd, Y, rvec, z, neg_out, n, acc);
  block_reduce_store<1>(acc, partials + blockIdx.x);
}

#define DD_DISPATCH(h, ...)                                                              \
  switch ((h)->d * 16 + (h)->r) {                                                        \
    case 2 * 16 + 2: { constexpr int D = 2, R = 2; __VA_ARGS__; } break;                 \
    case 2 * 16 + 3: { constexpr int D = 2, R = 3; __VA_ARGS__; } break;                 \
    case 2 * 16 + 4: { constexpr int D = 2, R = 4; __VA_ARGS__; } break;                 \
    case 2 * 16 + 5: { constexpr int D = 2, R = 5; __VA_ARGS__; } break;                 \
    case 3 * 16 + 3: { constexpr int D = 3, R = 3; __VA_ARGS__; } break;                 \
    case 3 * 16 + 4: { constexpr int D = 3, R = 4; __VA_ARGS__; } break;                 \
    case 3 * 16 + 5: { constexpr int D = 3, R = 5; __VA_ARGS__; } break;                 \
    case 3 * 16 + 6: { constexpr int D = 3, R = 6; __VA_ARGS__; } break;                 \
    default: dpgo::set_error("unsupported (d=%d, r=%d)", (h)->d, (h)->r); return DPGO_EINVAL; \
  }

DdView dd_view(const dpgo_dev *h) {
  const DdState *s = (const DdState *)h->dd;
  DdView v;
  v.P1 = DdStripSet{s->M1, s->strips1, s->cta1, s->chunks1};
  v.P3 = DdStripSet{s->M3, s->strips3, s->cta3, s->chunks3};
  v.V = s->V;
  v.nsplit1 = s->nsplit1; v.nsplit3 = s->nsplit3;
  v.A_SI = BsrView{s->si_rowptr, s->si_colidx, s->si_blocks};
  v.A_BS = BsrView{s->bs_rowptr, s->bs_colidx, s->bs_blocks};
  v.nS = s->nS; v.nB = s->nB;
  v.pcol = s->pcol; v.srow = s->srow; v.bcol = s->bcol; v.icol = s->icol;
  v.sep_col0 = s->sep_col0; v.pcols = s->pcols;
  v.y = s->y; v.t = s->t; v.zs = s->zs; v.u = s->u; v.w = s->w; v.rp = s->rp;
  v.prefetch = h->dd_prefetch;
  return v;
}

void dd_free(dpgo_dev *h) {
  DdState *s = (DdState *)h->dd;
  if (!s) return;
  void *ptrs[] = {s->M1, s->M3, s->strips1, s->strips3, s->cta1, s->cta3, s->chunks1, s->chunks3, s->pcol, s->srow,
                  s->bcol, s->icol, s->si_rowptr, s->si_colidx, s->bs_rowptr, s->bs_colidx, s->si_blocks, s->bs_blocks,
                  s->y, s->t, s->zs, s->u, s->w, s->rp, s->d_si_src, s->d_bs_src};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete s;
  h->dd = nullptr;
}

// a strip of an interior block is one wave of kDdStages chunks of kStageK inner indices
int two_level_max_domain_poses(int dh) { return std::max(8, kDdStages * kStageK / dh); }

double dd_bytes(const dpgo_dev *h) { return h->dd ? ((const DdState *)h->dd)->bytes_per_apply : 0.0; }

static int dd_numeric(dpgo_dev *h, DdState *s, bool refresh_couplings);

int dd_build(dpgo_dev *h) {
  const int n = h->n, dh = h->d + 1, R = h->r;
  // interior blocks of <= 320 scalars (5 column blocks of 64): a strip is one wave of kDdStages chunks
  const int thr = h->dd_max_domain > 0 ? h->dd_max_domain : two_level_max_domain_poses(dh);
  {
    // weight-only update (dpgo_update_weights: same pattern, new values): numeric part only
    DdState *old = (DdState *)h->dd;
    if (old && old->symbolic && h->weights_only_update && old->sym_n == n && old->sym_nnzb == h->nnzb &&
        old->sym_thr == thr && old->sym_split1 == h->dd_split1 && old->sym_split3 == h->dd_split3)
      return dd_numeric(h, old, true);
  }
  DPGO_TRY(sync_host_blocks(h));   // the couplings below are read from the host copy of Q
  dd_free(h);
  DdState *s = new DdState();
  h->dd = s;
  // ---- partition
  const std::vector<std::vector<int>> adj = bsr_adjacency(n, h->rowptr.data(), h->colidx.data());
  Dissector ds(adj, thr);
  {
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    ds.run(std::move(all));
  }
  std::sort(ds.sep.begin(), ds.sep.end());
  const int K = (int)ds.domains.size();
  s->K = K;
  s->nS = (int)ds.sep.size();
  // ---- permuted, padded column space
  std::vector<int> group(n, -1), lpos(n, 0), pcol(n, 0), irow, srow(ds.sep);
  std::vector<int> dom_off(K), dom_pad(K), dom_m(K);
  int col = 0;
  for (int k = 0; k < K; ++k) {
    auto &dom = ds.domains[k];
    std::sort(dom.begin(), dom.end());
    dom_m[k] = (int)dom.size() * dh;
    dom_pad[k] = ((dom_m[k] + kGemvCols - 1) / kGemvCols) * kGemvCols;
    dom_off[k] = col;
    for (size_t j = 0; j < dom.size(); ++j) {
      group[dom[j]] = k;
      lpos[dom[j]] = (int)j;
      pcol[dom[j]] = col + (int)j * dh;
      irow.push_back(dom[j]);
    }
    col += dom_pad[k];
  }
  s->nI = (int)irow.size();
  s->sep_col0 = col;
  const int mS = s->nS * dh;
  const int padS = ((mS + kGemvCols - 1) / kGemvCols) * kGemvCols;
  for (int j = 0; j < s->nS; ++j) {
    lpos[srow[j]] = j;
    pcol[srow[j]] = col + j * dh;
  }
  col += padS;
  s->pcols = std::max(col, kGemvCols);
  // ---- sparse couplings (blocks of Q between separator and interior poses)
  const int bs = dh * dh;
  std::vector<int> si_rowptr(s->nS + 1, 0), si_colidx, bs_rowptr(1, 0), bs_colidx, bcol;
  std::vector<double> si_blocks, bs_blocks;
  for (int j = 0; j < s->nS; ++j) {
    const int i = srow[j];
    for (int e = h->rowptr[i]; e < h->rowptr[i + 1]; ++e) {
      const int c = h->colidx[e];
      if (group[c] < 0) continue;
      si_colidx.push_back(pcol[c]);
      s->si_src.push_back(e);
      si_blocks.insert(si_blocks.end(), h->blocks.begin() + (size_t)e * bs, h->blocks.begin() + (size_t)(e + 1) * bs);
    }
    si_rowptr[j + 1] = (int)si_colidx.size();
  }
  // interior poses with at least one separator neighbour ("boundary" rows of A_IS)
  for (int a = 0; a < s->nI; ++a) {
    const int i = irow[a];
    const size_t before = bs_colidx.size();
    for (int e = h->rowptr[i]; e < h->rowptr[i + 1]; ++e) {
      const int c = h->colidx[e];
      if (group[c] >= 0) continue;
      bs_colidx.push_back(pcol[c]);
      s->bs_src.push_back(e);
      bs_blocks.insert(bs_blocks.end(), h->blocks.begin() + (size_t)e * bs, h->blocks.begin() + (size_t)(e + 1) * bs);
    }
    if (bs_colidx.size() > before) {
      bcol.push_back(pcol[i]);
      bs_rowptr.push_back((int)bs_colidx.size());
    }
  }
  s->nB = (int)bcol.size();
  {
    std::vector<int> icol(s->pcols, -1);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < dh; ++c) icol[pcol[i] + c] = i * dh + c;
    DPGO_TRY(upload_vec(&s->icol, icol));
  }
  DPGO_TRY(upload_vec(&s->pcol, pcol));
  DPGO_TRY(upload_vec(&s->srow, srow));
  DPGO_TRY(upload_vec(&s->bcol, bcol));
  DPGO_TRY(upload_vec(&s->si_rowptr, si_rowptr));
  DPGO_TRY(upload_vec(&s->si_colidx, si_colidx));
  DPGO_TRY(upload_vec(&s->si_blocks, si_blocks));
  DPGO_TRY(upload_vec(&s->bs_rowptr, bs_rowptr));
  DPGO_TRY(upload_vec(&s->bs_colidx, bs_colidx));
  DPGO_TRY(upload_vec(&s->bs_blocks, bs_blocks));
  DPGO_TRY(upload_vec(&s->d_si_src, s->si_src));
  DPGO_TRY(upload_vec(&s->d_bs_src, s->bs_src));
  // ---- strip tables.  A strip = 64 output columns x a run of 32-index chunks.  The apply is
  // latency bound (the matrices are L2 resident): one CTA per SM, each with about one pipeline
  // fill of work.  Interior strips are whole (one partial slot) and balanced over the V virtual
  // CTAs longest-first; the Schur block is split along the inner dimension into nsplit3 partial
  // slots so that its strips fill the V CTAs once.
  const int V = std::max(1, h->num_sms);
  s->V = V;
  auto balance = [&](std::vector<DdStrip> &strips, std::vector<int> &cta, std::vector<int> &chunks) {
    std::vector<int> order(strips.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return strips[a].nchunks > strips[b].nchunks; });
    std::vector<std::vector<int>> bins(V);
    chunks.assign(V, 0);
    for (int i : order) {
      int best = 0;
      for (int v = 1; v < V; ++v)
        if (chunks[v] < chunks[best]) best = v;
      bins[best].push_back(i);
      chunks[best] += strips[i].nchunks;
    }
    std::vector<DdStrip> sorted;
    cta.assign(V + 1, 0);
    for (int v = 0; v < V; ++v) {
      for (int i : bins[v]) sorted.push_back(strips[i]);
      cta[v + 1] = (int)sorted.size();
    }
    strips.swap(sorted);
  };
  std::vector<DdStrip> strips1, strips3;
  long long stage = 0;
  std::vector<long long> dom_base(K);
  double bytes = 0;
  int nsplit1 = std::max(1, h->dd_split1);
  int used1 = 1;
  for (int k = 0; k < K; ++k) {
    dom_base[k] = stage;
    const int nch = dom_pad[k] / kStageK, ncb = dom_pad[k] / kGemvCols;
    const int e = std::max(1, std::min(nsplit1, nch));
    const int cps = (nch + e - 1) / e;
    for (int cb = 0; cb < ncb; ++cb)
      for (int sp = 0, c0 = 0; c0 < nch; ++sp, c0 += cps) {
        strips1.push_back(DdStrip{dom_off[k] / kGemvCols + cb, dom_off[k] / kStageK + c0, std::min(cps, nch - c0), sp,
                                  stage + (long long)cb * nch + c0});
        used1 = std::max(used1, sp + 1);
      }
    stage += (long long)ncb * nch;
    bytes += 2.0 * (double)dom_m[k] * dom_m[k] * 8;
  }
  nsplit1 = used1;
  const long long stages1 = stage;
  const int nchS = padS / kStageK, ncbS = padS / kGemvCols;
  int nsplit3 = h->dd_split3;
  if (nsplit3 <= 0) {
    // waves of V CTAs x time of a strip, plus the cost of one more partial array for the consumers
    nsplit3 = 1;
    double best_cost = 1e300;
    for (int ns = 1; ns <= std::min(std::max(nchS, 1), 32); ++ns) {
      const int cps = (nchS + ns - 1) / ns;
      const long strips = (long)ncbS * ((nchS + cps - 1) / std::max(cps, 1));
      // per strip: fixed latency + rounds of 8 warps over (chunk, 8-row) units + extra waves
      const double per_strip = 4.0 + 2.0 * ((cps + 1) / 2) / 4.0 + 3.0 * ((cps + kDdStages - 1) / kDdStages - 1);
      const double cost = (double)((strips + V - 1) / V) * per_strip + 0.3 * ns;
      if (cost < best_cost - 1e-9) { best_cost = cost; nsplit3 = ns; }
    }
  }
  nsplit3 = std::max(1, std::min(nsplit3, std::max(nchS, 1)));
  const int cps = nchS > 0 ? (nchS + nsplit3 - 1) / nsplit3 : 1;
  nsplit3 = nchS > 0 ? (nchS + cps - 1) / cps : 1;
  for (int cb = 0; cb < ncbS; ++cb)
    for (int sp = 0; sp < nsplit3; ++sp) {
      const int c0 = sp * cps, nc = std::min(cps, nchS - c0);
      if (nc <= 0) continue;
      strips3.push_back(DdStrip{s->sep_col0 / kGemvCols + cb, s->sep_col0 / kStageK + c0, nc, sp,
                                (long long)cb * nchS + c0});
    }
  bytes += (double)mS * mS * 8;
  s->bytes_per_apply = bytes + 6.0 * R * h->N * 8;
  s->nsplit1 = nsplit1;
  s->nsplit3 = nsplit3;
  s->nstrips1 = (int)strips1.size();
  s->nstrips3 = (int)strips3.size();
  {
    std::vector<int> cta, chunks;
    balance(strips1, cta, chunks);
    DPGO_TRY(upload_vec(&s->strips1, strips1));
    DPGO_TRY(upload_vec(&s->cta1, cta));
    DPGO_TRY(upload_vec(&s->chunks1, chunks));
    balance(strips3, cta, chunks);
    DPGO_TRY(upload_vec(&s->strips3, strips3));
    DPGO_TRY(upload_vec(&s->cta3, cta));
    DPGO_TRY(upload_vec(&s->chunks3, chunks));
  }
  // ---- work arrays
  const size_t wlen = (size_t)R * s->pcols;
  double **arrs[] = {&s->y, &s->t, &s->u, &s->w};
  for (double **a : arrs) {
    const size_t mult = (a == &s->y || a == &s->w) ? (size_t)nsplit1 : 1;   // partial slots
    CUDA_TRY(cudaMalloc((void **)a, wlen * mult * sizeof(double)));
    CUDA_TRY(cudaMemset(*a, 0, wlen * mult * sizeof(double)));
  }
  CUDA_TRY(cudaMalloc((void **)&s->zs, wlen * nsplit3 * sizeof(double)));
  CUDA_TRY(cudaMemset(s->zs, 0, wlen * nsplit3 * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&s->rp, wlen * sizeof(double)));
  CUDA_TRY(cudaMemset(s->rp, 0, wlen * sizeof(double)));
  // ---- dense blocks on the device
  CUDA_TRY(cudaMalloc((void **)&s->M1, std::max<size_t>((size_t)stages1 * kStageDoubles, 1) * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&s->M3, std::max<size_t>((size_t)padS * padS, 1) * sizeof(double)));
  // S_k: separator positions with a neighbour in domain k (ascending); the coupling A_kS is non-zero only there
  std::vector<int> sk_ptr(K + 1, 0), sk;
  {
    std::vector<std::vector<int>> Sk(K);
    for (int j = 0; j < s->nS; ++j) {
      const int i = srow[j];
      for (int e = h->rowptr[i]; e < h->rowptr[i + 1]; ++e) {
        const int g = group[h->colidx[e]];
        if (g >= 0 && (Sk[g].empty() || Sk[g].back() != j)) Sk[g].push_back(j);
      }
    }
    for (int k = 0; k < K; ++k) {
      sk.insert(sk.end(), Sk[k].begin(), Sk[k].end());
      sk_ptr[k + 1] = (int)sk.size();
    }
  }
  s->h_group = group; s->h_lpos = lpos; s->h_dom_m = dom_m; s->h_dom_pad = dom_pad; s->h_dom_base = dom_base;
  s->h_sk_ptr = sk_ptr; s->h_sk = sk;
  s->mS = mS; s->padS = padS;
  s->sym_n = n; s->sym_nnzb = h->nnzb; s->sym_thr = thr; s->sym_split1 = h->dd_split1; s->sym_split3 = h->dd_split3;
  s->symbolic = true;
  return dd_numeric(h, s, false);
}

// coupling block k <- block src[k] of Q
__global__ void k_dd_regather(const double *blocks, const int *src, int count, int bs, double *dst) {
  const size_t total = (size_t)count * bs;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t k = idx / bs;
    dst[idx] = blocks[(size_t)src[k] * bs + (idx - k * bs)];
  }
}

// Numeric part of the two-level set-up: dense interior inverses, couplings on S_k, Schur complement and its inverse,
// strip layouts -- from the current values of Q.  refresh_couplings: also re-read the sparse coupling blocks
// (weight-only rebuild; the index structure on the device stays).
static int dd_numeric(dpgo_dev *h, DdState *s, bool refresh_couplings) {
  const int dh = h->d + 1, bs = dh * dh, K = s->K, mS = s->mS, padS = s->padS;
  if (refresh_couplings) {    // device copy of the re-weighted blocks of Q (the host copy may be stale)
    auto regather = [&](const std::vector<int> &src, const int *d_src, double *dst) {
      if (src.empty()) return;
      const size_t total = src.size() * (size_t)bs;
      const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
      k_dd_regather<<<std::max(grid, 1), 256, 0, h->stream>>>(h->d_blocks, d_src, (int)src.size(), bs, dst);
      h->launches++;
    };
    regather(s->si_src, s->d_si_src, s->si_blocks);
    regather(s->bs_src, s->d_bs_src, s->bs_blocks);
    CUDA_TRY(cudaPeekAtLastError());
  }
  double *Sg = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&Sg, (size_t)std::max(mS, 1) * std::max(mS, 1) * sizeof(double), h->stream));
  {
    Elimination E;
    int rc = eliminate_domains(h, s->h_group, s->h_lpos, s->h_dom_m, s->h_sk_ptr, s->h_sk, Sg, mS, E);
    for (int k = 0; k < K && rc == DPGO_OK; ++k) {
      const size_t total = (size_t)s->h_dom_pad[k] * s->h_dom_pad[k];
      const int lgrid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
      k_dd_layout<<<std::max(lgrid, 1), 256, 0, h->stream>>>(E.Ainv + E.a_off[k], s->h_dom_m[k], s->h_dom_m[k], s->h_dom_pad[k],
                                                           s->M1 + (size_t)s->h_dom_base[k] * kStageDoubles);
    }
    if (rc == DPGO_OK && mS > 0) rc = invert_schur(h, Sg, mS);
    if (rc == DPGO_OK && mS > 0) {
      const size_t total = (size_t)padS * padS;
      const int lgrid = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 8);
      k_dd_layout<<<std::max(lgrid, 1), 256, 0, h->stream>>>(Sg, mS, mS, padS, s->M3);
    }
    cudaFreeAsync(Sg, h->stream);
    const cudaError_t e1 = cudaPeekAtLastError(), e2 = cudaStreamSynchronize(h->stream);
    DPGO_TRY(rc);
    CUDA_TRY(e1);
    CUDA_TRY(e2);
  }
  return DPGO_OK;
}

template <int R>
static int strip_setup() {
  return cudaFuncSetAttribute(k_strip_gemv<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDdDynSmem) == cudaSuccess
             ? 0 : -1;
}

static int launch_strips(dpgo_dev *h, const DdStripSet &S, int nstrips, const double *vec, const int *icol,
                         double *out, size_t outstride) {
  DdState *s = (DdState *)h->dd;
  if (nstrips <= 0) return DPGO_OK;
  if (!s->configured) {
    int rc = -1;
    switch (h->r) {
      case 2: rc = strip_setup<2>(); break;
      case 3: rc = strip_setup<3>(); break;
      case 4: rc = strip_setup<4>(); break;
      case 5: rc = strip_setup<5>(); break;
      case 6: rc = strip_setup<6>(); break;
    }
    if (rc != 0) {
      set_error("two-level strip kernel does not fit on the device");
      return DPGO_ECUDA;
    }
    s->configured = true;
  }
  const int grid = s->V;
  switch (h->r) {
    case 2: k_strip_gemv<2><<<grid, kBlock, kDdDynSmem, h->stream>>>(S, s->V, vec, icol, out, outstride); break;
    case 3: k_strip_gemv<3><<<grid, kBlock, kDdDynSmem, h->stream>>>(S, s->V, vec, icol, out, outstride); break;
    case 4: k_strip_gemv<4><<<grid, kBlock, kDdDynSmem, h->stream>>>(S, s->V, vec, icol, out, outstride); break;
    case 5: k_strip_gemv<5><<<grid, kBlock, kDdDynSmem, h->stream>>>(S, s->V, vec, icol, out, outstride); break;
    case 6: k_strip_gemv<6><<<grid, kBlock, kDdDynSmem, h->stream>>>(S, s->V, vec, icol, out, outstride); break;
    default: set_error("unsupported r=%d", h->r); return DPGO_EINVAL;
  }
  h->launches++;
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

// grid for the warp-per-row sparse phases
static inline int warp_rows_grid(const dpgo_dev *h, int rows) {
  long blocks = ((long)rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  blocks = std::max(1L, std::min(blocks, (long)h->num_sms * 4));
  return (int)blocks;
}

static inline int rows_grid(const dpgo_dev *h, int rows) {
  const int gpw = 32 / (h->d + 1);
  const long warps = ((long)rows + gpw - 1) / gpw;
  long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  blocks = std::max(1L, std::min(blocks, (long)h->partial_blocks));
  return (int)blocks;
}

// the five streaming / sparse phases (everything but the final projection)
int dd_time_apply(dpgo_dev *h, const double *vec) {
  DdState *s = (DdState *)h->dd;
  const DdView dd = dd_view(h);
  const size_t zstride = (size_t)h->r * s->pcols;
  DPGO_TRY(launch_strips(h, dd.P1, s->nstrips1, vec, s->icol, s->y, zstride));   // gathers vec on the fly
  if (s->nS > 0) {
    DD_DISPATCH(h, k_dd_sep_rhs<R, D><<<warp_rows_grid(h, s->nS), kBlock, 0, h->stream>>>(dd, vec));
    h->launches++;
    DPGO_TRY(launch_strips(h, dd.P3, s->nstrips3, s->t, nullptr, s->zs, zstride));
    DD_DISPATCH(h, k_dd_back_rhs<R, D><<<warp_rows_grid(h, s->nB), kBlock, 0, h->stream>>>(dd));
    h->launches++;
    DPGO_TRY(launch_strips(h, dd.P1, s->nstrips1, s->u, nullptr, s->w, zstride));
  }
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

int op_precon_dd(dpgo_dev *h, const double *Y, const double *rvec, double *z, double *neg_out, double *z_r) {
  if (!h->dd) {
    set_error("two-level preconditioner not built");
    return DPGO_ESTATE;
  }
  DPGO_TRY(dd_time_apply(h, rvec));
  const DdView dd = dd_view(h);
  const int g2 = rows_grid(h, h->n);
  DD_DISPATCH(h, k_dd_finish<R, D><<<g2, kBlock, 0, h->stream>>>(dd, Y, rvec, z, neg_out, h->n, h->d_partials));
  h->launches++;
  CUDA_TRY(cudaPeekAtLastError());
  if (z_r) {
    double sc[1];
    DPGO_TRY(read_partials(h, g2, 1, sc));
    *z_r = sc[0];
  }
  return DPGO_OK;
}

}  // namespace dpgo
