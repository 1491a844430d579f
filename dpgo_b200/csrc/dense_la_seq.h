// Launch sequences of the in-tree dense linear algebra, written against a small back-end interface so
// that the same descriptor arithmetic runs on the device (dense_la.cu) and under the CPU emulation of the
// kernels (tests/native/dense_la_emu.cpp).  See dense_la.cuh for the algorithm.
//
// Backend concept:
//   void *alloc(size_t bytes);                 zero-initialised device (or host) memory, nullptr on failure
//   void release(void *p);
//   bool upload(void *dst, const void *src, size_t bytes);
//   bool download(void *dst, const void *src, size_t bytes);     (waits for queued work)
//   void diag(const SpdDesc *d, int panel, int count, int *info);
//   void gemm(const GemmDesc *g, int tiles_m, int tiles_n, int count, const GemmFlags &f);
//   void copy(const SpdDesc *d, int what, int tiles, int count);
#pragma once
#include <algorithm>
#include <vector>

#include "dense_la.cuh"

namespace dpgo {
namespace dla {

struct SeqItem {
  double *A;
  int n, lda;
};

struct Launch {
  int first, count;      // descriptor range
  int tiles_m, tiles_n;
  GemmFlags flags;
};

inline int tiles_of(int n) { return (n + TS - 1) / TS; }

template <class Backend>
int spd_inverse_seq(Backend &be, const SeqItem *items, int count, bool symmetrize) {
  if (count <= 0) return 0;
  int nmax = 0;
  size_t dinv_total = 0, tmp_total = 0;
  std::vector<size_t> dinv_off(count), tmp_off(count);
  for (int b = 0; b < count; ++b) {
    nmax = std::max(nmax, items[b].n);
    dinv_off[b] = dinv_total;
    tmp_off[b] = tmp_total;
    dinv_total += (size_t)tiles_of(items[b].n) * TS * TS;
    tmp_total += (size_t)items[b].n * items[b].n;
  }
  if (nmax <= 0) return 0;
  const int T = tiles_of(nmax);
  double *dinv = (double *)be.alloc(std::max<size_t>(dinv_total, 1) * sizeof(double));
  double *tmp = (double *)be.alloc(std::max<size_t>(tmp_total, 1) * sizeof(double));
  int *info = (int *)be.alloc(sizeof(int));
  SpdDesc *d_spd = (SpdDesc *)be.alloc((size_t)count * sizeof(SpdDesc));
  int rc = 0;
  GemmDesc *d_g = nullptr;
  do {
    if (!dinv || !tmp || !info || !d_spd) { rc = -1; break; }
    std::vector<SpdDesc> spd(count);
    for (int b = 0; b < count; ++b)
      spd[b] = SpdDesc{items[b].A, items[b].n, items[b].lda, dinv + dinv_off[b], tmp + tmp_off[b], std::max(items[b].n, 1)};
    if (!be.upload(d_spd, spd.data(), (size_t)count * sizeof(SpdDesc))) { rc = -1; break; }

    // ---- every GEMM of the whole inversion, described up front
    std::vector<GemmDesc> g;
    std::vector<Launch> trsm(T), syrk(T), lvl1, lvl2;
    for (int p = 0; p < T; ++p) {
      const int j0 = p * TS;
      Launch lt{(int)g.size(), 0, 0, 1, GemmFlags{0, 1, 0, K_FULL, 1.0, 0.0}};
      for (int b = 0; b < count; ++b) {
        const SpdDesc &s = spd[b];
        const int nb = std::min(TS, s.n - j0), n2 = s.n - j0 - nb;
        if (nb <= 0 || n2 <= 0) continue;
        double *A21 = s.A + (size_t)(j0 + nb) + (size_t)j0 * s.lda;
        // X = A21 * inv(L11)^T, in place (one column tile: a CTA reads only the rows it writes)
        g.push_back(GemmDesc{A21, s.dinv + (size_t)p * TS * TS, A21, n2, nb, nb, s.lda, TS, s.lda});
        lt.count++;
        lt.tiles_m = std::max(lt.tiles_m, tiles_of(n2));
      }
      trsm[p] = lt;
      Launch ls{(int)g.size(), 0, 0, 0, GemmFlags{0, 1, 1, K_FULL, -1.0, 1.0}};
      for (int b = 0; b < count; ++b) {
        const SpdDesc &s = spd[b];
        const int nb = std::min(TS, s.n - j0), n2 = s.n - j0 - nb;
        if (nb <= 0 || n2 <= 0) continue;
        double *A21 = s.A + (size_t)(j0 + nb) + (size_t)j0 * s.lda;
        double *A22 = s.A + (size_t)(j0 + nb) + (size_t)(j0 + nb) * s.lda;
        g.push_back(GemmDesc{A21, A21, A22, n2, n2, nb, s.lda, s.lda, s.lda});   // A22 -= X X^T (lower tiles)
        ls.count++;
        ls.tiles_m = std::max(ls.tiles_m, tiles_of(n2));
      }
      ls.tiles_n = ls.tiles_m;
      syrk[p] = ls;
    }
    for (int sz = TS; sz < nmax; sz *= 2) {
      Launch l1{(int)g.size(), 0, 0, sz / TS, GemmFlags{0, 0, 0, K_FROM_COL_TILE, 1.0, 0.0}};
      for (int b = 0; b < count; ++b) {
        const SpdDesc &s = spd[b];
        for (int base = 0; base + sz < s.n; base += 2 * sz) {
          const int s2 = std::min(sz, s.n - base - sz);
          const double *Bblk = s.A + (size_t)(base + sz) + (size_t)base * s.lda;
          const double *Ablk = s.A + (size_t)base + (size_t)base * s.lda;
          double *Tblk = s.tmp + (size_t)(base + sz) + (size_t)base * s.ldt;
          g.push_back(GemmDesc{Bblk, Ablk, Tblk, s2, sz, sz, s.lda, s.lda, s.ldt});          // T = B A^-1
          l1.count++;
          l1.tiles_m = std::max(l1.tiles_m, tiles_of(s2));
        }
      }
      Launch l2{(int)g.size(), 0, l1.tiles_m, sz / TS, GemmFlags{0, 0, 0, K_UPTO_ROW_TILE, -1.0, 0.0}};
      for (int b = 0; b < count; ++b) {
        const SpdDesc &s = spd[b];
        for (int base = 0; base + sz < s.n; base += 2 * sz) {
          const int s2 = std::min(sz, s.n - base - sz);
          double *Bblk = s.A + (size_t)(base + sz) + (size_t)base * s.lda;
          const double *Cblk = s.A + (size_t)(base + sz) + (size_t)(base + sz) * s.lda;
          const double *Tblk = s.tmp + (size_t)(base + sz) + (size_t)base * s.ldt;
          g.push_back(GemmDesc{Cblk, Tblk, Bblk, s2, sz, s2, s.lda, s.ldt, s.lda});          // B = -C^-1 T
          l2.count++;
        }
      }
      lvl1.push_back(l1);
      lvl2.push_back(l2);
    }
    Launch lau{(int)g.size(), count, T, T, GemmFlags{1, 0, 1, K_FROM_ROW_TILE, 1.0, 0.0}};
    for (int b = 0; b < count; ++b) {
      const SpdDesc &s = spd[b];
      g.push_back(GemmDesc{s.A, s.A, s.tmp, s.n, s.n, s.n, s.lda, s.lda, s.ldt});            // W^T W (lower tiles)
    }
    d_g = (GemmDesc *)be.alloc(g.size() * sizeof(GemmDesc));
    if (!d_g || !be.upload(d_g, g.data(), g.size() * sizeof(GemmDesc))) { rc = -1; break; }
    auto run = [&](const Launch &l) {
      if (l.count > 0 && l.tiles_m > 0 && l.tiles_n > 0) be.gemm(d_g + l.first, l.tiles_m, l.tiles_n, l.count, l.flags);
    };
    // ---- 1. Cholesky
    for (int p = 0; p < T; ++p) {
      be.diag(d_spd, p, count, info);
      run(trsm[p]);
      run(syrk[p]);
    }
    // ---- 2. W = L^-1
    be.copy(d_spd, 0, T, count);
    for (size_t l = 0; l < lvl1.size(); ++l) {
      run(lvl1[l]);
      run(lvl2[l]);
    }
    // ---- 3. A^-1 = W^T W
    run(lau);
    be.copy(d_spd, 1, T, count);
    if (symmetrize) be.copy(d_spd, 2, T, count);
    int hinfo = 0;
    if (!be.download(&hinfo, info, sizeof(int))) { rc = -1; break; }
    rc = hinfo;
  } while (false);
  be.release(d_g);
  be.release(d_spd);
  be.release(info);
  be.release(tmp);
  be.release(dinv);
  return rc;
}

}  // namespace dla
}  // namespace dpgo
