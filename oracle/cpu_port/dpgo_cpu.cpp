// CPU port of the reference's local-solve hot path -- TEST / BASELINE INFRASTRUCTURE (see
// oracle/__init__.py; the product never links this).  It restates, in plain C++ (-O3
// -march=native, single thread like the reference's per-agent path), the same algorithm the
// numpy oracle (oracle/pgo.py) restates:
//   * Q X as a block-CSR product                        (ref: src/QuadraticProblem.cpp:29-54)
//   * preconditioner = exact solve with Q + 0.1 I        (ref: src/PoseGraph.cpp:598-613, CHOLMOD
//     there; here a block sparse Cholesky with a minimum-degree ordering)
//   * Stiefel tangent projection / QF retraction          (ROPTLIB semantics, restated)
//   * RTR + Steihaug-Toint tCG restated here from oracle/pgo.py (tcg / rtr_run); no product header is included:
//     the port is test infrastructure and is validated against the numpy oracle (tests/test_cpu_port.py)
//     (ref: src/QuadraticOptimizer.cpp:26-108)
// It exists so that the "reference CPU" column of bench.py is a compiled implementation and not
// a Python one.  Validated against oracle/pgo.py by tests/test_cpu_port.py.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <set>
#include <vector>


namespace {

struct Problem {
  int n, d, r, dh, N, tile;
  std::vector<int> rowptr, colidx;
  std::vector<double> blocks;  // row-major dh x dh per block
  std::vector<double> G;       // r x N col-major
  // block sparse Cholesky of Q + 0.1 I in elimination order
  std::vector<int> perm, iperm;                 // perm[k] = pose eliminated k-th
  std::vector<std::vector<int>> Lrows;          // per column k: rows (elimination positions > k), ascending
  std::vector<std::vector<double>> Lblk;        // per column k: blocks L_ik (row-major dh x dh), same order
  std::vector<double> Ldiag;                    // per column k: lower Cholesky factor of the pivot block
  bool factored = false;
  long qx = 0, precon = 0;
};

// ---- Q X ----------------------------------------------------------------------------------------
void qx(Problem &P, const double *X, double *out) {
  const int r = P.r, dh = P.dh, tile = P.tile;
  P.qx++;
  for (int i = 0; i < P.n; ++i) {
    double acc[8 * 4] = {0};
    for (int e = P.rowptr[i]; e < P.rowptr[i + 1]; ++e) {
      const double *B = &P.blocks[(size_t)e * dh * dh];  // Q_ij; out_i += X_j Q_ij^T
      const double *xj = X + (size_t)P.colidx[e] * tile;
      for (int c = 0; c < dh; ++c)
        for (int k = 0; k < dh; ++k) {
          const double b = B[c * dh + k];
          for (int q = 0; q < r; ++q) acc[c * r + q] += xj[k * r + q] * b;
        }
    }
    memcpy(out + (size_t)i * tile, acc, sizeof(double) * tile);
  }
}

// ---- ordering + factorization ---------------------------------------------------------------------
void minimum_degree(const Problem &P, std::vector<int> &perm) {
  const int n = P.n;
  std::vector<std::set<int>> adj(n);
  for (int i = 0; i < n; ++i)
    for (int e = P.rowptr[i]; e < P.rowptr[i + 1]; ++e)
      if (P.colidx[e] != i) adj[i].insert(P.colidx[e]);
  std::vector<char> done(n, 0);
  std::set<std::pair<int, int>> heap;  // (degree, node)
  for (int i = 0; i < n; ++i) heap.insert({(int)adj[i].size(), i});
  perm.clear();
  while (!heap.empty()) {
    const int v = heap.begin()->second;
    heap.erase(heap.begin());
    done[v] = 1;
    perm.push_back(v);
    std::vector<int> nb(adj[v].begin(), adj[v].end());
    for (int a : nb) {
      heap.erase({(int)adj[a].size(), a});
      adj[a].erase(v);
    }
    for (size_t x = 0; x < nb.size(); ++x)
      for (size_t y = x + 1; y < nb.size(); ++y) {
        adj[nb[x]].insert(nb[y]);
        adj[nb[y]].insert(nb[x]);
      }
    for (int a : nb) heap.insert({(int)adj[a].size(), a});
    adj[v].clear();
  }
}

bool chol_small(double *A, int m) {  // in-place lower Cholesky of a row-major m x m SPD block
  for (int j = 0; j < m; ++j) {
    double s = A[j * m + j];
    for (int k = 0; k < j; ++k) s -= A[j * m + k] * A[j * m + k];
    if (s <= 0) return false;
    const double l = sqrt(s);
    A[j * m + j] = l;
    for (int i = j + 1; i < m; ++i) {
      double t = A[i * m + j];
      for (int k = 0; k < j; ++k) t -= A[i * m + k] * A[j * m + k];
      A[i * m + j] = t / l;
    }
    for (int k = j + 1; k < m; ++k) A[j * m + k] = 0.0;
  }
  return true;
}

bool factorize(Problem &P, double shift) {
  const int n = P.n, dh = P.dh, bs = dh * dh;
  minimum_degree(P, P.perm);
  P.iperm.assign(n, 0);
  for (int k = 0; k < n; ++k) P.iperm[P.perm[k]] = k;
  // symbolic: structure of column k = {permuted neighbours > k} union fill, by the usual
  // "parent merge" of elimination: struct(k) = adj_upper(k) U (struct(c) \ {k}) for children c
  std::vector<std::set<int>> st(n);
  for (int i = 0; i < n; ++i)
    for (int e = P.rowptr[i]; e < P.rowptr[i + 1]; ++e) {
      const int a = P.iperm[i], b = P.iperm[P.colidx[e]];
      if (b > a) st[a].insert(b);
    }
  for (int k = 0; k < n; ++k) {
    if (st[k].empty()) continue;
    const int parent = *st[k].begin();
    for (int x : st[k])
      if (x != parent) st[parent].insert(x);
  }
  P.Lrows.assign(n, {});
  P.Lblk.assign(n, {});
  P.Ldiag.assign((size_t)n * bs, 0.0);
  for (int k = 0; k < n; ++k) {
    P.Lrows[k].assign(st[k].begin(), st[k].end());
    P.Lblk[k].assign(P.Lrows[k].size() * bs, 0.0);
  }
  // scatter A = Q + shift I (lower part, permuted): block (i,j) of A with pos(i) > pos(j) goes to
  // column pos(j), row pos(i); stored row-major as A_ij (= Q block (i,j))
  for (int i = 0; i < n; ++i)
    for (int e = P.rowptr[i]; e < P.rowptr[i + 1]; ++e) {
      const int j = P.colidx[e];
      const int a = P.iperm[i], b = P.iperm[j];
      const double *B = &P.blocks[(size_t)e * bs];
      if (a == b) {
        for (int x = 0; x < bs; ++x) P.Ldiag[(size_t)a * bs + x] += B[x];
      } else if (a > b) {
        const auto &rows = P.Lrows[b];
        const size_t pos = std::lower_bound(rows.begin(), rows.end(), a) - rows.begin();
        for (int x = 0; x < bs; ++x) P.Lblk[b][pos * bs + x] += B[x];
      }
    }
  for (int k = 0; k < n; ++k)
    for (int x = 0; x < dh; ++x) P.Ldiag[(size_t)k * bs + x * dh + x] += shift;
  // numeric, right-looking
  std::vector<double> W;
  for (int k = 0; k < n; ++k) {
    double *D = &P.Ldiag[(size_t)k * bs];
    if (!chol_small(D, dh)) return false;
    const auto &rows = P.Lrows[k];
    const size_t m = rows.size();
    double *Lk = P.Lblk[k].data();
    // L_ik = A_ik D^{-T}: solve X D^T = A_ik row by row (forward substitution on columns)
    for (size_t p = 0; p < m; ++p) {
      double *B = Lk + p * bs;
      for (int a = 0; a < dh; ++a)
        for (int c = 0; c < dh; ++c) {
          double s = B[a * dh + c];
          for (int q = 0; q < c; ++q) s -= B[a * dh + q] * D[c * dh + q];
          B[a * dh + c] = s / D[c * dh + c];
        }
    }
    // trailing update: A_ij -= L_ik L_jk^T for i >= j in struct(k)
    for (size_t pj = 0; pj < m; ++pj) {
      const int j = rows[pj];
      const double *Lj = Lk + pj * bs;
      {  // diagonal block of column j
        double *Dj = &P.Ldiag[(size_t)j * bs];
        for (int a = 0; a < dh; ++a)
          for (int c = 0; c < dh; ++c) {
            double s = 0;
            for (int q = 0; q < dh; ++q) s += Lj[a * dh + q] * Lj[c * dh + q];
            Dj[a * dh + c] -= s;
          }
      }
      const auto &rj = P.Lrows[j];
      size_t cursor = 0;
      for (size_t pi = pj + 1; pi < m; ++pi) {
        const int i = rows[pi];
        while (rj[cursor] != i) ++cursor;  // struct(k) \ {j} is contained in struct(j)
        const double *Li = Lk + pi * bs;
        double *T = &P.Lblk[j][cursor * bs];
        for (int a = 0; a < dh; ++a)
          for (int c = 0; c < dh; ++c) {
            double s = 0;
            for (int q = 0; q < dh; ++q) s += Li[a * dh + q] * Lj[c * dh + q];
            T[a * dh + c] -= s;
          }
      }
    }
  }
  P.factored = true;
  return true;
}

// Z = V (Q + 0.1 I)^{-1}  (V, Z are r x N; each pose tile is the transpose of a dh x r RHS block)
void solve(Problem &P, const double *V, double *Z) {
  const int n = P.n, dh = P.dh, r = P.r, tile = P.tile, bs = dh * dh;
  P.precon++;
  std::vector<double> y((size_t)n * tile);
  for (int k = 0; k < n; ++k) memcpy(&y[(size_t)k * tile], V + (size_t)P.perm[k] * tile, sizeof(double) * tile);
  // forward: y_k = D^{-1} y_k ; y_i -= L_ik y_k.  Tile layout [c*r + q] = component c of RHS q.
  for (int k = 0; k < n; ++k) {
    double *yk = &y[(size_t)k * tile];
    const double *D = &P.Ldiag[(size_t)k * bs];
    for (int c = 0; c < dh; ++c)
      for (int q = 0; q < r; ++q) {
        double s = yk[c * r + q];
        for (int a = 0; a < c; ++a) s -= D[c * dh + a] * yk[a * r + q];
        yk[c * r + q] = s / D[c * dh + c];
      }
    const auto &rows = P.Lrows[k];
    const double *Lk = P.Lblk[k].data();
    for (size_t p = 0; p < rows.size(); ++p) {
      double *yi = &y[(size_t)rows[p] * tile];
      const double *B = Lk + p * bs;
      for (int a = 0; a < dh; ++a)
        for (int c = 0; c < dh; ++c) {
          const double b = B[a * dh + c];
          for (int q = 0; q < r; ++q) yi[a * r + q] -= b * yk[c * r + q];
        }
    }
  }
  // backward: y_k -= sum_i L_ik^T y_i ; y_k = D^{-T} y_k
  for (int k = n - 1; k >= 0; --k) {
    double *yk = &y[(size_t)k * tile];
    const auto &rows = P.Lrows[k];
    const double *Lk = P.Lblk[k].data();
    for (size_t p = 0; p < rows.size(); ++p) {
      const double *yi = &y[(size_t)rows[p] * tile];
      const double *B = Lk + p * bs;
      for (int a = 0; a < dh; ++a)
        for (int c = 0; c < dh; ++c) {
          const double b = B[a * dh + c];
          for (int q = 0; q < r; ++q) yk[c * r + q] -= b * yi[a * r + q];
        }
    }
    const double *D = &P.Ldiag[(size_t)k * bs];
    for (int c = dh - 1; c >= 0; --c)
      for (int q = 0; q < r; ++q) {
        double s = yk[c * r + q];
        for (int a = c + 1; a < dh; ++a) s -= D[a * dh + c] * yk[a * r + q];
        yk[c * r + q] = s / D[c * dh + c];
      }
  }
  for (int k = 0; k < n; ++k) memcpy(Z + (size_t)P.perm[k] * tile, &y[(size_t)k * tile], sizeof(double) * tile);
}

// ---- manifold ----------------------------------------------------------------------------------------
void tangent(const Problem &P, const double *X, double *V, double *S /* nullable: d x d per pose */) {
  const int d = P.d, r = P.r, tile = P.tile;
  for (int i = 0; i < P.n; ++i) {
    const double *Y = X + (size_t)i * tile;
    double *W = V + (size_t)i * tile;
    double M[9], Sy[9];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        double s = 0;
        for (int q = 0; q < r; ++q) s += Y[a * r + q] * W[b * r + q];
        M[a * d + b] = s;
      }
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) Sy[a * d + b] = 0.5 * (M[a * d + b] + M[b * d + a]);
    if (S) memcpy(S + (size_t)i * d * d, Sy, sizeof(double) * d * d);
    for (int b = 0; b < d; ++b)
      for (int a = 0; a < d; ++a)
        for (int q = 0; q < r; ++q) W[b * r + q] -= Y[a * r + q] * Sy[a * d + b];
  }
}

void retract(const Problem &P, const double *X, const double *Eta, double *Out) {
  const int d = P.d, r = P.r, tile = P.tile;
  for (int i = 0; i < P.n; ++i) {
    double *o = Out + (size_t)i * tile;
    for (int k = 0; k < tile; ++k) o[k] = X[(size_t)i * tile + k] + Eta[(size_t)i * tile + k];
    for (int k = 0; k < d; ++k) {
      for (int pass = 0; pass < 2; ++pass)
        for (int p = 0; p < k; ++p) {
          double s = 0;
          for (int q = 0; q < r; ++q) s += o[p * r + q] * o[k * r + q];
          for (int q = 0; q < r; ++q) o[k * r + q] -= s * o[p * r + q];
        }
      double nn = 0;
      for (int q = 0; q < r; ++q) nn += o[k * r + q] * o[k * r + q];
      nn = 1.0 / sqrt(nn);
      for (int q = 0; q < r; ++q) o[k * r + q] *= nn;
    }
  }
}

double dot(const std::vector<double> &a, const std::vector<double> &b) {
  double s = 0;
  for (size_t k = 0; k < a.size(); ++k) s += a[k] * b[k];
  return s;
}

struct State {
  std::vector<double> x, EG, grad, S;
  double f, gn2;
};

void fgrad(Problem &P, State &st) {
  const size_t len = (size_t)P.r * P.N;
  st.EG.resize(len);
  st.grad.resize(len);
  st.S.resize((size_t)P.n * P.d * P.d);
  qx(P, st.x.data(), st.EG.data());
  double f = 0;
  for (size_t k = 0; k < len; ++k) {
    st.EG[k] += P.G[k];
    f += 0.5 * (st.EG[k] + P.G[k]) * st.x[k];
  }
  st.grad = st.EG;
  tangent(P, st.x.data(), st.grad.data(), st.S.data());
  st.f = f;
  st.gn2 = dot(st.grad, st.grad);
}

void hess(Problem &P, const State &st, const std::vector<double> &V, std::vector<double> &HV) {
  const int d = P.d, r = P.r, tile = P.tile;
  HV.resize(V.size());
  qx(P, V.data(), HV.data());
  for (int i = 0; i < P.n; ++i) {
    const double *Si = &st.S[(size_t)i * d * d];
    const double *Vi = &V[(size_t)i * tile];
    double *Hi = &HV[(size_t)i * tile];
    for (int b = 0; b < d; ++b)
      for (int a = 0; a < d; ++a)
        for (int q = 0; q < r; ++q) Hi[b * r + q] -= Vi[a * r + q] * Si[a * d + b];
  }
  tangent(P, st.x.data(), HV.data(), nullptr);
}

void precondition(Problem &P, const State &st, const std::vector<double> &V, std::vector<double> &Z) {
  Z.resize(V.size());
  solve(P, V.data(), Z.data());
  tangent(P, st.x.data(), Z.data(), nullptr);
}

}  // namespace

extern "C" {

struct cpu_result {
  double f_init, gn_init, f_opt, gn_opt;
  int outer, inner, accepted, rejected, tcg_status, pad;
  long qx, precon;
};

void *cpu_create(int n, int d, int r, int nnzb, const int *rowptr, const int *colidx, const double *blocks) {
  Problem *P = new Problem();
  P->n = n; P->d = d; P->r = r; P->dh = d + 1; P->N = (d + 1) * n; P->tile = r * (d + 1);
  P->rowptr.assign(rowptr, rowptr + n + 1);
  P->colidx.assign(colidx, colidx + nnzb);
  P->blocks.assign(blocks, blocks + (size_t)nnzb * (d + 1) * (d + 1));
  P->G.assign((size_t)r * P->N, 0.0);
  return P;
}
void cpu_destroy(void *h) { delete (Problem *)h; }
void cpu_set_G(void *h, const double *G) {
  Problem *P = (Problem *)h;
  memcpy(P->G.data(), G, sizeof(double) * P->G.size());
}
int cpu_factorize(void *h) { return factorize(*(Problem *)h, 0.1) ? 0 : -1; }
long cpu_factor_blocks(void *h) {
  long s = 0;
  for (const auto &c : ((Problem *)h)->Lrows) s += (long)c.size() + 1;
  return s;
}
void cpu_qx(void *h, const double *X, double *out) { qx(*(Problem *)h, X, out); }
void cpu_solve(void *h, const double *V, double *Z) { solve(*(Problem *)h, V, Z); }

// QuadraticOptimizer::optimize with RTR (ref: src/QuadraticOptimizer.cpp:26-108)
int cpu_optimize(void *h, const double *X0, double *Xout, double gradnorm_tol, int max_outer, int max_inner,
                 double init_radius, cpu_result *res) {
  Problem &P = *(Problem *)h;
  if (!P.factored && !factorize(P, 0.1)) return -1;
  const size_t len = (size_t)P.r * P.N;
  P.qx = P.precon = 0;
  State s1, s2;
  s1.x.assign(X0, X0 + len);
  fgrad(P, s1);
  res->f_init = s1.f;
  res->gn_init = sqrt(s1.gn2);
  res->outer = res->inner = res->accepted = res->rejected = 0;
  // tCG status codes of the reference's ROPTResult (LCON, SCON, NEGCURVTURE, EXCREGION, MAXITER)
  enum { kLcon = 0, kScon = 1, kNegCurv = 2, kExcRegion = 3, kMaxIter = 4 };
  const double theta = 1.0, kappa = 0.1, accept_rho = 0.1, shrink = 0.25, magnify = 2.0;   // ROPTLIB defaults
  res->tcg_status = kMaxIter;
  const bool single = (max_outer == 1);                     // src/QuadraticOptimizer.cpp:80-98
  double radius = init_radius, Delta = init_radius, max_Delta = single ? init_radius : 5 * init_radius;
  int shrinks = 0;
  bool run = sqrt(s1.gn2) >= gradnorm_tol && max_outer > 0;
  std::vector<double> eta(len), r(len), z, delta(len), Hd;
  while (run) {
    // ---- Steihaug-Toint truncated CG from eta = 0 (oracle/pgo.py: tcg)
    r = s1.grad;
    const double norm_r0 = sqrt(s1.gn2);
    precondition(P, s1, r, z);
    double z_r = dot(z, r), d_Pd = z_r, e_Pd = 0.0, e_Pe = 0.0;
    int status = kMaxIter, inner = 0;
    for (size_t k = 0; k < len; ++k) { delta[k] = -z[k]; eta[k] = 0; }
    for (int j = 0; j < max_inner; ++j) {
      hess(P, s1, delta, Hd);
      inner = j + 1;
      const double d_Hd = dot(delta, Hd);
      const double alpha = z_r / d_Hd;
      const double e_Pe_next = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
      if (d_Hd <= 0.0 || e_Pe_next >= Delta * Delta) {      // leave along delta up to the boundary
        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
        for (size_t k = 0; k < len; ++k) eta[k] += tau * delta[k];
        status = d_Hd <= 0.0 ? kNegCurv : kExcRegion;
        break;
      }
      e_Pe = e_Pe_next;
      double r_r = 0;
      for (size_t k = 0; k < len; ++k) {
        eta[k] += alpha * delta[k];
        r[k] += alpha * Hd[k];
        r_r += r[k] * r[k];
      }
      const double lim = pow(norm_r0, theta);
      if (sqrt(r_r) <= norm_r0 * std::min(lim, kappa)) {
        status = kappa < lim ? kLcon : kScon;
        break;
      }
      precondition(P, s1, r, z);
      const double z_r_prev = z_r;
      z_r = dot(z, r);
      const double beta = z_r / z_r_prev;
      for (size_t k = 0; k < len; ++k) delta[k] = -z[k] + beta * delta[k];
      e_Pd = beta * (e_Pd + alpha * d_Pd);
      d_Pd = z_r + beta * beta * d_Pd;
    }
    res->inner += inner;
    res->tcg_status = status;
    // ---- candidate, ratio test, radius update (oracle/pgo.py: rtr_run)
    s2.x.resize(len);
    retract(P, s1.x.data(), eta.data(), s2.x.data());
    fgrad(P, s2);
    hess(P, s1, eta, Hd);
    const double rho = (s1.f - s2.f) / (-(dot(eta, s1.grad) + 0.5 * dot(eta, Hd)));
    if (rho > 0.75) {
      if (status == kExcRegion || status == kNegCurv) Delta = std::min(magnify * Delta, max_Delta);
    } else if (rho < 0.25) {
      Delta = shrink * Delta;
    }
    const double sqeps = sqrt(2.220446049250313e-16);
    const bool acc = rho > accept_rho || (fabs(s1.f - s2.f) / (fabs(s1.f) + 1.0) < sqeps && s2.f < s1.f);
    if (acc) { std::swap(s1, s2); res->accepted++; } else { res->rejected++; }
    res->outer++;
    if (single) {
      if (acc) run = false;
      else if (shrinks > 10) run = false;
      else { radius *= 0.25; shrinks++; Delta = radius; max_Delta = radius; }
    } else {
      run = res->outer < max_outer && !(sqrt(s1.gn2) < gradnorm_tol);
    }
  }
  res->f_opt = s1.f;
  res->gn_opt = sqrt(s1.gn2);
  res->qx = P.qx;
  res->precon = P.precon;
  memcpy(Xout, s1.x.data(), sizeof(double) * len);
  return 0;
}

}  // extern "C"
