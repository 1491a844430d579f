// Host-only: index plan of the three-phase form of the two-level exact preconditioner
// (precon_mode 3).  Used by the device set-up (precon_dd.cu) and exposed for inspection / CPU tests
// through dpgo_three_phase_plan (include/dpgo_b200.h): the CPU test fills the stage buffers from a
// numpy inverse according to this plan, replays the three strip phases and compares with the exact
// (Q + 0.1 I)^-1 -- so everything but the CUDA code itself is checked without a device.
//
// Algebra (A = Q + 0.1 I permuted to [D_1 .. D_K | S], M_k = A_k^-1, C_k = M_k A_kS, which is non-zero
// only in the columns of S_k = separator poses with a neighbour in D_k):
//   phase 1:  y_k = r_k M_k                    and   g_k = r_k C_k[:, S_k]      (same input slice)
//   phase 3:  t_S = r_S - sum_k scatter(g_k);        z_S = t_S Sigma^-1,  Sigma = A_SS - sum_k A_Sk C_k
//   phase 5:  z_k = y_k - z_S[S_k] C_k[:, S_k]^T
// i.e. the two sparse coupling phases of the five-phase form (mode 2) are folded into dense strips:
// three grid phases per application instead of five.
//
// Column spaces.  "y space": [D_1 | .. | D_K | S | T_1 | .. | T_K], every segment padded to 64
// scalars; T_k holds g_k in the compact order of S_k.  Phase-5 strips read z_S through a gather list
// (compact index of S_k -> scalar column of the S segment).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "dissect.h"

namespace dpgo {

constexpr int kPlanCols = 64;    // output columns per strip  (== kGemvCols)
constexpr int kPlanStageK = 32;  // inner indices per stage    (== kStageK)

struct PlanStrip {               // same members as DdStrip (kernels.cuh)
  int cb, kc0, nchunks, slot;
  long long data_off;
};

// what a strip multiplies with: kind 0 = M_k column block blk; 1 = C_k[:, S_k] T-block blk (phase 1);
// 2 = Sigma^-1 column block blk, inner split k; 3 = C_k[:, S_k]^T column block blk of domain k (phase 5)
struct PlanTile {
  int kind, k, blk;
};

struct ThreePhasePlan {
  int n = 0, dh = 0, K = 0, nS = 0, V = 1;
  int sep_col0 = 0, pcols = 0, ycols = 0, nsplit3 = 1;
  long long stages1 = 0, stages3 = 0, stages5 = 0;
  double bytes_per_apply = 0;     // dense blocks streamed by one application (without the vectors)
  std::vector<int> group;         // [n] domain of pose i, -1 = separator
  std::vector<int> lpos;          // [n] position of pose i inside its domain / inside the separator order
  std::vector<int> pcol;          // [n] scalar column of pose i in the permuted space
  std::vector<int> srow;          // [nS] pose id of separator position j
  std::vector<int> icol;          // [ycols] original scalar column of a permuted column, -1 = padding / T
  std::vector<int> dom_off, dom_m, dom_pad;   // [K] first column, scalars, padded scalars of D_k
  std::vector<int> t_off, t_m, t_pad;         // [K] the same for T_k
  std::vector<int> sk_ptr, sk;                // S_k as separator positions, CSR over the domains
  std::vector<int> tptr, tcol;    // per scalar column j of the padded S segment: y-space columns to subtract
  std::vector<int> gchunk;        // [K] first chunk of domain k in gidx
  std::vector<int> gidx;          // compact index -> scalar column of the permuted space, -1 = padding
  std::vector<PlanStrip> strips1, strips3, strips5;
  std::vector<PlanTile> tiles1, tiles3, tiles5;
  std::vector<int> cta1, cta3, cta5, chunks1, chunks3, chunks5;
};

// greedy longest-first balancing of the strips of one phase over V virtual CTAs (same rule as mode 2)
inline void plan_balance(int V, std::vector<PlanStrip> &strips, std::vector<PlanTile> &tiles, std::vector<int> &cta,
                         std::vector<int> &chunks) {
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
  std::vector<PlanStrip> s2;
  std::vector<PlanTile> t2;
  cta.assign(V + 1, 0);
  for (int v = 0; v < V; ++v) {
    for (int i : bins[v]) {
      s2.push_back(strips[i]);
      t2.push_back(tiles[i]);
    }
    cta[v + 1] = (int)s2.size();
  }
  strips.swap(s2);
  tiles.swap(t2);
}

// Balancing that keeps strips with the same input slice (same kc0 / nchunks: the strips of one domain in
// phases 1 and 5) on the same virtual CTAs, so that a CTA stages the slice once (phase_strip_gemv skips
// the staging of a strip whose slice is already in shared memory).  Every group gets at least one CTA,
// the remaining CTAs go one at a time to the group with the largest load per CTA (never more CTAs than
// strips); inside a group the strips are dealt longest-first.  Falls back to plan_balance when there
// are more groups than CTAs.
inline void plan_balance_affine(int V, std::vector<PlanStrip> &strips, std::vector<PlanTile> &tiles,
                                std::vector<int> &cta, std::vector<int> &chunks) {
  std::vector<std::pair<std::pair<int, int>, std::vector<int>>> groups;
  for (size_t i = 0; i < strips.size(); ++i) {
    const std::pair<int, int> key(strips[i].kc0, strips[i].nchunks);
    size_t g = 0;
    while (g < groups.size() && groups[g].first != key) ++g;
    if (g == groups.size()) groups.push_back({key, {}});
    groups[g].second.push_back((int)i);
  }
  const int G = (int)groups.size();
  if (G == 0 || G > V) {
    plan_balance(V, strips, tiles, cta, chunks);
    return;
  }
  // all strips of a group have the same length, so the busiest CTA of group g works
  // ceil(strips_g / nct_g) * nchunks_g chunks: give the spare CTAs to the group where that is largest
  std::vector<int> nct(G, 1);
  auto worst = [&](int g) {
    const int ns = (int)groups[g].second.size();
    return (long)((ns + nct[g] - 1) / nct[g]) * groups[g].first.second;
  };
  for (int left = V - G; left > 0; --left) {
    int best = -1;
    for (int g = 0; g < G; ++g) {
      if (nct[g] >= (int)groups[g].second.size()) continue;
      if (best < 0 || worst(g) > worst(best)) best = g;
    }
    if (best < 0) break;
    nct[best]++;
  }
  std::vector<PlanStrip> s2;
  std::vector<PlanTile> t2;
  cta.assign(1, 0);
  chunks.clear();
  for (int g = 0; g < G; ++g) {
    std::vector<int> order = groups[g].second;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return strips[a].nchunks > strips[b].nchunks; });
    std::vector<std::vector<int>> bins(nct[g]);
    std::vector<int> bl(nct[g], 0);
    for (int i : order) {
      int best = 0;
      for (int v = 1; v < nct[g]; ++v)
        if (bl[v] < bl[best]) best = v;
      bins[best].push_back(i);
      bl[best] += strips[i].nchunks;
    }
    for (int v = 0; v < nct[g]; ++v) {
      for (int i : bins[v]) {
        s2.push_back(strips[i]);
        t2.push_back(tiles[i]);
      }
      cta.push_back((int)s2.size());
      chunks.push_back(bl[v]);
    }
  }
  while ((int)chunks.size() < V) {   // idle virtual CTAs
    cta.push_back((int)s2.size());
    chunks.push_back(0);
  }
  strips.swap(s2);
  tiles.swap(t2);
}

inline int plan_round_up(int x, int m) { return (x + m - 1) / m * m; }

// n block rows of a symmetric block pattern; dh scalars per pose; interior domains of <= max_domain_poses;
// V virtual CTAs; split3 > 0 forces the inner split of the Schur strips; max_wave = stages of one wave.
// affine: phases 1 and 5 use plan_balance_affine (strips of one domain share CTAs) instead of plan_balance.
inline ThreePhasePlan build_three_phase_plan(int n, const int *rowptr, const int *colidx, int dh, int max_domain_poses,
                                             int V, int split3, int max_wave, bool affine = false) {
  ThreePhasePlan p;
  p.n = n; p.dh = dh; p.V = std::max(1, V);
  const std::vector<std::vector<int>> adj = bsr_adjacency(n, rowptr, colidx);
  Dissector ds(adj, max_domain_poses);
  {
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    ds.run(std::move(all));
  }
  const int K = (int)ds.domains.size();
  p.K = K;
  p.nS = (int)ds.sep.size();
  p.group.assign(n, -1);
  p.lpos.assign(n, 0);
  p.pcol.assign(n, 0);
  p.dom_off.resize(K); p.dom_m.resize(K); p.dom_pad.resize(K);
  int col = 0;
  for (int k = 0; k < K; ++k) {
    auto &dom = ds.domains[k];
    std::sort(dom.begin(), dom.end());
    p.dom_m[k] = (int)dom.size() * dh;
    p.dom_pad[k] = plan_round_up(p.dom_m[k], kPlanCols);
    p.dom_off[k] = col;
    for (size_t j = 0; j < dom.size(); ++j) {
      p.group[dom[j]] = k;
      p.lpos[dom[j]] = (int)j;
      p.pcol[dom[j]] = col + (int)j * dh;
    }
    col += p.dom_pad[k];
  }
  p.sep_col0 = col;
  // separator order: poses that touch the same set of domains are adjacent (pose id breaks ties), so
  // that the separator poses of one domain fall into few runs
  std::vector<std::vector<int>> sig(n);
  for (int v : ds.sep) {
    for (int u : adj[v])
      if (p.group[u] >= 0) sig[v].push_back(p.group[u]);
    std::sort(sig[v].begin(), sig[v].end());
    sig[v].erase(std::unique(sig[v].begin(), sig[v].end()), sig[v].end());
  }
  p.srow = ds.sep;
  std::sort(p.srow.begin(), p.srow.end(), [&](int a, int b) {
    if (sig[a] != sig[b]) return sig[a] < sig[b];
    return a < b;
  });
  const int mS = p.nS * dh, padS = plan_round_up(mS, kPlanCols);
  for (int j = 0; j < p.nS; ++j) {
    p.lpos[p.srow[j]] = j;
    p.pcol[p.srow[j]] = col + j * dh;
  }
  col += padS;
  p.pcols = std::max(col, kPlanCols);
  // S_k: separator positions with a neighbour in D_k, ascending
  std::vector<std::vector<int>> Sk(K);
  for (int j = 0; j < p.nS; ++j)
    for (int k : sig[p.srow[j]]) Sk[k].push_back(j);
  p.sk_ptr.assign(K + 1, 0);
  p.t_off.resize(K); p.t_m.resize(K); p.t_pad.resize(K);
  col = p.pcols;
  for (int k = 0; k < K; ++k) {
    p.sk_ptr[k + 1] = p.sk_ptr[k] + (int)Sk[k].size();
    p.sk.insert(p.sk.end(), Sk[k].begin(), Sk[k].end());
    p.t_m[k] = (int)Sk[k].size() * dh;
    p.t_pad[k] = plan_round_up(p.t_m[k], kPlanCols);
    p.t_off[k] = col;
    col += p.t_pad[k];
  }
  p.ycols = col;
  p.icol.assign(p.ycols, -1);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < dh; ++c) p.icol[p.pcol[i] + c] = i * dh + c;
  // t_S = r_S - sum_k scatter(g_k): per scalar column of the S segment, the T columns that hold its terms
  {
    std::vector<std::vector<int>> src(padS);
    for (int k = 0; k < K; ++k)
      for (size_t a = 0; a < Sk[k].size(); ++a)
        for (int c = 0; c < dh; ++c) src[Sk[k][a] * dh + c].push_back(p.t_off[k] + (int)a * dh + c);
    p.tptr.assign(padS + 1, 0);
    for (int j = 0; j < padS; ++j) {
      p.tptr[j + 1] = p.tptr[j] + (int)src[j].size();
      p.tcol.insert(p.tcol.end(), src[j].begin(), src[j].end());
    }
  }
  // gather lists of phase 5: compact index of S_k -> scalar column of the S segment
  p.gchunk.assign(K, 0);
  for (int k = 0; k < K; ++k) {
    p.gchunk[k] = (int)p.gidx.size() / kPlanStageK;
    const int len = plan_round_up(p.t_m[k], kPlanStageK);
    for (int i = 0; i < len; ++i)
      p.gidx.push_back(i < p.t_m[k] ? p.sep_col0 + Sk[k][i / dh] * dh + i % dh : -1);
  }
  // ---- strips
  long long stage = 0;
  double bytes = 0;
  for (int k = 0; k < K; ++k) {   // phase 1: M_k, then C_k[:, S_k]
    const int nch = p.dom_pad[k] / kPlanStageK;
    for (int cb = 0; cb < p.dom_pad[k] / kPlanCols; ++cb) {
      p.strips1.push_back(PlanStrip{p.dom_off[k] / kPlanCols + cb, p.dom_off[k] / kPlanStageK, nch, 0, stage});
      p.tiles1.push_back(PlanTile{0, k, cb});
      stage += nch;
    }
    for (int tb = 0; tb < p.t_pad[k] / kPlanCols; ++tb) {
      p.strips1.push_back(PlanStrip{p.t_off[k] / kPlanCols + tb, p.dom_off[k] / kPlanStageK, nch, 0, stage});
      p.tiles1.push_back(PlanTile{1, k, tb});
      stage += nch;
    }
    bytes += ((double)p.dom_m[k] + p.t_m[k]) * p.dom_m[k] * 8;
  }
  p.stages1 = stage;
  const int nchS = padS / kPlanStageK, ncbS = padS / kPlanCols;
  int nsplit3 = split3;
  if (nsplit3 <= 0) {   // same cost model as the five-phase form (precon_dd.cu)
    nsplit3 = 1;
    double best_cost = 1e300;
    for (int ns = 1; ns <= std::min(std::max(nchS, 1), 32); ++ns) {
      const int cps = (nchS + ns - 1) / ns;
      const long strips = (long)ncbS * ((nchS + cps - 1) / std::max(cps, 1));
      const double per_strip = 4.0 + 2.0 * ((cps + 1) / 2) / 4.0 + 3.0 * ((cps + max_wave - 1) / max_wave - 1);
      const double cost = (double)((strips + p.V - 1) / p.V) * per_strip + 0.3 * ns;
      if (cost < best_cost - 1e-9) { best_cost = cost; nsplit3 = ns; }
    }
  }
  nsplit3 = std::max(1, std::min(nsplit3, std::max(nchS, 1)));
  const int cps = nchS > 0 ? (nchS + nsplit3 - 1) / nsplit3 : 1;
  nsplit3 = nchS > 0 ? (nchS + cps - 1) / cps : 1;
  p.nsplit3 = nsplit3;
  for (int cb = 0; cb < ncbS; ++cb)
    for (int sp = 0; sp < nsplit3; ++sp) {
      const int c0 = sp * cps, nc = std::min(cps, nchS - c0);
      if (nc <= 0) continue;
      p.strips3.push_back(PlanStrip{p.sep_col0 / kPlanCols + cb, p.sep_col0 / kPlanStageK + c0, nc, sp,
                                    (long long)cb * nchS + c0});
      p.tiles3.push_back(PlanTile{2, sp, cb});
    }
  p.stages3 = (long long)ncbS * nchS;
  bytes += (double)mS * mS * 8;
  stage = 0;
  for (int k = 0; k < K && p.nS > 0; ++k) {   // phase 5: C_k[:, S_k]^T
    // a domain without separator neighbours (its own component) keeps empty strips (0 chunks): they
    // write w = 0 and, when the finish is fused into this phase, complete its poses with z = y
    const int nch = plan_round_up(p.t_m[k], kPlanStageK) / kPlanStageK;
    for (int cb = 0; cb < p.dom_pad[k] / kPlanCols; ++cb) {
      p.strips5.push_back(PlanStrip{p.dom_off[k] / kPlanCols + cb, p.gchunk[k], nch, 0, stage});
      p.tiles5.push_back(PlanTile{3, k, cb});
      stage += nch;
    }
    bytes += (double)p.dom_m[k] * p.t_m[k] * 8;
  }
  p.stages5 = stage;
  p.bytes_per_apply = bytes;
  if (affine) plan_balance_affine(p.V, p.strips1, p.tiles1, p.cta1, p.chunks1);
  else plan_balance(p.V, p.strips1, p.tiles1, p.cta1, p.chunks1);
  plan_balance(p.V, p.strips3, p.tiles3, p.cta3, p.chunks3);
  if (affine) plan_balance_affine(p.V, p.strips5, p.tiles5, p.cta5, p.chunks5);
  else plan_balance(p.V, p.strips5, p.tiles5, p.cta5, p.chunks5);
  return p;
}

// Flat int64 image of the plan for dpgo_three_phase_plan: out[0] = number of sections, then
// (offset, length) per section, then the sections in the order listed in include/dpgo_b200.h.
inline std::vector<int64_t> serialize_three_phase_plan(const ThreePhasePlan &p) {
  std::vector<std::vector<int64_t>> sec;
  auto ints = [&](const std::vector<int> &v) { sec.emplace_back(v.begin(), v.end()); };
  auto strips = [&](const std::vector<PlanStrip> &s, const std::vector<PlanTile> &t) {
    std::vector<int64_t> o;
    for (size_t i = 0; i < s.size(); ++i) {
      const int64_t row[8] = {s[i].cb, s[i].kc0, s[i].nchunks, s[i].slot, s[i].data_off, t[i].kind, t[i].k, t[i].blk};
      o.insert(o.end(), row, row + 8);
    }
    sec.push_back(std::move(o));
  };
  sec.push_back({p.n, p.dh, p.K, p.nS, p.V, p.sep_col0, p.pcols, p.ycols, p.nsplit3, p.stages1, p.stages3, p.stages5,
                 (int64_t)p.bytes_per_apply});
  ints(p.group); ints(p.pcol); ints(p.srow); ints(p.icol);
  ints(p.dom_off); ints(p.dom_m); ints(p.dom_pad);
  ints(p.t_off); ints(p.t_m); ints(p.t_pad);
  ints(p.sk_ptr); ints(p.sk); ints(p.tptr); ints(p.tcol); ints(p.gchunk); ints(p.gidx);
  strips(p.strips1, p.tiles1); strips(p.strips3, p.tiles3); strips(p.strips5, p.tiles5);
  ints(p.cta1); ints(p.cta3); ints(p.cta5);
  std::vector<int64_t> out;
  const int64_t ns = (int64_t)sec.size();
  out.push_back(ns);
  int64_t off = 1 + 2 * ns;
  for (const auto &s : sec) {
    out.push_back(off);
    out.push_back((int64_t)s.size());
    off += (int64_t)s.size();
  }
  for (const auto &s : sec) out.insert(out.end(), s.begin(), s.end());
  return out;
}

}  // namespace dpgo
