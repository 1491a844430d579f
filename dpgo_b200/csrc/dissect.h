// Host-only: vertex-separator nested dissection of a pose graph by BFS level sets.  Used by the
// two-level preconditioner set-up (precon_dd.cu) and exposed for inspection / CPU tests through
// dpgo_two_level_partition (include/dpgo_b200.h).
#pragma once
#include <algorithm>
#include <deque>
#include <vector>

namespace dpgo {

// Recursively splits the graph until every part has at most `thr` vertices: a BFS from a
// pseudo-peripheral vertex gives level sets; the smallest level that leaves at least 30 % of the
// vertices on either side becomes a separator, the two sides are dissected further.  Parts that
// cannot be split (BFS depth < 2) stay whole.  Result: `domains` (no edge joins two different
// domains) and `sep` (all separator vertices).
struct Dissector {
  const std::vector<std::vector<int>> &adj;
  int thr;
  std::vector<int> mark;  // scratch: generation stamps
  std::vector<int> level;
  int gen = 0;
  std::vector<std::vector<int>> domains;
  std::vector<int> sep;

  Dissector(const std::vector<std::vector<int>> &a, int t) : adj(a), thr(t), mark(a.size(), 0), level(a.size(), 0) {}

  // BFS inside `nodes` (identified by mark == gen_in) from start; returns visit order, fills level
  std::vector<int> bfs(int start, int gen_in) {
    const int g = ++gen;
    std::vector<int> order;
    std::deque<int> q;
    q.push_back(start);
    mark[start] = g;  // visited stamp (> gen_in)
    level[start] = 0;
    while (!q.empty()) {
      const int u = q.front();
      q.pop_front();
      order.push_back(u);
      for (int v : adj[u])
        if (mark[v] == gen_in) {
          mark[v] = g;
          level[v] = level[u] + 1;
          q.push_back(v);
        }
    }
    // restore membership stamps of the visited nodes
    for (int v : order) mark[v] = gen_in;
    return order;
  }

  void run(std::vector<int> nodes) {
    if ((int)nodes.size() <= thr) {
      if (!nodes.empty()) domains.push_back(std::move(nodes));
      return;
    }
    const int g = ++gen;
    for (int v : nodes) mark[v] = g;
    std::vector<int> order = bfs(nodes[0], g);
    if (order.size() < nodes.size()) {  // disconnected: split off the component
      const int gc = ++gen;
      for (int v : order) mark[v] = gc;
      std::vector<int> rest;
      for (int v : nodes)
        if (mark[v] != gc) rest.push_back(v);
      run(std::move(order));
      run(std::move(rest));
      return;
    }
    for (int pass = 0; pass < 2; ++pass) order = bfs(order.back(), g);  // pseudo-peripheral start
    int L = 0;
    for (int v : nodes) L = std::max(L, level[v]);
    if (L < 2) {
      domains.push_back(std::move(nodes));
      return;
    }
    std::vector<int> counts(L + 1, 0);
    for (int v : nodes) counts[level[v]]++;
    std::vector<long> cum(L + 1, 0);
    for (int l = 0; l <= L; ++l) cum[l] = counts[l] + (l ? cum[l - 1] : 0);
    int best = -1;
    const double nn = (double)nodes.size();
    for (int l = 1; l < L; ++l) {
      const long left = cum[l - 1], right = (long)nodes.size() - cum[l];
      if (std::min(left, right) >= 0.3 * nn && (best < 0 || counts[l] < counts[best])) best = l;
    }
    if (best < 0) {
      best = 1;
      while (best < L - 1 && cum[best] < nn / 2) ++best;
    }
    std::vector<int> lo, hi;
    for (int v : nodes) {
      if (level[v] == best) sep.push_back(v);
      else if (level[v] < best) lo.push_back(v);
      else hi.push_back(v);
    }
    run(std::move(lo));
    run(std::move(hi));
  }
};

// adjacency lists (without self loops) of a block-CSR pattern
inline std::vector<std::vector<int>> bsr_adjacency(int n, const int *rowptr, const int *colidx) {
  std::vector<std::vector<int>> adj(n);
  for (int i = 0; i < n; ++i)
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e)
      if (colidx[e] != i) adj[i].push_back(colidx[e]);
  return adj;
}

}  // namespace dpgo
