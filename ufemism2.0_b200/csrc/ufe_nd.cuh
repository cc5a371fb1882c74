// Elimination tree of the nested-dissection analysis (ufe_nd.cu), shared with the numeric phase (ufe_nd_numeric.cu).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

struct NdNode {
  int level = 0, parent = -1, child[2] = {-1, -1};
  std::vector<int> sep, bnd;          // 0-based triangle ids, ascending
  std::vector<int> up;                // position of bnd[k] in the parent's [sep; bnd] list
};

struct ufe_nd_tree {
  int nT = 0, n_levels = 0;
  std::vector<NdNode> nodes;          // post-order: children before their parent, root last
  std::vector<int> node_of;           // (nT) node that eliminates each triangle
  std::vector<int> pos_in_sep;        // (nT) position of the triangle in its node's sep list
  std::vector<int> entry_node, entry_row, entry_col;   // per block entry of the input pattern: front and block position
  std::vector<int> bptr, bind;        // the analysed block pattern (the numeric phase looks scalar entries up in it)
};


// owner rank / number of ranks below every node for a factorisation distributed over nranks ranks (ufe_nd.cu)
int ufe_nd_owner_map(const ufe_nd_tree *T, int nranks, std::vector<int> &owner, std::vector<int> &span);

// host threads for the once-per-mesh analysis (4 M unknowns: seconds of sorting and searching otherwise)
namespace ufe_nd_host {
inline int host_threads() {
  static const int n = [] {
    if (const char *e = getenv("UFE_ND_HOST_THREADS")) return std::max(1, atoi(e));
    return (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  }();
  return n;
}
template <class Fn> void parallel_for(int n, Fn fn) {           // fn(begin, end) on contiguous chunks
  const int nt = std::min(host_threads(), std::max(1, n / 4096));
  if (nt <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  for (int q = 0; q < nt; q++) th.emplace_back([=] { fn((int)((long long)n * q / nt), (int)((long long)n * (q + 1) / nt)); });
  for (auto &t : th) t.join();
}
}  // namespace ufe_nd_host
