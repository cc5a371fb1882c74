// Nested-dissection symbolic analysis of the stiffness matrix (host code; staged for the multifrontal exact solver
// that replaces the banded block cyclic reduction on wide meshes -- DESIGN.md section 9, tools/nd_prototype.py).
//
// Works on the block graph of the matrix: one node per triangle (its 2x2 (u,v) block), an edge where the stiffness
// matrix couples two triangles (solve_linearised_SSA_DIVA.f90:180-329: the 2-ring stencil of the M2_*_b_b operators).
// Recursive coordinate bisection on the triangle centroids; the separator of a cut is the set of triangles of the
// lower half that are coupled to the upper half.  Per tree node the analysis yields
//   sep   the triangles eliminated at the node (a leaf eliminates its whole sub-domain),
//   bnd   the not-yet-eliminated triangles coupled to the node's subtree (they live in ancestors' separators),
// i.e. the node's dense front is [sep; bnd] x [sep; bnd]; plus the index maps the numeric phase needs:
//   child -> parent positions for the extend-add of Schur complements,
//   matrix block entry -> (front, row, column) for the assembly.
// Everything is built once per mesh / BC configuration, like the stiffness pattern itself.
#include "ufe_internal.cuh"

#include <algorithm>
#include <atomic>
#include <future>
#include <numeric>

#include "ufe_nd.cuh"

using ufe_nd_host::host_threads;
using ufe_nd_host::parallel_for;

namespace {

struct Graph { std::vector<int> ptr, ind; };   // symmetric adjacency without self loops, 0-based, ascending

Graph symmetrise(int nT, const int *bptr, const int *bind) {
  std::vector<int> cnt(nT + 1, 0);
  for (int i = 0; i < nT; i++)
    for (int k = bptr[i]; k < bptr[i + 1]; k++) { const int j = bind[k]; if (j != i) { cnt[i + 1]++; cnt[j + 1]++; } }
  std::vector<int> ptr(nT + 1, 0);
  std::partial_sum(cnt.begin(), cnt.end(), ptr.begin());
  std::vector<int> ind(ptr[nT]), fill(ptr.begin(), ptr.end() - 1);
  for (int i = 0; i < nT; i++)
    for (int k = bptr[i]; k < bptr[i + 1]; k++) { const int j = bind[k]; if (j != i) { ind[fill[i]++] = j; ind[fill[j]++] = i; } }
  Graph G;
  G.ptr.assign(nT + 1, 0);
  std::vector<int> len(nT);
  parallel_for(nT, [&](int i0, int i1) {
    for (int i = i0; i < i1; i++) {
      std::sort(ind.begin() + ptr[i], ind.begin() + ptr[i + 1]);
      len[i] = (int)(std::unique(ind.begin() + ptr[i], ind.begin() + ptr[i + 1]) - (ind.begin() + ptr[i]));
    }
  });
  for (int i = 0; i < nT; i++) G.ptr[i + 1] = G.ptr[i] + len[i];
  G.ind.resize(G.ptr[nT]);
  parallel_for(nT, [&](int i0, int i1) {
    for (int i = i0; i < i1; i++) std::copy(ind.begin() + ptr[i], ind.begin() + ptr[i] + len[i], G.ind.begin() + G.ptr[i]);
  });
  return G;
}

struct Builder {
  const double *x, *y;
  const Graph &G;
  int leaf;
  std::vector<NdNode> &nodes;
  std::vector<char> &mark;            // scratch flags (nT), shared: the recursion touches the flags of its own triangles only
  int par_levels;                     // the two halves of a cut run concurrently above this level

  // appends the sub-tree of `idx` to `nodes` in post-order and returns the index of its root
  int dissect(std::vector<int> &idx, int level) {
    const int n = (int)idx.size();
    auto make_leaf = [&]() {
      NdNode nd; nd.level = level; nd.sep = idx; std::sort(nd.sep.begin(), nd.sep.end());
      nodes.push_back(std::move(nd));
      return (int)nodes.size() - 1;
    };
    if (n <= leaf) return make_leaf();
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int t : idx) { xmin = std::min(xmin, x[t]); xmax = std::max(xmax, x[t]); ymin = std::min(ymin, y[t]); ymax = std::max(ymax, y[t]); }
    const double *c = (xmax - xmin >= ymax - ymin) ? x : y;
    std::vector<double> v(n);
    for (int k = 0; k < n; k++) v[k] = c[idx[k]];
    std::nth_element(v.begin(), v.begin() + n / 2, v.end());
    double med = v[n / 2];
    if (n % 2 == 0) med = 0.5 * (med + *std::max_element(v.begin(), v.begin() + n / 2));
    std::vector<int> left, right;
    for (int t : idx) (c[t] <= med ? left : right).push_back(t);
    if (left.empty() || right.empty()) return make_leaf();
    for (int t : left) mark[t] = 1;
    std::vector<int> sep;
    for (int r : right)
      for (int k = G.ptr[r]; k < G.ptr[r + 1]; k++) { const int j = G.ind[k]; if (mark[j] == 1) { mark[j] = 2; sep.push_back(j); } }
    std::vector<int> rest;
    for (int t : left) { if (mark[t] == 1) rest.push_back(t); mark[t] = 0; }
    std::sort(sep.begin(), sep.end());
    idx.clear(); idx.shrink_to_fit();
    int c0, c1;
    if (level < par_levels && n > 8 * leaf) {
      // child 1 is built into a private vector by another thread and appended behind child 0's sub-tree
      std::vector<NdNode> other;
      Builder B1{x, y, G, leaf, other, mark, par_levels};
      auto fut = std::async(std::launch::async, [&] { return B1.dissect(right, level + 1); });
      c0 = dissect(rest, level + 1);
      const int r1 = fut.get(), off = (int)nodes.size();
      for (NdNode &o : other) {
        if (o.parent >= 0) o.parent += off;
        for (int &c : o.child) if (c >= 0) c += off;
        nodes.push_back(std::move(o));
      }
      c1 = r1 + off;
    } else {
      c0 = dissect(rest, level + 1);
      c1 = dissect(right, level + 1);
    }
    NdNode nd; nd.level = level; nd.sep = std::move(sep); nd.child[0] = c0; nd.child[1] = c1;
    nodes.push_back(std::move(nd));
    const int me = (int)nodes.size() - 1;
    nodes[c0].parent = me; nodes[c1].parent = me;
    return me;
  }
};

}  // namespace

// bptr / bind: block pattern of the matrix, 0-based CSR over triangles (row i couples to the listed triangles).
extern "C" int ufe_nd_analyse(int32_t nT, const double *centroid_x, const double *centroid_y, const int32_t *bptr,
                              const int32_t *bind, int32_t leaf_triangles, ufe_nd_tree **out) {
  if (!centroid_x || !centroid_y || !bptr || !bind || !out || nT <= 0 || leaf_triangles < 1) {
    ufe_set_error("ufe_nd_analyse: bad argument"); return UFE_ERR_INVALID;
  }
  for (int i = 0; i < nT; i++)
    for (int k = bptr[i]; k < bptr[i + 1]; k++)
      if (bind[k] < 0 || bind[k] >= nT) { ufe_set_error("ufe_nd_analyse: column out of range in row %d", i); return UFE_ERR_INVALID; }
  ufe_nd_tree *T = new ufe_nd_tree();
  T->nT = nT;
  T->bptr.assign(bptr, bptr + nT + 1); T->bind.assign(bind, bind + bptr[nT]);
  const Graph G = symmetrise(nT, bptr, bind);
  {
    std::vector<char> mark(nT, 0);
    int par_levels = 0;
    while ((1 << par_levels) < host_threads()) par_levels++;
    Builder B{centroid_x, centroid_y, G, leaf_triangles, T->nodes, mark, par_levels};
    std::vector<int> all(nT);
    std::iota(all.begin(), all.end(), 0);
    B.dissect(all, 0);
  }
  const int nn = (int)T->nodes.size();
  // symbolic structure, children before parents (= storage order)
  T->node_of.assign(nT, -1); T->pos_in_sep.assign(nT, -1);
  std::vector<char> eliminated(nT, 0), seen(nT, 0);
  for (int q = 0; q < nn; q++) {
    NdNode &nd = T->nodes[q];
    T->n_levels = std::max(T->n_levels, nd.level + 1);
    for (size_t k = 0; k < nd.sep.size(); k++) {
      const int t = nd.sep[k];
      if (T->node_of[t] != -1) { ufe_set_error("ufe_nd_analyse: triangle %d eliminated twice", t); delete T; return UFE_ERR_INVALID; }
      T->node_of[t] = q; T->pos_in_sep[t] = (int)k; eliminated[t] = 1;
    }
    std::vector<int> cand;
    auto add = [&](int j) { if (!eliminated[j] && !seen[j]) { seen[j] = 1; cand.push_back(j); } };
    for (int t : nd.sep) for (int k = G.ptr[t]; k < G.ptr[t + 1]; k++) add(G.ind[k]);
    for (int c : nd.child) if (c >= 0) for (int j : T->nodes[c].bnd) add(j);
    std::sort(cand.begin(), cand.end());
    for (int j : cand) seen[j] = 0;
    nd.bnd = std::move(cand);
  }
  for (int t = 0; t < nT; t++) if (T->node_of[t] < 0) { ufe_set_error("ufe_nd_analyse: triangle %d not eliminated", t); delete T; return UFE_ERR_INVALID; }
  // extend-add maps: child's bnd -> position in the parent's [sep; bnd]
  std::vector<int> loc(nT, -1);
  for (int q = 0; q < nn; q++) {
    NdNode &nd = T->nodes[q];
    const int ns = (int)nd.sep.size();
    for (int k = 0; k < ns; k++) loc[nd.sep[k]] = k;
    for (size_t k = 0; k < nd.bnd.size(); k++) loc[nd.bnd[k]] = ns + (int)k;
    for (int c : nd.child) {
      if (c < 0) continue;
      NdNode &ch = T->nodes[c];
      ch.up.resize(ch.bnd.size());
      for (size_t k = 0; k < ch.bnd.size(); k++) {
        ch.up[k] = loc[ch.bnd[k]];
        if (ch.up[k] < 0) { ufe_set_error("ufe_nd_analyse: child boundary not contained in the parent's front"); delete T; return UFE_ERR_INVALID; }
      }
    }
    for (int t : nd.sep) loc[t] = -1;
    for (int t : nd.bnd) loc[t] = -1;
  }
  // assembly map: block entry (i,j) belongs to the front of whichever of i, j is eliminated first
  const int nnzb = bptr[nT];
  T->entry_node.resize(nnzb); T->entry_row.resize(nnzb); T->entry_col.resize(nnzb);
  std::atomic<int> bad_i{-1}, bad_j{-1};
  parallel_for(nT, [&](int i0, int i1) {
  for (int i = i0; i < i1; i++) {
    for (int k = bptr[i]; k < bptr[i + 1]; k++) {
      const int j = bind[k];
      const int q = std::min(T->node_of[i], T->node_of[j]);      // post-order index: smaller = eliminated earlier
      const NdNode &nd = T->nodes[q];
      const int ns = (int)nd.sep.size();
      auto where = [&](int t) -> int {
        if (T->node_of[t] == q) return T->pos_in_sep[t];
        const auto it = std::lower_bound(nd.bnd.begin(), nd.bnd.end(), t);
        return (it != nd.bnd.end() && *it == t) ? ns + (int)(it - nd.bnd.begin()) : -1;
      };
      const int r = where(i), c = where(j);
      if (r < 0 || c < 0) { bad_i = i; bad_j = j; return; }
      T->entry_node[k] = q; T->entry_row[k] = r; T->entry_col[k] = c;
    }
  }
  });
  if (bad_i >= 0) { ufe_set_error("ufe_nd_analyse: entry (%d,%d) outside its front", bad_i.load(), bad_j.load()); delete T; return UFE_ERR_INVALID; }
  *out = T;
  return UFE_OK;
}

extern "C" int ufe_nd_tree_info(const ufe_nd_tree *T, int32_t *n_nodes, int32_t *n_levels, int32_t *max_front,
                                double *padded_front_bytes) {
  if (!T) { ufe_set_error("null tree"); return UFE_ERR_INVALID; }
  std::vector<long long> lmax(T->n_levels, 0), lcnt(T->n_levels, 0);
  int mf = 0;
  for (const NdNode &nd : T->nodes) {
    const int f = 2 * (int)(nd.sep.size() + nd.bnd.size());
    mf = std::max(mf, f);
    lmax[nd.level] = std::max<long long>(lmax[nd.level], f); lcnt[nd.level]++;
  }
  double bytes = 0.0;
  for (int l = 0; l < T->n_levels; l++) bytes += 8.0 * (double)lcnt[l] * (double)lmax[l] * (double)lmax[l];
  if (n_nodes) *n_nodes = (int)T->nodes.size();
  if (n_levels) *n_levels = T->n_levels;
  if (max_front) *max_front = mf;
  if (padded_front_bytes) *padded_front_bytes = bytes;
  return UFE_OK;
}

// node i of the post-order (children before parents, root last).  The pointers stay valid until ufe_nd_tree_free.
extern "C" int ufe_nd_tree_node(const ufe_nd_tree *T, int32_t i, int32_t *level, int32_t *parent, int32_t *n_sep,
                                int32_t *n_bnd, const int32_t **sep, const int32_t **bnd, const int32_t **up) {
  if (!T || i < 0 || i >= (int)T->nodes.size()) { ufe_set_error("bad node index"); return UFE_ERR_INVALID; }
  const NdNode &nd = T->nodes[i];
  if (level) *level = nd.level;
  if (parent) *parent = nd.parent;
  if (n_sep) *n_sep = (int)nd.sep.size();
  if (n_bnd) *n_bnd = (int)nd.bnd.size();
  if (sep) *sep = nd.sep.data();
  if (bnd) *bnd = nd.bnd.data();
  if (up) *up = nd.up.data();
  return UFE_OK;
}

// per block entry of the analysed pattern: the front it is assembled into and its block row / column there
extern "C" int ufe_nd_tree_entry_map(const ufe_nd_tree *T, const int32_t **node, const int32_t **row, const int32_t **col) {
  if (!T) { ufe_set_error("null tree"); return UFE_ERR_INVALID; }
  if (node) *node = T->entry_node.data();
  if (row) *row = T->entry_row.data();
  if (col) *col = T->entry_col.data();
  return UFE_OK;
}

// Owner rank of every tree node when the factorisation is distributed over `nranks` (a power of two) ranks: the
// sub-tree below the r-th node of level log2(nranks) belongs to rank r, the nodes above it to the left spine of their
// sub-tree (root: rank 0; level 1: ranks 0, nranks / 2; ...).  owner: n_nodes entries, post-order like ufe_nd_tree_node.
int ufe_nd_owner_map(const ufe_nd_tree *T, int nranks, std::vector<int> &owner, std::vector<int> &span) {
  const int nn = (int)T->nodes.size();
  if (nranks < 1 || (nranks & (nranks - 1)) != 0) { ufe_set_error("nd_lu: the number of ranks must be a power of two (got %d)", nranks); return UFE_ERR_INVALID; }
  owner.assign(nn, 0); span.assign(nn, 1);
  for (int q = nn - 1; q >= 0; q--) {            // post-order: parents have larger indices than their children
    const NdNode &nd = T->nodes[q];
    if (nd.parent < 0) { owner[q] = 0; span[q] = nranks; }
    if (nd.child[0] < 0) { if (span[q] > 1) { ufe_set_error("nd_lu: the elimination tree is too shallow for %d ranks", nranks); return UFE_ERR_INVALID; } continue; }
    const int half = span[q] / 2;
    owner[nd.child[0]] = owner[q]; span[nd.child[0]] = std::max(1, half);
    owner[nd.child[1]] = owner[q] + half; span[nd.child[1]] = std::max(1, half);
  }
  return UFE_OK;
}
extern "C" int ufe_nd_tree_owners(const ufe_nd_tree *T, int32_t nranks, int32_t *owner) {
  if (!T || !owner) { ufe_set_error("ufe_nd_tree_owners: null argument"); return UFE_ERR_INVALID; }
  std::vector<int> o, sp;
  UFE_TRY(ufe_nd_owner_map(T, nranks, o, sp));
  for (size_t q = 0; q < o.size(); q++) owner[q] = o[q];
  return UFE_OK;
}

extern "C" void ufe_nd_tree_free(ufe_nd_tree *T) { delete T; }
