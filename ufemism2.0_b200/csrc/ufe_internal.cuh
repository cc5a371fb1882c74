// Internal declarations shared by the translation units of libufe_diva.so.
// sm_100a only; fp64 throughout (the reference is real(dp) everywhere).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ufe_diva.h"

#define UFE_NZ_MAX 32
#define UFE_STACK_MAX 64        // local BFS stack capacity in the operator-construction kernels
#define UFE_RED_BLOCKS 592      // 4 x 148 SMs: fixed grid for reductions (deterministic sums)
#define UFE_RED_THREADS 256

void ufe_set_error(const char *fmt, ...);
extern thread_local int64_t g_launch_count;

#define UFE_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ufe_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                    cudaGetErrorString(e_));                                            \
      return UFE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define UFE_NCCL(call)                                                                  \
  do {                                                                                  \
    ncclResult_t r_ = (call);                                                           \
    if (r_ != ncclSuccess) {                                                            \
      ufe_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(r_)); \
      return UFE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define UFE_TRY(call)                 \
  do {                                \
    int rc_ = (call);                 \
    if (rc_ != UFE_OK) return rc_;    \
  } while (0)

#define UFE_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    g_launch_count++;                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      ufe_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                    cudaGetErrorString(e_));                                            \
      return UFE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

static inline int ufe_div_up(long long a, int b) { return (int)((a + b - 1) / b); }

// One operator family on the device: rows i1..i1+m_loc-1 (1-based global), shared
// pattern, nval value arrays.  ptr: local 1-based offsets; ind: global 1-based columns.
struct DevFamily {
  int m_loc = 0, m = 0, n = 0, i1 = 1, nnz = 0, nval = 0;
  int *ptr = nullptr, *ind = nullptr;
  double *val[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

// device mesh (full topology replicated per GPU, as the reference replicates it per rank)
struct DevMesh {
  int nV = 0, nTri = 0, nC_mem = 0, nz = 0;
  double xmin = 0, xmax = 0, ymin = 0, ymax = 0;
  double *V = nullptr, *TriGC = nullptr, *zeta = nullptr;
  int *Tri = nullptr, *TriC = nullptr, *C = nullptr, *nC = nullptr, *iTri = nullptr, *niTri = nullptr,
      *VBI = nullptr, *TriBI = nullptr;
};

struct PeerDev;
// Krylov scalars living in device memory
struct KrylovScalars {
  double dots[8];       // raw (all-reduced) dot products of the current stage
  double rho, alpha, omega, beta;
  double bnorm, ttol, rnorm, dtol_bnorm, abstol;
  int its, done, reason, maxits;   // reason: 2 rtol, 3 atol, -3 maxits, -4 dtol, -5 breakdown
  int jcount, finalized, pad0, pad1; // GMRES: steps completed in the current cycle; x updated after convergence
  const struct PeerDev *peer;        // several ranks with peer-memory communication, else nullptr
};

struct KrylovWork {
  int n_loc = 0;          // owned unknowns
  double *r = nullptr, *rhat = nullptr, *p = nullptr, *v = nullptr, *s = nullptr, *t = nullptr;  // owned-length
  double *pg = nullptr, *sg = nullptr;   // full-length, global-indexed SpMV inputs (owned part + halo valid)
  int N = 0;
  double *Vb = nullptr;   // GMRES basis (restart+1) x n_loc
  double *w = nullptr;
  double *partials = nullptr;   // UFE_RED_BLOCKS * 32
  double *dots_local = nullptr; // 40 doubles (pre-allreduce)
  unsigned *counter = nullptr;
  KrylovScalars *sc = nullptr;
  double *gm = nullptr;   // GMRES small dense state: H(31x30), cs, sn, g, y, hcol
  KrylovScalars *sc_host = nullptr;  // pinned mirror
  double *pctmp = nullptr, *bP = nullptr;   // explicit-preconditioner work vectors (owned length)
  bool ext_vecs = false;                    // pg / sg live in the symmetric peer buffer (not freed here)
  PeerDev *peer_dev = nullptr;              // device copy of the peer description
};

#define UFE_MAX_RANKS 8
#define UFE_PEER_DOTS 48        // doubles per rank per reduction slot

// Peer-memory view of the other ranks' Krylov vectors (NVLink P2P through CUDA IPC): every rank
// allocates one symmetric buffer  [pg (N) | sg (N) | x (N) | dots 2 x P x 48 | flags]  and maps
// the others'.  SpMV kernels read halo entries of the input vector straight from the owner's
// buffer; reductions are all-to-all stores into the peers' dot slots followed by a fixed-order
// sum, so no NCCL call is left inside a Krylov iteration.
struct PeerComm {
  int on = 0, P = 1, me = 0;
  double *base[UFE_MAX_RANKS] = {};     // mapped symmetric buffers (base[me] = own)
  long long off_pg = 0, off_sg = 0, off_x = 0, off_dots = 0, off_flags = 0;   // offsets in doubles
  int bounds[UFE_MAX_RANKS + 1] = {};   // triangle ownership: rank q owns [bounds[q], bounds[q+1])
  long long halo_epoch = 0, red_epoch = 0;   // host-side counters (identical on every rank)
};

// device-resident description of the peers (pointed to by KrylovScalars::peer)
struct PeerDev { int P, me; int *flags[UFE_MAX_RANKS]; double *dots[UFE_MAX_RANKS]; };
struct PeerFlagPtrs { int *f[UFE_MAX_RANKS]; };      // every rank's flag array (P2P mapped)
struct PeerDotPtrs { double *d[UFE_MAX_RANKS]; };    // every rank's dot-slot array (P2P mapped)

struct Comm {
  int rank = 0, nranks = 1;
  ncclComm_t nccl = nullptr;
  mutable PeerComm peer;
};

// A distributed square CSR system on the device (rows r1..r1+m_loc-1, 1-based global).
struct DevSystem {
  int N = 0, m_loc = 0, r1 = 1, nnz = 0;
  int *ptr = nullptr, *ind = nullptr;
  double *val = nullptr;     // as assembled (reference values)
  double *valS = nullptr;    // left-preconditioner folded in: B*A
  double *bb = nullptr, *bS = nullptr;
  double *x = nullptr;       // full-length (N) solution / SpMV input vector (global indexing)
  int jmin = 1, jmax = 0;    // column range touched by the owned rows (calc_j_node_range)
  // Krylov-loop copy of B*A for the DIVA/SSA stiffness matrix: 2x2 (u,v) blocks per triangle
  // pair in sliced ELL (slice = 32 block rows = one warp).  Entry e of block row s*32+l:
  //   bell_col[(off_s+e)*32 + l]            0-based global triangle index of the block column
  //   bell_val[(off_s+e)*128 + c*32 + l]    c = 0..3 -> a_uu, a_uv, a_vu, a_vv
  // off_s = bell_off[s] (entries), slice width = bell_off[s+1]-bell_off[s]; padding entries
  // point at the row's own triangle with zero values.  nullptr => CSR valS is used instead.
  int nslices = 0;
  long long bell_entries = 0;
  int *bell_off = nullptr, *bell_col = nullptr;
  double *bell_val = nullptr;
};

int ufe_spmv_launch(cudaStream_t st, int m_loc, int nnz, const int *ptr, const int *ind,
                    const double *val, const double *x, long long ldx, double *y, long long ldy,
                    int nlayers);

int ufe_krylov_alloc(KrylovWork &kw, int N, int n_loc, bool gmres, double *ext_pg = nullptr, double *ext_sg = nullptr);
void ufe_krylov_free(KrylovWork &kw);
// solves valS * x = bS on the owned rows; x full-length global vector (x[r1-1 ..] owned).
// halo: callback-free -- single GPU handled inline, multi-GPU through the exchange plan.
struct HaloPlan;
struct PcLU;     // block-Jacobi / block-tridiagonal LU preconditioner (ufe_pclu.cu); nullptr = none
int ufe_pclu_setup(cudaStream_t st, const DevSystem &S, int replicate, size_t max_bytes, PcLU **out,
                   const Comm *comm = nullptr, const HaloPlan *gather_all = nullptr);
int ufe_pclu_factor(cudaStream_t st, const DevSystem &S, PcLU *pc);
int ufe_pclu_apply(cudaStream_t st, PcLU *pc, const double *r, double *z);
void ufe_pclu_free(PcLU *pc);
bool ufe_pclu_exact_and_fresh(const PcLU *pc);   // whole-matrix factorisation of the current values: M^-1 b is the solution up to round-off
void ufe_pclu_mark_reused(PcLU *pc);
// multifrontal nested-dissection solver as the exact preconditioner (ufe_nd_numeric.cu; krylov_pc = UFE_PC_ND_LU)
int ufe_pclu_setup_nd(cudaStream_t st, const DevSystem &S, const Comm *comm, int nT, const double *gcx, const double *gcy, PcLU **out);
int ufe_nd_pc_create(cudaStream_t st, const DevSystem &S, const Comm *comm, int nT, const double *gcx, const double *gcy, int leaf, ufe_nd_solver **out);
void ufe_nd_pc_info(const ufe_nd_solver *S, double *front_bytes, double *flops, int *n_fronts);
void ufe_nd_pc_set_point_scaling(ufe_nd_solver *S);     // the Krylov operator is diag(A)^-1 A (generic L0 systems) instead of the 2x2-block scaling
void ufe_pclu_set_point_scaling(PcLU *pc);
int ufe_launch_scale_generic(cudaStream_t st, const DevSystem &S);
int ufe_nd_pc_factor(cudaStream_t st, ufe_nd_solver *S, const double *dval);
int ufe_nd_pc_apply(cudaStream_t st, ufe_nd_solver *S, const double *r, double *z);
int ufe_krylov_run(cudaStream_t st, const DevSystem &S, KrylovWork &kw, const Comm &comm,
                   const HaloPlan *halo, int method, double rtol, double abstol, int maxits,
                   int guess_nonzero, int *n_its, int *flags, PcLU *pc = nullptr);

// halo exchange plan for a global-indexed vector partitioned into contiguous ranges
struct HaloPlan {
  int nranks = 1, rank = 0;
  std::vector<int> own_lo, own_hi;     // 0-based [lo,hi) owned by each rank
  std::vector<int> need_lo, need_hi;   // 0-based [lo,hi) each rank needs (own + halo)
  int mult = 2;                        // vector entries per index (2: the (u,v) pair of a triangle; 1: generic systems)
};
int ufe_halo_exchange(cudaStream_t st, const Comm &comm, const HaloPlan &plan, double *x, long long ld,
                      int nlayers, int mult);
