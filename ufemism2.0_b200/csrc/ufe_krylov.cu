// Hand-written Krylov solvers replacing the PETSc KSP call
// (src/UPSY/basic/petsc_basic.f90:66-141, KSPSolve isolated at :143-168).
//
// The left preconditioner B (point Jacobi or 2x2 u-v block Jacobi) is folded into the
// matrix at assembly time (valS = B*A, bS = B*b), so the Krylov loop is unpreconditioned
// on the scaled system and its residual norm IS PETSc's default preconditioned norm:
//     stop when ||B r|| <= max(rtol*||B b||, abstol)      (KSPConvergedDefault)
//     diverged when ||B r|| >= dtol*||B b||, dtol = 1e5;   zero initial guess by default.
//
// Everything the iteration needs (dot products, alpha/omega/beta, the convergence test,
// the iteration counter) lives in device memory; the host only enqueues batches of
// iterations and polls a pinned copy of the scalars between batches.  After the device
// has set `done`, the remaining kernels of a batch return immediately, so the solution
// and the reported iteration count are exactly those of the converged iteration.
// Dot products are reduced in a fixed order (per-block partials summed by the last block)
// => bitwise reproducible run to run.  Multi-GPU (PeerComm, ufe_internal.cuh): the SpMV reads the
// halo entries of its input vector in place from the owners' IPC-mapped buffers, and the last
// block of every reducing kernel exchanges the partial sums with all peers by P2P stores and sums
// them in rank order -- no collective launch inside an iteration.  Without peer access the same
// loop runs over NCCL (ncclSend/Recv halo ranges, one ncclAllReduce per reduction stage).
// An explicit left preconditioner (ufe_pclu.cu) can be plugged in: the operator becomes
// M^-1 A and BiCGStab gains a half-step exit.
#include "ufe_internal.cuh"
#include "ufe_reduce.cuh"

#define KGRID 1184   // 8 x 148 persistent blocks for the fused SpMV kernels
#define GM_RESTART 30

// ---------------------------------------------------------------------------------
// scalar recurrences, run by exactly one thread after each reduction stage
// ---------------------------------------------------------------------------------
// gm layout: H[31*30] | cs[30] | sn[30] | g[31] | y[30] | hcol[32] | inv_hn
#define GM_H(gm) (gm)
#define GM_CS(gm) ((gm) + 31 * 30)
#define GM_SN(gm) ((gm) + 31 * 30 + 30)
#define GM_G(gm) ((gm) + 31 * 30 + 60)
#define GM_Y(gm) ((gm) + 31 * 30 + 60 + 31)
#define GM_HCOL(gm) ((gm) + 31 * 30 + 60 + 31 + 30)
#define GM_INVHN(gm) ((gm) + 31 * 30 + 60 + 31 + 30 + 32)
#define GM_SIZE (31 * 30 + 60 + 31 + 30 + 32 + 1)

enum { ST_INIT = 0, ST_A = 1, ST_B = 2, ST_C = 3, ST_GM_INIT = 4, ST_GM_H = 5, ST_GM_NORM = 6, ST_S = 7, ST_RICH = 8 };

__device__ void gmres_givens(KrylovScalars *sc, double *gm, double hn);

__device__ void post_reduce(int stage, KrylovScalars *sc, double *gm, double rtol, double abstol) {
  switch (stage) {
    case ST_INIT: {   // dots: (r,r), (b,b)
      sc->bnorm = sqrt(sc->dots[1]);
      sc->ttol = fmax(rtol * sc->bnorm, abstol);
      sc->dtol_bnorm = 1.0e5 * sc->bnorm;
      sc->rnorm = sqrt(sc->dots[0]);
      sc->rho = sc->dots[0];
      sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
      if (sc->rnorm <= sc->ttol) { sc->done = 1; sc->reason = (sc->rnorm <= abstol) ? 3 : 2; }
      break;
    }
    case ST_RICH: {   // dots: (r,r), (b,b) after the Richardson step x = M^-1 b with an exact M: one iteration if it passes
      sc->bnorm = sqrt(sc->dots[1]);
      sc->ttol = fmax(rtol * sc->bnorm, abstol);
      sc->rnorm = sqrt(sc->dots[0]);
      if (sc->rnorm <= sc->ttol) { sc->its = 1; sc->done = 1; sc->reason = (sc->rnorm <= abstol) ? 3 : 2; }
      break;
    }
    case ST_A: {      // dots: (rhat, v)
      const double rv = sc->dots[0];
      if (rv == 0.0) { sc->done = 1; sc->reason = -5; break; }
      sc->alpha = sc->rho / rv;
      break;
    }
    case ST_S: {      // dots: (s,s) -- half-step exit: x + alpha p already satisfies the stopping rule
      const double snorm = sqrt(sc->dots[0]);
      if (snorm <= sc->ttol) {
        sc->rnorm = snorm; sc->its += 1; sc->pad0 = 1; sc->done = 1; sc->reason = (snorm <= abstol) ? 3 : 2;
      }
      break;
    }
    case ST_B: {      // dots: (t,s), (t,t)
      const double tt = sc->dots[1];
      sc->omega = (tt == 0.0) ? 0.0 : sc->dots[0] / tt;
      break;
    }
    case ST_C: {      // dots: (r,r), (rhat,r)
      sc->rnorm = sqrt(sc->dots[0]);
      sc->its += 1;
      if (!(sc->rnorm == sc->rnorm)) { sc->done = 1; sc->reason = -5; break; }
      if (sc->rnorm <= sc->ttol) { sc->done = 1; sc->reason = (sc->rnorm <= abstol) ? 3 : 2; break; }
      if (sc->rnorm >= sc->dtol_bnorm) { sc->done = 1; sc->reason = -4; break; }
      if (sc->its >= sc->maxits) { sc->done = 1; sc->reason = -3; break; }
      const double rho_new = sc->dots[1];
      if (sc->omega == 0.0 || sc->rho == 0.0 || rho_new == 0.0) { sc->done = 1; sc->reason = -5; break; }
      sc->beta = (rho_new / sc->rho) * (sc->alpha / sc->omega);
      sc->rho = rho_new;
      break;
    }
    case ST_GM_INIT: {  // dots: (r,r)[, (b,b) on the first cycle -> dots[1] >= 0]
      if (sc->bnorm < 0.0) {
        sc->bnorm = sqrt(sc->dots[1]);
        sc->ttol = fmax(rtol * sc->bnorm, abstol);
        sc->dtol_bnorm = 1.0e5 * sc->bnorm;
      }
      sc->rnorm = sqrt(sc->dots[0]);
      sc->jcount = 0;
      for (int i = 0; i <= GM_RESTART; i++) GM_G(gm)[i] = 0.0;
      GM_G(gm)[0] = sc->rnorm;
      *GM_INVHN(gm) = (sc->rnorm > 0.0) ? 1.0 / sc->rnorm : 0.0;
      if (sc->rnorm <= sc->ttol) { sc->done = 1; sc->reason = (sc->rnorm <= abstol) ? 3 : 2; sc->finalized = 1; }
      break;
    }
    case ST_GM_NORM: {  // dots: (w,w)
      gmres_givens(sc, gm, sqrt(sc->dots[0]));
      break;
    }
  }
}

__device__ void gmres_givens(KrylovScalars *sc, double *gm, double hn) {
  const int j = sc->jcount;
  double *hcol = GM_HCOL(gm), *cs = GM_CS(gm), *sn = GM_SN(gm), *g = GM_G(gm);
  hcol[j + 1] = hn;
  *GM_INVHN(gm) = (hn != 0.0) ? 1.0 / hn : 0.0;
  for (int i = 0; i < j; i++) {
    const double a = hcol[i], c = hcol[i + 1];
    hcol[i] = cs[i] * a + sn[i] * c;
    hcol[i + 1] = -sn[i] * a + cs[i] * c;
  }
  const double a = hcol[j], c = hcol[j + 1], rr = hypot(a, c);
  if (rr == 0.0) { cs[j] = 1.0; sn[j] = 0.0; } else { cs[j] = a / rr; sn[j] = c / rr; }
  hcol[j] = rr; hcol[j + 1] = 0.0;
  g[j + 1] = -sn[j] * g[j];
  g[j] = cs[j] * g[j];
  for (int i = 0; i <= j; i++) GM_H(gm)[i + 31 * j] = hcol[i];
  sc->rnorm = fabs(g[j + 1]);
  sc->its += 1;
  sc->jcount = j + 1;
  if (!(sc->rnorm == sc->rnorm)) { sc->done = 1; sc->reason = -5; }
  else if (sc->rnorm <= sc->ttol) { sc->done = 1; sc->reason = (sc->rnorm <= sc->abstol) ? 3 : 2; }
  else if (sc->rnorm >= sc->dtol_bnorm) { sc->done = 1; sc->reason = -4; }
  else if (sc->its >= sc->maxits) { sc->done = 1; sc->reason = -3; }
  else if (hn == 0.0) { sc->done = 1; sc->reason = 2; }
}

__global__ void k_post(int stage, KrylovScalars *sc, double *gm, const double *dots_in, int nd, double rtol,
                       double abstol) {
  if (sc->done) return;
  for (int i = 0; i < nd; i++) sc->dots[i] = dots_in[i];
  post_reduce(stage, sc, gm, rtol, abstol);
}

// copy the (all-reduced) Gram-Schmidt coefficients of step j into the Hessenberg column
__global__ void k_gm_sethcol(const KrylovScalars *sc, double *gm, const double *dots_in, int cnt) {
  if (sc->done) return;
  for (int i = 0; i < cnt; i++) GM_HCOL(gm)[i] = dots_in[i];
}

// finish a reduction stage inside the producing kernel (single GPU) or leave the local
// partial sums in dots_local for the all-reduce (multi GPU)
// Bounded spin on a flag another GPU writes: a rank that died must not hang the others' GPUs.
// ~2^26 polls of a local volatile word (tens of seconds); returns false on time-out.
// time-out in wall time (%globaltimer, ns), not in polls: 20 s without the peer's flag means the peer is gone
__device__ __forceinline__ bool peer_wait(const volatile int *flag, int epoch) {
  unsigned long long t0 = 0;
  for (long long spins = 0;; spins++) {
    if (*flag >= epoch) return true;
    if ((spins & 1023) == 1023) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ULL) return false;
    }
  }
}

// `single`:  > 0 one rank: finish the stage here;  == 0 several ranks over NCCL: leave the local
// partial sums in dots_local for the all-reduce;  < 0 several ranks with peer memory, epoch =
// -single: the last block stores its partial sums into every rank's slot (NVLink P2P), waits
// for everybody's, sums them in rank order and runs the recurrence -- the collective is part of
// the producing kernel, there is no separate reduction launch.
template <int NV>
__device__ __forceinline__ void finish_stage(double (&acc)[NV], int stage, double *partials, unsigned *counter,
                                             double *dots_local, KrylovScalars *sc, double *gm, int single,
                                             double rtol, double abstol) {
  if (reduce_publish<NV>(acc, partials, counter, dots_local)) {
    if (single > 0) {
      if (threadIdx.x == 0) {
        for (int i = 0; i < NV; i++) sc->dots[i] = dots_local[i];
        post_reduce(stage, sc, gm, rtol, abstol);
      }
    } else if (single < 0) {
      const PeerDev *pd = sc->peer;
      const int epoch = -single, par = epoch & 1, P = pd->P, me = pd->me, t = threadIdx.x;
      if (t < P) {
        double *dst = pd->dots[t] + ((size_t)par * P + me) * UFE_PEER_DOTS;
        for (int i = 0; i < NV; i++) dst[i] = dots_local[i];
        __threadfence_system();
        *reinterpret_cast<volatile int *>(pd->flags[t] + P + par * P + me) = epoch;
        const volatile int *mine = reinterpret_cast<const volatile int *>(pd->flags[me] + P + par * P + t);
        if (!peer_wait(mine, epoch)) { sc->reason = -9; sc->done = 1; }
      }
      __syncthreads();
      if (t == 0 && sc->reason != -9) {
        __threadfence_system();
        const volatile double *src = pd->dots[me] + (size_t)par * P * UFE_PEER_DOTS;
        for (int i = 0; i < NV; i++) { double sum = 0.0; for (int q = 0; q < P; q++) sum += src[(size_t)q * UFE_PEER_DOTS + i]; sc->dots[i] = sum; }
        post_reduce(stage, sc, gm, rtol, abstol);
      }
    }
  }
}

// tail of a kernel that produces an SpMV input vector without reducing anything: the last block
// to finish publishes `epoch` to every peer (replaces a separate signal launch); epoch 0 = off
__device__ __forceinline__ void signal_peers_when_done(unsigned *counter, const KrylovScalars *sc, int epoch) {
  if (epoch <= 0) return;
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (!last) return;
  const PeerDev *pd = sc->peer;
  if ((int)threadIdx.x < pd->P && (int)threadIdx.x != pd->me) {
    __threadfence_system();
    *reinterpret_cast<volatile int *>(pd->flags[threadIdx.x] + pd->me) = epoch;
  }
}

// ---------------------------------------------------------------------------------
// fused SpMV (+ dot products) on the scaled matrix, T threads per row, persistent grid.
// MODE 0: y = A xg
// MODE 1: y = A xg ; dot0 = (z, y)                      [BiCGStab: v = A p, (rhat, v)]
// MODE 2: y = A xg ; dot0 = (y, xg_own), dot1 = (y, y)   [BiCGStab: t = A s, (t,s), (t,t)]
// ---------------------------------------------------------------------------------
template <int T, int MODE>
__global__ void __launch_bounds__(256)
k_kspmv(int m_loc, int r0, const int *__restrict__ ptr, const int *__restrict__ ind,
        const double *__restrict__ val, const double *__restrict__ xg, double *__restrict__ y,
        const double *__restrict__ z, int stage, double *partials, unsigned *counter, double *dots_local,
        KrylovScalars *sc, double *gm, int single, double rtol, double abstol) {
  if (sc->done) return;
  const int lane = threadIdx.x % T;
  const int rows_per_block = 256 / T;
  double acc[2] = {0.0, 0.0};
  for (int base = blockIdx.x * rows_per_block; base < m_loc; base += gridDim.x * rows_per_block) {
    const int row = base + threadIdx.x / T;
    double s = 0.0;
    if (row < m_loc) {
      const int k0 = ptr[row] - 1, k1 = ptr[row + 1] - 1;
      for (int k = k0 + lane; k < k1; k += T) s += __ldg(val + k) * __ldg(xg + (__ldg(ind + k) - 1));
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, T);
    if (row < m_loc && lane == 0) {
      y[row] = s;
      if (MODE == 1) acc[0] += z[row] * s;
      if (MODE == 2) { acc[0] += s * xg[r0 + row]; acc[1] += s * s; }
    }
  }
  if (MODE == 1) { double a1[1] = {acc[0]}; finish_stage<1>(a1, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol); }
  if (MODE == 2) finish_stage<2>(acc, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol);
}

// ---------------------------------------------------------------------------------
// Peer-memory communication inside the Krylov loop (PeerComm, ufe_internal.cuh)
// ---------------------------------------------------------------------------------
struct PeerView {             // by-value kernel argument
  int P, me, epoch;           // epoch: halo signal count the consumer must see from every rank
  int bounds[UFE_MAX_RANKS + 1];
  const double2 *xp[UFE_MAX_RANKS];   // the input vector of this SpMV in every rank's buffer
  const volatile int *flags;          // own flag array: [0, P) halo epochs, [P, 3P) reduction epochs (2 parities)
};

__device__ __forceinline__ double2 peer_load(const PeerView &pv, int c) {
  int q = 0;
#pragma unroll
  for (int k = 1; k < UFE_MAX_RANKS; k++) if (k < pv.P && c >= pv.bounds[k]) q = k;
  const double2 *p = pv.xp[q] + c;
  double2 v;
  asm volatile("ld.global.cv.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));   // never a stale cached line
  return v;
}

// every rank tells every other rank that its SpMV input vector is complete (P2P store of the epoch)
__global__ void k_peer_signal(int P, int me, int epoch, PeerFlagPtrs fp) {
  const int q = threadIdx.x;
  if (q >= P || q == me) return;
  __threadfence_system();
  *reinterpret_cast<volatile int *>(fp.f[q] + me) = epoch;
}

// all-to-all of the local partial dot products + fixed-order sum + scalar recurrence of the stage.
// stage < 0: GMRES Gram-Schmidt coefficients -> hcol (no recurrence).
__global__ void k_peer_reduce(int P, int me, int epoch, int nd, int stage, PeerFlagPtrs fp, PeerDotPtrs dp,
                              const double *__restrict__ dots_local, KrylovScalars *sc, double *gm, double rtol, double abstol) {
  if (sc->done) return;
  const int par = epoch & 1, t = threadIdx.x;
  if (t < P) {
    double *dst = dp.d[t] + ((size_t)par * P + me) * UFE_PEER_DOTS;
    for (int i = 0; i < nd; i++) dst[i] = dots_local[i];
    __threadfence_system();
    *reinterpret_cast<volatile int *>(fp.f[t] + P + par * P + me) = epoch;
  }
  if (t < P) {
    const volatile int *mine = reinterpret_cast<const volatile int *>(fp.f[me] + P + par * P + t);
    if (!peer_wait(mine, epoch)) { sc->reason = -9; sc->done = 1; }
  }
  __syncthreads();
  if (t == 0 && sc->reason != -9) {
    __threadfence_system();
    const volatile double *src = dp.d[me] + (size_t)par * P * UFE_PEER_DOTS;
    if (stage < 0) {
      for (int i = 0; i < nd; i++) { double sum = 0.0; for (int q = 0; q < P; q++) sum += src[(size_t)q * UFE_PEER_DOTS + i]; GM_HCOL(gm)[i] = sum; }
    } else {
      for (int i = 0; i < nd; i++) { double sum = 0.0; for (int q = 0; q < P; q++) sum += src[(size_t)q * UFE_PEER_DOTS + i]; sc->dots[i] = sum; }
      post_reduce(stage, sc, gm, rtol, abstol);
    }
  }
}

// ---------------------------------------------------------------------------------
// The same three modes on the blocked sliced-ELL copy of the DIVA/SSA stiffness matrix
// (2x2 u-v blocks per triangle pair, one warp per slice of 32 block rows, layout in
// DevSystem).  One thread owns a block row = matrix rows (2t, 2t+1): per block entry it
// loads one column index (coalesced, 128 B per warp), four values (4 x 256 B per warp)
// and gathers x as one 16-byte (u,v) pair; y is written as 16-byte pairs.  36 B per block
// instead of 48 B in CSR, no shuffles, loads of U consecutive entries are issued together.
// ---------------------------------------------------------------------------------
template <int MODE, int U, int PEER>
__global__ void __launch_bounds__(256)
k_kspmv_bell(int nt_loc, int t0, int nslices, const int *__restrict__ bell_off, const int *__restrict__ bcol,
             const double *__restrict__ bval, const double *__restrict__ xg, double *__restrict__ y,
             const double *__restrict__ z, int stage, double *partials, unsigned *counter, double *dots_local,
             KrylovScalars *sc, double *gm, int single, double rtol, double abstol, PeerView peer) {
  if (sc->done) return;
  if (PEER) {       // wait until every rank has published this input vector (flags are in own memory)
    if ((int)threadIdx.x < peer.P && (int)threadIdx.x != peer.me && !peer_wait(peer.flags + threadIdx.x, peer.epoch)) {
      sc->reason = -9; sc->done = 1;      // peer time-out: stop the solve, the host reports it
    }
    __syncthreads();
  }
  const int own_lo = PEER ? peer.bounds[peer.me] : 0, own_hi = PEER ? peer.bounds[peer.me + 1] : 0x7fffffff;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const double2 *__restrict__ x2 = reinterpret_cast<const double2 *>(xg);
  double acc[2] = {0.0, 0.0};
  for (int s = warp; s < nslices; s += nwarps) {
    const int off = __ldg(bell_off + s) - 1, w = __ldg(bell_off + s + 1) - 1 - off;
    const int *__restrict__ pc = bcol + (size_t)off * 32 + lane;
    const double *__restrict__ pv = bval + (size_t)off * 128 + lane;
    double yu = 0.0, yv = 0.0;
    int e = 0;
    for (; e + U <= w; e += U) {
      int c[U]; double a00[U], a01[U], a10[U], a11[U]; double2 xx[U];
#pragma unroll
      for (int q = 0; q < U; q++) {
        c[q] = __ldg(pc + (size_t)(e + q) * 32);
        const double *p = pv + (size_t)(e + q) * 128;
        a00[q] = __ldg(p); a01[q] = __ldg(p + 32); a10[q] = __ldg(p + 64); a11[q] = __ldg(p + 96);
      }
#pragma unroll
      for (int q = 0; q < U; q++) xx[q] = (!PEER || (c[q] >= own_lo && c[q] < own_hi)) ? __ldg(x2 + c[q]) : peer_load(peer, c[q]);
#pragma unroll
      for (int q = 0; q < U; q++) { yu += a00[q] * xx[q].x + a01[q] * xx[q].y; yv += a10[q] * xx[q].x + a11[q] * xx[q].y; }
    }
    for (; e < w; e++) {
      const int c = __ldg(pc + (size_t)e * 32);
      const double *p = pv + (size_t)e * 128;
      const double a00 = __ldg(p), a01 = __ldg(p + 32), a10 = __ldg(p + 64), a11 = __ldg(p + 96);
      const double2 xx = (!PEER || (c >= own_lo && c < own_hi)) ? __ldg(x2 + c) : peer_load(peer, c);
      yu += a00 * xx.x + a01 * xx.y; yv += a10 * xx.x + a11 * xx.y;
    }
    const int r = s * 32 + lane;
    if (r < nt_loc) {
      reinterpret_cast<double2 *>(y)[r] = make_double2(yu, yv);
      if (MODE == 1) { const double2 zz = reinterpret_cast<const double2 *>(z)[r]; acc[0] += zz.x * yu + zz.y * yv; }
      if (MODE == 2) { const double2 xo = x2[t0 + r]; acc[0] += yu * xo.x + yv * xo.y; acc[1] += yu * yu + yv * yv; }
    }
  }
  if (MODE == 1) { double a1[1] = {acc[0]}; finish_stage<1>(a1, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol); }
  if (MODE == 2) finish_stage<2>(acc, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol);
}

template <int MODE>
static int launch_kspmv(cudaStream_t st, const DevSystem &S, const double *xg, double *y, const double *z,
                        int stage, KrylovWork &kw, int single, double rtol, double abstol, const PeerView *pv = nullptr) {
  if (S.bell_val) {
    int blocks = ufe_div_up(S.nslices, 8);
    if (blocks > KGRID) blocks = KGRID;
    if (blocks < 1) blocks = 1;
    if (pv)
      k_kspmv_bell<MODE, 4, 1><<<blocks, 256, 0, st>>>(S.m_loc / 2, (S.r1 - 1) / 2, S.nslices, S.bell_off, S.bell_col, S.bell_val,
                                                       xg, y, z, stage, kw.partials, kw.counter, kw.dots_local, kw.sc, kw.gm,
                                                       single, rtol, abstol, *pv);
    else
      k_kspmv_bell<MODE, 4, 0><<<blocks, 256, 0, st>>>(S.m_loc / 2, (S.r1 - 1) / 2, S.nslices, S.bell_off, S.bell_col, S.bell_val,
                                                       xg, y, z, stage, kw.partials, kw.counter, kw.dots_local, kw.sc, kw.gm,
                                                       single, rtol, abstol, PeerView());
    UFE_LAUNCH_CHECK();
    return UFE_OK;
  }
  const double mean = S.m_loc > 0 ? (double)S.nnz / S.m_loc : 1.0;
  int T = 1;
  while (T < 32 && T * 4 < mean) T *= 2;
  const int rpb = 256 / T;
  int blocks = ufe_div_up(S.m_loc, rpb);
  if (blocks > KGRID) blocks = KGRID;
  if (blocks < 1) blocks = 1;
#define KS_ARGS S.m_loc, S.r1 - 1, S.ptr, S.ind, S.valS, xg, y, z, stage, kw.partials, kw.counter, kw.dots_local, kw.sc, kw.gm, single, rtol, abstol
  switch (T) {
    case 1: k_kspmv<1, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
    case 2: k_kspmv<2, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
    case 4: k_kspmv<4, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
    case 8: k_kspmv<8, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
    case 16: k_kspmv<16, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
    default: k_kspmv<32, MODE><<<blocks, 256, 0, st>>>(KS_ARGS); break;
  }
#undef KS_ARGS
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

__global__ void k_sc_reset(KrylovScalars *sc, int maxits, double abstol);
// plain y = A x outside a solve: clears the `done` flag a finished solve leaves behind
// (every Krylov kernel returns at once when it is set)
int ufe_kspmv_plain(cudaStream_t st, const DevSystem &S, const double *xg, double *y, KrylovWork &kw) {
  k_sc_reset<<<1, 1, 0, st>>>(kw.sc, 1, 0.0);
  UFE_LAUNCH_CHECK();
  return launch_kspmv<0>(st, S, xg, y, nullptr, 0, kw, 1, 0.0, 0.0);
}
int ufe_kspmv_only(cudaStream_t st, const DevSystem &S, const double *xg, double *y, KrylovWork &kw) {
  return launch_kspmv<0>(st, S, xg, y, nullptr, 0, kw, 1, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------
// Explicit left preconditioner (block-Jacobi / LU, ufe_pclu.cu): the operator of the Krylov
// loop becomes y = M^-1 (A x); the dot products the fused SpMV modes would have produced
// are taken afterwards by k_stage_dots.
// ---------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_stage_dots(int n, int r0, const double *__restrict__ y, const double *__restrict__ z, const double *__restrict__ xg,
             int stage, double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc, double *gm,
             int single, double rtol, double abstol) {
  if (sc->done) return;
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double yi = y[i];
    if (MODE == 1) acc[0] += z[i] * yi;
    if (MODE == 2) { acc[0] += yi * xg[r0 + i]; acc[1] += yi * yi; }
  }
  if (MODE == 1) { double a1[1] = {acc[0]}; finish_stage<1>(a1, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol); }
  if (MODE == 2) finish_stage<2>(acc, stage, partials, counter, dots_local, sc, gm, single, rtol, abstol);
}

template <int MODE>
static int apply_op(cudaStream_t st, const DevSystem &S, PcLU *pc, const double *xg, double *y, const double *z,
                    int stage, KrylovWork &kw, int single, double rtol, double abstol, const PeerView *pv = nullptr) {
  if (!pc) return launch_kspmv<MODE>(st, S, xg, y, z, stage, kw, single, rtol, abstol, pv);
  UFE_TRY(launch_kspmv<0>(st, S, xg, kw.pctmp, nullptr, 0, kw, single, rtol, abstol, pv));
  UFE_TRY(ufe_pclu_apply(st, pc, kw.pctmp, y));
  if (MODE != 0) {
    k_stage_dots<MODE><<<UFE_RED_BLOCKS, UFE_RED_THREADS, 0, st>>>(S.m_loc, S.r1 - 1, y, z, xg, stage, kw.partials, kw.counter,
                                                                 kw.dots_local, kw.sc, kw.gm, single, rtol, abstol);
    UFE_LAUNCH_CHECK();
  }
  return UFE_OK;
}

// ---------------------------------------------------------------------------------
// BiCGStab vector kernels (grid-stride, fixed grid => deterministic reductions)
// ---------------------------------------------------------------------------------
// r = b - Ax (Ax in r on entry if have_ax), rhat = r, p = r -> pg ; x unchanged.
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_init(int n, int r0, const double *__restrict__ b, double *__restrict__ r, double *__restrict__ rhat,
            double *__restrict__ pg, double *__restrict__ xg, int have_ax, double *partials, unsigned *counter,
            double *dots_local, KrylovScalars *sc, int single, double rtol, double abstol, int stage = ST_INIT) {
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = b[i];
    double ri = bi;
    if (have_ax) ri = bi - r[i]; else xg[r0 + i] = 0.0;
    r[i] = ri; rhat[i] = ri; pg[r0 + i] = ri;
    acc[0] += ri * ri; acc[1] += bi * bi;
  }
  finish_stage<2>(acc, stage, partials, counter, dots_local, sc, nullptr, single, rtol, abstol);
}

// p = r + beta (p - omega v)
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_p(int n, int r0, const double *__restrict__ r, const double *__restrict__ v, double *__restrict__ pg,
         const KrylovScalars *sc, unsigned *counter, int sig_epoch) {
  if (sc->done) return;
  const double beta = sc->beta, omega = sc->omega;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    pg[r0 + i] = r[i] + beta * (pg[r0 + i] - omega * v[i]);
  signal_peers_when_done(counter, sc, sig_epoch);
}

// s = r - alpha v
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_s(int n, int r0, const double *__restrict__ r, const double *__restrict__ v, double *__restrict__ sg,
         const KrylovScalars *sc, unsigned *counter, int sig_epoch) {
  if (sc->done) return;
  const double alpha = sc->alpha;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    sg[r0 + i] = r[i] - alpha * v[i];
  signal_peers_when_done(counter, sc, sig_epoch);
}

// s = r - alpha v with (s,s): used with an explicit preconditioner, where an iteration is
// expensive and the first half-step usually converges already
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_s_norm(int n, int r0, const double *__restrict__ r, const double *__restrict__ v, double *__restrict__ sg,
              double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc, int single, double rtol,
              double abstol) {
  if (sc->done) return;
  const double alpha = sc->alpha;
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double si = r[i] - alpha * v[i];
    sg[r0 + i] = si;
    acc[0] += si * si;
  }
  finish_stage<1>(acc, ST_S, partials, counter, dots_local, sc, nullptr, single, rtol, abstol);
}
// x += alpha p, once, if the half-step exit fired (pad0 == 1); k_half_ack then retires the flag
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_xhalf(int n, int r0, double *__restrict__ xg, const double *__restrict__ pg, const KrylovScalars *sc) {
  if (sc->pad0 != 1) return;
  const double alpha = sc->alpha;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xg[r0 + i] += alpha * pg[r0 + i];
}
__global__ void k_half_ack(KrylovScalars *sc) { if (sc->pad0 == 1) sc->pad0 = 2; }

// x += alpha p + omega s ; r = s - omega t ; dots (r,r), (rhat,r)
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_bicg_xr(int n, int r0, double *__restrict__ xg, const double *__restrict__ pg, const double *__restrict__ sg,
          const double *__restrict__ t, const double *__restrict__ rhat, double *__restrict__ r,
          double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc, int single, double rtol,
          double abstol) {
  if (sc->done) return;
  const double alpha = sc->alpha, omega = sc->omega;
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double si = sg[r0 + i];
    xg[r0 + i] += alpha * pg[r0 + i] + omega * si;
    const double ri = si - omega * t[i];
    r[i] = ri;
    acc[0] += ri * ri; acc[1] += rhat[i] * ri;
  }
  finish_stage<2>(acc, ST_C, partials, counter, dots_local, sc, nullptr, single, rtol, abstol);
}

__global__ void k_sc_reset(KrylovScalars *sc, int maxits, double abstol) {
  sc->abstol = abstol;
  sc->its = 0; sc->done = 0; sc->reason = 0; sc->maxits = maxits; sc->jcount = 0; sc->finalized = 0; sc->pad0 = 0;
  sc->bnorm = -1.0; sc->rnorm = 0.0; sc->ttol = 0.0; sc->rho = 1.0; sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
}

// ---------------------------------------------------------------------------------
// GMRES(30) kernels (PETSc-default twin: classical Gram-Schmidt, no refinement)
// ---------------------------------------------------------------------------------
// r = b - Ax (Ax in w on entry if have_ax) -> w ; dots (r,r), (b,b)
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_gm_resid(int n, int r0, const double *__restrict__ b, double *__restrict__ w, double *__restrict__ xg,
           int have_ax, int zero_x, double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc,
           double *gm, int single, double rtol, double abstol) {
  if (sc->done) return;
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bi = b[i];
    double ri = bi;
    if (have_ax) ri = bi - w[i];
    if (zero_x) xg[r0 + i] = 0.0;
    w[i] = ri;
    acc[0] += ri * ri; acc[1] += bi * bi;
  }
  finish_stage<2>(acc, ST_GM_INIT, partials, counter, dots_local, sc, gm, single, rtol, abstol);
}

// V_j = w * inv_hn -> Vb[j], pg
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_gm_scale(int n, int r0, const double *__restrict__ w, double *__restrict__ Vj, double *__restrict__ pg,
           const KrylovScalars *sc, const double *gm, int j, unsigned *counter, int sig_epoch) {
  if (sc->done || sc->jcount != j) return;
  const double inv = *GM_INVHN(gm);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double vi = w[i] * inv;
    Vj[i] = vi; pg[r0 + i] = vi;
  }
  signal_peers_when_done(counter, sc, sig_epoch);
}

// h_i = (V_i, w), i = i0 .. i0+NV-1 (<= j)
template <int NV>
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_gm_mdot(int n, const double *__restrict__ Vb, long long ldv, const double *__restrict__ w, int i0, int cnt,
          double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc) {
  if (sc->done) return;
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; q++) acc[q] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double wi = w[i];
#pragma unroll
    for (int q = 0; q < NV; q++) if (q < cnt) acc[q] += Vb[(size_t)(i0 + q) * ldv + i] * wi;
  }
  reduce_publish<NV>(acc, partials, counter, dots_local + i0);
}

// w -= sum_i h_i V_i ; dot (w,w)
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_gm_update(int n, const double *__restrict__ Vb, long long ldv, double *__restrict__ w, int j,
            double *partials, unsigned *counter, double *dots_local, KrylovScalars *sc, double *gm, int single,
            double rtol, double abstol) {
  if (sc->done) return;
  const double *h = GM_HCOL(gm);
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double wi = w[i];
    for (int q = 0; q <= j; q++) wi -= h[q] * Vb[(size_t)q * ldv + i];
    w[i] = wi;
    acc[0] += wi * wi;
  }
  finish_stage<1>(acc, ST_GM_NORM, partials, counter, dots_local, sc, gm, single, rtol, abstol);
}

// back-substitution for y (one thread), then x += sum_i y_i V_i
__global__ void k_gm_solve_y(KrylovScalars *sc, double *gm) {
  if (sc->finalized) return;
  const int j = sc->jcount;
  double *y = GM_Y(gm);
  const double *g = GM_G(gm), *H = GM_H(gm);
  for (int i = j - 1; i >= 0; i--) {
    double s = g[i];
    for (int k = i + 1; k < j; k++) s -= H[i + 31 * k] * y[k];
    y[i] = s / H[i + 31 * i];
  }
}
__global__ void __launch_bounds__(UFE_RED_THREADS)
k_gm_xupdate(int n, int r0, const double *__restrict__ Vb, long long ldv, double *__restrict__ xg,
             const KrylovScalars *sc, const double *gm) {
  if (sc->finalized) return;
  const int j = sc->jcount;
  const double *y = GM_Y(gm);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double xi = xg[r0 + i];
    for (int q = 0; q < j; q++) xi += y[q] * Vb[(size_t)q * ldv + i];
    xg[r0 + i] = xi;
  }
}
__global__ void k_gm_cycle_end(KrylovScalars *sc) {
  if (sc->done) sc->finalized = 1;
  sc->jcount = 0;
}

// ---------------------------------------------------------------------------------
// halo exchange of a global-indexed vector (contiguous owned ranges per rank)
// ---------------------------------------------------------------------------------
int ufe_halo_exchange(cudaStream_t st, const Comm &comm, const HaloPlan &plan, double *x, long long ld,
                      int nlayers, int mult) {
  if (comm.nranks <= 1) return UFE_OK;
  const int me = comm.rank;
  UFE_NCCL(ncclGroupStart());
  for (int q = 0; q < comm.nranks; q++) {
    if (q == me) continue;
    // what I need from q: intersection of my need range with q's owned range
    int lo = plan.need_lo[me] > plan.own_lo[q] ? plan.need_lo[me] : plan.own_lo[q];
    int hi = plan.need_hi[me] < plan.own_hi[q] ? plan.need_hi[me] : plan.own_hi[q];
    if (hi > lo)
      for (int l = 0; l < nlayers; l++)
        UFE_NCCL(ncclRecv(x + l * ld + (long long)lo * mult, (size_t)(hi - lo) * mult, ncclDouble, q, comm.nccl, st));
    // what q needs from me
    lo = plan.need_lo[q] > plan.own_lo[me] ? plan.need_lo[q] : plan.own_lo[me];
    hi = plan.need_hi[q] < plan.own_hi[me] ? plan.need_hi[q] : plan.own_hi[me];
    if (hi > lo)
      for (int l = 0; l < nlayers; l++)
        UFE_NCCL(ncclSend(x + l * ld + (long long)lo * mult, (size_t)(hi - lo) * mult, ncclDouble, q, comm.nccl, st));
  }
  UFE_NCCL(ncclGroupEnd());
  g_launch_count++;
  return UFE_OK;
}

// ---------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------
int ufe_krylov_alloc(KrylovWork &kw, int N, int n_loc, bool gmres, double *ext_pg, double *ext_sg) {
  kw.n_loc = n_loc; kw.N = N;
  size_t nb = (size_t)(n_loc > 0 ? n_loc : 1) * sizeof(double);
  size_t Nb = (size_t)(N > 0 ? N : 1) * sizeof(double);
  UFE_CUDA(cudaMalloc(&kw.r, nb)); UFE_CUDA(cudaMalloc(&kw.rhat, nb));
  UFE_CUDA(cudaMalloc(&kw.v, nb)); UFE_CUDA(cudaMalloc(&kw.t, nb));
  UFE_CUDA(cudaMalloc(&kw.w, nb));
  UFE_CUDA(cudaMalloc(&kw.pctmp, nb)); UFE_CUDA(cudaMalloc(&kw.bP, nb));
  if (ext_pg && ext_sg) { kw.pg = ext_pg; kw.sg = ext_sg; kw.ext_vecs = true; }
  else { UFE_CUDA(cudaMalloc(&kw.pg, Nb)); UFE_CUDA(cudaMalloc(&kw.sg, Nb)); }
  UFE_CUDA(cudaMemset(kw.pg, 0, Nb)); UFE_CUDA(cudaMemset(kw.sg, 0, Nb));
  if (gmres) UFE_CUDA(cudaMalloc(&kw.Vb, nb * (GM_RESTART + 1)));
  UFE_CUDA(cudaMalloc(&kw.partials, sizeof(double) * KGRID * 8));
  UFE_CUDA(cudaMalloc(&kw.dots_local, sizeof(double) * 48));
  UFE_CUDA(cudaMalloc(&kw.counter, sizeof(unsigned)));
  UFE_CUDA(cudaMemset(kw.counter, 0, sizeof(unsigned)));
  UFE_CUDA(cudaMalloc(&kw.sc, sizeof(KrylovScalars)));
  UFE_CUDA(cudaMemset(kw.sc, 0, sizeof(KrylovScalars)));
  UFE_CUDA(cudaMalloc(&kw.gm, sizeof(double) * GM_SIZE));
  UFE_CUDA(cudaMemset(kw.gm, 0, sizeof(double) * GM_SIZE));
  // callers run on non-blocking streams: the null-stream memsets above are not ordered before their kernels
  // (a stale reduction counter turns every dot product into garbage), so wait for them here
  UFE_CUDA(cudaStreamSynchronize(0));
  UFE_CUDA(cudaMallocHost(&kw.sc_host, sizeof(KrylovScalars)));
  return UFE_OK;
}

void ufe_krylov_free(KrylovWork &kw) {
  cudaFree(kw.r); cudaFree(kw.rhat); cudaFree(kw.v); cudaFree(kw.t); cudaFree(kw.w);
  if (!kw.ext_vecs) { cudaFree(kw.pg); cudaFree(kw.sg); }
  cudaFree(kw.Vb); cudaFree(kw.partials); cudaFree(kw.dots_local);
  cudaFree(kw.counter); cudaFree(kw.sc); cudaFree(kw.gm); cudaFree(kw.pctmp); cudaFree(kw.bP); cudaFree(kw.peer_dev);
  if (kw.sc_host) cudaFreeHost(kw.sc_host);
  kw = KrylovWork();
}

static PeerFlagPtrs peer_flags(const PeerComm &pc) {
  PeerFlagPtrs f;
  for (int q = 0; q < UFE_MAX_RANKS; q++) f.f[q] = q < pc.P ? reinterpret_cast<int *>(pc.base[q] + pc.off_flags) : nullptr;
  return f;
}
static PeerDotPtrs peer_dots(const PeerComm &pc) {
  PeerDotPtrs d;
  for (int q = 0; q < UFE_MAX_RANKS; q++) d.d[q] = q < pc.P ? pc.base[q] + pc.off_dots : nullptr;
  return d;
}

// Make the SpMV input vector `vec` (global-indexed, owned part valid) usable by the next SpMV:
// peer mode: publish an epoch to every rank and hand the kernel a PeerView (the halo is read in
// place from the owners' buffers); otherwise exchange the halo ranges through NCCL.
static int sync_input(cudaStream_t st, const Comm &comm, const HaloPlan *halo, double *vec, PeerView *pv, bool *use_pv,
                      int pre_epoch = 0) {
  *use_pv = false;
  if (!halo || comm.nranks <= 1) return UFE_OK;
  PeerComm &pc = comm.peer;
  if (!pc.on) return ufe_halo_exchange(st, comm, *halo, vec, 0, 1, halo->mult);
  int epoch = pre_epoch;          // > 0: the producing kernel has published this epoch itself
  if (epoch <= 0) {
    epoch = (int)(++pc.halo_epoch);
    k_peer_signal<<<1, 32, 0, st>>>(pc.P, pc.me, epoch, peer_flags(pc));
    UFE_LAUNCH_CHECK();
  }
  const long long off = vec - pc.base[pc.me];
  pv->P = pc.P; pv->me = pc.me; pv->epoch = epoch;
  for (int q = 0; q <= UFE_MAX_RANKS; q++) pv->bounds[q] = q <= pc.P ? pc.bounds[q] : 0x7fffffff;
  for (int q = 0; q < UFE_MAX_RANKS; q++) pv->xp[q] = q < pc.P ? reinterpret_cast<const double2 *>(pc.base[q] + off) : nullptr;
  pv->flags = reinterpret_cast<const volatile int *>(pc.base[pc.me] + pc.off_flags);
  *use_pv = true;
  return UFE_OK;
}

static int peer_reduce(cudaStream_t st, const Comm &comm, KrylovWork &kw, int stage, int nd, const double *dots_local,
                       double rtol, double abstol) {
  PeerComm &pc = comm.peer;
  const int epoch = (int)(++pc.red_epoch);
  k_peer_reduce<<<1, 32, 0, st>>>(pc.P, pc.me, epoch, nd, stage, peer_flags(pc), peer_dots(pc), dots_local, kw.sc, kw.gm, rtol, abstol);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

static int allreduce_stage(cudaStream_t st, const Comm &comm, KrylovWork &kw, int stage, int nd, double rtol,
                           double abstol) {
  if (comm.nranks <= 1) return UFE_OK;
  if (comm.peer.on) return UFE_OK;     // fused into the producing kernel (finish_stage, single < 0)
  UFE_NCCL(ncclAllReduce(kw.dots_local, kw.dots_local, nd, ncclDouble, ncclSum, comm.nccl, st));
  g_launch_count++;
  k_post<<<1, 1, 0, st>>>(stage, kw.sc, kw.gm, kw.dots_local, nd, rtol, abstol);
  UFE_LAUNCH_CHECK();
  return UFE_OK;
}

static int poll(cudaStream_t st, KrylovWork &kw) {
  UFE_CUDA(cudaMemcpyAsync(kw.sc_host, kw.sc, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, st));
  UFE_CUDA(cudaStreamSynchronize(st));
  return UFE_OK;
}

static int run_bicgstab(cudaStream_t st, const DevSystem &S, KrylovWork &kw, const Comm &comm,
                        const HaloPlan *halo, double rtol, double abstol, int maxits, int guess_nonzero, PcLU *pc) {
  const int n = S.m_loc, r0 = S.r1 - 1, single = comm.nranks <= 1;
  const bool peer = !single && comm.peer.on && halo;
  const int G = UFE_RED_BLOCKS, B = UFE_RED_THREADS;
  double *xg = S.x;
  PeerView pv; bool use_pv = false;
  // `single` argument of a reducing kernel: 1 one rank, 0 NCCL, -epoch peer memory (fused all-to-all)
  auto sarg = [&]() { return single ? 1 : (peer ? -(int)(++comm.peer.red_epoch) : 0); };
  auto sig = [&]() { return peer ? (int)(++comm.peer.halo_epoch) : 0; };   // epoch a producer kernel publishes itself
  k_sc_reset<<<1, 1, 0, st>>>(kw.sc, maxits, abstol); UFE_LAUNCH_CHECK();
  if (guess_nonzero) {
    UFE_TRY(sync_input(st, comm, halo, xg, &pv, &use_pv));
    UFE_TRY(apply_op<0>(st, S, pc, xg, kw.r, nullptr, 0, kw, single, rtol, abstol, use_pv ? &pv : nullptr));
  }
  const double *bS = S.bS;
  if (pc) { UFE_TRY(ufe_pclu_apply(st, pc, S.bS, kw.bP)); bS = kw.bP; }
  k_bicg_init<<<G, B, 0, st>>>(n, r0, bS, kw.r, kw.rhat, kw.pg, xg, guess_nonzero, kw.partials, kw.counter,
                               kw.dots_local, kw.sc, sarg(), rtol, abstol);
  UFE_LAUNCH_CHECK();
  UFE_TRY(allreduce_stage(st, comm, kw, ST_INIT, 2, rtol, abstol));
  int launched = 0, batch = pc ? 1 : 4;     // an exact block solve converges in the first (half) step
  while (true) {
    for (int b = 0; b < batch; b++, launched++) {
      int ep = 0;
      if (launched > 0) { ep = sig(); k_bicg_p<<<G, B, 0, st>>>(n, r0, kw.r, kw.v, kw.pg, kw.sc, kw.counter, ep); UFE_LAUNCH_CHECK(); }
      UFE_TRY(sync_input(st, comm, halo, kw.pg, &pv, &use_pv, ep));
      UFE_TRY(apply_op<1>(st, S, pc, kw.pg, kw.v, kw.rhat, ST_A, kw, sarg(), rtol, abstol, use_pv ? &pv : nullptr));
      UFE_TRY(allreduce_stage(st, comm, kw, ST_A, 1, rtol, abstol));
      ep = 0;
      if (pc) {
        k_bicg_s_norm<<<G, B, 0, st>>>(n, r0, kw.r, kw.v, kw.sg, kw.partials, kw.counter, kw.dots_local, kw.sc, sarg(), rtol, abstol);
        UFE_LAUNCH_CHECK();
        UFE_TRY(allreduce_stage(st, comm, kw, ST_S, 1, rtol, abstol));
        k_bicg_xhalf<<<G, B, 0, st>>>(n, r0, xg, kw.pg, kw.sc); UFE_LAUNCH_CHECK();
        k_half_ack<<<1, 1, 0, st>>>(kw.sc); UFE_LAUNCH_CHECK();
        // an exact block solve converges in this half step: look before paying for the second operator application
        // (SpMV + forward / backward sweep over the whole factorisation; the sweeps run from a replayed graph and
        // cannot return early on the device-side `done` flag)
        if (launched == 0) { UFE_TRY(poll(st, kw)); if (kw.sc_host->done) return UFE_OK; }     // same decision on every rank
      } else { ep = sig(); k_bicg_s<<<G, B, 0, st>>>(n, r0, kw.r, kw.v, kw.sg, kw.sc, kw.counter, ep); UFE_LAUNCH_CHECK(); }
      UFE_TRY(sync_input(st, comm, halo, kw.sg, &pv, &use_pv, ep));
      UFE_TRY(apply_op<2>(st, S, pc, kw.sg, kw.t, nullptr, ST_B, kw, sarg(), rtol, abstol, use_pv ? &pv : nullptr));
      UFE_TRY(allreduce_stage(st, comm, kw, ST_B, 2, rtol, abstol));
      k_bicg_xr<<<G, B, 0, st>>>(n, r0, xg, kw.pg, kw.sg, kw.t, kw.rhat, kw.r, kw.partials, kw.counter,
                                 kw.dots_local, kw.sc, sarg(), rtol, abstol);
      UFE_LAUNCH_CHECK();
      UFE_TRY(allreduce_stage(st, comm, kw, ST_C, 2, rtol, abstol));
    }
    UFE_TRY(poll(st, kw));
    if (kw.sc_host->done) break;
    if (batch < 64) batch *= 2;
  }
  return UFE_OK;
}

// An exact, fresh factorisation M = A: x = M^-1 b IS the solution up to round-off, so before any Krylov machinery take
// that one Richardson step from x = 0 and test the reference's stopping rule on the true residual of the scaled system,
// |B (b - A x)| <= max(rtol |B b|, abstol) (petsc_basic.f90:106-128 with B the Jacobi scaling) -- one preconditioner
// application and one SpMV per linear solve.  If the test fails (a perturbed pivot, an ill-conditioned front) the Krylov
// method continues from this x as a non-zero initial guess.  *done: the system is solved.
static int richardson_first(cudaStream_t st, const DevSystem &S, KrylovWork &kw, const Comm &comm, const HaloPlan *halo,
                            double rtol, double abstol, int maxits, PcLU *pc, bool *done) {
  const int n = S.m_loc, r0 = S.r1 - 1, single = comm.nranks <= 1;
  const bool peer = !single && comm.peer.on && halo;
  PeerView pv; bool use_pv = false;
  k_sc_reset<<<1, 1, 0, st>>>(kw.sc, maxits, abstol); UFE_LAUNCH_CHECK();
  UFE_TRY(ufe_pclu_apply(st, pc, S.bS, S.x + r0));
  UFE_TRY(sync_input(st, comm, halo, S.x, &pv, &use_pv));
  UFE_TRY(launch_kspmv<0>(st, S, S.x, kw.r, nullptr, 0, kw, single, rtol, abstol, use_pv ? &pv : nullptr));
  k_bicg_init<<<UFE_RED_BLOCKS, UFE_RED_THREADS, 0, st>>>(n, r0, S.bS, kw.r, kw.rhat, kw.pg, S.x, 1, kw.partials, kw.counter, kw.dots_local, kw.sc,
                                                          single ? 1 : (peer ? -(int)(++comm.peer.red_epoch) : 0), rtol, abstol, ST_RICH);
  UFE_LAUNCH_CHECK();
  UFE_TRY(allreduce_stage(st, comm, kw, ST_RICH, 2, rtol, abstol));
  UFE_TRY(poll(st, kw));
  *done = kw.sc_host->done != 0;
  return UFE_OK;
}

static int run_gmres(cudaStream_t st, const DevSystem &S, KrylovWork &kw, const Comm &comm, const HaloPlan *halo,
                     double rtol, double abstol, int maxits, int guess_nonzero, bool reset, PcLU *pc) {
  const int n = S.m_loc, r0 = S.r1 - 1, single = comm.nranks <= 1;
  const int G = UFE_RED_BLOCKS, B = UFE_RED_THREADS;
  const long long ldv = n;
  double *xg = S.x;
  const bool peer = !single && comm.peer.on && halo;
  auto sarg = [&]() { return single ? 1 : (peer ? -(int)(++comm.peer.red_epoch) : 0); };
  auto sig = [&]() { return peer ? (int)(++comm.peer.halo_epoch) : 0; };
  if (reset) { k_sc_reset<<<1, 1, 0, st>>>(kw.sc, maxits, abstol); UFE_LAUNCH_CHECK(); }
  const double *bS = S.bS;
  if (pc) { UFE_TRY(ufe_pclu_apply(st, pc, S.bS, kw.bP)); bS = kw.bP; }
  bool first = true;
  PeerView pv; bool use_pv = false;
  while (true) {
    const int have_ax = (!first || guess_nonzero) ? 1 : 0;
    if (have_ax) {
      UFE_TRY(sync_input(st, comm, halo, xg, &pv, &use_pv));
      UFE_TRY(apply_op<0>(st, S, pc, xg, kw.w, nullptr, 0, kw, single, rtol, abstol, use_pv ? &pv : nullptr));
    }
    k_gm_resid<<<G, B, 0, st>>>(n, r0, bS, kw.w, xg, have_ax, (first && !guess_nonzero) ? 1 : 0, kw.partials,
                                kw.counter, kw.dots_local, kw.sc, kw.gm, sarg(), rtol, abstol);
    UFE_LAUNCH_CHECK();
    UFE_TRY(allreduce_stage(st, comm, kw, ST_GM_INIT, 2, rtol, abstol));
    first = false;
    for (int j = 0; j < GM_RESTART; j++) {
      const int ep = sig();
      k_gm_scale<<<G, B, 0, st>>>(n, r0, kw.w, kw.Vb + (size_t)j * ldv, kw.pg, kw.sc, kw.gm, j, kw.counter, ep); UFE_LAUNCH_CHECK();
      UFE_TRY(sync_input(st, comm, halo, kw.pg, &pv, &use_pv, ep));
      UFE_TRY(apply_op<0>(st, S, pc, kw.pg, kw.w, nullptr, 0, kw, single, rtol, abstol, use_pv ? &pv : nullptr));
      for (int i0 = 0; i0 <= j; i0 += 8) {
        const int cnt = (j + 1 - i0) < 8 ? (j + 1 - i0) : 8;
        k_gm_mdot<8><<<G, B, 0, st>>>(n, kw.Vb, ldv, kw.w, i0, cnt, kw.partials, kw.counter, kw.dots_local, kw.sc);
        UFE_LAUNCH_CHECK();
      }
      if (peer) {
        UFE_TRY(peer_reduce(st, comm, kw, -1, j + 1, kw.dots_local, rtol, abstol));     // sums straight into hcol
      } else {
        if (!single) {
          UFE_NCCL(ncclAllReduce(kw.dots_local, kw.dots_local, j + 1, ncclDouble, ncclSum, comm.nccl, st));
          g_launch_count++;
        }
        k_gm_sethcol<<<1, 1, 0, st>>>(kw.sc, kw.gm, kw.dots_local, j + 1);
        UFE_LAUNCH_CHECK();
      }
      k_gm_update<<<G, B, 0, st>>>(n, kw.Vb, ldv, kw.w, j, kw.partials, kw.counter, kw.dots_local + 40, kw.sc,
                                   kw.gm, sarg(), rtol, abstol);
      UFE_LAUNCH_CHECK();
      if (!single && !peer) {
        UFE_NCCL(ncclAllReduce(kw.dots_local + 40, kw.dots_local + 40, 1, ncclDouble, ncclSum, comm.nccl, st));
        g_launch_count++;
        k_post<<<1, 1, 0, st>>>(ST_GM_NORM, kw.sc, kw.gm, kw.dots_local + 40, 1, rtol, abstol); UFE_LAUNCH_CHECK();
      }
    }
    k_gm_solve_y<<<1, 1, 0, st>>>(kw.sc, kw.gm); UFE_LAUNCH_CHECK();
    k_gm_xupdate<<<G, B, 0, st>>>(n, r0, kw.Vb, ldv, xg, kw.sc, kw.gm); UFE_LAUNCH_CHECK();
    k_gm_cycle_end<<<1, 1, 0, st>>>(kw.sc); UFE_LAUNCH_CHECK();
    UFE_TRY(poll(st, kw));
    if (kw.sc_host->done) break;
  }
  return UFE_OK;
}

int ufe_krylov_run(cudaStream_t st, const DevSystem &S, KrylovWork &kw, const Comm &comm, const HaloPlan *halo,
                   int method, double rtol, double abstol, int maxits, int guess_nonzero, int *n_its, int *flags,
                   PcLU *pc) {
  if (maxits <= 0) maxits = 10000;
  int fl = 0;
  if (kw.sc_host->reason == -9) {      // an earlier solve on this workspace timed out on a peer: blocks that started after
    UFE_CUDA(cudaMemsetAsync(kw.counter, 0, sizeof(unsigned), st));    // the time-out left the block counter mid-count
    kw.sc_host->reason = 0;
  }
  static const bool rich_off = getenv("UFE_RICHARDSON_FIRST") && atoi(getenv("UFE_RICHARDSON_FIRST")) == 0;
  bool solved = false;
  if (pc && !guess_nonzero && !rich_off && ufe_pclu_exact_and_fresh(pc)) {
    UFE_TRY(richardson_first(st, S, kw, comm, halo, rtol, abstol, maxits, pc, &solved));
    guess_nonzero = 1;
  }
  if (solved) {
  } else if (method == UFE_KRYLOV_GMRES) {
    if (!kw.Vb) { ufe_set_error("GMRES workspace not allocated"); return UFE_ERR_INVALID; }
    UFE_TRY(run_gmres(st, S, kw, comm, halo, rtol, abstol, maxits, guess_nonzero, true, pc));
  } else {
    UFE_TRY(run_bicgstab(st, S, kw, comm, halo, rtol, abstol, maxits, guess_nonzero, pc));
    if (kw.sc_host->reason == -5 && kw.Vb) {
      // BiCGStab breakdown: continue from the current iterate with GMRES(30)
      const int its0 = kw.sc_host->its;
      UFE_TRY(run_gmres(st, S, kw, comm, halo, rtol, abstol, maxits, 1, true, pc));
      kw.sc_host->its += its0;
    }
  }
  const int reason = kw.sc_host->reason;
  if (reason == -9) { ufe_set_error("peer-memory synchronisation timed out (another rank stopped?)"); return UFE_ERR_CUDA; }
  if (reason == -3) fl |= UFE_FLAG_KRYLOV_MAXIT;
  if (reason == -4 || reason == -5) fl |= UFE_FLAG_KRYLOV_DIVERGED;
  if (n_its) *n_its = kw.sc_host->its;
  if (flags) *flags = fl;
  return UFE_OK;
}
