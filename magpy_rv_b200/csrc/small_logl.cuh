// Fused small-N log-likelihood: ONE CTA PER WALKER, the covariance never leaves shared memory.
//
// Pipeline inside the CTA (reference call stack: mcmc.py:421-458 -> get_model -> GPLikelihood.LogL
// -> compute_kernel -> slogdet/cho_factor/cho_solve, gp_likelihood.py:250-278, 357-402):
//   1. mean models + residual r = y - model            (mean_model.cuh)
//   2. lower triangle of K built in smem from t, yerr  (common.cuh cov_eval / cov_diag)
//   3. in-place square-root-free Cholesky (LDL^T) of the matrix augmented with r as an extra row:
//        for j: A[i][c] -= A[i][j] * A[c][j] / A[j][j]     (c > j, i >= c, i == N is the r row)
//      one barrier per column; afterwards A[j][j] = d_j (pivots), A[N][j] = (L^-1 r)_j * sqrt(d_j)
//   4. logdet = sum ln d_j,  quad = sum A[N][j]^2 / d_j   (deterministic block reductions)
//   5. logL = -N/2 ln 2pi - logdet/2 - quad/2, then + priors in list order.
// Non-PD K shows up as a pivot d_j <= 0 exactly where LAPACK potrf reports info = j+1.
//
// smem: column-major (N+1) x N with an odd leading dimension (bank-conflict-free rows and
// columns) + t + r + small vectors.  N <= 168 fits the 227 KB opt-in limit.
#pragma once
#include "blocked.cuh"
#include "mean_model.cuh"

#define MRV_SMALL_THREADS 256

__host__ __device__ inline int small_ld(int N) { return (N + 1) | 1; }
__host__ inline size_t small_smem_bytes(int N) {
  return ((size_t)small_ld(N) * N + 4 * (size_t)N + MRV_MAX_MODPAR + MRV_MAX_HP + 40) * sizeof(double);
}

__global__ void __launch_bounds__(MRV_SMALL_THREADS)
fused_small_logl_kernel(ProblemDesc D, const double* __restrict__ t, const double* __restrict__ y,
                        const double* __restrict__ yerr, const double* __restrict__ flags,
                        const double* __restrict__ hp_all, const double* __restrict__ modpar_all,
                        const double* __restrict__ model_given, int to_ecc, double* __restrict__ logl_out,
                        int* __restrict__ status_out, SolveOut so) {
  extern __shared__ double smem[];
  const int N = D.N, ld = small_ld(N);
  double* S = smem;                         // ld x N, column major; row N = residual row
  double* ts = S + (size_t)ld * N;          // 3 N: t, phase cosine, phase sine of every point (cov_pair_n)
  double* r = ts + 3 * N;                   // N
  double* phys = r + N;                     // MRV_MAX_MODPAR
  double* hpv = phys + MRV_MAX_MODPAR;      // MRV_MAX_HP
  double* red = hpv + MRV_MAX_HP;           // 32
  int* ired = (int*)(red + 32);             // 32 ints
  const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

  if (tid == 0) {
    for (int i = 0; i < D.nh; ++i) hpv[i] = hp_all[(size_t)w * D.nh + i];
    for (int i = 0; i < D.nm; ++i) phys[i] = modpar_all[(size_t)w * D.nm + i];
    if (model_given == nullptr) convert_to_ecc(D, phys, to_ecc);
  }
  for (int i = tid; i < N; i += nt) ts[i] = t[i];
  __syncthreads();

  // 1. residual
  if (model_given != nullptr) {
    for (int i = tid; i < N; i += nt) r[i] = model_given[(size_t)w * N + i];
    __syncthreads();
  } else {
    cta_mean_model(D, ts, flags, phys, r, /*E scratch=*/S, red);
  }
  int bad = 0;
  for (int i = tid; i < N; i += nt) {
    const double v = so.given_is_resid ? r[i] : __dsub_rn(y[i], r[i]);
    r[i] = v;
    bad |= !isfinite(v);
  }
  __syncthreads();

  // 2. K (lower triangle) + residual row
  const CovParams cp = make_cov_params(D.kernel_id, D.variant, hpv);
  for (int i = tid; i < N; i += nt) cov_point_phase(cp, ts[i] - D.t_ref, ts[N + i], ts[2 * N + i]);
  __syncthreads();
  for (int c = warp; c < N; c += nwarps) {
    double* col = S + (size_t)c * ld;
    const double tc[1] = {ts[c]}, cc[1] = {ts[N + c]}, sc[1] = {ts[2 * N + c]};
    for (int i = c + lane; i < N; i += 32) {
      const double ti1[1] = {ts[i]}, ci1[1] = {ts[N + i]}, si1[1] = {ts[2 * N + i]};
      double ev[1];
      cov_pair_n<1>(cp, ti1, ci1, si1, tc, cc, sc, ev);
      const double v = (i == c) ? cov_diag(cp, yerr[i]) : ev[0];
      bad |= !isfinite(v);
      col[i] = v;
    }
    if (lane == 0) col[N] = r[c];
  }

  // 3. right-looking LDL^T with the residual row carried along
  for (int j = 0; j < N - 1; ++j) {
    __syncthreads();
    const double* cj = S + (size_t)j * ld;
    const double winv = 1.0 / cj[j];
    for (int c = j + 1 + warp; c < N; c += nwarps) {
      double* col = S + (size_t)c * ld;
      const double beta = cj[c] * winv;
      for (int i = c + lane; i <= N; i += 32) col[i] = fma(-cj[i], beta, col[i]);
    }
  }
  __syncthreads();

  // 4. reductions over the pivots
  double ld_part = 0.0, q_part = 0.0;
  int first_bad = 0x7fffffff;
  for (int j = tid; j < N; j += nt) {
    const double d = S[(size_t)j * ld + j];
    const double zr = S[(size_t)j * ld + N];
    if (!(d > 0.0)) first_bad = min(first_bad, j + 1);
    ld_part += log(d);
    q_part += (zr * zr) / d;
  }
  const double logdet = block_sum(ld_part, red);
  const double quad = block_sum(q_part, red);
  const int fb = -block_max_int(-first_bad, ired);
  const int anybad = block_max_int(bad, ired);
  if (so.z != nullptr)
    for (int j = tid; j < N; j += nt) so.z[(size_t)w * so.ldz + j] = S[(size_t)j * ld + N] / sqrt(S[(size_t)j * ld + j]);
  if (so.logdet != nullptr && tid == 0) so.logdet[w] = logdet;

  // 5. closed form + priors
  if (tid == 0) {
    double v = -((double)N / 2.0) * log(2.0 * MRV_PI) - 0.5 * logdet - 0.5 * quad;
    v = add_priors(D, hpv, phys, v);
    int st = 0;
    if (anybad) st = -1;
    else if (fb != 0x7fffffff) st = fb;
    logl_out[w] = v;
    status_out[w] = st;
  }
}


// ------------------------------------------------------------------------------------------------
// Same fused pipeline with the rank-16 blocked sweep (blocked.cuh, cta_block_sweep<false>): the
// trailing updates run on the FP64 tensor pipe (DMMA) straight out of shared memory.  N is padded
// to a multiple of 16 with an identity block; the residual row sits at row n_cols.  Fits N <= 144.
// ------------------------------------------------------------------------------------------------
#define MRV_SMALL_BLK_N_MAX 144
__host__ __device__ inline int sblk_cols(int N) { return (N + 15) & ~15; }
__host__ __device__ inline int sblk_ld(int N) {
  int ld = (sblk_cols(N) + 1 + 7) & ~7;
  if ((ld & 15) != 8) ld += 8;
  return ld;
}
__host__ inline size_t small_blk_smem_bytes(int N) {
  const int nc = sblk_cols(N), ld = sblk_ld(N);
  return (sweep_packed_doubles(ld, nc) + PB_NB * (size_t)(ld + 12) + PB_NB * (size_t)(nc + 4) + 2 * (PB_NB * PB_D_LD + PB_NB) + 4 * (size_t)N +
          MRV_MAX_MODPAR + MRV_MAX_HP + 40) * sizeof(double);
}

// -DSMALL_TIMING: clock64 per phase of CTA 0, thread 0 (tools/small_timing.py): 0 prologue + mean model + residual,
// 1 fill + phase table, 2 covariance generation, 3 sweep, 4 log det / priors / output
__device__ long long g_small_timing[8];
#ifdef SMALL_TIMING
#define SMALL_MARK(i)                                             \
  if (blockIdx.x == 0 && threadIdx.x == 0) {                      \
    const long long now_ = clock64();                             \
    g_small_timing[i] += now_ - tmark_;                           \
    tmark_ = now_;                                                \
  }
#else
#define SMALL_MARK(i)
#endif

// MINB = resident CTAs per SM the register cap is set for (packed factor storage: 2 CTAs fit up to N = 112,
// 3 up to N = 80, 4 up to N = 64; the caps cost 24..60 bytes of spills)
template <int MINB>
__global__ void __launch_bounds__(MRV_SMALL_THREADS, MINB)
fused_small_blk_kernel(ProblemDesc D, const double* __restrict__ t, const double* __restrict__ y,
                       const double* __restrict__ yerr, const double* __restrict__ flags,
                       const double* __restrict__ hp_all, const double* __restrict__ modpar_all,
                       const double* __restrict__ model_given, int to_ecc, double* __restrict__ logl_out,
                       int* __restrict__ status_out, SolveOut so) {
  extern __shared__ double smem[];
  const int N = D.N, nc = sblk_cols(N), ld = sblk_ld(N), pa_ld = ld + 12, pb_ld = nc + 4;
  double* S = smem;                          // packed column blocks (sweep_colp<true>); row nc = residual row
  double* PA = S + sweep_packed_doubles(ld, nc);
  double* PBm = PA + PB_NB * pa_ld;
  double* Dblk = PBm + PB_NB * pb_ld;         // 2 x (16 x 17): pivot blocks kb and kb + 1 (cta_block_sweep_cp)
  double* wv = Dblk + 2 * PB_NB * PB_D_LD;   // 2 x 16
  double* ts = wv + 2 * PB_NB;               // 3 N: t, phase cosine, phase sine of every point (cov_pair_n)
  double* r = ts + 3 * N;                    // N
  double* phys = r + N;                      // MRV_MAX_MODPAR
  double* hpv = phys + MRV_MAX_MODPAR;       // MRV_MAX_HP
  double* red = hpv + MRV_MAX_HP;            // 32
  int* ired = (int*)(red + 32);              // 32 ints
  const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
#ifdef SMALL_TIMING
  long long tmark_ = clock64();
#endif

  if (tid == 0) {
    for (int i = 0; i < D.nh; ++i) hpv[i] = hp_all[(size_t)w * D.nh + i];
    for (int i = 0; i < D.nm; ++i) phys[i] = modpar_all[(size_t)w * D.nm + i];
    if (model_given == nullptr) convert_to_ecc(D, phys, to_ecc);
  }
  for (int i = tid; i < N; i += nt) ts[i] = t[i];
  __syncthreads();

  if (model_given != nullptr) {
    for (int i = tid; i < N; i += nt) r[i] = model_given[(size_t)w * N + i];
    __syncthreads();
  } else {
    cta_mean_model(D, ts, flags, phys, r, /*E scratch=*/S, red);
  }
  int bad = 0;
  for (int i = tid; i < N; i += nt) {
    const double v = so.given_is_resid ? r[i] : __dsub_rn(y[i], r[i]);
    r[i] = v;
    bad |= !isfinite(v);
  }
  __syncthreads();

  SMALL_MARK(0)
  // (a) everything that is not a real lower-triangle entry: zeros, identity padding, residual row
  for (int c = warp; c < nc; c += nwarps) {
    double* col = sweep_colp<true>(S, ld, c);
    for (int i = (c & ~15) + lane; i < ld; i += 32) {   // rows above the column's 16-block are not stored
      if (i < N && c < N && i >= c) continue;   // written by (b)
      double v = 0.0;
      if (i == c) v = 1.0;                      // identity padding (c >= N)
      else if (i == nc && c < N) v = r[c];      // residual row
      col[i] = v;
    }
  }
  // (b) the N(N+1)/2 covariance evaluations, perfectly balanced: a warp takes the column pair
  //     (c, N-1-c), whose lower-triangle lengths add up to N+1, and every lane evaluates two
  //     independent entries per trip (ILP across the fp64 exp/sin dependency chains).
  const CovParams cp = make_cov_params(D.kernel_id, D.variant, hpv);
  for (int i = tid; i < N; i += nt) cov_point_phase(cp, ts[i] - D.t_ref, ts[N + i], ts[2 * N + i]);
  __syncthreads();
  // zero-lag value + jitter once per walker: K_ii = (k(0) + jit^2) + yerr_i^2, the order of cov_diag().  Inside
  // the `(i == c) ? cov_diag(...) : ev` selects the compiler evaluated k(0) -- a full exp / sin -- per element.
  const double kdiag0 = cov_from_dist(cp, 0.0, 0.0) + cp.jit2;
  SMALL_MARK(1)
  const int npairs = (N + 1) / 2;
  // four entries per lane and trip through the branch-free lock-step exp: with at most 16 warps per SM the library exp
  // (one rare-path branch per call, so its evaluations cannot interleave) left the generation latency-bound
  constexpr int GNV = MINB >= 4 ? 2 : 4;   // (the 64-register variant spills with four)
  for (int pc = warp; pc < npairs; pc += nwarps) {
    const int cA = pc, cB = N - 1 - pc;
    const int lenA = N - cA;
    const int total = (cA == cB) ? lenA : (N + 1);
    for (int v0 = lane; v0 < total; v0 += 32 * GNV) {
      int ii[GNV], cc[GNV];
      bool ok[GNV];
      double pti[GNV], pci[GNV], psi[GNV], ptj[GNV], pcj[GNV], psj[GNV], ev[GNV];
#pragma unroll
      for (int q = 0; q < GNV; ++q) {
        const int v = v0 + 32 * q;
        ok[q] = v < total;
        cc[q] = (!ok[q] || v < lenA) ? cA : cB;
        ii[q] = !ok[q] ? cA : ((v < lenA) ? cA + v : cB + (v - lenA));
        pti[q] = ts[ii[q]]; pci[q] = ts[N + ii[q]]; psi[q] = ts[2 * N + ii[q]];
        ptj[q] = ts[cc[q]]; pcj[q] = ts[N + cc[q]]; psj[q] = ts[2 * N + cc[q]];
      }
      cov_pair_n<GNV, true>(cp, pti, pci, psi, ptj, pcj, psj, ev);
#pragma unroll
      for (int q = 0; q < GNV; ++q) {
        double e = ev[q];
        if (ii[q] == cc[q]) {
          const double ye = yerr[ii[q]];
          e = kdiag0 + ye * ye;
        }
        if (ok[q]) {
          bad |= !isfinite(e);
          sweep_colp<true>(S, ld, cc[q])[ii[q]] = e;
        }
      }
    }
  }
  __syncthreads();
  SMALL_MARK(2)
#ifdef SMALL_SWEEP_NO_LOOKAHEAD
  cta_block_sweep<false, MRV_SMALL_THREADS / 32, true>(S, ld, nc, PA, pa_ld, PBm, pb_ld, Dblk, wv);
#else
  // look-ahead form (blocked.cuh).  Row tiles of 8: the MINB = 4 / 3 variants are only launched when 4 / 3 CTAs fit an SM,
  // i.e. N <= 64 / 80 (9 / 11 row tiles incl. the residual row); N <= 144: 19
#ifdef SMALL_SWEEP_LA
  cta_block_sweep_la<false, true, (MINB >= 4 ? 9 : (MINB == 3 ? 11 : 19))>(S, ld, nc, PA, pa_ld, PBm, pb_ld, Dblk, wv);
#else
  cta_block_sweep_cp<false, true, (MINB >= 4 ? 9 : (MINB == 3 ? 11 : 19))>(S, ld, nc, PA, pa_ld, PBm, pb_ld, Dblk, wv);
#endif
#endif

  SMALL_MARK(3)
  double ld_part = 0.0, q_part = 0.0;
  int first_bad = 0x7fffffff;
  for (int j = tid; j < N; j += nt) {
    const double d = sweep_colp<true>(S, ld, j)[j];
    const double zr = sweep_colp<true>(S, ld, j)[nc];
    if (!(d > 0.0)) first_bad = min(first_bad, j + 1);
    ld_part += log(d);
    q_part += (zr * zr) / d;
  }
  const double logdet = block_sum(ld_part, red);
  const double quad = block_sum(q_part, red);
  const int fb = -block_max_int(-first_bad, ired);
  const int anybad = block_max_int(bad, ired);
  if (so.z != nullptr)
    for (int j = tid; j < N; j += nt) so.z[(size_t)w * so.ldz + j] = sweep_colp<true>(S, ld, j)[nc] / sqrt(sweep_colp<true>(S, ld, j)[j]);
  if (so.logdet != nullptr && tid == 0) so.logdet[w] = logdet;
  if (tid == 0) {
    double v = -((double)N / 2.0) * log(2.0 * MRV_PI) - 0.5 * logdet - 0.5 * quad;
    v = add_priors(D, hpv, phys, v);
    int st = 0;
    if (anybad) st = -1;
    else if (fb != 0x7fffffff) st = fb;
    logl_out[w] = v;
    status_out[w] = st;
  }
  SMALL_MARK(4)
}
