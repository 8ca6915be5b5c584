// Dense element-wise maps: pairwise distances and covariance matrices written to HBM.
// These serve the reference's public per-object API (compute_distances kernels.py:51-78,
// Kernel.compute_covmatrix kernels.py:205.., GPLikelihood.compute_kernel gp_likelihood.py:173-230);
// the batched likelihood path never materialises these matrices.  HBM-bound: 8 B written per
// element (16 B read + 8 B written for the from-dist variant); grid-stride, coalesced along n2.
#pragma once
#include "common.cuh"

__global__ void distances_kernel(const double* __restrict__ t1, long long n1, const double* __restrict__ t2,
                                 long long n2, double* __restrict__ de, double* __restrict__ dse) {
  const long long total = n1 * n2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / n2, j = e - i * n2;
    const double dt = t1[i] - t2[j];
    // scipy cdist on one-dimensional points: sqeuclidean = dt^2, euclidean = sqrt(dt^2) = |dt|
    de[e] = fabs(dt);
    dse[e] = dt * dt;
  }
}

// mode 0: from x1/x2; mode 1: from dense dist_e/dist_se.
__global__ void covmatrix_kernel(CovParams cp, int from_dist, const double* __restrict__ x1, long long n1,
                                 const double* __restrict__ x2, long long n2, const double* __restrict__ dist_e,
                                 const double* __restrict__ dist_se, int diag_mode,
                                 const double* __restrict__ diag_add, double* __restrict__ K) {
  const long long total = n1 * n2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / n2, j = e - i * n2;
    double v;
    if (from_dist) {
      v = cov_from_dist(cp, dist_e[e], dist_se[e]);
    } else {
      const double dt = x1[i] - x2[j];
      v = cov_from_dist(cp, fabs(dt), dt * dt);
    }
    if (i == j && n1 == n2) {
      if (diag_mode & 1) v = v + cp.jit2;
      if (diag_mode & 2) v = v + diag_add[j];
    }
    K[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Gelman-Rubin statistic per parameter over the walker chains.  The
// reference's gelman_rubin_calc (src/magpy_rv/mcmc.py:571-662) mislabels the axes of the [steps,
// walkers, parameters] history and drops into pdb; this is the statistic it intends:
//   chain mean m_j, chain variance s_j^2 (ddof 1) over the L = steps - burn_in retained steps,
//   B = L / (J - 1) sum_j (m_j - mean(m))^2,  W = mean_j s_j^2,  R = ((1 - 1/L) W + B / L) / W,
// and R = 1 for a parameter that never moves (mcmc.py:632-638).  One CTA per parameter, one thread
// per chain (strided), fixed-order reductions => deterministic.
// ------------------------------------------------------------------------------------------------
__device__ inline double block_max_f64(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = -INFINITY;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) r = fmax(r, scratch[w]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256)
gelman_rubin_kernel(const double* __restrict__ chains, long long n_steps, long long n_chains, long long n_par,
                    long long burn_in, double* __restrict__ R_out) {
  __shared__ double red[32];
  const long long p = blockIdx.x;
  const long long L = n_steps - burn_in;
  double sum_m = 0.0, sum_v = 0.0, vmin = INFINITY, vmax = -INFINITY;
  for (long long j = threadIdx.x; j < n_chains; j += blockDim.x) {
    const double* c = chains + (burn_in * n_chains + j) * n_par + p;
    double m = 0.0;
    for (long long s = 0; s < L; ++s) m += c[s * n_chains * n_par];
    m /= (double)L;
    double v = 0.0;
    for (long long s = 0; s < L; ++s) {
      const double x = c[s * n_chains * n_par];
      v = fma(x - m, x - m, v);
      vmin = fmin(vmin, x);
      vmax = fmax(vmax, x);
    }
    sum_m += m;
    sum_v += v / (double)(L - 1);
  }
  const double grand = block_sum(sum_m, red) / (double)n_chains;
  const double W = block_sum(sum_v, red) / (double)n_chains;
  double dev = 0.0;
  for (long long j = threadIdx.x; j < n_chains; j += blockDim.x) {
    const double* c = chains + (burn_in * n_chains + j) * n_par + p;
    double m = 0.0;
    for (long long s = 0; s < L; ++s) m += c[s * n_chains * n_par];
    m /= (double)L;
    dev = fma(m - grand, m - grand, dev);
  }
  const double B = (double)L / (double)(n_chains - 1) * block_sum(dev, red);
  const double lo = -block_max_f64(-vmin, red);
  const double hi = block_max_f64(vmax, red);
  if (threadIdx.x == 0) {
    double R = ((1.0 - 1.0 / (double)L) * W + B / (double)L) / W;
    if (hi - lo == 0.0) R = 1.0;
    R_out[p] = R;
  }
}
