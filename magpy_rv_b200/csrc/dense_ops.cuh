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
