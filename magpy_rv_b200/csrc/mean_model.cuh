// Mean models, residuals and priors for one walker, evaluated cooperatively by one CTA.
//
// Reference: get_model (src/magpy_rv/mcmc_aux.py:52-97) sums the model curves in model_list
// order; Polynomial (models.py:370), Offset (models.py:467-479), Keplerian (models.py:618-697 with
// the whole-vector L1 Newton stop of SURVEY.md F6); residual y - model_y (gp_likelihood.py:244);
// priors summed in prior_list order on the parameter values AFTER the Sk,Ck -> ecc,omega
// conversion (mcmc_aux.py:84-88, auxiliary.py:206-211).
//
// The arithmetic uses the non-contracting intrinsics (__dmul_rn/__dadd_rn/...) so each NumPy
// ufunc step is mirrored by one correctly rounded operation (no FMA fusion).
#pragma once
#include "common.cuh"

// Convert the Keplerian Sk,Ck slots to ecc,omega in the walker's parameter copy (thread 0 only).
__device__ inline void convert_to_ecc(const ProblemDesc& D, double* phys, int to_ecc) {
  if (!to_ecc) return;
  for (int o = 0; o < D.n_ops; ++o) {
    if (D.ops[o].type != 3) continue;
    const int b = D.ops[o].param_offset;
    const double Sk = phys[b + 2], Ck = phys[b + 3];
    const double ecc = __dadd_rn(__dmul_rn(Sk, Sk), __dmul_rn(Ck, Ck));
    phys[b + 2] = ecc;
    phys[b + 3] = (ecc == 0.0) ? (MRV_PI / 2.0) : atan2(Sk, Ck);
  }
}

// Block-wide sum with ONE barrier per call (block_sum needs two): the warp partials go to scratch[buf][warp] and,
// after the barrier, every thread adds them in warp order.  The caller alternates `buf`, so the writes of call
// k + 2 are separated from the reads of call k by the barrier of call k + 1.  Fixed order => deterministic.
__device__ __forceinline__ double block_sum_1bar(double v, double* scratch /* 2 x 32 doubles */, int buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  double* s = scratch + 32 * buf;
  if (lane == 0) s[warp] = v;
  __syncthreads();
  // four interleaved partial sums (fixed association): the dependent-add chain is nwarps / 4 + 2 long instead of nwarps
  double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
  for (int w = 0; w + 3 < nwarps; w += 4) {
    r0 += s[w];
    r1 += s[w + 1];
    r2 += s[w + 2];
    r3 += s[w + 3];
  }
  for (int w = nwarps & ~3; w < nwarps; ++w) r0 += s[w];
  return (r0 + r1) + (r2 + r3);
}

// block_sum_1bar whose barrier also ANDs a per-thread predicate over the block (same single barrier).
__device__ __forceinline__ double block_sum_1bar_and(double v, double* scratch /* 2 x 32 doubles */, int buf, int pred, int* all) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  double* s = scratch + 32 * buf;
  if (lane == 0) s[warp] = v;
  *all = __syncthreads_and(pred);
  double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
  for (int w = 0; w + 3 < nwarps; w += 4) {
    r0 += s[w];
    r1 += s[w + 1];
    r2 += s[w + 2];
    r3 += s[w + 3];
  }
  for (int w = nwarps & ~3; w < nwarps; ++w) r0 += s[w];
  return (r0 + r1) + (r2 + r3);
}

// Newton loop of the eccentric anomaly with PPT points per thread held in registers (point i = tid + j * blockDim).
// The per-thread partial of ||E - E0||_1 is accumulated in increasing j, as the strided loop of the generic path does.
//
// The reference stops on the L1 norm of the WHOLE vector (models.py:643); with |E| in the hundreds and 500
// points the norm stalls at a few ulp per point above the 1e-10 threshold and the loop runs all 200 iterations although
// every point settled after ~6.  The iterates are then EXACTLY periodic: a point whose new value equals its current
// one is a fixed point, one whose new value equals its previous one alternates between two values (a rounding 2-cycle).
// Once every point of the vector is one or the other (one __syncthreads_and inside the reduction's barrier), every
// later iteration repeats the same partial sums, hence the same norm: the loop can never stop before max_itr, and the
// vector after iteration 199 is known -- fixed points as they are, 2-cycle points by the parity of the remaining
// count.  Bitwise the result of running the remaining iterations (NaN never compares equal, so those vectors still
// iterate to the end; so do vectors with a longer rounding cycle).
template <int PPT>
__device__ __forceinline__ void kepler_newton_regs(int N, const double* __restrict__ t, double P, double ecc, double t0,
                                                   double* E, double* kep_red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const double two_pi = 2.0 * MRV_PI;
  double M[PPT], E0[PPT], Ep[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int i = tid + j * nt;
    M[j] = i < N ? __ddiv_rn(__dmul_rn(two_pi, __dsub_rn(t[i], t0)), P) : 0.0;
    E0[j] = M[j];
    Ep[j] = __longlong_as_double(0x7ff8000000000000ll);   // no previous iterate yet
  }
  for (int it = 0; it < 200; ++it) {
    double part = 0.0;
    int periodic = 1;
    unsigned cyc = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (tid + j * nt < N) {
        double s, c;
        sincos(E0[j], &s, &c);
        const double f = __dsub_rn(__dsub_rn(E0[j], __dmul_rn(ecc, s)), M[j]);
        const double fp = __dsub_rn(1.0, __dmul_rn(ecc, c));
        const double En = __dsub_rn(E0[j], __ddiv_rn(f, fp));
        part += fabs(__dsub_rn(En, E0[j]));
        const bool fixed = En == E0[j], two = En == Ep[j];
        periodic &= (fixed || two) ? 1 : 0;
        if (two && !fixed) cyc |= 1u << j;
        Ep[j] = E0[j];
        E0[j] = En;
      }
    }
    int all_periodic;
    const double l1 = block_sum_1bar_and(part, kep_red, it & 1, periodic, &all_periodic);   // whole-vector ||E - E0||_1, models.py:643
    if (l1 <= 1.0e-10) break;                 // NaN compares false -> keeps iterating, as NumPy does
    if (all_periodic) {
      // every later iteration reproduces this norm (> 1e-10): the loop runs to it = 199; 2-cycle points end on the
      // other value when the number of remaining iterations is odd
      if ((199 - it) & 1) {
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (cyc & (1u << j)) E0[j] = Ep[j];
      }
      break;
    }
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j)
    if (tid + j * nt < N) E[tid + j * nt] = E0[j];
  __syncthreads();
}

// One Keplerian curve added into model[] (length N).  E is an N-double scratch array.
__device__ inline void cta_keplerian(int N, const double* __restrict__ t, double P, double K, double ecc,
                                     double omega, double t0, double* model, double* E, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const double two_pi = 2.0 * MRV_PI;
  __shared__ double kep_red[64];   // the Newton loop is a chain of up to 200 block-wide sums: one barrier each
  (void)red;
  // M = 2*pi*(t - t0)/P, E0 = M (not reduced mod 2pi), models.py:686, 635-636
  if (N <= 4 * nt) {
    // up to four points per thread: M and E stay in registers over the whole Newton loop (same operations, same order)
    if (N <= nt) kepler_newton_regs<1>(N, t, P, ecc, t0, E, kep_red);
    else if (N <= 2 * nt) kepler_newton_regs<2>(N, t, P, ecc, t0, E, kep_red);
    else kepler_newton_regs<4>(N, t, P, ecc, t0, E, kep_red);
  } else {   // generic path: E in the scratch array
    for (int i = tid; i < N; i += nt) E[i] = __ddiv_rn(__dmul_rn(two_pi, __dsub_rn(t[i], t0)), P);
    __syncthreads();
    for (int it = 0; it < 200; ++it) {
      double part = 0.0;
      for (int i = tid; i < N; i += nt) {
        const double M = __ddiv_rn(__dmul_rn(two_pi, __dsub_rn(t[i], t0)), P);
        const double E0 = E[i];
        double s, c;
        sincos(E0, &s, &c);
        const double f = __dsub_rn(__dsub_rn(E0, __dmul_rn(ecc, s)), M);
        const double fp = __dsub_rn(1.0, __dmul_rn(ecc, c));
        const double En = __dsub_rn(E0, __ddiv_rn(f, fp));
        E[i] = En;
        part += fabs(__dsub_rn(En, E0));
      }
      const double l1 = block_sum_1bar(part, kep_red, it & 1);   // whole-vector ||E - E0||_1, models.py:643
      if (l1 <= 1.0e-10) break;                 // NaN compares false -> keeps iterating, as NumPy does
    }
    __syncthreads();
  }
  const double fac = sqrt(__ddiv_rn(__dadd_rn(1.0, ecc), __dsub_rn(1.0, ecc)));
  const double ecw = __dmul_rn(ecc, cos(omega));
  for (int i = tid; i < N; i += nt) {
    const double nu = __dmul_rn(2.0, atan(__dmul_rn(fac, tan(__ddiv_rn(E[i], 2.0)))));
    const double rv = __dmul_rn(K, __dadd_rn(cos(__dadd_rn(omega, nu)), ecw));
    model[i] = __dadd_rn(model[i], rv);
  }
  __syncthreads();
}

// model[] <- sum of model curves (in op order).  phys[] = walker's model parameters (already
// converted).  flags may be null when no OFFSET op exists.
__device__ inline void cta_mean_model(const ProblemDesc& D, const double* __restrict__ t,
                                      const double* __restrict__ flags, const double* phys, double* model,
                                      double* E, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x, N = D.N;
  for (int i = tid; i < N; i += nt) model[i] = 0.0;
  __syncthreads();
  for (int o = 0; o < D.n_ops; ++o) {
    const int b = D.ops[o].param_offset;
    switch (D.ops[o].type) {
      case 0:  // No_Model: zeros
        break;
      case 1: {  // Polynomial: a0 + a1*t + a2*t**2 + a3*t**3, left to right
        const double a0 = phys[b], a1 = phys[b + 1], a2 = phys[b + 2], a3 = phys[b + 3];
        for (int i = tid; i < N; i += nt) {
          const double x = t[i];
          double v = __dadd_rn(a0, __dmul_rn(a1, x));
          v = __dadd_rn(v, __dmul_rn(a2, __dmul_rn(x, x)));
          v = __dadd_rn(v, __dmul_rn(a3, pow(x, 3.0)));
          model[i] = __dadd_rn(model[i], v);
        }
        break;
      }
      case 2: {  // Offset
        const double off = phys[b], fv = D.ops[o].flag_val;
        for (int i = tid; i < N; i += nt) model[i] = __dadd_rn(model[i], (flags[i] == fv) ? off : 0.0);
        break;
      }
      case 3:
        __syncthreads();
        cta_keplerian(N, t, phys[b], phys[b + 1], phys[b + 2], phys[b + 3], phys[b + 4], model, E, red);
        break;
    }
    __syncthreads();
  }
}

// Sum of log-priors in prior_list order, added sequentially onto `base` (gp_likelihood.py:392-400).
__device__ inline double add_priors(const ProblemDesc& D, const double* hp, const double* phys, double base) {
  double v = base;
  for (int i = 0; i < D.n_priors; ++i) {
    const PriorSpec& s = D.priors[i];
    const double x = (s.source == 0) ? hp[s.index] : phys[s.index];
    v = __dadd_rn(v, prior_logprob(s, x));
  }
  return v;
}

// Standalone kernel for the blocked path and mrv_model_eval: one CTA per walker.
//   resid_out [W, ldr]   y - model (or the model itself when y == nullptr); entries >= N zeroed up to ldr
//   phys_out  [W, nm]    converted parameters (nullable)
//   nonfinite [W]        1 when any residual is non-finite (nullable)
__global__ void mean_residual_kernel(ProblemDesc D, const double* __restrict__ t, const double* __restrict__ y,
                                     const double* __restrict__ flags, const double* __restrict__ modpar,
                                     const double* __restrict__ model_given, int to_ecc, double* resid_out,
                                     int ldr, double* E_scratch, double* phys_out, int* nonfinite) {
  __shared__ double phys[MRV_MAX_MODPAR];
  __shared__ double red[32];
  __shared__ int ired[32];
  const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, N = D.N;
  double* r = resid_out + (size_t)w * ldr;
  if (tid == 0) {
    for (int i = 0; i < D.nm; ++i) phys[i] = modpar[(size_t)w * D.nm + i];
    if (model_given == nullptr) convert_to_ecc(D, phys, to_ecc);
  }
  __syncthreads();
  if (phys_out != nullptr && tid < D.nm) phys_out[(size_t)w * D.nm + tid] = phys[tid];
  if (model_given != nullptr) {
    for (int i = tid; i < N; i += nt) r[i] = model_given[(size_t)w * N + i];
  } else {
    cta_mean_model(D, t, flags, phys, r, E_scratch + (size_t)w * ldr, red);
  }
  __syncthreads();
  int bad = 0;
  if (y != nullptr) {
    for (int i = tid; i < N; i += nt) {
      const double v = __dsub_rn(y[i], r[i]);
      r[i] = v;
      bad |= !isfinite(v);
    }
  } else {
    for (int i = tid; i < N; i += nt) bad |= !isfinite(r[i]);
  }
  for (int i = N + tid; i < ldr; i += nt) r[i] = 0.0;
  if (nonfinite != nullptr) {
    bad = block_max_int(bad, ired);
    if (tid == 0) nonfinite[w] = bad;
  }
}

// Keplerian.ecc_anomaly(M, ecc, max_itr) for a caller-supplied mean-anomaly vector (reference models.py:618-650):
// the same whole-vector Newton iteration as cta_keplerian, one CTA, E_out is the working array.
__global__ void kepler_anomaly_kernel(const double* __restrict__ M, int N, double ecc, int max_itr, double* E) {
  __shared__ double red[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N; i += nt) E[i] = M[i];
  __syncthreads();
  for (int it = 0; it < max_itr; ++it) {
    double part = 0.0;
    for (int i = tid; i < N; i += nt) {
      const double E0 = E[i];
      double s, c;
      sincos(E0, &s, &c);
      const double f = __dsub_rn(__dsub_rn(E0, __dmul_rn(ecc, s)), M[i]);
      const double fp = __dsub_rn(1.0, __dmul_rn(ecc, c));
      const double En = __dsub_rn(E0, __ddiv_rn(f, fp));
      E[i] = En;
      part += fabs(__dsub_rn(En, E0));
    }
    const double l1 = block_sum_1bar(part, red, it & 1);
    if (l1 <= 1.0e-10) break;
  }
}

// Keplerian.true_anomaly(E, ecc) = 2 atan(sqrt((1 + ecc) / (1 - ecc)) tan(E / 2)) (reference models.py:653-669).
__global__ void true_anomaly_kernel(const double* __restrict__ E, int N, double ecc, double* nu) {
  const double fac = sqrt(__ddiv_rn(__dadd_rn(1.0, ecc), __dsub_rn(1.0, ecc)));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
    nu[i] = __dmul_rn(2.0, atan(__dmul_rn(fac, tan(__ddiv_rn(E[i], 2.0)))));
}
