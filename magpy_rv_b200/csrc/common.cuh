// Shared device helpers: covariance functions, priors, block reductions.
// All arithmetic is IEEE binary64; no fast-math anywhere (compile WITHOUT --use_fast_math).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define MRV_PI 3.141592653589793
#define MRV_MAX_HP 8
#define MRV_MAX_OPS 16
#define MRV_MAX_MODPAR 48
#define MRV_MAX_PRIORS 48
#define MRV_STATUS_INTERNAL (-3)   // per-walker status: internal error (a dataflow dependency never arrived); never from valid input

struct ModelOp {
  int type;
  int param_offset;
  double flag_val;
};
struct PriorSpec {
  int type, source, index, pad;
  double p0, p1, p2;
};

// Optional extra outputs / input mode of a batched evaluation (mrv_solve_batch).
struct SolveOut {
  double* z;           // [W, ldz]  L^-1 r per walker (nullable)
  long long ldz;
  double* logdet;      // [W] ln det K (nullable)
  int given_is_resid;  // the caller-supplied [W, N] array is the right-hand side itself, not a model curve
};

// Everything a kernel needs to know about the problem besides the per-walker parameters.
struct ProblemDesc {
  int N;             // data points
  int kernel_id;     // MRV_KERNEL_*
  int variant;       // 0 reference-as-coded, 1 textbook Matern5
  int nh, nm;        // hyper-parameters / model parameters per walker
  int n_ops, n_priors;
  double t_ref;      // phase origin of the periodic covariance terms (a time stamp from the middle of the series)
  ModelOp ops[MRV_MAX_OPS];
  PriorSpec priors[MRV_MAX_PRIORS];
};

// ------------------------------------------------------------------------------------------------
// Covariance functions (reference src/magpy_rv/kernels.py:229, 356, 489, 629-630, 774-775, 909,
// 1042).  CovParams holds per-walker constants so the N^2/2 evaluations are a handful of fp64 ops
// plus the transcendental calls.  Constants fold the reference's divisions into reciprocals
// (<= 1 ulp change per entry, far inside the 1e-9 LogL tolerance) and the two QuasiPer exps into one.
// The periodic terms use sinpi / cospi of d / per instead of sin / cos of pi d / per: the range reduction is
// then exact (no Cody-Waite / Payne-Hanek step) and the evaluation 13 % cheaper (tools/micro/gen.cu: 1.08 ->
// 0.94 SM-clocks at 16 warps per SM); the argument carries the same one-ulp rounding as the reference's.
// ------------------------------------------------------------------------------------------------
struct CovParams {
  int kind;      // kernel_id (Matern5 textbook -> 7)
  double a2;     // amp^2
  double c1, c2, c3;
  double jit2;   // jitter^2 (0 when the kernel has none)
  double cph;    // periodic kernels: 1 / per (phase of a point = cph * (t - t_ref), in periods), else 0
};

__host__ __device__ inline CovParams make_cov_params(int kernel_id, int variant, const double* hp) {
  CovParams p;
  p.kind = kernel_id;
  p.a2 = p.c1 = p.c2 = p.c3 = p.jit2 = p.cph = 0.0;
  switch (kernel_id) {
    case 0:  // Cosine: hp = amp, per
      p.a2 = hp[0] * hp[0];
      p.c1 = 2.0 / hp[1];          // cospi(c1 d)
      p.cph = 1.0 / hp[1];
      break;
    case 1:  // ExpSquared: amp, timescale
      p.a2 = hp[0] * hp[0];
      p.c1 = -0.5 / (hp[1] * hp[1]);
      break;
    case 2:  // ExpSinSquared: amp, timescale, per
      p.a2 = hp[0] * hp[0];
      p.c1 = -2.0 / (hp[1] * hp[1]);
      p.c2 = 1.0 / hp[2];          // sinpi(c2 d)
      p.cph = p.c2;
      break;
    case 3:  // QuasiPer: per, perlength, explength, amp
    case 4:  // JitterQuasiPer: + jit
      p.a2 = hp[3] * hp[3];
      p.c1 = -1.0 / (hp[2] * hp[2]);
      p.c2 = 1.0 / hp[0];          // sinpi(c2 d)
      p.c3 = -1.0 / (hp[1] * hp[1]);
      p.cph = p.c2;
      if (kernel_id == 4) p.jit2 = hp[4] * hp[4];
      break;
    case 5:  // Matern5/2: amp, timescale
      p.a2 = hp[0] * hp[0];
      p.c1 = sqrt(5.0) / hp[1];
      if (variant == 1) {
        p.kind = 7;
        p.c2 = 5.0 / (3.0 * hp[1] * hp[1]);
      } else {
        p.c2 = hp[1] * hp[1];  // as coded: (5 d^4 / 3) * ts^2
      }
      break;
    case 6:  // Matern3/2: amp, timescale, jit
      p.a2 = hp[0] * hp[0];
      p.c1 = sqrt(3.0) / hp[1];
      p.jit2 = hp[2] * hp[2];
      break;
  }
  return p;
}

// k(d) for d = |t - t'| >= 0 and d2 = (t - t')^2 (off-diagonal part; no error/jitter terms).
__device__ __forceinline__ double cov_from_dist(const CovParams& p, double d, double d2) {
  switch (p.kind) {
    case 0:
      return p.a2 * cospi(p.c1 * d);
    case 1:
      return p.a2 * exp(p.c1 * (d * d));
    case 2: {
      double s = sinpi(p.c2 * d);
      return p.a2 * exp(p.c1 * (s * s));
    }
    case 3:
    case 4: {
      double s = sinpi(p.c2 * d);
      return p.a2 * exp(p.c1 * d2 + p.c3 * (s * s));
    }
    case 5: {  // reference as coded (kernels.py:909)
      double x = p.c1 * d;
      return p.a2 * (1.0 + x + ((5.0 * (d2 * d2)) / 3.0 * p.c2) * exp(-x));
    }
    case 7: {  // textbook Matern 5/2
      double x = p.c1 * d;
      return p.a2 * (1.0 + x + p.c2 * d2) * exp(-x);
    }
    case 6: {
      double x = p.c1 * d;
      return p.a2 * (1.0 + x) * exp(-x);
    }
  }
  return 0.0;
}

// NV independent evaluations with the kernel switch hoisted OUTSIDE the element loop, so the NV
// fp64 exp / sin dependency chains are interleaved by the compiler (instruction-level parallelism;
// a lone evaluation is a ~500-cycle serial chain on a pipe that could start a new op every 2).
template <int NV>
__device__ __forceinline__ void cov_from_dist_n(const CovParams& p, const double (&d)[NV], const double (&d2)[NV],
                                                double (&out)[NV]) {
  switch (p.kind) {
    case 0:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * cospi(p.c1 * d[e]);
      break;
    case 1:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * exp(p.c1 * (d[e] * d[e]));
      break;
    case 2: {
      double s[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) s[e] = sinpi(p.c2 * d[e]);
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * exp(p.c1 * (s[e] * s[e]));
      break;
    }
    case 3:
    case 4: {
      double s[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) s[e] = sinpi(p.c2 * d[e]);
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * exp(p.c1 * d2[e] + p.c3 * (s[e] * s[e]));
      break;
    }
    case 5:
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double x = p.c1 * d[e];
        out[e] = p.a2 * (1.0 + x + ((5.0 * (d2[e] * d2[e])) / 3.0 * p.c2) * exp(-x));
      }
      break;
    case 7:
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double x = p.c1 * d[e];
        out[e] = p.a2 * (1.0 + x + p.c2 * d2[e]) * exp(-x);
      }
      break;
    case 6:
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double x = p.c1 * d[e];
        out[e] = p.a2 * (1.0 + x) * exp(-x);
      }
      break;
    default:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Fast evaluation for the factorisation kernels (the N^2 / 2 entries of every walker): periodic terms by the
// angle-difference identity on PER-POINT phases.  With u_i = (t_i - t_ref) / per, C_i = cospi(2 u_i),
// S_i = sinpi(2 u_i) (N sincospi per walker instead of N^2 / 2 sinpi),
//        cos(2 pi d / per) = C_i C_j + S_i S_j,       sin^2(pi d / per) = (1 - cos(2 pi d / per)) / 2.
// The periodic term enters the QuasiPer / ExpSinSquared entries only through the EXPONENT, where an absolute error
// of ~1e-16 is a relative error of ~1e-16 of the entry; the phase of a point carries the rounding of u_i (|u_i| eps
// periods, i.e. a shift of the time stamp by 1 ulp of t - t_ref), the same size as the rounding of pi d / per in the
// reference for far-apart pairs.  Measured (tools/cov_rate_probe.py): QuasiPer 98 -> 44 FP64 flop-equivalents per entry.
// The dense per-object API (mrv_covmatrix*) and the diagonal keep the distance form above, which follows the
// reference's formula literally.
// ------------------------------------------------------------------------------------------------
// exp of NV values in LOCK STEP, branch-free: the library exp carries a rare-path branch per call, which cuts the
// NV evaluations into NV basic blocks that the compiler cannot interleave -- in the 8-warp tile kernels the exp chains
// then run one after the other at the latency of a dependent DFMA chain (ncu: the exp line was 12.5 % of all stall
// samples of the dataflow kernel, reason "wait"; generation 9.6 us per 128 x 128 tile against ~3 us of FP64 issue time).
// Range reduction by ln 2 (hi / lo), degree-11 polynomial (Chebyshev interpolant of exp on |r| <= ln 2 / 2, relative
// truncation error 4.2e-18; measured against expl: <= 1.5e-16 relative over [-700, 1]), scaling by 2^t built in the
// exponent field.  (A 32-entry table of 2^(j/32) with a degree-6 polynomial -- 12 FP64 instructions instead of 17 --
// measured no faster: the table load and the extra integer work cost what the five fmas save.)  Arguments below -700 are clamped (the result below 1e-304 is not flushed to zero); above +709 the result
// is inf (only reachable with negative time scales, whose covariance is not positive definite either way); NaN passes through.
template <int NV>
__device__ __forceinline__ void exp_lockstep(double (&x)[NV]) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the rounded integer sits in the low word
  double t[NV], p[NV];
  int ti[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) x[e] = (x[e] < -700.0) ? -700.0 : x[e];
#pragma unroll
  for (int e = 0; e < NV; ++e) t[e] = fma(x[e], 1.4426950408889634074, MAGIC);
  // large positive arguments: the power of two is capped at 2^1024 = inf in the exponent field (an integer min on the ALU
  // pipe instead of a second FP64 compare + select per value; the reduced argument stays exact far beyond 700)
#pragma unroll
  for (int e = 0; e < NV; ++e) ti[e] = min(__double2loint(t[e]), 1024);
#pragma unroll
  for (int e = 0; e < NV; ++e) t[e] -= MAGIC;
#pragma unroll
  for (int e = 0; e < NV; ++e) x[e] = fma(t[e], -6.93147180369123816490e-01, x[e]);
#pragma unroll
  for (int e = 0; e < NV; ++e) x[e] = fma(t[e], -1.90821492927058770002e-10, x[e]);
#pragma unroll
  for (int e = 0; e < NV; ++e) p[e] = fma(2.5110037605963777712e-8, x[e], 2.7632639639041029749e-7);
#define MRV_EXP_STEP(c)                \
  _Pragma("unroll") for (int e = 0; e < NV; ++e) p[e] = fma(p[e], x[e], c);
  MRV_EXP_STEP(2.7557240918578969823e-6)
  MRV_EXP_STEP(2.4801485482328492419e-5)
  MRV_EXP_STEP(1.9841269890047113707e-4)
  MRV_EXP_STEP(1.3888888952314774652e-3)
  MRV_EXP_STEP(8.3333333333196006109e-3)
  MRV_EXP_STEP(4.1666666666488095495e-2)
  MRV_EXP_STEP(1.6666666666666680806e-1)
  MRV_EXP_STEP(5.0000000000000183855e-1)
  MRV_EXP_STEP(1.0)
  MRV_EXP_STEP(1.0)
#undef MRV_EXP_STEP
#pragma unroll
  for (int e = 0; e < NV; ++e) x[e] = p[e] * __hiloint2double((1023 + ti[e]) << 20, 0);
}

// NV entries k(t_i, t_j) from per-point data: ti / tj times, (ci, si) / (cj, sj) the points' phase cosines / sines.
// LOCKSTEP: the exponentials through exp_lockstep (tile kernels: few warps, latency-bound) instead of the library exp
// (kernels with 16+ warps per SM, where the library function's shorter instruction count wins).
template <int NV, bool LOCKSTEP = false>
__device__ __forceinline__ void cov_pair_n(const CovParams& p, const double (&ti)[NV], const double (&ci)[NV],
                                           const double (&si)[NV], const double (&tj)[NV], const double (&cj)[NV],
                                           const double (&sj)[NV], double (&out)[NV]) {
  double arg[NV];
  switch (p.kind) {
    case 0:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * fma(ci[e], cj[e], si[e] * sj[e]);
      return;
    case 1:
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double dt = ti[e] - tj[e];
        arg[e] = p.c1 * (dt * dt);
      }
      break;
    case 2: {
      const double h = -0.5 * p.c1;   // c1 sin^2(pi d / per) = c1 (1 - cos) / 2 = h cos - h, one fma
#pragma unroll
      for (int e = 0; e < NV; ++e) arg[e] = fma(h, fma(ci[e], cj[e], si[e] * sj[e]), -h);
      break;
    }
    case 3:
    case 4: {
      const double h = -0.5 * p.c3;   // c3 sin^2(pi d / per) = h cos(2 pi d / per) - h
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double dt = ti[e] - tj[e];
        arg[e] = fma(p.c1, dt * dt, fma(h, fma(ci[e], cj[e], si[e] * sj[e]), -h));
      }
      break;
    }
    case 5:
    case 6:
    case 7:
#pragma unroll
      for (int e = 0; e < NV; ++e) arg[e] = -(p.c1 * fabs(ti[e] - tj[e]));
      break;
    default:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = 0.0;
      return;
  }
  if (LOCKSTEP) exp_lockstep<NV>(arg);
  else {
#pragma unroll
    for (int e = 0; e < NV; ++e) arg[e] = exp(arg[e]);
  }
  switch (p.kind) {
    case 5:   // reference as coded (kernels.py:909)
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double dt = ti[e] - tj[e], d2 = dt * dt;
        const double x = p.c1 * fabs(dt);
        out[e] = p.a2 * (1.0 + x + ((5.0 * (d2 * d2)) / 3.0 * p.c2) * arg[e]);
      }
      break;
    case 7:   // textbook Matern 5/2
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double dt = ti[e] - tj[e];
        const double x = p.c1 * fabs(dt);
        out[e] = p.a2 * (1.0 + x + p.c2 * (dt * dt)) * arg[e];
      }
      break;
    case 6:
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const double x = p.c1 * fabs(ti[e] - tj[e]);
        out[e] = p.a2 * (1.0 + x) * arg[e];
      }
      break;
    default:
#pragma unroll
      for (int e = 0; e < NV; ++e) out[e] = p.a2 * arg[e];
  }
}

// Phase of one point (only the periodic kernels read it).
__device__ __forceinline__ void cov_point_phase(const CovParams& p, double t_shifted, double& c, double& s) {
  if (p.cph != 0.0 || p.kind == 0 || (p.kind >= 2 && p.kind <= 4)) sincospi(2.0 * (p.cph * t_shifted), &s, &c);
  else {
    c = 1.0;
    s = 0.0;
  }
}

__device__ __forceinline__ double cov_eval(const CovParams& p, double ti, double tj) {
  double dt = ti - tj;
  return cov_from_dist(p, fabs(dt), dt * dt);
}

// Full K_ii: kernel value at zero lag, then + jit^2, then + yerr_i^2 -- the reference's order
// (kernels.py:778-789).
__device__ __forceinline__ double cov_diag(const CovParams& p, double yerr) {
  double k0 = cov_from_dist(p, 0.0, 0.0);
  return (k0 + p.jit2) + yerr * yerr;
}

// ------------------------------------------------------------------------------------------------
// Priors (reference src/magpy_rv/parameters.py:328, 395-402, 472-479, 539-544)
// ------------------------------------------------------------------------------------------------
__device__ inline double prior_logprob(const PriorSpec& s, double x) {
  switch (s.type) {
    case 0:  // Gaussian: -0.5((x-mu)/sigma)^2 - ln(sqrt(2pi) sigma)
    {
      double z = (x - s.p0) / s.p1;
      return -0.5 * (z * z) - log(sqrt(2.0 * MRV_PI) * s.p1);
    }
    case 1:  // Jeffrey
      if (x < s.p0 || x > s.p1) return -10e5;
      return log(1.0 / log(s.p1 / s.p0) * 1.0 / x);
    case 2:  // Modified Jeffrey
      if (x < s.p0 || x > s.p1) return -10e5;
      return log(1.0 / log((s.p1 + s.p2) / s.p2) * 1.0 / (x - s.p2));
    case 3:  // Uniform
      if (x < s.p0 || x > s.p1) return -10e5;
      return 0.0;
  }
  return 0.0;
}

// ------------------------------------------------------------------------------------------------
// Deterministic block-wide sum (fixed tree => bitwise reproducible, independent of shard count).
// `scratch` must hold >= 32 doubles.  All threads of the block must call; result on every thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ inline double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = (lane < nwarps) ? scratch[lane] : 0.0;
  r = warp_sum(r);
  return r;
}

__device__ inline int block_max_int(int v, int* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  int r = (lane < nwarps) ? scratch[lane] : INT_MIN;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = max(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}
