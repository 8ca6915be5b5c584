// Device-resident ensemble sampler (opt-in, NON-parity: SURVEY.md section 8(f) item 2).
//
// Same algorithm as the reference's MCMC.split_step / compute / compare / reset
// (src/magpy_rv/mcmc.py:286-567): every iteration ALL walkers propose a stretch move from the snapshot
// of the previous state (F3) against a random walker of a complementary split, redraw while the
// proposal is invalid (negative hyper-parameter, Keplerian with P, K, t0 < 0 or Sk^2 + Ck^2 > 1),
// the likelihoods are evaluated in one batch and each walker accepts if dlogL > 1, rejects if
// dlogL < -35 and otherwise accepts with probability exp(dlogL) (F4; the z^(n-1) factor is, as in the
// reference, not applied).  What differs is the random stream: counter-based Philox4x32-10 per
// (iteration, walker, draw) instead of the host's two Mersenne Twisters, so chains are reproducible
// for a seed but NOT comparable step by step with the reference -- the host sampler (mcmc.py) remains
// the parity path.  Nothing crosses PCIe inside the loop: the iteration cost is the likelihood kernel.
#pragma once
#include "common.cuh"

struct Philox {
  unsigned long long seed;
  __device__ __forceinline__ static void round(unsigned (&c)[4], unsigned k0, unsigned k1) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k0, n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k1, n3 = (unsigned)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // two uniforms in [0, 1) with 53 bits each for the counter (iteration, walker, draw, stream)
  __device__ __forceinline__ void uniform2(unsigned it, unsigned walker, unsigned draw, unsigned stream, double& u0,
                                           double& u1) const {
    unsigned c[4] = {it, walker, draw, stream};
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1], b = ((unsigned long long)c[2] << 32) | c[3];
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  }
};

struct SamplerArgs {
  int W, nh, nm, n_kep;
  int kep[MRV_MAX_OPS];
  unsigned char hp_vary[MRV_MAX_HP], mp_vary[MRV_MAX_MODPAR];
  double a;
  Philox rng;
};

// One thread per walker: partner from a complementary split, stretch factor, validity redraws.
__global__ void stretch_propose_kernel(SamplerArgs A, unsigned it, const unsigned char* __restrict__ split,
                                       const double* __restrict__ hp0, const double* __restrict__ mp0,
                                       double* __restrict__ hp, double* __restrict__ mp, int* __restrict__ error_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.W) return;
  const unsigned char s = split[i];
  unsigned draw = 0;
  int j = 0;
  for (int tries = 0;; ++tries) {        // uniform choice among the walkers of the other splits
    double u0, u1;
    A.rng.uniform2(it, (unsigned)i, draw++, 1u, u0, u1);
    j = min(A.W - 1, (int)(u0 * A.W));
    if (split[j] != s) break;
    j = min(A.W - 1, (int)(u1 * A.W));
    if (split[j] != s) break;
    if (tries > 4096) {                  // only if a split holds (almost) every walker
      atomicExch(error_flag, 1);
      j = i;
      break;
    }
  }
  const double* Xk = hp0 + (size_t)i * A.nh;
  const double* Xj = hp0 + (size_t)j * A.nh;
  const double* Mk = mp0 + (size_t)i * A.nm;
  const double* Mj = mp0 + (size_t)j * A.nm;
  double nh_[MRV_MAX_HP], nm_[MRV_MAX_MODPAR];
  for (int tries = 0;; ++tries) {
    double u0, u1;
    A.rng.uniform2(it, (unsigned)i, draw++, 2u, u0, u1);
    const double b = u0 * (A.a - 1.0) + 1.0;
    const double z = b * b / A.a;
    bool ok = true, has_nan = false, neg = false;
    for (int q = 0; q < A.nh; ++q) {
      const double v = Xj[q] + (Xk[q] - Xj[q]) * z;
      nh_[q] = v;
      if (isnan(v)) has_nan = true;
      else if (v < 0.0) neg = true;
    }
    if (neg && !has_nan) ok = false;
    for (int q = 0; q < A.nm; ++q) nm_[q] = Mj[q] + (Mk[q] - Mj[q]) * z;
    for (int k = 0; k < A.n_kep && ok; ++k) {
      const double* m = nm_ + A.kep[k];
      if (m[0] < 0.0 || m[1] < 0.0 || m[4] < 0.0 || m[2] * m[2] + m[3] * m[3] > 1.0) ok = false;
    }
    if (ok) break;
    if (tries > 100000) {                // a walker with no valid stretch at all: keep its position
      atomicExch(error_flag, 2);
      for (int q = 0; q < A.nh; ++q) nh_[q] = Xk[q];
      for (int q = 0; q < A.nm; ++q) nm_[q] = Mk[q];
      break;
    }
  }
  for (int q = 0; q < A.nh; ++q) hp[(size_t)i * A.nh + q] = A.hp_vary[q] ? nh_[q] : Xk[q];
  for (int q = 0; q < A.nm; ++q) mp[(size_t)i * A.nm + q] = A.mp_vary[q] ? nm_[q] : Mk[q];
}

// Accept / reject + bookkeeping, one thread per walker.  Row `row` of the histories receives the
// decision (zeros for a NaN difference, as the reference records); accepted proposals become the state.
__global__ void stretch_accept_kernel(SamplerArgs A, unsigned it, long long row, const double* __restrict__ hp,
                                      const double* __restrict__ mp, const double* __restrict__ logl,
                                      const int* __restrict__ status, double* __restrict__ hp0, double* __restrict__ mp0,
                                      double* __restrict__ logl0, double* __restrict__ hp_hist,
                                      double* __restrict__ mp_hist, double* __restrict__ logl_hist,
                                      double* __restrict__ acc_hist, int* __restrict__ fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.W) return;
  if (status[i] != 0) {                  // the reference's cho_factor would have raised here
    if (atomicCAS(&fail[0], 0, 1) == 0) {
      fail[1] = (int)it;
      fail[2] = status[i];
    }
  }
  const double diff = logl[i] - logl0[i];
  int decision = 0;                      // 0: NaN difference (zeros recorded), 1 accept, 2 reject
  if (diff > 1.0) decision = 1;
  else if (diff < -35.0) decision = 2;
  else if (diff >= -35.0 && diff <= 1.0) {
    double u0, u1;
    A.rng.uniform2(it, (unsigned)i, 0u, 3u, u0, u1);
    decision = (u0 <= exp(diff)) ? 1 : 2;
  }
  double* hh = hp_hist + ((size_t)row * A.W + i) * A.nh;
  double* mh = mp_hist + ((size_t)row * A.W + i) * A.nm;
  for (int q = 0; q < A.nh; ++q) {
    const double v = decision == 1 ? hp[(size_t)i * A.nh + q] : (decision == 2 ? hp0[(size_t)i * A.nh + q] : 0.0);
    hh[q] = v;
    if (decision == 1) hp0[(size_t)i * A.nh + q] = v;
  }
  for (int q = 0; q < A.nm; ++q) {
    const double v = decision == 1 ? mp[(size_t)i * A.nm + q] : (decision == 2 ? mp0[(size_t)i * A.nm + q] : 0.0);
    mh[q] = v;
    if (decision == 1) mp0[(size_t)i * A.nm + q] = v;
  }
  logl_hist[(size_t)row * A.W + i] = decision == 1 ? logl[i] : (decision == 2 ? logl0[i] : 0.0);
  acc_hist[(size_t)row * A.W + i] = decision == 1 ? 1.0 : 0.0;
  if (decision == 1) logl0[i] = logl[i];
}
