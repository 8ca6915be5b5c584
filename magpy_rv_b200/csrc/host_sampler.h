// Host-side helper of the ensemble sampler (no GPU work): the sequential part of one stretch-move
// split, MCMC.split_step (reference src/magpy_rv/mcmc.py:322-376).
//
// Chain c of the split consumes the np.random uniform stream in order: z = (u (a - 1) + 1)^2 / a,
// proposal = Xj + (Xk - Xj) z, and it KEEPS consuming while the proposal is invalid -- a negative
// hyper-parameter (`np.min(Xk_new) < 0`: a NaN makes np.min NaN and the test False) or a Keplerian
// with P, K or t0 < 0 or Sk^2 + Ck^2 > 1 (mcmc_aux.parameter_check, mcmc_aux.py:187-231).  The
// number of uniforms each chain sees therefore depends on all chains before it, which is why this
// scan is sequential; in Python it cost ~0.2 ms per redraw event (5 ms per iteration at 256 chains
// with a Keplerian).  Arithmetic mirrors the reference's scalar operations one for one (separate
// multiply and add, libm pow(x, 2.0) as Python's float ** 2): build WITHOUT fp contraction.
#pragma once
#include <math.h>
#include <stdint.h>

static inline bool mrv_proposal_valid(const double* hp, int nh, const double* mod, const int32_t* kep, int n_kep) {
  bool has_nan = false, neg = false;
  for (int i = 0; i < nh; ++i) {
    if (isnan(hp[i])) has_nan = true;
    else if (hp[i] < 0.0) neg = true;
  }
  if (neg && !has_nan) return false;
  for (int k = 0; k < n_kep; ++k) {
    const double* m = mod + kep[k];
    if (m[0] < 0.0 || m[1] < 0.0 || m[4] < 0.0) return false;
    const double s2 = m[2] * m[2], c2 = m[3] * m[3];
    if (s2 + c2 > 1.0) return false;
  }
  return true;
}

extern "C" int64_t mrv_stretch_scan(const double* u, int64_t n_u, int64_t chain_begin, int64_t n_chains, int32_t nh,
                                    int32_t nm, const double* Xj_hp, const double* dX_hp, const double* Xj_mod,
                                    const double* dX_mod, double a, const int32_t* kepler_offsets, int32_t n_kepler,
                                    int64_t* chains_done, double* z_out, double* new_hp, double* new_mod) {
  int64_t used = 0, c = chain_begin;
  const double am1 = a - 1.0;
  volatile double two = 2.0;
  for (; c < n_chains; ++c) {
    int64_t pos = used;
    bool ok = false;
    double* hp = new_hp + c * nh;
    double* md = new_mod + c * nm;
    while (pos < n_u) {
      volatile double s = u[pos] * am1;   // volatile: keep multiply and add separate operations
      const double base = s + 1.0;
      const double z = pow(base, two) / a;   // libm pow, as Python's float ** 2 (not folded into base * base)
      ++pos;
      for (int i = 0; i < nh; ++i) {
        volatile double prod = dX_hp[c * nh + i] * z;
        hp[i] = Xj_hp[c * nh + i] + prod;
      }
      for (int i = 0; i < nm; ++i) {
        volatile double prod = dX_mod[c * nm + i] * z;
        md[i] = Xj_mod[c * nm + i] + prod;
      }
      if (mrv_proposal_valid(hp, nh, md, kepler_offsets, n_kepler)) {
        z_out[c] = z;
        ok = true;
        break;
      }
    }
    if (!ok) break;   // ran out of uniforms inside this chain: its draws are not counted
    used = pos;
  }
  *chains_done = c;
  return used;
}
