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

// One whole split of the stretch move on the ensemble arrays themselves: chain c is walker own[c], stretches
// towards walker partner[c] and is stored at row dest[c] of the proposal arrays (reference mcmc.py:343-394;
// parameters with vary == False keep the chain's own value; rows are written in chain order, so a later
// chain overwrites an earlier one with the same destination, as the reference's loop does).  Same scan,
// refill protocol and arithmetic as mrv_stretch_scan; z_out is indexed by chain.
extern "C" int64_t mrv_stretch_split(const double* u, int64_t n_u, int64_t chain_begin, int64_t n_chains, int32_t nh,
                                     int32_t nm, const double* hp0, const double* mod0, const int64_t* own,
                                     const int64_t* partner, const int64_t* dest, const uint8_t* hp_vary,
                                     const uint8_t* mod_vary, double a, const int32_t* kepler_offsets,
                                     int32_t n_kepler, int64_t* chains_done, double* z_out, double* hp_out,
                                     double* mod_out) {
  enum { MAXP = 64 };
  if (nh > MAXP || nm > MAXP) return -1;
  int64_t used = 0, c = chain_begin;
  const double am1 = a - 1.0;
  volatile double two = 2.0;
  double hp[MAXP], md[MAXP], dh[MAXP], dm[MAXP];
  for (; c < n_chains; ++c) {
    const double* xk_h = hp0 + own[c] * nh;
    const double* xj_h = hp0 + partner[c] * nh;
    const double* xk_m = mod0 + own[c] * nm;
    const double* xj_m = mod0 + partner[c] * nm;
    for (int i = 0; i < nh; ++i) dh[i] = xk_h[i] - xj_h[i];
    for (int i = 0; i < nm; ++i) dm[i] = xk_m[i] - xj_m[i];
    int64_t pos = used;
    bool ok = false;
    while (pos < n_u) {
      volatile double s = u[pos] * am1;
      const double base = s + 1.0;
      const double z = pow(base, two) / a;
      ++pos;
      for (int i = 0; i < nh; ++i) {
        volatile double prod = dh[i] * z;
        hp[i] = xj_h[i] + prod;
      }
      for (int i = 0; i < nm; ++i) {
        volatile double prod = dm[i] * z;
        md[i] = xj_m[i] + prod;
      }
      if (mrv_proposal_valid(hp, nh, md, kepler_offsets, n_kepler)) {
        z_out[c] = z;
        ok = true;
        break;
      }
    }
    if (!ok) break;
    used = pos;
    double* oh = hp_out + dest[c] * nh;
    double* om = mod_out + dest[c] * nm;
    for (int i = 0; i < nh; ++i) oh[i] = hp_vary[i] ? hp[i] : xk_h[i];
    for (int i = 0; i < nm; ++i) om[i] = mod_vary[i] ? md[i] : xk_m[i];
  }
  *chains_done = c;
  return used;
}
