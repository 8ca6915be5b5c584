// Mid-N log-likelihood (144 < N <= 512, tutorial-2 / 51-Peg sizes): ONE CTA PER WALKER, one launch.
//
// The dense covariance K is never written anywhere.  A walker's CTA runs a left-looking blocked
// Cholesky over block columns of 32 (panel j = rows [32 j, R) x 32 columns, R = N padded to 32):
//   1. every thread GENERATES its share of the panel of K from t / yerr straight into its DMMA
//      accumulator registers (as -K, like the large-N GEMM: the epilogue is a negate).  Half of the
//      fragments are generated one step early, while a single warp factors the previous diagonal
//      block, the other half while the first operand slices are in flight -- the two latency-bound
//      sections hide behind the FP64-ALU-bound one;
//   2. acc += L(rows, 0:32 j) * L(block-j rows, 0:32 j)^T on the FP64 tensor pipe: the already
//      factored block columns stream back from a per-CTA workspace that stays L2-resident
//      (<= 1.1 MB per CTA slot), k8 slices of ALL panel rows as two 1-D bulk copies (TMA engine)
//      into a 6-stage mbarrier ring -- no CTA-wide barrier in the loop.  The B operand (the 32 rows
//      of block j) is simply the first 32 rows of the same slice, so nothing is loaded twice; one
//      warp per slice also folds the residual update r_j -= L(block-j rows, k) z_k into its pass
//      over the slice (fixed warp / order => bitwise reproducible);
//   3. the panel goes to shared memory in operand order; ONE warp factors the 32 x 32 diagonal block
//      in registers (row per lane, column broadcast through shared memory), inverts the factor
//      (column per lane) and forms z_j = L_jj^-1 r_j;
//   4. X = P inv(L_jj)^T for the rows below the block runs on the tensor pipe again and goes from the
//      accumulator registers straight to the workspace in operand order [k4 chunk][row][4 columns],
//      i.e. exactly the DMMA fragment layout step 2 reads back.
//   logdet = sum ln d_k (pivots before the square root), quad = z^T z, then the closed form and the
//   priors (reference gp_likelihood.py:250-278, 357-402).  A non-positive pivot is reported like
//   LAPACK potrf info (index of the leading minor); NaN pivots give -1.
//
// CTAs are persistent (grid = min(W, SMs x resident CTAs)): slot b owns workspace b and walks the
// walkers b, b + grid, ...  Mean models / residual / priors are the same device functions as in the
// other paths (mean_model.cuh), so their arithmetic is bit-identical across paths.
#pragma once
#include "blocked.cuh"
#include "mean_model.cuh"

#define MID_NB 32
#define MID_STAGES 6
#define MID_LOOKAHEAD 4
#define MID_N_MAX 512
#define MID_INV_LD 36   // == 4 (mod 16): conflict-free B-fragment loads of inv(L_jj)

__host__ __device__ inline int mid_npad(int N) { return (N + MID_NB - 1) & ~(MID_NB - 1); }
// first element of block column p in a workspace slot: sum_{p' < p} (R - 32 p') * 32
__host__ __device__ inline size_t mid_colblk_base(int R, int p) {
  return (size_t)32 * p * R - (size_t)512 * p * (p - 1);
}
__host__ inline size_t mid_ws_doubles(int N) {
  const int R = mid_npad(N);
  return mid_colblk_base(R, R / MID_NB);
}
__host__ inline int mid_warps(int N) { return mid_npad(N) <= 256 ? 8 : 16; }
// ring (MID_STAGES k8 slices of at most R - 32 rows) and panel (8 chunks of 4 R + 8 doubles) share one region
__host__ __device__ inline size_t mid_buf_doubles(int R) {
  const size_t ring = (size_t)MID_STAGES * 8 * (R > 32 ? R - 32 : 32);
  const size_t panel = (size_t)8 * (4 * R + 8);
  return ring > panel ? ring : panel;
}
__host__ inline size_t mid_smem_bytes(int N) {
  const int R = mid_npad(N), nw = mid_warps(N);
  return (mid_buf_doubles(R) + 4 * (size_t)R + (size_t)nw * 32 + 32 * MID_INV_LD + 32 + MRV_MAX_MODPAR + MRV_MAX_HP + 32 +
          16 + 2 * MID_STAGES) * sizeof(double);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
mid_logl_kernel(ProblemDesc D, const double* __restrict__ t, const double* __restrict__ y,
                const double* __restrict__ yerr, const double* __restrict__ flags,
                const double* __restrict__ hp_all, const double* __restrict__ modpar_all,
                const double* __restrict__ model_given, int to_ecc, int W, double* __restrict__ logl_out,
                int* __restrict__ status_out, SolveOut so, double* __restrict__ workspace, size_t ws_stride) {
  extern __shared__ double smem[];
  constexpr int NT = NW * 32;
  constexpr int FW = NW - 1;                // the warp that factors the diagonal blocks (fewest fragments)
  const int N = D.N, R = mid_npad(N), nblk = R / MID_NB;
  double* buf = smem;                       // ring of operand slices  |  panel in operand order
  double* ts = buf + mid_buf_doubles(R);    // R   t (zero padded)
  double* r = ts + R;                       // R   residual (zero padded)
  double* dvec = r + R;                     // R   pivots
  double* zvec = dvec + R;                  // R   z = L^-1 r
  double* rpart = zvec + R;                 // NW x 32  per-warp partials of the residual update
  double* invL = rpart + NW * 32;           // 32 x MID_INV_LD  inverse of the current diagonal factor
  double* rj = invL + 32 * MID_INV_LD;      // 32  updated residual block
  double* phys = rj + 32;                   // MRV_MAX_MODPAR
  double* hpv = phys + MRV_MAX_MODPAR;      // MRV_MAX_HP
  double* red = hpv + MRV_MAX_HP;           // 32
  int* ired = (int*)(red + 32);             // 32 ints
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(red + 32 + 16);
  unsigned long long* empty_bar = full_bar + MID_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tig = lane & 3;
  const int stage_stride = (R > 32 ? R - 32 : 32) * 8;   // doubles per ring stage
  double* ws = workspace + (size_t)blockIdx.x * ws_stride;

  if (tid == 0) {
    for (int s = 0; s < MID_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async;\n" ::: "memory");
  }
  unsigned qg = 0;   // slices consumed by this CTA so far (uniform): ring position and mbarrier phase

  for (int w = blockIdx.x; w < W; w += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < D.nh; ++i) hpv[i] = hp_all[(size_t)w * D.nh + i];
      for (int i = 0; i < D.nm; ++i) phys[i] = modpar_all[(size_t)w * D.nm + i];
      if (model_given == nullptr) convert_to_ecc(D, phys, to_ecc);
    }
    for (int i = tid; i < R; i += NT) {
      ts[i] = i < N ? t[i] : 0.0;
      zvec[i] = 0.0;
      dvec[i] = 1.0;
    }
    __syncthreads();

    // residual (same device functions and operation order as the other paths)
    if (model_given != nullptr) {
      for (int i = tid; i < N; i += NT) r[i] = model_given[(size_t)w * N + i];
      __syncthreads();
    } else {
      cta_mean_model(D, ts, flags, phys, r, /*E scratch=*/buf, red);
    }
    int bad = 0;
    for (int i = tid; i < R; i += NT) {
      double v = 0.0;
      if (i < N) {
        v = so.given_is_resid ? r[i] : __dsub_rn(y[i], r[i]);
        bad |= !isfinite(v);
      }
      r[i] = v;
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");   // buf was touched through the generic proxy
    __syncthreads();

    const CovParams cp = make_cov_params(D.kernel_id, D.variant, hpv);
    int fb = 0x7fffffff, fb_nan = 0;   // first non-positive pivot (tracked by warp FW)

    // -K for the m16 row fragment f = warp + NW s of panel jj, into acc[s]
    double acc[2][4][4];
    auto gen_slot = [&](int jj, int s) {
      const int rw0 = MID_NB * jj, Fj = (R - rw0) >> 4, f = warp + NW * s;
      if (f < Fj) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int gi0 = rw0 + 16 * f + gq, gj0 = rw0 + 8 * ni + 2 * tig;
          double frag[4];
          gen_fragment4(cp, N, gi0, gj0, ts[gi0], ts[gi0 + 8], ts[gj0], ts[gj0 + 1], yerr, frag);
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[s][ni][v] = -frag[v];
        }
      } else {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[s][ni][v] = 0.0;
      }
    };

    for (int j = 0; j < nblk; ++j) {
      const int row0 = MID_NB * j, M = R - row0, F = M >> 4, nsl = 4 * j;
      const int CS = 4 * M + 8;   // panel chunk stride (doubles): +8 keeps the fragment stores conflict-free
      const unsigned chunk_bytes = (unsigned)M * 4u * (unsigned)sizeof(double);

      // slice q of this step = k8 columns [8 (q&3), 8 (q&3) + 8) of block column q >> 2, rows [row0, R)
      auto issue = [&](int q, unsigned qglob) {
        const int p = q >> 2, c0 = 2 * (q & 3);
        const int Rp = R - MID_NB * p;
        const double* src = ws + mid_colblk_base(R, p) + (size_t)c0 * Rp * 4 + (size_t)(row0 - MID_NB * p) * 4;
        const unsigned st = qglob % MID_STAGES, u = qglob / MID_STAGES;
        if (u > 0) mbar_wait(&empty_bar[st], (u - 1) & 1u);
        double* dst = buf + (size_t)st * stage_stride;
        mbar_expect_tx(&full_bar[st], 2u * chunk_bytes);
        bulk_g2s(dst, src, chunk_bytes, &full_bar[st]);
        bulk_g2s(dst + (size_t)M * 4, src + (size_t)Rp * 4, chunk_bytes, &full_bar[st]);
      };
      if (tid == 0) {
        for (int q = 0; q < MID_LOOKAHEAD && q < nsl; ++q) issue(q, qg + q);
      }

      // 1. second half of this panel's generation (the first half was generated during the previous
      //    step's diagonal-block factorisation; the factoring warp catches up on both halves here)
      if (j == 0 || warp == FW) gen_slot(j, 0);
      gen_slot(j, 1);

      // 2. stream the factored block columns: acc += A B^T, residual block update on the side
      double racc = 0.0;
      for (int q = 0; q < nsl; ++q, ++qg) {
        if (tid == 0 && q + MID_LOOKAHEAD < nsl) issue(q + MID_LOOKAHEAD, qg + MID_LOOKAHEAD);
        __syncwarp();
        const unsigned st = qg % MID_STAGES;
        mbar_wait(&full_bar[st], (qg / MID_STAGES) & 1u);
        const double* Sl = buf + (size_t)st * stage_stride;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double* base = Sl + (size_t)h * M * 4;
          double b[4];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) b[ni] = base[ni * 32 + lane];
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int f = warp + NW * s;
            if (f < F) {   // warp-uniform
              const double a0 = base[(2 * f) * 32 + lane], a1 = base[(2 * f + 1) * 32 + lane];
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) dmma_1684(acc[s][ni], a0, a1, b[ni]);
            }
          }
        }
        if ((q % NW) == warp) {   // r_j[lane] -= sum_k L[row0 + lane][k] z[k] over this slice's 8 columns
          const int kbase = MID_NB * (q >> 2) + 8 * (q & 3);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              racc = fma(Sl[(size_t)h * M * 4 + lane * 4 + kk], zvec[kbase + 4 * h + kk], racc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
      }
      rpart[warp * 32 + lane] = racc;
      __syncthreads();   // every warp is done with the ring: buf becomes the panel

      // 3. panel -> shared memory in operand order: (row, col) at (col >> 2) CS + 4 row + (col & 3)
      double* P = buf;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int f = warp + NW * s;
        if (f < F) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              const int row = 16 * f + gq + 8 * v1;
              double2 o;
              o.x = -acc[s][ni][2 * v1 + 0];
              o.y = -acc[s][ni][2 * v1 + 1];
              *reinterpret_cast<double2*>(P + (size_t)(2 * ni + (tig >> 1)) * CS + 4 * row + 2 * (tig & 1)) = o;
            }
        }
      }
      if (tid < 32) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) s += rpart[ww * 32 + tid];   // fixed order
        rj[tid] = r[row0 + tid] - s;
      }
      __syncthreads();

      if (warp == FW) {
        // diagonal block: Cholesky of rows 0..31 in registers (lane = row); column k is broadcast
        // through the panel itself
        double a[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a[c] = P[(size_t)(c >> 2) * CS + 4 * lane + (c & 3)];
        double myrs = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const double dk = __shfl_sync(0xffffffffu, a[k], k);
          if (fb == 0x7fffffff && !(dk > 0.0)) {
            fb = row0 + k + 1;
            fb_nan = isnan(dk) ? 1 : 0;
          }
          const double rs = rsqrt(dk);
          const double lik = a[k] * rs;
          if (lane == k) {
            myrs = rs;
            dvec[row0 + k] = dk;
          }
          double* colk = P + (size_t)(k >> 2) * CS + (k & 3);
          if (lane >= k) colk[4 * lane] = lik;
          __syncwarp();
#pragma unroll
          for (int c = k + 1; c < 32; ++c) a[c] = fma(-lik, colk[4 * c], a[c]);
        }
        // inverse of the factor, column per lane: x[i] = inv(L)[i][lane]
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const double rdq = __shfl_sync(0xffffffffu, myrs, q);
          x[q] = (q >= lane) ? ((q == lane ? 1.0 : 0.0) + x[q]) * rdq : 0.0;
          const double* colq = P + (size_t)(q >> 2) * CS + (q & 3);
#pragma unroll
          for (int i = q + 1; i < 32; ++i) x[i] = fma(-colq[4 * i], x[q], x[i]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) invL[i * MID_INV_LD + lane] = x[i];
        __syncwarp();
        // z_j = inv(L_jj) r_j
        double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          z0 = fma(invL[lane * MID_INV_LD + c + 0], rj[c + 0], z0);
          z1 = fma(invL[lane * MID_INV_LD + c + 1], rj[c + 1], z1);
          z2 = fma(invL[lane * MID_INV_LD + c + 2], rj[c + 2], z2);
          z3 = fma(invL[lane * MID_INV_LD + c + 3], rj[c + 3], z3);
        }
        zvec[row0 + lane] = (z0 + z1) + (z2 + z3);
      } else if (j + 1 < nblk) {
        gen_slot(j + 1, 0);   // hides the single-warp section above
      }
      __syncthreads();

      // 4. rows below the diagonal block: X = P inv(L_jj)^T on the tensor pipe, straight to the workspace
      if (j + 1 < nblk) {
        double* dstb = ws + mid_colblk_base(R, j);
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          const int f = warp + NW * s;
          if (f < 2 || f >= F) continue;
          double X[4][4];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int v = 0; v < 4; ++v) X[ni][v] = 0.0;
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            const double* Pc = P + (size_t)kc * CS + (size_t)(2 * f) * 32 + lane;
            const double a0 = Pc[0], a1 = Pc[32];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
              if (2 * ni + 1 >= kc)   // inv(L) is lower triangular: columns 8 ni.. need k <= 8 ni + 7
                dmma_1684(X[ni], a0, a1, invL[(8 * ni + gq) * MID_INV_LD + 4 * kc + tig]);
          }
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              const int row = 16 * f + gq + 8 * v1;
              double2 o;
              o.x = X[ni][2 * v1 + 0];
              o.y = X[ni][2 * v1 + 1];
              *reinterpret_cast<double2*>(dstb + (size_t)(2 * ni + (tig >> 1)) * M * 4 + (size_t)row * 4 + 2 * (tig & 1)) = o;
            }
        }
      }
      asm volatile("fence.proxy.async;\n" ::: "memory");   // generic writes (workspace, buf) before the next bulk copies
      __syncthreads();
    }

    // reductions, closed form, priors
    double ld_part = 0.0, q_part = 0.0;
    for (int k = tid; k < N; k += NT) {
      ld_part += log(dvec[k]);
      q_part += zvec[k] * zvec[k];
    }
    const double logdet = block_sum(ld_part, red);
    const double quad = block_sum(q_part, red);
    const int anybad = block_max_int(bad, ired);
    if (so.z != nullptr)
      for (int k = tid; k < N; k += NT) so.z[(size_t)w * so.ldz + k] = zvec[k];
    if (tid == FW * 32) {
      if (so.logdet != nullptr) so.logdet[w] = logdet;
      double v = -((double)N / 2.0) * log(2.0 * MRV_PI) - 0.5 * logdet - 0.5 * quad;
      v = add_priors(D, hpv, phys, v);
      int st = 0;
      if (anybad) st = -1;
      else if (fb != 0x7fffffff) st = fb_nan ? -1 : fb;
      logl_out[w] = v;
      status_out[w] = st;
    }
  }
}
