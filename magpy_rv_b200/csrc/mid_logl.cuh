// Mid-N log-likelihood (144 < N <= 512, tutorial-2 / 51-Peg sizes): ONE CTA PER WALKER, one launch.
//
// The dense covariance K is never written anywhere.  A walker's CTA (NWM tensor warps + one factor
// warp) runs a left-looking blocked Cholesky over block columns of 32 (panel j = rows [32 j, R) x 32
// columns, R = N padded to 32):
//   1. the tensor warps GENERATE their share of the panel of K from t / yerr straight into their
//      DMMA accumulator registers (as -K, like the large-N GEMM: the epilogue is a negate).  The two
//      column halves of panel j + 1 are generated while the factor warp works on the two 16 x 16
//      diagonal blocks of panel j, so the latency-bound single-warp sections hide behind the
//      FP64-ALU-bound one;
//   2. acc += L(rows, 0:32 j) * L(block-j rows, 0:32 j)^T on the FP64 tensor pipe: the already
//      factored block columns stream back from a per-CTA workspace that stays L2-resident
//      (<= 1.1 MB per CTA slot), k16 slices of ALL panel rows as four 1-D bulk copies (TMA engine)
//      into an mbarrier ring -- no CTA-wide barrier in the loop.  The factor warp is the producer: it
//      runs as far ahead as the ring allows, so consumers wait for data only, never for each other
//      (with a consumer-side producer the loop spent half its time in empty-barrier waits).  The B operand (the 32 rows of
//      block j) is simply the first 32 rows of the same slice, so nothing is loaded twice; one warp
//      per slice also folds the residual update r_j -= L(block-j rows, k) z_k into its pass over the
//      slice (fixed warp / order => bitwise reproducible);
//   3. the panel goes to shared memory in operand order and is finished in two 16-column halves:
//      the factor warp factors the 16 x 16 diagonal block (rolled loops on shared memory: straight-
//      line code on a single warp was instruction-fetch bound), inverts it and forms z; the tensor
//      warps then apply X = P inv(L)^T to the rows below with DMMA, update the right half with the
//      rank-16 product (DMMA again) and repeat for the second half;
//   4. X goes from the accumulator registers straight to the workspace in operand order
//      [k4 chunk][row][4 columns], i.e. exactly the DMMA fragment layout step 2 reads back.
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
#define MID_STAGES 5      // most ring stages of any variant (sizes the mbarrier arrays)
// R > 256 (15 tensor warps): 3 stages of k16 slices (four 1-D bulk copies each); R <= 256 (7 tensor warps, two
// CTAs per SM): 5 stages of k8 slices -- measured: k16 is +2..9 % at N = 300..512, k8 +5 % at N = 256
#define MID_N_MAX 512
#define MID_INV_LD 36     // == 4 (mod 16): conflict-free B-fragment loads of inv(L)

__host__ __device__ inline int mid_npad(int N) { return (N + MID_NB - 1) & ~(MID_NB - 1); }
// first element of block column p in a workspace slot: sum_{p' < p} (R - 32 p') * 32
__host__ __device__ inline size_t mid_colblk_base(int R, int p) {
  return (size_t)32 * p * R - (size_t)512 * p * (p - 1);
}
__host__ inline size_t mid_ws_doubles(int N) {
  const int R = mid_npad(N);
  return mid_colblk_base(R, R / MID_NB);
}
__host__ inline int mid_warps(int N) { return mid_npad(N) <= 256 ? 7 : 15; }   // tensor warps (+ 1 factor warp)
__host__ __device__ inline int mid_stages(int R) { return R <= 256 ? 5 : 3; }
__host__ __device__ inline int mid_ks(int R) { return R <= 256 ? 8 : 16; }
// ring (k16 slices of at most R - 32 rows) and panel (8 chunks of 4 R + 8 doubles) share one region
__host__ __device__ inline size_t mid_buf_doubles(int R) {
  const size_t ring = (size_t)mid_stages(R) * mid_ks(R) * (R > 32 ? R - 32 : 32);
  const size_t panel = (size_t)8 * (4 * R + 8);
  return ring > panel ? ring : panel;
}
__host__ inline size_t mid_smem_bytes(int N) {
  const int R = mid_npad(N), nw = mid_warps(N);
  return (mid_buf_doubles(R) + 6 * (size_t)R + (size_t)nw * 32 + 32 * MID_INV_LD + 32 + 32 + 16 + 32 * 17 + MRV_MAX_MODPAR +
          MRV_MAX_HP + 32 + 16 + 2 * MID_STAGES) * sizeof(double);
}

// Optional phase timers (clock64 deltas of CTA 0, warps 0 and FW), read back via mrv_get_info("mid_t<i>").
__device__ long long g_mid_timing[32];
#ifdef MID_TIMING
#define MID_TICK(i)                              \
  do {                                           \
    const long long now_ = clock64();            \
    tacc_[i] += now_ - tick_;                    \
    tick_ = now_;                                \
  } while (0)
#else
#define MID_TICK(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// The factor warp's role, as its own non-inlined function: its register allocation is then free of the
// tensor warps' 64 accumulator registers (inlined into the kernel the compiler rematerialised
// addresses with S2R / LDC and spilled inside the pivot loop).  It meets the tensor warps at the
// seven CTA barriers of every step.
// ------------------------------------------------------------------------------------------------
struct MidShared {
  double *buf, *r, *dvec, *zvec, *rpart, *invL, *rj, *rdiag, *rhsb, *Dm;
  const double* ws;                         // this CTA's workspace slot (factored block columns)
  unsigned long long *full_bar, *empty_bar;
  int stage_stride;
  unsigned nst;
  int ch;                                   // k4 chunks per slice (2: k8 slices, 4: k16 slices)
};

// Factor the 16 x 16 diagonal sub-block at offset o of the panel: lane = row (rr) x column group (g)
// -- the two half-warps split the columns of every update.
__device__ __forceinline__ void mid_factor16(const MidShared& sh, const double* P, int CS, int row0, int o, int lane,
                                             int& fb, int& fb_nan) {
  double* const Dm = sh.Dm;
  double* const invL = sh.invL;
  double* const rdiag = sh.rdiag;
  double* const dvec = sh.dvec;
  double* const zvec = sh.zvec;
  double* const rhsb = sh.rhsb;
  const double* const rj = sh.rj;

      // History of this routine (each step measured in the kernel): Cholesky + inversion with rsqrt per step, 32 steps of
      // ~480 clocks (8 us per 16 x 16 block; the tensor warps waited for it at the barriers, ncu: 52 % of their stall
      // samples); square-root-free LDL^T + inverse of the unit factor, 16 + 15 lean steps; now ONE Gauss-Jordan sweep of
      // 15 steps (C2 +13 %, N = 160 x 1024 walkers +23 %).  Rolled loops with one __syncwarp per step; Dm[c * 17 + row].
      const int rr = lane & 15, g = lane >> 4;
      {
        const double* src = P + (size_t)((o + 8 * g) >> 2) * CS + 4 * (o + rr);
        double* dst = Dm + (8 * g) * 17 + rr;
#pragma unroll
        for (int u = 0; u < 8; ++u) dst[u * 17] = src[(size_t)(u >> 2) * CS + (u & 3)];
      }
      __syncwarp();
      // ONE square-root-free Gauss-Jordan sweep gives the pivots and the inverse factor together (15 steps with one
      // __syncwarp each; the LDL^T + inversion pair it replaces needed 16 + 15): step j, with w = 1 / d_j,
      //     beta = A[c][j] w;   A[j][c] = -beta;   A[r][c] -= A[r][j] beta   for c > j and (r >= c or r < j)
      // -- rows r >= c carry the factor, rows r < j the (unscaled) inverse, exactly the pivot-block rule of the tiled
      // path (blocked.cuh).  Afterwards inv(L)[n][k] = A[k][n] / sqrt(d_n) for k < n.  This lane: row rr, columns 8 g ...
      {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = Dm[(8 * g + u) * 17 + rr];
        const double* colj = Dm;
#pragma unroll 1
        for (int j = 0; j < 15; ++j, colj += 17) {
          const double dj = colj[j];
          if (fb == 0x7fffffff && !(dj > 0.0)) {
            fb = row0 + o + j + 1;
            fb_nan = isnan(dj) ? 1 : 0;
          }
          const double w = pivot_rcp(dj);
          const double drj = colj[rr];
          double dcj[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) dcj[u] = colj[8 * g + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = 8 * g + u;
            const bool on = c > j && (rr >= c || rr <= j);
            const double beta = dcj[u] * w;
            const double nv = (rr == j) ? -beta : fma(-drj, beta, v[u]);
            if (on) {
              v[u] = nv;
              Dm[c * 17 + rr] = nv;
            }
          }
          __syncwarp();
        }
        const double d15 = Dm[15 * 17 + 15];
        if (fb == 0x7fffffff && !(d15 > 0.0)) {
          fb = row0 + o + 16;
          fb_nan = isnan(d15) ? 1 : 0;
        }
      }
      if (lane < 16) {
        const double d = Dm[lane * 17 + lane];
        dvec[row0 + o + lane] = d;
        rdiag[lane] = rsqrt(d);
      }
      __syncwarp();
      double* X = invL + o * MID_INV_LD + o;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = 8 * g + u;                     // X[row][rr] = inv(L)[row][rr]
        const double sc = rdiag[row];
        X[row * MID_INV_LD + rr] = (rr < row) ? Dm[row * 17 + rr] * sc : (rr == row ? sc : 0.0);
      }
      __syncwarp();

      // z block: rhs = r_block (second half: minus L_BA z_A), z = inv(L) rhs
      if (g == 0) {
        double rhs = rj[o + rr];
        if (o != 0) {
          const double* lba = P + 4 * (16 + rr);
#pragma unroll
          for (int c = 0; c < 16; ++c) rhs = fma(-lba[(size_t)(c >> 2) * CS + (c & 3)], zvec[row0 + c], rhs);
        }
        rhsb[rr] = rhs;
      }
      __syncwarp();
      if (g == 0) {
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          z0 = fma(X[rr * MID_INV_LD + c], rhsb[c], z0);
          z1 = fma(X[rr * MID_INV_LD + c + 1], rhsb[c + 1], z1);
        }
        zvec[row0 + o + rr] = z0 + z1;
      }
      __syncwarp();
    }

template <int NWM>
__device__ __noinline__ void mid_factor_role(MidShared sh, int R, int nblk, unsigned* qg_io, int* fb_out, int* fb_nan_out) {
  const int lane = threadIdx.x & 31;
  int fb = 0x7fffffff, fb_nan = 0;
  unsigned qg = *qg_io;
  for (int j = 0; j < nblk; ++j) {
    const int row0 = MID_NB * j, M = R - row0, CS = 4 * M + 8, spb = 8 / sh.ch, nsl = spb * j;   // spb slices per block column
    // Producer of the operand stream while the tensor warps run the update loop of this step: slice q =
    // 4 ch columns of a block column (k8 or k16), rows [row0, R), one 1-D bulk copy per k4 chunk into ring
    // stage (qg + q) % nst.  A dedicated producer runs as far ahead as the ring allows (all stages in
    // flight), so the consumers only ever wait for data, never for each other.
    if (lane == 0) {
      const unsigned chunk_bytes = (unsigned)M * 4u * (unsigned)sizeof(double);
      for (int q = 0; q < nsl; ++q) {
        const unsigned qglob = qg + (unsigned)q, st = qglob % sh.nst, u = qglob / sh.nst;
        if (u > 0) mbar_wait(&sh.empty_bar[st], (u - 1) & 1u);
        const int p = q / spb, c0 = sh.ch * (q % spb), Rp = R - MID_NB * p;
        const double* src = sh.ws + mid_colblk_base(R, p) + (size_t)c0 * Rp * 4 + (size_t)(row0 - MID_NB * p) * 4;
        double* dst = sh.buf + (size_t)st * sh.stage_stride;
        mbar_expect_tx(&sh.full_bar[st], (unsigned)sh.ch * chunk_bytes);
        for (int h = 0; h < sh.ch; ++h) bulk_g2s(dst + (size_t)h * M * 4, src + (size_t)h * Rp * 4, chunk_bytes, &sh.full_bar[st]);
      }
    }
    qg += (unsigned)nsl;
    __syncwarp();
    __syncthreads();                                   // B1: update loop done, rpart complete
    {
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < NWM; ++ww) s += sh.rpart[ww * 32 + lane];   // fixed order
      sh.rj[lane] = sh.r[row0 + lane] - s;
    }
    __syncthreads();                                   // B2: panel is in shared memory
    mid_factor16(sh, sh.buf, CS, row0, 0, lane, fb, fb_nan);
    __syncthreads();                                   // B3
    __syncthreads();                                   // B4: X_A applied
    __syncthreads();                                   // B5: right half updated
    mid_factor16(sh, sh.buf, CS, row0, 16, lane, fb, fb_nan);
    __syncthreads();                                   // B6
    __syncthreads();                                   // B7: block column stored
  }
  *qg_io = qg;
  *fb_out = fb;
  *fb_nan_out = fb_nan;
}

template <int NWM>
__global__ void __launch_bounds__((NWM + 1) * 32, NWM == 7 ? 2 : 1)
mid_logl_kernel(ProblemDesc D, const double* __restrict__ t, const double* __restrict__ y,
                const double* __restrict__ yerr, const double* __restrict__ flags,
                const double* __restrict__ hp_all, const double* __restrict__ modpar_all,
                const double* __restrict__ model_given, int to_ecc, int W, double* __restrict__ logl_out,
                int* __restrict__ status_out, SolveOut so, double* __restrict__ workspace, size_t ws_stride,
                const double* __restrict__ resid_ready, int resid_ld, const double* __restrict__ phys_ready,
                const int* __restrict__ nonfinite_ready) {
  // resid_ready / phys_ready / nonfinite_ready (nullable): residuals, converted model parameters and the
  // non-finite flag precomputed by mean_residual_kernel for all walkers at once (Keplerian mean models:
  // the reference's whole-vector Newton stop means up to 200 block-wide reductions per walker, which would
  // otherwise run serially inside every walker's CTA before its factorisation can start).
  extern __shared__ double smem[];
  constexpr int NT = (NWM + 1) * 32;
  constexpr int FW = NWM;                   // the dedicated factor warp
  const int N = D.N, R = mid_npad(N), nblk = R / MID_NB;
  double* buf = smem;                       // ring of operand slices  |  panel in operand order
  double* ts = buf + mid_buf_doubles(R);    // 3 R  t (zero padded), phase cosine, phase sine of every point
  double* r = ts + 3 * R;                   // R   residual (zero padded)
  double* dvec = r + R;                     // R   pivots
  double* zvec = dvec + R;                  // R   z = L^-1 r
  double* rpart = zvec + R;                 // NWM x 32  per-warp partials of the residual update
  double* invL = rpart + NWM * 32;          // 32 x MID_INV_LD  inverses of the two 16 x 16 diagonal factors
  double* rj = invL + 32 * MID_INV_LD;      // 32  updated residual block
  double* rdiag = rj + 32;                  // 32  1 / L_kk of the current diagonal block
  double* rhsb = rdiag + 32;                // 16  right-hand side of the current 16-block solve
  double* Dm = rhsb + 16;                   // 32 x 17  the diagonal block while it is factored (+ padding columns)
  double* phys = Dm + 32 * 17;              // MRV_MAX_MODPAR
  double* hpv = phys + MRV_MAX_MODPAR;      // MRV_MAX_HP
  double* red = hpv + MRV_MAX_HP;           // 32
  int* ired = (int*)(red + 32);             // 32 ints
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(red + 32 + 16);
  unsigned long long* empty_bar = full_bar + MID_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tig = lane & 3;
  const bool is_fw = warp == FW;
  constexpr int KS = NWM == 7 ? 8 : 16, CH = KS / 4;        // columns / k4 chunks per slice
  const int stage_stride = (R > 32 ? R - 32 : 32) * KS;   // doubles per ring stage
  const unsigned NST = (unsigned)mid_stages(R);
  double* ws = workspace + (size_t)blockIdx.x * ws_stride;

  if (tid == 0) {
    for (int s = 0; s < MID_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NWM);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async;\n" ::: "memory");
  }
  unsigned qg = 0;   // slices consumed by this CTA so far (uniform): ring position and mbarrier phase

  for (int w = blockIdx.x; w < W; w += gridDim.x) {
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < D.nh; ++i) hpv[i] = hp_all[(size_t)w * D.nh + i];
      for (int i = 0; i < D.nm; ++i) phys[i] = (phys_ready ? phys_ready : modpar_all)[(size_t)w * D.nm + i];
      if (model_given == nullptr && phys_ready == nullptr) convert_to_ecc(D, phys, to_ecc);
    }
    for (int i = tid; i < R; i += NT) {
      ts[i] = i < N ? t[i] : 0.0;
      zvec[i] = 0.0;
      dvec[i] = 1.0;
    }
    __syncthreads();

    // residual (same device functions and operation order as the other paths)
    if (resid_ready != nullptr) {
      for (int i = tid; i < N; i += NT) r[i] = resid_ready[(size_t)w * resid_ld + i];
      __syncthreads();
    } else if (model_given != nullptr) {
      for (int i = tid; i < N; i += NT) r[i] = model_given[(size_t)w * N + i];
      __syncthreads();
    } else {
      cta_mean_model(D, ts, flags, phys, r, /*E scratch=*/buf, red);
    }
    int bad = nonfinite_ready ? nonfinite_ready[w] : 0;
    for (int i = tid; i < R; i += NT) {
      double v = 0.0;
      if (i < N) {
        v = (so.given_is_resid || resid_ready != nullptr) ? r[i] : __dsub_rn(y[i], r[i]);
        bad |= !isfinite(v);
      }
      r[i] = v;
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");   // buf was touched through the generic proxy
    __syncthreads();

    const CovParams cp = make_cov_params(D.kernel_id, D.variant, hpv);
    for (int i = tid; i < R; i += NT) {
      double c = 1.0, s = 0.0;
      if (i < N) cov_point_phase(cp, ts[i] - D.t_ref, c, s);
      ts[R + i] = c;
      ts[2 * R + i] = s;
    }
    __syncthreads();
    int fb = 0x7fffffff, fb_nan = 0;   // first non-positive pivot (tracked by the factor warp)

    // -K for column half h (columns 16 h .. 16 h + 15) of the m16 row fragments f = warp + NWM s of panel jj
    double acc[2][4][4];
    auto gen_half = [&](int jj, int h) {
      const int rw0 = MID_NB * jj, Fj = (R - rw0) >> 4;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int f = warp + NWM * s;
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
          const int ni = 2 * h + nh;
          if (f < Fj) {
            const int gi0 = rw0 + 16 * f + gq, gj0 = rw0 + 8 * ni + 2 * tig;
            if (f >= 2 && rw0 + 16 * f + 15 < N) {   // warp-uniform: below the diagonal block, no padding
              const Cov4 c4 = gen_interior4(cp.kind, cp.a2, cp.c1, cp.c2, cp.c3, ts, R, gi0, ts, R, gj0);
              acc[s][ni][0] = -c4.v0; acc[s][ni][1] = -c4.v1; acc[s][ni][2] = -c4.v2; acc[s][ni][3] = -c4.v3;
            } else {
              double frag[4];
              gen_fragment4(cp, N, gi0, gj0, ts, R, gi0, ts, R, gj0, yerr, frag);
#pragma unroll
              for (int v = 0; v < 4; ++v) acc[s][ni][v] = -frag[v];
            }
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[s][ni][v] = 0.0;
          }
        }
      }
    };
#ifdef MID_TIMING
    long long tacc_[16];
    for (int i_ = 0; i_ < 16; ++i_) tacc_[i_] = 0;
    long long tick_ = clock64();
#endif

    if (is_fw) {
      MidShared sh{buf, r, dvec, zvec, rpart, invL, rj, rdiag, rhsb, Dm, ws, full_bar, empty_bar, stage_stride, NST, CH};
      mid_factor_role<NWM>(sh, R, nblk, &qg, &fb, &fb_nan);
    } else
    for (int j = 0; j < nblk; ++j) {
      const int row0 = MID_NB * j, M = R - row0, F = M >> 4, nsl = (MID_NB / KS) * j;
      const int CS = 4 * M + 8;   // panel chunk stride (doubles): +8 keeps the fragment stores conflict-free
      const bool more = j + 1 < nblk;
      double* P = buf;
      double* dstb = ws + mid_colblk_base(R, j);

      // X = P[rows, 16 h .. 16 h + 15] inv(L_h)^T for the fragments f >= f_min; back into the panel
      // (first half: the right half still needs it) and, for rows below the diagonal block, to the workspace
      auto trsm_half = [&](int h, int f_min, bool to_panel) {
#pragma unroll 1
        for (int f = warp; f < F; f += NWM) {
          if (f < f_min) continue;
          double X[2][4];
#pragma unroll
          for (int nh = 0; nh < 2; ++nh)
#pragma unroll
            for (int v = 0; v < 4; ++v) X[nh][v] = 0.0;
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            const int kc = 4 * h + kq;
            const double* Pc = P + (size_t)kc * CS + (size_t)(2 * f) * 32 + lane;
            const double a0 = Pc[0], a1 = Pc[32];
#pragma unroll
            for (int nh = 0; nh < 2; ++nh)
              if (2 * nh + 1 >= kq)   // inv(L) is lower triangular
                dmma_1684(X[nh], a0, a1, invL[(16 * h + 8 * nh + gq) * MID_INV_LD + 4 * kc + tig]);
          }
#pragma unroll
          for (int nh = 0; nh < 2; ++nh)
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              const int row = 16 * f + gq + 8 * v1, chunk = 4 * h + 2 * nh + (tig >> 1);
              double2 o2;
              o2.x = X[nh][2 * v1 + 0];
              o2.y = X[nh][2 * v1 + 1];
              if (to_panel) *reinterpret_cast<double2*>(P + (size_t)chunk * CS + 4 * row + 2 * (tig & 1)) = o2;
              if (f >= 2 && more)
                *reinterpret_cast<double2*>(dstb + (size_t)chunk * M * 4 + (size_t)row * 4 + 2 * (tig & 1)) = o2;
            }
        }
      };

      {
        // (the operand slices of this step are streamed in by the factor warp, see mid_factor_role)
        MID_TICK(0);

        // 2. stream the factored block columns: acc += A B^T, residual block update on the side
        double racc = 0.0;
        for (int q = 0; q < nsl; ++q) {
          const unsigned qq = qg + q;
          MID_TICK(13);
          const unsigned st = qq % NST;
          mbar_wait(&full_bar[st], (qq / NST) & 1u);
          MID_TICK(14);
          const double* Sl = buf + (size_t)st * stage_stride;
#pragma unroll
          for (int h = 0; h < CH; ++h) {
            const double* base = Sl + (size_t)h * M * 4;
            double b[4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = base[ni * 32 + lane];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const int f = warp + NWM * s;
              if (f < F) {   // warp-uniform
                const double a0 = base[(2 * f) * 32 + lane], a1 = base[(2 * f + 1) * 32 + lane];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_1684(acc[s][ni], a0, a1, b[ni]);
              }
            }
          }
          if ((q % NWM) == warp) {   // r_j[lane] -= sum_k L[row0 + lane][k] z[k] over this slice's columns
            const int kbase = KS * q;   // slices walk the columns 0 .. 32 j in order
#pragma unroll
            for (int h = 0; h < CH; ++h)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                racc = fma(Sl[(size_t)h * M * 4 + lane * 4 + kk], zvec[kbase + 4 * h + kk], racc);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[st]);
          MID_TICK(15);
        }
        rpart[warp * 32 + lane] = racc;
      }
      qg += (unsigned)nsl;
      MID_TICK(1);
      __syncthreads();   // every warp is done with the ring: buf becomes the panel
      MID_TICK(2);

      // 3. panel -> shared memory in operand order: (row, col) at (col >> 2) CS + 4 row + (col & 3).
      //    Panel 0 has no update and up to 2 NWM + 2 row fragments: it is generated straight into the
      //    panel; every later panel has at most 2 NWM fragments and sits in the accumulators.
      if (j == 0) {
        for (int f = warp; f < F; f += NWM) {
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int gi0 = 16 * f + gq, gj0 = 8 * ni + 2 * tig;
            double frag[4];
            if (f >= 2 && 16 * f + 15 < N) {
              const Cov4 c4 = gen_interior4(cp.kind, cp.a2, cp.c1, cp.c2, cp.c3, ts, R, gi0, ts, R, gj0);
              frag[0] = c4.v0; frag[1] = c4.v1; frag[2] = c4.v2; frag[3] = c4.v3;
            } else {
              gen_fragment4(cp, N, gi0, gj0, ts, R, gi0, ts, R, gj0, yerr, frag);
            }
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              double2 o2;
              o2.x = frag[2 * v1 + 0];
              o2.y = frag[2 * v1 + 1];
              *reinterpret_cast<double2*>(P + (size_t)(2 * ni + (tig >> 1)) * CS + 4 * (gi0 + 8 * v1) + 2 * (tig & 1)) = o2;
            }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int f = warp + NWM * s;
          if (f < F) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
              for (int v1 = 0; v1 < 2; ++v1) {
                const int row = 16 * f + gq + 8 * v1;
                double2 o2;
                o2.x = -acc[s][ni][2 * v1 + 0];
                o2.y = -acc[s][ni][2 * v1 + 1];
                *reinterpret_cast<double2*>(P + (size_t)(2 * ni + (tig >> 1)) * CS + 4 * row + 2 * (tig & 1)) = o2;
              }
          }
        }
      }
      __syncthreads();
      MID_TICK(3);

      // first 16 columns: factor | generate, then X_A for all rows below (kept in the panel)
      if (more) gen_half(j + 1, 0);
      MID_TICK(4);
      __syncthreads();
      MID_TICK(5);
      trsm_half(0, 1, true);
      MID_TICK(8);
      __syncthreads();
      MID_TICK(9);
      // right half: P[rows >= 16, 16:32] -= X_A[rows] X_A[16:32]^T (rank 16, DMMA out of the panel)
      {
#pragma unroll 1
        for (int f = warp; f < F; f += NWM) {
          if (f < 1) continue;
          double C[2][4];
#pragma unroll
          for (int nh = 0; nh < 2; ++nh)
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              const int row = 16 * f + gq + 8 * v1;
              const double2 c2 = *reinterpret_cast<const double2*>(P + (size_t)(4 + 2 * nh + (tig >> 1)) * CS + 4 * row + 2 * (tig & 1));
              C[nh][2 * v1 + 0] = -c2.x;
              C[nh][2 * v1 + 1] = -c2.y;
            }
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            const double* Pc = P + (size_t)kc * CS + lane;
            const double a0 = Pc[(2 * f) * 32], a1 = Pc[(2 * f + 1) * 32];
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) dmma_1684(C[nh], a0, a1, Pc[(2 + nh) * 32]);
          }
#pragma unroll
          for (int nh = 0; nh < 2; ++nh)
#pragma unroll
            for (int v1 = 0; v1 < 2; ++v1) {
              const int row = 16 * f + gq + 8 * v1;
              double2 o2;
              o2.x = -C[nh][2 * v1 + 0];
              o2.y = -C[nh][2 * v1 + 1];
              *reinterpret_cast<double2*>(P + (size_t)(4 + 2 * nh + (tig >> 1)) * CS + 4 * row + 2 * (tig & 1)) = o2;
            }
        }
      }
      MID_TICK(10);
      __syncthreads();
      MID_TICK(11);
      // second 16 columns
      if (more) gen_half(j + 1, 1);
      MID_TICK(6);
      __syncthreads();
      MID_TICK(12);
      trsm_half(1, 2, false);
      asm volatile("fence.proxy.async;\n" ::: "memory");   // generic writes (workspace, buf) before the next bulk copies
      __syncthreads();
      MID_TICK(7);
    }

#ifdef MID_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1))
      for (int i_ = 0; i_ < 16; ++i_) g_mid_timing[(warp == 0 ? 0 : 16) + i_] += tacc_[i_];
#endif
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
