// Blocked right-looking Cholesky for large N, batched over a wave of walkers.
//
// Storage (per walker): lower-triangular TILES of 128 x 128 doubles, tile (i, j), i >= j, at index
// i(i+1)/2 + j; N is padded to nt*128 with an identity block (unit diagonal, zero residual) so the
// padding contributes nothing to log det or the quadratic form.  Inside a tile the elements are in
// "operand order": element (r, c) lives at ((c>>2)*16 + (r>>3))*32 + (r&7)*4 + (c&3).  With that
// one layout
//   * a k16 slice (16 consecutive c) of a tile is 16 KiB contiguous -> one bulk copy into smem,
//   * a warp's m16n8k16 DMMA A/B fragment loads are 32 consecutive doubles (conflict-free LDS.64),
//   * the accumulator fragment (row g / g+8, cols 2*tig, 2*tig+1) maps to aligned 16-byte
//     global accesses, so an updated C tile is directly usable as an operand later.
//
// Per tile-column k (host loop in abi.cu):
//   potrf_tile_kernel   one CTA per walker: Gauss-Jordan sweep of tile (k,k) in smem gives the
//                       pivots, inv(L_kk) (written back over the tile) and z_k = L_kk^-1 r_k;
//                       accumulates log det and z^T z; records the LAPACK-style status.
//   gemm_tile_kernel<TRSM>    X = C(i,k) * inv(L_kk)^T on the FP64 tensor pipe (DMMA), written
//                       back as L(i,k); fused GEMV r_i -= L(i,k) z_k.
//   gemm_tile_kernel<UPDATE>  C(i,j) -= sum_p L(i,p) L(j,p)^T over a panel of tile columns; on
//                       first touch C(i,j) is GENERATED from t / yerr instead of loaded (the
//                       covariance matrix is never materialised by a separate pass).
#pragma once
#include "mean_model.cuh"

#define TILE 128
#define TILE_ELEMS (TILE * TILE)
#define KSLICE 16
#define SLICE_ELEMS (TILE * KSLICE)       // 2048 doubles = 16 KiB
#define SLICES_PER_TILE (TILE / KSLICE)   // 8
#define GEMM_THREADS 256

__host__ __device__ __forceinline__ size_t tile_index(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }
__host__ __device__ __forceinline__ int tile_off(int r, int c) {
  return (((c >> 2) * 16 + (r >> 3)) << 5) + ((r & 7) << 2) + (c & 3);
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core MMA (DMMA): D(16x8) += A(16x16) * B(16x8), fragments as in the PTX ISA
//   a[v]: row = g + 8*(v&1), col = tig + 4*(v>>1);  b[v]: k = tig + 4*v, n = g;
//   c[v]: row = g + 8*(v>>1), col = 2*tig + (v&1);   g = lane>>2, tig = lane&3.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// D(16x8) += A(16x4) * B(4x8): a0 = A[g][tig], a1 = A[g+8][tig], b0 = B[tig][g]  (2 x DMMA.8x8x4)
__device__ __forceinline__ void dmma_1684(double (&d)[4], double a0, double a1, double b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a0), "d"(a1), "d"(b0));
}

// ------------------------------------------------------------------------------------------------
// Covariance fragments of the padded matrix, generated from per-point data in shared memory:
// point i of an array `p` with stride `st` is (t, C, S) = (p[i], p[i + st], p[i + 2 st]) -- the time stamp and the
// cosine / sine of its phase (common.cuh: cov_pair_n, cov_point_phase).
// (deliberately not inlined: the callers have 8..16 unrolled sites and the covariance switch with its fp64 exp
// expansions would otherwise be replicated at each of them)
// ------------------------------------------------------------------------------------------------
struct Cov4 { double v0, v1, v2, v3; };
struct Cov8 { double v0, v1, v2, v3, v4, v5, v6, v7; };

// INTERIOR fragments (no diagonal element, nothing in the padding): everything by value and the result in
// registers.  The general versions below take the output array by reference, which routes every value through
// local memory and adds per-element diagonal / padding tests.  Almost every fragment is interior.
// Accumulator order: rows R, R + 8 x columns C, C + 1 (then C + 8, C + 9 for the 8-entry form).
__device__ __noinline__ Cov4 gen_interior4(int kind, double a2, double c1, double c2, double c3, const double* __restrict__ pr,
                                           int sr, int R, const double* __restrict__ pc, int sc, int C) {
  CovParams p;
  p.kind = kind; p.a2 = a2; p.c1 = c1; p.c2 = c2; p.c3 = c3; p.jit2 = 0.0; p.cph = 0.0;
  double ti[4], ci[4], si[4], tj[4], cj[4], sj[4], o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = R + 8 * (e >> 1), c = C + (e & 1);
    ti[e] = pr[r]; ci[e] = pr[r + sr]; si[e] = pr[r + 2 * sr];
    tj[e] = pc[c]; cj[e] = pc[c + sc]; sj[e] = pc[c + 2 * sc];
  }
  cov_pair_n<4, true>(p, ti, ci, si, tj, cj, sj, o);
  Cov4 r;
  r.v0 = o[0]; r.v1 = o[1]; r.v2 = o[2]; r.v3 = o[3];
  return r;
}
__device__ __noinline__ Cov8 gen_interior8(int kind, double a2, double c1, double c2, double c3, const double* __restrict__ pr,
                                           int sr, int R, const double* __restrict__ pc, int sc, int C) {
  CovParams p;
  p.kind = kind; p.a2 = a2; p.c1 = c1; p.c2 = c2; p.c3 = c3; p.jit2 = 0.0; p.cph = 0.0;
  double ti[8], ci[8], si[8], tj[8], cj[8], sj[8], o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = R + 8 * ((e >> 1) & 1), c = C + 8 * (e >> 2) + (e & 1);
    ti[e] = pr[r]; ci[e] = pr[r + sr]; si[e] = pr[r + 2 * sr];
    tj[e] = pc[c]; cj[e] = pc[c + sc]; sj[e] = pc[c + 2 * sc];
  }
  cov_pair_n<8, true>(p, ti, ci, si, tj, cj, sj, o);
  Cov8 r;
  r.v0 = o[0]; r.v1 = o[1]; r.v2 = o[2]; r.v3 = o[3]; r.v4 = o[4]; r.v5 = o[5]; r.v6 = o[6]; r.v7 = o[7];
  return r;
}

// The same 8 interior entries INLINE for a kernel kind known at compile time, negated straight into two accumulator
// fragments: no call, no switch, and the compiler interleaves the exp chains of neighbouring sites.  Used for the
// QuasiPer kinds (every BASELINE configuration but the Matern one); values are bitwise those of gen_interior8.
template <int KIND>
__device__ __forceinline__ void gen_interior8_inline(const CovParams& cp, const double* __restrict__ pr, int R,
                                                     const double* __restrict__ pc, int C, double (&f0)[4], double (&f1)[4]) {
  CovParams p = cp;
  p.kind = KIND;
  double ti[8], ci[8], si[8], tj[8], cj[8], sj[8], o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = R + 8 * ((e >> 1) & 1), c = C + 8 * (e >> 2) + (e & 1);
    ti[e] = pr[r]; ci[e] = pr[r + TILE]; si[e] = pr[r + 2 * TILE];
    tj[e] = pc[c]; cj[e] = pc[c + TILE]; sj[e] = pc[c + 2 * TILE];
  }
  cov_pair_n<8, true>(p, ti, ci, si, tj, cj, sj, o);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    f0[v] = -o[v];
    f1[v] = -o[4 + v];
  }
}

// General fragments: global element (gi0 + 8 (e >> 1), gj0 + (e & 1)) etc.; diagonal elements get cov_diag, elements
// in the padding the identity.
__device__ __noinline__ void gen_fragment4(const CovParams& cp, int N, int gi0, int gj0, const double* __restrict__ pr, int sr,
                                           int R, const double* __restrict__ pc, int sc, int C,
                                           const double* __restrict__ yerr, double (&out)[4]) {
  double ti[4], ci[4], si[4], tj[4], cj[4], sj[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = R + 8 * (e >> 1), c = C + (e & 1);
    ti[e] = pr[r]; ci[e] = pr[r + sr]; si[e] = pr[r + 2 * sr];
    tj[e] = pc[c]; cj[e] = pc[c + sc]; sj[e] = pc[c + 2 * sc];
  }
  cov_pair_n<4, true>(cp, ti, ci, si, tj, cj, sj, out);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int gi = gi0 + 8 * (e >> 1), gj = gj0 + (e & 1);
    if (gi >= N || gj >= N) out[e] = (gi == gj) ? 1.0 : 0.0;
    else if (gi == gj) out[e] = cov_diag(cp, yerr[gi]);
  }
}
__device__ __noinline__ void gen_fragment8(const CovParams& cp, int N, int gi0, int gj0, const double* __restrict__ pr, int sr,
                                           int R, const double* __restrict__ pc, int sc, int C,
                                           const double* __restrict__ yerr, double (&out)[8]) {
  double ti[8], ci[8], si[8], tj[8], cj[8], sj[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int r = R + 8 * ((e >> 1) & 1), c = C + 8 * (e >> 2) + (e & 1);
    ti[e] = pr[r]; ci[e] = pr[r + sr]; si[e] = pr[r + 2 * sr];
    tj[e] = pc[c]; cj[e] = pc[c + sc]; sj[e] = pc[c + 2 * sc];
  }
  cov_pair_n<8, true>(cp, ti, ci, si, tj, cj, sj, out);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int gi = gi0 + 8 * ((e >> 1) & 1), gj = gj0 + 8 * (e >> 2) + (e & 1);
    if (gi >= N || gj >= N) out[e] = (gi == gj) ? 1.0 : 0.0;
    else if (gi == gj) out[e] = cov_diag(cp, yerr[gi]);
  }
}

// Per-walker phase table for the tiled paths: ph[w][0][i] = cospi(2 u), ph[w][1][i] = sinpi(2 u), u = cph_w (t_i - t_ref);
// N sincospi per walker.  Padding points get (1, 0); they are never used (padding entries are the identity).
// Also writes the walker's covariance constants (a2, c1, c2, c3, jit2, 0, 0, 0) for the dataflow kernel, which
// fetches them with the points instead of recomputing the reciprocals in every tile task.
// It is the first kernel of every batch of the tiled paths, so it also clears the per-walker accumulators
// (log det, quadratic form, factorisation status), the tile-row progress counters of the dataflow kernel and its
// task counters / abort flag -- five memset launches less per batch.
__global__ void __launch_bounds__(256)
phase_table_kernel(int N, int n_pad, int W, double t_ref, int kernel_id, int variant, int nh, const double* __restrict__ t,
                   const double* __restrict__ hp_all, double* __restrict__ ph, double* __restrict__ cpar,
                   double* __restrict__ acc_logdet, double* __restrict__ acc_quad, int* __restrict__ fstatus,
                   int* __restrict__ rowdone, int rowdone_per_walker, int* __restrict__ counters, int n_counters) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && counters != nullptr)
    for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0;
  for (int w = blockIdx.y; w < W; w += gridDim.y) {
    const CovParams cp = make_cov_params(kernel_id, variant, hp_all + (size_t)w * nh);
    if (blockIdx.x == 0) {
      if (threadIdx.x < 8) {
        const double v[8] = {cp.a2, cp.c1, cp.c2, cp.c3, cp.jit2, 0.0, 0.0, 0.0};
        cpar[(size_t)w * 8 + threadIdx.x] = v[threadIdx.x];
      }
      if (threadIdx.x == 8) {
        acc_logdet[w] = 0.0;
        acc_quad[w] = 0.0;
        fstatus[w] = 0;
      }
      if (rowdone != nullptr)
        for (int i = threadIdx.x; i < rowdone_per_walker; i += blockDim.x) rowdone[(size_t)w * rowdone_per_walker + i] = 0;
    }
    double* dst = ph + (size_t)w * 2 * n_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
      double c = 1.0, s = 0.0;
      if (i < N) cov_point_phase(cp, t[i] - t_ref, c, s);
      dst[i] = c;
      dst[n_pad + i] = s;
    }
  }
}

// Points of one tile (128 consecutive indices from `first`) into shared memory: p[0..127] = t, p[128..] = C, p[256..] = S.
__device__ __forceinline__ void load_tile_points(double* p, int first, int lane128, int N, int n_pad, const double* __restrict__ t,
                                                 const double* __restrict__ ph_w) {
  const int gi = first + lane128;
  p[lane128] = gi < N ? t[gi] : 0.0;
  p[TILE + lane128] = ph_w[gi];
  p[2 * TILE + lane128] = ph_w[n_pad + gi];
}

// Write the covariance tiles of tile-column `col` for every walker of the wave (needed before
// potrf/trsm touch a column that no UPDATE has generated yet, i.e. column 0).
__global__ void __launch_bounds__(256)
gen_column_kernel(int N, int nt, int col, int kernel_id, int variant, int nh, const double* __restrict__ t,
                  const double* __restrict__ yerr, const double* __restrict__ hp_wave, const double* __restrict__ ph_wave,
                  double* __restrict__ Abase, size_t walker_stride) {
  __shared__ double pi_s[3 * TILE], pj_s[3 * TILE];
  const int i = col + blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int n_pad = nt * TILE;
  const double* ph_w = ph_wave + (size_t)b * 2 * n_pad;
  if (tid < TILE) load_tile_points(pi_s, i * TILE, tid, N, n_pad, t, ph_w);
  else load_tile_points(pj_s, col * TILE, tid - TILE, N, n_pad, t, ph_w);
  __syncthreads();
  const CovParams cp = make_cov_params(kernel_id, variant, hp_wave + (size_t)b * nh);
  double* tile = Abase + (size_t)b * walker_stride + tile_index(i, col) * TILE_ELEMS;
  // four consecutive elements (same row R, columns C..C+3) per trip: one switch, ILP 4, 32-byte stores
  for (int e4 = tid; e4 < TILE_ELEMS / 4; e4 += 256) {
    const int e = e4 * 4;
    const int kg = e >> 9, m8 = (e >> 5) & 15, g = (e >> 2) & 7;
    const int R = m8 * 8 + g, C = kg * 4;
    double ti[4], ci[4], si[4], tj[4], cj[4], sj[4], out[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ti[u] = pi_s[R]; ci[u] = pi_s[TILE + R]; si[u] = pi_s[2 * TILE + R];
      tj[u] = pj_s[C + u]; cj[u] = pj_s[TILE + C + u]; sj[u] = pj_s[2 * TILE + C + u];
    }
    cov_pair_n<4, true>(cp, ti, ci, si, tj, cj, sj, out);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gi = i * TILE + R, gj = col * TILE + C + u;
      if (gi >= N || gj >= N) out[u] = (gi == gj) ? 1.0 : 0.0;
      else if (gi == gj) out[u] = cov_diag(cp, yerr[gi]);
      tile[e + u] = out[u];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile: pivots, inverse factor, partial forward solve.  One CTA (256 threads) per walker.
// smem: (TILE+1) x TILE column-major, ld = 129 (row 128 carries the residual block r_k).
//
// Gauss-Jordan sweep (square-root free): for j = 0..127, with w = 1/S[j][j] and beta_c = S[c][j]*w:
//     rows rho < j   (inverse part)   S[rho][c] -= S[rho][j] * beta_c      c > j
//     row  rho = j                    S[j][c]    = -beta_c                 c > j
//     rows rho >= c  (factor part, incl. the residual row)  S[rho][c] -= S[rho][j] * beta_c
// Afterwards, with d_c = S[c][c]:  L[i][c] = S[i][c]/sqrt(d_c) (i > c),
//     inv(L)[c][q] = S[q][c]/sqrt(d_c) (q < c), inv(L)[c][c] = 1/sqrt(d_c), z_c = S[128][c]/sqrt(d_c).
// ------------------------------------------------------------------------------------------------
#define POTRF_LD (TILE + 1)
#define POTRF_SMEM_BYTES ((size_t)POTRF_LD * TILE * sizeof(double) + (64 + TILE) * sizeof(double))

__global__ void __launch_bounds__(256)
potrf_tile_kernel(int k, int N, double* __restrict__ Abase, size_t walker_stride, double* __restrict__ resid,
                  int ldr, double* __restrict__ zbuf, double* __restrict__ acc_logdet,
                  double* __restrict__ acc_quad, int* __restrict__ status, double* __restrict__ zout,
                  long long ldz) {
  extern __shared__ double smem[];
  double* S = smem;
  double* red = smem + (size_t)POTRF_LD * TILE;
  int* ired = (int*)(red + 32);
  double* isd_s = red + 64;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = 8;
  double* tile = Abase + (size_t)b * walker_stride + tile_index(k, k) * TILE_ELEMS;
  double* rk = resid + (size_t)b * ldr + (size_t)k * TILE;

  // load the tile (operand order -> column-major smem); coalesced global reads
  for (int e = tid; e < TILE_ELEMS; e += 256) {
    const int kg = e >> 9, m8 = (e >> 5) & 15, g = (e >> 2) & 7, k3 = e & 3;
    const int R = m8 * 8 + g, C = kg * 4 + k3;
    S[R + C * POTRF_LD] = tile[e];
  }
  if (tid < TILE) S[TILE + tid * POTRF_LD] = rk[tid];

  for (int j = 0; j < TILE; ++j) {
    __syncthreads();
    const double* cj = S + j * POTRF_LD;
    const double winv = 1.0 / cj[j];
    for (int c = j + 1 + warp; c < TILE; c += nwarps) {
      double* col = S + c * POTRF_LD;
      const double beta = cj[c] * winv;
      // rows [0, j) then [c, 128]: one fused index space so all lanes stay busy
      const int n_inv = j, n_tot = j + (TILE + 1 - c);
      for (int v = lane; v < n_tot; v += 32) {
        const int rho = (v < n_inv) ? v : (c + v - n_inv);
        col[rho] = fma(-cj[rho], beta, col[rho]);
      }
      if (lane == 0) col[j] = -beta;
    }
  }
  __syncthreads();

  // pivots -> scale factors, status, log det, z
  double ld_part = 0.0, q_part = 0.0;
  int first_bad = 0x7fffffff, nan_seen = 0;
  double isd = 0.0;
  if (tid < TILE) {
    const double d = S[tid + tid * POTRF_LD];
    if (!(d > 0.0)) {
      first_bad = k * TILE + tid + 1;
      nan_seen = isnan(d) ? 1 : 0;
    }
    isd = 1.0 / sqrt(d);
    ld_part = log(d);
    const double z = S[TILE + tid * POTRF_LD] * isd;
    q_part = z * z;
    zbuf[(size_t)b * TILE + tid] = z;
    if (zout != nullptr && k * TILE + tid < N) zout[(size_t)b * ldz + (size_t)k * TILE + tid] = z;
  }
  const double logdet = block_sum(ld_part, red);
  const double quad = block_sum(q_part, red);
  const int fb = -block_max_int(-first_bad, ired);
  const int anynan = block_max_int(nan_seen, ired);
  if (tid == 0) {
    acc_logdet[b] += logdet;
    acc_quad[b] += quad;
    if (status[b] == 0 && fb != 0x7fffffff) status[b] = anynan ? -1 : fb;
  }
  if (tid < TILE) isd_s[tid] = isd;
  __syncthreads();
  // write inv(L_kk) over the tile in operand order: element (R = n, C = kk) = inv(L)[n][kk]
  // inv(L)[n][kk] = S[kk][n] / sqrt(d_n) for kk < n; 1/sqrt(d_n) on the diagonal; 0 above it.
  for (int e = tid; e < TILE_ELEMS; e += 256) {
    const int kg = e >> 9, m8 = (e >> 5) & 15, g = (e >> 2) & 7, k3 = e & 3;
    const int R = m8 * 8 + g, C = kg * 4 + k3;
    double v = 0.0;
    if (C <= R) {
      const double s = isd_s[R];
      v = (C == R) ? s : S[C + R * POTRF_LD] * s;
    }
    tile[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Blocked (rank-16) form of the same sweep, trailing updates on the FP64 tensor pipe.
//
// S is column-major with ld = 136 (== 8 mod 16: the transposed accumulator fragments below are
// conflict-free 16-byte accesses; thread-per-row column walks are conflict-free for any ld).
// Per block step kb (pivot columns K = [16 kb, 16 kb + 16)):
//   1. the 16 x 16 pivot block D is swept by the scalar algorithm (256 threads, one element each);
//   2. every other row rho (inverse rows < 16 kb, factor rows >= 16 kb + 16, residual row 128)
//      finishes its 16 panel entries with a 120-FMA triangular recurrence (one thread per row);
//      the panel is written back to S and staged as MMA operands
//         PA[rho][j]  = panel value (pivot rows: unit upper triangle of the inverse block)
//         PB[c][j]    = -S[c][j] / d_j
//   3. S[rho][c] += sum_j PA[rho][j] PB[c][j] for all trailing columns c (pivot rows start from 0:
//      assignment) as DMMA tiles, computed transposed (M = c, N = rho) so accumulators map to
//      contiguous 16-byte smem accesses; tiles lying wholly in the not-yet-activated part of the
//      upper triangle are skipped.
// ------------------------------------------------------------------------------------------------
#define PB_NB 16
#define PB_LD 136
#define PB_PA_LD 148   // == 4 mod 16: conflict-free B-fragment loads
#define PB_PB_LD 132   // == 4 mod 16: conflict-free A-fragment loads
#define PB_D_LD 17
#define POTRF_BLK_SMEM_DOUBLES (PB_LD * TILE + PB_NB * PB_PA_LD + PB_NB * PB_PB_LD + PB_NB * PB_D_LD + PB_NB + 64 + TILE)
#define POTRF_BLK_SMEM_BYTES ((size_t)POTRF_BLK_SMEM_DOUBLES * sizeof(double))

__device__ __forceinline__ void dmma_1684_neg0(double (&d)[4], double a0, double a1, double b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a0), "d"(a1), "d"(b0));
}

// Sweeps the n_cols x n_cols matrix held in S (column-major, ld = PB_LD, lower triangle valid, row
// `n_cols` = residual) in place.  n_cols must be a multiple of 16; ld == 8 (mod 16), ld >= rows rounded up to 8, ld <= 256;
// pa_ld, pb_ld == 4 (mod 16).  With WITH_INV the
// strict upper triangle receives the unscaled inverse factor (see potrf_tile_kernel above);
// without it only factor rows and the residual row are carried (fused small-N kernel).
// Column base pointer of the swept matrix.  PACKED (fused small-N kernel, factor rows only): the 16-column
// block b keeps rows >= 16 b only, with leading dimension ld - 16 b, blocks stored back to back -- half the
// shared memory of the square layout, so two walkers' CTAs fit an SM up to N = 112.  The returned pointer is
// pre-offset so that colp[row] is valid for every stored row (row >= 16 b).  ld == 8 (mod 16) holds for
// every block, which is what the conflict-free 16-byte accumulator accesses need.
template <bool PACKED>
__device__ __forceinline__ double* sweep_colp(double* S, int ld, int c) {
  if (!PACKED) return S + (size_t)c * ld;
  const int b = c >> 4;
  return S + (16 * b * ld - 128 * b * (b - 1)) + (c & 15) * (ld - 16 * b) - 16 * b;
}
__host__ __device__ inline size_t sweep_packed_doubles(int ld, int n_cols) {
  const int nb = n_cols / 16;
  return (size_t)16 * nb * ld - (size_t)128 * nb * (nb - 1);
}

// Phase timers of the dataflow kernel (only with -DDAG_TIMING; thread 0 of a CTA accumulates clock64 deltas in
// registers, tools/dag_timing.py reads the per-phase totals).  Without the flag every call is empty.
// Reciprocal of a pivot: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps in FMA, 5 dependent FP64 instructions
// instead of the ~25 of an IEEE division with its special-case paths (every thread of the pivot block divides after every
// barrier: in the fused small-N kernel that one line was 15 % of all stall samples, a quarter of them FP64-pipe
// contention).  Within 1 ulp of 1 / x for normal x; the sign and NaN pass through; x = 0 gives NaN where the division
// gave inf (either way the pivot itself, not its reciprocal, decides the status word).
__device__ __forceinline__ double pivot_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

struct PhaseTimer {
#ifdef DAG_TIMING
  long long* a;   // shared memory: 16 accumulators + the last time stamp (only thread 0 touches them)
  __device__ __forceinline__ void init(void* smem17) {
    a = reinterpret_cast<long long*>(smem17);
    if (threadIdx.x == 0) {
      for (int i = 0; i < 16; ++i) a[i] = 0;
      a[16] = clock64();
    }
  }
  __device__ __forceinline__ void mark(int slot) {
    if (threadIdx.x == 0) {
      const long long now = clock64();
      a[slot] += now - a[16];
      a[16] = now;
    }
  }
  __device__ __forceinline__ void count(int slot) {
    if (threadIdx.x == 0) a[slot] += 1;
  }
#else
  __device__ __forceinline__ void init(void*) {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void count(int) {}
#endif
};

__device__ __forceinline__ void bar_sync_pivot64() { asm volatile("bar.sync 1, 64;\n" ::: "memory"); }

// 16 x 16 pivot block at (k0, k0) by threads u = 0..63 (warps 0-1), four elements of a row each, synchronised by a named
// barrier of their own; the caller's __syncthreads follows.  (One element per thread on 256 threads, as in round 1, costs
// 1.96 us per block in the dataflow kernel, this form 1.61 us: the barrier spans two warps instead of eight.)
template <bool PACKED>
__device__ __forceinline__ void sweep_pivot_block64(double* S, const int ld, const int k0, double* D, double* wv, const int u) {
  const int r = u >> 2, c0 = (u & 3) * 4;
  double v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = sweep_colp<PACKED>(S, ld, k0 + c0 + e)[k0 + r];
    D[r * PB_D_LD + c0 + e] = v[e];
  }
  bar_sync_pivot64();
  for (int j = 0; j < PB_NB - 1; ++j) {
    {
      // branch-free over the thread's four elements (divergent per-element branches would run the four load -> multiply ->
      // fma -> store chains one after the other): all loads first, selects instead of branches, predicated stores
      const double w = pivot_rcp(D[j * PB_D_LD + j]);
      const double drj = D[r * PB_D_LD + j];
      double dcj[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) dcj[e] = D[(c0 + e) * PB_D_LD + j];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = c0 + e;
        const bool on = c > j && (r >= c || r <= j);
        const double beta = dcj[e] * w;
        const double nv = (r == j) ? -beta : fma(-drj, beta, v[e]);
        v[e] = on ? nv : v[e];
        if (on) D[r * PB_D_LD + c] = nv;
      }
    }
    bar_sync_pivot64();
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (r == c0 + e) wv[r] = pivot_rcp(v[e]);
    sweep_colp<PACKED>(S, ld, k0 + c0 + e)[k0 + r] = v[e];     // final block (lower: factor, upper: inverse part)
  }
}

template <bool WITH_INV, int NWARPS, bool PACKED = false>
__device__ inline void cta_block_sweep(double* S, const int ld, int n_cols, double* PA, const int pa_ld, double* PB,
                                       const int pb_ld, double* D, double* wv, PhaseTimer* ptm = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tig = lane & 3;
  const int n_rows = n_cols + 1;                 // + residual row
  const int n_rows8 = (n_rows + 7) & ~7;         // rows covered by n8 MMA tiles (padding rows are zero)
  const int nblk = n_cols / PB_NB;
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * PB_NB;
    // ---- 1. pivot block (warps 0-1; everybody else waits at the barrier) ------------------------
#ifdef SWEEP_PIVOT_256
    {
      const bool in_blk = tid < PB_NB * PB_NB;           // one thread per pivot-block element
      const int r = (tid >> 4) & 15, c = tid & 15;
      double v = 0.0;
      if (in_blk) {
        v = sweep_colp<PACKED>(S, ld, k0 + c)[k0 + r];
        D[r * PB_D_LD + c] = v;
      }
      __syncthreads();
      for (int j = 0; j < PB_NB - 1; ++j) {
        if (in_blk && c > j && (r >= c || r <= j)) {
          const double w = pivot_rcp(D[j * PB_D_LD + j]);
          const double beta = D[c * PB_D_LD + j] * w;
          if (r == j) v = -beta;
          else v = fma(-D[r * PB_D_LD + j], beta, v);
          D[r * PB_D_LD + c] = v;
        }
        __syncthreads();
      }
      if (in_blk) {
        if (r == c) wv[r] = pivot_rcp(v);
        sweep_colp<PACKED>(S, ld, k0 + c)[k0 + r] = v;     // final block (lower: factor, upper: inverse part)
      }
      __syncthreads();
    }
#else
    if (warp < 2) sweep_pivot_block64<PACKED>(S, ld, k0, D, wv, tid);
    __syncthreads();
#endif
    if (ptm) ptm->mark(8);
    // ---- 2. panel rows on the tensor pipe ------------------------------------------------------
    // A row's 16 panel entries x solve x U = a, U unit upper triangular with the block's multipliers; the swept pivot
    // block holds U^-1 =: G in its upper triangle (its own rows are e_r G), so X = A G for all rows at once:
    // X^T (16 x rows) = G^T (16 x 16) A^T (16 x rows) as m16n8k4 DMMAs, one n8 tile of 8 rows per step -- instead of
    // one thread per row running a 120-FMA triangular recurrence (11 of the 46 us of a 128 x 128 diagonal tile).
    {
      const int n_n8p = n_rows8 / 8;
      double gt0[4], gt1[4];   // G^T fragments: rows j = gq and gq + 8, columns jp = 4 kk + tig
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int jp = 4 * kk + tig;
        gt0[kk] = (jp < gq) ? D[jp * PB_D_LD + gq] : (jp == gq ? 1.0 : 0.0);
        gt1[kk] = (jp < gq + 8) ? D[jp * PB_D_LD + gq + 8] : (jp == gq + 8 ? 1.0 : 0.0);
      }
      for (int tn = warp; tn < n_n8p; tn += NWARPS) {
        const int rho0 = tn * 8;
        const bool pivot_tile = rho0 >= k0 && rho0 < k0 + PB_NB;
        const bool above = rho0 < k0;                          // inverse rows (kept only WITH_INV)
        const bool compute = !pivot_tile && (WITH_INV || !above);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (compute) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            dmma_1684_neg0(acc, gt0[kk], gt1[kk], sweep_colp<PACKED>(S, ld, k0 + 4 * kk + tig)[rho0 + gq]);
        }
        const int rhoA = rho0 + 2 * tig;                       // this lane's two rows: rhoA, rhoA + 1
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = gq + 8 * h;
          double2 x;
          x.x = acc[2 * h];
          x.y = acc[2 * h + 1];
          if (pivot_tile) {
            const int r0 = rhoA - k0, r1 = r0 + 1;
            x.x = (j == r0) ? 1.0 : (j > r0 ? D[r0 * PB_D_LD + j] : 0.0);
            x.y = (j == r1) ? 1.0 : (j > r1 ? D[r1 * PB_D_LD + j] : 0.0);
          } else if (compute) {
            if (rhoA >= n_rows) x.x = 0.0;                     // padding rows below the residual row stay zero
            if (rhoA + 1 >= n_rows) x.y = 0.0;
            *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, k0 + j) + rhoA) = x;
          }
          *reinterpret_cast<double2*>(PA + j * pa_ld + rhoA) = x;
          if (rhoA < n_cols) {
            const double w = wv[j];
            double2 y;
            y.x = (rhoA >= k0 + PB_NB) ? -x.x * w : 0.0;
            y.y = (rhoA + 1 >= k0 + PB_NB && rhoA + 1 < n_cols) ? -x.y * w : 0.0;
            *reinterpret_cast<double2*>(PB + j * pb_ld + rhoA) = y;
          }
        }
      }
    }
    __syncthreads();
    if (ptm) ptm->mark(9);
    // ---- 3. trailing update on the tensor pipe -------------------------------------------------
    const int c_begin = k0 + PB_NB;
    const int n_m16 = (n_cols - c_begin) / 16;
    const int n_n8 = n_rows8 / 8;
    // each warp owns row tiles tn = warp, warp + NWARPS, ... (n_n8 <= 24); their accumulator
    // chains are interleaved so consecutive DMMAs are independent
    constexpr int SLOTS = (24 + NWARPS - 1) / NWARPS;     // n_n8 <= 24 row tiles
    int rho0s[SLOTS];
    bool live[SLOTS], pivot[SLOTS];
    double bfr[SLOTS][4];
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int tn = warp + NWARPS * sl;
      rho0s[sl] = tn * 8;
      live[sl] = tn < n_n8 && (WITH_INV || rho0s[sl] + 7 >= c_begin);
      pivot[sl] = rho0s[sl] >= k0 && rho0s[sl] < k0 + PB_NB;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) bfr[sl][kk] = live[sl] ? PA[(kk * 4 + tig) * pa_ld + rho0s[sl] + gq] : 0.0;
    }
    for (int tm = 0; tm < n_m16; ++tm) {
      const int c0 = c_begin + tm * 16;
      double a0[4], a1[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        a0[kk] = PB[(kk * 4 + tig) * pb_ld + c0 + gq];
        a1[kk] = PB[(kk * 4 + tig) * pb_ld + c0 + gq + 8];
      }
      double acc[SLOTS][4];
      bool need[SLOTS];
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        // needed if the rows are inverse/pivot rows (rho < c_begin) or contain factor rows (rho >= c)
        need[sl] = live[sl] && (rho0s[sl] < c_begin || rho0s[sl] + 7 >= c0);
        if (need[sl] && !pivot[sl]) {
          const double2 u = *reinterpret_cast<const double2*>(sweep_colp<PACKED>(S, ld, c0 + gq) + (rho0s[sl] + 2 * tig));
          const double2 v = *reinterpret_cast<const double2*>(sweep_colp<PACKED>(S, ld, c0 + gq + 8) + (rho0s[sl] + 2 * tig));
          acc[sl][0] = u.x; acc[sl][1] = u.y; acc[sl][2] = v.x; acc[sl][3] = v.y;
        } else {
          acc[sl][0] = acc[sl][1] = acc[sl][2] = acc[sl][3] = 0.0;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl)
          if (need[sl]) dmma_1684_neg0(acc[sl], a0[kk], a1[kk], bfr[sl][kk]);   // warp-uniform predicate
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        if (need[sl]) {
          double2 u, v;
          u.x = acc[sl][0]; u.y = acc[sl][1]; v.x = acc[sl][2]; v.y = acc[sl][3];
          *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, c0 + gq) + (rho0s[sl] + 2 * tig)) = u;
          *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, c0 + gq + 8) + (rho0s[sl] + 2 * tig)) = v;
        }
      }
    }
    __syncthreads();
    if (ptm) ptm->mark(10);
  }
}

// ------------------------------------------------------------------------------------------------
// The same sweep with LOOK-AHEAD, for the dataflow kernel (256 threads, inverse kept, square storage).  In
// cta_block_sweep the three phases of a block step run one after the other, and the pivot block -- 15 barrier-separated
// scalar steps, pure latency -- leaves the FP64 pipes idle for a third of the sweep.  Here the trailing update of block
// step kb is split: first the 16 columns of the NEXT pivot block (all warps), then warps 0-1 sweep that pivot block
// (64 threads, four elements of a row each, a named barrier of their own) WHILE warps 2-7 update the remaining columns.
// Every element sees the same operations in the same order as in cta_block_sweep, so the results are bitwise identical.
// ------------------------------------------------------------------------------------------------
// Trailing update of the m16 column tiles tm0 .. tm1 - 1 behind pivot block k0 by NW warps (this warp: index wi).
template <int NW, int MAXT, bool WITH_INV, bool PACKED>
__device__ __forceinline__ void sweep_trailing_tiles(double* S, const int ld, const int n_n8, const int k0, const int tm0,
                                                     const int tm1, const double* PA, const int pa_ld, const double* PB,
                                                     const int pb_ld, const int wi, const int gq, const int tig,
                                                     const int tn_ex = -8) {   // row tiles tn_ex, tn_ex + 1 are somebody else's
  const int c_begin = k0 + PB_NB;
  constexpr int SLOTS = (MAXT + NW - 1) / NW;     // n_n8 <= MAXT row tiles (rows + the residual row)
  int rho0s[SLOTS];
  bool live[SLOTS], pivot[SLOTS];
  double bfr[SLOTS][4];
#pragma unroll
  for (int sl = 0; sl < SLOTS; ++sl) {
    const int tn = wi + NW * sl;
    rho0s[sl] = tn * 8;
    live[sl] = tn < n_n8 && (tn < tn_ex || tn > tn_ex + 1) && (WITH_INV || rho0s[sl] + 7 >= c_begin);
    pivot[sl] = rho0s[sl] >= k0 && rho0s[sl] < k0 + PB_NB;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) bfr[sl][kk] = live[sl] ? PA[(kk * 4 + tig) * pa_ld + rho0s[sl] + gq] : 0.0;
  }
  for (int tm = tm0; tm < tm1; ++tm) {
    const int c0 = c_begin + tm * 16;
    double a0[4], a1[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      a0[kk] = PB[(kk * 4 + tig) * pb_ld + c0 + gq];
      a1[kk] = PB[(kk * 4 + tig) * pb_ld + c0 + gq + 8];
    }
    double acc[SLOTS][4];
    bool need[SLOTS];
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      need[sl] = live[sl] && (rho0s[sl] < c_begin || rho0s[sl] + 7 >= c0);
      if (need[sl] && !pivot[sl]) {
        const double2 u = *reinterpret_cast<const double2*>(sweep_colp<PACKED>(S, ld, c0 + gq) + (rho0s[sl] + 2 * tig));
        const double2 v = *reinterpret_cast<const double2*>(sweep_colp<PACKED>(S, ld, c0 + gq + 8) + (rho0s[sl] + 2 * tig));
        acc[sl][0] = u.x; acc[sl][1] = u.y; acc[sl][2] = v.x; acc[sl][3] = v.y;
      } else {
        acc[sl][0] = acc[sl][1] = acc[sl][2] = acc[sl][3] = 0.0;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl)
        if (need[sl]) dmma_1684_neg0(acc[sl], a0[kk], a1[kk], bfr[sl][kk]);   // warp-uniform predicate
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      if (need[sl]) {
        double2 u, v;
        u.x = acc[sl][0]; u.y = acc[sl][1]; v.x = acc[sl][2]; v.y = acc[sl][3];
        *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, c0 + gq) + (rho0s[sl] + 2 * tig)) = u;
        *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, c0 + gq + 8) + (rho0s[sl] + 2 * tig)) = v;
      }
    }
  }
}

template <bool WITH_INV = true, bool PACKED = false, int MAXT = 17>
__device__ inline void cta_block_sweep_la(double* S, const int ld, int n_cols, double* PA, const int pa_ld, double* PB,
                                          const int pb_ld, double* D, double* wv, PhaseTimer* ptm = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tig = lane & 3;
  const int n_rows = n_cols + 1;                 // + residual row
  const int n_rows8 = (n_rows + 7) & ~7;
  const int n_n8 = n_rows8 / 8;
  const int nblk = n_cols / PB_NB;
  if (warp < 2) sweep_pivot_block64<PACKED>(S, ld, 0, D, wv, tid);
  __syncthreads();
  if (ptm) ptm->mark(8);
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * PB_NB;
    // ---- panel rows on the tensor pipe (all warps; as in cta_block_sweep) ----
    {
      double gt0[4], gt1[4];   // G^T fragments: rows j = gq and gq + 8, columns jp = 4 kk + tig
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int jp = 4 * kk + tig;
        gt0[kk] = (jp < gq) ? D[jp * PB_D_LD + gq] : (jp == gq ? 1.0 : 0.0);
        gt1[kk] = (jp < gq + 8) ? D[jp * PB_D_LD + gq + 8] : (jp == gq + 8 ? 1.0 : 0.0);
      }
      for (int tn = warp; tn < n_n8; tn += 8) {
        const int rho0 = tn * 8;
        const bool pivot_tile = rho0 >= k0 && rho0 < k0 + PB_NB;
        const bool compute = !pivot_tile && (WITH_INV || rho0 >= k0);   // rows above the pivot block: inverse rows (kept only WITH_INV)
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (compute) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) dmma_1684_neg0(acc, gt0[kk], gt1[kk], sweep_colp<PACKED>(S, ld, k0 + 4 * kk + tig)[rho0 + gq]);
        }
        const int rhoA = rho0 + 2 * tig;                       // this lane's two rows: rhoA, rhoA + 1
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = gq + 8 * h;
          double2 x;
          x.x = acc[2 * h];
          x.y = acc[2 * h + 1];
          if (pivot_tile) {
            const int r0 = rhoA - k0, r1 = r0 + 1;
            x.x = (j == r0) ? 1.0 : (j > r0 ? D[r0 * PB_D_LD + j] : 0.0);
            x.y = (j == r1) ? 1.0 : (j > r1 ? D[r1 * PB_D_LD + j] : 0.0);
          } else if (compute) {
            if (rhoA >= n_rows) x.x = 0.0;                     // padding rows below the residual row stay zero
            if (rhoA + 1 >= n_rows) x.y = 0.0;
            *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, k0 + j) + rhoA) = x;
          }
          *reinterpret_cast<double2*>(PA + j * pa_ld + rhoA) = x;
          if (rhoA < n_cols) {
            const double w = wv[j];
            double2 y;
            y.x = (rhoA >= k0 + PB_NB) ? -x.x * w : 0.0;
            y.y = (rhoA + 1 >= k0 + PB_NB && rhoA + 1 < n_cols) ? -x.y * w : 0.0;
            *reinterpret_cast<double2*>(PB + j * pb_ld + rhoA) = y;
          }
        }
      }
    }
    __syncthreads();
    if (ptm) ptm->mark(9);
    const int n_m16 = (n_cols - (k0 + PB_NB)) / 16;
    if (n_m16 == 0) break;
    // ---- trailing update, first the columns of the next pivot block (all warps) ... ----
    sweep_trailing_tiles<8, MAXT, WITH_INV, PACKED>(S, ld, n_n8, k0, 0, 1, PA, pa_ld, PB, pb_ld, warp, gq, tig);
    __syncthreads();
    // ---- ... then warps 0-1 sweep that pivot block while warps 2-7 update the rest ----
    if (warp < 2) sweep_pivot_block64<PACKED>(S, ld, k0 + PB_NB, D, wv, tid);
    else sweep_trailing_tiles<6, MAXT, WITH_INV, PACKED>(S, ld, n_n8, k0, 1, n_m16, PA, pa_ld, PB, pb_ld, warp - 2, gq, tig);
    __syncthreads();
    if (ptm) ptm->mark(10);
  }
}

// ------------------------------------------------------------------------------------------------
// CRITICAL-PATH form of the look-ahead sweep for the fused small-N kernel (factor rows only, 256 threads).  With a matrix
// of ~100 columns the tensor-pipe phases of a block step are short and the chain
//     pivot block kb -> panel rows -> next pivot block's 16 columns -> pivot block kb + 1
// is what a step costs.  Only the 16 rows / columns of the NEXT pivot block are on that chain, so warps 0-1 keep it to
// themselves: they finish the panel entries of those 16 rows (one 8-row tile each), update the 16 x 16 diagonal block and
// sweep it, synchronised by their own 64-thread barrier, while warps 2-7 do everything else of step kb (panel rows of the
// other rows, then -- once warps 0-1 have published their panel rows, named barrier 2 -- the trailing update).  The pivot
// block buffers D / wv are double-buffered (warps 2-7 still read block kb's while block kb + 1 is swept).  Every element
// sees the same operations in the same order as in cta_block_sweep: bitwise identical results.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_arrive_panel() { asm volatile("bar.arrive 2, 256;\n" ::: "memory"); }
__device__ __forceinline__ void bar_sync_panel() { asm volatile("bar.sync 2, 256;\n" ::: "memory"); }

// Panel entries of the 8-row tile rho0 (not a pivot tile) for pivot block k0: X = A G on the tensor pipe; writes the swept
// columns of S and the operand panels PA / PB (as step 2 of cta_block_sweep).
template <bool PACKED>
__device__ __forceinline__ void sweep_panel_tile(double* S, const int ld, const int n_cols, const int n_rows, const int k0,
                                                 const int rho0, const double (&gt0)[4], const double (&gt1)[4], double* PA,
                                                 const int pa_ld, double* PB, const int pb_ld, const double* D, const double* wv,
                                                 const int gq, const int tig) {
  const bool pivot_tile = rho0 >= k0 && rho0 < k0 + PB_NB;   // (only with the inverse kept: D = the swept pivot block)
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (!pivot_tile) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) dmma_1684_neg0(acc, gt0[kk], gt1[kk], sweep_colp<PACKED>(S, ld, k0 + 4 * kk + tig)[rho0 + gq]);
  }
  const int rhoA = rho0 + 2 * tig;                       // this lane's two rows: rhoA, rhoA + 1
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = gq + 8 * h;
    double2 x;
    x.x = acc[2 * h];
    x.y = acc[2 * h + 1];
    if (pivot_tile) {
      const int r0 = rhoA - k0, r1 = r0 + 1;
      x.x = (j == r0) ? 1.0 : (j > r0 ? D[r0 * PB_D_LD + j] : 0.0);
      x.y = (j == r1) ? 1.0 : (j > r1 ? D[r1 * PB_D_LD + j] : 0.0);
    } else {
      if (rhoA >= n_rows) x.x = 0.0;                     // padding rows below the residual row stay zero
      if (rhoA + 1 >= n_rows) x.y = 0.0;
      *reinterpret_cast<double2*>(sweep_colp<PACKED>(S, ld, k0 + j) + rhoA) = x;
    }
    *reinterpret_cast<double2*>(PA + j * pa_ld + rhoA) = x;
    if (rhoA < n_cols) {
      const double w = wv[j];
      double2 y;
      y.x = (rhoA >= k0 + PB_NB) ? -x.x * w : 0.0;
      y.y = (rhoA + 1 >= k0 + PB_NB && rhoA + 1 < n_cols) ? -x.y * w : 0.0;
      *reinterpret_cast<double2*>(PB + j * pb_ld + rhoA) = y;
    }
  }
}

template <bool WITH_INV, bool PACKED, int MAXT>
__device__ inline void cta_block_sweep_cp(double* S, const int ld, int n_cols, double* PA, const int pa_ld, double* PB,
                                          const int pb_ld, double* D2 /* 2 x PB_NB x PB_D_LD */, double* wv2 /* 2 x PB_NB */,
                                          PhaseTimer* ptm = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tig = lane & 3;
  const int n_rows = n_cols + 1;                 // + residual row
  const int n_rows8 = (n_rows + 7) & ~7;
  const int n_n8 = n_rows8 / 8;
  const int nblk = n_cols / PB_NB;
  if (warp < 2) sweep_pivot_block64<PACKED>(S, ld, 0, D2, wv2, tid);
  __syncthreads();
  if (ptm) ptm->mark(8);
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * PB_NB, c_begin = k0 + PB_NB;
    const double* D = D2 + (kb & 1) * (PB_NB * PB_D_LD);
    const double* wv = wv2 + (kb & 1) * PB_NB;
    const bool last = c_begin >= n_cols;           // no trailing columns: only the residual row's panel entries remain
    const int tn0 = c_begin / 8;                   // first row tile below the pivot block
    const int tn_first = WITH_INV ? 0 : tn0;       // factor rows only: nothing at or above the pivot block counts
    double gt0[4], gt1[4];   // G^T fragments: rows j = gq and gq + 8, columns jp = 4 kk + tig
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int jp = 4 * kk + tig;
      gt0[kk] = (jp < gq) ? D[jp * PB_D_LD + gq] : (jp == gq ? 1.0 : 0.0);
      gt1[kk] = (jp < gq + 8) ? D[jp * PB_D_LD + gq + 8] : (jp == gq + 8 ? 1.0 : 0.0);
    }
    if (last) {
      for (int tn = tn_first + warp; tn < n_n8; tn += 8) sweep_panel_tile<PACKED>(S, ld, n_cols, n_rows, k0, tn * 8, gt0, gt1, PA, pa_ld, PB, pb_ld, D, wv, gq, tig);
    } else if (warp < 2) {
      // ---- the chain: rows / columns of the next pivot block only ----
      const int rho0 = c_begin + 8 * warp;
      sweep_panel_tile<PACKED>(S, ld, n_cols, n_rows, k0, rho0, gt0, gt1, PA, pa_ld, PB, pb_ld, D, wv, gq, tig);
      __threadfence_block();
      bar_arrive_panel();                          // warps 2-7 may now read these panel rows
      bar_sync_pivot64();                          // ... and so may the other warp of the pair
      {
        // 8 x 16 piece of the next diagonal block: S[rho][c] += sum_j PA[rho][j] PB[c][j], c in [c_begin, c_begin + 16)
        double a0[4], a1[4], bfr[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          a0[kk] = PB[(kk * 4 + tig) * pb_ld + c_begin + gq];
          a1[kk] = PB[(kk * 4 + tig) * pb_ld + c_begin + gq + 8];
          bfr[kk] = PA[(kk * 4 + tig) * pa_ld + rho0 + gq];
        }
        double* p0 = sweep_colp<PACKED>(S, ld, c_begin + gq) + (rho0 + 2 * tig);
        double* p1 = sweep_colp<PACKED>(S, ld, c_begin + gq + 8) + (rho0 + 2 * tig);
        const double2 u = *reinterpret_cast<const double2*>(p0), v = *reinterpret_cast<const double2*>(p1);
        double acc[4] = {u.x, u.y, v.x, v.y};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dmma_1684_neg0(acc, a0[kk], a1[kk], bfr[kk]);
        double2 uo, vo;
        uo.x = acc[0]; uo.y = acc[1]; vo.x = acc[2]; vo.y = acc[3];
        *reinterpret_cast<double2*>(p0) = uo;
        *reinterpret_cast<double2*>(p1) = vo;
      }
      bar_sync_pivot64();
      sweep_pivot_block64<PACKED>(S, ld, c_begin, D2 + ((kb + 1) & 1) * (PB_NB * PB_D_LD), wv2 + ((kb + 1) & 1) * PB_NB, tid);
    } else {
      // ---- everything else of step kb ----
      for (int i = tn_first + (warp - 2); i < n_n8 - 2; i += 6) {
        const int tn = i < tn0 ? i : i + 2;        // every row tile but the chain's two
        sweep_panel_tile<PACKED>(S, ld, n_cols, n_rows, k0, tn * 8, gt0, gt1, PA, pa_ld, PB, pb_ld, D, wv, gq, tig);
      }
      __threadfence_block();
      bar_sync_panel();                            // all panel rows of the step are in PA / PB (named barrier 2: 64 arrive, 192 wait)
      const int n_m16 = (n_cols - c_begin) / 16;
      // the next pivot block's columns, rows below that block (its own 16 rows are the chain's) ...
      sweep_trailing_tiles<6, MAXT, WITH_INV, PACKED>(S, ld, n_n8, k0, 0, 1, PA, pa_ld, PB, pb_ld, warp - 2, gq, tig, tn0);
      // ... and every column behind it
      sweep_trailing_tiles<6, MAXT, WITH_INV, PACKED>(S, ld, n_n8, k0, 1, n_m16, PA, pa_ld, PB, pb_ld, warp - 2, gq, tig);
    }
    __syncthreads();
    if (ptm) ptm->mark(10);
  }
}

#define POTRF_BLK_THREADS 512
__global__ void __launch_bounds__(POTRF_BLK_THREADS)
potrf_tile_blk_kernel(int k, int N, double* __restrict__ Abase, size_t walker_stride, double* __restrict__ resid,
                      int ldr, double* __restrict__ zbuf, double* __restrict__ acc_logdet,
                      double* __restrict__ acc_quad, int* __restrict__ status, double* __restrict__ zout,
                  long long ldz) {
  extern __shared__ double smem[];
  double* S = smem;
  double* PA = S + (size_t)PB_LD * TILE;
  double* PBm = PA + PB_NB * PB_PA_LD;
  double* D = PBm + PB_NB * PB_PB_LD;
  double* wv = D + PB_NB * PB_D_LD;
  double* red = wv + PB_NB;
  int* ired = (int*)(red + 32);
  double* isd_s = red + 64;
  const int b = blockIdx.x, tid = threadIdx.x;
  double* tile = Abase + (size_t)b * walker_stride + tile_index(k, k) * TILE_ELEMS;
  double* rk = resid + (size_t)b * ldr + (size_t)k * TILE;

  // 16 independent 16-byte loads in flight per thread (the tile is 128 KiB of latency-bound traffic)
#pragma unroll 1
  for (int base = 0; base < TILE_ELEMS / 2; base += POTRF_BLK_THREADS * 16) {
    double2 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldcs(reinterpret_cast<const double2*>(tile) + base + u * POTRF_BLK_THREADS + tid);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = 2 * (base + u * POTRF_BLK_THREADS + tid);
      const int kg = e >> 9, m8 = (e >> 5) & 15, g = (e >> 2) & 7, k3 = e & 3;
      const int R = m8 * 8 + g, C = kg * 4 + k3;
      S[R + (size_t)C * PB_LD] = v[u].x;
      S[R + (size_t)(C + 1) * PB_LD] = v[u].y;
    }
  }
  if (tid < TILE) {
    S[TILE + (size_t)tid * PB_LD] = rk[tid];
#pragma unroll
    for (int r = TILE + 1; r < PB_LD; ++r) S[r + (size_t)tid * PB_LD] = 0.0;   // padding rows of the n8 tiles
  }
  __syncthreads();
  cta_block_sweep<true, POTRF_BLK_THREADS / 32>(S, PB_LD, TILE, PA, PB_PA_LD, PBm, PB_PB_LD, D, wv);

  double ld_part = 0.0, q_part = 0.0;
  int first_bad = 0x7fffffff, nan_seen = 0;
  double isd = 0.0;
  if (tid < TILE) {
    const double d = S[tid + (size_t)tid * PB_LD];
    if (!(d > 0.0)) {
      first_bad = k * TILE + tid + 1;
      nan_seen = isnan(d) ? 1 : 0;
    }
    isd = 1.0 / sqrt(d);
    ld_part = log(d);
    const double z = S[TILE + (size_t)tid * PB_LD] * isd;
    q_part = z * z;
    zbuf[(size_t)b * TILE + tid] = z;
    if (zout != nullptr && k * TILE + tid < N) zout[(size_t)b * ldz + (size_t)k * TILE + tid] = z;
  }
  const double logdet = block_sum(ld_part, red);
  const double quad = block_sum(q_part, red);
  const int fb = -block_max_int(-first_bad, ired);
  const int anynan = block_max_int(nan_seen, ired);
  if (tid == 0) {
    acc_logdet[b] += logdet;
    acc_quad[b] += quad;
    if (status[b] == 0 && fb != 0x7fffffff) status[b] = anynan ? -1 : fb;
  }
  if (tid < TILE) isd_s[tid] = isd;
  __syncthreads();
  for (int e2 = tid; e2 < TILE_ELEMS / 2; e2 += POTRF_BLK_THREADS) {
    const int e = 2 * e2;
    const int kg = e >> 9, m8 = (e >> 5) & 15, g = (e >> 2) & 7, k3 = e & 3;
    const int R = m8 * 8 + g, C = kg * 4 + k3;
    const double s = isd_s[R];
    double2 o;
    o.x = (C < R) ? S[C + (size_t)R * PB_LD] * s : (C == R ? s : 0.0);
    o.y = (C + 1 < R) ? S[C + 1 + (size_t)R * PB_LD] * s : (C + 1 == R ? s : 0.0);
    reinterpret_cast<double2*>(tile)[e2] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// DMMA tile GEMM: acc(128x128) = sum over operand tile pairs of A_p * B_p^T, K = 128 per pair.
// 8 warps as 2 (m) x 4 (n): warp tile 64 x 32 = 4 m16-fragments x 4 n8-fragments.
// 4-stage bulk-copy / mbarrier ring over k16 slices (16 KiB of A + 16 KiB of B per stage).
// ------------------------------------------------------------------------------------------------
#define GEMM_STAGE_DOUBLES (2 * SLICE_ELEMS)

enum { GEMM_TRSM = 0, GEMM_UPDATE = 1 };

struct GemmArgs {
  int N, nt;
  int mode;          // GEMM_TRSM / GEMM_UPDATE
  int k;             // TRSM: the tile column being solved
  int col_begin;     // UPDATE: first / one-past-last target tile column
  int col_end;
  int p_begin;       // UPDATE: panel of operand tile columns [p_begin, p_end)
  int p_end;
  int gen;           // UPDATE: 1 = target tiles are generated from t (first touch) instead of loaded
  int band;          // UPDATE rasterisation band height in tile rows (0 = default)
  int cs_store;      // UPDATE: streaming (evict-first) stores for the C tile
  int x_off;         // UPDATE: first target-tile index of this launch (diagonal-first split of a column)
  int kernel_id, variant, nh;
  double* Abase;
  size_t walker_stride;
  const double* t;
  const double* yerr;
  const double* hp_wave;
  const double* ph_wave;   // [wave, 2, nt * 128] phase cosines / sines of the points (phase_table_kernel)
  double* resid;     // [wave, ldr]
  int ldr;
  const double* zbuf;  // [wave, 128]
};

// ---- mbarrier / bulk-copy (TMA engine, 1-D) primitives -------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

#define GEMM_TMA_STAGES 4
#define GEMM_TMA_THREADS GEMM_THREADS
#define GEMM_TMA_SMEM_BYTES \
  ((size_t)GEMM_TMA_STAGES * GEMM_STAGE_DOUBLES * sizeof(double) + (6 * TILE + 4 * TILE + TILE) * sizeof(double) + 2 * GEMM_TMA_STAGES * sizeof(unsigned long long))

// L2-aware rasterisation of an UPDATE launch.  Target tiles (ti, tj), col_begin <= tj < col_end,
// ti >= tj, are walked in horizontal BANDS of UPDATE_BAND tile rows, column by column inside a band:
// the band's A operands (UPDATE_BAND panel rows, 1 MiB each at panel width 8) stay L2-resident while
// each B panel row is fetched once per band and shared by the <= UPDATE_BAND co-running CTAs below
// it.  (Plain column-major order re-fetched the whole 100 MiB panel for every target column:
// 127 GB of DRAM reads for 30 GB of algorithmic traffic at N = 16384, ncu r1e.)
#define UPDATE_BAND 64
__device__ __forceinline__ void decode_update_tile(const GemmArgs& g, int x, int& ti, int& tj) {
  int rem = x;
  const int band = g.band > 0 ? g.band : UPDATE_BAND;
  for (int r0 = g.col_begin; r0 < g.nt; r0 += band) {
    const int r1 = min(r0 + band, g.nt) - 1;          // last tile row of the band
    const int c_last = min(g.col_end - 1, r1);                // last target column with rows in the band
    for (int c = g.col_begin; c <= c_last; ++c) {
      const int top = max(c, r0);
      const int cnt = r1 - top + 1;
      if (rem < cnt) {
        tj = c;
        ti = top + rem;
        return;
      }
      rem -= cnt;
    }
  }
  ti = tj = g.nt - 1;  // unreachable for a correctly sized grid
}

// Operand feed: one elected lane streams the k16 operand slices with 1-D bulk copies (cp.async.bulk, the TMA
// engine) into a 4-stage ring guarded by full/empty mbarriers; the 8 consumer warps never meet at a CTA-wide
// barrier and fetch the first fragments of slice q+1 while the last MMAs of slice q issue.  (A 16-byte cp.async
// ring with one __syncthreads per slice and a persistent multi-tile variant were measured 25 % / 0-4 % slower
// in round 1 and removed.)
__global__ void __launch_bounds__(GEMM_TMA_THREADS, 1) gemm_tile_kernel(GemmArgs g) {
  extern __shared__ double smem[];
  constexpr int NST = GEMM_TMA_STAGES;
  double* stages = smem;
  double* tis = smem + NST * GEMM_STAGE_DOUBLES;  // 3 x 128: t, phase cosine, phase sine of the tile rows
  double* tjs = tis + 3 * TILE;                   // 3 x 128: the same of the tile cols
  double* part = tjs + 3 * TILE;                  // 4 x 128 GEMV partials
  double* zs = part + 4 * TILE;                   // 128: z_k
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(zs + TILE);
  unsigned long long* empty_bar = full_bar + GEMM_TMA_STAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 2 x 4 warp grid.  In TRSM the B operand inv(L_kk) is lower triangular, so the warp owning output
  // columns [32 wn, 32 wn + 32) only needs k16 slices 0 .. 2 wn + 1; the two warps that share an SM
  // sub-partition (warp, warp + 4) are given wn and 3 - wn so every sub-partition issues 10 of the
  // 16 warp-slices (balanced 37.5 % saving instead of an idle/busy split).
  const int wm = (warp >> 2) & 1;
  const int wn = (g.mode == GEMM_TRSM && wm) ? 3 - (warp & 3) : (warp & 3);
  const int q_last = (g.mode == GEMM_TRSM) ? 2 * wn + 1 : 0x7fffffff;
  const int gq = lane >> 2, tig = lane & 3;
  const int b = blockIdx.y;
  double* Aw = g.Abase + (size_t)b * g.walker_stride;

  // decode the target tile (ti, tj) and the operand tile list
  int ti, tj, np;
  if (g.mode == GEMM_TRSM) {
    ti = g.k + 1 + blockIdx.x;
    tj = g.k;
    np = 1;
  } else {
    decode_update_tile(g, (int)blockIdx.x + g.x_off, ti, tj);
    np = g.p_end - g.p_begin;
  }
  const int nslices = np * SLICES_PER_TILE;

  auto a_src = [&](int q) -> const double* {
    const int p = q >> 3, s = q & 7;
    const size_t tix = (g.mode == GEMM_TRSM) ? tile_index(ti, g.k) : tile_index(ti, g.p_begin + p);
    return Aw + tix * TILE_ELEMS + s * SLICE_ELEMS;
  };
  auto b_src = [&](int q) -> const double* {
    const int p = q >> 3, s = q & 7;
    const size_t tix = (g.mode == GEMM_TRSM) ? tile_index(g.k, g.k) : tile_index(tj, g.p_begin + p);
    return Aw + tix * TILE_ELEMS + s * SLICE_ELEMS;
  };

  {
    if (tid == 0) {
      for (int s = 0; s < GEMM_TMA_STAGES; ++s) {
        mbar_init(&full_bar[s], 1);   // one arrive (the producer's expect_tx) + 32 KiB of bytes
        mbar_init(&empty_bar[s], 8);  // one arrive per consumer warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncthreads();
    // No dedicated producer warp: 9 warps would be allocated registers as 12 (warp-allocation
    // granularity 4), capping the kernel at 168 registers and spilling the 64-double accumulator.
    // Instead one elected lane of warp 0 issues the two bulk copies per slice, two slices ahead.
    if (tid == 0) {
      for (int q = 0; q < 2 && q < nslices; ++q) {
        double* dst = stages + q * GEMM_STAGE_DOUBLES;
        mbar_expect_tx(&full_bar[q], 2 * SLICE_ELEMS * sizeof(double));
        bulk_g2s(dst, a_src(q), SLICE_ELEMS * sizeof(double), &full_bar[q]);
        bulk_g2s(dst + SLICE_ELEMS, b_src(q), SLICE_ELEMS * sizeof(double), &full_bar[q]);
      }
    }
  }

  if (g.mode == GEMM_UPDATE && g.gen) {
    const double* ph_w = g.ph_wave + (size_t)b * 2 * g.nt * TILE;
    if (tid < TILE) load_tile_points(tis, ti * TILE, tid, g.N, g.nt * TILE, g.t, ph_w);
    else load_tile_points(tjs, tj * TILE, tid - TILE, g.N, g.nt * TILE, g.t, ph_w);
  }
  if (g.mode == GEMM_TRSM && tid < TILE) zs[tid] = g.zbuf[(size_t)b * TILE + tid];

  // Accumulators start at -C_old (UPDATE: loaded from HBM, or generated from t on first touch) so the
  // epilogue is a pure negate-and-store: acc = A B^T - C_old  ->  C_new = -acc.  The 32 16-byte loads
  // per thread land directly in the accumulator registers (maximum memory-level parallelism) and
  // overlap with the operand pipeline fill.  TRSM starts from zero.
  double* Ct = Aw + tile_index(ti, tj) * TILE_ELEMS;
  double acc[4][4][4];
  if (g.mode == GEMM_UPDATE && !g.gen) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int v1 = 0; v1 < 2; ++v1) {
          const int R = wm * 64 + mi * 16 + gq + 8 * v1;
          const int C = wn * 32 + ni * 8 + 2 * tig;
          const double2 old = __ldcs(reinterpret_cast<const double2*>(Ct + tile_off(R, C)));
          acc[mi][ni][2 * v1 + 0] = -old.x;
          acc[mi][ni][2 * v1 + 1] = -old.y;
        }
  } else if (g.mode == GEMM_UPDATE) {
    consumer_sync();  // tis / tjs
    const CovParams cp = make_cov_params(g.kernel_id, g.variant, g.hp_wave + (size_t)b * g.nh);
    const bool interior = ti > tj && (ti + 1) * TILE <= g.N;   // off-diagonal tile, no padding rows
    if (interior && (cp.kind == 3 || cp.kind == 4)) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int np2 = 0; np2 < 2; ++np2)
          gen_interior8_inline<3>(cp, tis, wm * 64 + mi * 16 + gq, tjs, wn * 32 + np2 * 16 + 2 * tig, acc[mi][2 * np2], acc[mi][2 * np2 + 1]);
    } else if (interior) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int np2 = 0; np2 < 2; ++np2) {
          const int R = wm * 64 + mi * 16 + gq;
          const int C = wn * 32 + np2 * 16 + 2 * tig;
          const Cov8 f = gen_interior8(cp.kind, cp.a2, cp.c1, cp.c2, cp.c3, tis, TILE, R, tjs, TILE, C);
          acc[mi][2 * np2][0] = -f.v0; acc[mi][2 * np2][1] = -f.v1; acc[mi][2 * np2][2] = -f.v2; acc[mi][2 * np2][3] = -f.v3;
          acc[mi][2 * np2 + 1][0] = -f.v4; acc[mi][2 * np2 + 1][1] = -f.v5; acc[mi][2 * np2 + 1][2] = -f.v6;
          acc[mi][2 * np2 + 1][3] = -f.v7;
        }
    } else
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int np2 = 0; np2 < 2; ++np2)
        {
          const int R = wm * 64 + mi * 16 + gq;
          const int C = wn * 32 + np2 * 16 + 2 * tig;
          double frag[8];
          gen_fragment8(cp, g.N, ti * TILE + R, tj * TILE + C, tis, TILE, R, tjs, TILE, C, g.yerr, frag);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc[mi][2 * np2][v] = -frag[v];
            acc[mi][2 * np2 + 1][v] = -frag[4 + v];
          }
        }
  } else {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[mi][ni][v] = 0.0;
  }

  // main loop: per k16 slice four k4 steps; within a k4 step the 16 m16n8k4 MMAs (= 32 DMMA.8x8x4)
  // touch 16 different accumulators, so no DMMA ever waits on the previous one's result; the
  // fragments of step kk+1 are fetched from smem while step kk issues.
  const int a_lane = ((wm * 8) << 5) + lane;
  const int b_lane = ((wn * 4) << 5) + lane;
  double af[2][8], bf[2][4];
  {
    mbar_wait(&full_bar[0], 0u);
    const double* As = stages + a_lane;
    const double* Bs = stages + SLICE_ELEMS + b_lane;
#pragma unroll
    for (int v = 0; v < 8; ++v) af[0][v] = As[v << 5];
#pragma unroll
    for (int v = 0; v < 4; ++v) bf[0][v] = Bs[v << 5];
  }
  for (int q = 0; q < nslices; ++q) {
    const int st = q % NST;
    {
      // refill: slice q+2 goes into the stage consumed two iterations ago (one full iteration of
      // slack, so the wait on its empty barrier is normally already satisfied)
      const int qn = q + 2;
      if (tid == 0 && qn < nslices) {
        const int sn = qn % NST, u = qn / NST;
        if (u > 0) mbar_wait(&empty_bar[sn], (unsigned)((u - 1) & 1));
        double* dst = stages + sn * GEMM_STAGE_DOUBLES;
        mbar_expect_tx(&full_bar[sn], 2 * SLICE_ELEMS * sizeof(double));
        bulk_g2s(dst, a_src(qn), SLICE_ELEMS * sizeof(double), &full_bar[sn]);
        bulk_g2s(dst + SLICE_ELEMS, b_src(qn), SLICE_ELEMS * sizeof(double), &full_bar[sn]);
      }
      __syncwarp();
    }
    const double* As = stages + st * GEMM_STAGE_DOUBLES + a_lane;
    const double* Bs = stages + st * GEMM_STAGE_DOUBLES + SLICE_ELEMS + b_lane;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (kk < 3) {
#pragma unroll
        for (int v = 0; v < 8; ++v) af[nxt][v] = As[(((kk + 1) * 16 + v) << 5)];
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[nxt][v] = Bs[(((kk + 1) * 16 + v) << 5)];
      } else if (q + 1 < nslices) {
        // first fragments of the next slice (its stage was filled by the producer long ago)
        const int sn = (q + 1) % NST;
        mbar_wait(&full_bar[sn], (unsigned)(((q + 1) / NST) & 1));
        const double* An = stages + sn * GEMM_STAGE_DOUBLES + a_lane;
        const double* Bn = stages + sn * GEMM_STAGE_DOUBLES + SLICE_ELEMS + b_lane;
#pragma unroll
        for (int v = 0; v < 8; ++v) af[nxt][v] = An[v << 5];
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[nxt][v] = Bn[v << 5];
      }
      if (q <= q_last) {   // warp-uniform
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma_1684(acc[mi][ni], af[cur][2 * mi], af[cur][2 * mi + 1], bf[cur][ni]);
      }
    }
    // this warp has read everything it needs from stage `st`: hand it back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
  }

  // epilogue ------------------------------------------------------------------------------------
  // element (R, C): R = wm*64 + mi*16 + gq + 8*v1, C = wn*32 + ni*8 + 2*tig + v0
  if (g.mode == GEMM_UPDATE) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int v1 = 0; v1 < 2; ++v1) {
          const int R = wm * 64 + mi * 16 + gq + 8 * v1;
          const int C = wn * 32 + ni * 8 + 2 * tig;
          double2 o;
          o.x = -acc[mi][ni][2 * v1 + 0];
          o.y = -acc[mi][ni][2 * v1 + 1];
          if (g.cs_store) __stcs(reinterpret_cast<double2*>(Ct + tile_off(R, C)), o);   // not re-read before the next panel
          else *reinterpret_cast<double2*>(Ct + tile_off(R, C)) = o;
        }
  } else {
    // TRSM: X = acc is L(i,k); store it and fold r_i -= X z_k
    consumer_sync();  // zs
    double rowsum[4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int v1 = 0; v1 < 2; ++v1) {
        double s = 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int R = wm * 64 + mi * 16 + gq + 8 * v1;
          const int C = wn * 32 + ni * 8 + 2 * tig;
          double2 x;
          x.x = acc[mi][ni][2 * v1 + 0];
          x.y = acc[mi][ni][2 * v1 + 1];
          *reinterpret_cast<double2*>(Ct + tile_off(R, C)) = x;
          s = fma(x.x, zs[C], s);
          s = fma(x.y, zs[C + 1], s);
        }
        // reduce over the 4 lanes of a quad (columns), fixed order
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        rowsum[mi][v1] = s;
      }
    if (tig == 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int v1 = 0; v1 < 2; ++v1) part[wn * TILE + wm * 64 + mi * 16 + gq + 8 * v1] = rowsum[mi][v1];
    }
    consumer_sync();
    if (tid < TILE) {
      const double s = ((part[tid] + part[TILE + tid]) + part[2 * TILE + tid]) + part[3 * TILE + tid];
      double* ri = g.resid + (size_t)b * g.ldr + (size_t)ti * TILE;
      ri[tid] -= s;
    }
  }
}

// logL = -N/2 ln 2pi - logdet/2 - quad/2 (+ priors), status merge.  One thread per walker.
__global__ void finalize_logl_kernel(ProblemDesc D, int W, const double* __restrict__ hp,
                                     const double* __restrict__ phys, const double* __restrict__ acc_logdet,
                                     const double* __restrict__ acc_quad, const int* __restrict__ nonfinite,
                                     const int* __restrict__ fact_status, double* __restrict__ logl_out,
                                     int* __restrict__ status_out, double* __restrict__ logdet_out,
                                     const int* __restrict__ abort_flag) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double v = -((double)D.N / 2.0) * log(2.0 * MRV_PI) - 0.5 * acc_logdet[w] - 0.5 * acc_quad[w];
  v = add_priors(D, hp + (size_t)w * D.nh, phys + (size_t)w * D.nm, v);
  int st = fact_status[w];
  if (nonfinite[w]) st = -1;
  if (abort_flag != nullptr && *abort_flag != 0) st = MRV_STATUS_INTERNAL;   // the dataflow kernel gave up on a dependency
  logl_out[w] = v;
  status_out[w] = st;
  if (logdet_out != nullptr) logdet_out[w] = acc_logdet[w];
}
