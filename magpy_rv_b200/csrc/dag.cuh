// Persistent dataflow ("tile DAG") form of the tiled factorisation for 3 ... 32 tile columns (N <= 4096), batched over a wave.
//
// The launch-per-tile-column host loop of abi.cu:run_wave pays, per column, a serial chain
//     in-panel UPDATE -> diagonal-tile kernel (latency-bound, one CTA per walker) -> TRSM
// whose launches each end in a partially filled last round of CTAs.  Here ONE kernel of one CTA per SM runs the whole
// factorisation of a wave as a list of TILE TASKS pulled from an atomic counter; the dependencies between tasks
// are per-walker, per-tile-row progress counters in global memory (release / acquire), so different walkers sit in
// different tile columns at the same time and nothing waits for the slowest CTA of a launch.
//
// Left-looking by tiles, same storage (lower-triangular 128 x 128 tiles in operand order, blocked.cuh):
//   O(b, i, k), i > k   acc = -K(i,k) generated from the points (first and only touch) + sum_{p<k} L(i,p) L(k,p)^T  (DMMA,
//                       operands streamed through the bulk-copy / mbarrier ring; the k tiles of a row are contiguous in
//                       memory); C = -acc goes to SHARED memory in operand order (it never visits HBM); once the
//                       diagonal tile of column k is done: X = C inv(L_kk)^T (DMMA in 16 x 128 warp strips, A fragments
//                       straight from the C tile, inv(L_kk) streamed through a second small ring, the zero triangle skipped
//                       equally by every warp), L(i,k) = X stored, r_i -= X z_k; the progress counter is released one task
//                       later, when the stores have drained.
//   D(b, k)             acc = -K(k,k) + sum_{p<k} L(k,p) L(k,p)^T for the lower triangle only (9 accumulator fragments per
//                       warp); the block goes to shared memory and is swept by cta_block_sweep_la (pivots, inv(L_kk),
//                       z_k = L_kk^-1 r_k, log det, status), bitwise as potrf_tile_blk_kernel does.
// Row i of walker b has `rowdone[b][i]` = number of finished tiles (i,0..); O(b,i,k) needs rowdone[b][i] >= k and
// rowdone[b][k] >= k for its update and rowdone[b][k] >= k + 1 (diagonal tile done) for its solve.
//
// Task order (any order that keeps a walker's dependencies earlier in the list is deadlock-free: the grid is
// co-resident, tasks are taken in list order -- one task ahead at most --, so the smallest unfinished task is always
// running and everything it waits for is done):
//     D(.,0) | O(.,1,0)  O(.,2,0)  D(.,1)  O(.,i>=3,0) | O(.,2,1)  O(.,3,1)  D(.,2)  O(.,i>=4,1) | ...
// i.e. per column s first the two tiles that the next diagonal task and the next column's first task need, then
// the next diagonal tasks (look-ahead: they overlap with the rest of the column), then the rest walker-major (so
// that co-running CTAs share L(s, .) in L2).  A task's inputs that no task produces -- the time stamps / phases of
// its 2 x 128 points and the walker's covariance constants -- are fetched by bulk copies issued during the PREVIOUS
// task of the CTA (double-buffered), so no task starts with a global-memory round trip.
// Summation order per element is the order of the launch-per-column path (p ascending, k16 slices ascending;
// storing C = -acc and reloading it is exact), so results are bitwise identical to it.
#pragma once
#include "blocked.cuh"

struct DagArgs {
  int N, nt, B;
  int kind;                // CovParams::kind of the problem (kernel id, Matern5 textbook -> 7)
  int total_tasks;
  double* Abase;
  size_t walker_stride;
  const double* t;         // [nt * 128] time stamps, zero padded
  const double* yerr;
  const double* ph_wave;   // [B, 2, nt * 128] phase cosines / sines of the points (phase_table_kernel)
  const double* cpar_wave; // [B, 8] covariance constants a2, c1, c2, c3, jit2 (phase_table_kernel)
  double* resid;           // [B, ldr]
  int ldr;
  double* zcols;           // [B, nt, 128]  z_k per tile column
  int* rowdone;            // [B, nt]
  int* counter;            // this wave's task counter
  int* abort_flag;         // one per batch (sticky across its waves)
  double* acc_logdet;      // [B]
  double* acc_quad;
  int* status;
  double* zout;            // nullable [B, ldz]
  long long ldz;
};

#define DAG_THREADS 256
// automatic choice (measured, tools/dag_probe.py, one B200): with the lower-triangle diagonal update, the strip solve and
// the deferred release the dataflow kernel is ahead of the launch-per-column loop from 3 tile columns on (1.12-1.31x at
// N = 384, 1.06x at C3, 1.24x at 640, 1.22x at 1024, 1.11x at 2048, 1.08x at 4096) and bitwise identical to it
#define DAG_AUTO_MIN_TILES 3
#define DAG_MAX_TILES 128             // N <= 16384 (task decoding and the progress counters are sized for it)
#define DAG_BSTAGES 3                 // ring of inv(L_kk) k16 slices for the solve phase
#define DAG_SWEEP_DOUBLES (PB_LD * TILE + PB_NB * PB_PA_LD + PB_NB * PB_PB_LD + 2 * (PB_NB * PB_D_LD + PB_NB) + 64 + TILE)
#define DAG_R1_OFFSET (GEMM_TMA_STAGES * GEMM_STAGE_DOUBLES)   // doubles: end of the update ring / C tile
#define DAG_SOLVE_DOUBLES (DAG_R1_OFFSET + DAG_BSTAGES * SLICE_ELEMS)
#define DAG_MAIN_DOUBLES (DAG_SWEEP_DOUBLES > DAG_SOLVE_DOUBLES ? DAG_SWEEP_DOUBLES : DAG_SOLVE_DOUBLES)
#define DAG_PTS_DOUBLES (6 * TILE + 8)                          // one buffer of point data + covariance constants
#define DAG_AUX_DOUBLES (2 * DAG_PTS_DOUBLES + 4 * TILE + TILE)
#define DAG_NBARS (2 * GEMM_TMA_STAGES + 2 * DAG_BSTAGES + 2)
#define DAG_SMEM_BYTES ((size_t)(DAG_MAIN_DOUBLES + DAG_AUX_DOUBLES) * sizeof(double) + DAG_NBARS * sizeof(unsigned long long) + 16 * sizeof(int) + 17 * sizeof(long long))
#define DAG_SPIN_LIMIT (1ll << 35)   // ~17 s of SM clocks (a wave of 64 walkers at N = 16384 runs 2.7 s): a dependency that takes longer means a lost task, not a slow one

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// experiment switches (tools/build_variant.py): defaults are the measured best
#ifndef DAG_SWEEP_LOOKAHEAD
#define DAG_SWEEP_LOOKAHEAD 1   // diagonal-tile sweep: 1 = next pivot block overlapped with the trailing update; 2 = warps 0-1 own the whole pivot chain (blocked.cuh; measured +1.2...2.5 % up to 8 tile columns, -0.55 % at 32, and selecting between the two by size costs the large sizes 0.4 % through code size alone: not the default)
#endif
#ifdef DAG_KEEP_THREADFENCE
#define DAG_PRE_RELEASE_FENCE() __threadfence()
#else
#define DAG_PRE_RELEASE_FENCE()   // st.release.gpu is itself cumulative over everything ordered before it (bar.sync included)
#endif
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// Wait until *p >= need.  false = aborted (another CTA gave up, or this wait exceeded the limit).
__device__ __forceinline__ bool dag_wait(const int* p, int need, int* abort_flag) {
  if (ld_acquire_gpu(p) >= need) return true;
  const long long t0 = clock64();
  for (;;) {
    __nanosleep(100);
    if (ld_acquire_gpu(p) >= need) return true;
    if (ld_acquire_gpu(abort_flag) != 0) return false;
    if (clock64() - t0 > DAG_SPIN_LIMIT) {
      atomicExch(abort_flag, 1);
      return false;
    }
  }
}

enum { DAG_TASK_EXIT = 0, DAG_TASK_DIAG = 1, DAG_TASK_OFF = 2 };

// -DDAG_TIMING: SM clocks per phase summed over all CTAs (thread 0 of each), and task counts.
//  0 task fetch + dependency wait   1 generation (+ first bulk copies)   2 update main loop   3 C tile -> shared memory
//  4 wait for the diagonal tile     5 solve main loop                    6 solve epilogue + release
//  7 diagonal: acc -> S + residual  8 sweep: pivot blocks  9 sweep: panel rows  10 sweep: trailing DMMA update
// 11 pivots / log det / z           12 inv(L) write-back + release       13 # diagonal tasks  14 # off-diagonal tasks
// 15 update main loop of DIAGONAL tasks (slot 2 then holds the off-diagonal tasks only)
__device__ unsigned long long g_dag_timing[16];

__device__ __forceinline__ void dag_decode(const DagArgs& g, int tsk, int& type, int& b, int& i, int& k) {
  const int B = g.B, nt = g.nt;
  if (tsk < B) {
    type = DAG_TASK_DIAG;
    b = tsk;
    i = k = 0;
    return;
  }
  int r = tsk - B, s = 0;
  while (s < nt - 1 && r >= B * (nt - s)) {
    r -= B * (nt - s);
    ++s;
  }
  k = s;
  type = DAG_TASK_OFF;
  if (r < B) {                                 // the tile the next diagonal task needs
    b = r;
    i = s + 1;
    return;
  }
  r -= B;
  if (s + 2 < nt) {                            // the tile the next column's first task needs
    if (r < B) {
      b = r;
      i = s + 2;
      return;
    }
    r -= B;
  }
  if (r < B) {                                 // next diagonal tile (look-ahead)
    type = DAG_TASK_DIAG;
    b = r;
    i = k = s + 1;
    return;
  }
  r -= B;                                      // rest of column s, walker-major
  const int per = nt - s - 3;
  b = r / per;
  i = s + 3 + r % per;
}

// k16-slice main loop of the update of an OFF-DIAGONAL tile: A | B slice pairs through the 4-stage ring, 2 x 4 warp grid
// (warp tile 64 x 32), fragments of the next k4 step fetched while the current one issues.  `q0` is the ring's running
// slice counter (stage and barrier phase persist across tasks).
__device__ __forceinline__ void dag_mainloop(double (&acc)[4][4][4], int nslices, unsigned q0, const double* ring,
                                             unsigned long long* fbar, unsigned long long* ebar,
                                             const double* __restrict__ Asrc, const double* __restrict__ Bsrc, int a_lane,
                                             int b_lane, int tid, int lane) {
  constexpr unsigned NSTG = GEMM_TMA_STAGES;
  constexpr unsigned slice_bytes = SLICE_ELEMS * sizeof(double);
  auto a_ptr = [&](unsigned st) -> const double* { return ring + st * GEMM_STAGE_DOUBLES + a_lane; };
  auto b_ptr = [&](unsigned st) -> const double* { return ring + st * GEMM_STAGE_DOUBLES + SLICE_ELEMS + b_lane; };
  double af[2][8], bf[2][4];
  {
    const unsigned st0 = q0 % NSTG;
    mbar_wait(&fbar[st0], (q0 / NSTG) & 1);
    const double* As = a_ptr(st0);
    const double* Bs = b_ptr(st0);
#pragma unroll
    for (int v = 0; v < 8; ++v) af[0][v] = As[v << 5];
#pragma unroll
    for (int v = 0; v < 4; ++v) bf[0][v] = Bs[v << 5];
  }
  for (int q = 0; q < nslices; ++q) {
    const unsigned qa = q0 + q, st = qa % NSTG;
    {
      // refill: slice q + 2 goes into the stage consumed two iterations ago
      const int qn = q + 2;
      if (tid == 0 && qn < nslices) {
        const unsigned qq = q0 + qn, sn = qq % NSTG, u = qq / NSTG;
        if (u > 0) mbar_wait(&ebar[sn], (u - 1) & 1);
        double* dst = const_cast<double*>(ring) + sn * GEMM_STAGE_DOUBLES;
        mbar_expect_tx(&fbar[sn], 2 * slice_bytes);
        bulk_g2s(dst, Asrc + (size_t)qn * SLICE_ELEMS, slice_bytes, &fbar[sn]);
        bulk_g2s(dst + SLICE_ELEMS, Bsrc + (size_t)qn * SLICE_ELEMS, slice_bytes, &fbar[sn]);
      }
      __syncwarp();
    }
    const double* As = a_ptr(st);
    const double* Bs = b_ptr(st);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (kk < 3) {
#pragma unroll
        for (int v = 0; v < 8; ++v) af[nxt][v] = As[(((kk + 1) * 16 + v) << 5)];
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[nxt][v] = Bs[(((kk + 1) * 16 + v) << 5)];
      } else if (q + 1 < nslices) {
        const unsigned qn1 = qa + 1, sn = qn1 % NSTG;
        mbar_wait(&fbar[sn], (qn1 / NSTG) & 1);
        const double* An = a_ptr(sn);
        const double* Bn = b_ptr(sn);
#pragma unroll
        for (int v = 0; v < 8; ++v) af[nxt][v] = An[v << 5];
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[nxt][v] = Bn[v << 5];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma_1684(acc[mi][ni], af[cur][2 * mi], af[cur][2 * mi + 1], bf[cur][ni]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ebar[st]);
  }
}

// Update of a DIAGONAL tile, lower triangle only, dealt by m16 x n8 accumulator fragments: fragment (u, v) = rows
// [16 u, 16 u + 16) x columns [8 v, 8 v + 8) is needed for v <= 2 u + 1, 72 of the 128 fragments of the tile -- exactly 9 per
// warp, and 18 per SM sub-partition (warps w and w + 4):
//   warp          0        1              2        3              4        5              6        7
//   segment A   u7 v0-8  u7 v9-15       u6 v0-8  u6 v9-13       u5 v0-8  u2 v0-5        u4 v0-8  u3 v0-7
//   segment B            u0 v0-1                 u1 v0-3                 u5 v9-11                u4 v9
// so a diagonal tile's update costs 9/16 of a full tile product (the 2 x 4 warp grid: 16/16; whole fragment rows dealt
// 3,3,3,3 | 2,2,2,2: 12/16 -- a lone warp cannot keep its sub-partition's FP64 tensor pipe busy, so the largest share
// of any WARP is what counts).  Segment A lives in accumulator slots 0..8 (acc[s >> 2][s & 3]), segment B in acc[3][0..3].
// Each warp runs its own code copy (compile-time segment shapes: no predicates around the DMMAs); summation order per
// element is unchanged (k ascending), so the tile is bitwise what the full update gives.
__device__ __forceinline__ void dag_diag_segments(int warp, int& uA, int& vA0, int& nA, int& uB, int& vB0, int& nB) {
  const int sh = 4 * warp;
  uA = (0x34256677u >> sh) & 15;
  vA0 = (0x00009090u >> sh) & 15;
  nA = (0x89695979u >> sh) & 15;
  uB = (0x40501000u >> sh) & 15;
  vB0 = (0x90900000u >> sh) & 15;
  nB = (0x10304020u >> sh) & 15;
}
template <int UA, int VA0, int NA, int UB, int VB0, int NB>
__device__ __forceinline__ void dag_diag_loop(double (&acc)[4][4][4], int nslices, unsigned q0, const double* ring,
                                              unsigned long long* fbar, unsigned long long* ebar,
                                              const double* __restrict__ Asrc, int tid, int lane) {
  constexpr unsigned NSTG = GEMM_TMA_STAGES;
  constexpr unsigned slice_bytes = SLICE_ELEMS * sizeof(double);
  constexpr int NF = 2 + NA + (NB > 0 ? 2 + NB : 0);   // fragments of one k4 step: A rows of segment A, its B columns, same for B
  double fr[2][NF];
  auto load = [&](double (&f)[NF], const double* As, const double* Bs, int kk) {
    f[0] = As[(kk * 16 + 2 * UA) << 5];
    f[1] = As[(kk * 16 + 2 * UA + 1) << 5];
#pragma unroll
    for (int s2 = 0; s2 < NA; ++s2) f[2 + s2] = Bs[(kk * 16 + VA0 + s2) << 5];
    if (NB > 0) {
      f[2 + NA] = As[(kk * 16 + 2 * UB) << 5];
      f[3 + NA] = As[(kk * 16 + 2 * UB + 1) << 5];
#pragma unroll
      for (int s2 = 0; s2 < NB; ++s2) f[4 + NA + s2] = Bs[(kk * 16 + VB0 + s2) << 5];
    }
  };
  auto compute = [&](const double (&f)[NF]) {
#pragma unroll
    for (int s2 = 0; s2 < NA; ++s2) dmma_1684(acc[s2 >> 2][s2 & 3], f[0], f[1], f[2 + s2]);
    if (NB > 0) {
#pragma unroll
      for (int s2 = 0; s2 < NB; ++s2) dmma_1684(acc[3][s2], f[2 + NA], f[3 + NA], f[4 + NA + s2]);
    }
  };
  {
    const unsigned st0 = q0 % NSTG;
    mbar_wait(&fbar[st0], (q0 / NSTG) & 1);
    const double* As = ring + st0 * GEMM_STAGE_DOUBLES + lane;   // the slice of tile row k is both operands
    load(fr[0], As, As, 0);
  }
#pragma unroll 1
  for (int q = 0; q < nslices; ++q) {
    const unsigned qa = q0 + q, st = qa % NSTG;
    {
      const int qn = q + 2;   // refill: slice q + 2 goes into the stage consumed two iterations ago
      if (tid == 0 && qn < nslices) {
        const unsigned qq = q0 + qn, sn = qq % NSTG, u = qq / NSTG;
        if (u > 0) mbar_wait(&ebar[sn], (u - 1) & 1);
        double* dst = const_cast<double*>(ring) + sn * GEMM_STAGE_DOUBLES;
        mbar_expect_tx(&fbar[sn], slice_bytes);
        bulk_g2s(dst, Asrc + (size_t)qn * SLICE_ELEMS, slice_bytes, &fbar[sn]);
      }
      __syncwarp();
    }
    const double* As = ring + st * GEMM_STAGE_DOUBLES + lane;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (kk < 3) load(fr[nxt], As, As, kk + 1);
      else if (q + 1 < nslices) {
        const unsigned qn1 = qa + 1, sn = qn1 % NSTG;
        mbar_wait(&fbar[sn], (qn1 / NSTG) & 1);
        const double* An = ring + sn * GEMM_STAGE_DOUBLES + lane;
        load(fr[nxt], An, An, 0);
      }
      compute(fr[cur]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ebar[st]);
  }
}
__device__ __forceinline__ void dag_diag_update(double (&acc)[4][4][4], int nslices, unsigned q0, const double* ring,
                                                unsigned long long* fbar, unsigned long long* ebar,
                                                const double* __restrict__ Asrc, const double* __restrict__ Bsrc, int warp, int tid,
                                                int lane) {
  switch (warp) {   // one code copy per warp; only warp 0's contains a live producer branch
    case 0: dag_diag_loop<7, 0, 9, 0, 0, 0>(acc, nslices, q0, ring, fbar, ebar, Asrc, tid, lane); break;
    case 1: dag_diag_loop<7, 9, 7, 0, 0, 2>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    case 2: dag_diag_loop<6, 0, 9, 0, 0, 0>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    case 3: dag_diag_loop<6, 9, 5, 1, 0, 4>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    case 4: dag_diag_loop<5, 0, 9, 0, 0, 0>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    case 5: dag_diag_loop<2, 0, 6, 5, 9, 3>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    case 6: dag_diag_loop<4, 0, 9, 0, 0, 0>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
    default: dag_diag_loop<3, 0, 8, 4, 9, 1>(acc, nslices, q0, ring, fbar, ebar, Asrc, 1, lane); break;
  }
}

// Solve phase X = C inv(L_kk)^T in WARP STRIPS: warp w owns the 16 rows [16 w, 16 w + 16) of all 128 columns (16 n8
// accumulator fragments).  inv(L_kk) is lower triangular, so k16 slice q only feeds the columns >= 16 q, i.e. the
// fragment pairs vp >= q: every warp issues the same 36 / 64 of a full tile product -- the triangle is skipped without
// any imbalance between warps (the 2 x 4 warp grid of the update had the warps of column block 3 do all 8 slices, and a
// single warp cannot use its sub-partition's FP64 tensor pipe alone).  A fragments come straight from the resident C
// tile (slice q at q * SLICE_ELEMS), inv(L_kk) slices through the 3-stage ring.  One code copy per slice (Q is a compile
// time constant), so the skip costs no predicates.
template <int Q>
__device__ __forceinline__ void dag_solve_load(double (&f)[18], const double* __restrict__ As, const double* __restrict__ Bs, int kk) {
  f[0] = As[(kk * 16) << 5];
  f[1] = As[(kk * 16 + 1) << 5];
#pragma unroll
  for (int v = 2 * Q; v < 16; ++v) f[2 + v] = Bs[(kk * 16 + v) << 5];
}
template <int Q>
__device__ __forceinline__ void dag_solve_slice(double (&acc)[4][4][4], double (&fr)[2][18], unsigned q0, const double* Ctile,
                                                const double* ringB, unsigned long long* fbar, unsigned long long* ebar,
                                                const double* __restrict__ Bsrc, int a_lane, int tid, int lane) {
  constexpr unsigned NSTG = DAG_BSTAGES;
  constexpr unsigned slice_bytes = SLICE_ELEMS * sizeof(double);
  const unsigned qa = q0 + Q, st = qa % NSTG;
  if (Q + 2 < SLICES_PER_TILE) {   // refill: slice Q + 2 goes into the stage consumed one iteration ago
    if (tid == 0) {
      const unsigned qq = q0 + Q + 2, sn = qq % NSTG, u = qq / NSTG;
      if (u > 0) mbar_wait(&ebar[sn], (u - 1) & 1);
      mbar_expect_tx(&fbar[sn], slice_bytes);
      bulk_g2s(const_cast<double*>(ringB) + sn * SLICE_ELEMS, Bsrc + (size_t)(Q + 2) * SLICE_ELEMS, slice_bytes, &fbar[sn]);
    }
    __syncwarp();
  }
  const double* As = Ctile + (size_t)Q * SLICE_ELEMS + a_lane;
  const double* Bs = ringB + st * SLICE_ELEMS + lane;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int cur = (Q * 4 + kk) & 1, nxt = cur ^ 1;
    if (kk < 3) dag_solve_load<Q>(fr[nxt], As, Bs, kk + 1);
    else if (Q + 1 < SLICES_PER_TILE) {
      const unsigned qn1 = qa + 1, sn = qn1 % NSTG;
      mbar_wait(&fbar[sn], (qn1 / NSTG) & 1);
      dag_solve_load<(Q + 1 < SLICES_PER_TILE ? Q + 1 : Q)>(fr[nxt], Ctile + (size_t)(Q + 1) * SLICE_ELEMS + a_lane, ringB + sn * SLICE_ELEMS + lane, 0);
    }
#pragma unroll
    for (int v = 2 * Q; v < 16; ++v) dmma_1684(acc[v >> 2][v & 3], fr[cur][0], fr[cur][1], fr[cur][2 + v]);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(&ebar[st]);
}
__device__ __forceinline__ void dag_solve_strips(double (&acc)[4][4][4], unsigned q0, const double* Ctile, const double* ringB,
                                                 unsigned long long* fbar, unsigned long long* ebar,
                                                 const double* __restrict__ Bsrc, int warp, int tid, int lane) {
  const int a_lane = ((2 * warp) << 5) + lane;
  double fr[2][18];
  {
    const unsigned st0 = q0 % DAG_BSTAGES;
    mbar_wait(&fbar[st0], (q0 / DAG_BSTAGES) & 1);
    dag_solve_load<0>(fr[0], Ctile + a_lane, ringB + st0 * SLICE_ELEMS + lane, 0);
  }
  dag_solve_slice<0>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<1>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<2>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<3>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<4>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<5>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<6>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
  dag_solve_slice<7>(acc, fr, q0, Ctile, ringB, fbar, ebar, Bsrc, a_lane, tid, lane);
}

__global__ void __launch_bounds__(DAG_THREADS, 1) tile_dag_kernel(DagArgs g) {
  extern __shared__ double smem[];
  constexpr int NST = GEMM_TMA_STAGES;
  double* stages = smem;                          // R0: operand ring, 4 x 32 KiB | C tile of the solve phase | S
  double* S = smem;
  double* bring = smem + DAG_R1_OFFSET;           // R1 (behind the C tile): inv(L_kk) slices of the solve phase | rest of S, PA ...
  double* PA = S + (size_t)PB_LD * TILE;
  double* PBm = PA + PB_NB * PB_PA_LD;
  double* Dm = PBm + PB_NB * PB_PB_LD;
  double* wv = Dm + 2 * PB_NB * PB_D_LD;          // (pivot-block buffers double-buffered: cta_block_sweep_cp)
  double* red = wv + 2 * PB_NB;
  int* ired = (int*)(red + 32);
  double* isd_s = red + 64;
  double* pts = smem + DAG_MAIN_DOUBLES;          // never aliased: 2 x (rows t, C, S | cols t, C, S | constants)
  double* part = pts + 2 * DAG_PTS_DOUBLES;
  double* zs = part + 4 * TILE;
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(zs + TILE);
  unsigned long long* empty_bar = full_bar + NST;
  unsigned long long* fullB = empty_bar + NST;
  unsigned long long* emptyB = fullB + DAG_BSTAGES;
  unsigned long long* pt_bar = emptyB + DAG_BSTAGES;
  int* s_task = reinterpret_cast<int*>(pt_bar + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) & 1;
  const int gq = lane >> 2, tig = lane & 3;
  int* abort_flag = g.abort_flag;
  const int n_pad = g.nt * TILE;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    for (int s = 0; s < DAG_BSTAGES; ++s) {
      mbar_init(&fullB[s], 1);
      mbar_init(&emptyB[s], 8);
    }
    mbar_init(&pt_bar[0], 1);
    mbar_init(&pt_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  unsigned qg = 0;        // k16 slices streamed through the update ring so far (uniform across the CTA)
  unsigned qb = 0;        // inv(L) slices streamed through the solve ring so far
  unsigned n_local = 0;   // tasks this CTA has started: point buffer = n_local & 1
  PhaseTimer tm;
  tm.init(s_task + 16);

  // Point data of task `tsk` (thread 0): six 1 KiB bulk copies + the walker's constants into buffer `buf`.
  auto issue_points = [&](int tsk, unsigned buf) {
    int type, b, i, k;
    dag_decode(g, tsk, type, b, i, k);
    double* dst = pts + buf * DAG_PTS_DOUBLES;
    const double* ph_w = g.ph_wave + (size_t)b * 2 * n_pad;
    unsigned long long* bar = &pt_bar[buf];
    mbar_expect_tx(bar, (6 * TILE + 8) * sizeof(double));
    bulk_g2s(dst, g.t + i * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + TILE, ph_w + i * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + 2 * TILE, ph_w + n_pad + i * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + 3 * TILE, g.t + k * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + 4 * TILE, ph_w + k * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + 5 * TILE, ph_w + n_pad + k * TILE, TILE * sizeof(double), bar);
    bulk_g2s(dst + 6 * TILE, g.cpar_wave + (size_t)b * 8, 8 * sizeof(double), bar);
  };

  // Deferred release (thread 0): the progress counter of a finished off-diagonal tile is published after the NEXT task's
  // generation phase instead of right behind the tile's 128 KiB of stores -- the fence then finds them drained instead
  // of waiting for them.  It is flushed early whenever this CTA is about to wait (the awaited task may itself be waiting
  // for this counter) or to leave.  Diagonal tiles, which a whole tile column waits for, are released at once.
  int* pend_ptr = nullptr;
  int pend_val = 0;
  auto flush_release = [&]() {
    if (pend_ptr != nullptr) {
      DAG_PRE_RELEASE_FENCE();
      st_release_gpu(pend_ptr, pend_val);
      pend_ptr = nullptr;
    }
  };

  int next_tsk = 0x7fffffff;
  if (tid == 0) {
    if (ld_acquire_gpu(abort_flag) == 0) next_tsk = atomicAdd(g.counter, 1);
    if (next_tsk < g.total_tasks) issue_points(next_tsk, 0);
  }

  for (;;) {
    // ---- next task ---------------------------------------------------------------------------
    if (warp == 0) {
      // The three dependency words of the task (abort flag, progress of tile rows i and k) are acquired by three LANES in
      // one instruction -- one L2 round trip instead of three back to back (an acquire load completes before the next
      // one issues) -- and handed to lane 0 through the warp barrier of the shuffles.
      const int tsk = __shfl_sync(0xffffffffu, next_tsk, 0);
      int type = DAG_TASK_EXIT, b = 0, i = 0, k = 0, diag_ready = 0;
      int dep = 0;
      const bool live = tsk < g.total_tasks;
      if (live) {
        dag_decode(g, tsk, type, b, i, k);
        const int* rd = g.rowdone + (size_t)b * g.nt;
        if (lane < 3) dep = ld_acquire_gpu(lane == 0 ? abort_flag : (lane == 1 ? rd + i : rd + k));
      }
      __syncwarp();
      const int aborted = __shfl_sync(0xffffffffu, dep, 0), v1 = __shfl_sync(0xffffffffu, dep, 1), v2 = __shfl_sync(0xffffffffu, dep, 2);
      if (lane == 0) {
        next_tsk = atomicAdd(g.counter, 1);   // the task after this one: its round trip overlaps with this task
        if (live) {
          if (aborted != 0) type = DAG_TASK_EXIT;
          else {
            // update operands: rows i and k finished through column k - 1 (also orders the residual block r_i / r_k)
            const int* rd = g.rowdone + (size_t)b * g.nt;
            if (v1 < k || v2 < k) {
              flush_release();
              if (!(dag_wait(rd + i, k, abort_flag) && dag_wait(rd + k, k, abort_flag))) type = DAG_TASK_EXIT;
            }
            diag_ready = (type == DAG_TASK_OFF && v2 >= k + 1) ? 1 : 0;
            fence_proxy_async();
          }
          // this task's points were requested during the previous one: never leave with a bulk copy in flight
          if (type == DAG_TASK_EXIT) mbar_wait(&pt_bar[n_local & 1], (n_local >> 1) & 1);
        }
        if (type == DAG_TASK_EXIT) flush_release();
        s_task[0] = type;
        s_task[1] = b;
        s_task[2] = i;
        s_task[3] = k;
        s_task[4] = diag_ready;
      }
    }
    __syncthreads();
    const int type = s_task[0], b = s_task[1], ti = s_task[2], tk = s_task[3];
    if (type == DAG_TASK_EXIT) break;
    tm.mark(0);
    tm.count(type == DAG_TASK_DIAG ? 13 : 14);

    double* Aw = g.Abase + (size_t)b * g.walker_stride;
    const double* rowA = Aw + tile_index(ti, 0) * TILE_ELEMS;   // tiles (ti, 0..): contiguous
    const double* rowB = Aw + tile_index(tk, 0) * TILE_ELEMS;
    double* Ct = Aw + tile_index(ti, tk) * TILE_ELEMS;
    int* rd = g.rowdone + (size_t)b * g.nt;
    const double* pr = pts + (n_local & 1) * DAG_PTS_DOUBLES;   // points of the tile rows; + 3 TILE: of the tile columns
    const double* pc = pr + 3 * TILE;
    bool b_started = false;   // thread 0: the first inv(L) slices of the solve phase are already in flight

    // first operand slices of the update, and -- when the diagonal tile is already done -- of the solve
    if (tid == 0) {
      const int nsl = tk * SLICES_PER_TILE;
      for (int q = 0; q < 2 && q < nsl; ++q) {
        const unsigned qq = qg + q, sn = qq % NST, u = qq / NST;
        if (u > 0) mbar_wait(&empty_bar[sn], (u - 1) & 1);
        double* dst = stages + sn * GEMM_STAGE_DOUBLES;
        if (s_task[0] == DAG_TASK_DIAG) {   // rows i and k are the same tile row: one slice serves as A and B
          mbar_expect_tx(&full_bar[sn], SLICE_ELEMS * sizeof(double));
          bulk_g2s(dst, rowA + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
          continue;
        }
        mbar_expect_tx(&full_bar[sn], 2 * SLICE_ELEMS * sizeof(double));
        bulk_g2s(dst, rowA + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
        bulk_g2s(dst + SLICE_ELEMS, rowB + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
      }
      if (s_task[4]) {
        const double* Bsrc = rowB + (size_t)tk * TILE_ELEMS;
        for (int q = 0; q < 2; ++q) {
          const unsigned qq = qb + q, sn = qq % DAG_BSTAGES, u = qq / DAG_BSTAGES;
          if (u > 0) mbar_wait(&emptyB[sn], (u - 1) & 1);
          mbar_expect_tx(&fullB[sn], SLICE_ELEMS * sizeof(double));
          bulk_g2s(bring + sn * SLICE_ELEMS, Bsrc + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &fullB[sn]);
        }
        b_started = true;
      }
    }
    mbar_wait(&pt_bar[n_local & 1], (n_local >> 1) & 1);   // this task's points (issued during the previous task)

    // The residual block this task reads at its very end (r_i of an off-diagonal tile: lanes tig = 0 of warp w hold rows
    // 16 w + gq and + 8; r_k of a diagonal tile: one row per thread) is final once the task's dependencies hold, so it is
    // fetched now -- the L2 round trip then overlaps the whole task instead of sitting in its epilogue.
    double rpre0 = 0.0, rpre1 = 0.0, ld_pre = 0.0, q_pre = 0.0;
    int st_pre = 0;
    {
      const double* rblk = g.resid + (size_t)b * g.ldr + (size_t)ti * TILE;
      if (type == DAG_TASK_DIAG) {
        if (tid < TILE) rpre0 = __ldcg(rblk + tid);
        if (tid == 0) {   // the walker's running log det / quadratic form / status: only its diagonal tasks touch them, in order
          ld_pre = __ldcg(g.acc_logdet + b);
          q_pre = __ldcg(g.acc_quad + b);
          st_pre = __ldcg(g.status + b);
        }
      } else if (tig == 0) {
        rpre0 = __ldcg(rblk + warp * 16 + gq);
        rpre1 = __ldcg(rblk + warp * 16 + gq + 8);
      }
    }
    double acc[4][4][4];
    // off-diagonal tiles: 2 x 4 warp grid, warp tile = rows [rb, rb + 64) x columns [32 wn0, 32 wn0 + 32); diagonal tiles:
    // 9 accumulator fragments per warp (dag_diag_segments); the solve re-deals the tile in 16 x 128 warp strips
    const int wn0 = warp & 3, rb = wm * 64;
    {
      // ---- generation: acc = -K(ti, tk) ----
      const int wn = wn0;
      CovParams cp;
      cp.kind = g.kind;
      cp.a2 = pr[6 * TILE]; cp.c1 = pr[6 * TILE + 1]; cp.c2 = pr[6 * TILE + 2]; cp.c3 = pr[6 * TILE + 3];
      cp.jit2 = pr[6 * TILE + 4]; cp.cph = 0.0;
      const bool interior = ti > tk && (ti + 1) * TILE <= g.N;
      if (type == DAG_TASK_DIAG) {
        int uA, vA0, nA, uB, vB0, nB;
        dag_diag_segments(warp, uA, vA0, nA, uB, vB0, nB);
#pragma unroll
        for (int s2 = 0; s2 < 13; ++s2) {
          const bool segA = s2 < 9;
          if (segA ? (s2 >= nA) : (s2 - 9 >= nB)) continue;
          const int R = 16 * (segA ? uA : uB) + gq;
          const int C = 8 * (segA ? vA0 + s2 : vB0 + s2 - 9) + 2 * tig;
          double frag[4];
          gen_fragment4(cp, g.N, tk * TILE + R, tk * TILE + C, pr, TILE, R, pc, TILE, C, g.yerr, frag);
          const int slot = segA ? s2 : 12 + (s2 - 9);
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[slot >> 2][slot & 3][v] = -frag[v];
        }
      } else
      if (interior && (cp.kind == 3 || cp.kind == 4)) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int np2 = 0; np2 < 2; ++np2)
            gen_interior8_inline<3>(cp, pr, rb + mi * 16 + gq, pc, wn * 32 + np2 * 16 + 2 * tig, acc[mi][2 * np2],
                                    acc[mi][2 * np2 + 1]);
      } else if (interior) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int np2 = 0; np2 < 2; ++np2) {
            const int R = rb + mi * 16 + gq;
            const int C = wn * 32 + np2 * 16 + 2 * tig;
            const Cov8 f = gen_interior8(cp.kind, cp.a2, cp.c1, cp.c2, cp.c3, pr, TILE, R, pc, TILE, C);
            acc[mi][2 * np2][0] = -f.v0; acc[mi][2 * np2][1] = -f.v1; acc[mi][2 * np2][2] = -f.v2; acc[mi][2 * np2][3] = -f.v3;
            acc[mi][2 * np2 + 1][0] = -f.v4; acc[mi][2 * np2 + 1][1] = -f.v5; acc[mi][2 * np2 + 1][2] = -f.v6;
            acc[mi][2 * np2 + 1][3] = -f.v7;
          }
      } else {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int np2 = 0; np2 < 2; ++np2) {
            const int R = rb + mi * 16 + gq;
            const int C = wn * 32 + np2 * 16 + 2 * tig;
            double frag[8];
            gen_fragment8(cp, g.N, ti * TILE + R, tk * TILE + C, pr, TILE, R, pc, TILE, C, g.yerr, frag);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc[mi][2 * np2][v] = -frag[v];
              acc[mi][2 * np2 + 1][v] = -frag[4 + v];
            }
          }
      }
      // the next task's points: its index has long arrived, the other buffer was last read a task ago
      if (tid == 0) {
        flush_release();
        if (next_tsk < g.total_tasks) issue_points(next_tsk, (n_local + 1) & 1);
      }
      tm.mark(1);
    }
    // ---- update: acc += sum_p L(ti,p) L(tk,p)^T ----
    if (tk > 0) {
      if (type == DAG_TASK_DIAG)
        dag_diag_update(acc, tk * SLICES_PER_TILE, qg, stages, full_bar, empty_bar, rowA, rowB, warp, tid, lane);
      else
        dag_mainloop(acc, tk * SLICES_PER_TILE, qg, stages, full_bar, empty_bar, rowA, rowB, ((wm * 8) << 5) + lane,
                     ((wn0 * 4) << 5) + lane, tid, lane);
      qg += tk * SLICES_PER_TILE;
    }
    tm.mark(type == DAG_TASK_DIAG ? 15 : 2);

    bool aborted = false;
    if (type == DAG_TASK_OFF) {
      // ---- solve phase set-up: C = -acc into R0 (operand order), inv(L_kk) slices into R1 ----
      const double* Bsrc = rowB + (size_t)tk * TILE_ELEMS;
      __syncthreads();   // every warp has left the update ring: C aliases it
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int v1 = 0; v1 < 2; ++v1) {
            const int R = wm * 64 + mi * 16 + gq + 8 * v1;
            const int C = wn0 * 32 + ni * 8 + 2 * tig;
            double2 o;
            o.x = -acc[mi][ni][2 * v1 + 0];
            o.y = -acc[mi][ni][2 * v1 + 1];
            *reinterpret_cast<double2*>(stages + tile_off(R, C)) = o;
          }
      tm.mark(3);
      if (tid == 0 && !b_started) {   // (b_started: the acquire at the start of the task already saw the diagonal tile done)
        if (!dag_wait(rd + tk, tk + 1, abort_flag)) s_task[0] = DAG_TASK_EXIT;
        else {
          fence_proxy_async();
          for (int q = 0; q < 2; ++q) {
            const unsigned qq = qb + q, sn = qq % DAG_BSTAGES, u = qq / DAG_BSTAGES;
            if (u > 0) mbar_wait(&emptyB[sn], (u - 1) & 1);
            mbar_expect_tx(&fullB[sn], SLICE_ELEMS * sizeof(double));
            bulk_g2s(bring + sn * SLICE_ELEMS, Bsrc + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &fullB[sn]);
          }
        }
      }
      __syncthreads();   // C tile complete; diagonal tile known to be done (z_k visible)
      tm.mark(4);
      if (s_task[0] == DAG_TASK_EXIT) {
        // abort: the first inv(L) slices may be in flight -- let them land before this CTA can leave
        if (tid == 0 && b_started) {
          mbar_wait(&fullB[qb % DAG_BSTAGES], (qb / DAG_BSTAGES) & 1);
          mbar_wait(&fullB[(qb + 1) % DAG_BSTAGES], ((qb + 1) / DAG_BSTAGES) & 1);
        }
        aborted = true;
      } else {
        // values other CTAs produced are read past L1 (ld.global.cg): an SM's L1 may hold a stale line of them
        if (tid < TILE) zs[tid] = __ldcg(g.zcols + ((size_t)b * g.nt + tk) * TILE + tid);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[mi][ni][v] = 0.0;
        dag_solve_strips(acc, qb, stages, bring, fullB, emptyB, Bsrc, warp, tid, lane);
        qb += SLICES_PER_TILE;
      }
      tm.mark(5);
    }

    {
      if (aborted) {
        // nothing to store
      } else
      // ---- epilogues ---------------------------------------------------------------------------
      if (type == DAG_TASK_OFF) {
        // X = acc is L(i,k): store it and fold r_i -= X z_k
        __syncthreads();   // zs; every warp has read its last A fragments from the C tile
        // strip layout: fragment v = 4 cb + ni of the warp's 16 rows.  Row sums in the order of the 2 x 4 layout (per
        // column block cb: ni ascending, then the quad's lanes; then ((p0 + p1) + p2) + p3), so r_i is bitwise the same.
#pragma unroll
        for (int v1 = 0; v1 < 2; ++v1) {
          const int R = warp * 16 + gq + 8 * v1;
          double p4[4];
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            double s = 0.0;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
              const int C = cb * 32 + ni * 8 + 2 * tig;
              double2 x;
              x.x = acc[cb][ni][2 * v1 + 0];
              x.y = acc[cb][ni][2 * v1 + 1];
              *reinterpret_cast<double2*>(Ct + tile_off(R, C)) = x;
              s = fma(x.x, zs[C], s);
              s = fma(x.y, zs[C + 1], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            p4[cb] = s;
          }
          if (tig == 0) {
            double* ri = g.resid + (size_t)b * g.ldr + (size_t)ti * TILE;
            ri[R] = (v1 == 0 ? rpre0 : rpre1) - (((p4[0] + p4[1]) + p4[2]) + p4[3]);
          }
        }
        __syncthreads();
        // one fence after the barrier covers the whole CTA's stores (fence cumulativity), then the release
        if (tid == 0) {
#ifdef DAG_EAGER_RELEASE
          DAG_PRE_RELEASE_FENCE();
          st_release_gpu(rd + ti, tk + 1);
#else
          pend_ptr = rd + ti;
          pend_val = tk + 1;
#endif
        }
        tm.mark(6);
      } else {
        // ---- diagonal tile: S = -acc (column-major, ld = PB_LD), residual row, sweep ----
        // (only the fragment rows that reach the lower triangle exist; the sweep never reads what lies above it: the
        // strict upper triangle is where it ASSIGNS the inverse, blocked.cuh)
        __syncthreads();   // every warp has left the ring: S aliases it
        {
          int uA, vA0, nA, uB, vB0, nB;
          dag_diag_segments(warp, uA, vA0, nA, uB, vB0, nB);
#pragma unroll
          for (int s2 = 0; s2 < 13; ++s2) {
            const bool segA = s2 < 9;
            if (segA ? (s2 >= nA) : (s2 - 9 >= nB)) continue;
            const int R0 = 16 * (segA ? uA : uB) + gq;
            const int C0 = 8 * (segA ? vA0 + s2 : vB0 + s2 - 9) + 2 * tig;
            const int slot = segA ? s2 : 12 + (s2 - 9);
#pragma unroll
            for (int v = 0; v < 4; ++v) S[R0 + 8 * (v >> 1) + (size_t)(C0 + (v & 1)) * PB_LD] = -acc[slot >> 2][slot & 3][v];
          }
        }
        if (tid < TILE) {
          S[TILE + (size_t)tid * PB_LD] = rpre0;
#pragma unroll
          for (int r = TILE + 1; r < PB_LD; ++r) S[r + (size_t)tid * PB_LD] = 0.0;
        }
        __syncthreads();
        tm.mark(7);
#if DAG_SWEEP_LOOKAHEAD == 2
        cta_block_sweep_cp<true, false, 17>(S, PB_LD, TILE, PA, PB_PA_LD, PBm, PB_PB_LD, Dm, wv, &tm);
#elif DAG_SWEEP_LOOKAHEAD
        cta_block_sweep_la(S, PB_LD, TILE, PA, PB_PA_LD, PBm, PB_PB_LD, Dm, wv, &tm);
#else
        cta_block_sweep<true, DAG_THREADS / 32>(S, PB_LD, TILE, PA, PB_PA_LD, PBm, PB_PB_LD, Dm, wv, &tm);
#endif

        double ld_part = 0.0, q_part = 0.0;
        int first_bad = 0x7fffffff, nan_seen = 0;
        double isd = 0.0;
        if (tid < TILE) {
          const double d = S[tid + (size_t)tid * PB_LD];
          if (!(d > 0.0)) {
            first_bad = tk * TILE + tid + 1;
            nan_seen = isnan(d) ? 1 : 0;
          }
          isd = 1.0 / sqrt(d);
          ld_part = log(d);
          const double z = S[TILE + (size_t)tid * PB_LD] * isd;
          q_part = z * z;
          g.zcols[((size_t)b * g.nt + tk) * TILE + tid] = z;
          if (g.zout != nullptr && tk * TILE + tid < g.N) g.zout[(size_t)b * g.ldz + (size_t)tk * TILE + tid] = z;
        }
        const double logdet = block_sum(ld_part, red);
        const double quad = block_sum(q_part, red);
        const int fb = -block_max_int(-first_bad, ired);
        const int anynan = block_max_int(nan_seen, ired);
        if (tid == 0) {
          g.acc_logdet[b] = ld_pre + logdet;
          g.acc_quad[b] = q_pre + quad;
          if (st_pre == 0 && fb != 0x7fffffff) g.status[b] = anynan ? -1 : fb;
        }
        if (tid < TILE) isd_s[tid] = isd;
        __syncthreads();
        tm.mark(11);
        for (int e2 = tid; e2 < TILE_ELEMS / 2; e2 += DAG_THREADS) {
          const int e = 2 * e2;
          const int kg = e >> 9, m8 = (e >> 5) & 15, gg = (e >> 2) & 7, k3 = e & 3;
          const int R = m8 * 8 + gg, C = kg * 4 + k3;
          const double s = isd_s[R];
          double2 o;
          o.x = (C < R) ? S[C + (size_t)R * PB_LD] * s : (C == R ? s : 0.0);
          o.y = (C + 1 < R) ? S[C + 1 + (size_t)R * PB_LD] * s : (C + 1 == R ? s : 0.0);
          reinterpret_cast<double2*>(Ct)[e2] = o;
        }
        __syncthreads();       // (readers of inv(L_kk) fence on their side, see the solve epilogue)
        if (tid == 0) {
          DAG_PRE_RELEASE_FENCE();
          st_release_gpu(rd + tk, tk + 1);
        }
        tm.mark(12);
      }
    }
    ++n_local;
    // both epilogues end in a CTA barrier with nothing but thread 0's release behind it, and nothing reads s_task / zs after
    // that barrier: it doubles as the end-of-task barrier.  Only the aborted path comes here without one.
    if (aborted) __syncthreads();
    tm.mark(0);
  }
#ifdef DAG_TIMING
  if (tid == 0)
    for (int i = 0; i < 16; ++i) atomicAdd(&g_dag_timing[i], (unsigned long long)tm.a[i]);
#endif
}
