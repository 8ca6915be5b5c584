// Persistent dataflow ("tile DAG") form of the tiled factorisation for 512 < N <= 4096, batched over a wave.
//
// The launch-per-tile-column host loop of abi.cu:run_wave pays, per column, a serial chain
//     in-panel UPDATE -> diagonal-tile kernel (latency-bound, one CTA per walker) -> TRSM
// whose launches each end in a partially filled last round of CTAs; at N = 1024 x 128 walkers the tensor pipe idles
// more than half of the time (round 1: 44 % of the DMMA peak).  Here ONE kernel of one CTA per SM runs the whole
// factorisation of a wave as a list of TILE TASKS pulled from an atomic counter; the dependencies between tasks
// are per-walker, per-tile-row progress counters in global memory (release / acquire), so different walkers sit in
// different tile columns at the same time and nothing waits for the slowest CTA of a launch.
//
// Left-looking by tiles, same storage (lower-triangular 128 x 128 tiles in operand order, blocked.cuh):
//   O(b, i, k), i > k   acc = -K(i,k) generated from t (first and only touch) + sum_{p<k} L(i,p) L(k,p)^T  (DMMA, operands
//                       streamed through the bulk-copy / mbarrier ring; the k tiles of a row are contiguous in memory);
//                       C = -acc goes to the workspace; once the diagonal tile of column k is done: X = C inv(L_kk)^T
//                       (DMMA, triangular skip as in TRSM), L(i,k) = X stored, r_i -= X z_k.
//   D(b, k)             acc = -K(k,k) + sum_{p<k} L(k,p) L(k,p)^T; the 128 x 128 block goes to shared memory and is swept
//                       by cta_block_sweep (pivots, inv(L_kk), z_k = L_kk^-1 r_k, log det, status) exactly as
//                       potrf_tile_blk_kernel does.
// Row i of walker b has `rowdone[b][i]` = number of finished tiles (i,0..); O(b,i,k) needs rowdone[b][i] >= k and
// rowdone[b][k] >= k for its update and rowdone[b][k] >= k + 1 (diagonal tile done) for its solve.
//
// Task order (any order that keeps a walker's dependencies earlier in the list is deadlock-free: the grid is
// co-resident, tasks are taken in list order, so everything a task waits for is already running or done):
//     D(.,0) | O(.,1,0)  D(.,1)  O(.,i>=2,0) | O(.,2,1)  D(.,2)  O(.,i>=3,1) | ...
// i.e. the tile that the NEXT diagonal task needs is finished first and the next diagonal tasks overlap with the
// rest of the column (look-ahead).  Summation order per element is the order of the launch-per-column path
// (p ascending, k16 slices ascending; storing C = -acc and reloading acc = -C is exact), so results are bitwise
// identical to it.
#pragma once
#include "blocked.cuh"

struct DagArgs {
  int N, nt, B;
  int kernel_id, variant, nh;
  int total_tasks;
  double* Abase;
  size_t walker_stride;
  const double* t;
  const double* yerr;
  const double* hp_wave;   // [B, nh]
  double* resid;           // [B, ldr]
  int ldr;
  double* zcols;           // [B, nt, 128]  z_k per tile column
  int* rowdone;            // [B, nt]
  int* counter;            // [0] task counter, [1] abort flag
  double* acc_logdet;      // [B]
  double* acc_quad;
  int* status;
  double* zout;            // nullable [B, ldz]
  long long ldz;
};

#define DAG_THREADS 256
#define DAG_AUTO_MAX_TILES 0          // automatic choice: tile columns up to which the dataflow kernel replaces run_wave
#define DAG_SWEEP_DOUBLES (PB_LD * TILE + PB_NB * PB_PA_LD + PB_NB * PB_PB_LD + PB_NB * PB_D_LD + PB_NB + 64 + TILE)
#define DAG_AUX_DOUBLES (2 * TILE + 4 * TILE + TILE)
#define DAG_SMEM_BYTES ((size_t)(DAG_SWEEP_DOUBLES + DAG_AUX_DOUBLES) * sizeof(double) + 2 * GEMM_TMA_STAGES * sizeof(unsigned long long) + 16 * sizeof(int))
#define DAG_SPIN_LIMIT (1ll << 33)   // ~4 s of SM clocks: a dependency that takes longer means a lost task, not a slow one

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
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
  if (r < B) {                 // the tile the next diagonal task needs
    type = DAG_TASK_OFF;
    b = r;
    i = s + 1;
  } else if (r < 2 * B) {      // next diagonal tile (look-ahead)
    type = DAG_TASK_DIAG;
    b = r - B;
    i = k = s + 1;
  } else {                     // rest of column s, walker-major so that co-running CTAs share L(s, .) in L2
    r -= 2 * B;
    const int per = nt - s - 2;
    type = DAG_TASK_OFF;
    b = r / per;
    i = s + 2 + r % per;
  }
}

__global__ void __launch_bounds__(DAG_THREADS, 1) tile_dag_kernel(DagArgs g) {
  extern __shared__ double smem[];
  constexpr int NST = GEMM_TMA_STAGES;
  double* stages = smem;                          // operand ring, 4 x 32 KiB -- aliased by the sweep matrix S
  double* S = smem;
  double* PA = S + (size_t)PB_LD * TILE;
  double* PBm = PA + PB_NB * PB_PA_LD;
  double* Dm = PBm + PB_NB * PB_PB_LD;
  double* wv = Dm + PB_NB * PB_D_LD;
  double* red = wv + PB_NB;
  int* ired = (int*)(red + 32);
  double* isd_s = red + 64;
  double* tis = smem + DAG_SWEEP_DOUBLES;         // never aliased
  double* tjs = tis + TILE;
  double* part = tjs + TILE;
  double* zs = part + 4 * TILE;
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(zs + TILE);
  unsigned long long* empty_bar = full_bar + NST;
  int* s_task = reinterpret_cast<int*>(empty_bar + NST);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) & 1;
  const int gq = lane >> 2, tig = lane & 3;
  int* abort_flag = g.counter + 1;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  unsigned qg = 0;   // slices streamed through the ring so far (uniform across the CTA)

  for (;;) {
    // ---- next task ---------------------------------------------------------------------------
    if (tid == 0) {
      int type = DAG_TASK_EXIT, b = 0, i = 0, k = 0;
      const int tsk = atomicAdd(g.counter, 1);
      if (tsk < g.total_tasks && ld_acquire_gpu(abort_flag) == 0) {
        dag_decode(g, tsk, type, b, i, k);
        // update operands: rows i and k finished through column k - 1 (also orders the residual block r_i / r_k)
        const int* rd = g.rowdone + (size_t)b * g.nt;
        if (k > 0 && !(dag_wait(rd + i, k, abort_flag) && dag_wait(rd + k, k, abort_flag))) type = DAG_TASK_EXIT;
        __threadfence();
        fence_proxy_async();
      }
      s_task[0] = type;
      s_task[1] = b;
      s_task[2] = i;
      s_task[3] = k;
    }
    __syncthreads();
    const int type = s_task[0], b = s_task[1], ti = s_task[2], tk = s_task[3];
    if (type == DAG_TASK_EXIT) break;

    double* Aw = g.Abase + (size_t)b * g.walker_stride;
    const double* rowA = Aw + tile_index(ti, 0) * TILE_ELEMS;   // tiles (ti, 0..): contiguous
    const double* rowB = Aw + tile_index(tk, 0) * TILE_ELEMS;
    double* Ct = Aw + tile_index(ti, tk) * TILE_ELEMS;
    int* rd = g.rowdone + (size_t)b * g.nt;

    // t of the tile rows / columns for the generation
    if (tid < TILE) {
      const int gi = ti * TILE + tid;
      tis[tid] = gi < g.N ? g.t[gi] : 0.0;
    } else {
      const int gj = tk * TILE + (tid - TILE);
      tjs[tid - TILE] = gj < g.N ? g.t[gj] : 0.0;
    }

    double acc[4][4][4];
    const int nphase = (type == DAG_TASK_OFF) ? 2 : 1;
    for (int phase = 0; phase < nphase; ++phase) {
      // phase 0: update (K = 128 tk); phase 1: triangular solve against inv(L_kk) (K = 128, lower-triangular B)
      const int wn = (phase == 1 && wm) ? 3 - (warp & 3) : (warp & 3);
      const int q_last = (phase == 1) ? 2 * wn + 1 : 0x7fffffff;
      const int nslices = (phase == 1) ? SLICES_PER_TILE : tk * SLICES_PER_TILE;
      const double* Asrc = (phase == 1) ? Ct : rowA;
      const double* Bsrc = (phase == 1) ? rowB + (size_t)tk * TILE_ELEMS : rowB;

      if (phase == 1 && tid == 0) {
        // the diagonal tile of column tk (inv(L_kk), z_k); C was written through the generic proxy above
        if (!dag_wait(rd + tk, tk + 1, abort_flag)) s_task[0] = DAG_TASK_EXIT;
        __threadfence();
        fence_proxy_async();
      }
      if (phase == 1) {
        __syncthreads();
        if (s_task[0] == DAG_TASK_EXIT) break;
        // values other CTAs produced are read past L1 (ld.global.cg): an SM's L1 may hold a stale line of them
        if (tid < TILE) zs[tid] = __ldcg(g.zcols + ((size_t)b * g.nt + tk) * TILE + tid);
      }
      if (tid == 0) {
        for (int q = 0; q < 2 && q < nslices; ++q) {
          const unsigned qq = qg + q, sn = qq % NST, u = qq / NST;
          if (u > 0) mbar_wait(&empty_bar[sn], (u - 1) & 1);
          double* dst = stages + sn * GEMM_STAGE_DOUBLES;
          mbar_expect_tx(&full_bar[sn], 2 * SLICE_ELEMS * sizeof(double));
          bulk_g2s(dst, Asrc + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
          bulk_g2s(dst + SLICE_ELEMS, Bsrc + (size_t)q * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
        }
      }

      if (phase == 0) {
        consumer_sync();  // tis / tjs
        const CovParams cp = make_cov_params(g.kernel_id, g.variant, g.hp_wave + (size_t)b * g.nh);
        const bool interior = ti > tk && (ti + 1) * TILE <= g.N;
        if (interior) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int np2 = 0; np2 < 2; ++np2) {
              const int R = wm * 64 + mi * 16 + gq;
              const int C = wn * 32 + np2 * 16 + 2 * tig;
              const Cov8 f = gen_interior8(cp.kind, cp.a2, cp.c1, cp.c2, cp.c3, tis[R], tis[R + 8], tjs[C], tjs[C + 1],
                                           tjs[C + 8], tjs[C + 9]);
              acc[mi][2 * np2][0] = -f.v0; acc[mi][2 * np2][1] = -f.v1; acc[mi][2 * np2][2] = -f.v2; acc[mi][2 * np2][3] = -f.v3;
              acc[mi][2 * np2 + 1][0] = -f.v4; acc[mi][2 * np2 + 1][1] = -f.v5; acc[mi][2 * np2 + 1][2] = -f.v6;
              acc[mi][2 * np2 + 1][3] = -f.v7;
            }
        } else {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int np2 = 0; np2 < 2; ++np2) {
              const int R = wm * 64 + mi * 16 + gq;
              const int C = wn * 32 + np2 * 16 + 2 * tig;
              double frag[8];
              gen_fragment8(cp, g.N, ti * TILE + R, tk * TILE + C, tis[R], tis[R + 8], tjs[C], tjs[C + 1], tjs[C + 8],
                            tjs[C + 9], g.yerr, frag);
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                acc[mi][2 * np2][v] = -frag[v];
                acc[mi][2 * np2 + 1][v] = -frag[4 + v];
              }
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

      // ---- main loop (as gemm_tile_kernel; the ring and its barrier phases persist across tasks) ----
      if (nslices > 0) {
        const int a_lane = ((wm * 8) << 5) + lane;
        const int b_lane = ((wn * 4) << 5) + lane;
        double af[2][8], bf[2][4];
        {
          const unsigned st0 = qg % NST;
          mbar_wait(&full_bar[st0], (qg / NST) & 1);
          const double* As = stages + st0 * GEMM_STAGE_DOUBLES + a_lane;
          const double* Bs = stages + st0 * GEMM_STAGE_DOUBLES + SLICE_ELEMS + b_lane;
#pragma unroll
          for (int v = 0; v < 8; ++v) af[0][v] = As[v << 5];
#pragma unroll
          for (int v = 0; v < 4; ++v) bf[0][v] = Bs[v << 5];
        }
        for (int q = 0; q < nslices; ++q) {
          const unsigned qa = qg + q, st = qa % NST;
          {
            const int qn = q + 2;
            if (tid == 0 && qn < nslices) {
              const unsigned qq = qg + qn, sn = qq % NST, u = qq / NST;
              if (u > 0) mbar_wait(&empty_bar[sn], (u - 1) & 1);
              double* dst = stages + sn * GEMM_STAGE_DOUBLES;
              mbar_expect_tx(&full_bar[sn], 2 * SLICE_ELEMS * sizeof(double));
              bulk_g2s(dst, Asrc + (size_t)qn * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
              bulk_g2s(dst + SLICE_ELEMS, Bsrc + (size_t)qn * SLICE_ELEMS, SLICE_ELEMS * sizeof(double), &full_bar[sn]);
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
              const unsigned qb = qa + 1, sn = qb % NST;
              mbar_wait(&full_bar[sn], (qb / NST) & 1);
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
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
        qg += nslices;
      }

      // ---- epilogues ---------------------------------------------------------------------------
      if (type == DAG_TASK_OFF && phase == 0) {
        // C = -acc to the workspace (operand order): it is the A operand of the solve phase
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
              *reinterpret_cast<double2*>(Ct + tile_off(R, C)) = o;
            }
        fence_proxy_async();   // generic writes before this CTA's bulk copies read them back (barrier follows in phase 1)
      } else if (type == DAG_TASK_OFF) {
        // X = acc is L(i,k): store it and fold r_i -= X z_k
        __syncthreads();   // zs
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
        __syncthreads();
        if (tid < TILE) {
          const double s = ((part[tid] + part[TILE + tid]) + part[2 * TILE + tid]) + part[3 * TILE + tid];
          double* ri = g.resid + (size_t)b * g.ldr + (size_t)ti * TILE;
          ri[tid] = __ldcg(ri + tid) - s;
        }
        fence_proxy_async();
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(rd + ti, tk + 1);
      } else {
        // ---- diagonal tile: S = -acc (column-major, ld = PB_LD), residual row, sweep ----
        __syncthreads();   // every warp has left the ring: S aliases it
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int R = wm * 64 + mi * 16 + gq + 8 * (v >> 1);
              const int C = wn * 32 + ni * 8 + 2 * tig + (v & 1);
              S[R + (size_t)C * PB_LD] = -acc[mi][ni][v];
            }
        if (tid < TILE) {
          S[TILE + (size_t)tid * PB_LD] = __ldcg(g.resid + (size_t)b * g.ldr + (size_t)tk * TILE + tid);
#pragma unroll
          for (int r = TILE + 1; r < PB_LD; ++r) S[r + (size_t)tid * PB_LD] = 0.0;
        }
        __syncthreads();
        cta_block_sweep<true, DAG_THREADS / 32>(S, PB_LD, TILE, PA, PB_PA_LD, PBm, PB_PB_LD, Dm, wv);

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
          g.acc_logdet[b] = __ldcg(g.acc_logdet + b) + logdet;
          g.acc_quad[b] = __ldcg(g.acc_quad + b) + quad;
          if (__ldcg(g.status + b) == 0 && fb != 0x7fffffff) g.status[b] = anynan ? -1 : fb;
        }
        if (tid < TILE) isd_s[tid] = isd;
        __syncthreads();
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
        fence_proxy_async();   // S (generic) before the ring is refilled; inv(L_kk) before other CTAs' bulk copies
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(rd + tk, tk + 1);
      }
    }
    __syncthreads();   // s_task / tis / tjs / zs / part are rewritten by the next task
  }
}
