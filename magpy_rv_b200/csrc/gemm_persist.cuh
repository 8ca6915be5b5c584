// Persistent variant of the DMMA tile GEMM (blocked.cuh, gemm_tile_kernel<true>).
//
// One CTA per SM walks a static round-robin list of (walker, target tile) work items.  The operand
// ring (4 stages x 32 KiB, 1-D bulk copies + mbarriers) and its phase bookkeeping run CONTINUOUSLY
// across tiles: while a tile's last slices are being multiplied the elected producer lane is already
// streaming the first slices of the next tile, the first MMA fragments of the next tile are fetched
// before the epilogue, and the next tile's C block is prefetched into L2 a whole tile ahead.  What
// remains exposed between tiles is the epilogue's fire-and-forget stores plus an L2-hit load of C
// (about 1 us instead of the ~8 us launch/fill/drain bubble of the one-tile-per-CTA kernel).
#pragma once
#include "blocked.cuh"

struct TileWork {
  int b, ti, tj, nslices;
  double* Aw;
};

__device__ __forceinline__ TileWork decode_work(const GemmArgs& g, int w, int tiles_per_walker) {
  TileWork t;
  t.b = w / tiles_per_walker;
  int rem = w - t.b * tiles_per_walker;
  if (g.mode == GEMM_TRSM) {
    t.ti = g.k + 1 + rem;
    t.tj = g.k;
    t.nslices = SLICES_PER_TILE;
  } else {
    decode_update_tile(g, rem, t.ti, t.tj);
    t.nslices = (g.p_end - g.p_begin) * SLICES_PER_TILE;
  }
  t.Aw = g.Abase + (size_t)t.b * g.walker_stride;
  return t;
}

__device__ __forceinline__ const double* work_a_src(const GemmArgs& g, const TileWork& t, int q) {
  const int p = q >> 3, s = q & 7;
  const size_t tix = (g.mode == GEMM_TRSM) ? tile_index(t.ti, g.k) : tile_index(t.ti, g.p_begin + p);
  return t.Aw + tix * TILE_ELEMS + s * SLICE_ELEMS;
}
__device__ __forceinline__ const double* work_b_src(const GemmArgs& g, const TileWork& t, int q) {
  const int p = q >> 3, s = q & 7;
  const size_t tix = (g.mode == GEMM_TRSM) ? tile_index(g.k, g.k) : tile_index(t.tj, g.p_begin + p);
  return t.Aw + tix * TILE_ELEMS + s * SLICE_ELEMS;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_persist_kernel(GemmArgs g, int tiles_per_walker, int total_work, int* __restrict__ work_counter) {
  extern __shared__ double smem[];
  constexpr int NST = GEMM_TMA_STAGES;
  double* stages = smem;
  double* tis = smem + NST * GEMM_STAGE_DOUBLES;
  double* tjs = tis + TILE;
  double* part = tjs + TILE;
  double* zs = part + 4 * TILE;
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(zs + TILE);
  unsigned long long* empty_bar = full_bar + NST;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) & 1;
  const int wn = (g.mode == GEMM_TRSM && wm) ? 3 - (warp & 3) : (warp & 3);   // see gemm_tile_kernel
  const int q_last = (g.mode == GEMM_TRSM) ? 2 * wn + 1 : 0x7fffffff;
  const int gq = lane >> 2, tig = lane & 3;
  const int a_lane = ((wm * 8) << 5) + lane;
  const int b_lane = ((wn * 4) << 5) + lane;

  // Dynamic tile scheduler: the first tile of a CTA is blockIdx.x, every further one comes from a
  // global counter (static round-robin measured 5 % slower: SMs differ in L2 distance and the slowest
  // SM set the pace).  seq[] holds the CTA's tile sequence two tiles ahead of the consumer so the
  // producer lane can stream the next tile's operands before the current tile is finished.
  __shared__ int seq[4];
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async;\n" ::: "memory");
    seq[0] = blockIdx.x;
    seq[1] = ((int)blockIdx.x < total_work) ? atomicAdd(work_counter, 1) + (int)gridDim.x : total_work;
  }
  __syncthreads();

  // ---- producer cursor (used by tid 0 only): next slice to issue = (p_work, p_q), global index p_cnt
  int p_work = blockIdx.x, p_q = 0, p_cnt = 0, p_it = 0;
  TileWork p_tile;
  if (p_work < total_work) p_tile = decode_work(g, p_work, tiles_per_walker);
  auto produce_one = [&]() {
    if (p_work >= total_work) return;
    const int s = p_cnt % NST, u = p_cnt / NST;
    if (u > 0) mbar_wait(&empty_bar[s], (unsigned)((u - 1) & 1));
    double* dst = stages + s * GEMM_STAGE_DOUBLES;
    mbar_expect_tx(&full_bar[s], 2 * SLICE_ELEMS * sizeof(double));
    bulk_g2s(dst, work_a_src(g, p_tile, p_q), SLICE_ELEMS * sizeof(double), &full_bar[s]);
    bulk_g2s(dst + SLICE_ELEMS, work_b_src(g, p_tile, p_q), SLICE_ELEMS * sizeof(double), &full_bar[s]);
    ++p_cnt;
    if (++p_q == p_tile.nslices) {
      p_q = 0;
      ++p_it;
      p_work = seq[p_it & 3];
      if (p_work < total_work) p_tile = decode_work(g, p_work, tiles_per_walker);
    }
  };
  if (tid == 0) {
    produce_one();
    produce_one();
  }

  // ---- consumer state
  int c_cnt = 0;  // global index of the slice being consumed
  double af[2][8], bf[2][4];
  if (blockIdx.x < total_work) {
    mbar_wait(&full_bar[0], 0u);
    const double* As = stages + a_lane;
    const double* Bs = stages + SLICE_ELEMS + b_lane;
#pragma unroll
    for (int v = 0; v < 8; ++v) af[0][v] = As[v << 5];
#pragma unroll
    for (int v = 0; v < 4; ++v) bf[0][v] = Bs[v << 5];
  }

  int it = 0;
  for (int w = blockIdx.x; w < total_work; ++it) {
    const TileWork t = decode_work(g, w, tiles_per_walker);
    double* Ct = t.Aw + tile_index(t.ti, t.tj) * TILE_ELEMS;

    consumer_sync();  // previous tile's epilogue is done with tis / tjs / zs / part; seq[it+1] is visible
    const int w_next = seq[(it + 1) & 3];
    const bool has_next = w_next < total_work;
    if (tid == 0) seq[(it + 2) & 3] = has_next ? atomicAdd(work_counter, 1) + (int)gridDim.x : total_work;
    if (g.mode == GEMM_UPDATE && g.gen) {
      if (tid < TILE) {
        const int gi = t.ti * TILE + tid;
        tis[tid] = gi < g.N ? g.t[gi] : 0.0;
      } else {
        const int gj = t.tj * TILE + (tid - TILE);
        tjs[tid - TILE] = gj < g.N ? g.t[gj] : 0.0;
      }
    }
    if (g.mode == GEMM_TRSM && tid < TILE) zs[tid] = g.zbuf[(size_t)t.b * TILE + tid];
    if (g.mode == GEMM_UPDATE && !g.gen && has_next && (g.cs_store & 2)) {
      // pull the NEXT tile's C block into L2 while this tile computes (1024 lines of 128 B)
      const TileWork nx = decode_work(g, w_next, tiles_per_walker);
      const char* cn = reinterpret_cast<const char*>(nx.Aw + tile_index(nx.ti, nx.tj) * TILE_ELEMS);
#pragma unroll
      for (int u = 0; u < 4; ++u) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(cn + (size_t)(u * 256 + tid) * 128));
    }

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
      const CovParams cp = make_cov_params(g.kernel_id, g.variant, g.hp_wave + (size_t)t.b * g.nh);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          {
            const int R = wm * 64 + mi * 16 + gq;
            const int C = wn * 32 + ni * 8 + 2 * tig;
            double frag[4];
            gen_fragment4(cp, g.N, t.ti * TILE + R, t.tj * TILE + C, tis[R], tis[R + 8], tjs[C], tjs[C + 1], g.yerr, frag);
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[mi][ni][v] = -frag[v];
          }
    } else {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[mi][ni][v] = 0.0;
    }

    for (int q = 0; q < t.nslices; ++q, ++c_cnt) {
      const int st = c_cnt % NST;
      if (tid == 0) produce_one();   // keeps the ring two slices ahead, across tile boundaries
      __syncwarp();
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
        } else if (q + 1 < t.nslices || has_next) {
          // first fragments of the next slice -- possibly the first slice of the next tile
          const int sn = (c_cnt + 1) % NST;
          mbar_wait(&full_bar[sn], (unsigned)(((c_cnt + 1) / NST) & 1));
          const double* An = stages + sn * GEMM_STAGE_DOUBLES + a_lane;
          const double* Bn = stages + sn * GEMM_STAGE_DOUBLES + SLICE_ELEMS + b_lane;
#pragma unroll
          for (int v = 0; v < 8; ++v) af[nxt][v] = An[v << 5];
#pragma unroll
          for (int v = 0; v < 4; ++v) bf[nxt][v] = Bn[v << 5];
        }
        if (q <= q_last) {   // warp-uniform (TRSM skips the zero slices of inv(L))
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma_1684(acc[mi][ni], af[cur][2 * mi], af[cur][2 * mi + 1], bf[cur][ni]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }

    // epilogue
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
            *reinterpret_cast<double2*>(Ct + tile_off(R, C)) = o;
          }
    } else {
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
        double* ri = g.resid + (size_t)t.b * g.ldr + (size_t)t.ti * TILE;
        ri[tid] -= s;
      }
    }
    w = w_next;
  }
}
