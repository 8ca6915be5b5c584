// Single-warp 16x16 Cholesky + inverse in shared memory: where do ~1300 clk per pivot go?
#include <cstdio>
#include <cuda_runtime.h>
#define LD 36
__global__ void k(double* out, long long* t, int variant) {
  __shared__ double Dm[32 * 17];
  __shared__ double invL[64 * LD];
  __shared__ double rdiag[32];
  const int lane = threadIdx.x & 31;
  const int rr = lane & 15, g = lane >> 4;
  for (int rep = 0; rep < 3; ++rep) {
    // SPD test matrix: A = I*20 + 1/(1+|i-j|)
    for (int c = 0; c < 32; ++c)
      if (lane < 16) Dm[c * 17 + rr] = (c < 16) ? ((c == rr ? 20.0 : 0.0) + 1.0 / (1.0 + abs(c - rr))) : 0.0;
    __syncwarp();
    long long t0 = clock64();
    double* colk = Dm;
#pragma unroll 1
    for (int k = 0; k < 16; ++k, colk += 17) {
      const double dk = colk[k];
      const double rs = rsqrt(dk);
      const double lik = colk[rr] * rs;
      if (lane == k) rdiag[k] = rs;
      if (variant != 1) __syncwarp();
      if (lane >= k && lane < 16) colk[rr] = lik;
      if (variant != 1) __syncwarp();
      const int cb = k + 1 + 8 * g;
      const double* pl = colk + cb;
      double* pa = Dm + cb * 17 + rr;
      const int nst = min(16 - cb, rr - cb + 1);
      double lv[8], av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { lv[u] = pl[u]; av[u] = pa[u * 17]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) av[u] = fma(-lik, lv[u], av[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (u < nst) pa[u * 17] = av[u];
      if (variant != 1) __syncwarp();
    }
    long long t1 = clock64();
    double* X = invL;
#pragma unroll
    for (int u = 0; u < 8; ++u) X[(8 * g + u) * LD + rr] = 0.0;
    __syncwarp();
    const double* colq = Dm;
#pragma unroll 1
    for (int q = 0; q < 16; ++q, colq += 17) {
      const double xq = (q >= rr) ? ((q == rr ? 1.0 : 0.0) + X[q * LD + rr]) * rdiag[q] : 0.0;
      __syncwarp();
      if (g == 0) X[q * LD + rr] = xq;
      const int ib = q + 1 + 8 * g;
      const double* pl = colq + ib;
      double* px = X + ib * LD + rr;
      const int nst = 16 - ib;
      double lv[8], xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { lv[u] = pl[u]; xv[u] = px[u * LD]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = fma(-lv[u], xq, xv[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (u < nst) px[u * LD] = xv[u];
      __syncwarp();
    }
    long long t2 = clock64();
    if (lane == 0) { t[0] = t1 - t0; t[1] = t2 - t1; }
  }
  out[lane] = Dm[lane] + invL[lane];
}
int main() {
  double* out; long long* t;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&t, 16 * 8);
  for (int variant = 0; variant < 2; ++variant) {
    k<<<1, 32>>>(out, t, variant);
    long long h[2];
    cudaMemcpy(h, t, sizeof h, cudaMemcpyDeviceToHost);
    double o[32];
    cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
    printf("variant %d: potrf16 %lld clk (%.0f per pivot), inverse16 %lld clk (%.0f per step); L00=%.6f\n", variant, h[0], h[0] / 16.0, h[1], h[1] / 16.0, o[0]);
  }
  return 0;
}
