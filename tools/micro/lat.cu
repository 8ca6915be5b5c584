// Latency microbenchmarks (single warp, dependent chains): nvcc -arch=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double seed) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x;
  double x = seed + lane * 1e-3, y = 1.0000001;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (lane == 0) t[0] = t1 - t0;
  // 2. dependent rsqrt chain
  double r = fabs(x) + 1.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) r = rsqrt(r) + 1.5;
  t1 = clock64();
  if (lane == 0) t[1] = t1 - t0;
  // 3. sqrt + div chain
  double q = r + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) q = 1.0 / sqrt(q) + 1.5;
  t1 = clock64();
  if (lane == 0) t[2] = t1 - t0;
  // 4. smem store -> load chain (same thread)
  sm[lane] = q;
  __syncwarp();
  double v = q;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    sm[lane + 32 * (i & 7)] = v;
    v = sm[lane + 32 * (i & 7)] + 1.0;
  }
  t1 = clock64();
  if (lane == 0) t[3] = t1 - t0;
  // 5. smem store -> syncwarp -> load from another lane
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    sm[lane] = v;
    __syncwarp();
    v = sm[(lane + 1) & 31] + 1.0;
    __syncwarp();
  }
  t1 = clock64();
  if (lane == 0) t[4] = t1 - t0;
  // 6. shuffle chain (double)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31) + 1.0;
  t1 = clock64();
  if (lane == 0) t[5] = t1 - t0;
  // 7. dependent DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) v = v + y;
  t1 = clock64();
  if (lane == 0) t[6] = t1 - t0;
  // 8. dependent FFMA chain (fp32) for reference
  float f = (float)v;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
  t1 = clock64();
  if (lane == 0) t[7] = t1 - t0;
  // 9. independent DFMA x8 (throughput per warp)
  double a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = v + u;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = fma(a[u], y, 1e-9);
  t1 = clock64();
  if (lane == 0) t[8] = t1 - t0;
  double s = 0;
  for (int u = 0; u < 8; ++u) s += a[u];
  out[lane] = x + r + q + v + f + s;
}
int main() {
  double* out; long long* t;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&t, 16 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, t, 1.0);
  long long h[16];
  cudaMemcpy(h, t, sizeof h, cudaMemcpyDeviceToHost);
  printf("DFMA dep: %.1f clk/op\nrsqrt(+DADD) dep: %.1f clk\n1/sqrt(+DADD) dep: %.1f clk\nSTS->LDS same thread (+DADD): %.1f clk\nSTS,syncwarp,LDS other lane,syncwarp (+DADD): %.1f clk\nSHFL f64 (+DADD): %.1f clk\nDADD dep: %.1f clk/op\nFFMA dep: %.1f clk/op\nDFMA x8 independent: %.1f clk per 8 (%.1f per op)\n",
         h[0] / 256.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 256.0, h[7] / 256.0, h[8] / 64.0, h[8] / 512.0);
  return 0;
}
