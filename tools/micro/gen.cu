// Covariance-evaluation throughput (QuasiPer): sin vs sinpi, ILP 2/4/8, 8 or 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(const double* t, double* out, int n_iter, double a2, double c1, double c2, double c3) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  double ti = t[tid & 1023];
  for (int it = 0; it < n_iter; ++it) {
    double d[ILP], s[ILP], o[ILP];
#pragma unroll
    for (int e = 0; e < ILP; ++e) d[e] = fabs(ti - t[(it * ILP + e) & 1023]);
#pragma unroll
    for (int e = 0; e < ILP; ++e) s[e] = MODE == 0 ? sin(c2 * d[e]) : sinpi(c2 * d[e]);
#pragma unroll
    for (int e = 0; e < ILP; ++e) o[e] = a2 * exp(c1 * d[e] * d[e] + c3 * (s[e] * s[e]));
#pragma unroll
    for (int e = 0; e < ILP; ++e) acc += o[e];
  }
  out[tid] = acc;
}
template <int ILP, int MODE>
void run(const char* name, int threads, double* t, double* out) {
  const int blocks = 148, n_iter = 2048 / ILP;
  const double c2 = MODE == 0 ? 3.141592653589793 / 25.0 : 1.0 / 25.0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<ILP, MODE><<<blocks, threads>>>(t, out, n_iter, 25.0, -1.0 / 2500.0, c2, -4.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  const double evals = (double)blocks * threads * n_iter * ILP;
  printf("%-14s ILP %d, %2d warps/SM: %.3f ms, %.2f G evals/s, %.2f SM-clk per eval\n", name, ILP, threads / 32, best, evals / best / 1e6,
         best * 1e-3 * 1.965e9 * 148 / evals);
}
int main() {
  double *t, *out; cudaMalloc(&t, 1024 * 8); cudaMalloc(&out, 148 * 1024 * 8);
  double h[1024]; for (int i = 0; i < 1024; ++i) h[i] = i * 0.937 + (i % 7) * 0.01;
  cudaMemcpy(t, h, sizeof h, cudaMemcpyHostToDevice);
  run<2, 0>("sin", 256, t, out); run<4, 0>("sin", 256, t, out); run<8, 0>("sin", 256, t, out);
  run<2, 0>("sin", 512, t, out); run<4, 0>("sin", 512, t, out); run<8, 0>("sin", 512, t, out);
  run<4, 1>("sinpi", 256, t, out); run<8, 1>("sinpi", 256, t, out); run<4, 1>("sinpi", 512, t, out); run<8, 1>("sinpi", 512, t, out);
  run<4, 0>("sin", 1024, t, out); run<4, 1>("sinpi", 1024, t, out);
  return 0;
}
