// Dependent-chain latencies (cycles per op) for one warp alone on an SM.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void lat(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a + threadIdx.x; sm[threadIdx.x + 32] = b;
  __syncwarp();
  double x = a; long long t0, t1; int k = 0;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  // DMUL+DADD chain via division (1/x) chain
  t0 = clock64();
  for (int i = 0; i < N / 8; ++i) x = 1.0 / (x + 1.5);
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) * 8; k++;
  // DMMA dependent chain (accumulator dependency)
  double c[2] = {x, x};
  t0 = clock64();
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  x += c[0] + c[1];
  // LDS dependent chain (address depends on loaded value)
  int idx = threadIdx.x & 31;
  t0 = clock64();
  for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = ((int)v + i) & 31; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  x += idx;
  // syncwarp cost
  t0 = clock64();
  for (int i = 0; i < N; ++i) __syncwarp();
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  // STS -> syncwarp -> LDS round trip (hand-over through shared memory)
  t0 = clock64();
  for (int i = 0; i < N; ++i) { sm[(threadIdx.x + 1) & 31] = x; __syncwarp(); x += sm[threadIdx.x]; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  // shuffle chain (64-bit)
  t0 = clock64();
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  // fmax/fabs compare chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) x = fmax(fabs(x), b + i);
  t1 = clock64(); if (threadIdx.x == 0) cyc[k] = t1 - t0; k++;
  out[threadIdx.x] = x;
}
int main() {
  double *out; long long *cyc; cudaMalloc(&out, 256); cudaMallocManaged(&cyc, 64 * 8);
  lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
  lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
  const char *names[] = {"DFMA dependent", "1/x (fp64 divide) dependent", "DMMA dependent (accumulator)", "LDS dependent", "__syncwarp", "STS->sync->LDS->sync", "SHFL.64 dependent", "fmax(fabs) dependent"};
  for (int i = 0; i < 8; ++i) printf("%-32s %.1f cycles/op\n", names[i], (double)cyc[i] / N);
  return 0;
}
