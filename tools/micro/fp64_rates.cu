// Micro-benchmark: FP64 issue rates on this GPU -- DFMA, DMMA (mma.sync m8n8k4 f64),
// broadcast LDS.64 / LDS.128, SHFL -- in warp-instructions per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  for (int i = 0; i < ITERS; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma(double *out, double a, double b) {
  double c0[2] = {a, a}, c1[2] = {a, a}, c2[2] = {a, a}, c3[2] = {a, a};
  double fa = a + threadIdx.x, fb = b;
  for (int i = 0; i < ITERS; ++i) {
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" \
                            : "+d"(c[0]), "+d"(c[1]) : "d"(fa), "d"(fb));
    MMA(c0) MMA(c1) MMA(c2) MMA(c3)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

template <int W> __global__ void k_lds(double *out, int off) {
  __shared__ __align__(16) double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  int base = off; // same address for every lane: broadcast
  for (int i = 0; i < ITERS; ++i) {
    if (W == 8) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)(__cvta_generic_to_shared(sm + ((base + i) & 511)))));
      acc += v;
    } else {
      double v0, v1;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1)
                   : "r"((unsigned)(__cvta_generic_to_shared(sm + (((base + i) & 255) * 2)))));
      acc += v0 + v1;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_shfl(double *out, double a) {
  double x = a + threadIdx.x;
  for (int i = 0; i < ITERS; ++i) x += __shfl_sync(0xffffffffu, x, i & 31);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double *out; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  const int warps_per_sm = 32, blocks = sms * 4, threads = 256; // 8 warps x 4 CTAs
  auto rate = [&](float ms, double warp_instr_per_thread_iter) {
    double wi = (double)blocks * (threads / 32) * ITERS * warp_instr_per_thread_iter;
    double cyc = ms * 1e-3 * khz * 1e3;
    return wi / cyc / sms;
  };
  float t;
  t = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0, 0.999); });
  printf("DFMA   : %.3f warp-instr/clk/SM  (%.1f FMA/clk/SM)  %.3f ms\n", rate(t, 8), rate(t, 8) * 32, t);
  t = timeit([&] { k_dmma<<<blocks, threads>>>(out, 1.0, 0.999); });
  printf("DMMA884: %.3f warp-instr/clk/SM  (%.1f FMA/clk/SM)  %.3f ms\n", rate(t, 4), rate(t, 4) * 256, t);
  t = timeit([&] { k_lds<8><<<blocks, threads>>>(out, 3); });
  printf("LDS.64  broadcast: %.3f warp-instr/clk/SM (%.1f B/clk/SM delivered per lane-set)  %.3f ms\n", rate(t, 1), rate(t, 1) * 8, t);
  t = timeit([&] { k_lds<16><<<blocks, threads>>>(out, 3); });
  printf("LDS.128 broadcast: %.3f warp-instr/clk/SM  %.3f ms\n", rate(t, 1), t);
  t = timeit([&] { k_shfl<<<blocks, threads>>>(out, 1.0); });
  printf("SHFL(64-bit = 2x32): %.3f shfl64/clk/SM  %.3f ms\n", rate(t, 1), t);
  printf("SMs %d clock %.0f MHz (nominal; actual may differ)\n", sms, khz / 1e3);
  (void)warps_per_sm;
  return 0;
}
