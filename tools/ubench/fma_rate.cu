// Microbenchmark: fp32 FMA issue rate on sm_100a — scalar FFMA vs packed FFMA2 (fma.rn.f32x2) with register operands, with the
// operand pattern of the depthwise kernel (8 accumulators share one multiplier operand -> .reuse) and with all-distinct operands.
// Prints FMA/clk/SM (clock from cudaDevAttrClockRate, and from a clock64() delta) so that the depthwise kernel's roofline uses a
// measured number.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* __restrict__ out, int iters, long long* cyc) {
  float a[8], w[8];
  u64 A[8], W[8];
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; w[i] = in[64 + threadIdx.x + i]; A[i] = (u64(__float_as_uint(a[i])) << 32) | __float_as_uint(w[i]); W[i] = A[i] ^ 0x1234; }
  float x0 = in[300], x1 = in[301];
  u64 X = (u64(__float_as_uint(x0)) << 32) | __float_as_uint(x1);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {        // scalar FFMA, shared multiplier (acc = acc * x0 + w)   16 FMA per thread per iteration
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], x0, w[i]);
    } else if (MODE == 1) { // scalar FFMA, all operands distinct registers
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], w[(i + 1) & 7], w[i]);
    } else if (MODE == 2) { // FFMA2, dwconv pattern: acc += v * wv (wv shared by 8)   16 FMA2 = 32 FMA per thread per iteration
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) A[i] = ffma2(W[i], X, A[i]);
    } else {                // FFMA2, all distinct
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) A[i] = ffma2(W[i], W[(i + 3) & 7], A[i]);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)A[i]) + __uint_as_float((unsigned)(A[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int fma_per_iter, int ctas_per_sm) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, sizeof(float) * sms * 8 * 256); cudaMalloc(&cyc, 8);
  const int iters = 20000, grid = sms * ctas_per_sm;
  k<MODE><<<grid, 256>>>(in, out, 100, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(in, out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double fma = double(grid) * 256 * iters * fma_per_iter;
  printf("%-34s ctas/SM %d  %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (clock64: %lld cyc -> %.3f GHz effective)  [%s]\n", name, ctas_per_sm, ms, 2 * fma / ms / 1e9,
         fma / double(c) / sms, c, double(c) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int c : {2, 4, 8}) {
    run<0>("FFMA  shared multiplier (.reuse)", 16, c);
    run<1>("FFMA  distinct operands", 16, c);
    run<2>("FFMA2 shared multiplier (.reuse)", 32, c);
    run<3>("FFMA2 distinct operands", 32, c);
  }
  return 0;
}
