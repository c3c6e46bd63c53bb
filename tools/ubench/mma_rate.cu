// Microbenchmark: legacy warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) and ldmatrix.x4 issue rates on sm_100a — the building
// blocks of a tensor-core depthwise convolution (Toeplitz blocks).  Prints MMAs per clock per SM and the equivalent dense TFLOP/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE>  // 0: MMA only (4 independent accumulators), 1: ldmatrix.x4 + MMA per iteration
__global__ void __launch_bounds__(256) k(const uint32_t* in, float* out, int iters) {
  __shared__ __align__(16) uint16_t sm[16 * 22 * 40];
  for (int i = threadIdx.x; i < 16 * 22 * 40; i += 256) sm[i] = (uint16_t)(in[i & 255] >> 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t a[4] = {in[lane], in[lane + 32], in[lane + 64], in[lane + 96]}, b[2] = {in[lane + 128], in[lane + 160]};
  float d[4][4] = {};
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + ((lane & 7) + ((lane >> 3) & 1) * 8) * 80 + (lane >> 4) * 16;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 1) {
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (u + (it & 3)) * 80));
      }
      mma16816(d[u], a, b);
    }
  }
  float s = 0;
  for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) s += d[u][j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ctas) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* in; float* out; cudaMalloc(&in, 4096); cudaMemset(in, 0x3c, 4096); cudaMalloc(&out, 4 * sms * 8 * 256);
  const int iters = 20000, grid = sms * ctas;
  k<MODE><<<grid, 256>>>(in, out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<grid, 256>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = double(grid) * 8 * iters * 4;
  printf("%-28s ctas/SM %d: %.3f ms  %.2f MMA(m16n8k16)/ns/GPU = %.1f dense TFLOP/s  (%.3f MMA per SM per 1.965 GHz clock)  [%s]\n", name, ctas, ms,
         mmas / ms / 1e6, mmas * 4096 / ms / 1e9, mmas / ms / 1e6 / sms / 1.965, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (int c : {1, 2, 4}) { run<0>("mma.sync only", c); run<1>("ldmatrix.x4 + mma.sync", c); }
  return 0;
}
