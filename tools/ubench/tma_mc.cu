// Microbenchmark: chip-wide L2 -> shared-memory delivery rate of TMA when the CTAs of a cluster all need the SAME tile (the activation
// tile of N-neighbouring GEMM CTAs): every CTA fetches the whole tile itself (unicast) versus every CTA fetches 1/cluster of it and
// multicasts its slice to all CTAs of the cluster.  Tiles are 32 KB (256 rows x 128 B, 128B swizzle), the working set (64 MB) is L2
// resident, 4 tiles in flight per CTA.  Prints landed GB/s summed over all SMs.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t crank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
constexpr int kRows = 256, kTile = kRows * 128, kStages = 4;
template <int CS, int MC>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_slice, int rounds, int ntiles) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(sm + kStages * kTile);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) csync();
  const uint32_t r = CS > 1 ? crank() : 0, c = CS > 1 ? cid() : blockIdx.x;
  for (int it = 0; it < rounds; ++it) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStages; ++s) {
        const int tile = (c * 37 + it * kStages + s) % ntiles;
        mbar_expect(&bar[s], kTile);
        if (MC) {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                           s32(sm + s * kTile + r * (kTile / CS))), "l"((uint64_t)&tm_slice), "r"(s32(&bar[s])), "r"(0), "r"(tile * kRows + (int)r * (kRows / CS)), "h"((uint16_t)((1 << CS) - 1)) : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(sm + s * kTile)),
                       "l"((uint64_t)&tm_full), "r"(s32(&bar[s])), "r"(0), "r"(tile * kRows) : "memory");
        }
      }
      for (int s = 0; s < kStages; ++s) mbar_wait(&bar[s], it & 1);
    }
    __syncthreads();
    if (CS > 1 && MC) csync();  // a peer may not overwrite my stage before I have seen it complete
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn enc;
static CUtensorMap make(void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows}, strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}
template <int CS, int MC>
void run(void* buf, int ntiles) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms / CS * CS, smem = kStages * kTile + 1024 + 64, rounds = 400;
  CUtensorMap full = make(buf, (uint64_t)ntiles * kRows, kRows), slice = make(buf, (uint64_t)ntiles * kRows, kRows / CS);
  auto kern = k<CS, MC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, full, slice, 20, ntiles);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); cudaLaunchKernelEx(&cfg, kern, full, slice, rounds, ntiles); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double landed = (double)grid * rounds * kStages * kTile;
  printf("cluster %d %-9s: %7.3f ms, landed %8.1f GB/s chip-wide (%5.1f B/clk/SM at 1.9 GHz), L2 reads %8.1f GB/s  %s\n", CS, MC ? "multicast" : "unicast", ms,
         landed / ms / 1e6, landed / ms / 1e6 / grid / 1.9, landed / ms / 1e6 / (MC ? CS : 1), err == cudaSuccess ? "" : cudaGetErrorString(err));
}
int main() {
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int ntiles = 2048;  // 64 MB
  void* buf; cudaMalloc(&buf, (size_t)ntiles * kTile); cudaMemset(buf, 1, (size_t)ntiles * kTile);
  run<1, 0>(buf, ntiles);
  run<2, 0>(buf, ntiles); run<2, 1>(buf, ntiles);
  run<4, 0>(buf, ntiles); run<4, 1>(buf, ntiles);
  run<8, 0>(buf, ntiles); run<8, 1>(buf, ntiles);
  return 0;
}
