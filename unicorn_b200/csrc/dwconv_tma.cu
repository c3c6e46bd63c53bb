// Depthwise 7x7 convolution (pad 3) + bias of the ConvNeXt block (unicorn/models/backbone/convnext.py:43, :21), TMA staged.
//
//   y[b, oh, ow, c] = bias[c] + sum_{kh,kw} x[b, oh+kh-3, ow+kw-3, c] * w[kh*7+kw][c]          x, y NHWC bf16, w fp32 [49][C]
//
// Work item = a 16 x 4 output tile of one 64-channel chunk.  ONE elected thread asks the TMA engine for the item's
// (16+6) x (4+6) x 64-channel input box — a 4-D box {64 ch, 22, 10, 1} of the NHWC map whose out-of-map part (the zero padding
// of the convolution, negative coordinates included) is zero-filled by the hardware — and, when the chunk changed, for the
// chunk's 49 x 64 fp32 filter taps; both land in shared memory and complete on one mbarrier.  No thread computes an address or
// a bounds check for the staging.  256 threads = 8 warps; warp = one 8-pixel strip (half a tile row), lane = one channel pair:
// 14 staged inputs and 8 packed (channel pair) accumulators live in registers, one packed FFMA2 per tap and pixel.
//
// The kernel is bound by fp32 FMA issue (98 flop per output element, 15.6 GFLOP per 800x1280 frame), not by HBM: the grid is
// persistent with four CTAs per SM (40.8 KB of shared memory each) so that the loads of one CTA overlap the arithmetic of the other
// three, and the items are 64 pixels x 64 channels so that 148 x 4 CTAs stay balanced on maps as small as 50 x 80 (780 items).
//
// Optional per-pixel LayerNorm statistics (sum, sum of squares over C of the STORED bf16 values, int64 fixed point 2^22, integer
// atomics: order independent) feed the following pwconv1, which applies the normalisation in its epilogue (UcConv2d.row_stats).
#include "uc_ptx.cuh"
#include "uc_common.h"
#include "../../include/unicorn_b200.h"
#include <algorithm>
#include <stdlib.h>

namespace uc {

constexpr int kDwTW = 16, kDwTH = 4, kDwCCH = 64, kDwPX = 8;
constexpr int kDwHW = kDwTW + 6, kDwHH = kDwTH + 6;               // 22 x 10 input box
constexpr int kDwPixBytes = kDwCCH * 2;                             // 128 B per staged pixel
constexpr int kDwTileBytes = kDwHH * kDwHW * kDwPixBytes;           // 28160
constexpr int kDwWBytes = 49 * kDwCCH * 4;                          // 12544
constexpr int kDwThreads = (kDwTW / kDwPX) * kDwTH * 32;            // 256
constexpr int kDwSmem = kDwTileBytes + kDwWBytes + 128 + 128;       // + barrier + alignment slack
constexpr int kDwCtasPerSm = 4;

struct alignas(64) DwParams {
  CUtensorMap tmX, tmW;
  const float* bias;
  uint16_t* y;
  unsigned long long* ln_stats;
  int H, W, C, B, tiles_w, tiles_h, n_items;
};

__global__ void __launch_bounds__(kDwThreads, kDwCtasPerSm) dwconv7_tma_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* tile = dsm_raw + ((128u - (smem_u32(dsm_raw) & 127u)) & 127u);  // 128-byte aligned; [10][22][64] bf16 (pointer stays a shared-memory pointer: LDS, not LD)
  float* sw = reinterpret_cast<float*>(tile + kDwTileBytes);                                                   // [49][64] fp32
  uint64_t* bar = reinterpret_cast<uint64_t*>(tile + kDwTileBytes + kDwWBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hx = warp & 1, r = warp >> 1;  // this warp's strip: pixels 8 hx .. 8 hx + 7 of tile row r
  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmX);
    prefetch_tmap(&p.tmW);
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();
  // contiguous item range per CTA (chunk slowest): neighbouring tiles share their halo in L2 and mostly the filter chunk
  const long n = p.n_items;
  const int i0 = static_cast<int>(n * blockIdx.x / gridDim.x), i1 = static_cast<int>(n * (blockIdx.x + 1) / gridDim.x);
  const int tiles_img = p.tiles_w * p.tiles_h, tiles_chunk = tiles_img * p.B;
  const int C2 = p.C >> 1;
  uint32_t phase = 0;
  int last_chunk = -1;
  for (int item = i0; item < i1; ++item) {
    const int chunk = item / tiles_chunk, t = item - chunk * tiles_chunk;
    const int b = t / tiles_img, tt = t - b * tiles_img;
    const int oh0 = (tt / p.tiles_w) * kDwTH, ow0 = (tt % p.tiles_w) * kDwTW;
    const int c0 = chunk * kDwCCH;
    if (threadIdx.x == 0) {
      const bool load_w = chunk != last_chunk;
      mbar_arrive_expect_tx(bar, kDwTileBytes + (load_w ? kDwWBytes : 0));
      tma_load_4d(tile, &p.tmX, bar, c0, ow0 - 3, oh0 - 3, b);
      if (load_w) tma_load_2d(sw, &p.tmW, bar, c0, 0);
      last_chunk = chunk;
    }
    const int c = c0 + 2 * lane;  // this lane's channel pair (C is even: both channels are in range or neither)
    const bool c_ok = c < p.C;
    unsigned long long acc[kDwPX];
    {
      const float2 bv = c_ok ? __ldg(reinterpret_cast<const float2*>(p.bias + c)) : make_float2(0.f, 0.f);
      const unsigned long long bb = (static_cast<unsigned long long>(__float_as_uint(bv.y)) << 32) | __float_as_uint(bv.x);
#pragma unroll
      for (int q = 0; q < kDwPX; ++q) acc[q] = bb;
    }
    mbar_wait(bar, phase);
    phase ^= 1;
#pragma unroll 1
    for (int kh = 0; kh < 7; ++kh) {
      const uint32_t* rowp = reinterpret_cast<const uint32_t*>(tile + ((r + kh) * kDwHW + hx * kDwPX) * kDwPixBytes) + lane;
      unsigned long long v[kDwPX + 6];
#pragma unroll
      for (int j = 0; j < kDwPX + 6; ++j) {
        const uint32_t u = rowp[j * (kDwPixBytes / 4)];
        v[j] = (static_cast<unsigned long long>(u & 0xffff0000u) << 32) | (u << 16);  // (lo -> .x, hi -> .y) as fp32 bits
      }
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const unsigned long long wv = *reinterpret_cast<const unsigned long long*>(sw + (kh * 7 + kw) * kDwCCH + 2 * lane);
#pragma unroll
        for (int q = 0; q < kDwPX; ++q) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[q]) : "l"(v[q + kw]), "l"(wv));
      }
    }
    const int oh = oh0 + r;
    uint32_t packed[kDwPX];
#pragma unroll
    for (int q = 0; q < kDwPX; ++q)
      packed[q] = pack_bf16(__uint_as_float(static_cast<uint32_t>(acc[q] & 0xffffffffull)), __uint_as_float(static_cast<uint32_t>(acc[q] >> 32)));
    if (oh < p.H && c_ok) {
      uint32_t* yr = reinterpret_cast<uint32_t*>(p.y + (static_cast<long>(b) * p.H + oh) * p.W * p.C + c);
#pragma unroll
      for (int q = 0; q < kDwPX; ++q) {
        const int ow = ow0 + hx * kDwPX + q;
        if (ow < p.W) yr[static_cast<long>(ow) * C2] = packed[q];  // a warp writes the 128 contiguous bytes of one pixel's chunk
      }
    }
    if (p.ln_stats) {
      // 16 values per lane (8 pixels x {sum, sumsq} of its channel pair) summed over the 32 lanes with a halving butterfly
      // (16 shuffles instead of 80); lane 2k and 2k+1 end up with the total of value k.  Lanes of out-of-range channels hold 0.
      float a[16];
#pragma unroll
      for (int q = 0; q < kDwPX; ++q) {
        const float r0 = bf16lo(packed[q]), r1 = bf16hi(packed[q]);
        a[q] = r0 + r1;
        a[8 + q] = fmaf(r0, r0, r1 * r1);
      }
      float b8[8], b4[4], b2[2], b1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool up = lane & 16;
        const float recv = __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + 8], 16);
        b8[i] = (up ? a[i + 8] : a[i]) + recv;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool up = lane & 8;
        const float recv = __shfl_xor_sync(0xffffffffu, up ? b8[i] : b8[i + 4], 8);
        b4[i] = (up ? b8[i + 4] : b8[i]) + recv;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool up = lane & 4;
        const float recv = __shfl_xor_sync(0xffffffffu, up ? b4[i] : b4[i + 2], 4);
        b2[i] = (up ? b4[i + 2] : b4[i]) + recv;
      }
      {
        const bool up = lane & 2;
        const float recv = __shfl_xor_sync(0xffffffffu, up ? b2[0] : b2[1], 2);
        b1 = (up ? b2[1] : b2[0]) + recv;
      }
      b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
      // value index held by this lane: bit 3 <- lane bit 4, bit 2 <- lane bit 3, bit 1 <- lane bit 2, bit 0 <- lane bit 1
      const int k = (lane >> 1) & 15;
      const int q = k & 7, ow = ow0 + hx * kDwPX + q;
      if ((lane & 1) == 0 && oh < p.H && ow < p.W) {
        unsigned long long* dst = p.ln_stats + ((static_cast<long>(b) * p.H + oh) * p.W + ow) * 2 + (k >> 3);
        atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn(b1 * kGnFixedScale)));
      }
    }
    __syncthreads();  // every thread has read the staged box / taps: the next item's TMA may overwrite them
  }
}

static bool dw_use_tiled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UC_DW_TILED"); v = (e && e[0] == '1') ? 1 : 0; }
  return v != 0;
}

}  // namespace uc

using namespace uc;

extern "C" int uc_dwconv7_tiled(const void* x_bf16, const float* w49, const float* bias, void* y_bf16, int B, int H, int W, int C,
                                void* ln_stats, void* stream_v);

extern "C" int uc_dwconv7(const void* x_bf16, const float* w49, const float* bias, void* y_bf16, int B, int H, int W, int C,
                          void* ln_stats, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!x_bf16 || !w49 || !bias || !y_bf16) return set_error(UC_EINVAL, "uc_dwconv7: null pointer");
  if (x_bf16 == y_bf16) return set_error(UC_EINVAL, "uc_dwconv7: not an in-place operation");
  if (B < 1 || H < 1 || W < 1 || C < 8) return set_error(UC_EINVAL, "uc_dwconv7: bad sizes");
  if (dw_use_tiled() || C % 8 || (reinterpret_cast<uintptr_t>(x_bf16) & 15) || (reinterpret_cast<uintptr_t>(w49) & 15) ||
      (reinterpret_cast<uintptr_t>(bias) & 7))
    return uc_dwconv7_tiled(x_bf16, w49, bias, y_bf16, B, H, W, C, ln_stats, stream_v);  // cp.async kernel (no TMA alignment needs)
  int rc = ensure_driver();
  if (rc) return rc;
  DwParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(H) * W * C * 2};
    uint32_t box[4] = {kDwCCH, kDwHW, kDwHH, 1};
    rc = encode_tmap(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x_bf16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(C), 49};
    uint64_t strides[1] = {static_cast<uint64_t>(C) * 4};
    uint32_t box[2] = {kDwCCH, 49};
    rc = encode_tmap(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w49, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  p.bias = bias;
  p.y = static_cast<uint16_t*>(y_bf16);
  p.ln_stats = static_cast<unsigned long long*>(ln_stats);
  p.H = H; p.W = W; p.C = C; p.B = B;
  p.tiles_w = (W + kDwTW - 1) / kDwTW;
  p.tiles_h = (H + kDwTH - 1) / kDwTH;
  const long items = static_cast<long>(p.tiles_w) * p.tiles_h * B * ((C + kDwCCH - 1) / kDwCCH);
  if (items > 0x7fffffffL) return set_error(UC_EINVAL, "uc_dwconv7: too many tiles");
  p.n_items = static_cast<int>(items);
  static PerDeviceFlag attr_dev;
  bool& attr = attr_dev.get();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "uc_dwconv7: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int grid = static_cast<int>(std::min<long>(items, static_cast<long>(num_sms()) * kDwCtasPerSm));
  cudaError_t e = launch_pdl(dwconv7_tma_kernel, dim3(grid), dim3(kDwThreads), kDwSmem, stream, p);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "uc_dwconv7 launch: %s", cudaGetErrorString(e));
  return check_launch("uc_dwconv7");
}
