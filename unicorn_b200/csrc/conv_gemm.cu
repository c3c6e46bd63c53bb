// Implicit-GEMM convolution / linear layer on Blackwell tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile.  The 128 pixels are a tile_w x tile_h patch of
// the NHWC output map; for every filter tap the TMA engine fetches the shifted tile_w x tile_h x 64-channel
// box of the input straight into 128B-swizzled shared memory (out-of-bounds coordinates are zero-filled by
// the hardware, which is the convolution's zero padding), so no im2col buffer ever exists.  A single elected
// thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16, bf16/fp16 in, fp32 accumulate in TMEM); four epilogue
// warps read the accumulator back with tcgen05.ld and apply bias / activation / layer-scale+residual, and
// optionally accumulate GroupNorm statistics, before a vectorised NHWC store.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
// Reference call sites replaced: see include/unicorn_b200.h (uc_conv2d).
#include <algorithm>
#include "uc_ptx.cuh"
#include "uc_common.h"
#include "../../include/unicorn_b200.h"

namespace uc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 16-bit elements -> 128-byte rows
constexpr int kMaxTaps = 9;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kConvEpiWarps = 16;                       // four per TMEM lane quadrant
constexpr int kConvThreads = (2 + kConvEpiWarps) * 32;  // warp 0 TMA, warp 1 MMA, then the epilogue warps

struct ConvTap {
  int16_t map, dw, dh, tap;
};

struct alignas(64) ConvKernelParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmBh;  // half-height weight box for the 2-CTA multicast variant
  CUtensorMap tmC;  // output map for the TMA store (16-bit outputs)
  ConvTap taps[kMaxTaps];
  int ntaps, kchunks;
  int n_tiles, m_tiles;
  int tile_w, tile_h, tiles_w, tiles_h;
  int Wo, Ho, B, Cout;
  uint32_t idesc;
  const float* bias;
  const float* gamma;
  const void* res;
  int ldres;
  void* y;
  int ldy, y_dtype, act, tma_store;
  long long* gn_stats;  // fixed-point (2^22) accumulators: order-independent, hence deterministic
  int gn_groups, gn_gs;  // gs = Cout / groups
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact-erf GELU (nn.GELU()) with erf from Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below the bf16 output ulp):
// two MUFU ops (rcp, ex2) and ~12 FMAs instead of libdevice erff's long dependent chain — the epilogue warps have
// nobody to hide latency behind.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = fast_ex2(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-poly, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case UC_ACT_RELU: return fmaxf(x, 0.f);
    case UC_ACT_GELU: return gelu_erf(x);
    case UC_ACT_SILU: return __fdividef(x, 1.f + fast_ex2(-x * 1.4426950408889634f));
    case UC_ACT_SIGMOID: return __fdividef(1.f, 1.f + fast_ex2(-x * 1.4426950408889634f));
    default: return x;
  }
}

// Epilogue of one 32-column chunk of one tile, specialised on the activation so that the inner loops are branch-free.
template <int ACT>
__device__ __forceinline__ void epi_math(float (&f)[32], const ConvKernelParams& p, int cbase, int ncols, bool valid, int b,
                                         int lane, bool tile_ok) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
        f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
      }
    }
  }
  if (p.gn_stats) {
    // per-group partial sums of this warp's 32 rows x chunk columns; a group may span several chunks — partial sums
    // are simply added by the (order-independent) integer atomics.
    float gs_sum = 0.f, gs_sq = 0.f;
    int gs_left = p.gn_gs - (cbase % p.gn_gs);
    int gs_group = cbase / p.gn_gs;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) {
        const float x = valid ? f[j] : 0.f;
        gs_sum += x;
        gs_sq += x * x;
        if (--gs_left == 0 || j == ncols - 1) {
          float s1 = gs_sum, s2 = gs_sq;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0 && tile_ok) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.gn_stats) + (static_cast<size_t>(b) * p.gn_groups + gs_group) * 2;
            atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn(s1 * kGnFixedScale)));
            atomicAdd(dst + 1, static_cast<unsigned long long>(__float2ll_rn(s2 * kGnFixedScale)));
          }
          gs_sum = 0.f; gs_sq = 0.f;
          if (gs_left == 0) { gs_left = p.gn_gs; ++gs_group; }
        }
      }
    }
  }
  if (ACT != UC_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], ACT);
  }
  if (p.gamma) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + cbase + j));
        f[j] *= g.x; f[j + 1] *= g.y; f[j + 2] *= g.z; f[j + 3] *= g.w;
      }
    }
  }
}

// Persistent kernel: grid = min(#tiles, resident CTAs); every CTA walks tiles tile = blockIdx.x + i * gridDim.x (N tile
// fastest, so CTAs running side by side share the activation tile in L2).  The TMEM accumulator is double buffered:
// the MMA warp fills accumulator (i+1)&1 while the epilogue warps drain accumulator i&1.
// CLUSTER = 2 is the cta_group::2 variant: two CTAs (a cluster = one TPC's SM pair) with consecutive M tiles and the same N
// tile compute a 256 x BLOCK_N tile with ONE pair-MMA stream issued by the leader CTA.  Each CTA stages its own 128 rows
// of A and only HALF of the weight box (BLOCK_N/2 rows): shared-memory traffic per MAC (TMA writes + MMA operand reads,
// which is what bounds the single-CTA kernel on K-deep layers) drops by a third.  Barriers: the leader's full[stage]
// collects the bytes of all four loads; one tcgen05.commit.cta_group::2 multicast releases the stage / publishes the
// accumulator in both CTAs; the peer's epilogue warps hand their accumulator back with remote arrives on the leader.
template <int BLOCK_N, int STAGES, int CLUSTER>
__global__ void __launch_bounds__(kConvThreads) conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  constexpr int B_BYTES = (BLOCK_N / CLUSTER) * kBlockK * 2;  // per-CTA weight bytes per stage
  constexpr uint32_t ACC_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr int C_BLOCKS = (BLOCK_N % 64 == 0) ? BLOCK_N / 64 : 0;  // 64-channel staging blocks for the TMA store
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint8_t* sC = sB + STAGES * B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sC + C_BLOCKS * kABytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters = p.ntaps * p.kchunks;
  const int crank = CLUSTER > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  // work items: (N tile, group of CLUSTER consecutive M tiles); this CTA takes M tile group*CLUSTER + crank
  const int num_items = p.n_tiles * ((p.m_tiles + CLUSTER - 1) / CLUSTER);
  const int item0 = blockIdx.x / CLUSTER, item_step = gridDim.x / CLUSTER;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[0]);
    prefetch_tmap(&p.tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kConvEpiWarps * CLUSTER);  // pair mode: the peer's epilogue warps arrive remotely
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CLUSTER > 1) { tmem_alloc_2sm(tmem_slot, TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer: the whole warp walks the loop (converged), one elected lane issues
    int stage = 0, phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int n0 = (item % p.n_tiles) * BLOCK_N;
      const int mt = (item / p.n_tiles) * CLUSTER + crank;
      const int ow0 = (mt % p.tiles_w) * p.tile_w, oh0 = ((mt / p.tiles_w) % p.tiles_h) * p.tile_h;
      const int b = mt / (p.tiles_w * p.tiles_h);
      for (int t = 0; t < p.ntaps; ++t) {
        const ConvTap tp = p.taps[t];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            if (CLUSTER > 1) {
              // the leader arms its barrier for the bytes of both CTAs; the peer's loads are credited to it as well
              if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (kABytes + B_BYTES));
              tma_load_4d_2sm(sA + stage * kABytes, &p.tmA[tp.map], &full[stage], kc * kBlockK, ow0 + tp.dw, oh0 + tp.dh, b);
              tma_load_3d_2sm(sB + stage * B_BYTES, &p.tmBh, &full[stage], kc * kBlockK, tp.tap, n0 + crank * (BLOCK_N / 2));
            } else {
              mbar_arrive_expect_tx(&full[stage], kABytes + B_BYTES);
              tma_load_4d(sA + stage * kABytes, &p.tmA[tp.map], &full[stage], kc * kBlockK, ow0 + tp.dw, oh0 + tp.dh, b);
              tma_load_3d(sB + stage * B_BYTES, &p.tmB, &full[stage], kc * kBlockK, tp.tap, n0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (pair mode: leader CTA only): converged warp, one elected lane issues.
    // Shared-memory descriptors are built once per stage: inside the K loop only their start-address field advances.
    if (crank == 0) {
      const uint32_t idesc = p.idesc;
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA)), b_desc0 = umma_desc_sw128(smem_u32(sB));
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            // descriptor start address is in 16-byte units: stage offsets and the 32-byte K step are plain adds
            const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * kABytes) >> 4);
            const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (CLUSTER > 1) umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
              else umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
            }
            // frees this smem stage (in both CTAs of the pair) when the MMAs above have read it
            if (CLUSTER > 1) umma_commit_2sm_mc(&empty[stage], static_cast<uint16_t>(0x3));
            else umma_commit(&empty[stage]);
            if (it == kiters - 1) {  // accumulator complete (in both CTAs' tensor memory)
              if (CLUSTER > 1) umma_commit_2sm_mc(&tmem_full[acc], static_cast<uint16_t>(0x3));
              else umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> fused math -> NHWC global
    // kConvEpiWarps warps: warp w owns TMEM lane quadrant (w & 3) and the 32-column chunks ci with ci % 4 == (w-2)/4,
    // so four warps share each scheduler and hide each other's latencies.
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int wi = row % p.tile_w, hi = row / p.tile_w;
    int acc = 0, acc_phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int n0 = (item % p.n_tiles) * BLOCK_N;
      const int mt = (item / p.n_tiles) * CLUSTER + crank;
      const int ow0 = (mt % p.tiles_w) * p.tile_w, oh0 = ((mt / p.tiles_w) % p.tiles_h) * p.tile_h;
      const int b = mt / (p.tiles_w * p.tiles_h);
      const int ow = ow0 + wi, oh = oh0 + hi;
      const bool valid = (ow < p.Wo) && (oh < p.Ho) && (mt < p.m_tiles);  // mt >= m_tiles: padding tile of an odd pair
      const size_t pix = (static_cast<size_t>(b) * p.Ho + oh) * p.Wo + ow;
      const int limit = min(BLOCK_N, p.Cout - n0);  // valid columns of this tile (multiple of 8)
      if (C_BLOCKS > 0 && p.tma_store) {
        // the staging blocks are about to be overwritten: the previous tile's TMA stores must have read them
        if (lane == 0 && q == 0 && cg < 2) tma_store_wait_read();  // the two issuing threads (warps 4 and 8)
        asm volatile("bar.sync 1, %0;" ::"n"(kConvEpiWarps * 32) : "memory");
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int r0 = 0; r0 < BLOCK_N; r0 += 128) {
        const int c0 = r0 + cg * 32;
        const bool last_round = (r0 + 128 >= BLOCK_N);
        if (c0 < limit && c0 < BLOCK_N) {
          uint32_t v[32];
          if constexpr (BLOCK_N % 32 == 0) {
            tmem_ld_32x32(t_acc + c0, v);
          } else {
            uint32_t h[16];
            tmem_ld_32x16(t_acc + c0, h);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[j + 16] = 0; }
          }
          tmem_ld_wait();
          if (last_round) {  // this warp's last read of the accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CLUSTER > 1 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
              else mbar_arrive(&tmem_empty[acc]);
            }
          }
          const int cbase = n0 + c0;
          const int ncols = min(32, limit - c0);  // multiple of 8
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          switch (p.act) {
            case UC_ACT_GELU: epi_math<UC_ACT_GELU>(f, p, cbase, ncols, valid, b, lane, mt < p.m_tiles); break;
            case UC_ACT_RELU: epi_math<UC_ACT_RELU>(f, p, cbase, ncols, valid, b, lane, mt < p.m_tiles); break;
            case UC_ACT_SILU: epi_math<UC_ACT_SILU>(f, p, cbase, ncols, valid, b, lane, mt < p.m_tiles); break;
            case UC_ACT_SIGMOID: epi_math<UC_ACT_SIGMOID>(f, p, cbase, ncols, valid, b, lane, mt < p.m_tiles); break;
            default: epi_math<UC_ACT_NONE>(f, p, cbase, ncols, valid, b, lane, mt < p.m_tiles); break;
          }
          if (valid) {
            if (p.res) {
              const uint16_t* r = reinterpret_cast<const uint16_t*>(p.res) + pix * p.ldres + cbase;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (j < ncols) {
                  const uint4 rv = __ldg(reinterpret_cast<const uint4*>(r + j));
                  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    f[j + 2 * t] += bits16_to_float(rw[t] & 0xffffu, p.y_dtype == UC_F16 ? UC_F16 : UC_BF16);
                    f[j + 2 * t + 1] += bits16_to_float(rw[t] >> 16, p.y_dtype == UC_F16 ? UC_F16 : UC_BF16);
                  }
                }
              }
            }
            if (p.y_dtype == UC_F32) {
              float* yp = reinterpret_cast<float*>(p.y) + pix * p.ldy + cbase;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < ncols) *reinterpret_cast<float4*>(yp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              }
            } else if (!p.tma_store) {
              uint16_t* yp = reinterpret_cast<uint16_t*>(p.y) + pix * p.ldy + cbase;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (j < ncols) {
                  uint4 o;
                  o.x = pack2_16(f[j], f[j + 1], p.y_dtype);
                  o.y = pack2_16(f[j + 2], f[j + 3], p.y_dtype);
                  o.z = pack2_16(f[j + 4], f[j + 5], p.y_dtype);
                  o.w = pack2_16(f[j + 6], f[j + 7], p.y_dtype);
                  *reinterpret_cast<uint4*>(yp + j) = o;
                }
              }
            }
          }
          if (C_BLOCKS > 0 && p.tma_store) {
            // Stage in shared memory in the 128B-swizzled layout of a TMA box (64 channels per block); the TMA engine
            // then writes full lines and clips the out-of-range rows / channels of edge tiles.
            uint8_t* blk = sC + (c0 >> 6) * kABytes;
            const int kb = (c0 & 32) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack2_16(f[8 * j], f[8 * j + 1], p.y_dtype);
              o.y = pack2_16(f[8 * j + 2], f[8 * j + 3], p.y_dtype);
              o.z = pack2_16(f[8 * j + 4], f[8 * j + 5], p.y_dtype);
              o.w = pack2_16(f[8 * j + 6], f[8 * j + 7], p.y_dtype);
              *reinterpret_cast<uint4*>(blk + row * 128 + (((kb + j) ^ (row & 7)) << 4)) = o;
            }
          }
        } else if (last_round) {
          // nothing to read in the last round (narrow or edge tile): still release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CLUSTER > 1 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        if (C_BLOCKS > 0 && p.tma_store && r0 < limit) {  // uniform over the epilogue warps
          fence_proxy_async();
          asm volatile("bar.sync 1, %0;" ::"n"(kConvEpiWarps * 32) : "memory");
          if (lane == 0 && q == 0 && cg < 2) {  // two issuing threads: warps 4 (block r0/64) and 8 (block r0/64 + 1)
            const int blk = (r0 >> 6) + cg;
            if (blk * 64 < limit) {
              tma_store_4d(&p.tmC, sC + blk * kABytes, n0 + blk * 64, ow0, oh0, b);
              tma_store_commit();
            }
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (C_BLOCKS > 0 && p.tma_store && lane == 0 && q == 0 && cg < 2) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // no CTA may exit while its peer can still signal its barriers / write its smem
  if (warp == 1) {
    tc_fence_after();
    if (CLUSTER > 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------- host side

template <int BLOCK_N, int STAGES, int CLUSTER>
static int launch_conv(ConvKernelParams& p, cudaStream_t stream) {
  constexpr int c_blocks = (BLOCK_N % 64 == 0) ? BLOCK_N / 64 : 0;
  constexpr int smem = STAGES * (kABytes + (BLOCK_N / CLUSTER) * kBlockK * 2) + c_blocks * kABytes + 1024 + 256;
  static int per_sm = 0;
  auto kern = conv_gemm_kernel<BLOCK_N, STAGES, CLUSTER>;
  if (!per_sm) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "conv_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kConvThreads, smem);
    if (e != cudaSuccess || n < 1) return set_error(UC_EINVAL, "conv_gemm<%d,%d>: does not fit on an SM", BLOCK_N, STAGES);
    constexpr int acc_cols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
    per_sm = std::min(n, 512 / (2 * acc_cols));  // TMEM: 512 columns per SM
  }
  const int items = p.n_tiles * ((p.m_tiles + CLUSTER - 1) / CLUSTER);
  int grid = std::min(items * CLUSTER, num_sms() * per_sm);
  grid -= grid % CLUSTER;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CLUSTER > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "conv_gemm<%d,%d,%d> launch: %s", BLOCK_N, STAGES, CLUSTER, cudaGetErrorString(e));
  return UC_OK;
}

static int pick_block_n(int Cout, int m_tiles, int gn_gs) {
  // Heuristic default for the persistent kernel (one CTA per SM for the wide tiles): fewest waves of the widest tile
  // that does not waste more than a third of its columns.  unicorn_b200/engine.py autotunes block_n per layer on top
  // of this (plan-time timing of the candidates), so this only has to be reasonable.
  static const int cands[] = {256, 192, 128, 96, 64, 32, 16};
  const int sms = num_sms();
  int best = 0;
  double best_cost = -1.0;
  for (int bn : cands) {
    if (gn_gs > 0 && (bn % gn_gs) != 0) continue;  // GroupNorm groups must not straddle N tiles
    const int nt = (Cout + bn - 1) / bn;
    const long waste_cols = static_cast<long>(nt) * bn - Cout;
    if (waste_cols * 3 > static_cast<long>(nt) * bn && bn > 16 && gn_gs <= 0) continue;
    const long tiles = static_cast<long>(nt) * m_tiles;
    const long waves = (tiles + sms - 1) / sms;  // one persistent CTA per SM
    const double cost = static_cast<double>(waves) * (bn + 40);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace uc

using namespace uc;

extern "C" int uc_conv2d(const UcConv2d* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!d || !d->x || !d->w || !d->y) return set_error(UC_EINVAL, "uc_conv2d: null pointer");
  if (d->x_dtype != UC_BF16 && d->x_dtype != UC_F16) return set_error(UC_EINVAL, "uc_conv2d: x must be bf16/f16");
  if (d->Cin % 8 || d->Cout % 8 || d->ldx % 8 || d->ldy % 8 || d->ldx < d->Cin || d->ldy < d->Cout)
    return set_error(UC_EINVAL, "uc_conv2d: Cin/Cout/ldx/ldy must be multiples of 8 (Cin=%d Cout=%d ldx=%d ldy=%d)",
                     d->Cin, d->Cout, d->ldx, d->ldy);
  if (d->stride != 1 && d->stride != 2) return set_error(UC_EINVAL, "uc_conv2d: stride must be 1 or 2");
  if (d->KH * d->KW > kMaxTaps || d->KH < 1 || d->KW < 1) return set_error(UC_EINVAL, "uc_conv2d: at most 9 taps");
  if (d->pad < 0 || d->pad >= d->KH + 1) return set_error(UC_EINVAL, "uc_conv2d: bad pad");
  if (d->res && (d->ldres % 8 || d->y_dtype == UC_F32)) return set_error(UC_EINVAL, "uc_conv2d: residual needs 16-bit y, ldres%%8==0");
  if ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->w) | reinterpret_cast<uintptr_t>(d->y)) & 15)
    return set_error(UC_EINVAL, "uc_conv2d: pointers must be 16-byte aligned");
  if (d->gn_stats && (d->gn_groups <= 0 || d->Cout % d->gn_groups))
    return set_error(UC_EINVAL, "uc_conv2d: bad GroupNorm grouping");
  int rc = ensure_driver();
  if (rc) return rc;

  const int s = d->stride;
  const int Ho = (d->H + 2 * d->pad - d->KH) / s + 1;
  const int Wo = (d->W + 2 * d->pad - d->KW) / s + 1;
  if (Ho <= 0 || Wo <= 0) return set_error(UC_EINVAL, "uc_conv2d: empty output");

  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  const CUtensorMapDataType dt = d->x_dtype == UC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const bool flat = (d->KH == 1 && d->KW == 1 && s == 1 && d->pad == 0);
  int B = d->B, Hm = d->H, Wm = d->W;  // map geometry
  if (flat) { Wm = d->B * d->H * d->W; Hm = 1; B = 1; }
  p.Wo = flat ? Wm : Wo;
  p.Ho = flat ? 1 : Ho;
  p.B = B;
  // tile shape minimising the number of 128-pixel tiles
  {
    int best_tw = 128, best_th = 1;
    long best = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
      const int th = 128 / tw;
      const long n = static_cast<long>((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
      if (best < 0 || n < best) { best = n; best_tw = tw; best_th = th; }
    }
    p.tile_w = best_tw; p.tile_h = best_th;
  }
  p.tiles_w = (p.Wo + p.tile_w - 1) / p.tile_w;
  p.tiles_h = (p.Ho + p.tile_h - 1) / p.tile_h;
  const int m_tiles = p.tiles_w * p.tiles_h * B;

  // activation maps: one per stride phase
  const size_t es = 2;
  for (int ph = 0; ph < s; ++ph) {
    for (int pw = 0; pw < s; ++pw) {
      const int Wp = (Wm - pw + s - 1) / s, Hp = (Hm - ph + s - 1) / s;
      if (Wp <= 0 || Hp <= 0) continue;
      const uint8_t* base = reinterpret_cast<const uint8_t*>(d->x) + (static_cast<size_t>(ph) * Wm + pw) * d->ldx * es;
      uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(Wp), static_cast<uint64_t>(Hp), static_cast<uint64_t>(B)};
      uint64_t strides[3] = {static_cast<uint64_t>(s) * d->ldx * es, static_cast<uint64_t>(s) * Wm * d->ldx * es,
                             static_cast<uint64_t>(Hm) * Wm * d->ldx * es};
      uint32_t box[4] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
      rc = encode_tmap(&p.tmA[ph * s + pw], dt, 4, base, dims, strides, box);
      if (rc) return rc;
    }
  }
  int nt = 0;
  for (int kh = 0; kh < d->KH; ++kh) {
    for (int kw = 0; kw < d->KW; ++kw) {
      const int offh = kh - d->pad, offw = kw - d->pad;
      const int ph = ((offh % s) + s) % s, pw = ((offw % s) + s) % s;
      ConvTap t;
      t.map = static_cast<int16_t>(ph * s + pw);
      t.dh = static_cast<int16_t>((offh - ph) / s);
      t.dw = static_cast<int16_t>((offw - pw) / s);
      t.tap = static_cast<int16_t>(kh * d->KW + kw);
      p.taps[nt++] = t;
    }
  }
  p.ntaps = nt;
  p.kchunks = (d->Cin + kBlockK - 1) / kBlockK;

  const int gn_gs = d->gn_stats ? d->Cout / d->gn_groups : 0;
  // block_n >= 1000 selects the cta_group::2 pair variant (1128 / 1192 / 1256)
  const bool cluster2 = d->block_n >= 1000;
  const int bn = cluster2 ? d->block_n - 1000 : d->block_n ? d->block_n : pick_block_n(d->Cout, m_tiles, gn_gs);
  if (bn == 0) return set_error(UC_EINVAL, "uc_conv2d: no N tile compatible with GroupNorm group size %d", gn_gs);
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(nt), static_cast<uint64_t>(d->Cout)};
    uint64_t strides[2] = {static_cast<uint64_t>(d->Cin) * es, static_cast<uint64_t>(nt) * d->Cin * es};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1, static_cast<uint32_t>(bn)};
    rc = encode_tmap(&p.tmB, dt, 3, d->w, dims, strides, box);
    if (rc) return rc;
  }
  p.tma_store = 0;
  if (d->y_dtype != UC_F32 && (bn % 64) == 0) {  // every 64-channel store box must lie inside this CTA's N tile
    const CUtensorMapDataType dty = d->y_dtype == UC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    uint64_t dims[4] = {static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(p.Wo), static_cast<uint64_t>(p.Ho), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(d->ldy) * es, static_cast<uint64_t>(p.Wo) * d->ldy * es,
                           static_cast<uint64_t>(p.Ho) * p.Wo * d->ldy * es};
    uint32_t box[4] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
    rc = encode_tmap(&p.tmC, dty, 4, d->y, dims, strides, box);
    if (rc) return rc;
    p.tma_store = 1;
  }
  p.Cout = d->Cout;
  p.idesc = umma_idesc_f16(d->x_dtype == UC_BF16 ? 1u : 0u, kBlockM, static_cast<uint32_t>(bn));
  p.bias = d->bias; p.gamma = d->gamma; p.res = d->res; p.ldres = d->ldres;
  p.y = d->y; p.ldy = d->ldy; p.y_dtype = d->y_dtype; p.act = d->act;
  p.gn_stats = static_cast<long long*>(d->gn_stats); p.gn_groups = d->gn_groups;
  p.gn_gs = d->gn_stats ? d->Cout / d->gn_groups : 1 << 30;
  if (d->gn_stats && (bn % p.gn_gs) != 0)
    return set_error(UC_EINVAL, "uc_conv2d: N tile %d incompatible with GroupNorm group size %d", bn, p.gn_gs);
  p.n_tiles = (d->Cout + bn - 1) / bn;
  p.m_tiles = m_tiles;
  if (cluster2) {
    uint64_t dims[3] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(nt), static_cast<uint64_t>(d->Cout)};
    uint64_t strides[2] = {static_cast<uint64_t>(d->Cin) * es, static_cast<uint64_t>(nt) * d->Cin * es};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1, static_cast<uint32_t>(bn / 2)};
    rc = encode_tmap(&p.tmBh, dt, 3, d->w, dims, strides, box);
    if (rc) return rc;
    p.idesc = umma_idesc_f16(d->x_dtype == UC_BF16 ? 1u : 0u, 2 * kBlockM, static_cast<uint32_t>(bn));  // UMMA M = 256
    switch (bn) {
      case 256: return launch_conv<256, 4, 2>(p, stream);  // 4 x 32 KB ring + 64 KB staging
      case 192: return launch_conv<192, 5, 2>(p, stream);  // 5 x 28 KB + 48 KB
      case 128: return launch_conv<128, 6, 2>(p, stream);  // 6 x 24 KB + 32 KB
      default: return set_error(UC_EINVAL, "uc_conv2d: the cta_group::2 variant exists for block_n 128/192/256 only");
    }
  }
  switch (bn) {
    case 256: return launch_conv<256, 3, 1>(p, stream);
    case 192: return launch_conv<192, 4, 1>(p, stream);
    case 128: return launch_conv<128, 5, 1>(p, stream);
    case 96: return launch_conv<96, 3, 1>(p, stream);
    case 64: return launch_conv<64, 3, 1>(p, stream);
    case 32: return launch_conv<32, 4, 1>(p, stream);
    case 16: return launch_conv<16, 4, 1>(p, stream);
    default: return set_error(UC_EINVAL, "uc_conv2d: unsupported block_n %d", bn);
  }
}
