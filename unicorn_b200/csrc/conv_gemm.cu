// Implicit-GEMM convolution / linear layer on Blackwell tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile.  The 128 pixels are a tile_w x tile_h patch of
// the NHWC output map; for every filter tap the TMA engine fetches the shifted tile_w x tile_h x 64-channel
// box of the input straight into 128B-swizzled shared memory (out-of-bounds coordinates are zero-filled by
// the hardware, which is the convolution's zero padding), so no im2col buffer ever exists.  A single elected
// thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16, bf16/fp16 in, fp32 accumulate in a double-buffered TMEM
// accumulator); sixteen epilogue warps read the accumulator back with tcgen05.ld and apply bias / activation /
// layer-scale+residual, and optionally accumulate GroupNorm statistics, before 256-bit (one L2 sector per lane) stores.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..17 = epilogue.
// Reference call sites replaced: see include/unicorn_b200.h (uc_conv2d).
#include <algorithm>
#include <stdlib.h>
#include "uc_ptx.cuh"
#include "uc_common.h"
#include "../../include/unicorn_b200.h"
#include "uc_epilogue.cuh"

namespace uc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 16-bit elements -> 128-byte rows
constexpr int kMaxTaps = 9;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kConvEpiWarps = 16;                       // four per TMEM lane quadrant
constexpr int kConvThreads = (2 + kConvEpiWarps) * 32;  // warp 0 TMA, warp 1 MMA, then the epilogue warps
constexpr int kGnMaxLocal = 64;                         // GroupNorm groups per N tile (tile width 256 / group size >= 4)
constexpr int kGnSmemBytes = 2 * kGnMaxLocal * 2 * 8;   // two tile parities x {sum, sumsq} int64

struct ConvTap {
  int16_t map, dw, dh, tap;
};

struct alignas(64) ConvKernelParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmBh;  // half-height weight box for the 2-CTA multicast variant
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
  int ldy, y_dtype, act;
  int wide_store, wide_res;  // 256-bit stores / residual loads possible (32-byte aligned rows)
  int debug;  // tools only (UC_CONV_DEBUG): 1 = no MMA (operand feed rate alone), 2 = no TMA loads (MMA + epilogue alone)
  const long long* row_stats;  // LayerNorm folded into this 1x1 conv: per input pixel {sum, sumsq} (fixed point 2^22) ...
  const float* col_s;          // ... column sums of the folded weights, channel count and epsilon of the LayerNorm
  float row_inv, row_eps;  // row_inv = 1 / (2^22 * Cin)
  long long* gn_stats;  // fixed-point (2^22) accumulators: order-independent, hence deterministic
  int gn_groups, gn_gs;  // gs = Cout / groups
};

// GroupNorm partial sums of one epilogue item (this warp's 32 rows x ncols columns starting at channel cbase), added to the
// CTA's shared-memory accumulators of the current tile (fixed point, integer adds: order independent).  g0 = first group of
// the N tile (an N tile never splits a group).
__device__ __forceinline__ void gn_partial_sums(const float (&f)[16], const ConvKernelParams& p, int cbase, int ncols, bool valid,
                                             int lane, unsigned long long* acc_tile, int g0) {
  int c = 0;
#pragma unroll 1
  while (c < ncols) {
    const int g = (cbase + c) / p.gn_gs;
    const int end = min(ncols, (g + 1) * p.gn_gs - cbase);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = (valid && j >= c && j < end) ? f[j] : 0.f;
      s1 += x;
      s2 = fmaf(x, x, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      atomicAdd(acc_tile + (g - g0) * 2, static_cast<unsigned long long>(__float2ll_rn(s1 * kGnFixedScale)));
      atomicAdd(acc_tile + (g - g0) * 2 + 1, static_cast<unsigned long long>(__float2ll_rn(s2 * kGnFixedScale)));
    }
    c = end;
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
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  constexpr int B_BYTES = (BLOCK_N / CLUSTER) * kBlockK * 2;  // per-CTA weight bytes per stage
  constexpr uint32_t ACC_COLS = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // per-CTA GroupNorm accumulators (fixed point): the epilogue warps add into shared memory, ONE global atomic per group and
  // tile follows — the short-K GN convs were bound by ~36k global atomics on the 32 addresses of an image
  unsigned long long* gn_acc = reinterpret_cast<unsigned long long*>(smem + STAGES * (kABytes + B_BYTES) + 256);

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
  for (int i = threadIdx.x; i < 2 * kGnMaxLocal * 2; i += kConvThreads) gn_acc[i] = 0ull;
  if (warp == 1) {
    if (CLUSTER > 1) { tmem_alloc_2sm(tmem_slot, TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
  // tail of the previous kernel in the stream; global memory is touched only after it has completed.
  pdl_wait();
  pdl_launch_dependents();

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
            if (CLUSTER > 1 && p.debug == 2) {
              if (crank == 0) mbar_arrive(&full[stage]);
            } else if (CLUSTER > 1) {
              // the leader arms its barrier for the bytes of both CTAs; the peer's loads are credited to it as well
              if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (kABytes + B_BYTES));
              tma_load_4d_2sm(sA + stage * kABytes, &p.tmA[tp.map], &full[stage], kc * kBlockK, ow0 + tp.dw, oh0 + tp.dh, b);
              tma_load_3d_2sm(sB + stage * B_BYTES, &p.tmBh, &full[stage], kc * kBlockK, tp.tap, n0 + crank * (BLOCK_N / 2));
            } else if (p.debug == 2) {
              mbar_arrive(&full[stage]);
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
          if (CLUSTER == 1 && p.debug == 1) {
            if (elect_one()) {
              mbar_arrive(&empty[stage]);
              if (it == kiters - 1) mbar_arrive(&tmem_full[acc]);
            }
          } else if (elect_one()) {
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
    // 16 warps = 4 TMEM lane quadrants (q = warp % 4, fixed by the hardware) x 4 column groups (cg).  Work item = 32 rows x 16
    // channels: warp (q, cg) owns rows 32q..32q+31 and the channels 64 rd + 16 cg of round rd, so the warps are balanced for
    // every N tile.  A lane holds 16 consecutive channels of one pixel = 32 bytes of bf16 = exactly one L2 sector, written
    // with ONE 256-bit store (no partial sectors, no staging buffer, no barrier): the warps are completely independent and
    // drift apart, which is what hides the TMEM / L2 / MUFU latencies of one another.
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int wi = row % p.tile_w, hi = row / p.tile_w;
    const bool f16 = p.y_dtype == UC_F16;
    constexpr int ROUNDS = (BLOCK_N + 63) / 64;
    int acc = 0, acc_phase = 0, gpar = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int n0 = (item % p.n_tiles) * BLOCK_N;
      const int mt = (item / p.n_tiles) * CLUSTER + crank;
      const int ow0 = (mt % p.tiles_w) * p.tile_w, oh0 = ((mt / p.tiles_w) % p.tiles_h) * p.tile_h;
      const int b = mt / (p.tiles_w * p.tiles_h);
      const int ow = ow0 + wi, oh = oh0 + hi;
      const bool tile_ok = mt < p.m_tiles;  // false: padding tile of an odd pair
      const bool valid = (ow < p.Wo) && (oh < p.Ho) && tile_ok;
      const size_t pix = (static_cast<size_t>(b) * p.Ho + oh) * p.Wo + ow;
      const int limit = min(BLOCK_N, p.Cout - n0);  // valid columns of this tile (multiple of 8)
      // LayerNorm folded into the GEMM: y = rstd * (W' x) - rstd * mu * colsum(W') + c ; (mu, rstd) of this lane's pixel
      float r_rstd = 1.f, r_murstd = 0.f;
      if (p.row_stats && valid) {  // one 128-bit load; fp32 is enough here (|mu| <~ 10 sigma for a ConvNeXt block's depthwise output)
        const longlong2 st = __ldg(reinterpret_cast<const longlong2*>(p.row_stats) + pix);
        const float mu = static_cast<float>(st.x) * p.row_inv;
        const float var = fmaxf(fmaf(-mu, mu, static_cast<float>(st.y) * p.row_inv), 0.f);
        r_rstd = rsqrtf(var + p.row_eps);
        r_murstd = mu * r_rstd;
      }
      // this warp's last round with columns to read: the accumulator is handed back to the MMA warp right after it
      const int last_rd = (limit - 1 - cg * 16) >= 0 ? min(ROUNDS - 1, (limit - 1 - cg * 16) / 64) : -1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int rd = 0; rd <= last_rd; ++rd) {
        const int c0 = rd * 64 + cg * 16;
        const int cbase = n0 + c0;
        const int ncols = min(16, limit - c0);  // 8 or 16
        uint32_t v[16];
        tmem_ld_32x16(t_acc + c0, v);
        float4 bb[4];
        if (p.bias) {  // in flight together with the TMEM load
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = (4 * j < ncols) ? __ldg(reinterpret_cast<const float4*>(p.bias + cbase) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t rw[8];
        const bool has_res = p.res && valid;
        if (has_res) {  // residual: the same 32-byte sector of the shortcut tensor
          const uint16_t* r = reinterpret_cast<const uint16_t*>(p.res) + pix * p.ldres + cbase;
          if (p.wide_res && ncols == 16) {
            ldg_v8(r, rw);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4 rv = make_uint4(0, 0, 0, 0);
              if (8 * j < ncols) rv = __ldg(reinterpret_cast<const uint4*>(r) + j);
              rw[4 * j] = rv.x; rw[4 * j + 1] = rv.y; rw[4 * j + 2] = rv.z; rw[4 * j + 3] = rv.w;
            }
          }
        }
        tmem_ld_wait();
        if (rd == last_rd) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CLUSTER > 1 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
        f32x2 h[8];
        if (p.row_stats) {
          const f32x2 rs = pk2(r_rstd, r_rstd), nm = pk2(-r_murstd, -r_murstd);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 cs = (4 * j < ncols) ? __ldg(reinterpret_cast<const float4*>(p.col_s + cbase) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            h[2 * j] = fma2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), rs, fma2(nm, pk2(cs.x, cs.y), pk2(bb[j].x, bb[j].y)));
            h[2 * j + 1] = fma2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), rs, fma2(nm, pk2(cs.z, cs.w), pk2(bb[j].z, bb[j].w)));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            h[2 * j] = add2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pk2(bb[j].x, bb[j].y));
            h[2 * j + 1] = add2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pk2(bb[j].z, bb[j].w));
          }
        }
        if (p.act == UC_ACT_GELU && !p.gn_stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = gelu2(h[j]);
        } else if (p.gn_stats || p.act != UC_ACT_NONE) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) { f[2 * j] = lo2(h[j]); f[2 * j + 1] = hi2(h[j]); }
          if (p.gn_stats) gn_partial_sums(f, p, cbase, ncols, valid, lane, gn_acc + gpar * (kGnMaxLocal * 2), n0 / p.gn_gs);
          switch (p.act) {
            case UC_ACT_RELU:
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
              break;
            case UC_ACT_NONE: break;
            default:
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = apply_act(f[j], p.act);
              break;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = pk2(f[2 * j], f[2 * j + 1]);
        }
        if (p.gamma) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (4 * j < ncols) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + cbase) + j);
              h[2 * j] = mul2(h[2 * j], pk2(g.x, g.y));
              h[2 * j + 1] = mul2(h[2 * j + 1], pk2(g.z, g.w));
            }
          }
        }
        if (has_res) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const f32x2 rr = f16 ? pk2(bits16_to_float(rw[t] & 0xffffu, UC_F16), bits16_to_float(rw[t] >> 16, UC_F16))
                                 : pk2(bf16lo(rw[t]), bf16hi(rw[t]));
            h[t] = add2(h[t], rr);
          }
        }
        if (valid) {
          if (p.y_dtype == UC_F32) {
            float* yp = reinterpret_cast<float*>(p.y) + pix * p.ldy + cbase;
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[2 * j] = __float_as_uint(lo2(h[j])); o[2 * j + 1] = __float_as_uint(hi2(h[j])); }
            if (p.wide_store) {
              stg_v8(yp, o);
              if (ncols == 16) stg_v8(yp + 8, o + 8);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (4 * j < ncols) *(reinterpret_cast<uint4*>(yp) + j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              }
            }
          } else {
            uint16_t* yp = reinterpret_cast<uint16_t*>(p.y) + pix * p.ldy + cbase;
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = pack2_fast(lo2(h[j]), hi2(h[j]), f16);
            if (p.wide_store && ncols == 16) {
              stg_v8(yp, o);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                if (8 * j < ncols) *(reinterpret_cast<uint4*>(yp) + j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              }
            }
          }
        }
      }
      if (last_rd < 0) {  // narrow or edge tile: nothing to read for this warp, still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CLUSTER > 1 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      if (p.gn_stats) {
        // every epilogue warp has added its partial sums of this tile: one global atomic per group, then the slots are
        // cleared for the tile after next (the next tile uses the other parity, so no second barrier is needed)
        asm volatile("bar.sync 1, %0;" ::"n"(kConvEpiWarps * 32) : "memory");
        const int et = static_cast<int>(threadIdx.x) - 64;  // 0 .. 511 over the epilogue warps
        const int ng = (limit + p.gn_gs - 1) / p.gn_gs;
        if (et < 2 * ng) {
          unsigned long long* slot = gn_acc + gpar * (kGnMaxLocal * 2) + et;
          const unsigned long long v = *slot;
          *slot = 0ull;
          if (tile_ok && v != 0ull) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.gn_stats) +
                                      (static_cast<size_t>(b) * p.gn_groups + n0 / p.gn_gs) * 2 + et;
            atomicAdd(dst, v);
          }
        }
        gpar ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
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
  constexpr int smem = STAGES * (kABytes + (BLOCK_N / CLUSTER) * kBlockK * 2) + 1024 + 256 + kGnSmemBytes;
  static PerDeviceInt per_sm_dev;
  int& per_sm = per_sm_dev.get();
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
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CLUSTER > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CLUSTER;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
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
  {
    const size_t yes = d->y_dtype == UC_F32 ? 4 : 2;
    p.wide_store = ((d->ldy * yes) % 32 == 0) && (reinterpret_cast<uintptr_t>(d->y) % 32 == 0);
    p.wide_res = d->res && ((d->ldres * es) % 32 == 0) && (reinterpret_cast<uintptr_t>(d->res) % 32 == 0);
  }
  p.Cout = d->Cout;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("UC_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.idesc = umma_idesc_f16(d->x_dtype == UC_BF16 ? 1u : 0u, kBlockM, static_cast<uint32_t>(bn));
  p.bias = d->bias; p.gamma = d->gamma; p.res = d->res; p.ldres = d->ldres;
  p.y = d->y; p.ldy = d->ldy; p.y_dtype = d->y_dtype; p.act = d->act;
  p.row_stats = static_cast<const long long*>(d->row_stats); p.col_s = d->col_s; p.row_inv = 1.f / (kGnFixedScale * static_cast<float>(d->Cin)); p.row_eps = d->row_eps;
  if (d->row_stats && (!d->col_s || d->KH != 1 || d->KW != 1 || d->stride != 1 || d->pad != 0))
    return set_error(UC_EINVAL, "uc_conv2d: row_stats (folded LayerNorm) needs a 1x1 stride-1 conv and col_s");
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
      case 256: return launch_conv<256, 6, 2>(p, stream);  // 6 x 32 KB ring
      case 192: return launch_conv<192, 7, 2>(p, stream);  // 7 x 28 KB
      case 128: return launch_conv<128, 8, 2>(p, stream);  // 8 x 24 KB
      default: return set_error(UC_EINVAL, "uc_conv2d: the cta_group::2 variant exists for block_n 128/192/256 only");
    }
  }
  switch (bn) {
    case 256: return launch_conv<256, 4, 1>(p, stream);
    case 192: return launch_conv<192, 5, 1>(p, stream);
    case 128: return launch_conv<128, 6, 1>(p, stream);
    case 96: return launch_conv<96, 6, 1>(p, stream);
    case 64: return launch_conv<64, 8, 1>(p, stream);
    case 32: return launch_conv<32, 8, 1>(p, stream);
    case 16: return launch_conv<16, 8, 1>(p, stream);
    default: return set_error(UC_EINVAL, "uc_conv2d: unsupported block_n %d", bn);
  }
}
