// Fused reference<->current embedding correlation + softmax over reference positions + label propagation.
//
//   out[o, j] = sum_i V[o, i] * softmax_i( <K_i, Q_j> )        K = reference embedding, Q = current embedding
//
// replaces the three materialising passes of external/lib/test/tracker/unicorn_sot.py:95-100
// (torch.mm -> softmax(dim=0) -> values @ trans_mat; same in unicorn_vos.py:171-181): the (N_ref x N_cur) similarity
// matrix (512 MB in fp16 at 800x1280) never leaves the SM.  Flash-attention style: one CTA owns 128 current positions
// (rows of the TMEM accumulator), streams the reference positions in chunks of 256 through a TMA ring, computes the
// 128x256 similarity tile with tcgen05.mma (UMMA 128x256x16) into a double-buffered TMEM accumulator (all 512 columns), and
// sixteen softmax warps (four per TMEM lane quadrant, 64 columns each, no cross-thread reductions inside the loop) keep the
// running max / sum / weighted label sums; one exponential in four is evaluated on the FMA pipe (the kernel is MUFU bound).
// V has only n_obj (1..8) rows, so the P.V product is done with CUDA-core FMAs on the probabilities instead of
// wasting an MMA tile.
#include "uc_ptx.cuh"
#include "uc_common.h"
#include "../../include/unicorn_b200.h"

namespace uc {

constexpr int kCorrC = 128;       // embedding channels
constexpr int kCorrTile = 128;    // current positions per CTA (rows of the TMEM accumulator)
constexpr int kCorrChunk = 256;   // reference positions per MMA chunk (UMMA N = 256: half as many barrier hand-offs as 128)
constexpr int kCorrStages = 2;    // K ring (64 KB per stage)
constexpr int kCorrVSlots = 6;    // label-value slots, see the producer: chunk x's values may be written once every softmax warp has LOADED
                                  // S of chunk x-4 (it may still be computing chunk x-4, but is done with x-5): >= 5 slots are needed
constexpr int kCorrQBytes = kCorrTile * kCorrC * 2;    // 32 KB (two 128B-swizzled 64-channel halves)
constexpr int kCorrKBytes = kCorrChunk * kCorrC * 2;   // 64 KB
constexpr int kCorrSoftmaxWarps = 16;                  // 4 per TMEM lane quadrant, each owning 64 of the 256 chunk columns
constexpr int kCorrThreads = (2 + kCorrSoftmaxWarps) * 32;

struct alignas(64) CorrParams {
  CUtensorMap tmQ, tmK;
  const float* V;  // [n_obj, ldv]
  float* out;      // [n_obj, ldo]
  int ldv, ldo, n_cur, n_ref, n_obj;
  uint32_t idesc;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, 2^f by a degree-4 minimax polynomial
// (max relative error 2.9e-6 — below the bf16/fp16 rounding of the similarity itself), 2^n by an exponent-field add.  One element in
// four takes this path: the kernel is bound by the MUFU unit (16 ex2 per clock per SM), the FMA pipe has room for ~25 % of them.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;              // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);        // in [-0.5, 0.5]
  float p = fmaf(f, 9.582853e-3f, 5.5906426e-2f);  // weighted least-squares (near-minimax) fit on [-0.5, 0.5] with p(0) = 1
  p = fmaf(p, f, 2.4024099e-1f);
  p = fmaf(p, f, 6.9312418e-1f);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int NOBJ>
__global__ void __launch_bounds__(kCorrThreads, 1) corr_kernel(const __grid_constant__ CorrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kCorrQBytes;
  float* sV = reinterpret_cast<float*>(sK + kCorrStages * kCorrKBytes);  // [kCorrVSlots][NOBJ][256]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sV + kCorrVSlots * NOBJ * kCorrChunk);
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + kCorrStages;
  uint64_t* s_full = k_empty + kCorrStages;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * kCorrTile;
  const int nchunks = (p.n_ref + kCorrChunk - 1) / kCorrChunk;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmQ);
    prefetch_tmap(&p.tmK);
    mbar_init(q_full, 1);
    for (int i = 0; i < kCorrStages; ++i) { mbar_init(&k_full[i], 2); mbar_init(&k_empty[i], 1); }  // full: TMA bytes + label values
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kCorrSoftmaxWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);  // two 128 x 256 fp32 similarity buffers
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // barrier init / TMEM allocation above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    // Producer: TMA for the K chunk (one elected lane) and — all 32 lanes — the chunk's label values V[:, i0 .. i0+255] into their
    // slot.  The producer runs up to kCorrStages chunks ahead of the MMA, so the L2 latency of these loads is never on the softmax
    // warps' critical path (it was: they used to stage V themselves behind a 512-thread barrier every chunk).
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kCorrQBytes);
      tma_load_2d(sQ, &p.tmQ, q_full, 0, j0);
      tma_load_2d(sQ + kCorrQBytes / 2, &p.tmQ, q_full, 64, j0);
    }
    __syncwarp();
    int stage = 0, phase = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&k_empty[stage], phase ^ 1);
      const int i0 = c * kCorrChunk;
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[stage], kCorrKBytes);
        uint8_t* dst = sK + stage * kCorrKBytes;
        tma_load_2d(dst, &p.tmK, &k_full[stage], 0, i0);
        tma_load_2d(dst + kCorrKBytes / 2, &p.tmK, &k_full[stage], 64, i0);
      }
      float* vb = sV + (c % kCorrVSlots) * NOBJ * kCorrChunk;
#pragma unroll
      for (int o = 0; o < NOBJ; ++o) {
#pragma unroll
        for (int t = 0; t < kCorrChunk / 32; ++t) {
          const int i = i0 + t * 32 + lane;
          vb[o * kCorrChunk + t * 32 + lane] = (o < p.n_obj && i < p.n_ref) ? __ldg(p.V + static_cast<long>(o) * p.ldv + i) : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_full[stage]);  // release: the MMA warp acquires it, its commit publishes it to the softmax warps
      if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues; descriptors built once, only the address field advances
    mbar_wait(q_full, 0);
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ)), k_desc0 = umma_desc_sw128(smem_u32(sK));
    const uint32_t idesc = p.idesc;
    int stage = 0, phase = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&s_empty[buf], ((c >> 1) & 1) ^ 1);
      mbar_wait(&k_full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t k_desc = k_desc0 + static_cast<uint64_t>((stage * kCorrKBytes) >> 4);
#pragma unroll
        for (int ks = 0; ks < kCorrC / 16; ++ks) {
          const uint64_t qo = static_cast<uint64_t>(((ks >> 2) * (kCorrQBytes / 2) + (ks & 3) * 32) >> 4);
          const uint64_t ko = static_cast<uint64_t>(((ks >> 2) * (kCorrKBytes / 2) + (ks & 3) * 32) >> 4);
          umma_f16(tmem_base + buf * kCorrChunk, q_desc + qo, k_desc + ko, idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[stage]);
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
      if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ---------------- online softmax + label propagation
    // 16 warps: warp w owns TMEM lane quadrant (w & 3) (= 32 current positions) and columns [64*cg, 64*cg+64) of every 256-column
    // similarity chunk, cg = (w-2)/4.  Each thread keeps a private running (max, sum, weighted label sums) for its (position,
    // column group); the four partial states of a position are merged once at the end.  No block-level barrier inside the loop: the
    // warps drift apart and hide one another's tcgen05.ld / MUFU latencies.  Everything is in the log2 domain: p = 2^(s*log2e - m).
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int j = j0 + row;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int CW = kCorrChunk / 4;  // 64 columns per warp and chunk
    float m = -INFINITY;
    float l[4] = {0.f, 0.f, 0.f, 0.f};
    float acc[NOBJ][2];
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) acc[o][0] = acc[o][1] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      const int i0 = c * kCorrChunk;
      const float* vb = sV + (c % kCorrVSlots) * NOBJ * kCorrChunk;
      mbar_wait(&s_full[buf], (c >> 1) & 1);
      tc_fence_after();
      const int nvalid = min(kCorrChunk, p.n_ref - i0);
      const int c0 = cg * CW;
      uint32_t v[CW];
      {
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kCorrChunk + c0;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(ta, v0);
        tmem_ld_32x32(ta + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) { v[t] = v0[t]; v[32 + t] = v1[t]; }
      }
      // this warp's only read of the S buffer is done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[buf]);
      if (c0 >= nvalid) continue;  // warp-uniform: fully masked column group of the tail chunk
      float s[CW];
#pragma unroll
      for (int t = 0; t < CW; ++t) s[t] = __uint_as_float(v[t]);
      if (nvalid < kCorrChunk) {  // tail chunk only (warp-uniform)
#pragma unroll
        for (int t = 0; t < CW; ++t)
          if (c0 + t >= nvalid) s[t] = -INFINITY;
      }
      float mx[8];  // tree maximum (a 64-deep dependent chain of FMNMX would sit on the critical path of every chunk)
#pragma unroll
      for (int t = 0; t < 8; ++t) mx[t] = fmaxf(fmaxf(s[t], s[t + 8]), fmaxf(s[t + 16], s[t + 24]));
#pragma unroll
      for (int t = 0; t < 8; ++t) mx[t] = fmaxf(mx[t], fmaxf(fmaxf(s[t + 32], s[t + 40]), fmaxf(s[t + 48], s[t + 56])));
      const float cmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m, cmax * kLog2e);
      if (m_new != m) {  // the running maximum moves in the first few chunks only: skip the rescale otherwise
        const float scale = fast_exp2(m - m_new);
        m = m_new;
#pragma unroll
        for (int u = 0; u < 4; ++u) l[u] *= scale;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) { acc[o][0] *= scale; acc[o][1] *= scale; }
      }
      const float neg_m = -m_new;
#pragma unroll
      for (int t = 0; t < CW; t += 4) {
        float pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float x = fmaf(s[t + u], kLog2e, neg_m);  // <= 0 (-inf for masked columns -> p = 0 on both paths)
          pr[u] = (u == 3) ? poly_exp2(x) : fast_exp2(x);
          l[u] += pr[u];
        }
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
          const float4 vv = *reinterpret_cast<const float4*>(vb + o * kCorrChunk + c0 + t);  // same address in every lane: broadcast
          acc[o][0] = fmaf(pr[0], vv.x, acc[o][0]); acc[o][1] = fmaf(pr[1], vv.y, acc[o][1]);
          acc[o][0] = fmaf(pr[2], vv.z, acc[o][0]); acc[o][1] = fmaf(pr[3], vv.w, acc[o][1]);
        }
      }
    }
    // merge the four column groups of every position (the K ring is idle now: reuse it as scratch)
    float* part = reinterpret_cast<float*>(sK);  // [4][128][2 + NOBJ]
    float* mine = part + (cg * kCorrTile + row) * (2 + NOBJ);
    mine[0] = m; mine[1] = (l[0] + l[1]) + (l[2] + l[3]);
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) mine[2 + o] = acc[o][0] + acc[o][1];
    asm volatile("bar.sync 1, %0;" ::"n"(kCorrSoftmaxWarps * 32) : "memory");
    if (cg == 0 && j < p.n_cur) {
      float M = -INFINITY;
#pragma unroll
      for (int g = 0; g < 4; ++g) M = fmaxf(M, part[(g * kCorrTile + row) * (2 + NOBJ)]);
      float L = 0.f, A[NOBJ];
#pragma unroll
      for (int o = 0; o < NOBJ; ++o) A[o] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float* pg = part + (g * kCorrTile + row) * (2 + NOBJ);
        const float w = fast_exp2(pg[0] - M);  // a group that saw only masked columns has m = -inf -> weight 0
        L = fmaf(pg[1], w, L);
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) A[o] = fmaf(pg[2 + o], w, A[o]);
      }
      const float inv = 1.f / L;
#pragma unroll
      for (int o = 0; o < NOBJ; ++o)
        if (o < p.n_obj) p.out[static_cast<long>(o) * p.ldo + j] = A[o] * inv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NOBJ>
static int launch_corr(const CorrParams& p, int grid, cudaStream_t stream) {
  constexpr int smem = kCorrQBytes + kCorrStages * kCorrKBytes + kCorrVSlots * NOBJ * kCorrChunk * 4 + 256 + 1024;
  static PerDeviceFlag attr_dev;
  bool& attr_set = attr_dev.get();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(corr_kernel<NOBJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "corr: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_pdl(corr_kernel<NOBJ>, grid, kCorrThreads, smem, stream, p);
  return check_launch("uc_corr_propagate");
}

}  // namespace uc

using namespace uc;

extern "C" int uc_corr_propagate(const void* embed_ref, int ld_ref, int n_ref, const void* embed_cur, int ld_cur, int n_cur,
                                 int C, int dtype, const float* values, int ldv, int n_obj, float* out, int ldo,
                                 void* stream_v) {
  if (!embed_ref || !embed_cur || !values || !out) return set_error(UC_EINVAL, "uc_corr_propagate: null pointer");
  if (C != kCorrC) return set_error(UC_EINVAL, "uc_corr_propagate: embedding dim must be %d (got %d)", kCorrC, C);
  if (dtype != UC_BF16 && dtype != UC_F16) return set_error(UC_EINVAL, "uc_corr_propagate: embeddings must be bf16/f16");
  if (n_obj < 1 || n_obj > 8) return set_error(UC_EINVAL, "uc_corr_propagate: 1 <= n_obj <= 8 (got %d)", n_obj);
  if (ld_ref % 8 || ld_cur % 8 || n_ref < 1 || n_cur < 1 || ldv < n_ref || ldo < n_cur) return set_error(UC_EINVAL, "uc_corr_propagate: bad sizes/strides");
  int rc = ensure_driver();
  if (rc) return rc;
  CorrParams p;
  memset(&p, 0, sizeof(p));
  const CUtensorMapDataType dt = dtype == UC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(n_cur)};
    uint64_t strides[1] = {static_cast<uint64_t>(ld_cur) * 2};
    uint32_t box[2] = {64, kCorrTile};
    rc = encode_tmap(&p.tmQ, dt, 2, embed_cur, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(n_ref)};
    uint64_t strides[1] = {static_cast<uint64_t>(ld_ref) * 2};
    uint32_t box[2] = {64, kCorrChunk};
    rc = encode_tmap(&p.tmK, dt, 2, embed_ref, dims, strides, box);
    if (rc) return rc;
  }
  p.V = values; p.out = out; p.ldv = ldv; p.ldo = ldo; p.n_cur = n_cur; p.n_ref = n_ref; p.n_obj = n_obj;
  p.idesc = umma_idesc_f16(dtype == UC_BF16 ? 1u : 0u, kCorrTile, kCorrChunk);
  const int grid = (n_cur + kCorrTile - 1) / kCorrTile;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (n_obj == 1) return launch_corr<1>(p, grid, stream);
  if (n_obj == 2) return launch_corr<2>(p, grid, stream);
  if (n_obj <= 4) return launch_corr<4>(p, grid, stream);
  return launch_corr<8>(p, grid, stream);
}
