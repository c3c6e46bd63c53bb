// Fused reference<->current embedding correlation + softmax over reference positions + label propagation.
//
//   out[o, j] = sum_i V[o, i] * softmax_i( <K_i, Q_j> )        K = reference embedding, Q = current embedding
//
// replaces the three materialising passes of external/lib/test/tracker/unicorn_sot.py:95-100
// (torch.mm -> softmax(dim=0) -> values @ trans_mat; same in unicorn_vos.py:171-181): the (N_ref x N_cur) similarity
// matrix (512 MB in fp16 at 800x1280) never leaves the SM.  Flash-attention style: one CTA owns 128 current positions
// (rows of the TMEM accumulator), streams the reference positions in chunks of 128 through a TMA ring, computes the
// 128x128 similarity tile with tcgen05.mma into a double-buffered TMEM accumulator, and four softmax warps (one
// thread per current position, no cross-thread reductions) keep the running max / sum / weighted label sums.
// V has only n_obj (1..8) rows, so the P.V product is done with CUDA-core FMAs on the probabilities instead of
// wasting an MMA tile.
#include "uc_ptx.cuh"
#include "uc_common.h"
#include "../../include/unicorn_b200.h"

namespace uc {

constexpr int kCorrC = 128;      // embedding channels
constexpr int kCorrTile = 128;   // current positions per CTA == reference positions per chunk
constexpr int kCorrStages = 4;
constexpr int kCorrTileBytes = kCorrTile * kCorrC * 2;  // 32 KB (two 128B-swizzled 64-channel halves)
constexpr int kCorrSoftmaxWarps = 16;                  // 4 per TMEM lane quadrant, each owning 32 of the 128 chunk columns
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

template <int NOBJ>
__global__ void __launch_bounds__(kCorrThreads, 1) corr_kernel(const __grid_constant__ CorrParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kCorrTileBytes;
  float* sV = reinterpret_cast<float*>(sK + kCorrStages * kCorrTileBytes);  // [2][NOBJ][128]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sV + 2 * NOBJ * kCorrTile);
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + kCorrStages;
  uint64_t* s_full = k_empty + kCorrStages;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * kCorrTile;
  const int nchunks = (p.n_ref + kCorrTile - 1) / kCorrTile;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmQ);
    prefetch_tmap(&p.tmK);
    mbar_init(q_full, 1);
    for (int i = 0; i < kCorrStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kCorrSoftmaxWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // barrier init / TMEM allocation above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues (keeps addresses in uniform registers)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kCorrTileBytes);
      tma_load_2d(sQ, &p.tmQ, q_full, 0, j0);
      tma_load_2d(sQ + kCorrTileBytes / 2, &p.tmQ, q_full, 64, j0);
    }
    __syncwarp();
    int stage = 0, phase = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&k_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[stage], kCorrTileBytes);
        uint8_t* dst = sK + stage * kCorrTileBytes;
        tma_load_2d(dst, &p.tmK, &k_full[stage], 0, c * kCorrTile);
        tma_load_2d(dst + kCorrTileBytes / 2, &p.tmK, &k_full[stage], 64, c * kCorrTile);
      }
      __syncwarp();
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
        const uint64_t k_desc = k_desc0 + static_cast<uint64_t>((stage * kCorrTileBytes) >> 4);
#pragma unroll
        for (int ks = 0; ks < kCorrC / 16; ++ks) {
          const uint64_t o = static_cast<uint64_t>(((ks >> 2) * (kCorrTileBytes / 2) + (ks & 3) * 32) >> 4);
          umma_f16(tmem_base + buf * kCorrTile, q_desc + o, k_desc + o, idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[stage]);
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
      if (++stage == kCorrStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ---------------- online softmax + label propagation
    // 16 warps: warp w owns TMEM lane quadrant (w & 3) (= 32 current positions) and columns [32*cg, 32*cg+32) of every
    // 128-column similarity chunk, cg = (w-2)/4.  Each thread keeps a private running (max, sum, weighted label sums)
    // for its (position, column group); the four partial states of a position are merged once at the end.  Four warps
    // per scheduler hide the tcgen05.ld / MUFU latencies that a single warp per scheduler could not.
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int tid = threadIdx.x - 64;  // 0..511 among the softmax threads
    const int j = j0 + row;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    float acc[NOBJ];
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) acc[o] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      const int i0 = c * kCorrTile;
      float* vb = sV + buf * NOBJ * kCorrTile;
      if (tid < kCorrTile) {
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
          float v = 0.f;
          if (o < p.n_obj && i0 + tid < p.n_ref) v = __ldg(p.V + static_cast<long>(o) * p.ldv + i0 + tid);
          vb[o * kCorrTile + tid] = v;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kCorrSoftmaxWarps * 32) : "memory");
      mbar_wait(&s_full[buf], (c >> 1) & 1);
      tc_fence_after();
      const int nvalid = min(kCorrTile, p.n_ref - i0);
      const int c0 = cg * 32;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kCorrTile + c0, v);
      tmem_ld_wait();
      // this warp's only read of the S buffer is done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[buf]);
      if (c0 >= nvalid) continue;  // warp-uniform: fully masked column group of the tail chunk
      float s[32];
      float cmax = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; ++t) s[t] = __uint_as_float(v[t]) * kLog2e;
      if (nvalid < kCorrTile) {  // tail chunk only (warp-uniform)
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (c0 + t >= nvalid) s[t] = -INFINITY;
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) cmax = fmaxf(cmax, s[t]);
      const float m_new = fmaxf(m, cmax);
      const float scale = fast_exp2(m - m_new);
      m = m_new;
      l *= scale;
#pragma unroll
      for (int o = 0; o < NOBJ; ++o) acc[o] *= scale;
#pragma unroll
      for (int t = 0; t < 32; t += 4) {
        float pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { pr[u] = fast_exp2(s[t + u] - m_new); l += pr[u]; }
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
          const float4 vv = *reinterpret_cast<const float4*>(vb + o * kCorrTile + c0 + t);
          acc[o] = fmaf(pr[0], vv.x, acc[o]); acc[o] = fmaf(pr[1], vv.y, acc[o]);
          acc[o] = fmaf(pr[2], vv.z, acc[o]); acc[o] = fmaf(pr[3], vv.w, acc[o]);
        }
      }
    }
    // merge the four column groups of every position (the K ring is idle now: reuse it as scratch)
    float* part = reinterpret_cast<float*>(sK);  // [4][128][2 + NOBJ]
    float* mine = part + (cg * kCorrTile + row) * (2 + NOBJ);
    mine[0] = m; mine[1] = l;
#pragma unroll
    for (int o = 0; o < NOBJ; ++o) mine[2 + o] = acc[o];
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
    tmem_dealloc(tmem_base, 256);
  }
}

template <int NOBJ>
static int launch_corr(const CorrParams& p, int grid, cudaStream_t stream) {
  constexpr int smem = (1 + kCorrStages) * kCorrTileBytes + 2 * NOBJ * kCorrTile * 4 + 256 + 1024;
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
    uint32_t box[2] = {64, kCorrTile};
    rc = encode_tmap(&p.tmK, dt, 2, embed_ref, dims, strides, box);
    if (rc) return rc;
  }
  p.V = values; p.out = out; p.ldv = ldv; p.ldo = ldo; p.n_cur = n_cur; p.n_ref = n_ref; p.n_obj = n_obj;
  p.idesc = umma_idesc_f16(dtype == UC_BF16 ? 1u : 0u, kCorrTile, kCorrTile);
  const int grid = (n_cur + kCorrTile - 1) / kCorrTile;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (n_obj == 1) return launch_corr<1>(p, grid, stream);
  if (n_obj == 2) return launch_corr<2>(p, grid, stream);
  if (n_obj <= 4) return launch_corr<4>(p, grid, stream);
  return launch_corr<8>(p, grid, stream);
}
