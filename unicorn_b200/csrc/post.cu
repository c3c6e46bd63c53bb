// Detection post-processing on the device: head decode, score filter, sort, class-aware NMS.
//   uc_head_decode   unicorn_head.py:332-334 (cat[reg, sigmoid(obj), sigmoid(cls)]) + decode_outputs :467-482
//   uc_postprocess   unicorn/utils/boxes.py:33-77: cxcywh->xyxy, class_conf/pred = max/argmax over classes,
//                    keep obj*class_conf >= conf, torchvision.ops.batched_nms (greedy, IoU > thr suppresses,
//                    only within the same class), result ordered by descending score.
// Everything stays on the GPU; the host reads back one counter.  Decision arithmetic is fp32 in the reference's
// operation order so that thresholds flip only on exact ties.
#include "uc_common.h"
#include "../../include/unicorn_b200.h"
#include <algorithm>

namespace uc {

struct DecodeLevels {
  const float* regobj[3];  // [HW, ld_ro]: reg(4), obj logit
  const float* cls[3];     // [HW, ld_cls]: class logits
  int h[3], w[3], stride[3], start[3];
  int ld_ro, ld_cls, ncls, total;
};

__global__ void __launch_bounds__(256) head_decode_kernel(DecodeLevels lv, float* __restrict__ out) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lv.total) return;
  int k = 0;
  if (i >= lv.start[1]) k = 1;
  if (i >= lv.start[2]) k = 2;
  const int a = i - lv.start[k];
  const int x = a % lv.w[k], y = a / lv.w[k];
  const float s = static_cast<float>(lv.stride[k]);
  const float* ro = lv.regobj[k] + static_cast<long>(a) * lv.ld_ro;
  const float* cl = lv.cls[k] + static_cast<long>(a) * lv.ld_cls;
  float* o = out + static_cast<long>(i) * (5 + lv.ncls);
  o[0] = (ro[0] + x) * s;
  o[1] = (ro[1] + y) * s;
  o[2] = expf(ro[2]) * s;
  o[3] = expf(ro[3]) * s;
  o[4] = 1.f / (1.f + expf(-ro[4]));
  for (int c = 0; c < lv.ncls; ++c) o[5 + c] = 1.f / (1.f + expf(-cl[c]));
}

// ---- filter: deterministic compaction (anchor order) of candidates with obj*class_conf >= conf.
// det rows: x1,y1,x2,y2,obj,class_conf,class_pred ; key = (score bits << 32) | (0xffffffff - candidate index)
__global__ void __launch_bounds__(1024) det_filter_kernel(const float* __restrict__ pred, int A, int ncls, float conf,
                                                           float* __restrict__ det, unsigned long long* __restrict__ keys,
                                                           int* __restrict__ count, int cap, int* __restrict__ det_anchor) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  __shared__ int warp_cnt[32];
  __shared__ int warp_excl[32];
  __shared__ int base, round_total;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int a0 = 0; a0 < A; a0 += 1024) {
    const int a = a0 + threadIdx.x;
    bool pass = false;
    float r[7];
    float score = 0.f;
    if (a < A) {
      const float* p = pred + static_cast<long>(a) * (5 + ncls);
      float best = p[5];
      int bi = 0;
      for (int c = 1; c < ncls; ++c) {
        const float v = p[5 + c];
        if (v > best) { best = v; bi = c; }
      }
      score = p[4] * best;
      pass = score >= conf;
      r[0] = p[0] - p[2] / 2; r[1] = p[1] - p[3] / 2; r[2] = p[0] + p[2] / 2; r[3] = p[1] + p[3] / 2;
      r[4] = p[4]; r[5] = best; r[6] = static_cast<float>(bi);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      const int v = warp_cnt[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
      }
      warp_excl[lane] = incl - v;
      if (lane == 31) round_total = incl;
    }
    __syncthreads();
    if (pass) {
      const int idx = base + warp_excl[warp] + __popc(bal & ((1u << lane) - 1));
      if (idx < cap) {
        float* d = det + static_cast<long>(idx) * 7;
#pragma unroll
        for (int t = 0; t < 7; ++t) d[t] = r[t];
        keys[idx] = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) | (0xffffffffu - static_cast<unsigned>(idx));
        det_anchor[idx] = a;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base += round_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = min(base, cap);
}

// ---- sort keys descending (bitonic, one CTA, n2 = power of two >= count; pads with 0 keys)
__global__ void __launch_bounds__(1024) sort_desc_kernel(unsigned long long* __restrict__ keys, const int* __restrict__ count, int cap2) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int n = *count;
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  if (n2 > cap2) n2 = cap2;
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int i = ((t / j) * (j << 1)) + (t % j);
        const int ixj = i + j;
        const unsigned long long a = keys[i], b = keys[ixj];
        const bool desc = ((i & k) == 0);
        if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

// ---- gather rows in sorted order
__global__ void __launch_bounds__(256) det_gather_kernel(const float* __restrict__ det, const unsigned long long* __restrict__ keys,
                                                          const int* __restrict__ count, float* __restrict__ sorted,
                                                          const int* __restrict__ det_anchor, int* __restrict__ sorted_anchor) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned idx = 0xffffffffu - static_cast<unsigned>(keys[i] & 0xffffffffull);
    sorted_anchor[i] = det_anchor[idx];
#pragma unroll
    for (int t = 0; t < 7; ++t) sorted[static_cast<long>(i) * 7 + t] = det[static_cast<long>(idx) * 7 + t];
  }
}

// ---- greedy NMS without an N x N matrix (one CTA).  Candidates are visited in score order in chunks of 256:
//  1. every candidate of the chunk is tested against the boxes kept so far (4 threads per candidate split the list);
//  2. the survivors of the chunk are tested against each other (256 x 256 bits in shared memory);
//  3. one warp resolves the chunk greedily, jumping from survivor to survivor with ffs;
//  4. the newly kept boxes are appended to the kept list (shared memory, spilling to the output rows in global).
// Work ~ N x kept IoUs instead of N^2, and the sequential part is proportional to the number of kept boxes.
// IoU arithmetic is torchvision's devIoU (fp32, inter / (areaA + areaB - inter) > thr), same-class pairs only.
constexpr int kNmsChunk = 256;
constexpr int kNmsKeepSmem = 3072;

__device__ __forceinline__ bool nms_hit(const float4 a, const float4 b, float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(xx2 - xx1, 0.f), h = fmaxf(yy2 - yy1, 0.f);
  const float inter = w * h;
  const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
  return inter / (sa + sb - inter) > thr;
}

__global__ void __launch_bounds__(1024) nms_greedy_kernel(const float* __restrict__ sorted, const int* __restrict__ count, float thr,
                                                           float* __restrict__ out, int* __restrict__ out_count, int max_keep,
                                                           const int* __restrict__ sorted_anchor, int* __restrict__ out_anchor) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  extern __shared__ float4 kept_box[];                       // [kNmsKeepSmem]
  float* kept_cls = reinterpret_cast<float*>(kept_box + kNmsKeepSmem);  // [kNmsKeepSmem]
  __shared__ float4 cbox[kNmsChunk];
  __shared__ float ccls[kNmsChunk];
  __shared__ unsigned long long pair_mask[kNmsChunk][4];
  __shared__ unsigned long long alive_w[4], kept_w[4];
  __shared__ int nk_s;
  const int n = *count;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) nk_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += kNmsChunk) {
    const int nk = nk_s;
    if (nk >= max_keep) break;  // uniform: the first max_keep rows of the full result are already final
    const int ci = tid >> 2, sub = tid & 3;  // candidate within chunk, quarter of the kept list
    const int j = c0 + ci;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float cl = -1.f;
    if (j < n) {
      const float* d = sorted + static_cast<long>(j) * 7;
      bx = make_float4(d[0], d[1], d[2], d[3]);
      cl = d[6];
    }
    bool sup = (j >= n);
    for (int k = sub; k < nk && !sup; k += 4) {
      float4 kb;
      float kc;
      if (k < kNmsKeepSmem) { kb = kept_box[k]; kc = kept_cls[k]; }
      else { const float* d = out + static_cast<long>(k) * 7; kb = make_float4(d[0], d[1], d[2], d[3]); kc = d[6]; }
      if (kc == cl && nms_hit(kb, bx, thr)) sup = true;
    }
    sup |= __shfl_xor_sync(0xffffffffu, sup, 1) != 0;
    sup |= __shfl_xor_sync(0xffffffffu, sup, 2) != 0;
    if (sub == 0) { cbox[ci] = bx; ccls[ci] = cl; }
    // alive words: ballot over lanes with sub == 0 (8 candidates per warp)
    const unsigned bal = __ballot_sync(0xffffffffu, !sup && sub == 0);
    if (tid < 4) alive_w[tid] = 0ull;
    __syncthreads();
    if (lane == 0) {
      // compress ballot bits (every 4th lane) into 8 bits at position (warp*8)
      unsigned v = 0;
      for (int t = 0; t < 8; ++t) v |= ((bal >> (4 * t)) & 1u) << t;
      const int cbase = (tid >> 5) * 8;  // first candidate of this warp
      atomicOr(&alive_w[cbase >> 6], static_cast<unsigned long long>(v) << (cbase & 63));
    }
    __syncthreads();
    {  // pairwise bits inside the chunk: candidate ci vs candidates [sub*64, sub*64+64), later ones only
      unsigned long long bits = 0ull;
      const bool me = (alive_w[ci >> 6] >> (ci & 63)) & 1ull;
      if (me) {
        const unsigned long long aw = alive_w[sub];
        for (int t = 0; t < 64; ++t) {
          const int o = sub * 64 + t;
          if (o > ci && ((aw >> t) & 1ull) && ccls[o] == cl && nms_hit(bx, cbox[o], thr)) bits |= 1ull << t;
        }
      }
      pair_mask[ci][sub] = bits;
    }
    __syncthreads();
    if (tid < 32) {  // greedy resolution of the chunk (all lanes run the same scalar code)
      unsigned long long removed[4] = {0ull, 0ull, 0ull, 0ull};
      unsigned long long kept[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        unsigned long long cur = ~alive_w[w] | removed[w];
        while (~cur != 0ull) {
          const int t = __ffsll(static_cast<long long>(~cur)) - 1;
          kept[w] |= 1ull << t;
          const int r = w * 64 + t;
          cur |= pair_mask[r][w] | (1ull << t);
#pragma unroll
          for (int w2 = w + 1; w2 < 4; ++w2) removed[w2] |= pair_mask[r][w2];
        }
      }
      if (tid == 0) { kept_w[0] = kept[0]; kept_w[1] = kept[1]; kept_w[2] = kept[2]; kept_w[3] = kept[3]; }
    }
    __syncthreads();
    if (tid < kNmsChunk) {
      const int w = tid >> 6, t = tid & 63;
      if ((kept_w[w] >> t) & 1ull) {
        int pos = nk + __popcll(kept_w[w] & ((1ull << t) - 1ull));
        for (int w2 = 0; w2 < w; ++w2) pos += __popcll(kept_w[w2]);
        if (pos < kNmsKeepSmem) { kept_box[pos] = cbox[tid]; kept_cls[pos] = ccls[tid]; }
        const float* d = sorted + static_cast<long>(c0 + tid) * 7;
        float* o = out + static_cast<long>(pos) * 7;
#pragma unroll
        for (int q = 0; q < 7; ++q) o[q] = d[q];
        if (out_anchor) out_anchor[pos] = sorted_anchor[c0 + tid];
      }
    }
    __syncthreads();
    if (tid == 0) nk_s = nk + __popcll(kept_w[0]) + __popcll(kept_w[1]) + __popcll(kept_w[2]) + __popcll(kept_w[3]);
    __syncthreads();
  }
  if (tid == 0) *out_count = min(nk_s, max_keep);
}

}  // namespace uc

using namespace uc;

extern "C" int uc_head_decode(const float* const* regobj, const float* const* cls, const int* hw, const int* strides, int ld_ro,
                              int ld_cls, int ncls, float* out, void* stream_v) {
  if (!regobj || !cls || !hw || !strides || !out || ncls < 1 || ncls > ld_cls || ld_ro < 5) return set_error(UC_EINVAL, "uc_head_decode: bad arguments");
  DecodeLevels lv;
  int start = 0;
  for (int k = 0; k < 3; ++k) {
    lv.regobj[k] = regobj[k]; lv.cls[k] = cls[k];
    lv.h[k] = hw[2 * k]; lv.w[k] = hw[2 * k + 1]; lv.stride[k] = strides[k]; lv.start[k] = start;
    start += lv.h[k] * lv.w[k];
  }
  lv.ld_ro = ld_ro; lv.ld_cls = ld_cls; lv.ncls = ncls; lv.total = start;
  launch_pdl(head_decode_kernel, (start + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_v), lv, out);
  return check_launch("uc_head_decode");
}

extern "C" long uc_postprocess_workspace_bytes(int max_anchors) {
  const long A = max_anchors;
  long a2 = 1;
  while (a2 < A) a2 <<= 1;
  return A * 7 * 4 * 2 + a2 * 8 + A * 4 * 2 + 256;
}

extern "C" int uc_postprocess(const float* pred, int A, int ncls, float conf_thre, float nms_thre, int max_keep, void* workspace,
                              long workspace_bytes, float* out_dets, int* out_count, int* out_anchor, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!pred || !workspace || !out_dets || !out_count || A < 1 || ncls < 1) return set_error(UC_EINVAL, "uc_postprocess: bad arguments");
  if (workspace_bytes < uc_postprocess_workspace_bytes(A)) return set_error(UC_EINVAL, "uc_postprocess: workspace too small");
  long a2 = 1;
  while (a2 < A) a2 <<= 1;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* count = reinterpret_cast<int*>(ws);
  float* det = reinterpret_cast<float*>(ws + 256);
  float* sorted = det + static_cast<long>(A) * 7;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sorted + static_cast<long>(A) * 7);
  int* det_anchor = reinterpret_cast<int*>(keys + a2);
  int* sorted_anchor = det_anchor + A;
  launch_pdl(det_filter_kernel, 1, 1024, 0, stream, pred, A, ncls, conf_thre, det, keys, count, A, det_anchor);
  launch_pdl(sort_desc_kernel, 1, 1024, 0, stream, keys, count, static_cast<int>(a2));
  launch_pdl(det_gather_kernel, std::min(num_sms() * 4, (A + 255) / 256), 256, 0, stream, det, keys, count, sorted, det_anchor, sorted_anchor);
  constexpr int smem = kNmsKeepSmem * (16 + 4);
  static PerDeviceFlag attr_dev;
  bool& attr = attr_dev.get();
  if (!attr) {
    cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  launch_pdl(nms_greedy_kernel, 1, 1024, smem, stream, sorted, count, nms_thre, out_dets, out_count, max_keep > 0 ? max_keep : 0x7fffffff,
                                               sorted_anchor, out_anchor);
  return check_launch("uc_postprocess");
}
