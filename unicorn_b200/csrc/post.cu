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
                                                           int* __restrict__ count, int cap) {
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
                                                          const int* __restrict__ count, float* __restrict__ sorted) {
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned idx = 0xffffffffu - static_cast<unsigned>(keys[i] & 0xffffffffull);
#pragma unroll
    for (int t = 0; t < 7; ++t) sorted[static_cast<long>(i) * 7 + t] = det[static_cast<long>(idx) * 7 + t];
  }
}

// ---- suppression bit matrix: mask[i][w] bit j set iff j=64w+bit > i, same class, IoU(i,j) > thr
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ sorted, const int* __restrict__ count, float thr,
                                                       unsigned long long* __restrict__ mask, int words_cap) {
  const int n = *count;
  const int nb = (n + 63) / 64;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb >= nb || cb >= nb || cb < rb) return;
  __shared__ float cbx[64][5];
  const int cj = cb * 64 + threadIdx.x;
  if (cj < n) {
    const float* d = sorted + static_cast<long>(cj) * 7;
    cbx[threadIdx.x][0] = d[0]; cbx[threadIdx.x][1] = d[1]; cbx[threadIdx.x][2] = d[2]; cbx[threadIdx.x][3] = d[3];
    cbx[threadIdx.x][4] = d[6];
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  const float* d = sorted + static_cast<long>(i) * 7;
  const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3], cls = d[6];
  const float area = (x2 - x1) * (y2 - y1);
  unsigned long long bits = 0ull;
  const int lim = min(64, n - cb * 64);
  for (int t = (rb == cb ? threadIdx.x + 1 : 0); t < lim; ++t) {
    if (cbx[t][4] != cls) continue;
    const float xx1 = fmaxf(x1, cbx[t][0]), yy1 = fmaxf(y1, cbx[t][1]);
    const float xx2 = fminf(x2, cbx[t][2]), yy2 = fminf(y2, cbx[t][3]);
    const float w = fmaxf(xx2 - xx1, 0.f), h = fmaxf(yy2 - yy1, 0.f);
    const float inter = w * h;
    const float areab = (cbx[t][2] - cbx[t][0]) * (cbx[t][3] - cbx[t][1]);
    const float iou = inter / (area + areab - inter);
    if (iou > thr) bits |= 1ull << t;
  }
  mask[static_cast<long>(i) * words_cap + cb] = bits;
}

// ---- greedy scan over the sorted boxes (one CTA).  Work is proportional to the number of KEPT boxes:
// inside a 64-box block the next survivor is found with ffs on the running removed-word.
__global__ void __launch_bounds__(1024) nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ count,
                                                         int words_cap, const float* __restrict__ sorted, float* __restrict__ out,
                                                         int* __restrict__ out_count) {
  extern __shared__ unsigned long long removed[];  // [words]
  __shared__ unsigned long long kept_bits;
  __shared__ int n_out;
  const int n = *count;
  const int nb = (n + 63) / 64;
  for (int w = threadIdx.x; w < nb; w += blockDim.x) removed[w] = 0ull;
  if (threadIdx.x == 0) n_out = 0;
  __syncthreads();
  for (int b = 0; b < nb; ++b) {
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int i0 = b * 64 + lane, i1 = i0 + 32;
      const unsigned long long d0 = i0 < n ? mask[static_cast<long>(i0) * words_cap + b] : 0ull;
      const unsigned long long d1 = i1 < n ? mask[static_cast<long>(i1) * words_cap + b] : 0ull;
      unsigned long long cur = removed[b];
      if (n - b * 64 < 64) cur |= ~0ull << (n - b * 64);
      unsigned long long kept = 0ull;
      while (~cur != 0ull) {
        const int t = __ffsll(static_cast<long long>(~cur)) - 1;
        kept |= 1ull << t;
        const unsigned long long dl = __shfl_sync(0xffffffffu, t < 32 ? d0 : d1, t & 31);
        cur |= dl | (1ull << t);
      }
      if (lane == 0) kept_bits = kept;
    }
    __syncthreads();
    unsigned long long kb = kept_bits;
    const int base_out = n_out;
    const int nk = __popcll(kb);
    // OR the kept rows into the removed vector (words > b) and emit the kept rows
    int ord = 0;
    while (kb) {
      const int t = __ffsll(static_cast<long long>(kb)) - 1;
      kb &= kb - 1;
      const long i = static_cast<long>(b) * 64 + t;
      for (int w = b + 1 + threadIdx.x; w < nb; w += blockDim.x) {
        const unsigned long long mw = mask[i * words_cap + w];
        if (mw) removed[w] |= mw;
      }
      if (threadIdx.x < 7) out[static_cast<long>(base_out + ord) * 7 + threadIdx.x] = sorted[i * 7 + threadIdx.x];
      ++ord;
    }
    __syncthreads();
    if (threadIdx.x == 0) n_out = base_out + nk;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_count = n_out;
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
  head_decode_kernel<<<(start + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(lv, out);
  return check_launch("uc_head_decode");
}

extern "C" long uc_postprocess_workspace_bytes(int max_anchors) {
  const long A = max_anchors;
  long a2 = 1;
  while (a2 < A) a2 <<= 1;
  const long words = (A + 63) / 64;
  return A * 7 * 4 * 2 + a2 * 8 + A * words * 8 + 256;
}

extern "C" int uc_postprocess(const float* pred, int A, int ncls, float conf_thre, float nms_thre, void* workspace,
                              long workspace_bytes, float* out_dets, int* out_count, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!pred || !workspace || !out_dets || !out_count || A < 1 || ncls < 1) return set_error(UC_EINVAL, "uc_postprocess: bad arguments");
  if (workspace_bytes < uc_postprocess_workspace_bytes(A)) return set_error(UC_EINVAL, "uc_postprocess: workspace too small");
  long a2 = 1;
  while (a2 < A) a2 <<= 1;
  const int words = (A + 63) / 64;
  if (words * 8 > 200 * 1024) return set_error(UC_EINVAL, "uc_postprocess: too many anchors (%d)", A);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* count = reinterpret_cast<int*>(ws);
  float* det = reinterpret_cast<float*>(ws + 256);
  float* sorted = det + static_cast<long>(A) * 7;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sorted + static_cast<long>(A) * 7);
  unsigned long long* mask = keys + a2;
  det_filter_kernel<<<1, 1024, 0, stream>>>(pred, A, ncls, conf_thre, det, keys, count, A);
  sort_desc_kernel<<<1, 1024, 0, stream>>>(keys, count, static_cast<int>(a2));
  det_gather_kernel<<<std::min(num_sms() * 4, (A + 255) / 256), 256, 0, stream>>>(det, keys, count, sorted);
  dim3 g(static_cast<unsigned>(words), static_cast<unsigned>(words));
  nms_mask_kernel<<<g, 64, 0, stream>>>(sorted, count, nms_thre, mask, words);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  nms_scan_kernel<<<1, 1024, static_cast<size_t>(words) * 8, stream>>>(mask, count, words, sorted, out_dets, out_count);
  return check_launch("uc_postprocess");
}
