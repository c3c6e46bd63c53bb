// CondInst mask path (config 4): aligned-bilinear fusion of the mask branch, per-instance dynamic convolutions,
// RAFT-style convex upsampling, sigmoid, final aligned-bilinear upsample.
//   uc_aligned_bilinear_add   condinst/comm.py:5-27 + mask_branch.py:81-96 (x = x + aligned_bilinear(x_p, f))
//   uc_dynamic_masks          condinst/dynamic_mask_head.py:61-87 (parameter split), :172-225 (rel-coords + 3 grouped
//                             1x1 convs 10->8->8->1), :159-170 (convex upsample x up_rate), :284 (sigmoid);
//                             utils/boxes.py:138-145 (aligned_bilinear x d_rate of the scores)
// HBM-bound on the output (N x H x W fp32 masks); all arithmetic fp32.
#include "uc_common.h"
#include "../../include/unicorn_b200.h"
#include <algorithm>
#include <cmath>

namespace uc {

// aligned_bilinear(t, f)[i] samples the (replicate-padded) source at max(i - f/2, 0) / f with align_corners=True.
__device__ __forceinline__ void ab_coord(int i, int f, int n, int& i0, int& i1, float& frac) {
  const int ii = max(i - f / 2, 0);
  i0 = ii / f;
  frac = static_cast<float>(ii - i0 * f) / f;
  i1 = min(i0 + 1, n - 1);
  i0 = min(i0, n - 1);
}

__global__ void __launch_bounds__(256) aligned_bilinear_add_kernel(const uint16_t* __restrict__ src, int lds, int hs, int ws,
                                                                    uint16_t* __restrict__ dst, int ldd, int C, int f) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int C2 = C >> 1, hd = hs * f, wd = ws * f;
  const long total = static_cast<long>(hd) * wd * C2;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C2) * 2;
    const long pix = i / C2;
    const int x = static_cast<int>(pix % wd), y = static_cast<int>(pix / wd);
    int y0, y1, x0, x1;
    float fy, fx;
    ab_coord(y, f, hs, y0, y1, fy);
    ab_coord(x, f, ws, x0, x1, fx);
    const uint32_t a = *reinterpret_cast<const uint32_t*>(src + (static_cast<long>(y0) * ws + x0) * lds + c);
    const uint32_t b = *reinterpret_cast<const uint32_t*>(src + (static_cast<long>(y0) * ws + x1) * lds + c);
    const uint32_t cc = *reinterpret_cast<const uint32_t*>(src + (static_cast<long>(y1) * ws + x0) * lds + c);
    const uint32_t d = *reinterpret_cast<const uint32_t*>(src + (static_cast<long>(y1) * ws + x1) * lds + c);
    uint32_t* o = reinterpret_cast<uint32_t*>(dst + pix * ldd + c);
    const uint32_t e = *o;
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    const float lo = bf16lo(e) + w00 * bf16lo(a) + w01 * bf16lo(b) + w10 * bf16lo(cc) + w11 * bf16lo(d);
    const float hi = bf16hi(e) + w00 * bf16hi(a) + w01 * bf16hi(b) + w10 * bf16hi(cc) + w11 * bf16hi(d);
    *o = pack_bf16(lo, hi);
  }
}

struct MaskLevels {
  const float* dyn[3];  // per level [h*w, ld_dyn] controller outputs
  int h[3], w[3], stride[3], start[3];
  float soi[3];
};

// logits[n, y, x] for instance n (anchor index from the NMS output) — one thread per (instance, pixel)
__global__ void __launch_bounds__(256) mask_logits_kernel(const float* __restrict__ mask_feats, int h, int w, MaskLevels lv, int ld_dyn,
                                                           const int* __restrict__ anchors, const int* __restrict__ count, int n_max,
                                                           float* __restrict__ logits) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  __shared__ float prm[169];
  __shared__ float inst[3];
  const int n = min(*count, n_max);
  const int ins = blockIdx.y;
  if (ins >= n) return;
  if (threadIdx.x < 169 || threadIdx.x == 255) {
    const int a = anchors[ins];
    int k = 0;
    if (a >= lv.start[1]) k = 1;
    if (a >= lv.start[2]) k = 2;
    const int ai = a - lv.start[k];
    if (threadIdx.x < 169) prm[threadIdx.x] = lv.dyn[k][static_cast<long>(ai) * ld_dyn + threadIdx.x];
    else {
      inst[0] = ((ai % lv.w[k]) + 0.5f) * lv.stride[k];  // locations = (grid + 0.5) * stride (unicorn_head_mask.py:518)
      inst[1] = ((ai / lv.w[k]) + 0.5f) * lv.stride[k];
      inst[2] = lv.soi[k];
    }
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  float in[10];
  in[0] = (inst[0] - ((pix % w) * 8 + 4)) / inst[2];  // compute_locations: stride 8, + stride // 2 (comm.py:30-45)
  in[1] = (inst[1] - ((pix / w) * 8 + 4)) / inst[2];
  const float4 f0 = *reinterpret_cast<const float4*>(mask_feats + static_cast<long>(pix) * 8);
  const float4 f1 = *reinterpret_cast<const float4*>(mask_feats + static_cast<long>(pix) * 8 + 4);
  in[2] = f0.x; in[3] = f0.y; in[4] = f0.z; in[5] = f0.w; in[6] = f1.x; in[7] = f1.y; in[8] = f1.z; in[9] = f1.w;
  float h1[8], h2[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float s = prm[152 + o];
#pragma unroll
    for (int i = 0; i < 10; ++i) s += prm[o * 10 + i] * in[i];
    h1[o] = fmaxf(s, 0.f);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float s = prm[160 + o];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += prm[80 + o * 8 + i] * h1[i];
    h2[o] = fmaxf(s, 0.f);
  }
  float s = prm[168];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += prm[144 + i] * h2[i];
  logits[static_cast<long>(ins) * h * w + pix] = s;
}

// convex upsampling x up (softmax over the 9 neighbours, weights from up_masks [h,w,9*up*up]) + sigmoid
__global__ void __launch_bounds__(256) mask_convex_up_kernel(const float* __restrict__ logits, const float* __restrict__ up_masks, int h,
                                                              int w, int up, const int* __restrict__ count, int n_max,
                                                              float* __restrict__ out) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int n = min(*count, n_max);
  const int ins = blockIdx.y;
  if (ins >= n) return;
  const int H = h * up, W = w * up;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * W) return;
  const int X = t % W, Y = t / W;
  const int x = X / up, y = Y / up, j = X % up, i = Y % up;
  const float* um = up_masks + (static_cast<long>(y) * w + x) * (9 * up * up) + i * up + j;
  float m[9], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) { m[k] = um[k * up * up]; mx = fmaxf(mx, m[k]); }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { m[k] = expf(m[k] - mx); den += m[k]; }
  const float* lg = logits + static_cast<long>(ins) * h * w;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? lg[yy * w + xx] : 0.f;
    acc += m[k] / den * v;
  }
  out[static_cast<long>(ins) * H * W + t] = 1.f / (1.f + expf(-acc));
}

__global__ void __launch_bounds__(256) mask_final_up_kernel(const float* __restrict__ src, int hs, int ws, int f,
                                                             const int* __restrict__ count, int n_max, float* __restrict__ out) {
  pdl_wait();               // programmatic dependent launch: global memory is touched only after the predecessor completed
  pdl_launch_dependents();  // ... and the next kernel in the stream may become resident / run its prologue from here on
  const int n = min(*count, n_max);
  const int ins = blockIdx.y;
  if (ins >= n) return;
  const int hd = hs * f, wd = ws * f;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hd * wd) return;
  int y0, y1, x0, x1;
  float fy, fx;
  ab_coord(t / wd, f, hs, y0, y1, fy);
  ab_coord(t % wd, f, ws, x0, x1, fx);
  const float* s = src + static_cast<long>(ins) * hs * ws;
  out[static_cast<long>(ins) * hd * wd + t] = (1.f - fy) * ((1.f - fx) * s[y0 * ws + x0] + fx * s[y0 * ws + x1]) +
                                             fy * ((1.f - fx) * s[y1 * ws + x0] + fx * s[y1 * ws + x1]);
}


// ------------------------------------------------------------------------------------------------ VOS soft aggregation
// external/lib/test/tracker/unicorn_vos.py:129-155 (resize of every object's best mask to the original frame:
// F.interpolate(scale_factor=1/r, bilinear, align_corners=False)[:H, :W] into a zero map) and :105-121 (soft aggregation:
// background = prod_i (1 - m_i) in float32 in list order, argmax over [background, m_id...] with the lower channel winning
// ties, label = object id).  One thread per original-frame pixel; the resized soft masks are optional outputs.
constexpr int kVosMaxObj = 16;
struct VosObjs {
  const float* mask[kVosMaxObj];       // network-resolution soft mask [Hin, Win] or nullptr
  const uint8_t* init_mask[kVosMaxObj];  // original-frame label map [H, W]: object = (label == id), or nullptr
  int id[kVosMaxObj];
  int by_id[kVosMaxObj];  // object indices in ascending id order (argmax tie-breaking)
  int n;
};

__global__ void __launch_bounds__(256) vos_aggregate_kernel(VosObjs o, int Hin, int Win, int H, int W, int hm, int wm, float scale,
                                                             float* __restrict__ soft, uint8_t* __restrict__ seg) {
  pdl_wait();
  pdl_launch_dependents();
  const long total = static_cast<long>(H) * W;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
    float m[kVosMaxObj];
    const bool inside = y < hm && x < wm;
    int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
    float ly = 0.f, lx = 0.f;
    if (inside) {  // PyTorch's source index rule (see bilinear_kernel in misc_kernels.cu)
      const float fy = fmaxf((y + 0.5f) * scale - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * scale - 0.5f, 0.f);
      y0 = min(static_cast<int>(fy), Hin - 1); x0 = min(static_cast<int>(fx), Win - 1);
      y1 = min(y0 + 1, Hin - 1); x1 = min(x0 + 1, Win - 1);
      ly = fy - y0; lx = fx - x0;
    }
    float bg = 1.f;
#pragma unroll 1
    for (int k = 0; k < o.n; ++k) {
      float v = 0.f;
      if (o.init_mask[k]) {
        v = o.init_mask[k][i] == static_cast<uint8_t>(o.id[k]) ? 1.f : 0.f;
      } else if (o.mask[k] && inside) {
        const float* s = o.mask[k];
        v = (1.f - ly) * ((1.f - lx) * s[y0 * Win + x0] + lx * s[y0 * Win + x1]) + ly * ((1.f - lx) * s[y1 * Win + x0] + lx * s[y1 * Win + x1]);
      }
      m[k] = v;
      if (soft) soft[static_cast<long>(k) * total + i] = v;
      bg = bg * (1.f - v);
    }
    float best = bg;
    int label = 0;
#pragma unroll 1
    for (int t = 0; t < o.n; ++t) {
      const int k = o.by_id[t];
      if (m[k] > best) { best = m[k]; label = o.id[k]; }
    }
    seg[i] = static_cast<uint8_t>(label);
  }
}

}  // namespace uc

using namespace uc;

extern "C" int uc_aligned_bilinear_add(const void* src, int lds, int hs, int ws, void* dst, int ldd, int C, int factor, void* stream_v) {
  if (!src || !dst || C % 2 || lds % 2 || ldd % 2 || factor < 1) return set_error(UC_EINVAL, "uc_aligned_bilinear_add: bad arguments");
  const long total = static_cast<long>(hs) * factor * ws * factor * (C / 2);
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((total + 255) / 256, static_cast<long>(num_sms()) * 16)));
  launch_pdl(aligned_bilinear_add_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream_v), static_cast<const uint16_t*>(src), lds, hs, ws,
                                                                                     static_cast<uint16_t*>(dst), ldd, C, factor);
  return check_launch("uc_aligned_bilinear_add");
}

extern "C" int uc_dynamic_masks(const float* mask_feats, const float* up_masks, int h, int w, int up_rate, int d_rate,
                                const float* const* dyn_levels, int ld_dyn, const int* level_hw, const int* level_strides,
                                const float* level_soi, const int* anchors_dev, const int* count_dev, int n_max, float* scratch,
                                float* out_masks, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!mask_feats || !up_masks || !dyn_levels || !level_hw || !level_strides || !level_soi || !anchors_dev || !count_dev || !scratch || !out_masks)
    return set_error(UC_EINVAL, "uc_dynamic_masks: null pointer");
  if (n_max < 1 || ld_dyn < 169 || up_rate < 1 || d_rate < 1) return set_error(UC_EINVAL, "uc_dynamic_masks: bad sizes");
  MaskLevels lv;
  int start = 0;
  for (int k = 0; k < 3; ++k) {
    lv.dyn[k] = dyn_levels[k];
    lv.h[k] = level_hw[2 * k]; lv.w[k] = level_hw[2 * k + 1]; lv.stride[k] = level_strides[k]; lv.soi[k] = level_soi[k];
    lv.start[k] = start;
    start += lv.h[k] * lv.w[k];
  }
  float* logits = scratch;                                   // [n_max, h, w]
  float* mid = scratch + static_cast<long>(n_max) * h * w;    // [n_max, h*up, w*up]
  launch_pdl(mask_logits_kernel, dim3((h * w + 255) / 256, n_max), 256, 0, stream, mask_feats, h, w, lv, ld_dyn, anchors_dev, count_dev, n_max, logits);
  const int H1 = h * up_rate, W1 = w * up_rate;
  launch_pdl(mask_convex_up_kernel, dim3((H1 * W1 + 255) / 256, n_max), 256, 0, stream, logits, up_masks, h, w, up_rate, count_dev, n_max,
                                                                                d_rate == 1 ? out_masks : mid);
  if (d_rate != 1) {
    const int H2 = H1 * d_rate, W2 = W1 * d_rate;
    launch_pdl(mask_final_up_kernel, dim3((H2 * W2 + 255) / 256, n_max), 256, 0, stream, mid, H1, W1, d_rate, count_dev, n_max, out_masks);
  }
  return check_launch("uc_dynamic_masks");
}

extern "C" int uc_vos_aggregate(const UcVosObject* objs, int n, int Hin, int Win, int H, int W, float r, float* soft_out, uint8_t* seg_out,
                                void* stream_v) {
  if (!objs || !seg_out || n < 1 || n > kVosMaxObj) return set_error(UC_EINVAL, "uc_vos_aggregate: 1..%d objects", kVosMaxObj);
  if (Hin < 1 || Win < 1 || H < 1 || W < 1 || !(r > 0.f)) return set_error(UC_EINVAL, "uc_vos_aggregate: bad sizes");
  VosObjs o;
  memset(&o, 0, sizeof(o));
  o.n = n;
  for (int k = 0; k < n; ++k) {
    if (objs[k].id < 1 || objs[k].id > 255) return set_error(UC_EINVAL, "uc_vos_aggregate: object ids must be 1..255");
    o.mask[k] = objs[k].mask; o.init_mask[k] = objs[k].init_mask; o.id[k] = objs[k].id; o.by_id[k] = k;
  }
  std::stable_sort(o.by_id, o.by_id + n, [&](int a, int b) { return o.id[a] < o.id[b]; });
  // F.interpolate(scale_factor = 1/r): output size floor(in * (1/r)) in double precision, source scale 1 / (1/r) as a float
  const double sf = 1.0 / static_cast<double>(r);
  const int hm = std::min(H, static_cast<int>(std::floor(Hin * sf))), wm = std::min(W, static_cast<int>(std::floor(Win * sf)));
  const long total = static_cast<long>(H) * W;
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((total + 255) / 256, static_cast<long>(num_sms()) * 16)));
  launch_pdl(vos_aggregate_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream_v), o, Hin, Win, H, W, hm, wm, static_cast<float>(1.0 / sf), soft_out, seg_out);
  return check_launch("uc_vos_aggregate");
}
