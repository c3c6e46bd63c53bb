/* unicorn_b200 — C ABI of the B200-native Unicorn per-frame inference hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (passed as void* so the header
 * needs no CUDA include).  No entry point allocates, synchronises or keeps state between calls (apart from
 * a lazily resolved driver entry point), so all of them are CUDA-graph capturable.  All return 0 on success
 * and a negative UC_E* / positive cudaError_t code otherwise; uc_last_error() gives the text.
 * There is NO CPU fallback: on a machine without an sm_100 device every launch returns an error.
 *
 * Reference interfaces replaced (paths relative to MasterBin-IIAU/Unicorn):
 *   uc_msda_forward_*      unicorn/models/ops/src/ms_deform_attn.h:20-39  (ms_deform_attn_forward)
 *                          -> unicorn/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80
 *                          -> ms_deformable_im2col_gpu_kernel, ms_deform_im2col_cuda.cuh:237-299
 *   uc_conv2d              nn.Conv2d / nn.Linear call sites of the backbone, neck, heads:
 *                          backbone/convnext.py:41-54,82-87; network_blocks.py:50-51; unicorn.py:36-44;
 *                          unicorn_head.py:267-336; ops/modules/ms_deform_attn.py:94-113
 *   uc_stem_ln             backbone/convnext.py:77-80 (conv4x4s4 + channels_first LayerNorm)
 *   uc_dwconv7_ln          backbone/convnext.py:43-45 (dwconv 7x7 + LayerNorm)
 *   uc_layernorm           backbone/convnext.py:176-184; deformable_transformer.py:113,121
 *   uc_groupnorm_*         GroupNorm(16,eps 1e-3) from exp/unicorn_track.py:450-470; unicorn.py:38
 *   uc_corr_propagate      external/lib/test/tracker/unicorn_sot.py:95-100, unicorn_vos.py:171-181
 *   uc_head_decode         unicorn_head.py:332-334,467-482
 *   uc_nms_*               utils/boxes.py:33-77 (torchvision.ops.batched_nms)
 */
#ifndef UNICORN_B200_H_
#define UNICORN_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define UC_API __attribute__((visibility("default")))
#else
#define UC_API
#endif

#define UC_OK 0
#define UC_EINVAL (-1)   /* bad argument (shape / alignment / dtype) */
#define UC_EDRIVER (-2)  /* driver entry point (cuTensorMapEncodeTiled) unavailable */
#define UC_ENODEV (-3)   /* no sm_100 device */

/* dtypes of activation tensors */
#define UC_BF16 0
#define UC_F32 1
#define UC_F16 2

/* fused activations */
#define UC_ACT_NONE 0
#define UC_ACT_RELU 1
#define UC_ACT_GELU 2 /* exact erf GELU (nn.GELU()) */
#define UC_ACT_SILU 3
#define UC_ACT_SIGMOID 4

UC_API const char* uc_last_error(void);
UC_API int uc_version(void);
/* 0 if the current device is sm_100 and the driver entry points resolve, else a UC_E* code */
UC_API int uc_check_device(void);

/* Dense convolution / linear layer as an implicit GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *   x : NHWC activations, 16-bit (bf16 or f16 per x_dtype), pixel stride ldx elements (ldx >= Cin, ldx % 8 == 0)
 *   w : packed weights [Cout][KH*KW][Cin] in the same 16-bit type as x (K-major)
 *   y : NHWC output, pixel stride ldy, dtype y_dtype; y = act(conv(x) + bias) ; then y = res + gamma * y if given
 * A Linear layer on [M, Cin] rows is B=1, H=1, W=M, KH=KW=1.  stride in {1,2}; pad < KH.
 * Cin % 8 == 0, Cout % 8 == 0 (pad the weight rows / output channels otherwise).
 */
typedef struct UcConv2d {
  const void* x;
  int x_dtype;
  int B, H, W, Cin, ldx;
  const void* w;
  int Cout, KH, KW, stride, pad;
  const float* bias;  /* [Cout] or NULL */
  int act;            /* UC_ACT_* */
  const float* gamma; /* [Cout] layer scale or NULL */
  const void* res;    /* residual, 16-bit like x, rows of ldres elements, or NULL */
  int ldres;
  void* y;
  int ldy;
  int y_dtype;
  int block_n; /* 0 = auto; else force the N tile (16,32,64,96,128,192,256) */
  /* Optional GroupNorm statistics of the (pre-activation) output, accumulated per (image, group):
   * gn_stats[b][g] = {sum, sumsq}; must be zeroed by the caller; NULL = off. */
  float* gn_stats;
  int gn_groups;
} UcConv2d;
UC_API int uc_conv2d(const UcConv2d* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
