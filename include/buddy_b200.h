/*
 * buddy_b200 — C ABI of the B200-native (sm_100a) kernels behind the BUDDy reverse-diffusion
 * dereverberation hot path.  Plain C, plain pointers and sizes, no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - every entry point is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, negative = error; `buddy_last_error()` returns the message;
 *   - nothing is allocated on behalf of the caller: outputs and scratch are caller-owned.
 *   - activations are channels-last ("NHWC": [batch][H][W][C]); spectrogram images use H = frequency
 *     bins, W = STFT frames, exactly the (B, C, F, T) axes of the reference
 *     (networks/ncsnpp.py:281-297) with C moved innermost.
 *
 * Each entry point names the reference call site it replaces (file:line under the upstream repo).
 * The reference is pure Python/PyTorch; its own FFI (networks/ncsnpp_utils/op/upfirdn2d.cpp:12-23) is dead
 * code under the shipped configuration and is not on this path.  The Python-side binding a
 * maintainer adds is a ctypes stub: see INTEGRATION.md.
 */
#ifndef BUDDY_B200_H
#define BUDDY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BUDDY_OK 0
#define BUDDY_ERR_INVALID (-1)
#define BUDDY_ERR_CUDA (-2)
#define BUDDY_ERR_UNSUPPORTED (-3)

const char* buddy_last_error(void);
int buddy_version(void);
/* Number of kernels launched by this library since load / last reset (bench.py's gpu_launches). */
int64_t buddy_launch_count(void);
void buddy_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / batched GEMM on tcgen05 tensor cores (fp16 operands, fp32 accumulate
 * in TMEM, TMA-staged operands).  Replaces nn.Conv2d 3x3/1x1 (networks/ncsnpp_utils/layers.py:100-126),
 * NIN (layers.py:548-557) and the two attention einsums (layerspp.py:81,85), and their data-gradients
 * (torch.autograd through the same modules, testing/EulerHeunSamplerDPS.py:61-69).
 *
 *   out[b,h,w,n] = scale * ( sum_{tap,k} A[b, h+dy(tap), w+dx(tap), k] * Bw[tap][n][k]
 *                          + sum_{k} A2[b,h,w,k] * Bw2[n][k]            (optional fused 1x1 skip conv)
 *                          + bias[n] + bias_b[b][n] + resid[b,h,w,n] )
 *
 * A / A2 : fp16, logical dims (C, W, H, batch) with element strides (channel stride 1); out-of-image
 *          taps read zeros ("same" padding).  C must be a multiple of 64.
 * Bw     : fp16, [T][rows][K] with K contiguous; T = taps (b_batched = 0) or batch (b_batched = 1).
 * out    : fp32 or fp16 at out + pixel * ldc + col_off + n, pixel = (b*H + h)*W + w.
 * stats  : optional fp64 [batch][n_total/4][2] (sum, sum of squares of the written fp32 values per
 *          4-channel bundle) accumulated with atomics — GroupNorm statistics of the output
 *          (nn.GroupNorm, layerspp.py:219,231) for free.
 */
typedef struct buddy_gemm_desc {
  const void* a;
  int32_t a_c;
  int64_t a_stride_w, a_stride_h, a_stride_b;
  const void* a2; /* may be NULL */
  int32_t a2_c;
  int64_t a2_stride_w, a2_stride_h, a2_stride_b;
  const void* b;
  int32_t b_rows;
  int32_t b_t;
  int64_t b_stride_n, b_stride_t;
  const void* b2; /* may be NULL */
  int32_t b2_rows;
  int64_t b2_stride_n;
  int32_t batch, H, W;
  int32_t taps;      /* 1 or 9 */
  int32_t b_batched; /* 0: third B coordinate = tap; 1: = batch index */
  int32_t n_total;   /* valid output columns */
  int32_t n_tile;    /* MMA N per tile: multiple of 16, 16..256 */
  void* out;
  int32_t out_fp16;
  int64_t ldc;
  int32_t col_off;
  const float* bias;   /* [n_total] or NULL */
  const float* bias_b; /* [batch][n_total] or NULL */
  const float* resid;  /* fp32 [pixel][ld_res] or NULL */
  int64_t ld_res;
  float scale;
  double* stats; /* or NULL */
  int32_t max_ctas; /* 0 = one persistent CTA per SM */
} buddy_gemm_desc;

int buddy_conv_gemm(const buddy_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BUDDY_B200_H */
