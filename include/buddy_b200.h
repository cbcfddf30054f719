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
  /* Split-precision operands.  Contraction length per tap of Bw / Bw2 (0 = a_c / a2_c).  When larger than the
   * A tensor's channel count the A-side chunk index wraps (chunk % (a_c/64)): with A = [a_hi | a_lo] and
   * Bw = [w_hi | w_hi | w_lo] one launch accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in fp32 (fp32-class
   * products from fp16 tensor-core operands). */
  int32_t k_total, k2_total;
  /* Optional fp8 (e4m3) correction operands, accumulated BEFORE the fp16 products and folded with 2^-14
   * (tcgen05 scale-input-d): a8 = uint8 [batch][H][W][a8_c] = [e4m3(a_lo * 2^9) | e4m3(a_hi)],
   * b8 = uint8 [taps][b_rows][a8_c] = [e4m3(w_hi * 2^5) | e4m3(w_lo * 2^14)], so that
   * out = a_hi*w_hi + 2^-14 * (2^14 (a_lo*w_hi + a_hi*w_lo)) at the cost of ONE extra fp16-pass equivalent
   * (fp8 MMAs consume twice the K per instruction).  a8_2 / b8_2: same for the fused skip conv. */
  const void* a8;
  int32_t a8_c;
  int64_t a8_stride_w, a8_stride_h, a8_stride_b;
  const void* b8;
  const void* a8_2;
  int32_t a8_2_c;
  int64_t a8_2_stride_w, a8_2_stride_h, a8_2_stride_b;
  const void* b8_2;
  /* 1 = force the direct (register -> global) epilogue even where the staged TMA-store epilogue applies
   * (dense fp32 output, n_tile %% 32 == 0); testing / A-B timing only. */
  int32_t no_staged_epilogue;
  /* 1 = never pair CTAs (tcgen05 cta_group::2 over a 2-CTA cluster, the default whenever n_tile >= 32 and B is not
   * per-image); testing / A-B timing only. */
  int32_t no_cta_pairs;
  /* profiling experiments only (results are NOT written): 1 = skip the epilogue body, 2 = only read the accumulator */
  int32_t debug_flags;
  /* 1 = one weight tile per pipeline stage even where a kernel row of three fits (testing / A-B timing only) */
  int32_t one_tap_per_stage;
  /* Stacked tiles (two 16x8-pixel tiles per CTA and work item share every weight tile; plain 3x3 launches with
   * n_tile <= 128): 0 = automatic (when the launch has enough work items), 1 = never, 2 = whenever legal
   * (testing / A-B timing) */
  int32_t single_tile_per_cta;
  /* Fused GroupNorm-backward statistics (staged epilogue only).  When this launch is the data-gradient convolution
   * whose output `da` is the gradient w.r.t. act(GroupNorm(x)) of a tensor x of the SAME geometry (no resampling in
   * between, single tensor), the epilogue also accumulates, per (image, group),
   *     gnb_gsum[b][g][0] += sum dxh,  gnb_gsum[b][g][1] += sum dxh * xh,
   *     xh = (x - mean_g) * rstd_g,  dxh = da * act'(xh * gamma + beta) * gamma
   * i.e. pass 0 of buddy_gn_bwd (which is then called with pass0_done = 1): the statistics pass' 8 bytes per element
   * of HBM reads disappear behind the tensor-core mainloop.  gnb_stats = bundle sums of x (as for buddy_gn_apply),
   * gnb_gsum fp64 [batch][gnb_groups][2], zeroed by the caller. */
  const float* gnb_x; /* NULL = off */
  const double* gnb_stats;
  const float* gnb_gamma;
  const float* gnb_beta;
  double* gnb_gsum;
  int32_t gnb_groups;
  float gnb_eps;
  int32_t gnb_silu;
} buddy_gemm_desc;

int buddy_conv_gemm(const buddy_gemm_desc* d, void* stream);

/* One-off repacking of a convolution / linear weight into buddy_conv_gemm's B operand (replaces the eager
 * permute / cast / cat chain a PyTorch host would run; reference weights: nn.Conv2d.weight [Co][Ci][kh][kw],
 * networks/ncsnpp_utils/layers.py:100-126).  Element (t, n, k) of the packed operand is read from
 *     src[off0 + t*st + (n / ndiv)*sn_outer + (n % ndiv)*sn_inner + (k / kdiv)*sk_outer + (k % kdiv)*sk_inner]
 * (element strides of the fp32 source; n >= n_valid or k >= k_valid: zero padding), which covers forward 3x3
 * weights, their flipped / transposed data-gradient form and the im2col-ordered thin convolutions.
 *   w16 : fp16 [T][N][passes*K] = [hi] | [hi | hi] | [hi | hi | lo],  hi = fp16(w), lo = fp16(w - hi)
 *   w8  : optional uint8 [T][N][2K] = [e4m3(hi * 2^5) | e4m3((w - hi) * 2^14)]  (buddy_gemm_desc.b8) */
typedef struct buddy_pack_desc {
  const float* src;
  int64_t off0, st;
  int32_t ndiv;
  int64_t sn_outer, sn_inner;
  int32_t kdiv;
  int64_t sk_outer, sk_inner;
  int32_t T, N, K, n_valid, k_valid, passes;
  void* w16;
  void* w8; /* or NULL */
} buddy_pack_desc;
int buddy_pack_weights(const buddy_pack_desc* d, void* stream);

/* Weighted prediction error (WPE) dereverberation, one channel, per (utterance, frequency bin) — the
 * `wpe_scaled` warm start of the blind sampler.  Replaces nara_wpe.wpe.wpe(Y, taps, delay, iterations,
 * statistics_mode='full') as called by testing/EulerHeunSamplerDPS.py:32-54 (numpy, CPU, complex128).
 * Y, Z: fp32 [batch][F][T][2] (re, im); arithmetic in fp64.  taps <= 64, T <= 4096 (and 40 T + 16 taps^2 bytes of
 * shared memory <= 227 KB). */
int buddy_wpe(const float* Y, int batch, int F, int T, int taps, int delay, int iterations, float* Z, void* stream);

/* upfirdn2d: zero-insertion upsample -> pad / crop -> 2-D FIR -> downsample, fp32 — the reference's only native
 * operator (pybind11 `upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)`,
 * networks/ncsnpp_utils/op/upfirdn2d.cpp:12-23, op/upfirdn2d_kernel.cu:107-207), same contract:
 * in [major][in_h][in_w][minor], kernel [kh][kw] (<= 32 x 32),
 * out [major][(in_h*up_y + pad_y0 + pad_y1 - kh)/down_y + 1][(in_w*up_x + pad_x0 + pad_x1 - kw)/down_x + 1][minor].
 * Negative pads crop.  The data-gradient is the same operator with the flipped kernel and up / down exchanged. */
int buddy_upfirdn2d(const float* in, const float* kernel, int major, int in_h, int in_w, int minor, int kh, int kw,
                    int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                    float* out, void* stream);


/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) (+nearest-x2 / 2x2-mean resample) (+virtual channel concat) and its data-gradient.
 * Replaces nn.GroupNorm(min(C//4,32), C, eps=1e-6) + SiLU + naive_up/downsample_2d + torch.cat of
 * ResnetBlockBigGANpp.forward / AttnBlockpp.forward / pyramid heads
 * (networks/ncsnpp_utils/layerspp.py:75-76,242-263; up_or_down_sampling.py:59-69; ncsnpp.py:383,396-412).
 *
 * Statistics travel as per-4-channel-bundle fp64 (sum, sum of squares) accumulators [batch][C/4][2], produced
 * either by buddy_conv_gemm's epilogue (`stats`) or by buddy_gn_stats; any group size that is a multiple of
 * 4 channels — including groups straddling the concat boundary (384 = 256 + 128 -> 12 per group) — is
 * assembled from bundles by the consumer.
 */
typedef struct buddy_gn_desc {
  const float* xa; /* fp32 [batch][H][W][Ca] */
  const float* xb; /* fp32 [batch][H][W][Cb] or NULL (Cb = 0): virtual concat [xa | xb] */
  int32_t Ca, Cb;
  const double* stats_a; /* [batch][Ca/4][2] */
  const double* stats_b; /* [batch][Cb/4][2] or NULL */
  const float* gamma;    /* [Ca+Cb] */
  const float* beta;     /* [Ca+Cb] */
  int32_t batch, H, W;   /* resolution of x */
  int32_t groups;
  float eps;
  int32_t silu; /* 1: SiLU after the affine */
  int32_t mode; /* 0 none, 1 nearest x2 upsample after the activation, 2 2x2 mean after the activation */
  void* out;     /* fp16 [batch][H'][W'][Ca+Cb]   (split: [..][2*(Ca+Cb)] = [hi | lo]) */
  void* out_raw; /* optional fp16 copy of the (resampled) input x, operand of the 1x1 skip conv */
  int32_t split; /* 1: every fp16 output carries a second half lo = fp16(v - float(hi));
                    2: fp16 hi only + e4m3 pair [e4m3(lo * 2^9) | e4m3(hi)] in out8 / out_raw8 (see buddy_gemm_desc.a8) */
  void* out8;     /* uint8 [batch][H'][W'][2*(Ca+Cb)] when split == 2 */
  void* out_raw8;
} buddy_gn_desc;

typedef struct buddy_gn_bwd_desc {
  const float* da;    /* fp32 [batch][H'][W'][C]: gradient w.r.t. `out` (at the resampled resolution) */
  const float* dskip; /* optional fp32 [batch][H'][W'][C], added as R^T(dskip) * skip_scale */
  float skip_scale;
  const float* extra_a; /* optional fp32 [batch][H][W][Ca] added to dxa */
  const float* extra_b; /* optional fp32 [batch][H][W][Cb] added to dxb */
  double* gsum;         /* scratch fp64 [batch][groups][2] (zeroed by the call unless pass0_done) */
  float* dxa;           /* optional fp32 out [batch][H][W][Ca] */
  float* dxb;           /* optional fp32 out [batch][H][W][Cb] */
  void* g16a;           /* optional fp16 out = dxa * g16_scale (dgrad operand of the producer) */
  void* g16b;
  float g16_scale;
  void* g8a; /* e4m3 pairs of g16a / g16b when the gn desc has split == 2 */
  void* g8b;
  int32_t pass0_done; /* 1: gsum already holds the group sums (buddy_conv_gemm's gnb_* epilogue): skip pass 0 */
} buddy_gn_bwd_desc;

int buddy_gn_stats(const float* x, int batch, int64_t pixels, int C, double* stats /* += */, void* stream);
int buddy_gn_apply(const buddy_gn_desc* d, void* stream);
/* GroupNorm (+SiLU) of one tensor x [batch][pixels][C] written as fp32 — the activation that the `fir: True` blocks hand
 * to upfirdn2d before the convolution (ResnetBlockBigGANpp.forward with fir, layerspp.py:252-259). */
int buddy_gn_act32(const float* x, const double* stats, const float* gamma, const float* beta, int batch,
                   int64_t pixels, int C, int groups, float eps, int silu, float* out, void* stream);
int buddy_gn_bwd(const buddy_gn_desc* d, const buddy_gn_bwd_desc* g, void* stream);

/* 2-channel (re, im) image helpers — the thin ends of the U-Net.
 * im2col_c2: x fp32 [B][H][W][2] -> fp16 [B][H][W][64], K index = tap*2 + ci (3x3, zero padded), so the
 *            2->128 input conv (ncsnpp.py:331) and the dgrad of the 256->2 pyramid heads (:396-412) run as
 *            plain GEMMs on buddy_conv_gemm.  col2im_c2 is its adjoint (gathers 9 taps).
 * resample_c2 modes: 0 = 2x2 mean (pyramid_downsample, layerspp.py:156), 1 = nearest x2 (+add)
 *            (pyramid_upsample, layerspp.py:117), 2 = adjoint of 0, 3 = adjoint of 1.
 * combine_*: Combine.forward, method 'sum' (layerspp.py:52-59): out = h + Conv1x1(pyr) and d/dpyr.
 * affine_c2: 2x2 affine map per pixel = output_layer (ncsnpp.py:113,445) and its adjoint. */
int buddy_im2col_c2(const float* x, int B, int H, int W, void* col, int split /* 1: [64 hi | 64 lo]; 2: + col8 */,
                    void* col8, float in_scale, void* stream);
int buddy_col2im_c2(const float* dcol, int ld, int B, int H, int W, float* dx, int accumulate, void* stream);
int buddy_resample_c2(const float* in, int B, int Hin, int Win, int mode, const float* add, float* out,
                      int accumulate, void* stream);
int buddy_combine_fwd(const float* h, const float* pyr, const float* w, const float* bias, int64_t pixels, int C,
                      float* out, void* stream);
int buddy_combine_bwd(const float* dout, const float* w, int64_t pixels, int C, float* dpyr, void* stream);
int buddy_affine_c2(const float* x, int64_t pixels, const float* m_host /*[4]*/, const float* b_host /*[2]*/,
                    float* y, void* stream);

/* Attention helpers (AttnBlockpp.forward, layerspp.py:81-85): row softmax fp32 -> fp16 probabilities, its
 * backward dS = P*(dP - <dP,P>)*scale, batched fp16 transpose, and fp32 -> fp16 cast with scale. */
int buddy_softmax_fwd(const float* s, int64_t rows, int n, void* p, int ldp, void* stream);
int buddy_softmax_bwd(const void* p, int ldp, const float* dp, int64_t rows, int n, float scale, void* ds, int ldds,
                      void* stream);
int buddy_transpose_h(const void* in, int batch, int R, int C, int64_t ld_in, int64_t bs_in, void* out,
                      int64_t ld_out, int64_t bs_out, void* stream);
int buddy_cast_scale_h(const float* x, int64_t n, float scale, void* y, void* stream);
/* fp32 [batch][H][W][C] -> tensor-core operand of scale * x: fp16 `out16` (split 1: [hi | lo]) and, for split 2 with
 * out8 != NULL, the e4m3 pair (buddy_gemm_desc.a8); upsample = 1 goes through nearest-neighbour x2 first
 * (out [batch][2H][2W][..]).  Feeds the convolutions that take a RAW tensor: Downsample / Upsample with_conv of the
 * `resblock_type: ddpm` variant (networks/ncsnpp_utils/layerspp.py:93-160) and their data-gradients. */
int buddy_cast_operand(const float* x, int batch, int H, int W, int C, int upsample, float scale, void* out16,
                       void* out8, int split, void* stream);

/* ------------------------------------------------------------------------------------------------
 * STFT / iSTFT as fp32 DFT-GEMMs with exact frame indexing (torch.stft / torch.istft semantics of
 * NCSNppTime.stft/istft, networks/ncsnpp.py:473-496, and of the operators' apply_stft/apply_istft,
 * testing/operators/subband_filtering.py:41-65).  Spectrograms are [batch][bins][frames][2] (re, im) fp32 —
 * the channels-last image the network consumes.  `mat` is a host-built [2*bins][K] matrix (window, onesided
 * weights and normalisation folded in); the adjoints are the same two kernels with the matrices swapped.
 *   analysis : out[b][f][t][c] = sum_{n<K} mat[2f+c][n] * sig[b][t*hop + n]   (t < frames; zeros up to Tout)
 *   synthesis: fr[b][t][n]     = sum_{m<M} S[b][m/2][t][m%2] * mat[m][n]
 *   ola_gather: out[b][s] = tab[s+off] * scale_b[b] * sum_t fr[b][t][s+off-t*hop]   (overlap-add + envelope)
 *   pad_signal: zero (mode 0) / reflect (mode 1) padding with optional per-sample table and per-utterance scale
 *   reflect_fold: adjoint of the reflect padding.
 */
int buddy_dft_analysis(const float* sig, int64_t sig_ld, int batch, const float* mat, int M, int K, int hop,
                       int frames, int Tout, float* out, void* stream);
int buddy_dft_synthesis(const float* S, int batch, int Tin, const float* mat, int M, int K, int frames, float* fr,
                        void* stream);
/* The same two maps for the 1024-point transforms (likelihood STFT, blind operator: subband_filtering.py:41-65) as
 * shared-memory FFTs: mat[2f+c][n] == av[f] * wv[n] * (cos, -sin)(2 pi f n / 1024); tw1024 = exp(-2 pi i k / 1024),
 * k < 512, as float2.  ~20x fewer FLOPs than the DFT-matrix form and lower rounding error. */
int buddy_fft_analysis(const float* sig, int64_t sig_ld, int batch, const float* wv, const float* av,
                       const float* tw1024, int bins, int K, int hop, int frames, int Tout, float* out, void* stream);
int buddy_fft_synthesis(const float* S, int batch, int Tin, const float* wv, const float* av, const float* tw1024,
                        int bins, int K, int frames, float* fr, void* stream);
int buddy_ola_gather(const float* fr, int batch, int frames, int K, int hop, int off, int n_out, const float* tab,
                     const float* scale_b, float* out, int64_t out_ld, void* stream);
int buddy_pad_signal(const float* x, int64_t x_ld, int batch, int N, int left, int total, int mode, const float* tab,
                     const float* scale_b, float* out, void* stream);
int buddy_reflect_fold(const float* dxp, int batch, int N, int L, const float* scale_b, float* dx, int64_t dx_ld,
                       void* stream);

/* Compressed-spectrum likelihood "l2_comp_stft_summean" (utils/losses.py:59-64,74-76), per utterance:
 * loss[b] = weight/frames * sum_{f,t} |Yc - Xc|^2, Zc = (|Z|+1e-8)^c e^{j angle Z}; grad = dloss[b]/dX or NULL. */
int buddy_comp_loss(const float* Y, const float* X, int batch, int64_t bins_times_frames, int frames,
                    float compression, float weight, double* loss, float* grad, void* stream);
/* out[b] = (sum, sum of squares) of x[b][:n] in fp64 — .std() / torch.norm of EulerHeunSamplerDPS.py:30,68,129. */
int buddy_row_stats(const float* x, int64_t ld, int batch, int n, double* out, void* stream);

/* FFT convolution of fast_apply_RIR (utils/reverb_utils.py:25-60), L = 256 * 2^log2_n2 points.
 * mode 0: work <- spectrum of x (reusable as `H`); mode 1: y = real(ifft(fft(x) H))[:n_out]; mode 2: conj(H)
 * (the adjoint).  `tw512` = exp(-2 pi i k / 512), k < 256, as float2.  work: complex scratch [batch][L]. */
int buddy_fftconv(const float* x, int64_t x_ld, int batch, int n_in, int log2_n2, const float* tw512, float* work,
                  const float* H, int64_t h_batch_stride, int mode, float* y, int64_t y_ld, int n_out, void* stream);

/* Noise-level embedding (GaussianFourierProjection + Linear/SiLU, ncsnpp.py:299-318; Dense_0, layerspp.py:262). */
int buddy_fourier_features(const float* t, const float* W, int B, int E, float* out, void* stream);
int buddy_dense(const float* x, const float* W, const float* bias, int B, int In, int Out, int act_in, int act_out,
                float* y, void* stream);
/* Several dense layers on one input in ONE launch — the 21 per-ResBlock Dense_0(act(temb)) time biases
 * (layerspp.py:248-249).  W / bias: the layers' rows concatenated [Out][In] / [Out]; seg [Out][2] = (first row, row
 * count) of the layer each row belongs to; layer with first row r0 writes its own contiguous [B][rows] slab at
 * y + B * r0. */
int buddy_dense_seg(const float* x, const float* W, const float* bias, const int32_t* seg, int B, int In, int Out,
                    int act_in, float* y, void* stream);
/* Philox4x32-10 N(0,1), one stream per utterance (seed[b]), `draw` = running draw index of the sampler. */
int buddy_philox_normal(const int64_t* seeds, uint64_t draw, int batch, int n, float* out, int64_t ld, void* stream);
/* out[b][:] = ca[b] x[b][:] + cb[b] y[b][:] + cc[b] z[b][:] — the Euler/Heun/DPS update algebra. */
int buddy_lincomb3(const float* x, const float* y, const float* z, const float* ca, const float* cb, const float* cc,
                   int batch, int n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Blind reverb operator (BlindSubbandFiltering, testing/operators/subband_filtering.py:67-74,193-351;
 * utils/reverb_utils.py:3-23; optimiser of testing/EulerHeunSamplerDPS.py:71-113,198).  All batched over
 * utterances: every utterance owns H, its 25+25+513*100 parameters and its Adam state.
 *
 * subband_fir (complex, [batch][F][T] rows):
 *   mode 0: Y[t]  = sum_n H[n] X[t+pre-n]            (a = X, h_or_dy = H [batch][F][Nf], stride 0 = shared)
 *   mode 1: dX[s] = sum_n conj(H[n]) dY[s-pre+n]     (a = dY, h_or_dy = H)
 *   mode 2: dH[n] (+)= sum_t conj(X[t+pre-n]) dY[t]  (a = X, h_or_dy = dY)
 * blind_design_fwd/bwd: (decays, weights, phases) -> A [batch][F][Nf], H0 = A e^{j phase} as [batch][F][Nf+2]
 *   complex with a zero frame on each side; kidx/frac = piecewise-linear interpolation tables over the 27 EQ knots.
 * fft_mixed: complex (or real-input) FFT of length N1*256 (25 856 = 101*256), sign -1 forward / +1 inverse, unnormalised.
 * minphase_pw: pointwise stages of minimum_phase_version and of its backward (mode 0..7, see buddy_b200/blind.py).
 * adam_project: torch.optim.Adam update (lr, betas, eps; bias-corrected) + project_params clamps; per utterance layout
 *   [25 decays | 25 weights | phases].
 */
int buddy_subband_fir(const float* a, const float* h_or_dy, int64_t h_batch_stride, float* out, int batch, int F,
                      int Tx, int Nf, int pre, int mode, int accumulate, void* stream);
int buddy_blind_design_fwd(const float* decays, const float* weights, const float* phases, const int* kidx,
                           const float* frac, const float* corr, const float* dpmag, int batch, int F, int Nf,
                           float* A, float* H0, void* stream);
int buddy_blind_design_bwd(const float* decays, const float* weights, const float* phases, const float* A,
                           const int* kidx, const float* frac, const float* corr, const float* dpmag, const float* G,
                           int batch, int F, int Nf, float* dphases, float* ddecays, float* dweights,
                           float* scratch /* [batch][50][Nf], fixed-order reduction of the per-tap partials */,
                           void* stream);
int buddy_fft_mixed(const float* in, int in_real, float* work, float* out, int batch, int N1, int sign,
                    const float* tw512, void* stream);
int buddy_minphase_pw(int mode, const float* c0, const float* c1, const float* r0, const float* r1, float* oc,
                      float* or0, float* or1, int batch, int N, int T, int scale_inv_n, void* stream);
int buddy_adam_project(float* p, const float* g, float* m, float* v, int batch, int n_per, int step, float lr,
                       float beta1, float beta2, float eps, float dmin, float dmax, float wmin, float wmax,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BUDDY_B200_H */
