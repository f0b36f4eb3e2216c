/*
 * b200seg.h — C ABI of libb200seg.so: the sm_100a kernels behind the CutMix mean-teacher
 * hot path (reference: train_seg_semisup_mask_mt.py:287-476).
 *
 * The reference has no FFI; its boundary is its Python API.  This header is the
 * boundary a maintainer would bind with ctypes: plain pointers + sizes, no
 * torch types.  Every entry point
 *   - returns 0 on success, a non-zero B2_ERR_* code otherwise (message: b2_last_error()),
 *   - takes DEVICE pointers unless the argument is documented "host",
 *   - enqueues work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream),
 *   - never allocates or frees device memory: outputs/workspaces are caller-owned.
 *
 * Tensor layouts: "NCHW" = PyTorch contiguous; "NHWC" = channels-last, channel stride 1 with an
 * explicit leading dimension `ld` (floats per pixel) so ops can write into channel slices of a
 * wider buffer (this is how torch.cat is eliminated).
 */
#ifndef B200SEG_H
#define B200SEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_INVALID 1   /* bad shape / alignment / unsupported configuration */
#define B2_ERR_CUDA 2      /* CUDA runtime / driver error at launch or map encode */
#define B2_ERR_UNSUPPORTED 3

const char* b2_last_error(void);  /* thread-local message of the last failing call */
int b2_version(void);             /* ABI version (monotonic) */
int b2_num_sms(void);             /* SM count of the current device (0 if no device) */

/* ------------------------------------------------------------------------------------------
 * E1  EMA teacher update — replaces optim_weight_ema.py:21-25
 *   t = fl(fl(t*alpha) + fl(s*one_minus_alpha))  (three fp32 roundings, no FMA: bit-exact with
 *   `t.mul_(a); t.add_(s*(1-a))`).  Multi-tensor: `table` is a DEVICE array of n_chunks entries
 *   {tgt ptr, src ptr, count} (struct b2_ema_chunk), one thread block per entry.
 * ------------------------------------------------------------------------------------------ */
typedef struct b2_ema_chunk {
  float* tgt;
  const float* src;
  int64_t count;
} b2_ema_chunk;
int b2_ema_step(const b2_ema_chunk* table, int64_t n_chunks, float alpha, float one_minus_alpha,
                void* stream);
/* Flat fast path: one contiguous state buffer per network. */
int b2_ema_step_flat(float* tgt, const float* src, int64_t count, float alpha,
                     float one_minus_alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * O1 + E1  Fused optimiser step (student) + EMA step (teacher) -- replaces train_seg_semisup_mask_mt.py:465-467
 *   (torch.optim.Adam / SGD on the groups built at :90-98, then optim_weight_ema.py:21-25) with ONE launch.
 *   `table`: DEVICE array, one thread block per entry.  An entry covers <= 8192 consecutive elements of one tensor:
 *     p  student parameter (updated in place)      g  its gradient
 *     m  exp_avg (Adam) / momentum buffer (SGD)     v  exp_avg_sq (Adam; unused for SGD)
 *     t  teacher copy (EMA target; NULL = none)     group  index into lr_groups
 *     k  how many times the reference's parameter group lists this tensor: the per-tensor loop of torch.optim
 *        applies k sequential updates per step() with shared state and step += k (DeepLab v2: deeplab2.py:224-230,
 *        up to k = 4).  k == 0: no optimiser update, EMA only (BatchNorm running statistics: p = student buffer).
 *   lr_groups: DEVICE doubles (one per group; the LR schedule rewrites them, also between CUDA-graph replays).
 *   iter_dev:  DEVICE int64, optimiser steps completed so far; incremented by the call (bias corrections use
 *              step = iter * k + j + 1 for the j-th sequential update).
 *   algo 0 = Adam(beta1, beta2, eps; no weight decay, no amsgrad: the reference's defaults :91-93),
 *        1 = SGD(momentum, weight_decay, nesterov; dampening 0: :95-98).
 *   Arithmetic: fp32, operation order of torch.optim's single-tensor implementations, no FMA contraction; EMA as in
 *   b2_ema_step (three roundings), applied to the parameter value just written.
 * ------------------------------------------------------------------------------------------ */
#define B2_OPT_MAX_K 8
#define B2_OPT_CHUNK 8192
typedef struct b2_opt_chunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* t;
  int32_t count;
  int16_t group;
  int16_t k;
} b2_opt_chunk;
int b2_opt_ema_step(const b2_opt_chunk* table, int64_t n_chunks, const double* lr_groups, int64_t* iter_dev,
                    int algo, double beta1, double beta2, float eps, float sgd_momentum, float sgd_weight_decay,
                    int sgd_nesterov, int do_ema, float ema_alpha, float ema_one_minus_alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation -- replaces the validation loop's argmax + D2H + numpy confusion matrices
 * (train_seg_semisup_mask_mt.py:484-517, evaluation.py:6-62) with one pass over the logits.
 *   logits (N,C,H,W) fp32, labels (N,H,W) int64 (`ignore` and out-of-range labels are skipped),
 *   cm: DEVICE int64 [C*C], ACCUMULATED: cm[truth * C + prediction] += 1 (may be NULL when only pred_out is wanted);
 *   pred_out: optional (N,H,W) int64 argmax (ties -> lowest class index, like torch.argmax).
 *   evaluation.EvaluatorIoU's per-class intersection / union follow exactly from cm:
 *   I[c] = cm[c][c], U[c] = sum_p cm[c][p] + sum_t cm[t][c] - cm[c][c].
 * ------------------------------------------------------------------------------------------ */
int b2_argmax_confusion(const float* logits, const int64_t* labels, int n, int c, int64_t hw, int64_t ignore,
                        int64_t* cm, int64_t* pred_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * M1/M2  Box mask rasterisation — device half of mask_gen.py:110-117
 *   boxes: int32 (N, n_boxes, 4) = [y0, y1, x0, x1) half-open ranges ALREADY resolved with numpy
 *   slice semantics (host side, mask_gen.BoxMaskGenerator.generate_boxes); each rectangle XOR-
 *   toggles the mask which starts at `init` (0 if invert else 1).  out: fp32 (N,1,H,W).
 * ------------------------------------------------------------------------------------------ */
int b2_box_mask_rasterize(const int32_t* boxes, int n, int n_boxes, int h, int w, float init,
                          float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * X1  CutMix image / valid-mask mix — train_seg_semisup_mask_mt.py:350-351
 *   out = fl(fl(a*fl(1-m)) + fl(b*m)), m broadcast over channels.  a,b,out: (N,C,H,W); m: (N,1,H,W).
 *   b == NULL  →  "cut" mode (line 389/401): out = fl(a*m).
 * ------------------------------------------------------------------------------------------ */
int b2_mix(const float* a, const float* b, const float* m, float* out, int n, int c, int64_t hw,
           void* stream);

/* ------------------------------------------------------------------------------------------
 * ICT (interpolation consistency training) -- train_seg_semisup_ict.py:306-392
 *   b2_mix_per_sample: out = fl(fl(a*fl(1-f[i])) + fl(b*f[i])) with ONE factor per sample i (:310-311);
 *     factors: DEVICE fp32 [N] (np.random.beta draws of the host, :306-307).
 *   b2_ict_consistency_fwd_bwd: the consistency block with ICT's teacher target -- softmax of both teacher views, then
 *     probabilities (:329), confidences (:340-342) and logits (:328) mixed with the per-sample factor; same five loss
 *     functions, outputs and partials layout as b2_consistency_fwd_bwd (finish with b2_consistency_finalize).
 *   b2_ict_conf_mean: confbar[p] = mean over the batch of the confidence mask at pixel p.  REQUIRED when
 *     conf_per_pixel && conf_thresh > 0: the reference indexes `conf_mask[:, None, :, :]` (:344), the product with the
 *     (N,1,H,W) loss mask broadcasts to (N,N,1,H,W), i.e. every sample is weighted by the batch-mean mask of the pixel;
 *     conf_rate stays the mean of the mask itself.  confbar: DEVICE fp32 [H*W].
 * ------------------------------------------------------------------------------------------ */
int b2_mix_per_sample(const float* a, const float* b, const float* factors, float* out, int n, int c, int64_t hw,
                      void* stream);
int b2_ict_conf_mean(const float* l0, const float* l1, const float* factors, float* confbar, int n, int c, int64_t hw,
                     float conf_thresh, void* stream);
int b2_ict_consistency_fwd_bwd(const float* l0, const float* l1, const float* ls, const float* factors,
                               const float* lmask, const float* confbar, float* dls, double* partials, int n, int c,
                               int64_t hw, int loss_fn, float conf_thresh, int conf_per_pixel, void* stream);

/* ------------------------------------------------------------------------------------------
 * Augmentation-driven consistency -- train_seg_semisup_aug_mt.py:291-391
 *   The teacher predicts on view 0, the student on view 1; theta: DEVICE fp32 [N,2,3] = the batch's `xf0_to_1` (:281),
 *   the affine map from student-space to teacher-space normalised coordinates.
 *   b2_affine_grid_sample: y = F.grid_sample(x, F.affine_grid(theta, (N,C,OH,OW), align_corners=True), align_corners=True)
 *     (bilinear, zero padding: :302-306, :312).  x: (N,C,IH,IW), y: (N,C,OH,OW), NCHW fp32.
 *   b2_aug_consistency_fwd_bwd: the whole block in one pass -- per student pixel the sampling position, the four teacher
 *     pixels, their soft-maxes, interpolated teacher logits / probabilities / valid mask (loss mask = sampled um0 * um1,
 *     :306), confidence threshold on the interpolated probabilities (:345-352), the five loss functions (:366-387; the
 *     reference's `logits_var` branch reads an unassigned variable and raises -- the host wrapper raises the same
 *     error, the kernel itself implements the formula of the sibling scripts), un-scaled student gradient and partials.
 *     ltea, ls, dls: (N,C,H,W); um0, um1: (N,1,H,W); outputs / partials layout as b2_consistency_fwd_bwd (count:
 *     b2_consistency_num_partials(n, h*w)); finish with b2_consistency_finalize.
 * ------------------------------------------------------------------------------------------ */
int b2_affine_grid_sample(const float* x, const float* theta, float* y, int n, int c, int ih, int iw, int oh, int ow,
                          void* stream);
int b2_aug_consistency_fwd_bwd(const float* ltea, const float* ls, const float* theta, const float* um0,
                               const float* um1, float* dls, double* partials, int n, int c, int h, int w, int loss_fn,
                               float conf_thresh, int conf_per_pixel, void* stream);

/* ------------------------------------------------------------------------------------------
 * VAT (virtual adversarial training) -- train_seg_semisup_vat_mt.py:214-301, 392
 *   The perturbation direction is the input gradient of a consistency loss (network passes + b2_consistency_fwd_bwd);
 *   these entry points are the rest of the block:
 *   b2_col2im: adjoint of b2_im2col -- dx (N,H,W,ldx; C <= 4 channels written) (+)= fold of dcol (N*OH*OW, kpad), the
 *     input gradient of the Cin = 3 stem convolution (autograd's conv backward-data for `x_hat + eps`, :247-268).
 *   b2_sample_l2norm: mag[i] = sqrt(sum_j x[i,j]^2) per sample (normalize_eps :217-219); per = elements per sample;
 *     partials: DEVICE double [n * b2_sample_reduce_blocks(per) * 2] scratch.
 *   b2_vat_adaptive_radius: radius[i] = vat_radius * sqrt(sum (x[..,y+2,x]-x[..,y,x])^2 + sum (x[..,y,x+2]-x[..,y,x])^2) * 0.5
 *     over an (N,C,H,W) batch (:289-296); same scratch with per = C*H*W.
 *   b2_add_scaled_per_sample: out = x + (e / (mag[i] + 1e-12)) * r_i with r_i = radius[i] (DEVICE fp32 [n]) or, when
 *     radius == NULL, the scalar radius_host; every operation rounded to fp32 separately like the reference's tensor
 *     expression (:220 normalize, :223 / :301 scale, :247 / :392 add).  x == NULL: out = the scaled direction only.
 * ------------------------------------------------------------------------------------------ */
int b2_col2im(const float* dcol, float* dx, int n, int h, int w, int c, int ldx, int kh, int kw, int stride, int pad,
              int dil, int oh, int ow, int kpad, int accumulate, void* stream);
int64_t b2_sample_reduce_blocks(int64_t per);
int b2_sample_l2norm(const float* x, int n, int64_t per, double* partials, float* mag, void* stream);
int b2_vat_adaptive_radius(const float* x, int n, int c, int h, int w, float vat_radius, double* partials, float* radius,
                           void* stream);
int b2_add_scaled_per_sample(const float* x, const float* e, const float* mag, const float* radius, float radius_host,
                             float* out, int n, int64_t per, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-format boundary -- datapipe/seg_transforms_cv.py:587-672 (SegCVTransformNormalizeToTensor)
 *   The last DataLoader stage moved behind the host-to-device copy: batches cross PCIe as uint8.
 *   b2_normalize_to_tensor: img DEVICE uint8 (N,H,W,cin), cin = 3, or 4 with the padding alpha channel ->
 *     out DEVICE fp32 (N,3,H,W) = float32((img * (1/255) - mean [* alpha]) / std), float64 arithmetic like numpy (:596-614).
 *     mean3 / std3: HOST pointers to 3 doubles each (net.MEAN / net.STD), both NULL = no standardisation (:609).
 *   b2_u8_to_tensor: mode 0 labels uint8 -> int64 (:617), mode 1 valid mask uint8 -> float32(m * (1/255)) (:620).
 * ------------------------------------------------------------------------------------------ */
int b2_normalize_to_tensor(const uint8_t* img, int n, int h, int w, int cin, const double* mean3, const double* std3,
                           float* out, void* stream);
int b2_u8_to_tensor(const uint8_t* src, int64_t count, int mode, void* dst, void* stream);

/* Device crop + flip + normalise: SegCVTransformPad (datapipe/seg_transforms_cv.py:30-99) + SegCVTransformRandomCrop (:102-167)
 * + SegCVTransformRandomFlip (:445-520) + SegCVTransformNormalizeToTensor (:587-672) in one gather pass over variable-size uint8
 * source images.  The random parameters are drawn on the host in the reference's order (cutmix_semisup_seg_b200/input_pipeline.py);
 * `table` is a DEVICE array of n entries.  Outputs: image fp32 (n,3,out_h,out_w); labels int64 (n,1,out_h,out_w) or NULL
 * (255 in the padding and when an entry has no labels); mask fp32 (n,1,out_h,out_w) or NULL (0 in the padding).
 * out_h x out_w is the crop size (transposed samples, flip_d, need a square crop). */
typedef struct {
  const uint8_t* image;        /* DEVICE (h0, w0, 3) uint8 */
  const uint8_t* labels;       /* DEVICE (h0, w0) uint8 or NULL */
  const uint8_t* mask;         /* DEVICE (h0, w0) uint8 or NULL */
  int32_t h0, w0;              /* source size */
  int32_t pad_top, pad_left;   /* rows / columns of padding in front of the source (pad_h // 2, pad_w // 2) */
  int32_t padded;              /* 1: the reference would have padded this sample (alpha-channel standardisation) */
  int32_t pos_y, pos_x;        /* crop origin in padded coordinates */
  int32_t crop_h, crop_w;
  int32_t flip_x, flip_y, flip_d;
} b2_crop_entry;               /* 72 bytes */
int b2_crop_flip_normalize(const b2_crop_entry* table, int n, int out_h, int out_w, const double* mean3, const double* std3,
                           float* image, int64_t* labels, float* mask, void* stream);
/* The same gather with uint8 RGBA pixels as output (n, out_h, out_w, 4): alpha = 255 inside the source image, 0 in the padding
 * (the strong-colour branch: b2_crop_flip_u8 -> b2_colour_jitter -> b2_normalize_to_tensor with cin = 4). */
int b2_crop_flip_u8(const b2_crop_entry* table, int n, int out_h, int out_w, uint8_t* image_rgba, int64_t* labels, float* mask,
                    void* stream);

/* Scale / rotation crops on uint8 pixels -- SegCVTransformRandomCropScaleHung (datapipe/seg_transforms_cv.py:169-303: cv2.resize
 * INTER_LINEAR / INTER_NEAREST of a window of the padded image) and SegCVTransformRandomCropRotateScale (:305-449: cv2.warpAffine,
 * BORDER_REFLECT_101 image, BORDER_CONSTANT labels 255 / mask 0), then SegCVTransformRandomFlip.flip_image (:467-474): one gather
 * per output pixel in OpenCV's fixed-point arithmetic (byte-identical to cv2).  `tables`: DEVICE int32, 3 * (out_h + out_w) entries
 * per sample at tab_off (layout: csrc/input.cu; built by input_pipeline.resize_tables / warp_tables).  Outputs as b2_crop_flip_u8. */
typedef struct {
  const uint8_t* image;        /* (h0, w0, 3) */
  const uint8_t* labels;       /* (h0, w0) or NULL */
  const uint8_t* mask;         /* (h0, w0) or NULL */
  int32_t h0, w0;
  int32_t mode;                /* 0: crop window + resize, 1: warpAffine */
  int32_t pad_top, pad_left, padded;          /* mode 0: virtual padding (SegCVTransformPad) */
  int32_t pos_y, pos_x, src_h, src_w;         /* mode 0: window in padded coordinates */
  int32_t image_interp, mask_interp;          /* 0 nearest, 1 linear, 2 exact 2x decimation (2x2 average); labels: nearest */
  int32_t tab_off;
  int32_t flip_x, flip_y, flip_d;
} b2_geom_entry;               /* 88 bytes */
int b2_geom_u8(const b2_geom_entry* table, const int32_t* tables, int n, int out_h, int out_w, uint8_t* image_rgba, int64_t* labels,
               float* mask, void* stream);

/* Colour jitter on uint8 pixels, in place: torchvision ColorJitter + RandomGrayscale on PIL images as applied by the reference's
 * SegCVTransformTVT (datapipe/seg_transforms_cv.py:541-586; assembled in train_seg_semisup_mask_mt.py:169-179), byte-identical to
 * Pillow's arithmetic.  Per image up to 4 operations in the drawn order -- op 0 brightness, 1 contrast, 2 saturation (factor =
 * blend alpha), 3 hue (hue_shift = np.int32(hue_factor * 255).astype(uint8)) -- then grey = luma if `grey`.
 * img: DEVICE uint8 (n, h, w, cstride), cstride 3 or 4 (alpha untouched); table_dev / table_host: the same n entries in DEVICE and
 * HOST memory (the host copy plans the passes); workspace: DEVICE n x uint64. */
typedef struct {
  int32_t n_ops;
  int32_t op[4];
  float factor[4];
  int32_t hue_shift[4];
  int32_t grey;
} b2_colour_entry;             /* 56 bytes */
int b2_colour_jitter(uint8_t* img, int n, int h, int w, int cstride, const b2_colour_entry* table_dev,
                     const b2_colour_entry* table_host, unsigned long long* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * U-Net decoder operators -- architectures/resunet.py:10-34, 57-92 (identical in denseunet.py:10-34, 98-124); NHWC fp32
 *   b2_upsample2x_add: y (N,2H,2W,C; ldy) = nearest x2 up-sampling of x (N,H,W,C; ldx) [+ skip (N,2H,2W,C; lds)]   (:31-32, :87)
 *   b2_upsample2x_bwd: dx (+)= sum over each 2x2 block of dy (adjoint of the up-sampling; d(skip) = dy)
 *   b2_mul_mask: y = x * mask * scale, mask dense (rows, C): nn.Dropout on a raw convolution output (:88), forward and backward
 * ------------------------------------------------------------------------------------------ */
int b2_upsample2x_add(const float* x, int ldx, const float* skip, int lds, float* y, int ldy, int n, int h, int w, int c,
                      void* stream);
int b2_upsample2x_bwd(const float* dy, int lddy, float* dx, int lddx, int n, int h, int w, int c, int accumulate, void* stream);
int b2_mul_mask(const float* x, int ldx, const float* mask, float scale, float* y, int ldy, int64_t rows, int c, void* stream);
/* DenseNet encoder glue (architectures/denseunet.py: torchvision densenet161 features):
 *   b2_avgpool2x2 / _bwd: nn.AvgPool2d(2, 2) of the transition layers (floor output size), NHWC with leading dimensions;
 *   b2_scale_channels: dst[r,c] (+)= g[r,c] * scale[c], the data gradient of a stand-alone eval-mode BatchNorm (DenseNet's
 *     pre-activation norms act on a concatenation), accumulated into a channel prefix of the concatenation's gradient. */
int b2_avgpool2x2(const float* x, int ldx, float* y, int ldy, int n, int ih, int iw, int c, void* stream);
int b2_avgpool2x2_bwd(const float* dy, int lddy, float* dx, int lddx, int n, int ih, int iw, int c, int accumulate, void* stream);
int b2_scale_channels(const float* g, int ldg, const float* scale, float* dst, int ldd, int64_t rows, int c, int accumulate,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * L1  Fused CutMix consistency loss — train_seg_semisup_mask_mt.py:363-367,406-420,428-459
 *   Inputs (NCHW fp32): l0, l1 teacher logits of the two views (l1 == NULL → cut mode, l_t = l0),
 *   ls student logits, m mix mask (N,1,H,W) (NULL → no logit mixing), lmask per-pixel loss mask
 *   (N,1,H,W) (um_mixed for mix mode, cut_mask*um for cut mode).
 *   One pass: l_t = l0*(1-m)+l1*m; p_t, p_s = softmax; conf = max p_t >= conf_thresh;
 *   q = per-pixel loss (loss_fn); writes the UNSCALED gradient g = lmask * dq/dls (times conf if
 *   conf_per_pixel) to dls and per-block partials to `partials` (3 doubles per block).
 *   b2_consistency_finalize (1 block, fixed order) then produces out[0..3]:
 *     out[0] loss  (= mean(q*lmask*conf) * ramp, the value the reference logs, line 461)
 *     out[1] conf_rate
 *     out[2] grad_scale: true dls = g * out[2]   ( = conf/(N*H*W) * ramp * cons_weight )
 *     out[3] unsup_loss = out[0] * cons_weight
 *   loss_fn: 0 var, 1 logits_var, 2 logits_smoothl1, 3 bce, 4 kld.   C <= 64.
 *   conf_thresh <= 0 disables confidence masking (line 407/419).
 * ------------------------------------------------------------------------------------------ */
int64_t b2_consistency_num_partials(int n, int64_t hw);
int b2_consistency_fwd_bwd(const float* l0, const float* l1, const float* ls, const float* m,
                           const float* lmask, float* dls, double* partials, int n, int c,
                           int64_t hw, int loss_fn, float conf_thresh, int conf_per_pixel,
                           void* stream);
int b2_consistency_finalize(const double* partials, int64_t n_partials, int64_t n_pixels,
                            float conf_thresh, int conf_per_pixel, float ramp, float cons_weight,
                            float* out4, void* stream);

/* ------------------------------------------------------------------------------------------
 * L2  Per-pixel cross-entropy, ignore_index — train_seg_semisup_mask_mt.py:126,300-301
 *   logits NCHW fp32, labels int64 (N,H,W).  Writes UNSCALED gradient (softmax - onehot)*valid to
 *   dlogits, partials (2 doubles/block: sum nll, n_valid).  Finalize: out[0] = mean nll over valid,
 *   out[1] = n_valid, out[2] = grad_scale = 1/n_valid.
 * ------------------------------------------------------------------------------------------ */
int64_t b2_ce_num_partials(int n, int64_t hw);
int b2_ce_fwd_bwd(const float* logits, const int64_t* labels, float* dlogits, double* partials,
                  int n, int c, int64_t hw, int64_t ignore_index, void* stream);
int b2_ce_finalize(const double* partials, int64_t n_partials, float* out3, void* stream);

/* x[i] *= scale_dev[0] * scale_host   (applies a device-resident gradient scale) */
int b2_scale_inplace(float* x, int64_t count, const float* scale_dev, float scale_host,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * A3/A4/A5  Convolution as tcgen05 implicit GEMM (fprop, dgrad, wgrad) — replaces
 *   nn.Conv2d forward/backward in architectures/deeplab2.py:65-128,140-150 and torchvision
 *   ResNet/ASPP + architectures/deeplab3plus.py:29-48.
 *
 * b2_conv_gemm:  D[pix, n] = epilogue( sum_{tap, k} A[pix@tap, k] * B[n, tap, k] )
 *   A: NHWC activations (N, IH, IW, K) with leading dim lda;  B: (NB, T_b, K) K-contiguous
 *   (fprop: weights KRSC; dgrad: transposed weights CRSK).  TF32 tensor-core math, fp32 accum.
 *   Taps: table of n_taps entries (dh, dw, b_tap): input pixel for output (h,w) is
 *   (h*istride + dh, w*istride + dw); out-of-range pixels contribute zero (TMA OOB fill).
 *   Output pixel (n,h,w) of the OH x OW grid is stored at out[((n*FH + h*ostride+ooh)*FW +
 *   w*ostride+oow)*ldd + n_col]  (FH x FW = full output buffer; ostride>1 = scatter for strided
 *   dgrad).  Epilogue, in order:  v = acc*scale[n] + shift[n] (NULL = identity);
 *   v += addend[pix, n] (ld = ld_add);  relu;  v = (gate[pix, n] > 0) ? v : 0 (ld = ld_gate);
 *   v *= scale2[n];  if accumulate v += D_old.
 *   n_split > 1 = 3xTF32 precision mode: a_lo/b_lo hold the low parts (see b2_split_tf32) and the
 *   kernel accumulates A*B + A_lo*B + A*B_lo (+ A_lo*B_lo when n_split == 4).
 * ------------------------------------------------------------------------------------------ */
typedef struct b2_conv_params {
  const float* a; const float* a_lo;      /* activations (and low part, or NULL) */
  const float* b; const float* b_lo;      /* weights (NB, T_b, K) */
  float* d;                               /* output */
  int32_t n, ih, iw, k, lda;              /* A dims: batch, in H, in W, channels K, leading dim */
  int32_t nb, tb, ldb;                    /* B dims: NB rows (GEMM-N), T_b taps, floats per (n,tap) row (>= K, %4) */
  int32_t oh, ow;                         /* logical output grid */
  int32_t fh, fw, ldd, ostride, ooh, oow; /* output buffer geometry */
  int32_t istride;
  int32_t n_taps;
  const int32_t* taps;                    /* HOST array, 3*n_taps: dh, dw, b_tap */
  const float* scale; const float* shift; /* per-output-channel, or NULL */
  const float* addend; int32_t ld_add;    /* laid out like D (same pixel mapping) */
  const float* gate; int32_t ld_gate;
  const float* scale2;
  int32_t relu, accumulate, n_split;
  int32_t max_ctas;                       /* 0 = one persistent CTA per SM */
  /* Optional fused column statistics of the stored value v (needs gate, no accumulate, nb % 4 == 0): partial sums
   * over row blocks of <= 32 pixels, stats[(blk*2 + j)*ld_stats + ch], j = 0: sum v, j = 1: sum v*(gate - stats_sub),
   * blk < b2_conv_stats_rows(p).  Every (blk, ch < nb) entry is written exactly once (deterministic); reduce them with
   * b2_bn_eval_param_grad_from_stats.  Replaces a separate b2_bn_eval_param_grad pass over g and y. */
  float* stats; int32_t ld_stats;
  const float* stats_sub; int32_t ld_stats_sub;   /* laid out like D, or NULL */
} b2_conv_params;
int b2_conv_gemm(const b2_conv_params* p, void* stream);
int64_t b2_conv_stats_rows(const b2_conv_params* p);
/* Host-only description of the launch the CTA-pair kernel (output-channel tiles of 256) runs for `p` on `sms` SMs; no GPU, the
 * tensor pointers are not dereferenced.  out = int64[8] {bw, bh, bn (pixel box of an M tile), tile pairs, pipeline stages of the
 * whole launch with padding-only taps skipped, stages of the busiest CTA pair, CTA pairs, tcgen05.mma instructions per stage}.
 * A stage keeps the tensor pipe of both SMs of a pair busy for instructions x 128 cycles (kind::tf32, M256 x N256 x K8), so
 * stages x 512 / sm__cycles_elapsed is the launch's tensor-pipe occupancy (tools/tensor_busy.py). */
int b2_conv_gemm_plan(const b2_conv_params* p, int32_t sms, int64_t* out);

/* wgrad:  dW[m, tap, c] (+)= sum_pix dY[pix, m] * X[pix@tap, c]
 *   dY: NHWC (N, OH, OW, M) ld = ldy;  X: NHWC (N, IH, IW, C) ld = ldx;  dW: (M, T, C) fp32.
 *   Split over pixel ranges: `workspace` holds n_splits partial (M,T,C) slabs reduced in fixed
 *   order by a second kernel (deterministic).  b2_conv_wgrad_workspace returns the bytes needed. */
typedef struct b2_wgrad_params {
  const float* dy; const float* dy_lo;
  const float* x;  const float* x_lo;
  float* dw;
  int32_t n, oh, ow, m, ldy;
  int32_t ih, iw, c, ldx;
  int32_t istride;
  int32_t n_taps; const int32_t* taps;    /* HOST array 3*n_taps: dh, dw, w_tap */
  int32_t tw;                             /* taps in dW's layout */
  int32_t accumulate, n_split;
  void* workspace; size_t workspace_bytes;
  int32_t max_ctas;
  const float* row_scale;                 /* per-m scale applied to the gradient rows, or NULL */
  int32_t kchunk;                         /* >0: at most this many pixels per TMEM accumulation (precision mode) */
} b2_wgrad_params;
size_t b2_conv_wgrad_workspace(const b2_wgrad_params* p);
int b2_conv_wgrad(const b2_wgrad_params* p, void* stream);
/* Host-only self check of the launch plan (no GPU, pointers are not dereferenced): out = int64[5] {splits, work units,
 * pipeline stages, units whose two stage counts (producer walk vs closed form of the MMA issuer) disagree, CTA-pair kernel}. */
int b2_conv_wgrad_plan_check(const b2_wgrad_params* p, int64_t* out);
/* Host-only view of the plan's load balance: out = int64[5] {pixel splits, tap rotation per split, pipeline stages of the busiest
 * worker (CTA or CTA pair), pipeline stages of the whole launch, workers}.  Dilated layers skip the pixel boxes whose input lies
 * in the padding, so the units of different taps differ in length; the planner picks more pixel splits and rotates the tap from
 * split to split so that every worker gets a mix of taps (debug knob 14 = 0 turns that off). */
int b2_conv_wgrad_plan_balance(const b2_wgrad_params* p, int64_t* out);

/* Debug knobs (tests / timing experiments only): key 1 = wgrad smem-descriptor variant; 2 = 1 forces the single-CTA conv kernel;
 * 3 = epilogue timing bits; 4 = PF-build K limit; 5 = 1 forces the single-CTA wgrad kernel; 6 = wgrad timing bits;
 * 7 = 1 restores the tap-outer / K-block-inner producer order of the conv kernels (A/B timing of the L2 working set);
 * 8 = 0 disables the TMA epilogue of the CTA-pair conv kernel (register epilogue everywhere);
 * 10 = 5 selects the 5-stage build of the compute-bound CTA-pair kernel (default 6 stages; A/B timing);
 * 11 = programmatic dependent launch of the tensor-core kernels (1 on, 0 off; environment B200SEG_PDL);
 * 12 = consecutive fprop / dgrad launches walk their tiles in alternating directions (L2 reuse of the tensor written last;
 *      1 on, 0 off; environment B200SEG_ALT_DIR);
 * 13 = operand ring depth of the CTA-pair weight-gradient kernel (6 or 7 stages; environment B200SEG_WGRAD_STAGES);
 * 14 = 0 disables the load-balanced work decomposition of the weight-gradient kernel for dilated layers (more pixel splits,
 *      taps rotated from split to split), 1 (default) enables it; 15 / 16 = force the pixel splits / the tap rotation (experiments);
 * 17 = 0: per-tile tap masks of the conv kernels by a loop over the taps instead of the host-built row / column tables (A/B);
 * 18 = L2 prefetch of the A stream of 1x1 layers (K >= 512) in 256-channel boxes by the CTA-pair kernel (environment
 *      B200SEG_WIDE_PF; 1 on, 0 off);
 * 20 = 1: flat-index forms of the stem im2col and the NHWC bilinear resize instead of the row-based kernels (bit-identical). */
void b2_debug_set(int key, int value);
/* Diagnostics: device buffer of 2048 int64 receiving clock64 stamps of CTA 0's pipeline roles in the 2-CTA conv
 * kernel ([0,512) producer stage issue, [512,1024) MMA stage acquired, [1024,1536) MMA tile begin/accumulator
 * acquired, [1536,2048) epilogue warp tile wait begin/end); NULL disables.  Not part of the product path. */
void b2_debug_trace(void* buf);

/* hi = x with the 13 low mantissa bits cleared (exact TF32), lo = x - hi (exact). */
int b2_split_tf32(const float* x, float* hi, float* lo, int64_t count, void* stream);

/* (A, T, B) -> (B, T, ldd) transpose of a weight tensor (KRSC -> CRSK), rows padded with zeros to
 * ldd >= A; optional per-A scale (folded BN scale of the output channel) applied on the fly. */
int b2_transpose_w(const float* src, float* dst, int a, int t, int b, int ldd, const float* scale_a,
                   void* stream);

/* Multi-tensor forms (ONE launch per network and pass; tables in DEVICE memory, built once by the caller).  Replace the
 * per-layer launches of b2_bn_fold / b2_transpose_w in the iteration (reference: the BatchNorm modules in eval mode under
 * freeze_batchnorm, deeplab2.py:244-245 / deeplab3plus.py:120-121, and autograd's transposed-weight convolutions). */
typedef struct {
  const float* gamma; const float* beta; const float* mean; const float* var;
  float* scale; float* shift;
  int32_t c; float eps;
} b2_bn_fold_entry;                       /* 56 bytes */
int b2_bn_fold_multi(const b2_bn_fold_entry* table, int n_entries, int max_c, void* stream);
typedef struct {
  const float* src; float* dst; const float* scale;      /* scale: per-A factor or NULL */
  int32_t a, t, b, ldd;
  int64_t block_begin;                    /* prefix sum of ceil(b/32) * ceil(ldd/32) * t over the preceding entries */
} b2_transpose_entry;                     /* 48 bytes */
int b2_transpose_w_multi(const b2_transpose_entry* table, int n_entries, int64_t total_blocks, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-bound network ops (NHWC fp32).
 * ------------------------------------------------------------------------------------------ */
int b2_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, int ldd, void* stream);
int b2_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, int lds, void* stream);
/* im2col for the Cin=3 stem: out (N*OH*OW, kpad) rows [r][s][c], zero padded to kpad. */
int b2_im2col(const float* x, float* col, int n, int h, int w, int c, int ldx, int kh, int kw,
              int stride, int pad, int dil, int oh, int ow, int kpad, void* stream);
/* 3x3 stride-2 pad-1 max pooling (deeplab2.py:146 ceil_mode / torchvision floor): caller passes
 * oh, ow.  idx: uint8 argmax tap (0..8) for the backward. */
int b2_maxpool3x3s2(const float* x, float* y, uint8_t* idx, int n, int h, int w, int c, int oh,
                    int ow, void* stream);
int b2_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int n, int h, int w,
                        int c, int oh, int ow, void* stream);
/* Bilinear resize (F.interpolate, deeplab2.py:204 align_corners=True; deeplab3plus.py:54,77
 * align_corners=False).  NHWC -> NHWC (ldd) or NHWC -> NCHW (to_nchw).  Backward: dy layout
 * mirrors the forward output; `scale_dev` (may be NULL) * scale_host multiplies the gradient. */
int b2_bilinear_fwd(const float* x, float* y, int n, int ih, int iw, int c, int ldx, int oh, int ow,
                    int ldy, int align_corners, int to_nchw, void* stream);
int b2_bilinear_bwd(const float* dy, float* dx, int n, int ih, int iw, int c, int ldx, int oh,
                    int ow, int ldy, int align_corners, int from_nchw, const float* scale_dev,
                    float scale_host, int accumulate, void* stream);
/* Separable form of the same backward for an NCHW dY (the final resize to the input resolution, reference
 * deeplab3plus.py:77 / deeplab2.py:204): horizontal pass into `workspace` (b2_bilinear_bwd_nchw_workspace_floats(n, c, iw, oh)
 * floats), then the vertical pass; identical results to b2_bilinear_bwd(from_nchw = 1). */
int64_t b2_bilinear_bwd_nchw_workspace_floats(int n, int c, int iw, int oh);
int b2_bilinear_bwd_nchw(const float* dy, float* dx, float* workspace, int n, int ih, int iw, int c, int ldx, int oh, int ow,
                         int align_corners, const float* scale_dev, float scale_host, int accumulate, void* stream);
/* Global average pool (ASPPPooling) and its backward (broadcast /HW). */
int b2_gap_fwd(const float* x, float* y, int n, int hw, int c, int ldx, void* stream);
int b2_gap_bwd(const float* dy, float* dx, int n, int hw, int c, int ldx, int accumulate,
               void* stream);
/* Broadcast a (N, C) vector over HW pixels into an NHWC slice (bilinear from 1x1), + backward. */
int b2_bcast_fwd(const float* v, float* y, int n, int hw, int c, int ldy, void* stream);
int b2_bcast_bwd(const float* dy, float* dv, int n, int hw, int c, int ldy, void* stream);
/* Train-mode batch norm over NHWC (N*H*W rows, C channels), DLv3+ head.
 *   stats: mean[c], rstd[c] (biased var, eps), updates running stats with momentum (unbiased var).
 *   apply: y = relu?((x-mean)*rstd*gamma+beta + residual?) * dropmask? ; y may be a channel slice (ldy).
 *   bwd: given dy (w.r.t. y), y-side gate (relu: y>0), computes dx, dgamma, dbeta; g_out (optional)
 *        receives the gated gradient (the residual branch's gradient). */
int b2_bn_stats(const float* x, int64_t rows, int c, int ldx, float eps, float momentum,
                float* mean, float* rstd, float* running_mean, float* running_var,
                double* workspace, void* stream);
int64_t b2_bn_workspace_doubles(int64_t rows, int c);
int b2_bn_apply(const float* x, int64_t rows, int c, int ldx, const float* mean, const float* rstd,
                const float* gamma, const float* beta, int relu, const float* dropmask,
                float drop_scale, float* y, int ldy, const float* residual, int ldr, void* stream);
int b2_bn_bwd(const float* dy, int lddy, const float* x, int ldx, const float* y, int ldy,
              int64_t rows, int c, const float* mean, const float* rstd, const float* gamma,
              int relu, const float* dropmask, float drop_scale, float* dx, int lddx,
              float* dgamma, float* dbeta, int accumulate_params, float* g_out, int ldgo,
              const float* gate_beta, double* workspace, void* stream);
/* gate_beta != NULL (relu layers WITHOUT a residual): the ReLU gate is recomputed from x as gamma*(rstd*(x-mean))+gate_beta > 0 with
 * the instruction sequence of b2_bn_apply (bit-identical sign) and y is not read: 5 instead of 7 tensor passes over HBM. */
/* Eval-mode (frozen) BN folded constants: scale = gamma*rsqrt(var+eps), shift = beta - mean*scale. */
int b2_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
               float eps, float* scale, float* shift, int c, void* stream);
/* Frozen-statistics BN with trainable affine (DLv3+ backbone under freeze_batchnorm):
 *   dbeta[c] (+)= sum g, dgamma[c] (+)= sum g*xhat, g = dy where gate > 0 (gate NULL = always),
 *   xhat = (ybn - beta)/gamma recovered from the stored BN output `ybn`. */
int b2_bn_eval_param_grad(const float* dy, int lddy, const float* ybn, int ldy, int64_t rows, int c,
                          const float* gamma, const float* beta, const float* gate, int ldg,
                          const float* sub, int lds, float* dgamma, float* dbeta, int accumulate,
                          double* workspace, void* stream);
/* Same result from the partial sums a b2_conv_gemm epilogue wrote (b2_conv_params.stats): fixed-order reduction over
 * the `stat_rows` row blocks in double, then dbeta (+)= S0, dgamma (+)= (S1 - beta*S0)/gamma. */
int b2_bn_eval_param_grad_from_stats(const float* stats, int64_t stat_rows, int ld_stats, int c, const float* gamma,
                                     const float* beta, float* dgamma, float* dbeta, int accumulate,
                                     double* workspace, void* stream);   /* >= b2_bn_stats_workspace_doubles(c) */
int64_t b2_bn_stats_workspace_doubles(int c);
/* Frozen-BN parameter gradients from the WEIGHT gradient of the convolution feeding the BN (no pass over the
 * activations): with y = scale*conv(x, W) + shift and gw = scale*dWraw (what b2_conv_wgrad stores with
 * row_scale = scale = gamma*rsqrt(var+eps)),  sum_pix g*conv = <W[c,:], dWraw[c,:]>  gives
 *   dbeta[c] = (accumulate ? dbeta[c] : 0) + sum_pix g[pix,c];   dgamma[c] = <W[c], gw[c]>/gamma - rsqrt(var+eps)*mean*dbeta[c].
 * dgamma is SET from the accumulated totals (gw and dbeta accumulate over backward passes), which equals the accumulated
 * gradient whenever W.grad, gamma.grad and beta.grad were zeroed together.  w, gw: (c, row_len) contiguous rows.
 * Replaces autograd's native_batch_norm_backward for eval-mode BN with trainable affine (torchvision backbone under
 * freeze_batchnorm, reference deeplab3plus.py:120-121).  `_from_stats` takes sum_pix g from the partial sums a
 * b2_conv_gemm epilogue wrote (b2_conv_params.stats, entry j = 0); the other form reduces dy itself.
 * workspace: >= b2_bn_stats_workspace_doubles(c) / b2_bn_workspace_doubles(rows, c) doubles. */
int b2_bn_eval_param_grad_wdot_from_stats(const float* stats, int64_t stat_rows, int ld_stats, int c, const float* w,
                                          const float* gw, int64_t row_len, const float* gamma, const float* mean,
                                          const float* var, float eps, float* dgamma, float* dbeta, int accumulate,
                                          double* workspace, void* stream);
int b2_bn_eval_param_grad_wdot(const float* dy, int lddy, int64_t rows, int c, const float* w, const float* gw,
                               int64_t row_len, const float* gamma, const float* mean, const float* var, float eps,
                               float* dgamma, float* dbeta, int accumulate, double* workspace, void* stream);
/* dropout keep-mask: mask[i] = (hash(seed, offset + *offset_dev + i) >= p) ? 1 : 0; offset_dev (may be NULL) is a
 * device counter so that CUDA-graph replays advance the random stream. */
int b2_dropout_mask(float* mask, int64_t count, float p, uint64_t seed, uint64_t offset,
                    const uint64_t* offset_dev, void* stream);
/* elementwise helpers used by the backward pass */
int b2_relu_gate(float* g, int ldg, const float* y, int ldy, int64_t rows, int c, void* stream);
int b2_slice_copy(float* dst, int ldd, const float* src, int lds, int64_t rows, int c, int accumulate,
                  void* stream);
int b2_add_inplace(float* dst, const float* src, int64_t count, void* stream);
int b2_fill(float* dst, float value, int64_t count, void* stream);
/* per-column bias gradient: db[c] (+)= sum_rows dy[row, c] */
int b2_colsum(const float* dy, int ld, int64_t rows, int c, float* out, int accumulate,
              double* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H */
