// Fused loss kernels of the CutMix mean-teacher hot path (NCHW fp32 logits, one thread per pixel,
// channel loop strided by H*W so every warp access is a coalesced 128 B line):
//   L1  consistency block   train_seg_semisup_mask_mt.py:363-367, 406-420, 428-459
//   L2  supervised CE        train_seg_semisup_mask_mt.py:126, 300-301
//   L1' ICT consistency block train_seg_semisup_ict.py:306-392: per-sample Beta mix factors, teacher
//       PROBABILITIES and confidences mixed (not logits), same five loss functions
// Both write the UNSCALED logit gradient in the same pass and leave the global scalars
// (conf_rate, 1/n_valid, ramp, cons_weight) to a 1-block finalize kernel + the gradient consumer,
// so no host synchronisation is needed (the reference syncs at :413, :461, :469).
#include "common.cuh"
#include <math_constants.h>

constexpr int LOSS_THREADS = 256;

enum { LOSS_VAR = 0, LOSS_LOGITS_VAR = 1, LOSS_LOGITS_SMOOTHL1 = 2, LOSS_BCE = 3, LOSS_KLD = 4 };

// Per-pixel loss q and dq/d(student logits) (g) of the five consistency loss functions, given the teacher target (logits lt,
// probabilities pt) and the student's logits st / soft-max ps (with its row maximum ms and exponential sum sum_s).  Shared by
// the CutMix / ICT kernel and the augmentation-consistency kernel below.
template <int MAXC>
__device__ __forceinline__ float consistency_q_and_grad(const int C, const int loss_fn, const float (&lt)[MAXC],
                                                        const float (&st)[MAXC], const float (&pt)[MAXC],
                                                        const float (&ps)[MAXC], const float ms, const float sum_s,
                                                        float (&g)[MAXC]) {
  float q = 0.f;
  if (loss_fn == LOSS_VAR) {                       // lines 428-431
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float d = ps[c] - pt[c];
      q += d * d;
      g[c] = 2.0f * d;
      dot += g[c] * ps[c];
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) g[c] = ps[c] * (g[c] - dot);
  } else if (loss_fn == LOSS_LOGITS_VAR) {         // lines 432-435
    const float inv = 1.0f / sqrtf((float)C);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float d = st[c] - lt[c];
      q += d * d;
      g[c] = 2.0f * d * inv;
    }
    q *= inv;
  } else if (loss_fn == LOSS_LOGITS_SMOOTHL1) {    // lines 436-439
    const float inv = 1.0f / sqrtf((float)C);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float d = st[c] - lt[c];
      const float ad = fabsf(d);
      if (ad < 1.0f) { q += 0.5f * d * d; g[c] = d * inv; }
      else { q += ad - 0.5f; g[c] = (d > 0.f ? 1.0f : -1.0f) * inv; }
    }
    q *= inv;
  } else if (loss_fn == LOSS_BCE) {                // lines 440-443, network_architectures.py:115-118
    const float eps = 1e-6f;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float t = pt[c], pr = ps[c];
      const float inv_t = 1.0f - t;
      const float inv_p = 1.0f - pr + eps;
      q += -(t * logf(pr + eps) + inv_t * logf(inv_p));
      g[c] = -(t / (pr + eps) - inv_t / inv_p);
      dot += g[c] * pr;
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) g[c] = ps[c] * (g[c] - dot);
  } else {                                         // kld, lines 444-446
    const float lse = logf(sum_s);
    float st_sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float t = pt[c];
      const float logp = (st[c] - ms) - lse;
      const float tlogt = t > 0.f ? t * logf(t) : 0.f;
      q += tlogt - t * logp;
      st_sum += t;
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) g[c] = ps[c] * st_sum - pt[c];
  }
  return q;
}

// EXACT: the class count is the compile-time constant MAXC (19 / 21 / 2: the reference's data sets), so the channel
// loops carry no run-time guards: every load of a pixel's 3*C logits is issued before the first use (with the guarded
// form the compiler serialised them -- three loads in flight per thread, 1.6 ms for 1.3 GB; profiles/r01_v6_*).
// ICT = true (train_seg_semisup_ict.py): `m` holds ONE mix factor per sample (:306-307); the teacher's two softmaxes are
// mixed as probabilities (:329) and as confidences (:340-342), the logits mix (:328) feeds the two logit losses.
// `confbar` (ICT, per-pixel confidence mask): mean over the batch of the confidence mask at this pixel -- the reference's
// `conf_mask[:, None, :, :]` (:344) turns (N,1,H,W) into (N,1,1,H,W), so `loss_mask * conf_mask` broadcasts to
// (N,N,1,H,W) and every sample's loss is weighted by the batch-mean confidence of the pixel (b2_ict_conf_mean).
template <int MAXC, bool EXACT, bool ICT>
__global__ void __launch_bounds__(LOSS_THREADS)
consistency_kernel(const float* __restrict__ l0, const float* __restrict__ l1, const float* __restrict__ ls,
                   const float* __restrict__ m, const float* __restrict__ lmask, const float* __restrict__ confbar,
                   float* __restrict__ dls, double* __restrict__ partials, int C_rt, int64_t hw, int loss_fn,
                   float conf_thresh, int conf_per_pixel) {
  const int C = EXACT ? MAXC : C_rt;
  __shared__ double red[32];
  const int img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  double s_conf = 0.0, s_q = 0.0, s_qc = 0.0;
  if (p < hw) {
    const int64_t base = (int64_t)img * C * hw + p;
    const int64_t pm = (int64_t)img * hw + p;
    float lt[MAXC], st[MAXC];
    const float mv = ICT ? __ldg(m + img) : (m ? __ldg(m + pm) : 0.0f);
    const float om = __fsub_rn(1.0f, mv);
    const float w = lmask ? __ldg(lmask + pm) : 1.0f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { lt[c] = __ldg(l0 + base + (int64_t)c * hw); st[c] = __ldg(ls + base + (int64_t)c * hw); }
    float pt[MAXC], ps[MAXC];
    float pmax = 0.f;
    if (ICT) {
      // teacher: softmax of each view (:322-323), then mix probabilities (:329), confidences (:340-342) and logits (:328)
      float lb[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) lb[c] = __ldg(l1 + base + (int64_t)c * hw);
      float ma = -CUDART_INF_F, mb = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) { ma = fmaxf(ma, lt[c]); mb = fmaxf(mb, lb[c]); }
      float pb[MAXC];
      float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) {
        pt[c] = expf(lt[c] - ma); sum_a += pt[c];
        pb[c] = expf(lb[c] - mb); sum_b += pb[c];
      }
      const float inv_a = 1.0f / sum_a, inv_b = 1.0f / sum_b;
      float conf_a = 0.f, conf_b = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) {
        const float pa = pt[c] * inv_a, pbv = pb[c] * inv_b;
        conf_a = fmaxf(conf_a, pa); conf_b = fmaxf(conf_b, pbv);
        pt[c] = __fadd_rn(__fmul_rn(pa, om), __fmul_rn(pbv, mv));
        lt[c] = __fadd_rn(__fmul_rn(lt[c], om), __fmul_rn(lb[c], mv));
      }
      pmax = __fadd_rn(__fmul_rn(conf_a, om), __fmul_rn(conf_b, mv));
    } else if (l1) {
      float lb[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) lb[c] = __ldg(l1 + base + (int64_t)c * hw);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) lt[c] = __fadd_rn(__fmul_rn(lt[c], om), __fmul_rn(lb[c], mv));  // line 363
    }
    // softmax statistics (lines 366-367)
    float mt = -CUDART_INF_F, ms = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { mt = fmaxf(mt, lt[c]); ms = fmaxf(ms, st[c]); }
    float sum_t = 0.f, sum_s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      if (!ICT) { pt[c] = expf(lt[c] - mt); sum_t += pt[c]; }
      ps[c] = expf(st[c] - ms); sum_s += ps[c];
    }
    // one reciprocal per softmax (the sums lie in [1, C]: always the fast path) instead of C divisions: IEEE division
    // falls into its slow subroutine whenever the numerator is zero / denormal, which is most classes of a confident
    // pixel (3x kernel time on peaked logits, profiles/r01_v8_launch_list_summary.txt); <= 1 ulp from exp / sum
    const float inv_t = ICT ? 1.0f : 1.0f / sum_t, inv_s = 1.0f / sum_s;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      ps[c] = ps[c] * inv_s;
      if (!ICT) { pt[c] = pt[c] * inv_t; pmax = fmaxf(pmax, pt[c]); }
    }
    const float conf = (conf_thresh > 0.0f) ? (pmax >= conf_thresh ? 1.0f : 0.0f) : 1.0f;  // lines 407-411 / ict :343
    // per-pixel loss q and dq/dls (g)
    float g[MAXC];
    const float q = consistency_q_and_grad<MAXC>(C, loss_fn, lt, st, pt, ps, ms, sum_s, g);
    const float cw = (ICT && confbar) ? __ldg(confbar + p) : conf;      // weight of the per-pixel confidence mask
    const float gw = conf_per_pixel ? w * cw : w;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) dls[base + (int64_t)c * hw] = g[c] * gw;
    s_conf = conf;
    s_q = (double)q * (double)w;
    s_qc = s_q * cw;
  }
  const int64_t bid = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  double r;
  r = block_sum_d(s_conf, red); if (threadIdx.x == 0) partials[bid * 3 + 0] = r;
  r = block_sum_d(s_q, red);    if (threadIdx.x == 0) partials[bid * 3 + 1] = r;
  r = block_sum_d(s_qc, red);   if (threadIdx.x == 0) partials[bid * 3 + 2] = r;
}

extern "C" int64_t b2_consistency_num_partials(int n, int64_t hw) {
  return (int64_t)n * ceil_div64(hw, LOSS_THREADS);
}

extern "C" int b2_consistency_fwd_bwd(const float* l0, const float* l1, const float* ls, const float* m,
                                      const float* lmask, float* dls, double* partials, int n, int c,
                                      int64_t hw, int loss_fn, float conf_thresh, int conf_per_pixel,
                                      void* stream) {
  B2_REQUIRE(l0 && ls && dls && partials && n > 0 && c > 0 && hw > 0, "b2_consistency_fwd_bwd: bad args");
  B2_REQUIRE(c <= 64, "b2_consistency_fwd_bwd: C=%d > 64 unsupported", c);
  B2_REQUIRE(loss_fn >= 0 && loss_fn <= 4, "b2_consistency_fwd_bwd: unknown loss_fn %d", loss_fn);
  B2_REQUIRE(!(l1 && !m), "b2_consistency_fwd_bwd: l1 given without mix mask");
  B2_REQUIRE(n <= 65535, "b2_consistency_fwd_bwd: n too large");
  dim3 grid((unsigned)ceil_div64(hw, LOSS_THREADS), n);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(MC, EX) consistency_kernel<MC, EX, false><<<grid, LOSS_THREADS, 0, s>>>(l0, l1, ls, m, lmask, nullptr, dls, partials, c, hw, loss_fn, conf_thresh, conf_per_pixel)
  if (c == 19) LAUNCH(19, true);            // Cityscapes
  else if (c == 21) LAUNCH(21, true);       // Pascal VOC
  else if (c == 2) LAUNCH(2, true);         // ISIC
  else if (c <= 8) LAUNCH(8, false);
  else if (c <= 24) LAUNCH(24, false);
  else if (c <= 32) LAUNCH(32, false);
  else LAUNCH(64, false);
#undef LAUNCH
  B2_LAUNCH_CHECK("consistency_kernel");
  return B2_OK;
}

// ---- ICT (train_seg_semisup_ict.py:306-392) ----------------------------------------------------------------------------
// confbar[p] = mean over the batch of (mix(conf_u0, conf_u1) >= thresh) at pixel p: the weight the reference's
// (N,N,1,H,W) broadcast gives every sample at that pixel when --conf_per_pixel is set (see consistency_kernel).
template <int MAXC, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS)
ict_conf_mean_kernel(const float* __restrict__ l0, const float* __restrict__ l1, const float* __restrict__ lam,
                     float* __restrict__ confbar, int n, int C_rt, int64_t hw, float conf_thresh) {
  const int C = EXACT ? MAXC : C_rt;
  const int64_t p = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  if (p >= hw) return;
  float acc = 0.f;
  for (int img = 0; img < n; ++img) {
    const int64_t base = (int64_t)img * C * hw + p;
    const float mv = __ldg(lam + img), om = __fsub_rn(1.0f, mv);
    float la[MAXC], lb[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { la[c] = __ldg(l0 + base + (int64_t)c * hw); lb[c] = __ldg(l1 + base + (int64_t)c * hw); }
    float ma = -CUDART_INF_F, mb = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { ma = fmaxf(ma, la[c]); mb = fmaxf(mb, lb[c]); }
    float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { la[c] = expf(la[c] - ma); sum_a += la[c]; lb[c] = expf(lb[c] - mb); sum_b += lb[c]; }
    const float inv_a = 1.0f / sum_a, inv_b = 1.0f / sum_b;
    float conf_a = 0.f, conf_b = 0.f;                  // same expressions as consistency_kernel<.., ICT>: identical mask
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { conf_a = fmaxf(conf_a, la[c] * inv_a); conf_b = fmaxf(conf_b, lb[c] * inv_b); }
    const float cf = __fadd_rn(__fmul_rn(conf_a, om), __fmul_rn(conf_b, mv));
    acc += cf >= conf_thresh ? 1.0f : 0.0f;
  }
  confbar[p] = acc / (float)n;
}

extern "C" int b2_ict_conf_mean(const float* l0, const float* l1, const float* lam, float* confbar, int n, int c,
                                int64_t hw, float conf_thresh, void* stream) {
  B2_REQUIRE(l0 && l1 && lam && confbar && n > 0 && c > 0 && hw > 0, "b2_ict_conf_mean: bad args");
  B2_REQUIRE(c <= 64, "b2_ict_conf_mean: C=%d > 64 unsupported", c);
  const unsigned grid = (unsigned)ceil_div64(hw, LOSS_THREADS);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(MC, EX) ict_conf_mean_kernel<MC, EX><<<grid, LOSS_THREADS, 0, s>>>(l0, l1, lam, confbar, n, c, hw, conf_thresh)
  if (c == 19) LAUNCH(19, true);
  else if (c == 21) LAUNCH(21, true);
  else if (c == 2) LAUNCH(2, true);
  else if (c <= 8) LAUNCH(8, false);
  else if (c <= 24) LAUNCH(24, false);
  else if (c <= 32) LAUNCH(32, false);
  else LAUNCH(64, false);
#undef LAUNCH
  B2_LAUNCH_CHECK("ict_conf_mean_kernel");
  return B2_OK;
}

extern "C" int b2_ict_consistency_fwd_bwd(const float* l0, const float* l1, const float* ls, const float* lam,
                                          const float* lmask, const float* confbar, float* dls, double* partials, int n,
                                          int c, int64_t hw, int loss_fn, float conf_thresh, int conf_per_pixel,
                                          void* stream) {
  B2_REQUIRE(l0 && l1 && ls && lam && dls && partials && n > 0 && c > 0 && hw > 0, "b2_ict_consistency_fwd_bwd: bad args");
  B2_REQUIRE(c <= 64, "b2_ict_consistency_fwd_bwd: C=%d > 64 unsupported", c);
  B2_REQUIRE(loss_fn >= 0 && loss_fn <= 4, "b2_ict_consistency_fwd_bwd: unknown loss_fn %d", loss_fn);
  B2_REQUIRE(n <= 65535, "b2_ict_consistency_fwd_bwd: n too large");
  B2_REQUIRE(!(conf_per_pixel && conf_thresh > 0.0f && !confbar),
             "b2_ict_consistency_fwd_bwd: the per-pixel confidence mask needs confbar (b2_ict_conf_mean)");
  dim3 grid((unsigned)ceil_div64(hw, LOSS_THREADS), n);
  cudaStream_t s = (cudaStream_t)stream;
  const float* cb = (conf_per_pixel && conf_thresh > 0.0f) ? confbar : nullptr;
#define LAUNCH(MC, EX) consistency_kernel<MC, EX, true><<<grid, LOSS_THREADS, 0, s>>>(l0, l1, ls, lam, lmask, cb, dls, partials, c, hw, loss_fn, conf_thresh, conf_per_pixel)
  if (c == 19) LAUNCH(19, true);
  else if (c == 21) LAUNCH(21, true);
  else if (c == 2) LAUNCH(2, true);
  else if (c <= 8) LAUNCH(8, false);
  else if (c <= 24) LAUNCH(24, false);
  else if (c <= 32) LAUNCH(32, false);
  else LAUNCH(64, false);
#undef LAUNCH
  B2_LAUNCH_CHECK("consistency_kernel<ict>");
  return B2_OK;
}

// ---- augmentation-driven consistency (train_seg_semisup_aug_mt.py:291-391) --------------------------
// The teacher sees view 0, the student view 1; `theta` (N,2,3) maps student-space normalised coordinates to teacher-space
// ones (the DataLoader's `xf0_to_1`).  The reference builds F.affine_grid(theta, x.shape, align_corners=True) and runs
// F.grid_sample (bilinear, zero padding, align_corners=True) three times: over the teacher's logits (:304), its valid mask
// (:306) and its soft-max probabilities (:312).  Here each student pixel computes its sampling position, gathers the (up to)
// four teacher pixels, takes the soft-max of each and interpolates logits, probabilities and mask in registers.

struct BilinearTap { int off[4]; float wt[4]; };     // teacher pixel offsets (y*W + x; -1 = outside: contributes zero)

// torch.linspace(-1, 1, n)[i] as at::linspace computes it (first half counted from the start, second half from the end);
// affine_grid uses 0 for a single-element axis.
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  if (n <= 1) return 0.0f;
  const float step = 2.0f / (float)(n - 1);
  return i < n / 2 ? -1.0f + step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

// Sampling position of output pixel (y, x) of an (OH, OW) grid in an (IH, IW) image and its bilinear taps in the order
// nw, ne, sw, se (grid_sampler: unnormalise with align_corners=True, floor, weights (x_e - x)(y_s - y) ...).
__device__ __forceinline__ BilinearTap affine_bilinear_tap(const float* __restrict__ th, int y, int x, int OH, int OW,
                                                           int IH, int IW) {
  const float xb = linspace_pm1(x, OW), yb = linspace_pm1(y, OH);
  const float gx = xb * __ldg(th + 0) + yb * __ldg(th + 1) + __ldg(th + 2);      // base_grid (x, y, 1) . theta^T
  const float gy = xb * __ldg(th + 3) + yb * __ldg(th + 4) + __ldg(th + 5);
  const float ix = ((gx + 1.0f) * 0.5f) * (float)(IW - 1);
  const float iy = ((gy + 1.0f) * 0.5f) * (float)(IH - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const float we = ix - fx, ww = 1.0f - we, ws = iy - fy, wn = 1.0f - ws;
  // clamp before the int conversion: positions far outside (or NaN) must not overflow; they only ever select "outside"
  const int x0 = (int)fminf(fmaxf(fx, -2.0f), (float)IW), y0 = (int)fminf(fmaxf(fy, -2.0f), (float)IH);
  BilinearTap t;
  t.wt[0] = wn * ww; t.wt[1] = wn * we; t.wt[2] = ws * ww; t.wt[3] = ws * we;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    t.off[k] = (xx >= 0 && xx < IW && yy >= 0 && yy < IH) ? yy * IW + xx : -1;
  }
  return t;
}

// y[n, c, oy, ox] = bilinear sample of x[n, c] at theta[n] . (ox, oy, 1): F.grid_sample(x, F.affine_grid(theta, (N,C,OH,OW),
// align_corners=True), align_corners=True) with the default bilinear mode and zero padding.
__global__ void __launch_bounds__(LOSS_THREADS)
affine_grid_sample_kernel(const float* __restrict__ x, const float* __restrict__ theta, float* __restrict__ y, int C,
                          int IH, int IW, int OH, int OW) {
  const int img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  const int64_t ohw = (int64_t)OH * OW, ihw = (int64_t)IH * IW;
  if (p >= ohw) return;
  const BilinearTap t = affine_bilinear_tap(theta + img * 6, (int)(p / OW), (int)(p % OW), OH, OW, IH, IW);
  for (int c = 0; c < C; ++c) {
    const float* xc = x + ((int64_t)img * C + c) * ihw;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (t.off[k] >= 0) acc += __ldg(xc + t.off[k]) * t.wt[k];
    y[((int64_t)img * C + c) * ohw + p] = acc;
  }
}

extern "C" int b2_affine_grid_sample(const float* x, const float* theta, float* y, int n, int c, int ih, int iw, int oh,
                                     int ow, void* stream) {
  B2_REQUIRE(x && theta && y && n > 0 && c > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "b2_affine_grid_sample: bad args");
  B2_REQUIRE(n <= 65535, "b2_affine_grid_sample: n too large");
  B2_REQUIRE((int64_t)ih * iw < (1ll << 31) && (int64_t)oh * ow < (1ll << 31), "b2_affine_grid_sample: image too large");
  dim3 grid((unsigned)ceil_div64((int64_t)oh * ow, LOSS_THREADS), n);
  affine_grid_sample_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>(x, theta, y, c, ih, iw, oh, ow);
  B2_LAUNCH_CHECK("affine_grid_sample_kernel");
  return B2_OK;
}

template <int MAXC, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS)
aug_consistency_kernel(const float* __restrict__ ltea, const float* __restrict__ ls, const float* __restrict__ theta,
                       const float* __restrict__ um0, const float* __restrict__ um1, float* __restrict__ dls,
                       double* __restrict__ partials, int C_rt, int H, int W, int loss_fn, float conf_thresh,
                       int conf_per_pixel) {
  const int C = EXACT ? MAXC : C_rt;
  __shared__ double red[32];
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  double s_conf = 0.0, s_q = 0.0, s_qc = 0.0;
  if (p < hw) {
    const int64_t base = (int64_t)img * C * hw + p;
    const int64_t pm = (int64_t)img * hw + p;
    const BilinearTap t = affine_bilinear_tap(theta + img * 6, (int)(p / W), (int)(p % W), H, W, H, W);
    float st[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) st[c] = __ldg(ls + base + (int64_t)c * hw);
    // teacher target in student space: logits (:304), probabilities (:309, :312) and valid mask (:306) interpolated from the
    // four teacher pixels, accumulated in grid_sample's order nw, ne, sw, se
    float lt[MAXC], pt[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { lt[c] = 0.f; pt[c] = 0.f; }
    float mt0 = 0.f;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      if (t.off[k] < 0) continue;
      const float wk = t.wt[k];
      const float* src = ltea + (int64_t)img * C * hw + t.off[k];
      float lk[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) lk[c] = __ldg(src + (int64_t)c * hw);
      mt0 += __ldg(um0 + (int64_t)img * hw + t.off[k]) * wk;
      float mk = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) mk = fmaxf(mk, lk[c]);
      float sum_k = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) { lt[c] += lk[c] * wk; lk[c] = expf(lk[c] - mk); sum_k += lk[c]; }
      const float inv_k = 1.0f / sum_k;              // one reciprocal per soft-max, see consistency_kernel
#pragma unroll
      for (int c = 0; c < MAXC; ++c) if (c < C) pt[c] += (lk[c] * inv_k) * wk;
    }
    const float w = mt0 * __ldg(um1 + pm);           // mask_tea_in_stu (:306) = loss_mask (:341)
    // student soft-max (:310)
    float ms = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) ms = fmaxf(ms, st[c]);
    float ps[MAXC];
    float sum_s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { ps[c] = expf(st[c] - ms); sum_s += ps[c]; }
    const float inv_s = 1.0f / sum_s;
    float pmax = 0.f;                                // pt >= 0: max over classes of the interpolated probabilities (:347)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { ps[c] = ps[c] * inv_s; pmax = fmaxf(pmax, pt[c]); }
    const float conf = (conf_thresh > 0.0f) ? (pmax >= conf_thresh ? 1.0f : 0.0f) : 1.0f;      // :345-352
    float g[MAXC];
    const float q = consistency_q_and_grad<MAXC>(C, loss_fn, lt, st, pt, ps, ms, sum_s, g);    // :366-387
    const float gw = conf_per_pixel ? w * conf : w;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) dls[base + (int64_t)c * hw] = g[c] * gw;
    s_conf = conf;
    s_q = (double)q * (double)w;
    s_qc = s_q * conf;
  }
  const int64_t bid = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  double r;
  r = block_sum_d(s_conf, red); if (threadIdx.x == 0) partials[bid * 3 + 0] = r;
  r = block_sum_d(s_q, red);    if (threadIdx.x == 0) partials[bid * 3 + 1] = r;
  r = block_sum_d(s_qc, red);   if (threadIdx.x == 0) partials[bid * 3 + 2] = r;
}

extern "C" int b2_aug_consistency_fwd_bwd(const float* ltea, const float* ls, const float* theta, const float* um0,
                                          const float* um1, float* dls, double* partials, int n, int c, int h, int w,
                                          int loss_fn, float conf_thresh, int conf_per_pixel, void* stream) {
  B2_REQUIRE(ltea && ls && theta && um0 && um1 && dls && partials && n > 0 && c > 0 && h > 0 && w > 0,
             "b2_aug_consistency_fwd_bwd: bad args");
  B2_REQUIRE(c <= 64, "b2_aug_consistency_fwd_bwd: C=%d > 64 unsupported", c);
  B2_REQUIRE(loss_fn >= 0 && loss_fn <= 4, "b2_aug_consistency_fwd_bwd: unknown loss_fn %d", loss_fn);
  B2_REQUIRE(n <= 65535, "b2_aug_consistency_fwd_bwd: n too large");
  B2_REQUIRE((int64_t)h * w < (1ll << 31), "b2_aug_consistency_fwd_bwd: image too large");
  dim3 grid((unsigned)ceil_div64((int64_t)h * w, LOSS_THREADS), n);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(MC, EX) aug_consistency_kernel<MC, EX><<<grid, LOSS_THREADS, 0, s>>>(ltea, ls, theta, um0, um1, dls, partials, c, h, w, loss_fn, conf_thresh, conf_per_pixel)
  if (c == 19) LAUNCH(19, true);
  else if (c == 21) LAUNCH(21, true);
  else if (c == 2) LAUNCH(2, true);
  else if (c <= 8) LAUNCH(8, false);
  else if (c <= 24) LAUNCH(24, false);
  else if (c <= 32) LAUNCH(32, false);
  else LAUNCH(64, false);
#undef LAUNCH
  B2_LAUNCH_CHECK("aug_consistency_kernel");
  return B2_OK;
}

__global__ void __launch_bounds__(1024)
consistency_finalize_kernel(const double* __restrict__ partials, int64_t n_partials, int64_t n_pixels,
                            float conf_thresh, int conf_per_pixel, float ramp, float cons_weight,
                            float* __restrict__ out4) {
  __shared__ double red[32];
  double a = 0, b = 0, c = 0;
  for (int64_t i = threadIdx.x; i < n_partials; i += blockDim.x) {
    a += partials[i * 3 + 0]; b += partials[i * 3 + 1]; c += partials[i * 3 + 2];
  }
  a = block_sum_d(a, red); b = block_sum_d(b, red); c = block_sum_d(c, red);
  if (threadIdx.x == 0) {
    const double P = (double)n_pixels;
    const double conf_rate = a / P;
    double loss, gscale;
    if (conf_thresh > 0.0f && !conf_per_pixel) {   // scalar-mean mask, line 415-416
      loss = conf_rate * (b / P);
      gscale = conf_rate / P;
    } else {                                       // per-pixel mask (or no thresholding: conf == 1)
      loss = c / P;
      gscale = 1.0 / P;
    }
    loss *= (double)ramp;                          // line 454-455
    out4[0] = (float)loss;
    out4[1] = (float)conf_rate;
    out4[2] = (float)(gscale * (double)ramp * (double)cons_weight);
    out4[3] = (float)(loss * (double)cons_weight); // line 458
  }
}

extern "C" int b2_consistency_finalize(const double* partials, int64_t n_partials, int64_t n_pixels,
                                       float conf_thresh, int conf_per_pixel, float ramp, float cons_weight,
                                       float* out4, void* stream) {
  B2_REQUIRE(partials && out4 && n_partials > 0 && n_pixels > 0, "b2_consistency_finalize: bad args");
  consistency_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n_partials, n_pixels, conf_thresh,
                                                                   conf_per_pixel, ramp, cons_weight, out4);
  B2_LAUNCH_CHECK("consistency_finalize_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// Cross entropy with ignore_index.
// ------------------------------------------------------------------------------------------
template <int MAXC, bool EXACT>
__global__ void __launch_bounds__(LOSS_THREADS)
ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ dlogits,
          double* __restrict__ partials, int C_rt, int64_t hw, int64_t ignore_index) {
  const int C = EXACT ? MAXC : C_rt;
  __shared__ double red[32];
  const int img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  double s_nll = 0.0, s_valid = 0.0;
  if (p < hw) {
    const int64_t base = (int64_t)img * C * hw + p;
    const int64_t lab = __ldg(labels + (int64_t)img * hw + p);
    const bool valid = lab != ignore_index;
    float x[MAXC];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { x[c] = __ldg(logits + base + (int64_t)c * hw); mx = fmaxf(mx, x[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) { x[c] = x[c] - mx; sum += expf(x[c]); }
    const float lse = logf(sum);
    float xl = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) if (c < C) {
      const float logp = x[c] - lse;
      float gsm = valid ? expf(logp) : 0.f;
      if (valid && (int64_t)c == lab) { xl = logp; gsm -= 1.0f; }
      dlogits[base + (int64_t)c * hw] = gsm;
    }
    if (valid) { s_nll = -(double)xl; s_valid = 1.0; }
  }
  const int64_t bid = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  double r;
  r = block_sum_d(s_nll, red);   if (threadIdx.x == 0) partials[bid * 2 + 0] = r;
  r = block_sum_d(s_valid, red); if (threadIdx.x == 0) partials[bid * 2 + 1] = r;
}

extern "C" int64_t b2_ce_num_partials(int n, int64_t hw) { return (int64_t)n * ceil_div64(hw, LOSS_THREADS); }

extern "C" int b2_ce_fwd_bwd(const float* logits, const int64_t* labels, float* dlogits, double* partials,
                             int n, int c, int64_t hw, int64_t ignore_index, void* stream) {
  B2_REQUIRE(logits && labels && dlogits && partials && n > 0 && c > 0 && hw > 0, "b2_ce_fwd_bwd: bad args");
  B2_REQUIRE(c <= 64, "b2_ce_fwd_bwd: C=%d > 64 unsupported", c);
  B2_REQUIRE(n <= 65535, "b2_ce_fwd_bwd: n too large");
  dim3 grid((unsigned)ceil_div64(hw, LOSS_THREADS), n);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(MC, EX) ce_kernel<MC, EX><<<grid, LOSS_THREADS, 0, s>>>(logits, labels, dlogits, partials, c, hw, ignore_index)
  if (c == 19) LAUNCH(19, true);
  else if (c == 21) LAUNCH(21, true);
  else if (c == 2) LAUNCH(2, true);
  else if (c <= 8) LAUNCH(8, false);
  else if (c <= 24) LAUNCH(24, false);
  else if (c <= 32) LAUNCH(32, false);
  else LAUNCH(64, false);
#undef LAUNCH
  B2_LAUNCH_CHECK("ce_kernel");
  return B2_OK;
}

__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const double* __restrict__ partials, int64_t n_partials, float* __restrict__ out3) {
  __shared__ double red[32];
  double a = 0, b = 0;
  for (int64_t i = threadIdx.x; i < n_partials; i += blockDim.x) { a += partials[i * 2 + 0]; b += partials[i * 2 + 1]; }
  a = block_sum_d(a, red); b = block_sum_d(b, red);
  if (threadIdx.x == 0) {
    out3[0] = (float)(a / b);          // mean over non-ignored pixels (NaN if none, like PyTorch)
    out3[1] = (float)b;
    out3[2] = b > 0 ? (float)(1.0 / b) : 0.f;
  }
}

extern "C" int b2_ce_finalize(const double* partials, int64_t n_partials, float* out3, void* stream) {
  B2_REQUIRE(partials && out3 && n_partials > 0, "b2_ce_finalize: bad args");
  ce_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n_partials, out3);
  B2_LAUNCH_CHECK("ce_finalize_kernel");
  return B2_OK;
}
