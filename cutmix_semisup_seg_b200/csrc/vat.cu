// Kernels of the VAT variant of the loop (train_seg_semisup_vat_mt.py:214-301): everything the virtual
// adversarial perturbation needs besides the network passes and the consistency kernel (csrc/losses.cu):
//   b2_col2im                 input gradient of the Cin = 3 stem (the stem runs as im2col + GEMM, csrc/netops.cu): the adjoint
//                             of b2_im2col, gather form (no atomics, deterministic)
//   b2_sample_l2norm          mag[i] = sqrt(sum of squares of sample i)            normalize_eps, :217-220
//   b2_vat_adaptive_radius    radius[i] = vat_radius * sqrt(|dx/dv|^2 + |dx/dh|^2) * 0.5 from central differences, :277-296
//   b2_add_scaled_per_sample  out = x + (e / (mag[i] + 1e-12)) * radius[i]        :220, :222-223, :247, :301, :392
// All HBM-bound, one pass each; reductions are two-stage with a fixed summation order (doubles), so results do not depend on
// the launch geometry.
#include "common.cuh"

constexpr int VAT_THREADS = 256;
constexpr int VAT_ITEMS = 16;                       // elements per thread of a reduction block (4096 per block)

// ---------------------------------------------------------------------------------------------------- col2im
// dx[n, y, x, ch] (+)= sum over the filter taps (r, s) whose output position (oy, ox) = ((y + pad - r*dil) / stride,
// (x + pad - s*dil) / stride) exists of dcol[(n*OH + oy)*OW + ox][(r*KW + s)*C + ch]   (column order of b2_im2col).
template <int MAXC>
__global__ void __launch_bounds__(VAT_THREADS)
col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int n, int h, int w, int c, int ldx, int kh, int kw,
              int stride, int pad, int dil, int oh, int ow, int kpad, int accumulate) {
  const int64_t total = (int64_t)n * h * w;
  const int64_t p = (int64_t)blockIdx.x * VAT_THREADS + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % w);
  const int64_t t = p / w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  float acc[MAXC];
#pragma unroll
  for (int ch = 0; ch < MAXC; ++ch) acc[ch] = 0.f;
  for (int r = 0; r < kh; ++r) {
    const int ty = y + pad - r * dil;
    if (ty < 0 || ty % stride) continue;
    const int oy = ty / stride;
    if (oy >= oh) continue;
    for (int s = 0; s < kw; ++s) {
      const int tx = x + pad - s * dil;
      if (tx < 0 || tx % stride) continue;
      const int ox = tx / stride;
      if (ox >= ow) continue;
      const float* src = dcol + (((int64_t)img * oh + oy) * ow + ox) * kpad + (r * kw + s) * c;
#pragma unroll
      for (int ch = 0; ch < MAXC; ++ch) if (ch < c) acc[ch] += __ldg(src + ch);
    }
  }
  float* dst = dx + p * ldx;
#pragma unroll
  for (int ch = 0; ch < MAXC; ++ch) if (ch < c) dst[ch] = accumulate ? dst[ch] + acc[ch] : acc[ch];
}

extern "C" int b2_col2im(const float* dcol, float* dx, int n, int h, int w, int c, int ldx, int kh, int kw, int stride,
                         int pad, int dil, int oh, int ow, int kpad, int accumulate, void* stream) {
  B2_REQUIRE(dcol && dx && n > 0 && h > 0 && w > 0 && c > 0 && ldx >= c && kh > 0 && kw > 0 && stride > 0 && dil > 0 &&
             pad >= 0 && oh > 0 && ow > 0 && kpad >= kh * kw * c, "b2_col2im: bad args");
  B2_REQUIRE(c <= 4, "b2_col2im: C=%d > 4 unsupported (stem convolutions only)", c);
  const int64_t total = (int64_t)n * h * w;
  col2im_kernel<4><<<(unsigned)ceil_div64(total, VAT_THREADS), VAT_THREADS, 0, (cudaStream_t)stream>>>(
      dcol, dx, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad, accumulate);
  B2_LAUNCH_CHECK("col2im_kernel");
  return B2_OK;
}

// ---------------------------------------------------------------------------------------------------- per-sample reductions
extern "C" int64_t b2_sample_reduce_blocks(int64_t per) { return ceil_div64(per, (int64_t)VAT_THREADS * VAT_ITEMS); }

// partials[(img * blocks + block) * 2 + {0, 1}]
__global__ void __launch_bounds__(VAT_THREADS)
sample_sumsq_kernel(const float* __restrict__ x, int64_t per, double* __restrict__ partials) {
  __shared__ double red[32];
  const int img = blockIdx.y;
  const float* xs = x + (int64_t)img * per;
  const int64_t i0 = (int64_t)blockIdx.x * VAT_THREADS * VAT_ITEMS + threadIdx.x;
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < VAT_ITEMS; ++k) {
    const int64_t i = i0 + (int64_t)k * VAT_THREADS;
    if (i < per) { const float v = __ldg(xs + i); acc += (double)v * (double)v; }
  }
  const double r = block_sum_d(acc, red);
  if (threadIdx.x == 0) {
    const int64_t b = ((int64_t)img * gridDim.x + blockIdx.x) * 2;
    partials[b] = r; partials[b + 1] = 0.0;
  }
}

// central differences with step 2 (:289-290): vertical x[.., y+2, x] - x[.., y, x] for y < H-2, horizontal likewise
__global__ void __launch_bounds__(VAT_THREADS)
jacobian_sumsq_kernel(const float* __restrict__ x, int h, int w, int64_t per, double* __restrict__ partials) {
  __shared__ double red[32];
  const int img = blockIdx.y;
  const float* xs = x + (int64_t)img * per;
  const int64_t i0 = (int64_t)blockIdx.x * VAT_THREADS * VAT_ITEMS + threadIdx.x;
  double av = 0.0, ah = 0.0;
#pragma unroll 4
  for (int k = 0; k < VAT_ITEMS; ++k) {
    const int64_t i = i0 + (int64_t)k * VAT_THREADS;
    if (i < per) {
      const int xx = (int)(i % w), yy = (int)((i / w) % h);
      const float v = __ldg(xs + i);
      if (yy + 2 < h) { const float d = __fsub_rn(__ldg(xs + i + 2 * (int64_t)w), v); av += (double)d * (double)d; }
      if (xx + 2 < w) { const float d = __fsub_rn(__ldg(xs + i + 2), v); ah += (double)d * (double)d; }
    }
  }
  const double rv = block_sum_d(av, red);
  const double rh = block_sum_d(ah, red);
  if (threadIdx.x == 0) {
    const int64_t b = ((int64_t)img * gridDim.x + blockIdx.x) * 2;
    partials[b] = rv; partials[b + 1] = rh;
  }
}

// mode 0: out[i] = sqrt(S0)   (normalize_eps' `mag`);  mode 1: out[i] = vat_radius * sqrt(S0 + S1) * 0.5  (:296), each
// product rounded to fp32 like the reference's tensor expression
__global__ void sample_reduce_finalize_kernel(const double* __restrict__ partials, int64_t blocks, int n, int mode,
                                              float vat_radius, float* __restrict__ out) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n) return;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t b = 0; b < blocks; ++b) { s0 += partials[((int64_t)img * blocks + b) * 2]; s1 += partials[((int64_t)img * blocks + b) * 2 + 1]; }
  if (mode == 0) out[img] = sqrtf((float)s0);
  else out[img] = __fmul_rn(__fmul_rn(vat_radius, sqrtf(__fadd_rn((float)s0, (float)s1))), 0.5f);
}

extern "C" int b2_sample_l2norm(const float* x, int n, int64_t per, double* partials, float* mag, void* stream) {
  B2_REQUIRE(x && partials && mag && n > 0 && per > 0, "b2_sample_l2norm: bad args");
  B2_REQUIRE(n <= 65535, "b2_sample_l2norm: n too large");
  const int64_t blocks = b2_sample_reduce_blocks(per);
  cudaStream_t s = (cudaStream_t)stream;
  sample_sumsq_kernel<<<dim3((unsigned)blocks, n), VAT_THREADS, 0, s>>>(x, per, partials);
  B2_LAUNCH_CHECK("sample_sumsq_kernel");
  sample_reduce_finalize_kernel<<<(n + 127) / 128, 128, 0, s>>>(partials, blocks, n, 0, 0.f, mag);
  B2_LAUNCH_CHECK("sample_reduce_finalize_kernel");
  return B2_OK;
}

extern "C" int b2_vat_adaptive_radius(const float* x, int n, int c, int h, int w, float vat_radius, double* partials,
                                      float* radius, void* stream) {
  B2_REQUIRE(x && partials && radius && n > 0 && c > 0 && h > 0 && w > 0, "b2_vat_adaptive_radius: bad args");
  B2_REQUIRE(n <= 65535, "b2_vat_adaptive_radius: n too large");
  const int64_t per = (int64_t)c * h * w;
  const int64_t blocks = b2_sample_reduce_blocks(per);
  cudaStream_t s = (cudaStream_t)stream;
  jacobian_sumsq_kernel<<<dim3((unsigned)blocks, n), VAT_THREADS, 0, s>>>(x, h, w, per, partials);
  B2_LAUNCH_CHECK("jacobian_sumsq_kernel");
  sample_reduce_finalize_kernel<<<(n + 127) / 128, 128, 0, s>>>(partials, blocks, n, 1, vat_radius, radius);
  B2_LAUNCH_CHECK("sample_reduce_finalize_kernel");
  return B2_OK;
}

// ---------------------------------------------------------------------------------------------------- perturbation
// out = fl(x + fl(fl(e / fl(mag[i] + 1e-12)) * r_i)),  r_i = radius[i] or the host scalar: one rounding per operation of the
// reference's tensor expression (normalize_eps, `* adv_radius` / `* scale`, `x + perturbation`).  x == NULL: out = the scaled
// direction alone.
__global__ void __launch_bounds__(VAT_THREADS)
add_scaled_per_sample_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ mag,
                             const float* __restrict__ radius, float radius_host, float* __restrict__ out, int64_t per) {
  const int img = blockIdx.y;
  const float den = __fadd_rn(__ldg(mag + img), 1e-12f);
  const float r = radius ? __ldg(radius + img) : radius_host;
  const int64_t base = (int64_t)img * per;
  for (int64_t i = (int64_t)blockIdx.x * VAT_THREADS + threadIdx.x; i < per; i += (int64_t)gridDim.x * VAT_THREADS) {
    const float d = __fmul_rn(__fdiv_rn(__ldg(e + base + i), den), r);
    out[base + i] = x ? __fadd_rn(__ldg(x + base + i), d) : d;
  }
}

extern "C" int b2_add_scaled_per_sample(const float* x, const float* e, const float* mag, const float* radius,
                                        float radius_host, float* out, int n, int64_t per, void* stream) {
  B2_REQUIRE(e && mag && out && n > 0 && per > 0, "b2_add_scaled_per_sample: bad args");
  B2_REQUIRE(n <= 65535, "b2_add_scaled_per_sample: n too large");
  int64_t bx = ceil_div64(per, VAT_THREADS);
  if (bx > 4096) bx = 4096;
  add_scaled_per_sample_kernel<<<dim3((unsigned)bx, n), VAT_THREADS, 0, (cudaStream_t)stream>>>(x, e, mag, radius, radius_host,
                                                                                              out, per);
  B2_LAUNCH_CHECK("add_scaled_per_sample_kernel");
  return B2_OK;
}
