// Decoder operators of the reference's U-Net style architectures (architectures/resunet.py:10-34, 57-92 and the identical
// DecoderBlock / final layers of architectures/denseunet.py:10-34, 98-124), NHWC fp32 with a leading dimension:
//   b2_upsample2x_add       y = nearest-neighbour x2 up-sampling of x (nn.Upsample(scale_factor=2)) [+ skip]   resunet.py:31-32
//   b2_upsample2x_bwd       dx (+)= sum of the 2x2 block of dy (the adjoint); d(skip) = dy needs no kernel
//   b2_mul_mask             y = x * mask * scale: nn.Dropout applied to a raw convolution output (resunet.py:88), fwd and bwd
// All HBM-bound single passes; float4 paths when channels and leading dimensions allow.
#include "common.cuh"

constexpr int UN_THREADS = 256;

template <int VEC>
__global__ void __launch_bounds__(UN_THREADS)
upsample2x_add_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ skip, int lds, float* __restrict__ y,
                      int ldy, int n, int h, int w, int c) {
  const int cv = c / VEC;
  const int64_t total = (int64_t)n * (2 * h) * (2 * w) * cv;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int ch = (int)(i % cv) * VEC;
    const int64_t pix = i / cv;                       // output pixel index (n, oy, ox)
    const int ox = (int)(pix % (2 * w));
    const int64_t t = pix / (2 * w);
    const int oy = (int)(t % (2 * h));
    const int img = (int)(t / (2 * h));
    const int64_t src = ((int64_t)img * h + (oy >> 1)) * w + (ox >> 1);
    if (VEC == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x + src * ldx + ch));
      if (skip) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(skip + pix * lds + ch));
        v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
      }
      *reinterpret_cast<float4*>(y + pix * ldy + ch) = v;
    } else {
      float v = __ldg(x + src * ldx + ch);
      if (skip) v += __ldg(skip + pix * lds + ch);
      y[pix * ldy + ch] = v;
    }
  }
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline unsigned grid_for(int64_t total) {
  int64_t b = ceil_div64(total, UN_THREADS);
  const int64_t cap = (int64_t)148 * 32;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

extern "C" int b2_upsample2x_add(const float* x, int ldx, const float* skip, int lds, float* y, int ldy, int n, int h, int w,
                                 int c, void* stream) {
  B2_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && ldx >= c && ldy >= c && (!skip || lds >= c),
             "b2_upsample2x_add: bad args");
  const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16p(x) && al16p(y) && (!skip || (lds % 4 == 0 && al16p(skip)));
  const int64_t total = (int64_t)n * 4 * h * w * (vec ? c / 4 : c);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) upsample2x_add_kernel<4><<<grid_for(total), UN_THREADS, 0, s>>>(x, ldx, skip, lds, y, ldy, n, h, w, c);
  else upsample2x_add_kernel<1><<<grid_for(total), UN_THREADS, 0, s>>>(x, ldx, skip, lds, y, ldy, n, h, w, c);
  B2_LAUNCH_CHECK("upsample2x_add_kernel");
  return B2_OK;
}

template <int VEC>
__global__ void __launch_bounds__(UN_THREADS)
upsample2x_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx, int n, int h, int w, int c,
                      int accumulate) {
  const int cv = c / VEC;
  const int64_t total = (int64_t)n * h * w * cv;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int ch = (int)(i % cv) * VEC;
    const int64_t pix = i / cv;                       // input pixel index (n, iy, ix)
    const int ix = (int)(pix % w);
    const int64_t t = pix / w;
    const int iy = (int)(t % h);
    const int img = (int)(t / h);
    const int64_t o00 = ((int64_t)img * 2 * h + 2 * iy) * (2 * w) + 2 * ix;      // top-left output pixel of the 2x2 block
    const int64_t o10 = o00 + 2 * w;
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(dy + o00 * lddy + ch));
      const float4 b = __ldg(reinterpret_cast<const float4*>(dy + (o00 + 1) * lddy + ch));
      const float4 cc = __ldg(reinterpret_cast<const float4*>(dy + o10 * lddy + ch));
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (o10 + 1) * lddy + ch));
      float4 r = make_float4((a.x + b.x) + (cc.x + d.x), (a.y + b.y) + (cc.y + d.y), (a.z + b.z) + (cc.z + d.z),
                             (a.w + b.w) + (cc.w + d.w));
      float4* dst = reinterpret_cast<float4*>(dx + pix * lddx + ch);
      if (accumulate) { const float4 o = *dst; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
      *dst = r;
    } else {
      float r = (__ldg(dy + o00 * lddy + ch) + __ldg(dy + (o00 + 1) * lddy + ch)) +
                (__ldg(dy + o10 * lddy + ch) + __ldg(dy + (o10 + 1) * lddy + ch));
      float* dst = dx + pix * lddx + ch;
      *dst = accumulate ? *dst + r : r;
    }
  }
}

extern "C" int b2_upsample2x_bwd(const float* dy, int lddy, float* dx, int lddx, int n, int h, int w, int c, int accumulate,
                                 void* stream) {
  B2_REQUIRE(dy && dx && n > 0 && h > 0 && w > 0 && c > 0 && lddy >= c && lddx >= c, "b2_upsample2x_bwd: bad args");
  const bool vec = c % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && al16p(dy) && al16p(dx);
  const int64_t total = (int64_t)n * h * w * (vec ? c / 4 : c);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) upsample2x_bwd_kernel<4><<<grid_for(total), UN_THREADS, 0, s>>>(dy, lddy, dx, lddx, n, h, w, c, accumulate);
  else upsample2x_bwd_kernel<1><<<grid_for(total), UN_THREADS, 0, s>>>(dy, lddy, dx, lddx, n, h, w, c, accumulate);
  B2_LAUNCH_CHECK("upsample2x_bwd_kernel");
  return B2_OK;
}

// y[r, ch] = x[r, ch] * mask[r * c + ch] * scale (mask dense (rows, c))
__global__ void __launch_bounds__(UN_THREADS)
mul_mask_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mask, float scale, float* __restrict__ y,
                int ldy, int64_t rows, int c) {
  const int64_t total = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    y[r * ldy + ch] = __fmul_rn(__fmul_rn(__ldg(x + r * ldx + ch), __ldg(mask + i)), scale);
  }
}

extern "C" int b2_mul_mask(const float* x, int ldx, const float* mask, float scale, float* y, int ldy, int64_t rows, int c,
                           void* stream) {
  B2_REQUIRE(x && mask && y && rows > 0 && c > 0 && ldx >= c && ldy >= c, "b2_mul_mask: bad args");
  mul_mask_kernel<<<grid_for(rows * c), UN_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, mask, scale, y, ldy, rows, c);
  B2_LAUNCH_CHECK("mul_mask_kernel");
  return B2_OK;
}

// ---------------------------------------------------------------------------------------------------- DenseNet encoder glue
// (architectures/denseunet.py: torchvision densenet161 `features`)
//   b2_avgpool2x2       transition layers' nn.AvgPool2d(kernel_size=2, stride=2): y[n,h,w,c] = mean of the 2x2 block of x
//   b2_avgpool2x2_bwd   dx[n,2h+a,2w+b,c] (+)= 0.25 * dy[n,h,w,c]
//   b2_scale_channels   dst[r,c] (+)= g[r,c] * scale[c]: backward of a stand-alone eval-mode BatchNorm (the pre-activation
//                       norm1 / transition norm of DenseNet act on a concatenation and cannot be folded into a producing
//                       convolution), accumulating into a channel prefix of the concatenation's gradient
template <int VEC>
__global__ void __launch_bounds__(UN_THREADS)
avgpool2x2_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int n, int oh, int ow, int ih, int iw,
                  int c) {
  const int cv = c / VEC;
  const int64_t total = (int64_t)n * oh * ow * cv;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int ch = (int)(i % cv) * VEC;
    const int64_t pix = i / cv;
    const int ox = (int)(pix % ow);
    const int64_t t = pix / ow;
    const int oy = (int)(t % oh);
    const int img = (int)(t / oh);
    const int64_t i00 = ((int64_t)img * ih + 2 * oy) * iw + 2 * ox;
    const int64_t i10 = i00 + iw;
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + i00 * ldx + ch));
      const float4 b = __ldg(reinterpret_cast<const float4*>(x + (i00 + 1) * ldx + ch));
      const float4 cc = __ldg(reinterpret_cast<const float4*>(x + i10 * ldx + ch));
      const float4 d = __ldg(reinterpret_cast<const float4*>(x + (i10 + 1) * ldx + ch));
      *reinterpret_cast<float4*>(y + pix * ldy + ch) = make_float4(
          ((a.x + b.x) + (cc.x + d.x)) * 0.25f, ((a.y + b.y) + (cc.y + d.y)) * 0.25f, ((a.z + b.z) + (cc.z + d.z)) * 0.25f,
          ((a.w + b.w) + (cc.w + d.w)) * 0.25f);
    } else {
      y[pix * ldy + ch] = ((__ldg(x + i00 * ldx + ch) + __ldg(x + (i00 + 1) * ldx + ch)) +
                           (__ldg(x + i10 * ldx + ch) + __ldg(x + (i10 + 1) * ldx + ch))) * 0.25f;
    }
  }
}

extern "C" int b2_avgpool2x2(const float* x, int ldx, float* y, int ldy, int n, int ih, int iw, int c, void* stream) {
  B2_REQUIRE(x && y && n > 0 && ih >= 2 && iw >= 2 && c > 0 && ldx >= c && ldy >= c, "b2_avgpool2x2: bad args");
  const int oh = ih / 2, ow = iw / 2;                       // floor, like nn.AvgPool2d(2, 2)
  const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16p(x) && al16p(y);
  const int64_t total = (int64_t)n * oh * ow * (vec ? c / 4 : c);
  cudaStream_t s = (cudaStream_t)stream;
  if (vec) avgpool2x2_kernel<4><<<grid_for(total), UN_THREADS, 0, s>>>(x, ldx, y, ldy, n, oh, ow, ih, iw, c);
  else avgpool2x2_kernel<1><<<grid_for(total), UN_THREADS, 0, s>>>(x, ldx, y, ldy, n, oh, ow, ih, iw, c);
  B2_LAUNCH_CHECK("avgpool2x2_kernel");
  return B2_OK;
}

__global__ void __launch_bounds__(UN_THREADS)
avgpool2x2_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx, int n, int oh, int ow, int ih,
                      int iw, int c, int accumulate) {
  const int64_t total = (int64_t)n * ih * iw * c;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int ch = (int)(i % c);
    const int64_t pix = i / c;
    const int ix = (int)(pix % iw);
    const int64_t t = pix / iw;
    const int iy = (int)(t % ih);
    const int img = (int)(t / ih);
    float v = 0.f;                                          // an odd last row / column is not covered by any window
    if ((iy >> 1) < oh && (ix >> 1) < ow) v = 0.25f * __ldg(dy + (((int64_t)img * oh + (iy >> 1)) * ow + (ix >> 1)) * lddy + ch);
    float* dst = dx + pix * lddx + ch;
    *dst = accumulate ? *dst + v : v;
  }
}

extern "C" int b2_avgpool2x2_bwd(const float* dy, int lddy, float* dx, int lddx, int n, int ih, int iw, int c, int accumulate,
                                 void* stream) {
  B2_REQUIRE(dy && dx && n > 0 && ih >= 2 && iw >= 2 && c > 0 && lddy >= c && lddx >= c, "b2_avgpool2x2_bwd: bad args");
  const int64_t total = (int64_t)n * ih * iw * c;
  avgpool2x2_bwd_kernel<<<grid_for(total), UN_THREADS, 0, (cudaStream_t)stream>>>(dy, lddy, dx, lddx, n, ih / 2, iw / 2, ih, iw, c,
                                                                                 accumulate);
  B2_LAUNCH_CHECK("avgpool2x2_bwd_kernel");
  return B2_OK;
}

__global__ void __launch_bounds__(UN_THREADS)
scale_channels_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ scale, float* __restrict__ dst, int ldd,
                      int64_t rows, int c, int accumulate) {
  const int64_t total = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * UN_THREADS) {
    const int64_t r = i / c;
    const int ch = (int)(i - r * c);
    const float v = __ldg(g + r * ldg + ch) * __ldg(scale + ch);
    float* d = dst + r * ldd + ch;
    *d = accumulate ? *d + v : v;
  }
}

extern "C" int b2_scale_channels(const float* g, int ldg, const float* scale, float* dst, int ldd, int64_t rows, int c,
                                 int accumulate, void* stream) {
  B2_REQUIRE(g && scale && dst && rows > 0 && c > 0 && ldg >= c && ldd >= c, "b2_scale_channels: bad args");
  scale_channels_kernel<<<grid_for(rows * c), UN_THREADS, 0, (cudaStream_t)stream>>>(g, ldg, scale, dst, ldd, rows, c, accumulate);
  B2_LAUNCH_CHECK("scale_channels_kernel");
  return B2_OK;
}
