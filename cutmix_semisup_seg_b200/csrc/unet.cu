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
