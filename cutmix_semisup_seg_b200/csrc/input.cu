// Device side of the data-format boundary: the last stage of the reference's DataLoader pipeline,
// SegCVTransformNormalizeToTensor (datapipe/seg_transforms_cv.py:587-672), moved behind the host-to-device copy so that a
// batch crosses PCIe as uint8 HWC pixels (3 B per pixel and image) instead of standardised fp32 planes (12 B):
//   image  uint8 (N,H,W,3|4)  ->  fp32 (N,3,H,W):  img_as_float (x * (1/255) in float64; scikit-image 0.16.2,
//          skimage/util/dtype.py `convert`: np.multiply(image, 1. / 255, dtype=float64)), then (v - mean) / std in float64
//          (:610; with a padding alpha channel (v - mean * alpha) / std, :606), then .astype(float32) (:614)
//   labels uint8 (N,H,W)      ->  int64 (N,1,H,W)   (:617)
//   mask   uint8 (N,H,W)      ->  fp32 (N,1,H,W) = float32(m * (1/255))   (:620)
// Float64 arithmetic on purpose: the results are bit-identical to the numpy pipeline.  HBM-bound, one pass.
#include "common.cuh"

constexpr int IN_THREADS = 256;

struct NormArgs { double mean[3]; double inv_scale; double stdv[3]; int has_norm; };

__global__ void __launch_bounds__(IN_THREADS)
normalize_to_tensor_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int64_t hw, int cin, NormArgs a) {
  const int n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
  if (p >= hw) return;
  const uint8_t* px = img + ((int64_t)n * hw + p) * cin;
  const double alpha = cin == 4 ? __dmul_rn((double)px[3], a.inv_scale) : 1.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = __dmul_rn((double)px[c], a.inv_scale);                 // img_as_float
    if (a.has_norm) {
      const double m = cin == 4 ? __dmul_rn(a.mean[c], alpha) : a.mean[c];
      v = __ddiv_rn(__dsub_rn(v, m), a.stdv[c]);                      // (image - mean [* alpha]) / std
    }
    out[((int64_t)n * 3 + c) * hw + p] = (float)v;                    // .astype(np.float32): round to nearest
  }
}

extern "C" int b2_normalize_to_tensor(const uint8_t* img, int n, int h, int w, int cin, const double* mean3,
                                      const double* std3, float* out, void* stream) {
  B2_REQUIRE(img && out && n > 0 && h > 0 && w > 0, "b2_normalize_to_tensor: bad args");
  B2_REQUIRE(cin == 3 || cin == 4, "b2_normalize_to_tensor: image should have 3 channels (or 4 with a padding alpha), not %d", cin);
  B2_REQUIRE((mean3 == nullptr) == (std3 == nullptr), "b2_normalize_to_tensor: mean and std must be given together");
  B2_REQUIRE(n <= 65535, "b2_normalize_to_tensor: n too large");
  NormArgs a;
  a.inv_scale = 1.0 / 255.0;
  a.has_norm = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean3 ? mean3[c] : 0.0; a.stdv[c] = std3 ? std3[c] : 1.0; }
  const int64_t hw = (int64_t)h * w;
  dim3 grid((unsigned)ceil_div64(hw, IN_THREADS), n);
  normalize_to_tensor_kernel<<<grid, IN_THREADS, 0, (cudaStream_t)stream>>>(img, out, hw, cin, a);
  B2_LAUNCH_CHECK("normalize_to_tensor_kernel");
  return B2_OK;
}

template <int MODE>
__global__ void __launch_bounds__(IN_THREADS)
u8_to_tensor_kernel(const uint8_t* __restrict__ src, void* __restrict__ dst, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x; i < count; i += (int64_t)gridDim.x * IN_THREADS) {
    if (MODE == 0) reinterpret_cast<int64_t*>(dst)[i] = (int64_t)src[i];
    else reinterpret_cast<float*>(dst)[i] = (float)__dmul_rn((double)src[i], 1.0 / 255.0);
  }
}

// mode 0: labels, uint8 -> int64 (:617);  mode 1: valid mask, uint8 -> float32(m * (1/255)) (:620)
extern "C" int b2_u8_to_tensor(const uint8_t* src, int64_t count, int mode, void* dst, void* stream) {
  B2_REQUIRE(src && dst && count > 0 && (mode == 0 || mode == 1), "b2_u8_to_tensor: bad args");
  int64_t blocks = ceil_div64(count, IN_THREADS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (mode == 0) u8_to_tensor_kernel<0><<<(unsigned)blocks, IN_THREADS, 0, (cudaStream_t)stream>>>(src, dst, count);
  else u8_to_tensor_kernel<1><<<(unsigned)blocks, IN_THREADS, 0, (cudaStream_t)stream>>>(src, dst, count);
  B2_LAUNCH_CHECK("u8_to_tensor_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Device crop + flip + normalise (rest of the input pipeline): SegCVTransformPad.pad_single / pad_pair
// (datapipe/seg_transforms_cv.py:30-62, 64-99), SegCVTransformRandomCrop (:102-167), SegCVTransformRandomFlip.flip_image
// (:467-474) and SegCVTransformNormalizeToTensor (:587-672) in ONE pass: the DataLoader workers keep the decoded uint8 image at
// its original size, the random parameters are drawn on the host exactly like the reference draws them
// (input_pipeline.DeviceCropFlipNormalize), and each output pixel is gathered from its source pixel:
//   output (oy, ox)  -- undo the transposition / vertical / horizontal flip -->  crop (cy, cx)
//   -> padded image (pos_y + cy, pos_x + cx) -> source (.. - pad_top, .. - pad_left), or padding when outside:
//      image 0 with alpha 0 (so the standardised value is exactly 0), labels 255, mask 0.
// A padded sample carries the reference's alpha channel, i.e. (v - mean * alpha) / std; inside the image alpha = 255/255 = 1.0
// exactly, so the arithmetic is the same as for an unpadded one.  Float64, one rounding to float32: bit-identical to numpy.
struct CropArgs { double mean[3]; double inv_scale; double stdv[3]; int has_norm; int out_h, out_w; };

__global__ void __launch_bounds__(IN_THREADS)
crop_flip_normalize_kernel(const b2_crop_entry* __restrict__ table, CropArgs a, float* __restrict__ image,
                           int64_t* __restrict__ labels, float* __restrict__ mask) {
  const int n = blockIdx.y;
  const int64_t hw = (int64_t)a.out_h * a.out_w;
  const int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
  if (p >= hw) return;
  const b2_crop_entry e = table[n];
  int oy = (int)(p / a.out_w), ox = (int)(p % a.out_w);
  int cy = oy, cx = ox;
  if (e.flip_d) { cy = ox; cx = oy; }                 // np.swapaxes(img, 0, 1) was applied last
  if (e.flip_y) cy = e.crop_h - 1 - cy;               // img[::-1, ...]
  if (e.flip_x) cx = e.crop_w - 1 - cx;               // img[:, ::-1]
  const int sy = e.pos_y + cy - e.pad_top, sx = e.pos_x + cx - e.pad_left;
  const bool inside = sy >= 0 && sy < e.h0 && sx >= 0 && sx < e.w0;
  const int64_t sp = (int64_t)sy * e.w0 + sx;
  const double alpha = (e.padded && !inside) ? 0.0 : 1.0;     // img_as_float(255) == 255 * (1/255) == 1.0 exactly
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = inside ? __dmul_rn((double)e.image[sp * 3 + c], a.inv_scale) : 0.0;
    if (a.has_norm) {
      const double m = e.padded ? __dmul_rn(a.mean[c], alpha) : a.mean[c];
      v = __ddiv_rn(__dsub_rn(v, m), a.stdv[c]);
    }
    image[((int64_t)n * 3 + c) * hw + p] = (float)v;
  }
  if (labels) labels[(int64_t)n * hw + p] = (e.labels && inside) ? (int64_t)e.labels[sp] : 255;
  if (mask) mask[(int64_t)n * hw + p] = (e.mask && inside) ? (float)__dmul_rn((double)e.mask[sp], a.inv_scale) : 0.0f;
}

extern "C" int b2_crop_flip_normalize(const b2_crop_entry* table, int n, int out_h, int out_w, const double* mean3,
                                      const double* std3, float* image, int64_t* labels, float* mask, void* stream) {
  B2_REQUIRE(table && image && n > 0 && n <= 65535 && out_h > 0 && out_w > 0, "b2_crop_flip_normalize: bad args");
  B2_REQUIRE((mean3 == nullptr) == (std3 == nullptr), "b2_crop_flip_normalize: mean and std must be given together");
  CropArgs a;
  a.inv_scale = 1.0 / 255.0;
  a.has_norm = mean3 != nullptr;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean3 ? mean3[c] : 0.0; a.stdv[c] = std3 ? std3[c] : 1.0; }
  a.out_h = out_h; a.out_w = out_w;
  dim3 grid((unsigned)ceil_div64((int64_t)out_h * out_w, IN_THREADS), n);
  crop_flip_normalize_kernel<<<grid, IN_THREADS, 0, (cudaStream_t)stream>>>(table, a, image, labels, mask);
  B2_LAUNCH_CHECK("crop_flip_normalize_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Crop + flip to uint8 RGBA (the strong-colour branch keeps the crop as pixels for the colour jitter below; alpha = 255 inside
// the source image, 0 in the padding: b2_normalize_to_tensor(cin = 4) then reproduces the reference's alpha standardisation).
__global__ void __launch_bounds__(IN_THREADS)
crop_flip_u8_kernel(const b2_crop_entry* __restrict__ table, int out_h, int out_w, uint8_t* __restrict__ image,
                    int64_t* __restrict__ labels, float* __restrict__ mask) {
  const int n = blockIdx.y;
  const int64_t hw = (int64_t)out_h * out_w;
  const int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
  if (p >= hw) return;
  const b2_crop_entry e = table[n];
  const int oy = (int)(p / out_w), ox = (int)(p % out_w);
  int cy = oy, cx = ox;
  if (e.flip_d) { cy = ox; cx = oy; }
  if (e.flip_y) cy = e.crop_h - 1 - cy;
  if (e.flip_x) cx = e.crop_w - 1 - cx;
  const int sy = e.pos_y + cy - e.pad_top, sx = e.pos_x + cx - e.pad_left;
  const bool inside = sy >= 0 && sy < e.h0 && sx >= 0 && sx < e.w0;
  const int64_t sp = (int64_t)sy * e.w0 + sx;
  uchar4 px = make_uchar4(0, 0, 0, 0);
  if (inside) px = make_uchar4(e.image[sp * 3], e.image[sp * 3 + 1], e.image[sp * 3 + 2], 255);
  reinterpret_cast<uchar4*>(image)[(int64_t)n * hw + p] = px;
  if (labels) labels[(int64_t)n * hw + p] = (e.labels && inside) ? (int64_t)e.labels[sp] : 255;
  if (mask) mask[(int64_t)n * hw + p] = (e.mask && inside) ? (float)__dmul_rn((double)e.mask[sp], 1.0 / 255.0) : 0.0f;
}

extern "C" int b2_crop_flip_u8(const b2_crop_entry* table, int n, int out_h, int out_w, uint8_t* image_rgba, int64_t* labels,
                               float* mask, void* stream) {
  B2_REQUIRE(table && image_rgba && n > 0 && n <= 65535 && out_h > 0 && out_w > 0, "b2_crop_flip_u8: bad args");
  B2_REQUIRE((reinterpret_cast<uintptr_t>(image_rgba) & 3) == 0, "b2_crop_flip_u8: output must be 4-byte aligned");
  dim3 grid((unsigned)ceil_div64((int64_t)out_h * out_w, IN_THREADS), n);
  crop_flip_u8_kernel<<<grid, IN_THREADS, 0, (cudaStream_t)stream>>>(table, out_h, out_w, image_rgba, labels, mask);
  B2_LAUNCH_CHECK("crop_flip_u8_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Scale / rotation crops on uint8 pixels: SegCVTransformRandomCropScaleHung (datapipe/seg_transforms_cv.py:
// 169-303: pad, crop a scaled window, cv2.resize to the crop size) and SegCVTransformRandomCropRotateScale (:305-449: cv2.warpAffine
// of the whole image into the crop) followed by SegCVTransformRandomFlip.flip_image (:467-474), as ONE gather per output pixel.
// OpenCV's uint8 paths are integer algorithms (imgproc/resize.cpp, imgwarp.cpp; pinned 3.4.2, same results from the installed
// 4.13), restated here; the host (input_pipeline.resize_tables / warp_tables) prepares the per-sample integer tables:
//   mode 0  resize of the window [pos, pos + src) of the virtually padded image.  tables (int32, W = out_w, H = out_h):
//           [0,W) nearest source column, [W,2W) left column of the linear filter, [2W,3W) its coefficients a0 | a1 << 16 (11 bit),
//           [3W,3W+H) nearest row, [3W+H,3W+2H) upper row (clipped at use), [3W+2H,3W+3H) b0 | b1 << 16.
//           linear: rows are filtered horizontally into ints (S[x0]*a0 + S[x1]*a1), then
//           (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;  interp 2 = exact 2x decimation = (sum of 2x2 + 2) >> 2.
//   mode 1  warpAffine.  tables: [0,W) adelta, [W,2W) bdelta, [3W,3W+H) X0, [3W+H,3W+2H) Y0 (10-bit fixed point of the inverted
//           matrix).  nearest: X = (X0 + 512 + adelta) >> 10; linear: X = (X0 + 16 + adelta) >> 5, pixel X >> 5, 5-bit fractions,
//           weights (32-fy)(32-fx)*32 ... (sum 2^15), (acc + 2^14) >> 15.  Borders: image BORDER_REFLECT_101, labels
//           BORDER_CONSTANT 255, mask BORDER_CONSTANT 0.
// Output: RGBA uint8 (alpha = 255 inside the image, 0 in the padding, interpolated like a colour channel: the reference resizes
// the 4-channel padded image), labels int64, mask float32(m * (1/255)).
__device__ __forceinline__ int geom_reflect101(int p, int len) {
  if (len == 1) return 0;
  while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}
__device__ __forceinline__ int geom_sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

struct GeomSrc {
  const b2_geom_entry* e;
  // pixel (y, x) of the window of the virtually padded image (mode 0): channel c of RGBA / labels / mask
  __device__ __forceinline__ bool inside(int y, int x, int64_t* sp) const {
    const int sy = e->pos_y + y - e->pad_top, sx = e->pos_x + x - e->pad_left;
    *sp = (int64_t)sy * e->w0 + sx;
    return sy >= 0 && sy < e->h0 && sx >= 0 && sx < e->w0;
  }
  __device__ __forceinline__ void rgba(int y, int x, int (&v)[4]) const {
    int64_t sp;
    if (inside(y, x, &sp)) { v[0] = e->image[sp * 3]; v[1] = e->image[sp * 3 + 1]; v[2] = e->image[sp * 3 + 2]; v[3] = 255; }
    else { v[0] = v[1] = v[2] = v[3] = 0; }
  }
  __device__ __forceinline__ int plane(const uint8_t* p, int y, int x, int outside) const {
    int64_t sp;
    return inside(y, x, &sp) ? (int)p[sp] : outside;
  }
};

__device__ __forceinline__ int geom_vresize(int r0, int r1, int b0, int b1) {
  return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

__global__ void __launch_bounds__(IN_THREADS)
geom_u8_kernel(const b2_geom_entry* __restrict__ table, const int32_t* __restrict__ tables, int out_h, int out_w,
               uint8_t* __restrict__ image, int64_t* __restrict__ labels, float* __restrict__ mask) {
  const int n = blockIdx.y;
  const int64_t hw = (int64_t)out_h * out_w;
  const int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
  if (p >= hw) return;
  const b2_geom_entry e = table[n];
  const int32_t* t = tables + e.tab_off;
  const int W = out_w, H = out_h;
  const int oy = (int)(p / out_w), ox = (int)(p % out_w);
  int ry = oy, rx = ox;                                   // pixel of the un-flipped crop
  if (e.flip_d) { ry = ox; rx = oy; }
  if (e.flip_y) ry = H - 1 - ry;
  if (e.flip_x) rx = W - 1 - rx;
  int px[4] = {0, 0, 0, 255};
  int lab = 255, msk = 0;
  if (e.mode == 0) {
    GeomSrc s{&e};
    const int xn = t[rx], yn = t[3 * W + ry];
    const int xl = t[W + rx], xa = t[2 * W + rx], yl = t[3 * W + H + ry], yb = t[3 * W + 2 * H + ry];
    const int x1 = min(xl + 1, e.src_w - 1);
    const int y0 = min(max(yl, 0), e.src_h - 1), y1 = min(max(yl + 1, 0), e.src_h - 1);
    const int a0 = xa & 0xffff, a1 = (xa >> 16) & 0xffff, b0 = yb & 0xffff, b1 = (yb >> 16) & 0xffff;
    if (e.image_interp == 0) {
      s.rgba(yn, xn, px);
    } else if (e.image_interp == 1) {
      int v00[4], v01[4], v10[4], v11[4];
      s.rgba(y0, xl, v00); s.rgba(y0, x1, v01); s.rgba(y1, xl, v10); s.rgba(y1, x1, v11);
#pragma unroll
      for (int c = 0; c < 4; ++c) px[c] = geom_vresize(v00[c] * a0 + v01[c] * a1, v10[c] * a0 + v11[c] * a1, b0, b1);
    } else {
      int v00[4], v01[4], v10[4], v11[4];
      s.rgba(2 * ry, 2 * rx, v00); s.rgba(2 * ry, 2 * rx + 1, v01); s.rgba(2 * ry + 1, 2 * rx, v10); s.rgba(2 * ry + 1, 2 * rx + 1, v11);
#pragma unroll
      for (int c = 0; c < 4; ++c) px[c] = (v00[c] + v01[c] + v10[c] + v11[c] + 2) >> 2;
    }
    if (labels && e.labels) lab = s.plane(e.labels, yn, xn, 255);
    if (mask && e.mask) {
      if (e.mask_interp == 0) msk = s.plane(e.mask, yn, xn, 0);
      else if (e.mask_interp == 1)
        msk = geom_vresize(s.plane(e.mask, y0, xl, 0) * a0 + s.plane(e.mask, y0, x1, 0) * a1,
                           s.plane(e.mask, y1, xl, 0) * a0 + s.plane(e.mask, y1, x1, 0) * a1, b0, b1);
      else
        msk = (s.plane(e.mask, 2 * ry, 2 * rx, 0) + s.plane(e.mask, 2 * ry, 2 * rx + 1, 0) + s.plane(e.mask, 2 * ry + 1, 2 * rx, 0) +
               s.plane(e.mask, 2 * ry + 1, 2 * rx + 1, 0) + 2) >> 2;
    }
  } else {
    const int ad = t[rx], bd = t[W + rx], X0 = t[3 * W + ry], Y0 = t[3 * W + H + ry];
    const int h0 = e.h0, w0 = e.w0;
    // nearest coordinates (labels always; image / mask when their interpolation is nearest)
    const int nx = geom_sat_short((X0 + 512 + ad) >> 10), ny = geom_sat_short((Y0 + 512 + bd) >> 10);
    const bool n_in = (unsigned)nx < (unsigned)w0 && (unsigned)ny < (unsigned)h0;
    // linear coordinates
    const int LX = (X0 + 16 + ad) >> 5, LY = (Y0 + 16 + bd) >> 5;
    const int sx = geom_sat_short(LX >> 5), sy = geom_sat_short(LY >> 5);
    const int fx = LX & 31, fy = LY & 31;
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    if (e.image_interp == 0) {
      const int yy = geom_reflect101(ny, h0), xx = geom_reflect101(nx, w0);
      const int64_t sp = (int64_t)yy * w0 + xx;
      px[0] = e.image[sp * 3]; px[1] = e.image[sp * 3 + 1]; px[2] = e.image[sp * 3 + 2];
    } else {
      const int xa = geom_reflect101(sx, w0), xb = geom_reflect101(sx + 1, w0);
      const int ya = geom_reflect101(sy, h0), yb = geom_reflect101(sy + 1, h0);
      const uint8_t* p00 = e.image + ((int64_t)ya * w0 + xa) * 3; const uint8_t* p01 = e.image + ((int64_t)ya * w0 + xb) * 3;
      const uint8_t* p10 = e.image + ((int64_t)yb * w0 + xa) * 3; const uint8_t* p11 = e.image + ((int64_t)yb * w0 + xb) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) px[c] = (p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11 + (1 << 14)) >> 15;
    }
    if (labels && e.labels) lab = n_in ? (int)e.labels[(int64_t)ny * w0 + nx] : 255;
    if (mask && e.mask) {
      if (e.mask_interp == 0) msk = n_in ? (int)e.mask[(int64_t)ny * w0 + nx] : 0;
      else if (sx >= w0 || sx + 1 < 0 || sy >= h0 || sy + 1 < 0) msk = 0;
      else {
        auto tap = [&](int yy, int xx) -> int {
          return ((unsigned)xx < (unsigned)w0 && (unsigned)yy < (unsigned)h0) ? (int)e.mask[(int64_t)yy * w0 + xx] : 0;
        };
        msk = (tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11 + (1 << 14)) >> 15;
      }
    }
  }
  reinterpret_cast<uchar4*>(image)[(int64_t)n * hw + p] = make_uchar4((unsigned char)px[0], (unsigned char)px[1], (unsigned char)px[2],
                                                                      (unsigned char)px[3]);
  if (labels) labels[(int64_t)n * hw + p] = e.labels ? (int64_t)lab : 255;
  if (mask) mask[(int64_t)n * hw + p] = e.mask ? (float)__dmul_rn((double)msk, 1.0 / 255.0) : 0.0f;
}

extern "C" int b2_geom_u8(const b2_geom_entry* table, const int32_t* tables, int n, int out_h, int out_w, uint8_t* image_rgba,
                          int64_t* labels, float* mask, void* stream) {
  B2_REQUIRE(table && tables && image_rgba && n > 0 && n <= 65535 && out_h > 0 && out_w > 0, "b2_geom_u8: bad args");
  B2_REQUIRE((reinterpret_cast<uintptr_t>(image_rgba) & 3) == 0, "b2_geom_u8: output must be 4-byte aligned");
  dim3 grid((unsigned)ceil_div64((int64_t)out_h * out_w, IN_THREADS), n);
  geom_u8_kernel<<<grid, IN_THREADS, 0, (cudaStream_t)stream>>>(table, tables, out_h, out_w, image_rgba, labels, mask);
  B2_LAUNCH_CHECK("geom_u8_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// Colour jitter on uint8 pixels: torchvision ColorJitter + RandomGrayscale on PIL images (the reference's strong colour
// augmentation, train_seg_semisup_mask_mt.py:169-179 -> datapipe/seg_transforms_cv.py:541-586), i.e. Pillow's C arithmetic
// (libImaging/Blend.c, Convert.c; ImageEnhance.py), restated with its float / double mixing and explicit roundings so that the
// bytes are identical (tests/colour_recipe.py states the same algorithm in numpy; pinned against the installed Pillow /
// torchvision, exhaustively for the HSV conversions).  Up to four operations per image in a per-image random order:
//   brightness  blend(0, x, a);  contrast  blend(mean L, x, a) (mean over the image: a reduction pass);  saturation
//   blend(L(x), x, a);  hue  RGB -> HSV, H += shift (mod 256), HSV -> RGB;  then optionally grey = L(x).
__device__ __forceinline__ int colour_lum(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

__device__ __forceinline__ int colour_blend(int d, int x, float a) {
  const float t = __fadd_rn((float)d, __fmul_rn(a, (float)(x - d)));
  if (a >= 0.0f && a <= 1.0f) return (int)t;                 // (UINT8) cast of a value in [0, 255]: truncation
  return t <= 0.0f ? 0 : (t >= 255.0f ? 255 : (int)t);
}

__device__ __forceinline__ void colour_rgb2hsv(int r, int g, int b, int* uh, int* us, int* uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  *uv = maxc;
  if (minc == maxc) { *uh = 0; *us = 0; return; }
  const float cr = (float)(maxc - minc);
  const float s = __fdiv_rn(cr, (float)maxc);
  const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = (float)__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc);
  else h = (float)__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc);
  const double hh = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);
  h = (float)(hh - floor(hh));                                // fmod(x, 1.0) for x >= 0: exact
  int ih = (int)__dmul_rn((double)h, 255.0), is = (int)__dmul_rn((double)s, 255.0);
  *uh = min(max(ih, 0), 255); *us = min(max(is, 0), 255);
}

__device__ __forceinline__ void colour_hsv2rgb(int h, int s, int v, int* r, int* g, int* b) {
  if (s == 0) { *r = *g = *b = v; return; }
  const double h6 = __ddiv_rn(__dmul_rn((double)(float)h, 6.0), 255.0);
  const int i = (int)floor(h6);
  const float f = (float)__dsub_rn(h6, (double)(float)i);
  const float fs = (float)__ddiv_rn((double)(float)s, 255.0);
  const double vf = (double)(float)v;
  const int p = (int)round(__dmul_rn(vf, __dsub_rn(1.0, (double)fs)));
  const int q = (int)round(__dmul_rn(vf, __dsub_rn(1.0, (double)__fmul_rn(fs, f))));
  const int t = (int)round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn((double)fs, __dsub_rn(1.0, (double)f)))));
  const int up = min(max(p, 0), 255), uq = min(max(q, 0), 255), ut = min(max(t, 0), 255);
  switch (i % 6) {
    case 0: *r = v; *g = ut; *b = up; break;
    case 1: *r = uq; *g = v; *b = up; break;
    case 2: *r = up; *g = v; *b = ut; break;
    case 3: *r = up; *g = uq; *b = v; break;
    case 4: *r = ut; *g = up; *b = v; break;
    default: *r = v; *g = up; *b = uq; break;
  }
}

// sums[n] += sum of L over the pixels of image n, for the images whose k-th operation is a contrast change
__global__ void __launch_bounds__(IN_THREADS)
colour_lum_sum_kernel(const uint8_t* __restrict__ img, int cstride, int64_t hw, const b2_colour_entry* __restrict__ table, int k,
                      unsigned long long* __restrict__ sums) {
  const int n = blockIdx.y;
  if (k >= table[n].n_ops || table[n].op[k] != 1) return;
  unsigned long long acc = 0;
  for (int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x; p < hw; p += (int64_t)gridDim.x * IN_THREADS) {
    const uint8_t* px = img + ((int64_t)n * hw + p) * cstride;
    acc += (unsigned long long)colour_lum(px[0], px[1], px[2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[n], acc);          // integer sum: order-independent
}

__global__ void __launch_bounds__(IN_THREADS)
colour_apply_kernel(uint8_t* __restrict__ img, int cstride, int64_t hw, const b2_colour_entry* __restrict__ table, int k,
                    const unsigned long long* __restrict__ sums, int last) {
  const int n = blockIdx.y;
  const b2_colour_entry e = table[n];
  const int64_t p = (int64_t)blockIdx.x * IN_THREADS + threadIdx.x;
  if (p >= hw) return;
  const int op = k < e.n_ops ? e.op[k] : -1;
  const bool grey = last && e.grey;
  if (op < 0 && !grey) return;
  uint8_t* px = img + ((int64_t)n * hw + p) * cstride;
  int r = px[0], g = px[1], b = px[2];
  if (op == 0) {
    const float a = e.factor[k];
    r = colour_blend(0, r, a); g = colour_blend(0, g, a); b = colour_blend(0, b, a);
  } else if (op == 1) {
    const float a = e.factor[k];
    const int m = (int)__dadd_rn(__ddiv_rn((double)sums[n], (double)hw), 0.5);
    r = colour_blend(m, r, a); g = colour_blend(m, g, a); b = colour_blend(m, b, a);
  } else if (op == 2) {
    const float a = e.factor[k];
    const int l = colour_lum(r, g, b);
    r = colour_blend(l, r, a); g = colour_blend(l, g, a); b = colour_blend(l, b, a);
  } else if (op == 3) {
    int h, s, v;
    colour_rgb2hsv(r, g, b, &h, &s, &v);
    h = (h + e.hue_shift[k]) & 255;
    colour_hsv2rgb(h, s, v, &r, &g, &b);
  }
  if (grey) { const int l = colour_lum(r, g, b); r = g = b = l; }
  px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b;
}

extern "C" int b2_colour_jitter(uint8_t* img, int n, int h, int w, int cstride, const b2_colour_entry* table_dev,
                                const b2_colour_entry* table_host, unsigned long long* workspace, void* stream) {
  B2_REQUIRE(img && table_dev && table_host && workspace && n > 0 && n <= 65535 && h > 0 && w > 0, "b2_colour_jitter: bad args");
  B2_REQUIRE(cstride == 3 || cstride == 4, "b2_colour_jitter: pixel stride must be 3 (RGB) or 4 (RGBA)");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t hw = (int64_t)h * w;
  int max_ops = 0; bool any_grey = false;
  for (int i = 0; i < n; ++i) {
    B2_REQUIRE(table_host[i].n_ops >= 0 && table_host[i].n_ops <= 4, "b2_colour_jitter: n_ops out of range");
    for (int k = 0; k < table_host[i].n_ops; ++k)
      B2_REQUIRE(table_host[i].op[k] >= 0 && table_host[i].op[k] <= 3, "b2_colour_jitter: unknown operation %d", table_host[i].op[k]);
    if (table_host[i].n_ops > max_ops) max_ops = table_host[i].n_ops;
    any_grey = any_grey || table_host[i].grey;
  }
  if (max_ops == 0 && !any_grey) return B2_OK;
  const int passes = max_ops > 0 ? max_ops : 1;
  dim3 grid((unsigned)ceil_div64(hw, IN_THREADS), n);
  for (int k = 0; k < passes; ++k) {
    bool contrast = false;
    for (int i = 0; i < n; ++i) contrast = contrast || (k < table_host[i].n_ops && table_host[i].op[k] == 1);
    if (contrast) {
      B2_CUDA(cudaMemsetAsync(workspace, 0, sizeof(unsigned long long) * n, s));
      int64_t bx = ceil_div64(hw, IN_THREADS * 8); if (bx > 1024) bx = 1024; if (bx < 1) bx = 1;
      colour_lum_sum_kernel<<<dim3((unsigned)bx, n), IN_THREADS, 0, s>>>(img, cstride, hw, table_dev, k, workspace);
      B2_LAUNCH_CHECK("colour_lum_sum_kernel");
    }
    colour_apply_kernel<<<grid, IN_THREADS, 0, s>>>(img, cstride, hw, table_dev, k, workspace, k == passes - 1);
    B2_LAUNCH_CHECK("colour_apply_kernel");
  }
  return B2_OK;
}
