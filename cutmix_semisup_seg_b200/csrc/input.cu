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
