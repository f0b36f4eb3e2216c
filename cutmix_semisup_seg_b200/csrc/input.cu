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
