// On-device evaluation: per-pixel argmax of the logits fused with the confusion-matrix update --
// replaces `torch.argmax(...).cpu().numpy()` + per-image numpy confusion matrices in the validation loop
// (train_seg_semisup_mask_mt.py:484-517, evaluation.py:6-62).  One pass over the NCHW logits (HBM-bound: C*4 + 8 bytes
// per pixel), one thread per pixel, channel loop strided by H*W (coalesced 128 B lines), block-private histogram in
// shared memory, 64-bit global accumulation.  Integer arithmetic: results are exact.
#include "common.cuh"

namespace {

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_MAX_C = 64;

// cm[t * C + p] += 1 for every pixel whose label t != ignore (labels outside [0, C) are skipped like ignored ones).
// argmax ties resolve to the lowest class index (torch.argmax / numpy.argmax on the first maximum); NaN logits never win
// against a number, an all-NaN pixel predicts class 0.
__global__ void __launch_bounds__(EVAL_THREADS)
argmax_confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C, int64_t hw,
                        int64_t ignore, unsigned long long* __restrict__ cm, int64_t* __restrict__ pred_out) {
  extern __shared__ unsigned int hist[];            // C * C
  for (int i = threadIdx.x; i < C * C; i += EVAL_THREADS) hist[i] = 0u;
  __syncthreads();
  const int img = blockIdx.y;
  const int64_t stride = (int64_t)gridDim.x * EVAL_THREADS;
  for (int64_t p = (int64_t)blockIdx.x * EVAL_THREADS + threadIdx.x; p < hw; p += stride) {
    const float* l = logits + (int64_t)img * C * hw + p;
    float best = __ldg(l);
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(l + (int64_t)c * hw);
      if (v > best || (best != best && v == v)) { best = v; arg = c; }
    }
    const int64_t t = labels ? labels[(int64_t)img * hw + p] : -1;
    if (pred_out) pred_out[(int64_t)img * hw + p] = arg;
    if (labels && t != ignore && t >= 0 && t < C) atomicAdd(&hist[(int)t * C + arg], 1u);
  }
  __syncthreads();
  if (cm)
    for (int i = threadIdx.x; i < C * C; i += EVAL_THREADS)
      if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

}  // namespace

extern "C" int b2_argmax_confusion(const float* logits, const int64_t* labels, int n, int c, int64_t hw, int64_t ignore,
                                   int64_t* cm, int64_t* pred_out, void* stream) {
  B2_REQUIRE(logits && n > 0 && c > 0 && hw > 0, "b2_argmax_confusion: bad args");
  B2_REQUIRE(c <= EVAL_MAX_C, "b2_argmax_confusion: C=%d > %d unsupported", c, EVAL_MAX_C);
  B2_REQUIRE(n <= 65535, "b2_argmax_confusion: n too large");
  B2_REQUIRE((cm && labels) || pred_out, "b2_argmax_confusion: nothing to compute");
  B2_REQUIRE(hw < (1ll << 31) * EVAL_THREADS, "b2_argmax_confusion: image too large");
  int64_t bx = ceil_div64(hw, EVAL_THREADS);
  const int64_t cap = (int64_t)b2_sm_count_cached() * 8;       // grid-stride: a few resident blocks per SM
  if (cap > 0 && bx * n > cap) bx = ceil_div64(cap, n);
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, n);
  argmax_confusion_kernel<<<grid, EVAL_THREADS, (size_t)c * c * sizeof(unsigned int), (cudaStream_t)stream>>>(
      logits, labels, c, hw, ignore, reinterpret_cast<unsigned long long*>(cm), pred_out);
  B2_LAUNCH_CHECK("argmax_confusion_kernel");
  return B2_OK;
}
