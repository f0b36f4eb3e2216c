// Table-driven multi-tensor forms of the per-layer derived-weight kernels (one launch per network and pass instead of one
// per layer; the CUDA-graph of the cfg3 iteration held 208 bn_fold and 113 transpose_w nodes of ~1 us of work each):
//   b2_bn_fold_multi      scale / shift of every frozen (eval-mode) BatchNorm of a network
//   b2_transpose_w_multi  (Cout, taps, Cin) -> (Cin, taps, pad4(Cout)) dgrad operands of every convolution of a backward pass,
//                         with the folded BatchNorm scale applied per output channel
// Same arithmetic as the single-tensor kernels (bn_fold_kernel in netops.cu, transpose_w_kernel in elementwise.cu): results are
// bit-identical.  Tables live in device memory and are built once per (network, tape shape) by the host (engine.py).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128) bn_fold_multi_kernel(const b2_bn_fold_entry* __restrict__ table) {
  const b2_bn_fold_entry e = table[blockIdx.y];
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= e.c) return;
  const float s = e.gamma[ch] / sqrtf(e.var[ch] + e.eps);
  e.scale[ch] = s;
  e.shift[ch] = e.beta[ch] - e.mean[ch] * s;
}

__global__ void __launch_bounds__(256) transpose_w_multi_kernel(const b2_transpose_entry* __restrict__ table, int n_entries) {
  __shared__ float tile[32][33];
  // binary search: last entry whose first block is <= blockIdx.x
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].block_begin <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const b2_transpose_entry e = table[lo];
  int64_t r = (int64_t)blockIdx.x - e.block_begin;
  const int gx = (e.b + 31) / 32, gy = (e.ldd + 31) / 32;
  const int bx = (int)(r % gx); r /= gx;
  const int by = (int)(r % gy);
  const int t = (int)(r / gy);
  const int a0 = by * 32, b0 = bx * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int a = a0 + i, b = b0 + tx;
    tile[i][tx] = (a < e.a && b < e.b) ? e.src[((int64_t)a * e.t + t) * e.b + b] * (e.scale ? e.scale[a] : 1.0f) : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int b = b0 + i, a = a0 + tx;
    if (a < e.ldd && b < e.b) e.dst[((int64_t)b * e.t + t) * e.ldd + a] = a < e.a ? tile[tx][i] : 0.0f;
  }
}

}  // namespace

extern "C" int b2_bn_fold_multi(const b2_bn_fold_entry* table, int n_entries, int max_c, void* stream) {
  if (n_entries == 0) return B2_OK;
  B2_REQUIRE(table && n_entries > 0 && n_entries <= 65535 && max_c > 0, "b2_bn_fold_multi: bad args");
  bn_fold_multi_kernel<<<dim3((max_c + 127) / 128, n_entries), 128, 0, (cudaStream_t)stream>>>(table);
  B2_LAUNCH_CHECK("bn_fold_multi_kernel");
  return B2_OK;
}

extern "C" int b2_transpose_w_multi(const b2_transpose_entry* table, int n_entries, int64_t total_blocks, void* stream) {
  if (n_entries == 0) return B2_OK;
  B2_REQUIRE(table && n_entries > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "b2_transpose_w_multi: bad args");
  transpose_w_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(table, n_entries);
  B2_LAUNCH_CHECK("transpose_w_multi_kernel");
  return B2_OK;
}
