// HBM-bound NHWC network operators around the tensor-core convolutions:
// layout changes, stem im2col, 3x3/s2 max-pool, bilinear resize, global-average-pool / broadcast,
// batch-norm (train-mode statistics, apply, backward; frozen-BN folding and parameter gradients),
// per-channel reductions.  All reductions are two-level with a fixed summation order (deterministic).
#include <stdlib.h>
#include "common.cuh"
#include <math_constants.h>

// b2_debug_set(20, 1): the flat-index forms of the stem im2col and of the NHWC bilinear resize instead of the row-based kernels
// (bit-identical results; A/B timing and the bit-exactness tests)
int g_netops_flat_kernels = 0;

// ------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int c, int64_t hw, int ldd) {
  const int n = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (int64_t)gridDim.x * blockDim.x) {
    float* d = dst + ((int64_t)n * hw + p) * ldd;
    for (int ch = 0; ch < ldd; ++ch) d[ch] = ch < c ? __ldg(src + ((int64_t)n * c + ch) * hw + p) : 0.0f;
  }
}
extern "C" int b2_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, int ldd, void* stream) {
  B2_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && ldd >= c, "b2_nchw_to_nhwc: bad args");
  const int64_t hw = (int64_t)h * w;
  int bx = (int)((hw + 255) / 256); if (bx > 4096) bx = 4096;
  nchw_to_nhwc_kernel<<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(src, dst, c, hw, ldd);
  B2_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return B2_OK;
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int c, int64_t hw, int lds) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32; const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i; const int ch = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && ch < c) ? src[((int64_t)n * hw + p) * lds + ch] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i; const int64_t p = p0 + threadIdx.x;
    if (p < hw && ch < c) dst[((int64_t)n * c + ch) * hw + p] = tile[threadIdx.x][i];
  }
}
extern "C" int b2_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, int lds, void* stream) {
  B2_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && lds >= c, "b2_nhwc_to_nchw: bad args");
  const int64_t hw = (int64_t)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (c + 31) / 32, n);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, c, hw, lds);
  B2_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ im2col
// Thread per (output pixel, group of 4 consecutive K columns): one 16 B store per thread, consecutive threads write
// consecutive 16 B pieces of the column matrix (write-bound: the matrix is ~13x the image).  K index = (r*kw + s)*c + ch;
// columns >= kh*kw*c are the zero padding.  Requires kpad % 4 == 0 and a 16 B aligned matrix (host-checked).
template <typename I>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int n, int h, int w, int c,
                                                     int ldx, int kh, int kw, int stride, int pad, int dil, int oh, int ow, int kpad) {
  const int groups = kpad >> 2;
  const int kreal = kh * kw * c;
  const I total = (I)((int64_t)n * oh * ow * groups);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const I row = i / groups;
    const int grp = (int)(i - row * groups);
    const int x_o = (int)(row % ow); const I t = row / ow; const int y_o = (int)(t % oh); const int img = (int)(t / oh);
    int k = grp * 4;
    int tap = k / c, ch = k - tap * c;
    int r = tap / kw, s_ = tap - r * kw;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float val = 0.f;
      if (k < kreal) {
        const int iy = y_o * stride - pad + r * dil, ix = x_o * stride - pad + s_ * dil;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) val = __ldg(x + (((int64_t)img * h + iy) * w + ix) * ldx + ch);
      }
      v[e] = val;
      ++k;
      if (++ch == c) { ch = 0; if (++s_ == kw) { s_ = 0; ++r; } }
    }
    *reinterpret_cast<float4*>(col + (int64_t)row * kpad + grp * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// The stem of both hot-path networks (7x7, 3 channels, K padded 147 -> 160): the same mapping with every divisor a
// compile-time constant and the (image, output row) taken from blockIdx.y -- the generic kernel spends most of its time
// in seven run-time integer divisions per 16 B store (1.7 TB/s; profiles/r01_v11_launch_list_summary.txt).
template <int C, int KH, int KW, int KPAD>
__global__ void __launch_bounds__(256) im2col_const_kernel(const float* __restrict__ x, float* __restrict__ col, int h, int w,
                                                           int ldx, int stride, int pad, int dil, int oh, int ow) {
  constexpr int GROUPS = KPAD / 4, KREAL = KH * KW * C;
  const int y_o = blockIdx.y % oh, img = blockIdx.y / oh;
  const float* __restrict__ ximg = x + (int64_t)img * h * w * ldx;
  float* __restrict__ crow = col + (int64_t)blockIdx.y * ow * KPAD;
  const int per_row = ow * GROUPS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const int x_o = i / GROUPS, grp = i - x_o * GROUPS;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = grp * 4 + e;
      const int tap = k / C, ch = k - tap * C;
      const int r = tap / KW, s_ = tap - r * KW;
      float val = 0.f;
      if (k < KREAL) {
        const int iy = y_o * stride - pad + r * dil, ix = x_o * stride - pad + s_ * dil;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) val = __ldg(ximg + ((int64_t)iy * w + ix) * ldx + ch);
      }
      v[e] = val;
    }
    *reinterpret_cast<float4*>(crow + (int64_t)i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// The same stem matrix from a shared-memory copy of the input window: a block owns PX consecutive output pixels of one output row,
// stages the 7 input rows x ((PX - 1) * stride + 7) pixels it needs (16 B per pixel: 3 channels + the pad channel, ldx == 4; zeros
// outside the image) with one 16 B load per pixel, and builds its PX x 40 sixteen-byte column groups from shared memory.  The
// kernel above issues four scalar 4 B global loads with index arithmetic per 16 B store and ran at 2.8 TB/s of writes
// (profiles/r02_v22_launch_list_summary.txt); values are copied, so the matrix is bit-identical.
template <int PX>
__global__ void __launch_bounds__(256) im2col_stem_smem_kernel(const float* __restrict__ x, float* __restrict__ col, int h, int w,
                                                               int stride, int pad, int oh, int ow) {
  constexpr int KH = 7, KW = 7, C = 3, KPAD = 160, GROUPS = KPAD / 4, KREAL = KH * KW * C;
  extern __shared__ float4 win[];                      // [KH][wpx]
  __shared__ short lut[KPAD];                          // column k -> float offset of (r, s, ch) in the window; -1 = zero padding
  const int wpx = (PX - 1) * stride + KW;
  if (threadIdx.x < KPAD) {
    const int k = threadIdx.x;
    const int tap = k / C, ch = k - tap * C;
    const int r = tap / KW, s_ = tap - r * KW;
    lut[k] = k < KREAL ? (short)((r * wpx + s_) * 4 + ch) : (short)-1;
  }
  const int y_o = blockIdx.y % oh, img = blockIdx.y / oh;
  const int x0 = blockIdx.x * PX;
  const int ix0 = x0 * stride - pad, iy0 = y_o * stride - pad;
  const float4* __restrict__ ximg = reinterpret_cast<const float4*>(x) + (int64_t)img * h * w;
  for (int i = threadIdx.x; i < KH * wpx; i += blockDim.x) {
    const int r = i / wpx, px = i - r * wpx;
    const int iy = iy0 + r, ix = ix0 + px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(ximg + (int64_t)iy * w + ix);
    win[i] = v;
  }
  __syncthreads();
  const float* wf = reinterpret_cast<const float*>(win);
  const int npx = min(PX, ow - x0);
  float* __restrict__ crow = col + ((int64_t)blockIdx.y * ow + x0) * KPAD;
  for (int i = threadIdx.x; i < npx * GROUPS; i += blockDim.x) {
    const int xl = i / GROUPS, grp = i - xl * GROUPS;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = lut[grp * 4 + e];
      v[e] = o >= 0 ? wf[o + xl * stride * 4] : 0.f;
    }
    *reinterpret_cast<float4*>(crow + (int64_t)i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

extern "C" int b2_im2col(const float* x, float* col, int n, int h, int w, int c, int ldx, int kh, int kw, int stride, int pad,
                         int dil, int oh, int ow, int kpad, void* stream) {
  B2_REQUIRE(x && col && n > 0 && h > 0 && w > 0 && c > 0 && kpad >= kh * kw * c, "b2_im2col: bad args");
  B2_REQUIRE(kpad % 4 == 0 && (reinterpret_cast<uintptr_t>(col) & 15) == 0, "b2_im2col: kpad must be a multiple of 4 and col 16 B aligned");
  if (c == 3 && kh == 7 && kw == 7 && kpad == 160 && ldx == 4 && dil == 1 && stride >= 1 && stride <= 2 && (int64_t)n * oh <= 65535 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && !g_netops_flat_kernels) {
    constexpr int PX = 64;
    const int wpx = (PX - 1) * stride + 7;
    dim3 grid((unsigned)((ow + PX - 1) / PX), (unsigned)(n * oh));
    im2col_stem_smem_kernel<PX><<<grid, 256, 7 * wpx * sizeof(float4), (cudaStream_t)stream>>>(x, col, h, w, stride, pad, oh, ow);
    B2_LAUNCH_CHECK("im2col_stem_smem_kernel");
    return B2_OK;
  }
  if (c == 3 && kh == 7 && kw == 7 && kpad == 160 && (int64_t)n * oh <= 65535 && (int64_t)ow * 40 < (1ll << 30)) {
    // 8 stores per thread: one-store blocks are bound by the block launch rate (330 k blocks per stem at 512 x 512)
    dim3 grid((unsigned)((ow * 40 + 256 * 8 - 1) / (256 * 8)), (unsigned)(n * oh));
    im2col_const_kernel<3, 7, 7, 160><<<grid, 256, 0, (cudaStream_t)stream>>>(x, col, h, w, ldx, stride, pad, dil, oh, ow);
    B2_LAUNCH_CHECK("im2col_const_kernel");
    return B2_OK;
  }
  const int64_t total = (int64_t)n * oh * ow * (kpad / 4);
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 64) blocks = 148 * 64;
  if (total < (1ll << 31)) im2col_kernel<int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, col, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad);
  else im2col_kernel<int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, col, n, h, w, c, ldx, kh, kw, stride, pad, dil, oh, ow, kpad);
  B2_LAUNCH_CHECK("im2col_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ max pool 3x3 s2 p1
// V = 4: thread per (output pixel, 4 channels), 16 B loads / stores and one 32-bit store of the four argmax bytes
// (needs c % 4 == 0 and 16 B aligned tensors); V = 1: any channel count.  First maximum wins, NaN propagates (PyTorch).
template <int V, typename I>
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx,
                                                      int n, int h, int w, int c, int oh, int ow) {
  const int cg = c / V;
  const I total = (I)((int64_t)n * oh * ow * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; I t = i / cg;
    const int xo = (int)(t % ow); t /= ow; const int yo = (int)(t % oh); const int img = (int)(t / oh);
    float best[V]; int bi[V]; bool any = false;
#pragma unroll
    for (int e = 0; e < V; ++e) { best[e] = -CUDART_INF_F; bi[e] = 0; }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iy = yo * 2 - 1 + r, ix = xo * 2 - 1 + s;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          const float* src = x + (((int64_t)img * h + iy) * w + ix) * c + ch;
          float v[V];
          if (V == 4) { const float4 q = __ldg(reinterpret_cast<const float4*>(src)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
          else v[0] = __ldg(src);
#pragma unroll
          for (int e = 0; e < V; ++e)
            if (!any || v[e] > best[e] || v[e] != v[e]) { best[e] = v[e]; bi[e] = r * 3 + s; }
          any = true;
        }
      }
    const int64_t o = (int64_t)i * V;
    if (V == 4) {
      *reinterpret_cast<float4*>(y + o) = make_float4(best[0], best[1], best[2], best[3]);
      *reinterpret_cast<uint32_t*>(idx + o) = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    } else {
      y[o] = best[0]; idx[o] = (uint8_t)bi[0];
    }
  }
}
extern "C" int b2_maxpool3x3s2(const float* x, float* y, uint8_t* idx, int n, int h, int w, int c, int oh, int ow, void* stream) {
  B2_REQUIRE(x && y && idx && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "b2_maxpool3x3s2: bad args");
  const bool vec = c % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(idx) & 3) == 0;
  const int64_t total = (int64_t)n * oh * ow * (vec ? c / 4 : c);
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
  const bool i32 = total < (1ll << 31);
  if (vec && i32) maxpool_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, idx, n, h, w, c, oh, ow);
  else if (vec) maxpool_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, idx, n, h, w, c, oh, ow);
  else maxpool_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, idx, n, h, w, c, oh, ow);
  B2_LAUNCH_CHECK("maxpool_kernel");
  return B2_OK;
}
// gather form: each input pixel looks at the (up to 4) windows that contain it -> no atomics.
template <int V, typename I>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                          float* __restrict__ dx, int n, int h, int w, int c, int oh, int ow) {
  const int cg = c / V;
  const I total = (I)((int64_t)n * h * w * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; I t = i / cg;
    const int ix = (int)(t % w); t /= w; const int iy = (int)(t % h); const int img = (int)(t / h);
    float g[V];
#pragma unroll
    for (int e = 0; e < V; ++e) g[e] = 0.f;
    const int yo_lo = iy / 2, yo_hi = (iy + 1) / 2;   // windows with yo*2-1 <= iy <= yo*2+1
    const int xo_lo = ix / 2, xo_hi = (ix + 1) / 2;
    // up to 4 windows (visited in the order yo, xo: the summation order): all argmax words are requested first, then all matching
    // gradients -- two rounds of independent loads instead of four dependent index -> gradient chains (the serial form ran at
    // 1.4 TB/s, latency bound)
    int64_t off[4]; uint32_t want[4]; bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yo = (k >> 1) ? yo_hi : yo_lo, xo = (k & 1) ? xo_hi : xo_lo;
      const int r = iy - (yo * 2 - 1), s = ix - (xo * 2 - 1);
      ok[k] = yo < oh && xo < ow && r >= 0 && r <= 2 && s >= 0 && s <= 2 && !((k >> 1) && yo_hi == yo_lo) && !((k & 1) && xo_hi == xo_lo);
      off[k] = (((int64_t)img * oh + yo) * ow + xo) * c + ch;
      want[k] = (uint32_t)(r * 3 + s);
    }
    if (V == 4) {
      uint32_t kw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) kw[k] = ok[k] ? __ldg(reinterpret_cast<const uint32_t*>(idx + off[k])) : 0xffffffffu;
      float4 q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w_ = want[k];
        const bool any = ok[k] && (((kw[k] & 255u) == w_) | (((kw[k] >> 8) & 255u) == w_) | (((kw[k] >> 16) & 255u) == w_) | ((kw[k] >> 24) == w_));
        ok[k] = any;
        q[k] = any ? __ldg(reinterpret_cast<const float4*>(dy + off[k])) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ok[k]) {
          const uint32_t w_ = want[k];
          if ((kw[k] & 255u) == w_) g[0] += q[k].x;
          if (((kw[k] >> 8) & 255u) == w_) g[1] += q[k].y;
          if (((kw[k] >> 16) & 255u) == w_) g[2] += q[k].z;
          if ((kw[k] >> 24) == w_) g[3] += q[k].w;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k] && idx[off[k]] == want[k]) g[0] += __ldg(dy + off[k]);
    }
    if (V == 4) *reinterpret_cast<float4*>(dx + (int64_t)i * 4) = make_float4(g[0], g[1], g[2], g[3]);
    else dx[i] = g[0];
  }
}
extern "C" int b2_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int n, int h, int w, int c, int oh, int ow,
                                   void* stream) {
  B2_REQUIRE(dy && idx && dx && n > 0, "b2_maxpool3x3s2_bwd: bad args");
  const bool vec = c % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(idx) & 3) == 0;
  const int64_t total = (int64_t)n * h * w * (vec ? c / 4 : c);
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
  const bool i32 = total < (1ll << 31);
  if (vec && i32) maxpool_bwd_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, n, h, w, c, oh, ow);
  else if (vec) maxpool_bwd_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, n, h, w, c, oh, ow);
  else maxpool_bwd_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, idx, dx, n, h, w, c, oh, ow);
  B2_LAUNCH_CHECK("maxpool_bwd_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ bilinear
// PyTorch upsample_bilinear2d index math (UpSample.h area_pixel_compute_source_index), fp32.
struct LinCoef { int i0, i1; float l0, l1; };
__device__ __forceinline__ float lin_scale(int in, int out, int align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}
__device__ __forceinline__ LinCoef lin_coef(int dst, int in, float scale, int align) {
  float src;
  if (align) src = scale * dst;
  else { src = scale * (dst + 0.5f) - 0.5f; if (src < 0.f) src = 0.f; }
  LinCoef k;
  k.i0 = (int)src; if (k.i0 > in - 1) k.i0 = in - 1;
  k.i1 = k.i0 + (k.i0 < in - 1 ? 1 : 0);
  k.l1 = src - (float)k.i0; k.l0 = 1.f - k.l1;
  return k;
}

// NHWC -> NHWC : thread per (output pixel, V channels), channel fastest (V = 4: 16 B loads / stores).
template <int V, typename I>
__global__ void __launch_bounds__(256) bilinear_fwd_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int ih, int iw,
                                                                int c, int ldx, int oh, int ow, int ldy, int align) {
  const float sh = lin_scale(ih, oh, align), sw = lin_scale(iw, ow, align);
  const int cg = c / V;
  const I total = (I)((int64_t)n * oh * ow * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; I t = i / cg;
    const int xo = (int)(t % ow); t /= ow; const int yo = (int)(t % oh); const int img = (int)(t / oh);
    const LinCoef ky = lin_coef(yo, ih, sh, align), kx = lin_coef(xo, iw, sw, align);
    const float* b = x + (int64_t)img * ih * iw * ldx + ch;
    const float* p00 = b + ((int64_t)ky.i0 * iw + kx.i0) * ldx; const float* p01 = b + ((int64_t)ky.i0 * iw + kx.i1) * ldx;
    const float* p10 = b + ((int64_t)ky.i1 * iw + kx.i0) * ldx; const float* p11 = b + ((int64_t)ky.i1 * iw + kx.i1) * ldx;
    float* o = y + (((int64_t)img * oh + yo) * ow + xo) * ldy + ch;
    if (V == 4) {
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(p00)), v01 = __ldg(reinterpret_cast<const float4*>(p01));
      const float4 v10 = __ldg(reinterpret_cast<const float4*>(p10)), v11 = __ldg(reinterpret_cast<const float4*>(p11));
      float4 r;
      r.x = ky.l0 * (kx.l0 * v00.x + kx.l1 * v01.x) + ky.l1 * (kx.l0 * v10.x + kx.l1 * v11.x);
      r.y = ky.l0 * (kx.l0 * v00.y + kx.l1 * v01.y) + ky.l1 * (kx.l0 * v10.y + kx.l1 * v11.y);
      r.z = ky.l0 * (kx.l0 * v00.z + kx.l1 * v01.z) + ky.l1 * (kx.l0 * v10.z + kx.l1 * v11.z);
      r.w = ky.l0 * (kx.l0 * v00.w + kx.l1 * v01.w) + ky.l1 * (kx.l0 * v10.w + kx.l1 * v11.w);
      *reinterpret_cast<float4*>(o) = r;
    } else {
      *o = ky.l0 * (kx.l0 * __ldg(p00) + kx.l1 * __ldg(p01)) + ky.l1 * (kx.l0 * __ldg(p10) + kx.l1 * __ldg(p11));
    }
  }
}
// NHWC -> NCHW : thread per output pixel (xo fastest), loop over channels: coalesced plane writes.  V = 4 reads the four
// corner pixels 16 B at a time (needs ldx % 4 == 0 and a 16 B aligned x; channels past c inside the pitch are read, not stored).
template <int V>
__global__ void __launch_bounds__(128) bilinear_fwd_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int ih, int iw,
                                                                int c, int ldx, int oh, int ow, int align) {
  const float sh = lin_scale(ih, oh, align), sw = lin_scale(iw, ow, align);
  const int64_t ohw = (int64_t)oh * ow;
  const int64_t total = (int64_t)n * ohw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % ow); int64_t t = i / ow; const int yo = (int)(t % oh); const int img = (int)(t / oh);
    const LinCoef ky = lin_coef(yo, ih, sh, align), kx = lin_coef(xo, iw, sw, align);
    const float* b = x + (int64_t)img * ih * iw * ldx;
    const float* p00 = b + ((int64_t)ky.i0 * iw + kx.i0) * ldx; const float* p01 = b + ((int64_t)ky.i0 * iw + kx.i1) * ldx;
    const float* p10 = b + ((int64_t)ky.i1 * iw + kx.i0) * ldx; const float* p11 = b + ((int64_t)ky.i1 * iw + kx.i1) * ldx;
    float* o = y + (int64_t)img * c * ohw + (int64_t)yo * ow + xo;
    if (V == 4) {
      for (int ch = 0; ch < c; ch += 4) {
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(p00 + ch)), v01 = __ldg(reinterpret_cast<const float4*>(p01 + ch));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(p10 + ch)), v11 = __ldg(reinterpret_cast<const float4*>(p11 + ch));
        o[(int64_t)ch * ohw] = ky.l0 * (kx.l0 * v00.x + kx.l1 * v01.x) + ky.l1 * (kx.l0 * v10.x + kx.l1 * v11.x);
        if (ch + 1 < c) o[(int64_t)(ch + 1) * ohw] = ky.l0 * (kx.l0 * v00.y + kx.l1 * v01.y) + ky.l1 * (kx.l0 * v10.y + kx.l1 * v11.y);
        if (ch + 2 < c) o[(int64_t)(ch + 2) * ohw] = ky.l0 * (kx.l0 * v00.z + kx.l1 * v01.z) + ky.l1 * (kx.l0 * v10.z + kx.l1 * v11.z);
        if (ch + 3 < c) o[(int64_t)(ch + 3) * ohw] = ky.l0 * (kx.l0 * v00.w + kx.l1 * v01.w) + ky.l1 * (kx.l0 * v10.w + kx.l1 * v11.w);
      }
    } else {
      for (int ch = 0; ch < c; ++ch)
        o[(int64_t)ch * ohw] = ky.l0 * (kx.l0 * __ldg(p00 + ch) + kx.l1 * __ldg(p01 + ch)) + ky.l1 * (kx.l0 * __ldg(p10 + ch) + kx.l1 * __ldg(p11 + ch));
    }
  }
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// The same arithmetic with the thread -> channel group mapping fixed for the whole kernel: a block owns a run of output pixels of
// ONE output row (image and row from blockIdx.y, vertical coefficients computed once), thread = (channel group, pixel lane); no
// integer division in the pixel loop (the flat-index kernel above spends six run-time divisions per 16 B store: 2.25 TB/s at the
// 64 x 64 -> 128 x 128 resize of the DeepLab v3+ decoder, profiles/r02_v22_launch_list_summary.txt).  Needs 256 % (c / 4) == 0.
__global__ void __launch_bounds__(256) bilinear_fwd_nhwc_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int ih, int iw,
                                                                     int ldx, int oh, int ow, int ldy, int align, int cg, int ppb,
                                                                     int px_per_block) {
  const int lc = threadIdx.x % cg, lp = threadIdx.x / cg;
  const int yo = blockIdx.y % oh, img = blockIdx.y / oh;
  const float sh = lin_scale(ih, oh, align), sw = lin_scale(iw, ow, align);
  const LinCoef ky = lin_coef(yo, ih, sh, align);
  const float* b = x + (int64_t)img * ih * iw * ldx + lc * 4;
  const float* row0 = b + (int64_t)ky.i0 * iw * ldx;
  const float* row1 = b + (int64_t)ky.i1 * iw * ldx;
  float* orow = y + ((int64_t)img * oh + yo) * ow * ldy + lc * 4;
  const int x_begin = blockIdx.x * px_per_block;
  int x_end = x_begin + px_per_block; if (x_end > ow) x_end = ow;
  for (int xo = x_begin + lp; xo < x_end; xo += ppb) {
    const LinCoef kx = lin_coef(xo, iw, sw, align);
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(row0 + (int64_t)kx.i0 * ldx)), v01 = __ldg(reinterpret_cast<const float4*>(row0 + (int64_t)kx.i1 * ldx));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(row1 + (int64_t)kx.i0 * ldx)), v11 = __ldg(reinterpret_cast<const float4*>(row1 + (int64_t)kx.i1 * ldx));
    float4 r;
    r.x = ky.l0 * (kx.l0 * v00.x + kx.l1 * v01.x) + ky.l1 * (kx.l0 * v10.x + kx.l1 * v11.x);
    r.y = ky.l0 * (kx.l0 * v00.y + kx.l1 * v01.y) + ky.l1 * (kx.l0 * v10.y + kx.l1 * v11.y);
    r.z = ky.l0 * (kx.l0 * v00.z + kx.l1 * v01.z) + ky.l1 * (kx.l0 * v10.z + kx.l1 * v11.z);
    r.w = ky.l0 * (kx.l0 * v00.w + kx.l1 * v01.w) + ky.l1 * (kx.l0 * v10.w + kx.l1 * v11.w);
    *reinterpret_cast<float4*>(orow + (int64_t)xo * ldy) = r;
  }
}

extern "C" int b2_bilinear_fwd(const float* x, float* y, int n, int ih, int iw, int c, int ldx, int oh, int ow, int ldy,
                               int align_corners, int to_nchw, void* stream) {
  B2_REQUIRE(x && y && n > 0 && ih > 0 && iw > 0 && c > 0 && oh > 0 && ow > 0 && ldx >= c, "b2_bilinear_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (to_nchw) {
    const int64_t total = (int64_t)n * oh * ow;
    int64_t blocks = ceil_div64(total, 128); if (blocks > 148 * 64) blocks = 148 * 64;
    if (ldx % 4 == 0 && al16(x) && ldx >= (c + 3) / 4 * 4)
      bilinear_fwd_nchw_kernel<4><<<(unsigned)blocks, 128, 0, s>>>(x, y, n, ih, iw, c, ldx, oh, ow, align_corners);
    else
      bilinear_fwd_nchw_kernel<1><<<(unsigned)blocks, 128, 0, s>>>(x, y, n, ih, iw, c, ldx, oh, ow, align_corners);
  } else {
    B2_REQUIRE(ldy >= c, "b2_bilinear_fwd: ldy < c");
    const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y);
    const int64_t total = (int64_t)n * oh * ow * (vec ? c / 4 : c);
    int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
    const int cg = c / 4;
    if (vec && cg >= 1 && cg <= 256 && 256 % cg == 0 && (int64_t)n * oh <= 65535 && !g_netops_flat_kernels) {
      const int ppb = 256 / cg;                      // pixel lanes of a block
      const int px_per_block = ppb * 8;              // 8 pixels per thread
      dim3 grid((unsigned)((ow + px_per_block - 1) / px_per_block), (unsigned)(n * oh));
      bilinear_fwd_nhwc_rows_kernel<<<grid, 256, 0, s>>>(x, y, ih, iw, ldx, oh, ow, ldy, align_corners, cg, ppb, px_per_block);
    } else if (vec && total < (1ll << 31)) bilinear_fwd_nhwc_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, s>>>(x, y, n, ih, iw, c, ldx, oh, ow, ldy, align_corners);
    else if (vec) bilinear_fwd_nhwc_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, s>>>(x, y, n, ih, iw, c, ldx, oh, ow, ldy, align_corners);
    else bilinear_fwd_nhwc_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, s>>>(x, y, n, ih, iw, c, ldx, oh, ow, ldy, align_corners);
  }
  B2_LAUNCH_CHECK("bilinear_fwd");
  return B2_OK;
}

// Backward, gather form: input index i receives from every output o whose (i0 == i) or (i1 == i).
// Candidate range [lo, hi] of outputs is derived from the inverse map and widened; the exact
// forward coefficients decide membership, so the result is independent of the widening.
__device__ __forceinline__ void out_range(int i, int in, int out, float scale, int align, int* lo, int* hi) {
  if (scale <= 0.f) { *lo = 0; *hi = out - 1; return; }
  float a, b;
  if (align) { a = ((float)i - 1.f) / scale; b = ((float)i + 1.f) / scale; }
  else { a = ((float)i - 1.f + 0.5f) / scale - 0.5f; b = ((float)i + 1.f + 0.5f) / scale - 0.5f; }
  int l = (int)floorf(a) - 1, h = (int)ceilf(b) + 1;
  if (l < 0) l = 0; if (h > out - 1) h = out - 1;
  *lo = l; *hi = h;
}

// The forward coefficients of every output row / column are tabulated once per block in shared memory
// (tab[o] = {i0, i1, l1}; rows first, then columns), so the gather loops cost two LDS per candidate instead of
// re-deriving lin_coef (float -> int conversions) for every (row, column) pair.  The summation order (rows outer,
// columns inner, ascending) is the same for both source layouts.
struct LinTab { int i0, i1; float l1; };
constexpr int BIL_MAX_TAB = 3072;     // oh + ow entries (36 KB); larger outputs use the direct kernel below
__device__ __forceinline__ float tab_weight(const LinTab& k, int i, bool* hit) {
  float wgt = 0.f; *hit = false;
  if (k.i0 == i) { wgt += 1.f - k.l1; *hit = true; }
  if (k.i1 == i) { wgt += k.l1; *hit = true; }
  return wgt;
}
// FROM_NCHW: thread per input pixel (xi fastest: coalesced plane reads), loop over channels, each thread writes its
// pixel's contiguous channel vector.  Otherwise: thread per (input pixel, V channels), channel fastest.
template <bool FROM_NCHW, int V>
__global__ void __launch_bounds__(256) bilinear_bwd_tab_kernel(const float* __restrict__ dy, float* __restrict__ dx, int n, int ih, int iw,
                                                               int c, int ldx, int oh, int ow, int ldy, int align,
                                                               const float* __restrict__ scale_dev, float scale_host, int accumulate) {
  extern __shared__ LinTab tab[];
  const float sh = lin_scale(ih, oh, align), sw = lin_scale(iw, ow, align);
  for (int o = threadIdx.x; o < oh + ow; o += blockDim.x) {
    const LinCoef k = o < oh ? lin_coef(o, ih, sh, align) : lin_coef(o - oh, iw, sw, align);
    tab[o].i0 = k.i0; tab[o].i1 = k.i1; tab[o].l1 = k.l1;
  }
  __syncthreads();
  const LinTab* ytab = tab; const LinTab* xtab = tab + oh;
  const float gs = (scale_dev ? scale_dev[0] : 1.f) * scale_host;
  const int64_t ohw = (int64_t)oh * ow;
  const int cg = FROM_NCHW ? 1 : c / V;
  const int64_t total = (int64_t)n * ih * iw * cg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ch0 = 0, xi, yi, img;
    if (FROM_NCHW) { xi = (int)(i % iw); int64_t t = i / iw; yi = (int)(t % ih); img = (int)(t / ih); }
    else { ch0 = (int)(i % cg) * V; int64_t t = i / cg; xi = (int)(t % iw); t /= iw; yi = (int)(t % ih); img = (int)(t / ih); }
    int ylo, yhi, xlo, xhi;
    out_range(yi, ih, oh, sh, align, &ylo, &yhi);
    out_range(xi, iw, ow, sw, align, &xlo, &xhi);
    bool hit;
    // trim the candidate ranges to the outputs that really touch this input (membership by the exact coefficients)
    while (ylo <= yhi) { tab_weight(ytab[ylo], yi, &hit); if (hit) break; ++ylo; }
    while (yhi >= ylo) { tab_weight(ytab[yhi], yi, &hit); if (hit) break; --yhi; }
    while (xlo <= xhi) { tab_weight(xtab[xlo], xi, &hit); if (hit) break; ++xlo; }
    while (xhi >= xlo) { tab_weight(xtab[xhi], xi, &hit); if (hit) break; --xhi; }
    float* d = dx + (((int64_t)img * ih + yi) * iw + xi) * ldx + ch0;
    const int ny = yhi - ylo + 1, nx = xhi - xlo + 1;
    // The outputs touching one input form a contiguous run (the source index is monotone), so after trimming every
    // candidate is a hit; runs of <= 8 (any up-sampling factor <= 4) keep their weights in registers.
    const bool small = ny <= 8 && nx <= 8;
    float wyr[8], wxr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wyr[j] = (small && j < ny) ? tab_weight(ytab[ylo + j], yi, &hit) : 0.f;
      wxr[j] = (small && j < nx) ? tab_weight(xtab[xlo + j], xi, &hit) : 0.f;
    }
    if (FROM_NCHW) {
      for (int ch = 0; ch < c; ++ch) {
        const float* plane = dy + ((int64_t)img * c + ch) * ohw;
        float acc = 0.f;
        if (small) {
#pragma unroll
          for (int jy = 0; jy < 8; ++jy) {
            if (jy < ny) {
              const float* rowp = plane + (int64_t)(ylo + jy) * ow + xlo;
              float rowacc = 0.f;
#pragma unroll
              for (int jx = 0; jx < 8; ++jx)
                if (jx < nx) rowacc += wxr[jx] * __ldg(rowp + jx);
              acc += wyr[jy] * rowacc;
            }
          }
        } else {
          for (int yo = ylo; yo <= yhi; ++yo) {
            const float wy = tab_weight(ytab[yo], yi, &hit);
            if (!hit) continue;
            float rowacc = 0.f;
            for (int xo = xlo; xo <= xhi; ++xo) {
              const float wx = tab_weight(xtab[xo], xi, &hit);
              if (!hit) continue;
              rowacc += wx * __ldg(plane + (int64_t)yo * ow + xo);
            }
            acc += wy * rowacc;
          }
        }
        d[ch] = accumulate ? d[ch] + acc * gs : acc * gs;
      }
    } else {
      float acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
      if (small) {
#pragma unroll
        for (int jy = 0; jy < 8; ++jy) {
          if (jy < ny) {
            const float* rowp = dy + (((int64_t)img * oh + ylo + jy) * ow + xlo) * ldy + ch0;
            float rowacc[V];
#pragma unroll
            for (int e = 0; e < V; ++e) rowacc[e] = 0.f;
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
              if (jx < nx) {
                if (V == 4) {
                  const float4 g = __ldg(reinterpret_cast<const float4*>(rowp + (int64_t)jx * ldy));
                  rowacc[0] += wxr[jx] * g.x; rowacc[1] += wxr[jx] * g.y; rowacc[2] += wxr[jx] * g.z; rowacc[3] += wxr[jx] * g.w;
                } else {
                  rowacc[0] += wxr[jx] * __ldg(rowp + (int64_t)jx * ldy);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] += wyr[jy] * rowacc[e];
          }
        }
      } else {
        for (int yo = ylo; yo <= yhi; ++yo) {
          const float wy = tab_weight(ytab[yo], yi, &hit);
          if (!hit) continue;
          float rowacc[V];
#pragma unroll
          for (int e = 0; e < V; ++e) rowacc[e] = 0.f;
          for (int xo = xlo; xo <= xhi; ++xo) {
            const float wx = tab_weight(xtab[xo], xi, &hit);
            if (!hit) continue;
            const float* src = dy + (((int64_t)img * oh + yo) * ow + xo) * ldy + ch0;
            if (V == 4) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(src));
              rowacc[0] += wx * g.x; rowacc[1] += wx * g.y; rowacc[2] += wx * g.z; rowacc[3] += wx * g.w;
            } else {
              rowacc[0] += wx * __ldg(src);
            }
          }
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += wy * rowacc[e];
        }
      }
      if (V == 4) {
        float4 o = make_float4(acc[0] * gs, acc[1] * gs, acc[2] * gs, acc[3] * gs);
        if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(d); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
        *reinterpret_cast<float4*>(d) = o;
      } else {
        *d = accumulate ? *d + acc[0] * gs : acc[0] * gs;
      }
    }
  }
}

// Separable backward of the final NHWC -> NCHW resize (dY is NCHW, 15x larger than dX): a horizontal pass reduces
// every output row to the input width (coalesced reads of dY, tmp[n,c,yo,xi] = sum_xo wx * dY[n,c,yo,xo]), a vertical
// pass finishes dX[n,yi,xi,c] = gs * sum_yo wy * tmp[n,c,yo,xi].  Same grouping and order of the sums as the
// gather kernel above (rows of x-sums, then the y-sum), i.e. bit-identical results, with ~1.6x the bytes of dY moved
// instead of every input pixel re-reading an 8 x 8 window per channel.
// The run of outputs touching input column xi and its weights depend on xi only: they are worked out ONCE per block into shared
// memory (first output xlo[xi], count nx[xi], weights w[j][xi] -- the very floats tab_weight() returns, so the sums below are
// bit-identical to the per-element form), and the row loop is 8 predicated loads + FMAs per element instead of two trimming loops
// and eight table evaluations (the per-element form ran at 0.64 TB/s, instruction bound; profiles/r02_v11_device_timeline_*).
__global__ void __launch_bounds__(256) bilinear_bwd_h_kernel(const float* __restrict__ dy, float* __restrict__ tmp, int64_t planes_rows,
                                                             int iw, int ow, int align) {
  extern __shared__ LinTab tab[];
  int* s_xlo = reinterpret_cast<int*>(tab + ow);
  int* s_nx = s_xlo + iw;
  float* s_w = reinterpret_cast<float*>(s_nx + iw);          // [8][iw]
  const float sw = lin_scale(iw, ow, align);
  for (int o = threadIdx.x; o < ow; o += blockDim.x) {
    const LinCoef k = lin_coef(o, iw, sw, align);
    tab[o].i0 = k.i0; tab[o].i1 = k.i1; tab[o].l1 = k.l1;
  }
  __syncthreads();
  for (int xi = threadIdx.x; xi < iw; xi += blockDim.x) {
    int xlo, xhi; bool hit;
    out_range(xi, iw, ow, sw, align, &xlo, &xhi);
    while (xlo <= xhi) { tab_weight(tab[xlo], xi, &hit); if (hit) break; ++xlo; }
    while (xhi >= xlo) { tab_weight(tab[xhi], xi, &hit); if (hit) break; --xhi; }
    const int nx = xhi - xlo + 1;
    s_xlo[xi] = xlo; s_nx[xi] = nx;
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) s_w[jx * iw + xi] = (nx <= 8 && jx < nx) ? tab_weight(tab[xlo + jx], xi, &hit) : 0.f;
  }
  __syncthreads();
  const int64_t total = planes_rows * iw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xi = (int)(i % iw); const int64_t pr = i / iw;
    const int xlo = s_xlo[xi], nx = s_nx[xi];
    const float* rowp = dy + pr * ow;
    float rowacc = 0.f;
    if (nx <= 8) {
#pragma unroll
      for (int jx = 0; jx < 8; ++jx)
        if (jx < nx) rowacc += s_w[jx * iw + xi] * __ldg(rowp + xlo + jx);
    } else {
      bool hit;
      for (int xo = xlo; xo < xlo + nx; ++xo) {
        const float wx = tab_weight(tab[xo], xi, &hit);
        if (hit) rowacc += wx * __ldg(rowp + xo);
      }
    }
    tmp[i] = rowacc;
  }
}
__global__ void __launch_bounds__(128) bilinear_bwd_v_kernel(const float* __restrict__ tmp, float* __restrict__ dx, int n, int ih, int iw, int c,
                                                             int ldx, int oh, int align, const float* __restrict__ scale_dev,
                                                             float scale_host, int accumulate) {
  extern __shared__ LinTab tab[];
  const float sh = lin_scale(ih, oh, align);
  for (int o = threadIdx.x; o < oh; o += blockDim.x) {
    const LinCoef k = lin_coef(o, ih, sh, align);
    tab[o].i0 = k.i0; tab[o].i1 = k.i1; tab[o].l1 = k.l1;
  }
  __syncthreads();
  const float gs = (scale_dev ? scale_dev[0] : 1.f) * scale_host;
  const int64_t total = (int64_t)n * ih * iw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xi = (int)(i % iw); int64_t t = i / iw; const int yi = (int)(t % ih); const int img = (int)(t / ih);
    int ylo, yhi; bool hit;
    out_range(yi, ih, oh, sh, align, &ylo, &yhi);
    while (ylo <= yhi) { tab_weight(tab[ylo], yi, &hit); if (hit) break; ++ylo; }
    while (yhi >= ylo) { tab_weight(tab[yhi], yi, &hit); if (hit) break; --yhi; }
    const int ny = yhi - ylo + 1;
    const bool small = ny <= 8;
    float wyr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wyr[j] = (small && j < ny) ? tab_weight(tab[ylo + j], yi, &hit) : 0.f;
    float* d = dx + (((int64_t)img * ih + yi) * iw + xi) * ldx;
    for (int ch = 0; ch < c; ++ch) {
      const float* col = tmp + (((int64_t)img * c + ch) * oh) * iw + xi;
      float acc = 0.f;
      if (small) {
#pragma unroll
        for (int jy = 0; jy < 8; ++jy)
          if (jy < ny) acc += wyr[jy] * __ldg(col + (int64_t)(ylo + jy) * iw);
      } else {
        for (int yo = ylo; yo <= yhi; ++yo) {
          const float wy = tab_weight(tab[yo], yi, &hit);
          if (hit) acc += wy * __ldg(col + (int64_t)yo * iw);
        }
      }
      d[ch] = accumulate ? d[ch] + acc * gs : acc * gs;
    }
  }
}
extern "C" int64_t b2_bilinear_bwd_nchw_workspace_floats(int n, int c, int iw, int oh) { return (int64_t)n * c * oh * iw; }
extern "C" int b2_bilinear_bwd_nchw(const float* dy, float* dx, float* workspace, int n, int ih, int iw, int c, int ldx, int oh, int ow,
                                    int align_corners, const float* scale_dev, float scale_host, int accumulate, void* stream) {
  B2_REQUIRE(dy && dx && workspace && n > 0 && ih > 0 && iw > 0 && c > 0 && oh > 0 && ow > 0 && ldx >= c, "b2_bilinear_bwd_nchw: bad args");
  B2_REQUIRE(oh <= BIL_MAX_TAB && ow <= BIL_MAX_TAB, "b2_bilinear_bwd_nchw: output larger than %d", BIL_MAX_TAB);
  B2_REQUIRE((size_t)ow * sizeof(LinTab) + (size_t)iw * 10 * sizeof(int) <= 48 * 1024,
             "b2_bilinear_bwd_nchw: tables of %d output / %d input columns exceed 48 KB (use b2_bilinear_bwd)", ow, iw);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t planes_rows = (int64_t)n * c * oh;
  int64_t b1 = ceil_div64(planes_rows * iw, 256); if (b1 > 148 * 32) b1 = 148 * 32;
  bilinear_bwd_h_kernel<<<(unsigned)b1, 256, (size_t)ow * sizeof(LinTab) + (size_t)iw * 10 * sizeof(int), s>>>(dy, workspace, planes_rows, iw, ow, align_corners);
  B2_LAUNCH_CHECK("bilinear_bwd_h_kernel");
  int64_t b2 = ceil_div64((int64_t)n * ih * iw, 128); if (b2 > 148 * 16) b2 = 148 * 16;
  bilinear_bwd_v_kernel<<<(unsigned)b2, 128, (size_t)oh * sizeof(LinTab), s>>>(workspace, dx, n, ih, iw, c, ldx, oh, align_corners, scale_dev,
                                                                             scale_host, accumulate);
  B2_LAUNCH_CHECK("bilinear_bwd_v_kernel");
  return B2_OK;
}

// Direct form (no tables): any output size.
template <bool FROM_NCHW>
__global__ void bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int n, int ih, int iw, int c, int ldx,
                                    int oh, int ow, int ldy, int align, const float* __restrict__ scale_dev, float scale_host,
                                    int accumulate) {
  const float sh = lin_scale(ih, oh, align), sw = lin_scale(iw, ow, align);
  const float gs = (scale_dev ? scale_dev[0] : 1.f) * scale_host;
  const int64_t total = (int64_t)n * ih * iw * c;
  const int64_t ohw = (int64_t)oh * ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ch, xi, yi, img;
    if (FROM_NCHW) {  // (n, c, yi, xi), xi fastest: coalesced reads of dy planes
      xi = (int)(i % iw); int64_t t = i / iw; yi = (int)(t % ih); t /= ih; ch = (int)(t % c); img = (int)(t / c);
    } else {          // (n, yi, xi, c), c fastest
      ch = (int)(i % c); int64_t t = i / c; xi = (int)(t % iw); t /= iw; yi = (int)(t % ih); img = (int)(t / ih);
    }
    int ylo, yhi, xlo, xhi;
    out_range(yi, ih, oh, sh, align, &ylo, &yhi);
    out_range(xi, iw, ow, sw, align, &xlo, &xhi);
    float acc = 0.f;
    for (int yo = ylo; yo <= yhi; ++yo) {
      const LinCoef ky = lin_coef(yo, ih, sh, align);
      float wy = 0.f;
      if (ky.i0 == yi) wy += ky.l0;
      if (ky.i1 == yi) wy += ky.l1;
      if (wy == 0.f && !(ky.i0 == yi || ky.i1 == yi)) continue;
      float rowacc = 0.f;
      for (int xo = xlo; xo <= xhi; ++xo) {
        const LinCoef kx = lin_coef(xo, iw, sw, align);
        float wx = 0.f; bool hit = false;
        if (kx.i0 == xi) { wx += kx.l0; hit = true; }
        if (kx.i1 == xi) { wx += kx.l1; hit = true; }
        if (!hit) continue;
        const float g = FROM_NCHW ? __ldg(dy + ((int64_t)img * c + ch) * ohw + (int64_t)yo * ow + xo)
                                  : __ldg(dy + (((int64_t)img * oh + yo) * ow + xo) * ldy + ch);
        rowacc += wx * g;
      }
      acc += wy * rowacc;
    }
    float* d = dx + (((int64_t)img * ih + yi) * iw + xi) * ldx + ch;
    *d = accumulate ? *d + acc * gs : acc * gs;
  }
}
extern "C" int b2_bilinear_bwd(const float* dy, float* dx, int n, int ih, int iw, int c, int ldx, int oh, int ow, int ldy,
                               int align_corners, int from_nchw, const float* scale_dev, float scale_host, int accumulate,
                               void* stream) {
  B2_REQUIRE(dy && dx && n > 0 && ih > 0 && iw > 0 && c > 0 && oh > 0 && ow > 0 && ldx >= c, "b2_bilinear_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (oh + ow <= BIL_MAX_TAB) {
    const size_t smem = (size_t)(oh + ow) * sizeof(LinTab);
    if (from_nchw) {
      const int64_t total = (int64_t)n * ih * iw;
      int64_t blocks = ceil_div64(total, 128); if (blocks > 148 * 16) blocks = 148 * 16;
      bilinear_bwd_tab_kernel<true, 1><<<(unsigned)blocks, 128, smem, s>>>(dy, dx, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, scale_dev, scale_host, accumulate);
    } else {
      const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(dy) && al16(dx);
      const int64_t total = (int64_t)n * ih * iw * (vec ? c / 4 : c);
      int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 16) blocks = 148 * 16;
      if (vec) bilinear_bwd_tab_kernel<false, 4><<<(unsigned)blocks, 256, smem, s>>>(dy, dx, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, scale_dev, scale_host, accumulate);
      else bilinear_bwd_tab_kernel<false, 1><<<(unsigned)blocks, 256, smem, s>>>(dy, dx, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, scale_dev, scale_host, accumulate);
    }
    B2_LAUNCH_CHECK("bilinear_bwd_tab_kernel");
    return B2_OK;
  }
  const int64_t total = (int64_t)n * ih * iw * c;
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (from_nchw)
    bilinear_bwd_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(dy, dx, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, scale_dev, scale_host, accumulate);
  else
    bilinear_bwd_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(dy, dx, n, ih, iw, c, ldx, oh, ow, ldy, align_corners, scale_dev, scale_host, accumulate);
  B2_LAUNCH_CHECK("bilinear_bwd_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ column reductions
// partial[chunk][c][k] = sum over the chunk's rows of f_k(row, c), k < 2, in double.
// MODE 0: (x, -)            colsum
// MODE 1: (x, x*x)          BN statistics
// MODE 2: (g, g*xhat)       BN backward, g = dy * gate(y) * drop, xhat = (x - mean) * rstd
// MODE 3: (g, g*y)          frozen-BN parameter gradients (y = BN output proxy), g = dy gated by gate>0
__device__ __forceinline__ float bn_affine(float x, float mean, float rstd, float gamma, float beta);
constexpr int RED_ROWS_PER_CHUNK = 2048;
// Rows per chunk for a (rows x c) reduction: 2048, halved (down to 256) until the grid has >= 4 blocks per SM -- with 2048 rows
// per block a 65536 x 256 tensor gave 256 blocks = 1.7 per SM, too few bytes in flight to cover the HBM latency
// (col_reduce_kernel at 3.8 TB/s in profiles/r02_v7_launch_list_summary.txt).
static inline int red_rows_per_chunk(int64_t rows, int c) {
  int rpc = RED_ROWS_PER_CHUNK;
  const int64_t cb = (c + 31) / 32;
  while (rpc > 256 && ((rows + rpc - 1) / rpc) * cb < 148 * 4) rpc >>= 1;
  return rpc;
}
struct RedArgs {
  const float* a; int lda;      // dy or x
  const float* b; int ldb;      // x (mode 2) / y (mode 3)
  const float* gate; int ldg;   // relu gate tensor (y > 0) or NULL
  const float* drop; float drop_scale;  // dropout mask laid out like dy (ld = lda) or NULL
  const float* mean; const float* rstd;
  const float* gamma; const float* gate_beta;   // mode 2, gate_beta != NULL: gate = bn_affine(x) > 0 recomputed from x (no gate tensor)
  const float* sub; int lds;    // mode 3: o = b - sub (residual removed from the block output)
  int64_t rows; int c;
  int vec;                      // every pointer 16 B aligned and every ld / c a multiple of 4
  int rpc;                      // rows per chunk (red_rows_per_chunk)
};
// Block = 32 channels x one chunk of rows.  Thread (cg = tid & 7, rl = tid >> 3) owns 4 consecutive channels and walks
// rows rl, rl+32, ... with 16 B loads: 8 lanes cover a 128 B row segment, a warp covers 4 rows per instruction, and the
// loop is unrolled so that 8-12 independent 16 B loads per thread are in flight (HBM-bound kernel).
template <int MODE>
__global__ void __launch_bounds__(256) col_reduce_kernel(RedArgs r, double* __restrict__ partial) {
  __shared__ double sm[8][32][2];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch0 = blockIdx.y * 32 + cg * 4;
  const int64_t r0 = (int64_t)blockIdx.x * r.rpc;
  int64_t r1 = r0 + r.rpc; if (r1 > r.rows) r1 = r.rows;
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float mean[4] = {0, 0, 0, 0}, rstd[4] = {1, 1, 1, 1}, gam[4] = {0, 0, 0, 0}, bet[4] = {0, 0, 0, 0};
  const bool regate = MODE == 2 && r.gate_beta != nullptr;
  if (MODE == 2) {
#pragma unroll
    for (int e = 0; e < 4; ++e) if (ch0 + e < r.c) {
      mean[e] = r.mean[ch0 + e]; rstd[e] = r.rstd[ch0 + e];
      if (regate) { gam[e] = r.gamma[ch0 + e]; bet[e] = r.gate_beta[ch0 + e]; }
    }
  }
  if (MODE <= 1 && r.vec && ch0 + 3 < r.c) {
    // one tensor: four independent 16 B loads per thread in flight (rows rl, rl + 32, rl + 64, rl + 96 of a 128-row group)
    int64_t row = r0 + rl;
    for (; row + 96 < r1; row += 128) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const float4*>(r.a + (row + 32 * u) * r.lda + ch0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float va[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { s0[e] += va[e]; if (MODE == 1) s1[e] += (double)va[e] * (double)va[e]; }
      }
    }
    for (; row < r1; row += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(r.a + row * r.lda + ch0));
      const float va[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { s0[e] += va[e]; if (MODE == 1) s1[e] += (double)va[e] * (double)va[e]; }
    }
  } else if (r.vec && ch0 + 3 < r.c) {
#pragma unroll 2
    for (int64_t row = r0 + rl; row < r1; row += 32) {
      float4 v = __ldg(reinterpret_cast<const float4*>(r.a + row * r.lda + ch0));
      float va[4] = {v.x, v.y, v.z, v.w};
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s0[e] += va[e];
      } else if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { s0[e] += va[e]; s1[e] += (double)va[e] * (double)va[e]; }
      } else {
        const float4 bq = __ldg(reinterpret_cast<const float4*>(r.b + row * r.ldb + ch0));
        float vb[4] = {bq.x, bq.y, bq.z, bq.w};
        if (regate) {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (!(bn_affine(vb[e], mean[e], rstd[e], gam[e], bet[e]) > 0.f)) va[e] = 0.f;
        } else if (r.gate) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(r.gate + row * r.ldg + ch0));
          if (!(g.x > 0.f)) va[0] = 0.f; if (!(g.y > 0.f)) va[1] = 0.f; if (!(g.z > 0.f)) va[2] = 0.f; if (!(g.w > 0.f)) va[3] = 0.f;
        }
        if (r.drop) {
          const float4 d = __ldg(reinterpret_cast<const float4*>(r.drop + row * r.lda + ch0));
          va[0] *= d.x * r.drop_scale; va[1] *= d.y * r.drop_scale; va[2] *= d.z * r.drop_scale; va[3] *= d.w * r.drop_scale;
        }
        if (MODE == 3 && r.sub) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(r.sub + row * r.lds + ch0));
          vb[0] -= q.x; vb[1] -= q.y; vb[2] -= q.z; vb[3] -= q.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float o = MODE == 2 ? (vb[e] - mean[e]) * rstd[e] : vb[e];
          s0[e] += va[e]; s1[e] += (double)va[e] * (double)o;
        }
      }
    }
  } else {
    for (int64_t row = r0 + rl; row < r1; row += 32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = ch0 + e;
        if (ch >= r.c) continue;
        float v = __ldg(r.a + row * r.lda + ch);
        if (MODE == 0) { s0[e] += v; }
        else if (MODE == 1) { s0[e] += v; s1[e] += (double)v * (double)v; }
        else {
          float bv = __ldg(r.b + row * r.ldb + ch);
          if (regate) { if (!(bn_affine(bv, mean[e], rstd[e], gam[e], bet[e]) > 0.f)) v = 0.f; }
          else if (r.gate && !(__ldg(r.gate + row * r.ldg + ch) > 0.f)) v = 0.f;
          if (r.drop) v *= __ldg(r.drop + row * r.lda + ch) * r.drop_scale;
          if (MODE == 3 && r.sub) bv -= __ldg(r.sub + row * r.lds + ch);
          const float o = MODE == 2 ? (bv - mean[e]) * rstd[e] : bv;
          s0[e] += v; s1[e] += (double)v * (double)o;
        }
      }
    }
  }
  // lanes l, l^8, l^16, l^24 hold the same channels (different rows): fixed-order shuffle tree, then 8 warps via smem
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 8);  s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
    s0[e] += __shfl_xor_sync(0xffffffffu, s0[e], 16); s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { sm[warp][lane * 4 + e][0] = s0[e]; sm[warp][lane * 4 + e][1] = s1[e]; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int ch = blockIdx.y * 32 + threadIdx.x;
    if (ch < r.c) {
      double t0 = 0, t1 = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { t0 += sm[k][threadIdx.x][0]; t1 += sm[k][threadIdx.x][1]; }
      partial[((int64_t)blockIdx.x * r.c + ch) * 2 + 0] = t0;
      partial[((int64_t)blockIdx.x * r.c + ch) * 2 + 1] = t1;
    }
  }
}
static inline int64_t red_chunks(int64_t rows, int c) { return ceil_div64(rows, red_rows_per_chunk(rows, c)); }
extern "C" int64_t b2_bn_workspace_doubles(int64_t rows, int c) { return red_chunks(rows, c) * c * 2 + 2 * (int64_t)c; }

template <int MODE>
static int launch_col_reduce(const RedArgs& r0, double* ws, cudaStream_t s) {
  RedArgs r = r0;
  auto ok = [](const void* p, int ld) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0); };
  r.vec = (r.c % 4 == 0) && ok(r.a, r.lda) && ok(r.b, r.ldb) && ok(r.gate, r.ldg) && ok(r.sub, r.lds) && ok(r.drop, r.lda);
  r.rpc = red_rows_per_chunk(r.rows, r.c);
  dim3 grid((unsigned)red_chunks(r.rows, r.c), (r.c + 31) / 32);
  col_reduce_kernel<MODE><<<grid, 256, 0, s>>>(r, ws);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return b2_fail(B2_ERR_CUDA, "col_reduce launch failed: %s", cudaGetErrorString(e));
  return B2_OK;
}

// final[c][k] = sum over chunks, fixed order: block = 32 channels x 8 chunk lanes (lane j sums chunks j, j+8, ...; the
// 8 lane sums are then added in ascending order) -- deterministic, and 8x less serial latency than a thread per channel.
__global__ void __launch_bounds__(256) col_finalize_kernel(const double* __restrict__ partial, int64_t chunks, int c,
                                                           double* __restrict__ fin) {
  __shared__ double sm[8][32][2];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  double t0 = 0, t1 = 0;
  if (ch < c) {
    for (int64_t k = ry; k < chunks; k += 8) {
      const double2 v = *reinterpret_cast<const double2*>(partial + (k * c + ch) * 2);
      t0 += v.x; t1 += v.y;
    }
  }
  sm[ry][cx][0] = t0; sm[ry][cx][1] = t1;
  __syncthreads();
  if (ry == 0 && ch < c) {
    double a0 = 0, a1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a0 += sm[k][cx][0]; a1 += sm[k][cx][1]; }
    fin[ch * 2] = a0; fin[ch * 2 + 1] = a1;
  }
}
static inline void launch_col_finalize(const double* partial, int64_t chunks, int c, double* fin, cudaStream_t s) {
  col_finalize_kernel<<<(c + 31) / 32, 256, 0, s>>>(partial, chunks, c, fin);
}

__global__ void colsum_out_kernel(const double* __restrict__ fin, int c, float* __restrict__ out, int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch < c) out[ch] = (accumulate ? out[ch] : 0.f) + (float)fin[ch * 2];
}
extern "C" int b2_colsum(const float* dy, int ld, int64_t rows, int c, float* out, int accumulate, double* workspace, void* stream) {
  B2_REQUIRE(dy && out && workspace && rows > 0 && c > 0 && ld >= c, "b2_colsum: bad args");
  RedArgs r{}; r.a = dy; r.lda = ld; r.rows = rows; r.c = c;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_col_reduce<0>(r, workspace, s); if (rc) return rc;
  double* fin = workspace + red_chunks(rows, c) * c * 2;
  launch_col_finalize(workspace, red_chunks(rows, c), c, fin, s);
  colsum_out_kernel<<<(c + 127) / 128, 128, 0, s>>>(fin, c, out, accumulate);
  B2_LAUNCH_CHECK("colsum");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ batch norm (train mode)
__global__ void bn_stats_out_kernel(const double* __restrict__ fin, int c, int64_t rows, float eps, float momentum,
                                    float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ rm, float* __restrict__ rv) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double n = (double)rows;
  const double m = fin[ch * 2] / n;
  double var = fin[ch * 2 + 1] / n - m * m; if (var < 0) var = 0;
  mean[ch] = (float)m;
  rstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (rm) rm[ch] = (1.f - momentum) * rm[ch] + momentum * (float)m;
  if (rv) {
    const double unb = rows > 1 ? var * n / (n - 1.0) : var;
    rv[ch] = (1.f - momentum) * rv[ch] + momentum * (float)unb;
  }
}
extern "C" int b2_bn_stats(const float* x, int64_t rows, int c, int ldx, float eps, float momentum, float* mean, float* rstd,
                           float* running_mean, float* running_var, double* workspace, void* stream) {
  B2_REQUIRE(x && mean && rstd && workspace && rows > 0 && c > 0 && ldx >= c, "b2_bn_stats: bad args");
  RedArgs r{}; r.a = x; r.lda = ldx; r.rows = rows; r.c = c;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_col_reduce<1>(r, workspace, s); if (rc) return rc;
  double* fin = workspace + red_chunks(rows, c) * c * 2;
  launch_col_finalize(workspace, red_chunks(rows, c), c, fin, s);
  bn_stats_out_kernel<<<(c + 127) / 128, 128, 0, s>>>(fin, c, rows, eps, momentum, mean, rstd, running_mean, running_var);
  B2_LAUNCH_CHECK("bn_stats");
  return B2_OK;
}

// y = gamma * (rstd * (x - mean)) + beta with the roundings bn_apply_kernel has always had (subtract, multiply, fused
// multiply-add); bn_bwd recomputes the sign of y from x with the SAME three instructions instead of reading y (gate_beta).
__device__ __forceinline__ float bn_affine(float x, float mean, float rstd, float gamma, float beta) {
  return __fmaf_rn(gamma, __fmul_rn(rstd, __fsub_rn(x, mean)), beta);
}

template <int V, typename I>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, int64_t rows, int c, int ldx,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                       const float* __restrict__ drop, float drop_scale, float* __restrict__ y, int ldy,
                                                       const float* __restrict__ res, int ldr) {
  const int cg = c / V;
  const I total = (I)(rows * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; const int64_t row = (int64_t)(i / cg);
    float v[V], m[V], r[V], g[V], b[V];
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(x + row * ldx + ch)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      const float4 qm = __ldg(reinterpret_cast<const float4*>(mean + ch)); m[0] = qm.x; m[1] = qm.y; m[2] = qm.z; m[3] = qm.w;
      const float4 qr = __ldg(reinterpret_cast<const float4*>(rstd + ch)); r[0] = qr.x; r[1] = qr.y; r[2] = qr.z; r[3] = qr.w;
      const float4 qg = __ldg(reinterpret_cast<const float4*>(gamma + ch)); g[0] = qg.x; g[1] = qg.y; g[2] = qg.z; g[3] = qg.w;
      const float4 qb = __ldg(reinterpret_cast<const float4*>(beta + ch)); b[0] = qb.x; b[1] = qb.y; b[2] = qb.z; b[3] = qb.w;
    } else {
      v[0] = x[row * ldx + ch]; m[0] = mean[ch]; r[0] = rstd[ch]; g[0] = gamma[ch]; b[0] = beta[ch];
    }
#pragma unroll
    for (int e = 0; e < V; ++e) v[e] = bn_affine(v[e], m[e], r[e], g[e], b[e]);
    if (res) {
      if (V == 4) { const float4 q = __ldg(reinterpret_cast<const float4*>(res + row * ldr + ch)); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
      else v[0] += res[row * ldr + ch];
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    if (drop) {
      if (V == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(drop + row * c + ch));
        v[0] *= q.x * drop_scale; v[1] *= q.y * drop_scale; v[2] *= q.z * drop_scale; v[3] *= q.w * drop_scale;
      } else v[0] *= drop[row * c + ch] * drop_scale;
    }
    if (V == 4) *reinterpret_cast<float4*>(y + row * ldy + ch) = make_float4(v[0], v[1], v[2], v[3]);
    else y[row * ldy + ch] = v[0];
  }
}
extern "C" int b2_bn_apply(const float* x, int64_t rows, int c, int ldx, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, int relu, const float* dropmask, float drop_scale, float* y, int ldy,
                           const float* residual, int ldr, void* stream) {
  B2_REQUIRE(x && y && mean && rstd && gamma && beta && rows > 0 && c > 0, "b2_bn_apply: bad args");
  const bool vec = c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y) && al16(mean) && al16(rstd) && al16(gamma) &&
                   al16(beta) && (!residual || (ldr % 4 == 0 && al16(residual))) && (!dropmask || al16(dropmask));
  const int64_t total = rows * (vec ? c / 4 : c);
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec && total < (1ll << 31)) bn_apply_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, c, ldx, mean, rstd, gamma, beta, relu, dropmask, drop_scale, y, ldy, residual, ldr);
  else if (vec) bn_apply_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, c, ldx, mean, rstd, gamma, beta, relu, dropmask, drop_scale, y, ldy, residual, ldr);
  else bn_apply_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, c, ldx, mean, rstd, gamma, beta, relu, dropmask, drop_scale, y, ldy, residual, ldr);
  B2_LAUNCH_CHECK("bn_apply_kernel");
  return B2_OK;
}

template <int V, typename I>
__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                        const float* __restrict__ y, int ldy, int64_t rows, int c,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, int relu, const float* __restrict__ drop,
                                                        float drop_scale, const double* __restrict__ fin, float* __restrict__ dx, int lddx,
                                                        float* __restrict__ g_out, int ldgo, const float* __restrict__ gate_beta) {
  const int cg = c / V;
  const I total = (I)(rows * cg);
  const double inv_n = 1.0 / (double)rows;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; const int64_t row = (int64_t)(i / cg);
    float g[V], xv[V], yv[V], dv[V];
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + ch)); g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
      const float4 qx = __ldg(reinterpret_cast<const float4*>(x + row * ldx + ch)); xv[0] = qx.x; xv[1] = qx.y; xv[2] = qx.z; xv[3] = qx.w;
      if (relu && !gate_beta) { const float4 qy = __ldg(reinterpret_cast<const float4*>(y + row * ldy + ch)); yv[0] = qy.x; yv[1] = qy.y; yv[2] = qy.z; yv[3] = qy.w; }
      if (drop) { const float4 qd = __ldg(reinterpret_cast<const float4*>(drop + row * c + ch)); dv[0] = qd.x; dv[1] = qd.y; dv[2] = qd.z; dv[3] = qd.w; }
    } else {
      g[0] = dy[row * lddy + ch]; xv[0] = x[row * ldx + ch];
      if (relu && !gate_beta) yv[0] = y[row * ldy + ch];
      if (drop) dv[0] = drop[row * c + ch];
    }
    float o[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (relu && gate_beta) yv[e] = bn_affine(xv[e], mean[ch + e], rstd[ch + e], gamma[ch + e], gate_beta[ch + e]);
      if (relu && !(yv[e] > 0.f)) g[e] = 0.f;
      if (drop) g[e] *= dv[e] * drop_scale;
      const float xhat = (xv[e] - mean[ch + e]) * rstd[ch + e];
      const float mdb = (float)(fin[(ch + e) * 2] * inv_n), mdg = (float)(fin[(ch + e) * 2 + 1] * inv_n);
      o[e] = gamma[ch + e] * rstd[ch + e] * (g[e] - mdb - xhat * mdg);
    }
    if (V == 4) {
      if (g_out) *reinterpret_cast<float4*>(g_out + row * ldgo + ch) = make_float4(g[0], g[1], g[2], g[3]);
      *reinterpret_cast<float4*>(dx + row * lddx + ch) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      if (g_out) g_out[row * ldgo + ch] = g[0];
      dx[row * lddx + ch] = o[0];
    }
  }
}
// Same arithmetic with the thread -> channel mapping fixed for the whole kernel (block = rpb rows x c/4 channel groups,
// c/4 <= 256): the per-channel constants (mean, rstd, gamma, the two batch means of the reduction) live in registers and
// the row loop carries no index division, no double-precision multiplies and no per-element constant loads -- the
// flat-index kernel above is instruction-bound at 2.8 TB/s (profiles/r01_v11_launch_list_summary.txt).  Two rows per
// iteration keep 6-8 independent 16 B loads in flight per thread.
__global__ void __launch_bounds__(256) bn_bwd_dx_rows_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                             const float* __restrict__ y, int ldy, int64_t rows, int c,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, int relu, const float* __restrict__ drop,
                                                             float drop_scale, const double* __restrict__ fin, float* __restrict__ dx, int lddx,
                                                             float* __restrict__ g_out, int ldgo, int cg, int rpb,
                                                             const float* __restrict__ gate_beta) {
  const int lc = threadIdx.x % cg, lr = threadIdx.x / cg;          // blockDim.x == cg * rpb
  const int ch = lc * 4;
  const double inv_n = 1.0 / (double)rows;
  float mu[4], rs[4], ga[4], mdb[4], mdg[4], be[4];
  const bool regate = relu && gate_beta != nullptr;        // gate from x (bn_affine), y is not read
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    mu[e] = mean[ch + e]; rs[e] = rstd[ch + e]; ga[e] = gamma[ch + e]; be[e] = regate ? gate_beta[ch + e] : 0.f;
    mdb[e] = (float)(fin[(ch + e) * 2] * inv_n); mdg[e] = (float)(fin[(ch + e) * 2 + 1] * inv_n);
  }
  const int64_t step = (int64_t)gridDim.x * rpb;
  auto one_row = [&](int64_t row, const float4 q, const float4 qx, const float4 qy, const float4 qd) {
    float g[4] = {q.x, q.y, q.z, q.w};
    const float xv[4] = {qx.x, qx.y, qx.z, qx.w}, dv[4] = {qd.x, qd.y, qd.z, qd.w};
    float yv[4] = {qy.x, qy.y, qy.z, qy.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (regate) yv[e] = bn_affine(xv[e], mu[e], rs[e], ga[e], be[e]);
      if (relu && !(yv[e] > 0.f)) g[e] = 0.f;
      if (drop) g[e] *= dv[e] * drop_scale;
      const float xhat = (xv[e] - mu[e]) * rs[e];
      o[e] = ga[e] * rs[e] * (g[e] - mdb[e] - xhat * mdg[e]);
    }
    if (g_out) *reinterpret_cast<float4*>(g_out + row * ldgo + ch) = make_float4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<float4*>(dx + row * lddx + ch) = make_float4(o[0], o[1], o[2], o[3]);
  };
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t row = (int64_t)blockIdx.x * rpb + lr;
  for (; row + step < rows; row += 2 * step) {
    const int64_t r2 = row + step;
    const float4 qa = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + ch));
    const float4 qb = __ldg(reinterpret_cast<const float4*>(dy + r2 * lddy + ch));
    const float4 xa = __ldg(reinterpret_cast<const float4*>(x + row * ldx + ch));
    const float4 xb = __ldg(reinterpret_cast<const float4*>(x + r2 * ldx + ch));
    const float4 ya = (relu && !regate) ? __ldg(reinterpret_cast<const float4*>(y + row * ldy + ch)) : zero;
    const float4 yb = (relu && !regate) ? __ldg(reinterpret_cast<const float4*>(y + r2 * ldy + ch)) : zero;
    const float4 da = drop ? __ldg(reinterpret_cast<const float4*>(drop + row * c + ch)) : zero;
    const float4 db = drop ? __ldg(reinterpret_cast<const float4*>(drop + r2 * c + ch)) : zero;
    one_row(row, qa, xa, ya, da);
    one_row(r2, qb, xb, yb, db);
  }
  if (row < rows) {
    const float4 qa = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + ch));
    const float4 xa = __ldg(reinterpret_cast<const float4*>(x + row * ldx + ch));
    const float4 ya = (relu && !regate) ? __ldg(reinterpret_cast<const float4*>(y + row * ldy + ch)) : zero;
    const float4 da = drop ? __ldg(reinterpret_cast<const float4*>(drop + row * c + ch)) : zero;
    one_row(row, qa, xa, ya, da);
  }
}
__global__ void bn_param_out_kernel(const double* __restrict__ fin, int c, float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  if (dbeta) dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + (float)fin[ch * 2];
  if (dgamma) dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + (float)fin[ch * 2 + 1];
}
extern "C" int b2_bn_bwd(const float* dy, int lddy, const float* x, int ldx, const float* y, int ldy, int64_t rows, int c,
                         const float* mean, const float* rstd, const float* gamma, int relu, const float* dropmask, float drop_scale,
                         float* dx, int lddx, float* dgamma, float* dbeta, int accumulate_params, float* g_out, int ldgo,
                         const float* gate_beta, double* workspace, void* stream) {
  B2_REQUIRE(dy && x && dx && mean && rstd && gamma && workspace && rows > 0 && c > 0, "b2_bn_bwd: bad args");
  B2_REQUIRE(!relu || y || gate_beta, "b2_bn_bwd: relu gate needs y or gate_beta");
  if (!relu) gate_beta = nullptr;
  B2_REQUIRE(!dropmask || lddy == c, "b2_bn_bwd: dropout mask requires dense dy");
  RedArgs r{}; r.a = dy; r.lda = lddy; r.b = x; r.ldb = ldx; r.gate = (relu && !gate_beta) ? y : nullptr; r.ldg = ldy;
  r.gamma = gamma; r.gate_beta = gate_beta;
  r.drop = dropmask; r.drop_scale = drop_scale; r.mean = mean; r.rstd = rstd; r.rows = rows; r.c = c;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_col_reduce<2>(r, workspace, s); if (rc) return rc;
  double* fin = workspace + red_chunks(rows, c) * c * 2;
  launch_col_finalize(workspace, red_chunks(rows, c), c, fin, s);
  const bool vec = c % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0 && lddx % 4 == 0 && al16(dy) && al16(x) && al16(dx) &&
                   (!relu || gate_beta || (ldy % 4 == 0 && al16(y))) && (!dropmask || al16(dropmask)) && (!g_out || (ldgo % 4 == 0 && al16(g_out)));
  const int64_t total = rows * (vec ? c / 4 : c);
  int64_t blocks = ceil_div64(total, 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec && c / 4 <= 256) {
    const int cg = c / 4, rpb = 256 / cg;
    int64_t nb = ceil_div64(rows, rpb); if (nb > 148 * 16) nb = 148 * 16;
    bn_bwd_dx_rows_kernel<<<(unsigned)nb, cg * rpb, 0, s>>>(dy, lddy, x, ldx, y, ldy, rows, c, mean, rstd, gamma, relu, dropmask, drop_scale, fin, dx, lddx, g_out, ldgo, cg, rpb, gate_beta);
  } else if (vec && total < (1ll << 31)) bn_bwd_dx_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, s>>>(dy, lddy, x, ldx, y, ldy, rows, c, mean, rstd, gamma, relu, dropmask, drop_scale, fin, dx, lddx, g_out, ldgo, gate_beta);
  else if (vec) bn_bwd_dx_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, s>>>(dy, lddy, x, ldx, y, ldy, rows, c, mean, rstd, gamma, relu, dropmask, drop_scale, fin, dx, lddx, g_out, ldgo, gate_beta);
  else bn_bwd_dx_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, s>>>(dy, lddy, x, ldx, y, ldy, rows, c, mean, rstd, gamma, relu, dropmask, drop_scale, fin, dx, lddx, g_out, ldgo, gate_beta);
  bn_param_out_kernel<<<(c + 127) / 128, 128, 0, s>>>(fin, c, dgamma, dbeta, accumulate_params);
  B2_LAUNCH_CHECK("bn_bwd");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------ frozen BN
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ scale, float* __restrict__ shift, int c) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float s = gamma[ch] / sqrtf(var[ch] + eps);
  scale[ch] = s; shift[ch] = beta[ch] - mean[ch] * s;
}
extern "C" int b2_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                          float* shift, int c, void* stream) {
  B2_REQUIRE(gamma && beta && mean && var && scale && shift && c > 0, "b2_bn_fold: bad args");
  bn_fold_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, c);
  B2_LAUNCH_CHECK("bn_fold_kernel");
  return B2_OK;
}

// Frozen-statistics BN with trainable affine (torchvision backbone under freeze_batchnorm):
//   y = xhat*gamma + beta  =>  dbeta = sum g, dgamma = sum g*xhat with xhat = (y - beta)/gamma recovered
//   from the stored BN output proxy `ybn` (only needed where the gate is open).  g = dy gated by gate>0.
__global__ void bn_eval_param_out_kernel(const double* __restrict__ fin, int c, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double sg = fin[ch * 2], sgy = fin[ch * 2 + 1];
  const double ga = gamma[ch], be = beta[ch];
  const double dg = ga != 0.0 ? (sgy - be * sg) / ga : 0.0;
  dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + (float)sg;
  dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + (float)dg;
}
extern "C" int b2_bn_eval_param_grad(const float* dy, int lddy, const float* ybn, int ldy, int64_t rows, int c, const float* gamma,
                                     const float* beta, const float* gate, int ldg, const float* sub, int lds, float* dgamma,
                                     float* dbeta, int accumulate, double* workspace, void* stream) {
  B2_REQUIRE(dy && ybn && gamma && beta && dgamma && dbeta && workspace && rows > 0 && c > 0, "b2_bn_eval_param_grad: bad args");
  RedArgs r{}; r.a = dy; r.lda = lddy; r.b = ybn; r.ldb = ldy; r.gate = gate; r.ldg = ldg; r.sub = sub; r.lds = lds; r.rows = rows; r.c = c;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_col_reduce<3>(r, workspace, s); if (rc) return rc;
  double* fin = workspace + red_chunks(rows, c) * c * 2;
  launch_col_finalize(workspace, red_chunks(rows, c), c, fin, s);
  bn_eval_param_out_kernel<<<(c + 127) / 128, 128, 0, s>>>(fin, c, gamma, beta, dgamma, dbeta, accumulate);
  B2_LAUNCH_CHECK("bn_eval_param_grad");
  return B2_OK;
}

// Reduction of the partial sums written by the conv epilogue (b2_conv_params.stats), two fixed-order stages:
// stage 1: grid (c/32, STAT_SPLITS); block = 32 channels x 8 row lanes over one contiguous slice of the row blocks
//          -> workspace[split][c][2] (double);  stage 2: thread per channel sums the splits and writes the gradients.
constexpr int STAT_SPLITS = 16;
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(const float* __restrict__ stats, int64_t rows, int ld, int c,
                                                              double* __restrict__ ws) {
  __shared__ double sm[8][32][2];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  const int64_t per = (rows + STAT_SPLITS - 1) / STAT_SPLITS;
  const int64_t k0 = blockIdx.y * per;
  int64_t k1 = k0 + per; if (k1 > rows) k1 = rows;
  double s0 = 0, s1 = 0;
  if (ch < c) {
#pragma unroll 4
    for (int64_t k = k0 + ry; k < k1; k += 8) {
      s0 += (double)__ldg(stats + (k * 2) * ld + ch);
      s1 += (double)__ldg(stats + (k * 2 + 1) * ld + ch);
    }
  }
  sm[ry][cx][0] = s0; sm[ry][cx][1] = s1;
  __syncthreads();
  if (ry == 0 && ch < c) {
    double t0 = 0, t1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t0 += sm[k][cx][0]; t1 += sm[k][cx][1]; }
    ws[((int64_t)blockIdx.y * c + ch) * 2] = t0;
    ws[((int64_t)blockIdx.y * c + ch) * 2 + 1] = t1;
  }
}
__global__ void bn_eval_param_from_stats_out_kernel(const double* __restrict__ ws, int c, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ dgamma,
                                                    float* __restrict__ dbeta, int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double sg = 0, sgy = 0;
#pragma unroll
  for (int k = 0; k < STAT_SPLITS; ++k) { sg += ws[((int64_t)k * c + ch) * 2]; sgy += ws[((int64_t)k * c + ch) * 2 + 1]; }
  const double ga = gamma[ch], be = beta[ch];
  const double dg = ga != 0.0 ? (sgy - be * sg) / ga : 0.0;
  dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + (float)sg;
  dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + (float)dg;
}
extern "C" int64_t b2_bn_stats_workspace_doubles(int c) { return (int64_t)64 * 2 * c; }   // max(STAT_SPLITS, WDOT_SPLITS)
extern "C" int b2_bn_eval_param_grad_from_stats(const float* stats, int64_t stat_rows, int ld_stats, int c, const float* gamma,
                                                const float* beta, float* dgamma, float* dbeta, int accumulate,
                                                double* workspace, void* stream) {
  B2_REQUIRE(stats && gamma && beta && dgamma && dbeta && workspace && stat_rows > 0 && c > 0 && ld_stats >= c,
             "b2_bn_eval_param_grad_from_stats: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  bn_stats_partial_kernel<<<dim3((c + 31) / 32, STAT_SPLITS), 256, 0, s>>>(stats, stat_rows, ld_stats, c, workspace);
  bn_eval_param_from_stats_out_kernel<<<(c + 127) / 128, 128, 0, s>>>(workspace, c, gamma, beta, dgamma, dbeta, accumulate);
  B2_LAUNCH_CHECK("bn_eval_param_grad_from_stats");
  return B2_OK;
}

// Frozen-BN parameter gradients from the WEIGHT gradient (no pass over activations at all):
//   y = scale*conv(x, W) + shift, scale = gamma*invstd  =>  dgamma = invstd * (sum_pix g*conv - mean * sum_pix g)
//   and  sum_pix g[pix,c]*conv[pix,c] = <W[c,:], dWraw[c,:]>  with dWraw the un-scaled weight gradient.  The wgrad
//   epilogue stores gw = scale*dWraw, hence  dgamma = <W[c], gw[c]>/gamma - invstd*mean*dbeta,  dbeta = sum_pix g.
// Because gw accumulates over backward passes, dgamma is SET from the accumulated totals (gw, dbeta); this is the
// accumulated value as long as W.grad, gamma.grad and beta.grad were zeroed together (they always are: zero_grad).
// ws[k][c][2] (double): n_splits partial column sums of g (entry 0 used).  One block per channel.
__global__ void __launch_bounds__(128) bn_wdot_out_kernel(const double* __restrict__ ws, int n_splits, int c,
                                                          const float* __restrict__ w, const float* __restrict__ gw,
                                                          int64_t row_len, const float* __restrict__ gamma,
                                                          const float* __restrict__ mean, const float* __restrict__ var,
                                                          float eps, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                          int accumulate, int vec) {
  __shared__ double sm[4];
  const int ch = blockIdx.x;
  const float* wr = w + (int64_t)ch * row_len;
  const float* gr = gw + (int64_t)ch * row_len;
  double acc = 0.0;
  if (vec) {
    const int64_t n4 = row_len >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(wr) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(gr) + i);
      acc += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
    }
  } else {
    for (int64_t i = threadIdx.x; i < row_len; i += 128) acc += (double)__ldg(wr + i) * (double)__ldg(gr + i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double dot = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    double sg = 0.0;
    for (int k = 0; k < n_splits; ++k) sg += ws[((int64_t)k * c + ch) * 2];
    const float db = (accumulate ? dbeta[ch] : 0.f) + (float)sg;
    dbeta[ch] = db;
    const double ga = gamma[ch];
    const double invstd = 1.0 / sqrt((double)var[ch] + (double)eps);
    dgamma[ch] = (float)((ga != 0.0 ? dot / ga : 0.0) - invstd * (double)mean[ch] * (double)db);
  }
}
// Column sums of entry j = 0 of the epilogue statistics (sum_pix g), WDOT_SPLITS contiguous slices of the row blocks:
// block = 32 channels x 8 row lanes, every lane issues all its loads before the first add (<= 8 rows per lane per split at
// the sizes of the hot path), fixed summation order.
constexpr int WDOT_SPLITS = 64;
__global__ void __launch_bounds__(256) stats_colsum_partial_kernel(const float* __restrict__ stats, int64_t rows, int ld, int c,
                                                                  double* __restrict__ ws) {
  __shared__ double sm[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx;
  const int64_t per = (rows + WDOT_SPLITS - 1) / WDOT_SPLITS;
  const int64_t k0 = blockIdx.y * per;
  int64_t k1 = k0 + per; if (k1 > rows) k1 = rows;
  double s0 = 0;
  if (ch < c) {
    int64_t k = k0 + ry;
    for (; k + 24 < k1; k += 32) {          // 4 independent loads in flight
      const float v0 = __ldg(stats + (k * 2) * ld + ch), v1 = __ldg(stats + ((k + 8) * 2) * ld + ch);
      const float v2 = __ldg(stats + ((k + 16) * 2) * ld + ch), v3 = __ldg(stats + ((k + 24) * 2) * ld + ch);
      s0 += (double)v0; s0 += (double)v1; s0 += (double)v2; s0 += (double)v3;
    }
    for (; k < k1; k += 8) s0 += (double)__ldg(stats + (k * 2) * ld + ch);
  }
  sm[ry][cx] = s0;
  __syncthreads();
  if (ry == 0 && ch < c) {
    double t0 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t0 += sm[k][cx];
    ws[((int64_t)blockIdx.y * c + ch) * 2] = t0;
    ws[((int64_t)blockIdx.y * c + ch) * 2 + 1] = 0.0;
  }
}
static int launch_wdot(const double* ws, int n_splits, int c, const float* w, const float* gw, int64_t row_len,
                       const float* gamma, const float* mean, const float* var, float eps, float* dgamma, float* dbeta,
                       int accumulate, cudaStream_t s) {
  const int vec = (row_len % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0) && ((reinterpret_cast<uintptr_t>(gw) & 15) == 0);
  bn_wdot_out_kernel<<<c, 128, 0, s>>>(ws, n_splits, c, w, gw, row_len, gamma, mean, var, eps, dgamma, dbeta, accumulate, vec);
  B2_LAUNCH_CHECK("bn_wdot_out_kernel");
  return B2_OK;
}
extern "C" int b2_bn_eval_param_grad_wdot_from_stats(const float* stats, int64_t stat_rows, int ld_stats, int c, const float* w,
                                                     const float* gw, int64_t row_len, const float* gamma, const float* mean,
                                                     const float* var, float eps, float* dgamma, float* dbeta, int accumulate,
                                                     double* workspace, void* stream) {
  B2_REQUIRE(stats && w && gw && gamma && mean && var && dgamma && dbeta && workspace && stat_rows > 0 && c > 0 &&
             ld_stats >= c && row_len > 0, "b2_bn_eval_param_grad_wdot_from_stats: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  stats_colsum_partial_kernel<<<dim3((c + 31) / 32, WDOT_SPLITS), 256, 0, s>>>(stats, stat_rows, ld_stats, c, workspace);
  B2_LAUNCH_CHECK("stats_colsum_partial_kernel");
  return launch_wdot(workspace, WDOT_SPLITS, c, w, gw, row_len, gamma, mean, var, eps, dgamma, dbeta, accumulate, s);
}
extern "C" int b2_bn_eval_param_grad_wdot(const float* dy, int lddy, int64_t rows, int c, const float* w, const float* gw,
                                          int64_t row_len, const float* gamma, const float* mean, const float* var, float eps,
                                          float* dgamma, float* dbeta, int accumulate, double* workspace, void* stream) {
  B2_REQUIRE(dy && w && gw && gamma && mean && var && dgamma && dbeta && workspace && rows > 0 && c > 0 && lddy >= c &&
             row_len > 0, "b2_bn_eval_param_grad_wdot: bad args");
  RedArgs r{}; r.a = dy; r.lda = lddy; r.rows = rows; r.c = c;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_col_reduce<0>(r, workspace, s); if (rc) return rc;
  double* fin = workspace + red_chunks(rows, c) * c * 2;
  launch_col_finalize(workspace, red_chunks(rows, c), c, fin, s);
  B2_LAUNCH_CHECK("col_finalize_kernel");
  return launch_wdot(fin, 1, c, w, gw, row_len, gamma, mean, var, eps, dgamma, dbeta, accumulate, s);
}

// ------------------------------------------------------------------------------------------ GAP / broadcast
__global__ void gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int hw, int c, int ldx, float mul) {
  __shared__ float sm[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cx; const int n = blockIdx.y;
  float s = 0.f;
  if (ch < c) for (int p = ry; p < hw; p += 8) s += __ldg(x + ((int64_t)n * hw + p) * ldx + ch);
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    y[(int64_t)n * c + ch] = t * mul;
  }
}
extern "C" int b2_gap_fwd(const float* x, float* y, int n, int hw, int c, int ldx, void* stream) {
  B2_REQUIRE(x && y && n > 0 && hw > 0 && c > 0 && ldx >= c, "b2_gap_fwd: bad args");
  gap_fwd_kernel<<<dim3((c + 31) / 32, n), 256, 0, (cudaStream_t)stream>>>(x, y, hw, c, ldx, 1.0f / (float)hw);
  B2_LAUNCH_CHECK("gap_fwd_kernel");
  return B2_OK;
}
template <int V>
__global__ void __launch_bounds__(256) bcast_kernel(const float* __restrict__ v, float* __restrict__ y, int hw, int c, int ldy, float mul,
                                                    int accumulate) {
  const int n = blockIdx.y;
  const int cg = c / V;
  const int64_t total = (int64_t)hw * cg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; const int64_t p = i / cg;
    float* d = y + ((int64_t)n * hw + p) * ldy + ch;
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(v + (int64_t)n * c + ch));
      float4 o = make_float4(q.x * mul, q.y * mul, q.z * mul, q.w * mul);
      if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(d); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
      *reinterpret_cast<float4*>(d) = o;
    } else {
      const float val = v[(int64_t)n * c + ch] * mul;
      *d = accumulate ? *d + val : val;
    }
  }
}
static int launch_bcast(const float* v, float* y, int n, int hw, int c, int ldy, float mul, int accumulate, cudaStream_t s) {
  const bool vec = c % 4 == 0 && ldy % 4 == 0 && al16(v) && al16(y);
  int bx = (int)(((int64_t)hw * (vec ? c / 4 : c) + 255) / 256); if (bx > 2048) bx = 2048;
  if (vec) bcast_kernel<4><<<dim3(bx, n), 256, 0, s>>>(v, y, hw, c, ldy, mul, accumulate);
  else bcast_kernel<1><<<dim3(bx, n), 256, 0, s>>>(v, y, hw, c, ldy, mul, accumulate);
  return B2_OK;
}
extern "C" int b2_gap_bwd(const float* dy, float* dx, int n, int hw, int c, int ldx, int accumulate, void* stream) {
  B2_REQUIRE(dy && dx && n > 0 && hw > 0 && c > 0, "b2_gap_bwd: bad args");
  launch_bcast(dy, dx, n, hw, c, ldx, 1.0f / (float)hw, accumulate, (cudaStream_t)stream);
  B2_LAUNCH_CHECK("gap_bwd");
  return B2_OK;
}
extern "C" int b2_bcast_fwd(const float* v, float* y, int n, int hw, int c, int ldy, void* stream) {
  B2_REQUIRE(v && y && n > 0 && hw > 0 && c > 0, "b2_bcast_fwd: bad args");
  launch_bcast(v, y, n, hw, c, ldy, 1.0f, 0, (cudaStream_t)stream);
  B2_LAUNCH_CHECK("bcast_fwd");
  return B2_OK;
}
extern "C" int b2_bcast_bwd(const float* dy, float* dv, int n, int hw, int c, int ldy, void* stream) {
  B2_REQUIRE(dy && dv && n > 0 && hw > 0 && c > 0, "b2_bcast_bwd: bad args");
  gap_fwd_kernel<<<dim3((c + 31) / 32, n), 256, 0, (cudaStream_t)stream>>>(dy, dv, hw, c, ldy, 1.0f);
  B2_LAUNCH_CHECK("bcast_bwd");
  return B2_OK;
}
