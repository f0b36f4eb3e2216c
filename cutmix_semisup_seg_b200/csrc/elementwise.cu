// HBM-bound elementwise kernels of the CutMix mean-teacher hot path:
//   EMA teacher update        (optim_weight_ema.py:21-25)
//   box mask rasterisation    (mask_gen.py:110-117)
//   image / valid-mask mix    (train_seg_semisup_mask_mt.py:350-351, 389)
// plus small utilities (scale, add, fill, tf32 split, weight transpose, dropout mask).
#include "common.cuh"

thread_local char g_b2_err[512] = "";

extern "C" const char* b2_last_error(void) { return g_b2_err; }
extern "C" int b2_version(void) { return 3; }

int b2_sm_count_cached() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    sms = v;
  }
  return sms;
}
extern "C" int b2_num_sms(void) { return b2_sm_count_cached(); }

// ------------------------------------------------------------------------------------------
// EMA.  Arithmetic contract (verified against the reference):
//   t <- fl( fl(t * a) + fl(s * (1-a)) ), no FMA contraction.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ema1(float t, float s, float a, float oma) {
  return __fadd_rn(__fmul_rn(t, a), __fmul_rn(s, oma));
}

constexpr int EMA_THREADS = 256;
constexpr int EMA_CHUNK = 16384;  // elements per chunk (64 KB in, 64 KB out per tensor)

__device__ __forceinline__ void ema_range(float* __restrict__ t, const float* __restrict__ s,
                                          int64_t count, float a, float oma) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
  if (aligned) {
    const int64_t n4 = count >> 2;
    float4* t4 = reinterpret_cast<float4*>(t);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    // 4 independent 16B loads per thread per tensor in flight
    int64_t i = threadIdx.x;
    for (; i + 3 * EMA_THREADS < n4; i += 4 * EMA_THREADS) {
      float4 tv[4], sv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { tv[j] = t4[i + j * EMA_THREADS]; sv[j] = __ldg(&s4[i + j * EMA_THREADS]); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tv[j].x = ema1(tv[j].x, sv[j].x, a, oma); tv[j].y = ema1(tv[j].y, sv[j].y, a, oma);
        tv[j].z = ema1(tv[j].z, sv[j].z, a, oma); tv[j].w = ema1(tv[j].w, sv[j].w, a, oma);
        t4[i + j * EMA_THREADS] = tv[j];
      }
    }
    for (; i < n4; i += EMA_THREADS) {
      float4 tv = t4[i], sv = __ldg(&s4[i]);
      tv.x = ema1(tv.x, sv.x, a, oma); tv.y = ema1(tv.y, sv.y, a, oma);
      tv.z = ema1(tv.z, sv.z, a, oma); tv.w = ema1(tv.w, sv.w, a, oma);
      t4[i] = tv;
    }
    for (int64_t k = (n4 << 2) + threadIdx.x; k < count; k += EMA_THREADS) t[k] = ema1(t[k], s[k], a, oma);
  } else {
    for (int64_t k = threadIdx.x; k < count; k += EMA_THREADS) t[k] = ema1(t[k], s[k], a, oma);
  }
}

__global__ void __launch_bounds__(EMA_THREADS) ema_table_kernel(const b2_ema_chunk* __restrict__ table,
                                                                float a, float oma) {
  const b2_ema_chunk c = table[blockIdx.x];
  ema_range(c.tgt, c.src, c.count, a, oma);
}

__global__ void __launch_bounds__(EMA_THREADS) ema_flat_kernel(float* __restrict__ t,
                                                               const float* __restrict__ s,
                                                               int64_t count, float a, float oma) {
  const int64_t start = (int64_t)blockIdx.x * EMA_CHUNK;
  int64_t n = count - start;
  if (n > EMA_CHUNK) n = EMA_CHUNK;
  ema_range(t + start, s + start, n, a, oma);
}

extern "C" int b2_ema_step(const b2_ema_chunk* table, int64_t n_chunks, float alpha,
                           float one_minus_alpha, void* stream) {
  if (n_chunks == 0) return B2_OK;
  B2_REQUIRE(table != nullptr && n_chunks > 0 && n_chunks < (1ll << 31), "b2_ema_step: bad table");
  ema_table_kernel<<<(unsigned)n_chunks, EMA_THREADS, 0, (cudaStream_t)stream>>>(table, alpha, one_minus_alpha);
  B2_LAUNCH_CHECK("ema_table_kernel");
  return B2_OK;
}

extern "C" int b2_ema_step_flat(float* tgt, const float* src, int64_t count, float alpha,
                                float one_minus_alpha, void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(tgt && src && count > 0, "b2_ema_step_flat: bad args");
  const int64_t blocks = ceil_div64(count, EMA_CHUNK);
  B2_REQUIRE(blocks < (1ll << 31), "b2_ema_step_flat: too large");
  ema_flat_kernel<<<(unsigned)blocks, EMA_THREADS, 0, (cudaStream_t)stream>>>(tgt, src, count, alpha, one_minus_alpha);
  B2_LAUNCH_CHECK("ema_flat_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// Box mask rasterisation.  Each rectangle toggles (x -> 1-x) the pixels of its half-open range.
// The toggle count parity decides the final value: exact for values in {0,1}.
// ------------------------------------------------------------------------------------------
__global__ void box_mask_kernel(const int32_t* __restrict__ boxes, int n_boxes, int h, int w,
                                float init, float* __restrict__ out) {
  const int img = blockIdx.y;
  const int64_t hw = (int64_t)h * w;
  const int32_t* bx = boxes + (int64_t)img * n_boxes * 4;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
    float v = init;
    for (int b = 0; b < n_boxes; ++b) {
      const int y0 = bx[b * 4 + 0], y1 = bx[b * 4 + 1], x0 = bx[b * 4 + 2], x1 = bx[b * 4 + 3];
      if (y >= y0 && y < y1 && x >= x0 && x < x1) v = 1.0f - v;
    }
    out[(int64_t)img * hw + p] = v;
  }
}

extern "C" int b2_box_mask_rasterize(const int32_t* boxes, int n, int n_boxes, int h, int w,
                                     float init, float* out, void* stream) {
  B2_REQUIRE(boxes && out && n > 0 && n_boxes >= 0 && h > 0 && w > 0, "b2_box_mask_rasterize: bad args");
  const int64_t hw = (int64_t)h * w;
  int bx = (int)((hw + 255) / 256);
  if (bx > 1024) bx = 1024;
  dim3 grid(bx, n);
  box_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, n_boxes, h, w, init, out);
  B2_LAUNCH_CHECK("box_mask_kernel");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// Mix.  out = fl(fl(a*fl(1-m)) + fl(b*m))   (reference line 350: separate roundings, no select,
// so -0.0 / inf / NaN propagate exactly as in PyTorch).   Cut: out = fl(a*m).
// grid.y = n*c planes; mask plane index = plane / c.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mix1(float a, float b, float m) {
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, m)), __fmul_rn(b, m));
}

// PER_SAMPLE: `m` holds one factor per image (ICT mix factors, train_seg_semisup_ict.py:306-311) instead of a mask plane.
template <bool CUT, bool PER_SAMPLE>
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  const float* __restrict__ m, float* __restrict__ out,
                                                  int c, int64_t hw) {
  const int plane = blockIdx.y;
  const int img = plane / c;
  const float* ap = a + (int64_t)plane * hw;
  const float* bp = CUT ? nullptr : b + (int64_t)plane * hw;
  const float* mp = PER_SAMPLE ? m : m + (int64_t)img * hw;
  const float ms = PER_SAMPLE ? __ldg(m + img) : 0.0f;
  float* op = out + (int64_t)plane * hw;
  const bool vec = (hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | (PER_SAMPLE ? 0 : reinterpret_cast<uintptr_t>(m)) |
                                      reinterpret_cast<uintptr_t>(out) | (CUT ? 0 : reinterpret_cast<uintptr_t>(b))) & 15) == 0;
  if (vec) {
    const int64_t n4 = hw >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(ap) + i);
      const float4 mv = PER_SAMPLE ? make_float4(ms, ms, ms, ms) : __ldg(reinterpret_cast<const float4*>(mp) + i);
      float4 r;
      if (CUT) {
        r.x = __fmul_rn(av.x, mv.x); r.y = __fmul_rn(av.y, mv.y); r.z = __fmul_rn(av.z, mv.z); r.w = __fmul_rn(av.w, mv.w);
      } else {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bp) + i);
        r.x = mix1(av.x, bv.x, mv.x); r.y = mix1(av.y, bv.y, mv.y); r.z = mix1(av.z, bv.z, mv.z); r.w = mix1(av.w, bv.w, mv.w);
      }
      reinterpret_cast<float4*>(op)[i] = r;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
      const float mi = PER_SAMPLE ? ms : mp[i];
      op[i] = CUT ? __fmul_rn(ap[i], mi) : mix1(ap[i], bp[i], mi);
    }
  }
}

extern "C" int b2_mix(const float* a, const float* b, const float* m, float* out, int n, int c,
                      int64_t hw, void* stream) {
  B2_REQUIRE(a && m && out && n > 0 && c > 0 && hw > 0, "b2_mix: bad args");
  B2_REQUIRE((int64_t)n * c <= 65535, "b2_mix: n*c too large");
  int bx = (int)((hw / 4 + 255) / 256);
  if (bx < 1) bx = 1;
  if (bx > 2048) bx = 2048;
  dim3 grid(bx, n * c);
  if (b) mix_kernel<false, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, m, out, c, hw);
  else   mix_kernel<true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, nullptr, m, out, c, hw);
  B2_LAUNCH_CHECK("mix_kernel");
  return B2_OK;
}

extern "C" int b2_mix_per_sample(const float* a, const float* b, const float* factors, float* out, int n, int c,
                                 int64_t hw, void* stream) {
  B2_REQUIRE(a && b && factors && out && n > 0 && c > 0 && hw > 0, "b2_mix_per_sample: bad args");
  B2_REQUIRE((int64_t)n * c <= 65535, "b2_mix_per_sample: n*c too large");
  int bx = (int)((hw / 4 + 255) / 256);
  if (bx < 1) bx = 1;
  if (bx > 2048) bx = 2048;
  dim3 grid(bx, n * c);
  mix_kernel<false, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, factors, out, c, hw);
  B2_LAUNCH_CHECK("mix_kernel<per sample>");
  return B2_OK;
}

// ------------------------------------------------------------------------------------------
// Utilities
// ------------------------------------------------------------------------------------------
__global__ void scale_kernel(float* __restrict__ x, int64_t count, const float* __restrict__ scale_dev, float scale_host) {
  const float s = (scale_dev ? scale_dev[0] : 1.0f) * scale_host;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) x[i] *= s;
}
extern "C" int b2_scale_inplace(float* x, int64_t count, const float* scale_dev, float scale_host, void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(x && count > 0, "b2_scale_inplace: bad args");
  int64_t blocks = ceil_div64(count, 256 * 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, count, scale_dev, scale_host);
  B2_LAUNCH_CHECK("scale_kernel");
  return B2_OK;
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] += src[i];
}
extern "C" int b2_add_inplace(float* dst, const float* src, int64_t count, void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(dst && src && count > 0, "b2_add_inplace: bad args");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, count);
  B2_LAUNCH_CHECK("add_kernel");
  return B2_OK;
}

__global__ void fill_kernel(float* __restrict__ dst, float v, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = v;
}
extern "C" int b2_fill(float* dst, float value, int64_t count, void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(dst && count > 0, "b2_fill: bad args");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, value, count);
  B2_LAUNCH_CHECK("fill_kernel");
  return B2_OK;
}

// TF32 split: hi keeps sign, exponent and the 10 top mantissa bits; lo = x - hi is exact in fp32.
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = __fsub_rn(v, h);
  }
}
extern "C" int b2_split_tf32(const float* x, float* hi, float* lo, int64_t count, void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(x && hi && lo && count > 0, "b2_split_tf32: bad args");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, count);
  B2_LAUNCH_CHECK("split_tf32_kernel");
  return B2_OK;
}

// (A, T, B) -> (B, T, A): 32x32 smem tile transpose per tap.
__global__ void transpose_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int T, int B, int ldd,
                                   const float* __restrict__ scale_a) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int a = a0 + i, b = b0 + threadIdx.x;
    tile[i][threadIdx.x] = (a < A && b < B) ? src[((int64_t)a * T + t) * B + b] * (scale_a ? scale_a[a] : 1.0f) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, a = a0 + threadIdx.x;
    if (a < ldd && b < B) dst[((int64_t)b * T + t) * ldd + a] = a < A ? tile[threadIdx.x][i] : 0.0f;
  }
}
extern "C" int b2_transpose_w(const float* src, float* dst, int a, int t, int b, int ldd, const float* scale_a, void* stream) {
  B2_REQUIRE(src && dst && a > 0 && t > 0 && b > 0 && ldd >= a, "b2_transpose_w: bad args");
  dim3 grid((b + 31) / 32, (ldd + 31) / 32, t), block(32, 8);
  B2_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "b2_transpose_w: too large");
  transpose_w_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, a, t, b, ldd, scale_a);
  B2_LAUNCH_CHECK("transpose_w_kernel");
  return B2_OK;
}

// Counter-based dropout mask (splitmix64 hash of (seed, index)); keep with prob 1-p.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void dropout_mask_kernel(float* __restrict__ mask, int64_t count, float p, uint64_t seed, uint64_t offset,
                                    const uint64_t* __restrict__ offset_dev) {
  if (offset_dev) offset += offset_dev[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const uint64_t r = splitmix64(seed ^ splitmix64(offset + (uint64_t)i));
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);  // [0,1)
    mask[i] = u >= p ? 1.0f : 0.0f;
  }
}
extern "C" int b2_dropout_mask(float* mask, int64_t count, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                               void* stream) {
  if (count == 0) return B2_OK;
  B2_REQUIRE(mask && count > 0 && p >= 0.f && p < 1.f, "b2_dropout_mask: bad args");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, count, p, seed, offset, offset_dev);
  B2_LAUNCH_CHECK("dropout_mask_kernel");
  return B2_OK;
}

// g[row, c] = (y[row, c] > 0) ? g[row, c] : 0   (explicit ReLU backward where it cannot be fused)
static inline bool ew_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
template <int V, typename I>
__global__ void __launch_bounds__(256) relu_gate_kernel(float* __restrict__ g, int ldg, const float* __restrict__ y, int ldy, int64_t rows, int c) {
  const int cg = c / V;
  const I total = (I)(rows * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; const int64_t row = (int64_t)(i / cg);
    if (V == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(y + row * ldy + ch));
      float4* gp = reinterpret_cast<float4*>(g + row * ldg + ch);
      float4 v = *gp;
      if (!(q.x > 0.0f)) v.x = 0.0f; if (!(q.y > 0.0f)) v.y = 0.0f; if (!(q.z > 0.0f)) v.z = 0.0f; if (!(q.w > 0.0f)) v.w = 0.0f;
      *gp = v;
    } else {
      if (!(y[row * ldy + ch] > 0.0f)) g[row * ldg + ch] = 0.0f;
    }
  }
}
extern "C" int b2_relu_gate(float* g, int ldg, const float* y, int ldy, int64_t rows, int c, void* stream) {
  B2_REQUIRE(g && y && rows > 0 && c > 0 && ldg >= c && ldy >= c, "b2_relu_gate: bad args");
  const bool vec = c % 4 == 0 && ldg % 4 == 0 && ldy % 4 == 0 && ew_al16(g) && ew_al16(y);
  int64_t blocks = ceil_div64(rows * (vec ? c / 4 : c), 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec && rows * (c / 4) < (1ll << 31)) relu_gate_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, y, ldy, rows, c);
  else if (vec) relu_gate_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, y, ldy, rows, c);
  else relu_gate_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, y, ldy, rows, c);
  B2_LAUNCH_CHECK("relu_gate_kernel");
  return B2_OK;
}
// strided copy / add of an NHWC slice: dst[row, c] (+)= src[row, c]
template <int V, typename I>
__global__ void __launch_bounds__(256) slice_copy_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int64_t rows, int c, int accumulate) {
  const int cg = c / V;
  const I total = (I)(rows * cg);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cg) * V; const int64_t row = (int64_t)(i / cg);
    if (V == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src + row * lds + ch));
      float4* d = reinterpret_cast<float4*>(dst + row * ldd + ch);
      if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *d = v;
    } else {
      const float v = src[row * lds + ch];
      float* d = dst + row * ldd + ch;
      *d = accumulate ? *d + v : v;
    }
  }
}
extern "C" int b2_slice_copy(float* dst, int ldd, const float* src, int lds, int64_t rows, int c, int accumulate, void* stream) {
  B2_REQUIRE(dst && src && rows > 0 && c > 0 && ldd >= c && lds >= c, "b2_slice_copy: bad args");
  const bool vec = c % 4 == 0 && ldd % 4 == 0 && lds % 4 == 0 && ew_al16(dst) && ew_al16(src);
  int64_t blocks = ceil_div64(rows * (vec ? c / 4 : c), 256); if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec && rows * (c / 4) < (1ll << 31)) slice_copy_kernel<4, int32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, ldd, src, lds, rows, c, accumulate);
  else if (vec) slice_copy_kernel<4, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, ldd, src, lds, rows, c, accumulate);
  else slice_copy_kernel<1, int64_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, ldd, src, lds, rows, c, accumulate);
  B2_LAUNCH_CHECK("slice_copy_kernel");
  return B2_OK;
}
