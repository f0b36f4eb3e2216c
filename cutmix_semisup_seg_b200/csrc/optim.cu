// Fused multi-tensor optimiser step of the student network + EMA step of the teacher network:
//   O1  student_optim.step()   train_seg_semisup_mask_mt.py:465 (Adam :91-93, SGD :95-98 on the reference's groups)
//   E1  teacher_optim.step()   train_seg_semisup_mask_mt.py:466-467 -> optim_weight_ema.py:21-25
// ONE launch updates every parameter (Adam or SGD-momentum), and -- in the same pass, from the value just written --
// the teacher's EMA copy; BatchNorm buffers (EMA only, no optimiser state) are chunks with k == 0.  HBM-bound:
// Adam reads p, g, m, v, t and writes p, m, v, t = 36 B per element, against 28 B (torch multi-tensor Adam) + 12 B (EMA).
//
// Reference quirk reproduced here: DeepLab v2's "pretrained" group lists a tensor once per enclosing module
// (deeplab2.py:224-230: 314 entries, 104 unique).  torch.optim's per-tensor loop therefore applies k SEQUENTIAL updates
// per step() to such a tensor, with shared exp_avg / exp_avg_sq and its step counter advancing by k;
// torch's multi-tensor (foreach / fused) kernels do not.  Each chunk carries its multiplicity k and the kernel runs the
// k updates in registers.
#include "common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

struct StepConsts {
  float step_size_neg[B2_OPT_MAX_K];    // -(lr / (1 - beta1^t))           per sequential update
  float bc2_sqrt[B2_OPT_MAX_K];         // sqrt(1 - beta2^t)
};

// Arithmetic follows torch.optim.adam._single_tensor_adam (fp32 tensors, python-double scalars rounded to fp32 at the
// kernel boundary), operation by operation and without FMA contraction:
//   exp_avg.lerp_(g, 1 - beta1)                         -> m + w1 * (g - m)            (|w1| < 0.5 branch of lerp)
//   exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1-b2)   -> fl(v * b2) + fl(fl(w2 * g) * g)
//   denom = (exp_avg_sq.sqrt() / bc2_sqrt).add_(eps)
//   param.addcdiv_(exp_avg, denom, value=-step_size)    -> p + fl(step_size_neg * fl(m / denom))
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float w1, float b2, float w2, float eps,
                                      const StepConsts& sc, int k) {
#pragma unroll 1
  for (int j = 0; j < k; ++j) {
    m = __fadd_rn(m, __fmul_rn(w1, __fsub_rn(g, m)));
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(w2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), sc.bc2_sqrt[j]), eps);
    p = __fadd_rn(p, __fmul_rn(sc.step_size_neg[j], __fdiv_rn(m, denom)));
  }
}

// torch.optim.sgd._single_tensor_sgd: d = g + wd * p; buf = first ? d : buf * mu + d (dampening 0);
// d = nesterov ? d + mu * buf : buf; p = p - lr * d.  On the very first step torch (2.x) collects the momentum buffers
// of all list entries BEFORE its loop (all None), so every one of the k visits of a duplicated tensor starts a fresh
// buffer from its own d; from the second step on the visits share one buffer.  (torch 1.4 looked the state up inside the
// loop; the oracle this build is pinned against is the reference run under the torch of this image.)
__device__ __forceinline__ void sgd1(float& p, float g, float& buf, float lr, float mu, float wd, int nesterov,
                                     bool first, int k) {
#pragma unroll 1
  for (int j = 0; j < k; ++j) {
    float d = wd != 0.0f ? __fadd_rn(g, __fmul_rn(wd, p)) : g;
    if (mu != 0.0f) {
      buf = first ? d : __fadd_rn(__fmul_rn(buf, mu), d);
      d = nesterov ? __fadd_rn(d, __fmul_rn(mu, buf)) : buf;
    }
    p = __fadd_rn(p, __fmul_rn(-lr, d));
  }
}

__device__ __forceinline__ float ema1(float t, float s, float a, float oma) {
  return __fadd_rn(__fmul_rn(t, a), __fmul_rn(s, oma));          // optim_weight_ema.py:22-25, three roundings
}

template <int ALGO>     // 0 = Adam, 1 = SGD
__global__ void __launch_bounds__(OPT_THREADS)
opt_ema_kernel(const b2_opt_chunk* __restrict__ table, const double* __restrict__ lr_groups,
               const int64_t* __restrict__ iter_dev, double beta1, double beta2, float eps, float sgd_mu, float sgd_wd,
               int sgd_nesterov, int do_ema, float ema_a, float ema_oma) {
  const b2_opt_chunk c = table[blockIdx.x];
  const int k = c.k < B2_OPT_MAX_K ? c.k : B2_OPT_MAX_K;
  __shared__ StepConsts sc;
  const int64_t it = *iter_dev;                       // optimiser steps completed so far
  const double lr_d = k > 0 ? lr_groups[c.group] : 0.0;
  const float lr = (float)lr_d;
  if (ALGO == 0 && threadIdx.x < k) {
    const double t = (double)(it * k + threadIdx.x + 1);        // this tensor's step counter for its j-th visit
    const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
    sc.step_size_neg[threadIdx.x] = (float)(-(lr_d / bc1));
    sc.bc2_sqrt[threadIdx.x] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2);
  const bool first = it == 0;
  float* __restrict__ p = c.p;
  const float* __restrict__ g = c.g;
  float* __restrict__ m = c.m;
  float* __restrict__ v = c.v;
  float* __restrict__ t = do_ema ? c.t : nullptr;
  const int n = c.count;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
  const int n4 = vec ? (n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    if (k > 0) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mv = reinterpret_cast<float4*>(m)[i];
      if (ALGO == 0) {
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam1(pv.x, gv.x, mv.x, vv.x, w1, b2, w2, eps, sc, k); adam1(pv.y, gv.y, mv.y, vv.y, w1, b2, w2, eps, sc, k);
        adam1(pv.z, gv.z, mv.z, vv.z, w1, b2, w2, eps, sc, k); adam1(pv.w, gv.w, mv.w, vv.w, w1, b2, w2, eps, sc, k);
        reinterpret_cast<float4*>(v)[i] = vv;
      } else {
        sgd1(pv.x, gv.x, mv.x, lr, sgd_mu, sgd_wd, sgd_nesterov, first, k); sgd1(pv.y, gv.y, mv.y, lr, sgd_mu, sgd_wd, sgd_nesterov, first, k);
        sgd1(pv.z, gv.z, mv.z, lr, sgd_mu, sgd_wd, sgd_nesterov, first, k); sgd1(pv.w, gv.w, mv.w, lr, sgd_mu, sgd_wd, sgd_nesterov, first, k);
      }
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(p)[i] = pv;
    }
    if (t) {
      float4 tv = reinterpret_cast<float4*>(t)[i];
      tv.x = ema1(tv.x, pv.x, ema_a, ema_oma); tv.y = ema1(tv.y, pv.y, ema_a, ema_oma);
      tv.z = ema1(tv.z, pv.z, ema_a, ema_oma); tv.w = ema1(tv.w, pv.w, ema_a, ema_oma);
      reinterpret_cast<float4*>(t)[i] = tv;
    }
  }
  for (int i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) {
    float pv = p[i];
    if (k > 0) {
      float mv = m[i];
      if (ALGO == 0) {
        float vv = v[i];
        adam1(pv, g[i], mv, vv, w1, b2, w2, eps, sc, k);
        v[i] = vv;
      } else {
        sgd1(pv, g[i], mv, lr, sgd_mu, sgd_wd, sgd_nesterov, first, k);
      }
      m[i] = mv;
      p[i] = pv;
    }
    if (t) t[i] = ema1(t[i], pv, ema_a, ema_oma);
  }
}

__global__ void opt_tick_kernel(int64_t* iter_dev) { *iter_dev += 1; }

}  // namespace

extern "C" int b2_opt_ema_step(const b2_opt_chunk* table, int64_t n_chunks, const double* lr_groups, int64_t* iter_dev,
                               int algo, double beta1, double beta2, float eps, float sgd_momentum,
                               float sgd_weight_decay, int sgd_nesterov, int do_ema, float ema_alpha,
                               float ema_one_minus_alpha, void* stream) {
  B2_REQUIRE(table && lr_groups && iter_dev && n_chunks > 0 && n_chunks < (1ll << 31), "b2_opt_ema_step: bad table");
  B2_REQUIRE(algo == 0 || algo == 1, "b2_opt_ema_step: algo must be 0 (adam) or 1 (sgd)");
  cudaStream_t s = (cudaStream_t)stream;
  if (algo == 0)
    opt_ema_kernel<0><<<(unsigned)n_chunks, OPT_THREADS, 0, s>>>(table, lr_groups, iter_dev, beta1, beta2, eps, 0.f, 0.f, 0,
                                                                do_ema, ema_alpha, ema_one_minus_alpha);
  else
    opt_ema_kernel<1><<<(unsigned)n_chunks, OPT_THREADS, 0, s>>>(table, lr_groups, iter_dev, 0.0, 0.0, 0.f, sgd_momentum,
                                                                sgd_weight_decay, sgd_nesterov, do_ema, ema_alpha,
                                                                ema_one_minus_alpha);
  B2_LAUNCH_CHECK("opt_ema_kernel");
  opt_tick_kernel<<<1, 1, 0, s>>>(iter_dev);
  B2_LAUNCH_CHECK("opt_tick_kernel");
  return B2_OK;
}
