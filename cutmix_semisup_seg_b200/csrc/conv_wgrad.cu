// Weight-gradient convolution on tcgen05 (kind::tf32, fp32 accumulate in TMEM), operands MN-major.
//   dW[m, tap, c] (+)= sum_pixels dY[pixel, m] * X[pixel @ tap, c]
// Backward of every student nn.Conv2d on the hot path (autograd of deeplab2.py:92-100 etc.,
// triggered at train_seg_semisup_mask_mt.py:301,459).
//
// GEMM view: M = output channels (tile 128), N = input channels (tile <= 256), K = pixels.
// Both operands are "pixel-major rows of channels" in HBM (NHWC), i.e. MN-major for this GEMM, so
// TMA boxes of (32 channels x 32 pixels) land in smem exactly as the canonical MN-major 128B-swizzle (32 B atom)
// atoms the tensor core reads; no transposes anywhere.  K is split over CTAs (pixel ranges) and the
// partial slabs are reduced in a fixed order by a second kernel => deterministic gradients.
#include <stdlib.h>
#include "tc_common.cuh"

namespace {

constexpr int W_BLOCK_M = 128;
constexpr int W_MAX_BLOCK_N = 256;
constexpr int W_KPIX = 32;                              // pixels (GEMM-K) per stage
constexpr int W_STAGES = 4;
constexpr int W_SLOT_BYTES = W_KPIX * 128;              // one 32-channel block: 32 pixel rows x 128 B
constexpr int W_A_STAGE_BYTES = (W_BLOCK_M / 32) * W_SLOT_BYTES;      // 16 KB
constexpr int W_B_STAGE_BYTES = (W_MAX_BLOCK_N / 32) * W_SLOT_BYTES;  // 32 KB
constexpr int W_SMEM_BYTES = W_STAGES * (W_A_STAGE_BYTES + W_B_STAGE_BYTES) + 1024 + 256;
constexpr int W_MAX_TAPS = 16;
constexpr int W_THREADS = 256;
constexpr int W_TMEM_COLS = 512;

struct WgradKArgs {
  int n, oh, ow, m;
  int ih, iw, c;
  int istride;
  int bw, bh, bn;            // pixel box, bw*bh*bn == kpix (multiple of 8)
  int kpix;
  int tiles_w, tiles_h, tiles_n, num_pb;
  int m_tiles, c_tiles, block_n;
  int n_taps;
  short dh[W_MAX_TAPS], dw[W_MAX_TAPS], wtap[W_MAX_TAPS];
  // per tap: the box columns [wlo, whi] / box rows [hlo, hhi] whose input pixels are not all padding (empty: lo > hi)
  int wlo[W_MAX_TAPS], whi[W_MAX_TAPS], hlo[W_MAX_TAPS], hhi[W_MAX_TAPS];
  int tw;
  int out_tiles, n_splits, num_units;
  int tap_step;              // split s works on tap (t + s * tap_step) % n_taps: see balance_units
  int n_pass;
  float* out;                // dW (n_splits == 1) or workspace slabs
  int64_t slab_elems;
  int accumulate;
  int desc_variant;
  const float* row_scale;   // per output channel m (folded BN scale), applied before accumulation
  int use5_a, use5_b;       // operand fetched with ONE 5-D TMA per stage (channel count multiple of 32)
  int m_tile_rows;          // 128 (single-CTA kernel) or 256 (CTA-pair kernel)
  int dbg;                  // debug knob 6 (timing experiments only): 8 = no TMA traffic, 16 = no MMAs
};

struct UnitInfo {
  int m0, c0, tap, split, pb_begin, pb_end;
};

__host__ __device__ __forceinline__ UnitInfo decode_unit(const WgradKArgs& a, int u) {
  UnitInfo r;
  const int ot = u % a.out_tiles;
  r.split = u / a.out_tiles;
  int t = ot;
  const int ct = t % a.c_tiles; t /= a.c_tiles;
  r.tap = (t % a.n_taps + r.split * a.tap_step) % a.n_taps;
  const int mt = t / a.n_taps;
  r.m0 = mt * a.m_tile_rows; r.c0 = ct * a.block_n;
  r.pb_begin = (int)(((int64_t)a.num_pb * r.split) / a.n_splits);
  r.pb_end = (int)(((int64_t)a.num_pb * (r.split + 1)) / a.n_splits);
  return r;
}

// Walks the pixel boxes [pb_begin, pb_end) of one work unit in order.  One division per unit (init), additions per box:
// the producer and the MMA issuer run this once per pipeline stage, and the issuing thread's instruction stream is what
// bounds this kernel (see tc_common.cuh "Warp-uniform role code").
struct PbWalk {
  int w0, h0, n0;      // output-space corner of the box (dY coordinates)
  int xw, xh;          // input-space corner for the unit's tap (X coordinates; may lie in the padding)
  int wt, ht;
  __host__ __device__ __forceinline__ void init(const WgradKArgs& a, int pb, int tap) {
    wt = pb % a.tiles_w; const int t = pb / a.tiles_w;
    ht = t % a.tiles_h;
    w0 = wt * a.bw; h0 = ht * a.bh; n0 = (t / a.tiles_h) * a.bn;
    xw = w0 * a.istride + a.dw[tap]; xh = h0 * a.istride + a.dh[tap];
  }
  // false: every input pixel of the box lies in the padding (contributes zero)
  __host__ __device__ __forceinline__ bool active(const WgradKArgs& a) const {
    return xh + (a.bh - 1) * a.istride >= 0 && xh < a.ih && xw + (a.bw - 1) * a.istride >= 0 && xw < a.iw;
  }
  __host__ __device__ __forceinline__ void next(const WgradKArgs& a, int tap) {
    if (++wt < a.tiles_w) { w0 += a.bw; xw += a.bw * a.istride; return; }
    wt = 0; w0 = 0; xw = a.dw[tap];
    if (++ht < a.tiles_h) { h0 += a.bh; xh += a.bh * a.istride; return; }
    ht = 0; h0 = 0; xh = a.dh[tap]; n0 += a.bn;
  }
};

// Number of pipeline stages (pixel boxes that are not all padding) of a unit, in closed form: the MMA issuer walks no
// boxes, its loop is wait / 4 MMAs / commit.  Agrees with PbWalk::active by construction of wlo..hhi (plan_wgrad).
__host__ __device__ __forceinline__ int active_before(const WgradKArgs& a, int tap, int pb) {   // active boxes with index < pb
  const int nw = max(a.whi[tap] - a.wlo[tap] + 1, 0), nh = max(a.hhi[tap] - a.hlo[tap] + 1, 0);
  const int wt = pb % a.tiles_w, t = pb / a.tiles_w;
  const int ht = t % a.tiles_h, nt = t / a.tiles_h;
  int f = nt * nh * nw + min(max(ht - a.hlo[tap], 0), nh) * nw;
  if (ht >= a.hlo[tap] && ht <= a.hhi[tap]) f += min(max(wt - a.wlo[tap], 0), nw);
  return f;
}
__host__ __device__ __forceinline__ int unit_boxes(const WgradKArgs& a, const UnitInfo& ui) {
  const int n = active_before(a, ui.tap, ui.pb_end) - active_before(a, ui.tap, ui.pb_begin);
  return n > 0 ? n : 1;        // a unit with no contributing box still runs one (all-zero X) box: see the producer
}

// Operand descriptors of the MMA issuer as (low word at stage 0, constant high word, low-word step per K step of 8
// pixels).  Product: MN-major SWIZZLE_128B_BASE32B atoms of (4 pixel rows x 128 B), LBO = stride between the 32-channel
// column blocks (one TMA box each), SBO = stride between 4-row groups along K, 8 pixels = 1024 B.  Debug knob 6 bits
// 32 / 64 (timing only, garbage results) pretend the operand is a K-major SWIZZLE_128B tile.
struct OperandDesc { uint32_t lo0, hi, step; };
__device__ __forceinline__ OperandDesc make_operand_desc(const WgradKArgs& a, uint32_t smem_addr, bool k_major_dbg) {
  OperandDesc d;
  if (k_major_dbg) {
    const uint64_t v = tc::make_smem_desc_sw128(smem_addr, 16, 1024);
    d.lo0 = (uint32_t)v; d.hi = (uint32_t)(v >> 32); d.step = 32 >> 4;
  } else {
    const uint32_t lbo = a.desc_variant == 1 ? 512u : (uint32_t)(a.kpix * 128);
    const uint32_t sbo = a.desc_variant == 1 ? (uint32_t)(a.kpix * 128) : 512u;
    const uint64_t v = tc::make_smem_desc(smem_addr, lbo, sbo, 1);
    d.lo0 = (uint32_t)v; d.hi = (uint32_t)(v >> 32); d.step = 1024 >> 4;
  }
  return d;
}

__global__ void __launch_bounds__(W_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYlo,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXlo,
                  const __grid_constant__ CUtensorMap tmY5, const __grid_constant__ CUtensorMap tmY5lo,
                  const __grid_constant__ CUtensorMap tmX5, const __grid_constant__ CUtensorMap tmX5lo,
                  const __grid_constant__ WgradKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + W_STAGES * W_A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_STAGES * (W_A_STAGE_BYTES + W_B_STAGE_BYTES));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + W_STAGES;
  uint64_t* tfull_bar = bars + 2 * W_STAGES;
  uint64_t* tempty_bar = bars + 2 * W_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * W_STAGES + 4);

  const int warp = (int)tc::uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;     // uniform: role branches converge

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmY); tc::tma_prefetch_desc(&tmX);
    if (a.n_pass > 1) { tc::tma_prefetch_desc(&tmYlo); tc::tma_prefetch_desc(&tmXlo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < W_STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_ptr, W_TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc::pdl_wait();                 // everything above overlaps the predecessor's tail (tc_common.cuh)
  tc::pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    const uint32_t el = tc::elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int u = blockIdx.x; u < a.num_units; u += gridDim.x) {
      const UnitInfo ui = decode_unit(a, u);
      int na = (a.m - ui.m0 + 31) / 32; if (na > W_BLOCK_M / 32) na = W_BLOCK_M / 32;
      int nbk = (a.c - ui.c0 + 31) / 32; if (nbk > a.block_n / 32) nbk = a.block_n / 32;
      const int slot = a.kpix * 128;   // bytes of one 32-channel block: kpix pixel rows x 128 B
      // 5-D boxes always carry the full block count of the tile (blocks past the tensor are zero-filled)
      const uint32_t tx = (uint32_t)((a.use5_a ? W_BLOCK_M / 32 : na) + (a.use5_b ? a.block_n / 32 : nbk)) * (uint32_t)slot;
      bool any = false;
      PbWalk b; b.init(a, ui.pb_begin, ui.tap);
      for (int pb = ui.pb_begin; pb < ui.pb_end; ++pb, b.next(a, ui.tap)) {
        // a unit with no contributing pixel box still runs one (all-zero X) box so that the
        // accumulator is written
        if (!b.active(a) && !(pb == ui.pb_end - 1 && !any)) continue;
        any = true;
        for (int p = 0; p < a.n_pass; ++p) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (el) {
            if (a.dbg & 8) {
              tc::mbar_arrive(&full_bar[stage]);
            } else {
              tc::mbar_expect_tx(&full_bar[stage], tx);
              uint8_t* sa = smem_a + stage * W_A_STAGE_BYTES;
              uint8_t* sb = smem_b + stage * W_B_STAGE_BYTES;
              if (a.use5_a) {
                tc::tma_load_5d(sa, (p & 1) ? &tmY5lo : &tmY5, &full_bar[stage], 0, b.w0, b.h0, b.n0, ui.m0 / 32);
              } else {
                const CUtensorMap* my = (p & 1) ? &tmYlo : &tmY;
                for (int j = 0; j < na; ++j)
                  tc::tma_load_4d(sa + j * slot, my, &full_bar[stage], ui.m0 + 32 * j, b.w0, b.h0, b.n0);
              }
              if (a.use5_b) {
                tc::tma_load_5d(sb, (p & 2) ? &tmX5lo : &tmX5, &full_bar[stage], 0, b.xw, b.xh, b.n0, ui.c0 / 32);
              } else {
                const CUtensorMap* mx = (p & 2) ? &tmXlo : &tmX;
                for (int j = 0; j < nbk; ++j)
                  tc::tma_load_4d(sb + j * slot, mx, &full_bar[stage], ui.c0 + 32 * j, b.xw, b.xh, b.n0);
              }
            }
          }
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t el = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_tf32(W_BLOCK_M, a.block_n, (a.dbg & 32) ? 0 : 1, (a.dbg & 64) ? 0 : 1);
    const OperandDesc da = make_operand_desc(a, tc::smem_u32(smem_a), a.dbg & 32);
    const OperandDesc db = make_operand_desc(a, tc::smem_u32(smem_b), a.dbg & 64);
    const bool no_mma = a.dbg & 16;
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const int ksteps = a.kpix / 8;
    for (int u = blockIdx.x; u < a.num_units; u += gridDim.x) {
      const UnitInfo ui = decode_unit(a, u);
      tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * W_MAX_BLOCK_N;
      uint32_t accum = 0;
      const int iters = unit_boxes(a, ui) * a.n_pass;
      {
        for (int it = 0; it < iters; ++it) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t a_lo = da.lo0 + (uint32_t)stage * (W_A_STAGE_BYTES >> 4);
          const uint32_t b_lo = db.lo0 + (uint32_t)stage * (W_B_STAGE_BYTES >> 4);
          if (el) {
            if (!no_mma) {
              if (ksteps == 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc::mma_tf32(tmem_d, ((uint64_t)da.hi << 32) | (a_lo + ks * da.step), ((uint64_t)db.hi << 32) | (b_lo + ks * db.step),
                               idesc, ks == 0 ? accum : 1u);
              } else {
                for (int ks = 0; ks < ksteps; ++ks)
                  tc::mma_tf32(tmem_d, ((uint64_t)da.hi << 32) | (a_lo + ks * da.step), ((uint64_t)db.hi << 32) | (b_lo + ks * db.step),
                               idesc, ks == 0 ? accum : 1u);
              }
            }
            tc::mma_commit(&empty_bar[stage]);
          }
          accum = 1;
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (el) tc::mma_commit(&tfull_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < a.num_units; u += gridDim.x) {
      const UnitInfo ui = decode_unit(a, u);
      const int mrow = ui.m0 + row;
      const bool valid = mrow < a.m;
      float* orow = a.out + (int64_t)ui.split * a.slab_elems + ((int64_t)mrow * a.tw + a.wtap[ui.tap]) * a.c;
      const float rs = (a.row_scale && valid) ? __ldg(a.row_scale + mrow) : 1.0f;
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * W_MAX_BLOCK_N;
      const int nchunks = a.block_n / 32;
      const bool vec = (a.c % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
      for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t r[32];
        tc::tmem_ld_x32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        if (valid) {
          const int col0 = ui.c0 + ch * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = col0 + j * 4;
            if (c >= a.c) break;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(r[j * 4 + e]) * rs;
            if (vec && c + 3 < a.c) {
              float4 o = make_float4(v[0], v[1], v[2], v[3]);
              if (a.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(orow + c);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              *reinterpret_cast<float4*>(orow + c) = o;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c + e < a.c) orow[c + e] = a.accumulate ? orow[c + e] + v[e] : v[e];
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, W_TMEM_COLS);
}

// =====================================================================================================================
// 2-CTA variant (cta_group::2, M = 256 output channels per CTA pair).  Same reason as conv_gemm2.cu: with fp32 operands a
// single-CTA 128 x 256 x 8 MMA reads 12 KB of shared memory per K step while TMA writes the same 12 KB, ~190 B/cycle
// against a ~128 B/cycle port, and the single-CTA kernel saturates at ~65 % of the tf32 rate (450-520 TFLOP/s on the
// layer3/4 shapes, profiles/r01_v7_*).  Here each CTA of the pair holds its own 128 output channels of dY and HALF of the
// input-channel tile of X, so per K step shared memory sees 8 KB of reads + 8 KB of TMA writes per CTA.
// Protocol as in conv_gemm2.cu (leader = cluster rank 0 issues the MMAs; TMA loads of both CTAs complete on the leader's
// full barrier; tcgen05.commit multicasts to both CTAs; both epilogues arrive on the leader's accumulator-empty barrier).
// Requirements (host-checked, else the single-CTA kernel runs): M % 256 == 0, C % 64 == 0 (5-D TMA for both operands).
constexpr int W2_STAGES_DEFAULT = 6;
constexpr int W2_STAGES_DEEP = 7;                                                 // b2_debug_set(13, 7): A/B of the operand ring depth
constexpr int W2_A_STAGE_BYTES = (W_BLOCK_M / 32) * W_SLOT_BYTES;              // 16 KB: this CTA's 128 output channels
constexpr int W2_B_STAGE_BYTES = (W_MAX_BLOCK_N / 64) * W_SLOT_BYTES;          // 16 KB: this CTA's half of the N tile
constexpr int w2_smem_bytes(int stages) { return stages * (W2_A_STAGE_BYTES + W2_B_STAGE_BYTES) + 1024 + 256; }
static_assert(w2_smem_bytes(W2_STAGES_DEEP) <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t w2_cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void w2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t W2_PEER_BIT_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared-window address -> leader CTA
__device__ __forceinline__ void w2_tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar) & W2_PEER_BIT_MASK),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void w2_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void w2_commit_mcast(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void w2_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(tc::smem_u32(bar)));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int W2_STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W_THREADS, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmY5, const __grid_constant__ CUtensorMap tmY5lo,
                   const __grid_constant__ CUtensorMap tmX5, const __grid_constant__ CUtensorMap tmX5lo,
                   const __grid_constant__ WgradKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + W2_STAGES * W2_A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W2_STAGES * (W2_A_STAGE_BYTES + W2_B_STAGE_BYTES));
  uint64_t* full_bar = bars;                          // [W2_STAGES]  leader's copies are the live ones
  uint64_t* empty_bar = bars + W2_STAGES;             // [W2_STAGES]  per CTA
  uint64_t* tfull_bar = bars + 2 * W2_STAGES;         // [2]          per CTA
  uint64_t* tempty_bar = bars + 2 * W2_STAGES + 2;    // [2]          leader's copies are the live ones
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * W2_STAGES + 4);

  const int warp = (int)tc::uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;     // uniform: role branches converge
  const uint32_t rank = tc::uniform(w2_cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmY5); tc::tma_prefetch_desc(&tmX5);
    if (a.n_pass > 1) { tc::tma_prefetch_desc(&tmY5lo); tc::tma_prefetch_desc(&tmX5lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < W2_STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 8); }   // 4 warps x 2 CTAs
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_ptr)), "r"(W_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  w2_cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc::pdl_wait();                 // everything above overlaps the predecessor's tail (tc_common.cuh)
  tc::pdl_launch_dependents();

  const int half_n = a.block_n / 2;                   // this CTA's input channels of the N tile (multiple of 32)
  const int slot = a.kpix * 128;
  const uint32_t stage_tx = 2u * (uint32_t)((W_BLOCK_M / 32) + half_n / 32) * (uint32_t)slot;     // both CTAs' bytes

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, one elected lane issues) =====================
    const uint32_t el = tc::elect_one();
    int stage = 0; uint32_t phase = 0;
    for (int u = cluster_id; u < a.num_units; u += num_clusters) {
      const UnitInfo ui = decode_unit(a, u);
      bool any = false;
      PbWalk b; b.init(a, ui.pb_begin, ui.tap);
      for (int pb = ui.pb_begin; pb < ui.pb_end; ++pb, b.next(a, ui.tap)) {
        if (!b.active(a) && !(pb == ui.pb_end - 1 && !any)) continue;
        any = true;
        for (int p = 0; p < a.n_pass; ++p) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (el) {
            if (a.dbg & 8) {
              if (leader) tc::mbar_arrive(&full_bar[stage]);
            } else {
              if (leader) tc::mbar_expect_tx(&full_bar[stage], stage_tx);
              w2_tma_load_5d(smem_a + stage * W2_A_STAGE_BYTES, (p & 1) ? &tmY5lo : &tmY5, &full_bar[stage], 0, b.w0, b.h0, b.n0,
                             (ui.m0 + (int)rank * W_BLOCK_M) / 32);
              w2_tma_load_5d(smem_b + stage * W2_B_STAGE_BYTES, (p & 2) ? &tmX5lo : &tmX5, &full_bar[stage], 0, b.xw, b.xh, b.n0,
                             (ui.c0 + (int)rank * half_n) / 32);
            }
          }
          if (++stage == W2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; converged warp, one elected lane issues) =====================
    if (leader) {
      const uint32_t el = tc::elect_one();
      const uint32_t idesc = tc::make_idesc_tf32(2 * W_BLOCK_M, a.block_n, (a.dbg & 32) ? 0 : 1, (a.dbg & 64) ? 0 : 1);
      const OperandDesc da = make_operand_desc(a, tc::smem_u32(smem_a), a.dbg & 32);
      const OperandDesc db = make_operand_desc(a, tc::smem_u32(smem_b), a.dbg & 64);
      const bool no_mma = a.dbg & 16;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const int ksteps = a.kpix / 8;
      for (int u = cluster_id; u < a.num_units; u += num_clusters) {
        const UnitInfo ui = decode_unit(a, u);
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * W_MAX_BLOCK_N;
        uint32_t accum = 0;
        const int iters = unit_boxes(a, ui) * a.n_pass;
        {
          for (int it = 0; it < iters; ++it) {
            tc::mbar_wait(&full_bar[stage], phase);
            tc::tc_fence_after();
            const uint32_t a_lo = da.lo0 + (uint32_t)stage * (W2_A_STAGE_BYTES >> 4);
            const uint32_t b_lo = db.lo0 + (uint32_t)stage * (W2_B_STAGE_BYTES >> 4);
            if (el) {
              if (!no_mma) {
                if (ksteps == 4) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    w2_mma_tf32(tmem_d, ((uint64_t)da.hi << 32) | (a_lo + ks * da.step), ((uint64_t)db.hi << 32) | (b_lo + ks * db.step),
                                idesc, ks == 0 ? accum : 1u);
                } else {
                  for (int ks = 0; ks < ksteps; ++ks)
                    w2_mma_tf32(tmem_d, ((uint64_t)da.hi << 32) | (a_lo + ks * da.step), ((uint64_t)db.hi << 32) | (b_lo + ks * db.step),
                                idesc, ks == 0 ? accum : 1u);
                }
              }
              w2_commit_mcast(&empty_bar[stage]);
            }
            accum = 1;
            if (++stage == W2_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (el) w2_commit_mcast(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: own 128 output channels) =====================
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = cluster_id; u < a.num_units; u += num_clusters) {
      const UnitInfo ui = decode_unit(a, u);
      const int mrow = ui.m0 + (int)rank * W_BLOCK_M + row;
      const bool valid = mrow < a.m;
      float* orow = a.out + (int64_t)ui.split * a.slab_elems + ((int64_t)mrow * a.tw + a.wtap[ui.tap]) * a.c;
      const float rs = (a.row_scale && valid) ? __ldg(a.row_scale + mrow) : 1.0f;
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * W_MAX_BLOCK_N;
      const int nchunks = a.block_n / 32;
      for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t r[32];
        tc::tmem_ld_x32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        if (valid) {
          const int col0 = ui.c0 + ch * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = col0 + j * 4;
            if (c >= a.c) break;
            float4 o = make_float4(__uint_as_float(r[j * 4]) * rs, __uint_as_float(r[j * 4 + 1]) * rs,
                                   __uint_as_float(r[j * 4 + 2]) * rs, __uint_as_float(r[j * 4 + 3]) * rs);
            if (a.accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(orow + c);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(orow + c) = o;       // C % 64 == 0 and a 16 B aligned output (host-checked)
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) w2_arrive_leader(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  w2_cluster_sync();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(W_TMEM_COLS) : "memory");
}

// dw[i] = (accumulate ? dw[i] : 0) + sum_s slabs[s][i], fixed order.
__global__ void wgrad_reduce_kernel(const float* __restrict__ slabs, int n_splits, int64_t elems,
                                    float* __restrict__ dw, int accumulate) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  tc::pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += stride) {
    float s = accumulate ? dw[i] : 0.0f;
    for (int k = 0; k < n_splits; ++k) s += slabs[(int64_t)k * elems + i];
    dw[i] = s;
  }
}

int g_wgrad_desc_variant = 0;
int g_wgrad_dbg = 0;              // b2_debug_set(6, v)
int g_wgrad_force_1cta = 0;       // b2_debug_set(5, 1): always the single-CTA kernel
int g_wgrad2_stages = -1;         // b2_debug_set(13, 6 | 7) / environment B200SEG_WGRAD_STAGES: operand ring depth of the CTA-pair kernel

struct WPlan {
  WgradKArgs a;
};

// Pixel box with bw*bh*bn == kpix in {32,24,16,8}: every smem K-row must hold real (or TMA-zero) data.
void choose_kbox(int ow, int oh, int n, int istride, int* bw_o, int* bh_o, int* bn_o, int* kpix_o) {
  double best = -1.0; int bbw = 8, bbh = 1, bbn = 1, bk = 8;
  const int max_b = 256 / istride;
  for (int bw = 1; bw <= 32 && bw <= max_b; ++bw)
    for (int bh = 1; bw * bh <= 32 && bh <= max_b; ++bh)
      for (int bn = 1; bw * bh * bn <= 32; ++bn) {
        const int kp = bw * bh * bn;
        if (kp % 8) continue;
        if (bn > 1 && !(bw >= ow && bh >= oh)) continue;
        const double tiles = (double)((ow + bw - 1) / bw) * ((oh + bh - 1) / bh) * ((n + bn - 1) / bn);
        const double eff = ((double)ow * oh * n) / (tiles * kp);    // useful rows / loaded rows
        const double score = eff * (0.75 + 0.25 * kp / 32.0) + 1e-6 * bw;
        if (score > best) { best = score; bbw = bw; bbh = bh; bbn = bn; bk = kp; }
      }
  *bw_o = bbw; *bh_o = bbh; *bn_o = bbn; *kpix_o = bk;
}

// Load balance of dilated layers.  A work unit is (output tile, tap, pixel range); pixel boxes whose input lies in the padding are
// skipped, so the units of an off-centre tap are shorter than those of the centre tap (ASPP, 64 x 64 map: dilation 12 skips 19 %
// of the boxes of six taps, dilation 36 up to 78 %).  With 72 units on 74 CTA pairs every pair owns ONE unit and the launch lasts
// as long as the centre tap (4096 stages) although the mean unit is 3487 / 2989 / 1661 stages (d = 12 / 24 / 36).  More pixel
// splits give every pair several units, and rotating the tap by `tap_step` from split to split makes them units of DIFFERENT
// taps.  The pair (splits, tap_step) minimising  [stages of the busiest worker + cost of the extra slabs]  is searched once per
// geometry (closed-form stage counts, a few thousand integer evaluations) and cached.  b2_debug_set(14, 0) disables it.
int g_wgrad_balance = 1;
int g_wgrad_force_splits = 0, g_wgrad_force_step = -1;     // b2_debug_set(15 / 16, v): experiments (pixel splits, tap rotation)
struct BalanceKey {
  int n, oh, ow, ih, iw, istride, m_tiles, c_tiles, n_taps, workers, base_splits, bw, bh, bn;
  short dh[W_MAX_TAPS], dw[W_MAX_TAPS];
  int64_t slab_elems;
};
struct BalanceEntry { BalanceKey key; int splits, tap_step; };
static int64_t busiest_worker(WgradKArgs& a, int workers, int64_t* total) {
  int64_t load[1024];
  const int w = workers < 1024 ? workers : 1024;
  for (int i = 0; i < w; ++i) load[i] = 0;
  int64_t sum = 0;
  for (int u = 0; u < a.num_units; ++u) {
    const int64_t b = unit_boxes(a, decode_unit(a, u));
    load[u % w] += b; sum += b;
  }
  int64_t worst = 0;
  for (int i = 0; i < w; ++i) if (load[i] > worst) worst = load[i];
  if (total) *total = sum;
  return worst;
}
static void balance_units(WgradKArgs& a, int workers, int max_splits) {
  if (a.n_taps < 2 || workers < 1) return;
  static thread_local BalanceEntry cache[64];
  static thread_local int n_cached = 0;
  BalanceKey key;
  memset(&key, 0, sizeof(key));
  key.n = a.n; key.oh = a.oh; key.ow = a.ow; key.ih = a.ih; key.iw = a.iw; key.istride = a.istride; key.m_tiles = a.m_tiles;
  key.c_tiles = a.c_tiles; key.n_taps = a.n_taps; key.workers = workers; key.base_splits = a.n_splits; key.bw = a.bw; key.bh = a.bh;
  key.bn = a.bn; key.slab_elems = a.slab_elems;
  for (int i = 0; i < a.n_taps; ++i) { key.dh[i] = a.dh[i]; key.dw[i] = a.dw[i]; }
  for (int i = 0; i < n_cached; ++i)
    if (memcmp(&cache[i].key, &key, sizeof(key)) == 0) {
      a.n_splits = cache[i].splits; a.tap_step = cache[i].tap_step; a.num_units = a.out_tiles * a.n_splits;
      return;
    }
  const int base = a.n_splits;
  int64_t total = 0;
  const int64_t worst0 = busiest_worker(a, workers, &total);
  int best_s = base, best_step = 0;
  // one pipeline stage ~ 0.35 us (512 tensor cycles); a slab costs a write by the kernel and a read by the reduction (~5 TB/s)
  const double slab_stages = 2.0 * (double)a.slab_elems * 4.0 / 5.0e12 / 0.35e-6;
  auto cost = [&](int64_t worst, int s) { return (double)worst + (s > 1 ? slab_stages * s : 0.0); };
  double best = cost(worst0, base);
  const int64_t mean = total / (workers < a.num_units ? workers : a.num_units);
  if (worst0 * 100 > mean * 105) {              // otherwise already balanced within 5 %
    static const int mult[] = {1, 2, 3, 4, 5, 6, 8, 9};
    for (int mi = 0; mi < 8; ++mi) {
      const int sp = base * mult[mi];
      if (sp > max_splits || sp > 32 || (int64_t)a.out_tiles * sp > 16384) continue;
      for (int step = 0; step < a.n_taps; ++step) {
        if (sp == base && step == 0) continue;
        a.n_splits = sp; a.tap_step = step; a.num_units = a.out_tiles * sp;
        const double c = cost(busiest_worker(a, workers, nullptr), sp);
        if (c < best * 0.97) { best = c; best_s = sp; best_step = step; }      // change the plan only for a real gain
      }
    }
  }
  a.n_splits = best_s; a.tap_step = best_step; a.num_units = a.out_tiles * best_s;
  if (n_cached < 64) { cache[n_cached].key = key; cache[n_cached].splits = best_s; cache[n_cached].tap_step = best_step; ++n_cached; }
}

int plan_wgrad(const b2_wgrad_params* p, WgradKArgs* out) {
  B2_REQUIRE(p && p->dy && p->x && p->dw, "b2_conv_wgrad: null tensor");
  B2_REQUIRE(p->n > 0 && p->oh > 0 && p->ow > 0 && p->m > 0 && p->ih > 0 && p->iw > 0 && p->c > 0, "b2_conv_wgrad: bad dims");
  B2_REQUIRE(p->ldy % 4 == 0 && p->ldy >= p->m, "b2_conv_wgrad: ldy=%d must be a multiple of 4 and >= M", p->ldy);
  B2_REQUIRE(p->ldx % 4 == 0 && p->ldx >= p->c, "b2_conv_wgrad: ldx=%d must be a multiple of 4 and >= C", p->ldx);
  B2_REQUIRE(p->n_taps >= 1 && p->n_taps <= W_MAX_TAPS && p->taps, "b2_conv_wgrad: n_taps out of range");
  B2_REQUIRE(p->istride >= 1 && p->istride <= 8, "b2_conv_wgrad: bad istride");
  B2_REQUIRE(p->n_split == 1 || p->n_split == 3 || p->n_split == 4, "b2_conv_wgrad: n_split must be 1, 3 or 4");
  B2_REQUIRE(p->n_split == 1 || (p->dy_lo && p->x_lo), "b2_conv_wgrad: split mode needs dy_lo and x_lo");
  WgradKArgs a;
  memset(&a, 0, sizeof(a));
  a.n = p->n; a.oh = p->oh; a.ow = p->ow; a.m = p->m; a.ih = p->ih; a.iw = p->iw; a.c = p->c; a.istride = p->istride;
  choose_kbox(p->ow, p->oh, p->n, p->istride, &a.bw, &a.bh, &a.bn, &a.kpix);
  a.tiles_w = (p->ow + a.bw - 1) / a.bw; a.tiles_h = (p->oh + a.bh - 1) / a.bh; a.tiles_n = (p->n + a.bn - 1) / a.bn;
  const int64_t num_pb = (int64_t)a.tiles_w * a.tiles_h * a.tiles_n;
  B2_REQUIRE(num_pb < (1ll << 31), "b2_conv_wgrad: too many pixel boxes");
  a.num_pb = (int)num_pb;
  // CTA-pair kernel (M = 256 per pair, each CTA loads half of the N tile): needs whole 256-channel M tiles, N tiles that
  // split into two multiples of 32, the 5-D TMA form of both operands and vector stores.
  const bool two_cta = !g_wgrad_force_1cta && p->m % 256 == 0 && p->c % 64 == 0 && p->max_ctas != 1 &&
                       (reinterpret_cast<uintptr_t>(p->dw) & 15) == 0 &&
                       (!p->workspace || (reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0);
  a.m_tile_rows = two_cta ? 2 * W_BLOCK_M : W_BLOCK_M;
  a.m_tiles = (p->m + a.m_tile_rows - 1) / a.m_tile_rows;
  // balanced N tiles: C = 304 -> two tiles of 160 channels instead of 256 + 48
  const int gran = two_cta ? 64 : 32;
  const int cblk = (p->c + gran - 1) / gran;                         // channel blocks of `gran`
  a.c_tiles = (cblk * gran + W_MAX_BLOCK_N - 1) / W_MAX_BLOCK_N;
  const int block_n = ((cblk + a.c_tiles - 1) / a.c_tiles) * gran;
  a.block_n = block_n;
  a.n_taps = p->n_taps;
  for (int i = 0; i < p->n_taps; ++i) {
    a.dh[i] = (short)p->taps[i * 3 + 0]; a.dw[i] = (short)p->taps[i * 3 + 1]; a.wtap[i] = (short)p->taps[i * 3 + 2];
    B2_REQUIRE(p->taps[i * 3 + 2] >= 0 && p->taps[i * 3 + 2] < p->tw, "b2_conv_wgrad: tap index out of range");
  }
  for (int i = 0; i < p->n_taps; ++i) {      // same predicate as PbWalk::active, per axis (the active set is an interval)
    int wlo = 1, whi = 0, hlo = 1, hhi = 0; bool fw = false, fh = false;
    for (int wt = 0; wt < a.tiles_w; ++wt) {
      const int xw = wt * a.bw * a.istride + a.dw[i];
      if (xw + (a.bw - 1) * a.istride >= 0 && xw < a.iw) { if (!fw) { wlo = wt; fw = true; } whi = wt; }
    }
    for (int ht = 0; ht < a.tiles_h; ++ht) {
      const int xh = ht * a.bh * a.istride + a.dh[i];
      if (xh + (a.bh - 1) * a.istride >= 0 && xh < a.ih) { if (!fh) { hlo = ht; fh = true; } hhi = ht; }
    }
    a.wlo[i] = wlo; a.whi[i] = whi; a.hlo[i] = hlo; a.hhi[i] = hhi;
  }
  a.tw = p->tw;
  a.out_tiles = a.m_tiles * a.n_taps * a.c_tiles;
  int sms = b2_sm_count_cached();
  if (sms <= 0) sms = 148;
  if (p->max_ctas > 0 && p->max_ctas < sms) sms = p->max_ctas;
  if (two_cta) sms /= 2;            // work units are processed by CTA pairs
  int splits = sms / a.out_tiles;
  if (splits < 1) splits = 1;
  // keep at least ~8 pixel boxes per split so the pipeline fills
  const int max_splits = (int)((num_pb + 7) / 8);
  if (splits > max_splits) splits = max_splits < 1 ? 1 : max_splits;
  if (p->kchunk > 0) {
    // precision mode: the tensor core accumulates in fp32 with truncation, so the error of one accumulator grows
    // with the number of pixels it sums; bound that count and let the fp32 (round-to-nearest) slab reduction
    // combine the partial sums.
    int64_t want = (num_pb * a.kpix + p->kchunk - 1) / p->kchunk;
    if (want > num_pb) want = num_pb;
    if (want > 4096) want = 4096;
    if (want > splits) splits = (int)want;
  }
  a.n_splits = splits;
  a.num_units = a.out_tiles * splits;
  a.n_pass = p->n_split;
  a.slab_elems = (int64_t)p->m * p->tw * p->c;
  if (p->kchunk == 0 && g_wgrad_balance != 0) balance_units(a, sms, max_splits);
  if (p->kchunk == 0 && g_wgrad_force_splits > 0 && g_wgrad_force_splits <= max_splits) {
    a.n_splits = g_wgrad_force_splits; a.num_units = a.out_tiles * a.n_splits;
    a.tap_step = g_wgrad_force_step >= 0 ? g_wgrad_force_step % a.n_taps : 0;
  }
  a.accumulate = p->accumulate;
  a.desc_variant = g_wgrad_desc_variant;
  a.dbg = g_wgrad_dbg;
  a.row_scale = p->row_scale;
  *out = a;
  return B2_OK;
}

}  // namespace

extern int g_conv_force_1cta;
extern int g_conv_epi_debug;
extern int g_conv_pf_max_k;
extern int g_conv_tap_outer;
extern int g_conv_tma_epi;
extern int g_conv_main_stages;
extern int g_conv_pdl;
extern int g_conv_alt_dir;
extern int g_conv_tap_tables;
extern int g_conv_wide_pf;
extern int g_netops_flat_kernels;
extern "C" void b2_debug_set(int key, int value) {
  if (key == 11) g_conv_pdl = value;
  if (key == 12) g_conv_alt_dir = value;
  if (key == 13) g_wgrad2_stages = value;
  if (key == 14) g_wgrad_balance = value;
  if (key == 17) g_conv_tap_tables = value;
  if (key == 18) g_conv_wide_pf = value;
  if (key == 20) g_netops_flat_kernels = value;
  if (key == 15) g_wgrad_force_splits = value;
  if (key == 16) g_wgrad_force_step = value;
  if (key == 5) g_wgrad_force_1cta = value;
  if (key == 6) g_wgrad_dbg = value;
  if (key == 1) g_wgrad_desc_variant = value;
  if (key == 2) g_conv_force_1cta = value;
  if (key == 3) g_conv_epi_debug = value;
  if (key == 4) g_conv_pf_max_k = value;
  if (key == 7) g_conv_tap_outer = value;
  if (key == 8) g_conv_tma_epi = value;
  if (key == 10) g_conv_main_stages = value;
}

extern "C" size_t b2_conv_wgrad_workspace(const b2_wgrad_params* p) {
  WgradKArgs a;
  if (plan_wgrad(p, &a) != B2_OK) return 0;
  return a.n_splits > 1 ? (size_t)a.n_splits * a.slab_elems * sizeof(float) : 0;
}

// Host-side self check of the work decomposition (no GPU needed, no tensor is touched): for every work unit of the plan the
// MMA issuer's closed-form stage count (unit_boxes) must equal the number of boxes the producer's walk (PbWalk) loads --
// a disagreement would leave one of the two pipeline roles waiting forever.  out[0] = splits, out[1] = work units,
// out[2] = pipeline stages of the whole launch, out[3] = units that disagree, out[4] = 1 if the CTA-pair kernel is chosen.
extern "C" int b2_conv_wgrad_plan_check(const b2_wgrad_params* p, int64_t* out) {
  B2_REQUIRE(out, "b2_conv_wgrad_plan_check: null output");
  WgradKArgs a;
  int rc = plan_wgrad(p, &a);
  if (rc) return rc;
  int64_t stages = 0, bad = 0;
  for (int u = 0; u < a.num_units; ++u) {
    const UnitInfo ui = decode_unit(a, u);
    int walked = 0; bool any = false;
    PbWalk b; b.init(a, ui.pb_begin, ui.tap);
    for (int pb = ui.pb_begin; pb < ui.pb_end; ++pb, b.next(a, ui.tap)) {
      if (!b.active(a) && !(pb == ui.pb_end - 1 && !any)) continue;
      any = true; ++walked;
    }
    if (walked != unit_boxes(a, ui) || ui.pb_end <= ui.pb_begin) ++bad;
    stages += (int64_t)walked * a.n_pass;
  }
  out[0] = a.n_splits; out[1] = a.num_units; out[2] = stages; out[3] = bad; out[4] = a.m_tile_rows == 2 * W_BLOCK_M;
  return B2_OK;
}

// Host-only view of the load balance of the plan: out = int64[5] {pixel splits, tap rotation per split, pipeline stages of the
// busiest worker (CTA or CTA pair), pipeline stages of the whole launch, workers}.
extern "C" int b2_conv_wgrad_plan_balance(const b2_wgrad_params* p, int64_t* out) {
  B2_REQUIRE(out, "b2_conv_wgrad_plan_balance: null output");
  WgradKArgs a;
  int rc = plan_wgrad(p, &a);
  if (rc) return rc;
  int workers = b2_sm_count_cached();
  if (workers <= 0) workers = 148;
  if (p->max_ctas > 0 && p->max_ctas < workers) workers = p->max_ctas;
  if (a.m_tile_rows == 2 * W_BLOCK_M) workers /= 2;
  if (workers > a.num_units) workers = a.num_units;
  int64_t total = 0;
  const int64_t worst = busiest_worker(a, workers, &total);
  out[0] = a.n_splits; out[1] = a.tap_step; out[2] = worst * a.n_pass; out[3] = total * a.n_pass; out[4] = workers;
  return B2_OK;
}

extern "C" int b2_conv_wgrad(const b2_wgrad_params* p, void* stream) {
  WgradKArgs a;
  int rc = plan_wgrad(p, &a);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (a.n_splits > 1) {
    const size_t need = (size_t)a.n_splits * a.slab_elems * sizeof(float);
    B2_REQUIRE(p->workspace && p->workspace_bytes >= need, "b2_conv_wgrad: workspace too small (%zu < %zu)", p->workspace_bytes, need);
    a.out = reinterpret_cast<float*>(p->workspace);
    a.accumulate = 0;
  } else {
    a.out = p->dw;
    a.slab_elems = 0;
  }

  CUtensorMap tmY, tmYlo, tmX, tmXlo, tmY5, tmY5lo, tmX5, tmX5lo;
  a.use5_a = (p->m % 32 == 0) ? 1 : 0;
  a.use5_b = (p->c % 32 == 0) ? 1 : 0;
  {
    const uint64_t dims[4] = {(uint64_t)p->m, (uint64_t)p->ow, (uint64_t)p->oh, (uint64_t)p->n};
    const uint64_t strides[3] = {(uint64_t)p->ldy * 4, (uint64_t)p->ow * p->ldy * 4, (uint64_t)p->oh * p->ow * p->ldy * 4};
    const uint32_t box[4] = {32, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn};
    const uint32_t es[4] = {1, 1, 1, 1};
    rc = tc::make_tmap_f32(&tmY, p->dy, 4, dims, strides, box, es, true); if (rc) return rc;
    rc = tc::make_tmap_f32(&tmYlo, p->dy_lo ? p->dy_lo : p->dy, 4, dims, strides, box, es, true); if (rc) return rc;
    tmY5 = tmY; tmY5lo = tmYlo;
    if (a.use5_a) {   // (32, OW, OH, N, M/32): channel blocks as an outer dimension with a 128 B stride
      const uint64_t d5[5] = {32, (uint64_t)p->ow, (uint64_t)p->oh, (uint64_t)p->n, (uint64_t)(p->m / 32)};
      const uint64_t s5[4] = {(uint64_t)p->ldy * 4, (uint64_t)p->ow * p->ldy * 4, (uint64_t)p->oh * p->ow * p->ldy * 4, 128};
      const uint32_t b5[5] = {32, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn, (uint32_t)(W_BLOCK_M / 32)};
      const uint32_t e5[5] = {1, 1, 1, 1, 1};
      if (tc::make_tmap_f32(&tmY5, p->dy, 5, d5, s5, b5, e5, true) != B2_OK ||
          tc::make_tmap_f32(&tmY5lo, p->dy_lo ? p->dy_lo : p->dy, 5, d5, s5, b5, e5, true) != B2_OK) {
        a.use5_a = 0; tmY5 = tmY; tmY5lo = tmYlo;
      }
    }
  }
  {
    const uint64_t dims[4] = {(uint64_t)p->c, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n};
    const uint64_t strides[3] = {(uint64_t)p->ldx * 4, (uint64_t)p->iw * p->ldx * 4, (uint64_t)p->ih * p->iw * p->ldx * 4};
    const uint32_t box[4] = {32, (uint32_t)(a.bw * p->istride), (uint32_t)(a.bh * p->istride), (uint32_t)a.bn};
    const uint32_t es[4] = {1, (uint32_t)p->istride, (uint32_t)p->istride, 1};
    rc = tc::make_tmap_f32(&tmX, p->x, 4, dims, strides, box, es, true); if (rc) return rc;
    rc = tc::make_tmap_f32(&tmXlo, p->x_lo ? p->x_lo : p->x, 4, dims, strides, box, es, true); if (rc) return rc;
    tmX5 = tmX; tmX5lo = tmXlo;
    if (a.use5_b) {
      const uint64_t d5[5] = {32, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n, (uint64_t)(p->c / 32)};
      const uint64_t s5[4] = {(uint64_t)p->ldx * 4, (uint64_t)p->iw * p->ldx * 4, (uint64_t)p->ih * p->iw * p->ldx * 4, 128};
      const uint32_t b5[5] = {32, (uint32_t)(a.bw * p->istride), (uint32_t)(a.bh * p->istride), (uint32_t)a.bn, (uint32_t)(a.block_n / 32)};
      const uint32_t e5[5] = {1, (uint32_t)p->istride, (uint32_t)p->istride, 1, 1};
      if (tc::make_tmap_f32(&tmX5, p->x, 5, d5, s5, b5, e5, true) != B2_OK ||
          tc::make_tmap_f32(&tmX5lo, p->x_lo ? p->x_lo : p->x, 5, d5, s5, b5, e5, true) != B2_OK) {
        a.use5_b = 0; tmX5 = tmX; tmX5lo = tmXlo;
      }
    }
  }

  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES));
    B2_CUDA(cudaFuncSetAttribute(conv_wgrad2_kernel<W2_STAGES_DEFAULT>, cudaFuncAttributeMaxDynamicSharedMemorySize, w2_smem_bytes(W2_STAGES_DEFAULT)));
    B2_CUDA(cudaFuncSetAttribute(conv_wgrad2_kernel<W2_STAGES_DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, w2_smem_bytes(W2_STAGES_DEEP)));
    attr_set = true;
  }
  if (a.m_tile_rows == 2 * W_BLOCK_M) {
    // CTA-pair kernel: 5-D maps whose outermost box covers this CTA's share of the tile (4 blocks of dY, block_n/64 of X)
    B2_REQUIRE(a.use5_a && a.use5_b, "b2_conv_wgrad: 5-D tensor maps unavailable for the CTA-pair kernel");
    CUtensorMap y5, y5lo, x5, x5lo;
    {
      const uint64_t d5[5] = {32, (uint64_t)p->ow, (uint64_t)p->oh, (uint64_t)p->n, (uint64_t)(p->m / 32)};
      const uint64_t s5[4] = {(uint64_t)p->ldy * 4, (uint64_t)p->ow * p->ldy * 4, (uint64_t)p->oh * p->ow * p->ldy * 4, 128};
      const uint32_t b5[5] = {32, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn, (uint32_t)(W_BLOCK_M / 32)};
      const uint32_t e5[5] = {1, 1, 1, 1, 1};
      rc = tc::make_tmap_f32(&y5, p->dy, 5, d5, s5, b5, e5, true); if (rc) return rc;
      rc = tc::make_tmap_f32(&y5lo, p->dy_lo ? p->dy_lo : p->dy, 5, d5, s5, b5, e5, true); if (rc) return rc;
    }
    {
      const uint64_t d5[5] = {32, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n, (uint64_t)(p->c / 32)};
      const uint64_t s5[4] = {(uint64_t)p->ldx * 4, (uint64_t)p->iw * p->ldx * 4, (uint64_t)p->ih * p->iw * p->ldx * 4, 128};
      const uint32_t b5[5] = {32, (uint32_t)(a.bw * p->istride), (uint32_t)(a.bh * p->istride), (uint32_t)a.bn, (uint32_t)(a.block_n / 64)};
      const uint32_t e5[5] = {1, (uint32_t)p->istride, (uint32_t)p->istride, 1, 1};
      rc = tc::make_tmap_f32(&x5, p->x, 5, d5, s5, b5, e5, true); if (rc) return rc;
      rc = tc::make_tmap_f32(&x5lo, p->x_lo ? p->x_lo : p->x, 5, d5, s5, b5, e5, true); if (rc) return rc;
    }
    int clusters = b2_sm_count_cached() / 2;
    if (clusters <= 0) return b2_fail(B2_ERR_CUDA, "b2_conv_wgrad: no CUDA device");
    if (p->max_ctas > 0 && p->max_ctas / 2 < clusters) clusters = p->max_ctas / 2 > 0 ? p->max_ctas / 2 : 1;
    if (clusters > a.num_units) clusters = a.num_units;
    if (g_wgrad2_stages < 0) { const char* e = getenv("B200SEG_WGRAD_STAGES"); g_wgrad2_stages = e ? atoi(e) : W2_STAGES_DEFAULT; }
    if (g_wgrad2_stages == W2_STAGES_DEEP)
      tc::launch(conv_wgrad2_kernel<W2_STAGES_DEEP>, clusters * 2, W_THREADS, w2_smem_bytes(W2_STAGES_DEEP), s, y5, y5lo, x5, x5lo, a);
    else
      tc::launch(conv_wgrad2_kernel<W2_STAGES_DEFAULT>, clusters * 2, W_THREADS, w2_smem_bytes(W2_STAGES_DEFAULT), s, y5, y5lo, x5, x5lo, a);
    B2_LAUNCH_CHECK("conv_wgrad2_kernel");
    if (a.n_splits > 1) {
      const int64_t elems = (int64_t)p->m * p->tw * p->c;
      int64_t blocks = ceil_div64(elems, 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      tc::launch(wgrad_reduce_kernel, (unsigned)blocks, 256, 0, s, reinterpret_cast<const float*>(p->workspace), a.n_splits, elems, p->dw, p->accumulate);
      B2_LAUNCH_CHECK("wgrad_reduce_kernel");
    }
    return B2_OK;
  }
  int grid = b2_sm_count_cached();
  if (grid <= 0) return b2_fail(B2_ERR_CUDA, "b2_conv_wgrad: no CUDA device");
  if (p->max_ctas > 0 && p->max_ctas < grid) grid = p->max_ctas;
  if (grid > a.num_units) grid = a.num_units;
  tc::launch(conv_wgrad_kernel, (unsigned)grid, W_THREADS, W_SMEM_BYTES, s, tmY, tmYlo, tmX, tmXlo, tmY5, tmY5lo, tmX5, tmX5lo, a);
  B2_LAUNCH_CHECK("conv_wgrad_kernel");
  if (a.n_splits > 1) {
    const int64_t elems = (int64_t)p->m * p->tw * p->c;
    int64_t blocks = ceil_div64(elems, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    tc::launch(wgrad_reduce_kernel, (unsigned)blocks, 256, 0, s, reinterpret_cast<const float*>(p->workspace), a.n_splits, elems, p->dw, p->accumulate);
    B2_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  return B2_OK;
}
