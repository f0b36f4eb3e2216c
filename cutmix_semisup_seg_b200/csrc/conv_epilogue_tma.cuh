// TMA epilogue of the CTA-pair convolution kernel (conv_gemm2.cu, PF build): TMEM accumulator -> HBM with every byte of
// HBM traffic moved by the TMA engine.
//
// Why (profiles/r01_v11_pipeline_trace_1x1.log, r01_v6_epilogue_experiment.log): the 1x1 layers with a 1024 / 2048-channel side
// (K <= 512; ~35 ms of the 146 ms cfg3 iteration) are bound by their epilogue, and the register-path epilogue
// (conv_epilogue.cuh) is bound by instruction issue: ~590 SASS instructions per 32 x 32 chunk (64-bit address arithmetic and
// predicates per row and operand, transposing staging pass, per-lane 16 B loads / stores); it ran at the same speed with its
// HBM stores disabled.  Here a lane keeps ITS accumulator row (tcgen05.ld 32x32b: lane = pixel, register = channel):
//   * the residual addend / ReLU gate tiles of the chunk arrive as 128B-swizzled 32 pixel x 32 channel TMA boxes in a per-warp
//     slot (one mbarrier per warp; the next chunk's boxes -- of the next tile if need be -- are requested as soon as the
//     current ones have been read, so 8 warps x 8 KB are in flight per SM),
//   * the lane combines its row with the operand rows (conflict-free 16 B shared-memory accesses: chunk j of row r lives at
//     16 B slot j ^ (r & 7)), per-channel scale / shift come from a 384 B per-warp table,
//   * the finished 4 KB tile leaves through ONE cp.async.bulk.tensor store (reduce-add for accumulating launches); rows /
//     channels outside the tensor are clipped by the TMA unit, so there is no per-row predicate or address at all.
// ~200 instructions per chunk, no per-lane global access.  Fused column statistics (frozen-BN parameter gradients) re-read the
// staged tile column-wise (same swizzle, conflict-free) and reduce over rows with two shuffles, as before.
//
// Usable when the 32 accumulator rows of a lane quarter form a rectangular box of the output tensor (host: epi2_box), the
// output is written with unit stride / no phase offset, operands are 16 B aligned with 16 B pitches, and no `sub` operand is
// used; everything else keeps the register path.
#pragma once
#include "conv_epilogue.cuh"

namespace epi2 {

constexpr int TILE_BYTES = 32 * 128;                     // 32 rows x 32 fp32 channels
constexpr int WARP_TILES = 3;                            // staging, addend slot, gate slot
constexpr int WARP_BYTES = WARP_TILES * TILE_BYTES;      // 12 KB, 1024 B aligned
constexpr int WARP_AUX_BYTES = 3 * 32 * 4 + 16 + 32 * 4; // scale/shift/scale2 table, mbarrier (+pad), row validity
constexpr int NUM_WARPS = 8;
constexpr int BYTES = NUM_WARPS * (WARP_BYTES + WARP_AUX_BYTES);

struct Geo {
  int on;              // 1: this launch uses the TMA epilogue
  int ex, ey, ez;      // pixel box of one lane quarter: ex * ey * ez == 32 (w, h, image extents)
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// Pixel-box corner of one lane quarter of an M tile, and what the quarter has to do.
struct Quarter {
  int x, y, z;         // output coordinates (w, h, image) of accumulator row 32 * quarter
  int active;          // 0: the quarter holds no row of the tile (tile with < 128 rows, phantom tile): nothing is read or written
};

// One epilogue warp's persistent state (operand pipeline across chunks AND tiles).
// Flavours with BOTH operands keep one request in flight (8 KB per warp): both slots and the first mbarrier.  Flavours with ONE
// operand (forward BN + residual + ReLU: the bulk of the HBM-bound 1x1 time) would have 4 KB per warp = 32 KB per SM in flight,
// ~2.9 TB/s of operand stream at DRAM latency; they use the idle second slot as a second buffer with its own mbarrier and keep TWO
// requests in flight (`head` = slot of the next item to consume, `requested` = items requested and not yet consumed, 0..2).
struct WarpState {
  uint32_t stage, slot_add, slot_gate, tab, bar;   // shared-memory addresses; the second mbarrier sits at bar + 8
  uint32_t phase;                                   // parity of the operand barrier for the NEXT wait (two-operand flavours)
  uint32_t ph0, ph1;                                // the same per slot (one-operand flavours)
  int head;
  int requested;                                    // operand boxes of the next item(s) to process have been requested
};

template <bool ADD, bool GATE>
__device__ __forceinline__ void request_operands(const CUtensorMap* tm_add, const CUtensorMap* tm_gate, const WarpState& w,
                                                 int c0, const Quarter& q) {
  mbar_expect_tx_u32(w.bar, (uint32_t)((ADD ? 1 : 0) + (GATE ? 1 : 0)) * TILE_BYTES);
  if (ADD) tma_load_tile(w.slot_add, tm_add, w.bar, c0, q.x, q.y, q.z);
  if (GATE) tma_load_tile(w.slot_gate, tm_gate, w.bar, c0, q.x, q.y, q.z);
}

// one-operand flavours: the box of one item into slot `k` (0: slot_add, 1: slot_gate), completing on that slot's mbarrier
__device__ __forceinline__ void request_single(const CUtensorMap* tm, const WarpState& w, int k, int c0, const Quarter& q) {
  const uint32_t bar = w.bar + 8u * (uint32_t)k;
  mbar_expect_tx_u32(bar, (uint32_t)TILE_BYTES);
  tma_load_tile(k ? w.slot_gate : w.slot_add, tm, bar, c0, q.x, q.y, q.z);
}

// Drains this warp's chunks (32-column blocks `half`, `half + 2`, ... of the N tile starting at channel n0) of one accumulator.
//   taddr      TMEM address of the accumulator with the warp's lane quarter applied
//   q          pixel box of the quarter for THIS tile; qn / n0n / have_next: the same for the next tile this CTA processes
//   release()  called once, when the accumulator has been read completely
template <bool ADD, bool GATE, bool STATS, class Release>
__device__ __forceinline__ void drain_tile_t(const epi::Params& p, const CUtensorMap* tm_d, const CUtensorMap* tm_add,
                                             const CUtensorMap* tm_gate, WarpState& w, uint32_t taddr, int block_n, int n0,
                                             const Quarter& q, bool have_next, int n0n, const Quarter& qn, int lane, int half,
                                             int stat_row, Release release) {
  constexpr bool OPS = ADD || GATE;
  const int nchunks = block_n / 32;
  // this warp's chunks with at least one real channel
  int cnt = 0;
  for (int ch = half; ch < nchunks && n0 + ch * 32 < p.nb; ch += 2) ++cnt;
  if (cnt == 0 || !q.active) {
    release();
    if (STATS && q.active == 0 && stat_row >= 0) {      // rows of the statistics buffer are summed unconditionally: write zeros
      for (int ch = half; ch < nchunks && n0 + ch * 32 < p.nb; ch += 2) {
        const int c = n0 + ch * 32 + lane;
        if (c < p.nb) {
          p.stats[(long long)stat_row * 2 * p.ld_stats + c] = 0.f;
          p.stats[((long long)stat_row * 2 + 1) * p.ld_stats + c] = 0.f;
        }
      }
    }
    return;
  }
  const uint32_t swz = (uint32_t)(lane & 7);
  const uint32_t row_off = (uint32_t)lane * 128u;
  const int sub_r = lane >> 3, sub_c4 = lane & 7;
  constexpr bool DUAL = (ADD != GATE) && !STATS;   // one operand: two slots, two requests in flight (see WarpState)
  const CUtensorMap* tm_one = ADD ? tm_add : tm_gate;
  int next_cnt = 0;
  if (have_next && qn.active)
    for (int ch = half; ch < nchunks && n0n + ch * 32 < p.nb; ch += 2) ++next_cnt;
  // item k of this warp, counted from the first chunk of this tile (k >= cnt: chunks of the next tile)
  auto request_item = [&](int k, int slot) {
    if (k < cnt) request_single(tm_one, w, slot, n0 + (half + 2 * k) * 32, q);
    else request_single(tm_one, w, slot, n0n + (half + 2 * (k - cnt)) * 32, qn);
  };
  if (DUAL) {
    // top up to two requests: first item of the kernel, after a skipped tile, or a predecessor tile with a single chunk
    while (w.requested < 2 && w.requested < cnt + next_cnt) {
      if (lane == 0) request_item(w.requested, (w.head + w.requested) & 1);
      ++w.requested;
    }
  } else if (OPS && !w.requested) {              // first item of the kernel (or after a tile this warp skipped)
    if (lane == 0) request_operands<ADD, GATE>(tm_add, tm_gate, w, n0 + half * 32, q);
    w.requested = 1;
  }
#pragma unroll 1
  for (int i = 0; i < cnt; ++i) {
    const int ch = half + 2 * i;
    const int c0 = n0 + ch * 32;
    uint32_t r[32];
    tc::tmem_ld_x32(taddr + ch * 32, r);
    // per-channel epilogue constants of the chunk -> the warp's table (lane = channel); read back as broadcast float4s
    {
      const int c = c0 + lane;
      const bool ok = c < p.nb;
      if (p.scale) sts32(w.tab + lane * 4, ok ? __ldg(p.scale + c) : 1.f);
      if (p.shift) sts32(w.tab + 128 + lane * 4, ok ? __ldg(p.shift + c) : 0.f);
      if (p.scale2) sts32(w.tab + 256 + lane * 4, ok ? __ldg(p.scale2 + c) : 1.f);
    }
    tc::tmem_ld_wait();
    if (i == cnt - 1) release();
    if (lane == 0) bulk_wait_read0();            // the previous tile's store has read the staging tile
    uint32_t slot_cur_add = w.slot_add, slot_cur_gate = w.slot_gate;
    if (DUAL) {
      mbar_wait_u32(w.bar + 8u * (uint32_t)w.head, w.head ? w.ph1 : w.ph0);
      if (w.head) w.ph1 ^= 1; else w.ph0 ^= 1;
      slot_cur_add = slot_cur_gate = w.head ? w.slot_gate : w.slot_add;
    } else if (OPS) { mbar_wait_u32(w.bar, w.phase); w.phase ^= 1; }
    __syncwarp();
    const float relu_floor = p.relu ? 0.f : -3.402823466e38f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t off = row_off + (((uint32_t)j ^ swz) << 4);
      float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                             __uint_as_float(r[4 * j + 3]));
      if (p.scale) {
        const float4 s = epi::lds128(w.tab + j * 16);
        if (p.shift) {                            // folded BatchNorm: one fused multiply-add, as in the register path
          const float4 t = epi::lds128(w.tab + 128 + j * 16);
          v.x = fmaf(v.x, s.x, t.x); v.y = fmaf(v.y, s.y, t.y); v.z = fmaf(v.z, s.z, t.z); v.w = fmaf(v.w, s.w, t.w);
        } else { v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w; }
      } else if (p.shift) {
        const float4 t = epi::lds128(w.tab + 128 + j * 16); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      if (ADD) { const float4 a = epi::lds128(slot_cur_add + off); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
      v.x = fmaxf(v.x, relu_floor); v.y = fmaxf(v.y, relu_floor); v.z = fmaxf(v.z, relu_floor); v.w = fmaxf(v.w, relu_floor);
      if (GATE) {
        const float4 g = epi::lds128(slot_cur_gate + off);
        v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f; v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
      }
      if (p.scale2) { const float4 s = epi::lds128(w.tab + 256 + j * 16); v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w; }
      epi::sts128(w.stage + off, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
    }
    tc::fence_proxy_async();                     // generic-proxy writes of the staging tile -> visible to the TMA unit
    __syncwarp();
    if (STATS) {
      // column sums over the quarter's 32 rows of o and o * gate: lane = (row group sub_r, 4 columns sub_c4)
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t rr = (uint32_t)(sub_r + 4 * k);
        const uint32_t off = rr * 128u + (((uint32_t)sub_c4 ^ (rr & 7u)) << 4);
        const float4 o = epi::lds128(w.stage + off);
        const float4 y = epi::lds128(w.slot_gate + off);
        a0.x += o.x; a0.y += o.y; a0.z += o.z; a0.w += o.w;
        a1.x = fmaf(o.x, y.x, a1.x); a1.y = fmaf(o.y, y.y, a1.y); a1.z = fmaf(o.z, y.z, a1.z); a1.w = fmaf(o.w, y.w, a1.w);
      }
      float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
      }
      const int c = c0 + sub_c4 * 4;
      if (sub_r == 0 && stat_row >= 0 && c < p.nb) {
        float* srow = p.stats + (long long)stat_row * 2 * p.ld_stats + c;
        *reinterpret_cast<float4*>(srow) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(srow + p.ld_stats) = make_float4(v[4], v[5], v[6], v[7]);
      }
      tc::fence_proxy_async();
      __syncwarp();                              // the gate slot may be refilled now
    }
    if (DUAL) {
      // item i has been consumed (all lanes have read the slot: the __syncwarp after the staging writes): its slot takes the
      // item two ahead, i.e. item i + 1 + (requests still in flight)
      --w.requested;
      const int nxt = i + 1 + w.requested;
      if (nxt < cnt + next_cnt) {
        if (lane == 0) request_item(nxt, w.head);
        ++w.requested;
      }
      w.head ^= 1;
    }
    if (lane == 0) {
      if (OPS && !DUAL) {                        // operand slots are free: request the next item's boxes
        if (i + 1 < cnt) request_operands<ADD, GATE>(tm_add, tm_gate, w, c0 + 64, q);
        else if (next_cnt > 0) request_operands<ADD, GATE>(tm_add, tm_gate, w, n0n + half * 32, qn);
      }
      if (p.accumulate) tma_reduce_add_4d(tm_d, w.stage, c0, q.x, q.y, q.z);
      else tma_store_4d(tm_d, w.stage, c0, q.x, q.y, q.z);
      bulk_commit();
    }
  }
  if (OPS && !DUAL) w.requested = next_cnt > 0 ? 1 : 0;
}

template <class Release>
__device__ __forceinline__ void drain_tile(const epi::Params& p, const CUtensorMap* tm_d, const CUtensorMap* tm_add,
                                           const CUtensorMap* tm_gate, WarpState& w, uint32_t taddr, int block_n, int n0,
                                           const Quarter& q, bool have_next, int n0n, const Quarter& qn, int lane, int half,
                                           int stat_row, Release release) {
  if (p.stats) {                      // statistics imply a gate (host-checked)
    if (p.addend) drain_tile_t<true, true, true>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
    else drain_tile_t<false, true, true>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
  } else if (p.addend) {
    if (p.gate) drain_tile_t<true, true, false>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
    else drain_tile_t<true, false, false>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
  } else {
    if (p.gate) drain_tile_t<false, true, false>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
    else drain_tile_t<false, false, false>(p, tm_d, tm_add, tm_gate, w, taddr, block_n, n0, q, have_next, n0n, qn, lane, half, stat_row, release);
  }
}

}  // namespace epi2
