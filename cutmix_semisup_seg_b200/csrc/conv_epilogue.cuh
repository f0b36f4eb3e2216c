// Shared epilogue of the convolution GEMM kernels: TMEM accumulator (128 rows x block_n columns per CTA) -> HBM.
//
// Design notes (from the ncu captures and the clock64 pipeline traces under profiles/): with two TMEM accumulators the
// MMA of tile i+2 waits for the epilogue of tile i, so for the small-K 1x1 layers (4k MMA cycles per tile) the
// epilogue IS the kernel.  v1 ran ~2200 SASS instructions per 32-column chunk, v2 ~590 (64-bit pixel offsets, a
// stack copy of the parameter block read back with LDL, every optional term executed) and measured 14k cycles per tile
// even with loads and stores disabled, i.e. issue/latency-bound.  This version
//   * reads its parameters straight from the kernel's __grid_constant__ block (uniform loads, no stack frame),
//   * keeps one 32-bit pixel index per row; an address is one IMAD.WIDE,
//   * is specialised on <residual addend, gate> so that absent terms cost no instructions,
//   * issues the residual and gate loads of all 8 row groups of a chunk before touching them (16 independent 16 B
//     loads in flight per lane),
//   * keeps the rare paths (channel count not a multiple of 4, unaligned pointers) out of line.
// Rows are staged through an XOR-swizzled smem tile so that 8 consecutive lanes own 32 consecutive channels of one pixel:
// loads and stores are full 128 B segments.
#pragma once
#include "tc_common.cuh"

namespace epi {

// explicit shared-space 128-bit accesses (a generic float* compiles to ST.E / LD.E with 64-bit addressing)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Ampere-style asynchronous 16 B global -> shared copies (LDGSTS): no destination register, so a lane can keep many
// more bytes in flight than its register budget allows; completion is tracked per thread with commit / wait groups.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Asynchronous operand prefetch (kernels built with the PF layout): the residual addend / ReLU gate values a lane needs
// for its NEXT chunk are copied into a per-lane shared-memory slot while it works on the current chunk.  Each lane reads
// back exactly the 16 B pieces it copied itself (entry i of lane l at (i*32 + l)*16), so no cross-lane synchronisation is
// involved; a piece is re-issued for the next chunk right after its current value has been consumed.  Two commit groups
// per chunk (row groups 0-3 / 4-7), always `wait_group 1`.  16 pieces x 16 B in flight per lane = 64 KB per SM, against
// ~16 KB for the register-limited direct loads (profiles/r01_v6_l3_1x1_insitu_flavours*: the HBM-bound 1x1 layers ran at
// half the bandwidth they need).
constexpr int PF_OPERAND_BYTES = 8 * 32 * 16;       // 8 row groups x 32 lanes x 16 B
constexpr int PF_WARP_BYTES = 2 * PF_OPERAND_BYTES; // addend + gate

constexpr int ROW_FLOATS = 32;                      // 128 B rows, 16 B chunk j of row r stored at chunk j ^ (r & 7): conflict-free
                                                    // 128-bit accesses without padding (the 4 KB saved pay for a sixth operand stage)
constexpr int WARP_BYTES = 32 * ROW_FLOATS * 4;
constexpr int NUM_WARPS = 8;                        // two warps per TMEM lane quarter (even / odd 32-column chunks)
constexpr int BYTES = NUM_WARPS * WARP_BYTES + NUM_WARPS * 32 * 4;  // staging tiles + per-row pixel indices
constexpr int PF_BYTES = NUM_WARPS * PF_WARP_BYTES;                // prefetch slots (PF kernels only)

struct Params {
  float* d; int ldd;
  const float* scale; const float* shift; const float* scale2;
  const float* addend; int ld_add;
  const float* gate; int ld_gate;
  int relu, accumulate, vec_ok, nb;
  // Optional fused column statistics of the stored gradient (frozen-BN parameter gradients, see b200seg.h):
  // stats[(row_block*2 + j)*ld_stats + ch], j = 0: sum v, j = 1: sum v*(gate - sub); row_block = m_tile*4 + lane quarter
  float* stats; int ld_stats;
  const float* sub; int ld_sub;
  int dbg;          // debug knob 3 (timing experiments only): low bits 1 = skip HBM stores, 2 = also skip the TMEM
                    // loads; bit 2 (value 4) = L2-prefetch the next tile's epilogue operands (measured slower: see below)
};

// EXPERIMENT (off by default, debug knob 3 bit 2): L2 prefetch of the epilogue's read operands (residual addend / ReLU gate / statistics `sub`) for one output row of an
// upcoming tile: `cols` channels starting at `c0` (one contiguous run per row).  The loads themselves are issued only
// after the accumulator is ready, 4-8 row groups at a time (~16-32 KB in flight per SM, far below the ~90 KB that
// HBM latency x 1/148 of HBM bandwidth needs), so without this the HBM-bound 1x1 layers run latency-bound; L2
// prefetches (one per 128 B line; no registers, no smem) should turn those loads into L2 hits.  Measured in situ (run 19/20,
// profiles/): both this and cp.async.bulk.prefetch.L2 made the k256->n1024 layers 14-16% SLOWER, so it is disabled.
__device__ __forceinline__ void prefetch_lines(const float* p, int floats) {
  for (int o = 0; o < floats; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}
__device__ __forceinline__ void prefetch_row(const Params& p, int pix, int c0, int cols) {
  if (pix < 0) return;
  int n = p.nb - c0; if (n > cols) n = cols;
  if (n <= 0) return;
  if (p.addend) prefetch_lines(p.addend + (long long)pix * p.ld_add + c0, n);
  if (p.gate) prefetch_lines(p.gate + (long long)pix * p.ld_gate + c0, n);
  if (p.stats && p.sub) prefetch_lines(p.sub + (long long)pix * p.ld_sub + c0, n);
  if (p.accumulate) prefetch_lines(p.d + (long long)pix * p.ldd + c0, n);
}

// Issue the asynchronous copies of row groups [4b, 4b+4) of the chunk whose first column (for this lane) is `c` into the
// slot regions `sa` (addend) / `sg` (gate), and close the commit group (always, so every lane counts the same groups).
template <bool ADD, bool GATE>
__device__ __forceinline__ void pf_issue(const Params& p, const int (&od)[8], int c, bool lane_ok, int b, uint32_t sa, uint32_t sg, int lane) {
  if (lane_ok) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = b * 4 + j;
      if (od[i] >= 0) {
        if (ADD) cp_async16(sa + (i * 32 + lane) * 16, p.addend + (long long)od[i] * p.ld_add + c);
        if (GATE) cp_async16(sg + (i * 32 + lane) * 16, p.gate + (long long)od[i] * p.ld_gate + c);
      }
    }
  }
  cp_async_commit();
}
// Slot regions of the q-th chunk a warp processes in a tile: with both operands each owns one region (prefetch distance
// 1 chunk); a single operand alternates between the two regions (distance 2), so 16 pieces stay in flight either way.
template <bool ADD, bool GATE>
__device__ __forceinline__ void pf_regions(uint32_t slot, int q, uint32_t* sa, uint32_t* sg) {
  if (ADD && GATE) { *sa = slot; *sg = slot + PF_OPERAND_BYTES; }
  else { *sa = *sg = slot + (q & 1) * PF_OPERAND_BYTES; }
}
template <bool ADD, bool GATE>
__device__ __forceinline__ void pf_prologue_t(const Params& p, int block_n, int n0, const int (&od)[8], int lane, int half, uint32_t slot) {
  constexpr int DIST = (ADD && GATE) ? 1 : 2;
  const int sub_c = (lane & 7) * 4;
#pragma unroll
  for (int q = 0; q < DIST; ++q) {
    const int ch = half + 2 * q;
    const int c = n0 + ch * 32 + sub_c;
    const bool ok = ch < block_n / 32 && p.vec_ok && c + 3 < p.nb;
    uint32_t sa, sg;
    pf_regions<ADD, GATE>(slot, q, &sa, &sg);
    pf_issue<ADD, GATE>(p, od, c, ok, 0, sa, sg, lane);
    pf_issue<ADD, GATE>(p, od, c, ok, 1, sa, sg, lane);
  }
}
// Tile prologue of a PF kernel: start the copies for this warp's first chunk(s) (called before waiting for the accumulator).
__device__ __forceinline__ void pf_prologue(const Params& p, int block_n, int n0, const int* rowpix, int lane, int half, uint32_t slot) {
  if (!(p.addend || p.gate)) return;
  if (half >= block_n / 32 || n0 + half * 32 >= p.nb) return;          // this warp has no chunk in the tile
  const int sub_r = lane >> 3;
  int od[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) od[i] = rowpix[i * 4 + sub_r];
  if (p.addend) { if (p.gate) pf_prologue_t<true, true>(p, block_n, n0, od, lane, half, slot); else pf_prologue_t<true, false>(p, block_n, n0, od, lane, half, slot); }
  else pf_prologue_t<false, true>(p, block_n, n0, od, lane, half, slot);
}

// Out-of-line general path for one lane's 4 columns of one row: any channel count / alignment, accumulate.
static __device__ __noinline__ void slow_store(const Params& p, float4 v, int pix, int c) {
  const float vv[4] = {v.x, v.y, v.z, v.w};
  float* drow = p.d + (long long)pix * p.ldd;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int cc = c + e;
    if (cc < p.nb) {
      float o = vv[e];
      if (p.scale) o *= __ldg(p.scale + cc);
      if (p.shift) o += __ldg(p.shift + cc);
      if (p.addend) o += __ldg(p.addend + (long long)pix * p.ld_add + cc);
      if (p.relu) o = fmaxf(o, 0.f);
      if (p.gate) o = __ldg(p.gate + (long long)pix * p.ld_gate + cc) > 0.f ? o : 0.f;
      if (p.scale2) o *= __ldg(p.scale2 + cc);
      if (p.accumulate) o += drow[cc];
      drow[cc] = o;
    }
  }
}

// One epilogue warp drains lanes [32*ew, 32*ew+32) of the accumulator at TMEM address `taddr` (lane quarter already
// applied).  rowpix[32]: output pixel index of each of the warp's rows (-1 = row not stored).  `release()` is called
// once the accumulator has been completely read (so the MMA warp may overwrite it).
// `half` (0/1) selects the even or odd chunks: two warps share a lane quarter.
template <bool ADD, bool GATE, bool STATS, class Release>
__device__ __forceinline__ void drain_tile_t(const Params& p, uint32_t taddr, int block_n, int n0, float* stg,
                                             const int* rowpix, int lane, int half, int stat_row, uint32_t pf_slot,
                                             Release release) {
  const bool pf = (ADD || GATE) && pf_slot != 0;       // operands arrive through the asynchronous prefetch slots
  constexpr int PF_DIST = (ADD && GATE) ? 1 : 2;
  const int sub_r = lane >> 3;          // row within a group of 4
  const int sub_c = (lane & 7) * 4;     // first of this lane's 4 columns inside a chunk
  int od[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) od[i] = rowpix[i * 4 + sub_r];
  const float relu_floor = p.relu ? 0.f : -3.402823466e38f;
  const bool fast = p.vec_ok;
  const bool acc = p.accumulate != 0;
  const int nchunks = block_n / 32;
  const uint32_t st_w = tc::smem_u32(stg + lane * ROW_FLOATS);          // this lane's row; chunk q goes to slot q ^ (lane & 7)
  const uint32_t st_swz = (uint32_t)(lane & 7);
  // read side: row i * 4 + sub_r, chunk lane & 7 -> slot (lane & 7) ^ (row & 7), row & 7 = (i & 1) * 4 + sub_r
  const uint32_t st_r0 = tc::smem_u32(stg + sub_r * ROW_FLOATS) + (((uint32_t)(lane & 7) ^ (uint32_t)sub_r) << 4);
  const uint32_t st_r1 = tc::smem_u32(stg + sub_r * ROW_FLOATS) + (((uint32_t)(lane & 7) ^ (uint32_t)(sub_r + 4)) << 4);
  if (half >= nchunks) release();       // nothing to read for this warp: still owes its arrival
#pragma unroll 1                        // keep the body resident in the instruction cache (4 specialisations x 2 kernels)
  for (int ch = half; ch < nchunks; ch += 2) {
    uint32_t r[32];
    if ((p.dbg & 3) < 2) {
      tc::tmem_ld_x32(taddr + ch * 32, r);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) r[q] = 0;
    }
    if (ch + 2 >= nchunks) release();
    const int col0 = n0 + ch * 32;
    if (col0 >= p.nb) continue;
    const int c = col0 + sub_c;
    const bool lane_fast = fast && c + 3 < p.nb;
    uint32_t pf_sa = 0, pf_sg = 0;
    if (pf) pf_regions<ADD, GATE>(pf_slot, (ch - half) >> 1, &pf_sa, &pf_sg);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;     // STATS: this lane's column sums over its 8 rows
#pragma unroll
    for (int q = 0; q < 8; ++q)
      sts128(st_w + (((uint32_t)q ^ st_swz) << 4), r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
    __syncwarp();
    if (lane_fast) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), s2 = sc;
      if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + c));
      if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + c));
      if (p.scale2) s2 = __ldg(reinterpret_cast<const float4*>(p.scale2 + c));
      // two batches of 4 row groups: up to 8 independent 16 B residual / gate loads in flight per lane without
      // spilling (all 16 at once needs > 168 registers)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float4 ad[4], gt[4], sb[4];
        if (STATS && p.sub) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = b * 4 + j;
            sb[j] = od[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.sub + (long long)od[i] * p.ld_sub + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (pf) {
          cp_async_wait<2 * PF_DIST - 1>();   // this batch's group has landed (only the later ones may still be pending)
          if (ADD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = b * 4 + j;
              ad[j] = od[i] >= 0 ? lds128(pf_sa + (i * 32 + lane) * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (GATE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = b * 4 + j;
              gt[j] = od[i] >= 0 ? lds128(pf_sg + (i * 32 + lane) * 16) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
          }
        } else {
          if (ADD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = b * 4 + j;
              ad[j] = od[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.addend + (long long)od[i] * p.ld_add + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (GATE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = b * 4 + j;
              gt[j] = od[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.gate + (long long)od[i] * p.ld_gate + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = b * 4 + j;
          const float4 v = lds128(((i & 1) ? st_r1 : st_r0) + i * 4 * ROW_FLOATS * 4);
          float4 o;
          o.x = fmaf(v.x, sc.x, sh.x); o.y = fmaf(v.y, sc.y, sh.y); o.z = fmaf(v.z, sc.z, sh.z); o.w = fmaf(v.w, sc.w, sh.w);
          if (ADD) { o.x += ad[j].x; o.y += ad[j].y; o.z += ad[j].z; o.w += ad[j].w; }
          o.x = fmaxf(o.x, relu_floor); o.y = fmaxf(o.y, relu_floor); o.z = fmaxf(o.z, relu_floor); o.w = fmaxf(o.w, relu_floor);
          if (GATE) {
            o.x = gt[j].x > 0.f ? o.x : 0.f; o.y = gt[j].y > 0.f ? o.y : 0.f; o.z = gt[j].z > 0.f ? o.z : 0.f; o.w = gt[j].w > 0.f ? o.w : 0.f;
          }
          o.x *= s2.x; o.y *= s2.y; o.z *= s2.z; o.w *= s2.w;
          if (STATS && od[i] >= 0) {
            float4 yv = gt[j];
            if (p.sub) { yv.x -= sb[j].x; yv.y -= sb[j].y; yv.z -= sb[j].z; yv.w -= sb[j].w; }
            a0.x += o.x; a0.y += o.y; a0.z += o.z; a0.w += o.w;
            a1.x = fmaf(o.x, yv.x, a1.x); a1.y = fmaf(o.y, yv.y, a1.y); a1.z = fmaf(o.z, yv.z, a1.z); a1.w = fmaf(o.w, yv.w, a1.w);
          }
          if (od[i] >= 0 && !(p.dbg & 3)) {
            float4* dst = reinterpret_cast<float4*>(p.d + (long long)od[i] * p.ldd + c);
            if (acc) {                    // several dgrads summing into one input gradient (ASPP branches, phases)
              const float4 old = *dst;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *dst = o;
          }
        }
        if (pf) {
          // the values of this batch have been consumed: re-use their slot entries for the chunk PF_DIST ahead
          const int nch = ch + 2 * PF_DIST;
          const int ncol = n0 + nch * 32 + sub_c;
          const bool more = nch < nchunks && ncol + 3 < p.nb;
          pf_issue<ADD, GATE>(p, od, ncol, more, b, pf_sa, pf_sg, lane);
        }
      }
    } else if (c < p.nb) {
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        if (od[i] < 0) continue;
        slow_store(p, lds128(((i & 1) ? st_r1 : st_r0) + i * 4 * ROW_FLOATS * 4), od[i], c);
      }
    }
    if (STATS) {
      // fixed-order reduction over the 4 row groups (lanes differing in bits 3, 4): deterministic partials
      float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
        v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
      }
      if (sub_r == 0 && stat_row >= 0 && c < p.nb) {
        float* srow = p.stats + (long long)stat_row * 2 * p.ld_stats + c;
        *reinterpret_cast<float4*>(srow) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(srow + p.ld_stats) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncwarp();
  }
}

template <class Release>
__device__ __forceinline__ void drain_tile(const Params& p, uint32_t taddr, int block_n, int n0, float* stg,
                                           const int* rowpix, int lane, int half, int stat_row, uint32_t pf_slot, Release release) {
  // the specialisations are selected once per tile (uniform branch); stats imply a gate (host-checked)
  if (p.stats) {
    if (p.addend) drain_tile_t<true, true, true>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
    else drain_tile_t<false, true, true>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
  } else if (p.addend) {
    if (p.gate) drain_tile_t<true, true, false>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
    else drain_tile_t<true, false, false>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
  } else {
    if (p.gate) drain_tile_t<false, true, false>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
    else drain_tile_t<false, false, false>(p, taddr, block_n, n0, stg, rowpix, lane, half, stat_row, pf_slot, release);
  }
}

}  // namespace epi
