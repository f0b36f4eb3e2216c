// 2-CTA variant of the implicit-GEMM convolution (see conv_gemm.cu for the single-CTA kernel and the GEMM view).
//
// Why: with kind::tf32 a 128x256x8 MMA reads 12 KB of operands from shared memory while TMA writes the same 12 KB,
// i.e. ~192 B/cycle against a ~128 B/cycle shared-memory port: the single-CTA kernel saturates at ~63 % tensor-pipe
// (ncu, profiles/r01_v1_conv_gemm_aspp_d12.txt; the MMA issuer never waits for data, the producer waits for slots).
// Here the two CTAs of a cluster (one SM pair) execute ONE tcgen05.mma.cta_group::2 of shape M256 x N256 x K8: each CTA
// holds its own 128 pixel rows of A and HALF of the weight tile (128 of the 256 output channels), so per CTA and K step
// shared memory sees 8 KB of reads + 8 KB of TMA writes, and L2 -> SM traffic per MAC drops by a third.
//
// Protocol (leader = cluster rank 0):
//   * both CTAs run a TMA producer; their loads carry .cta_group::2 and complete on the LEADER's full barrier, which the
//     leader arms with the byte count of both CTAs;
//   * only the leader issues MMAs; tcgen05.commit.cta_group::2 ... multicast::cluster releases the smem stage in both
//     CTAs and publishes the finished accumulator to both epilogues;
//   * each CTA's TMEM holds its 128 rows of the 256-row accumulator; both epilogues arrive (the peer remotely) on the
//     leader's "accumulator empty" barrier.
#include <stdlib.h>
#include "tc_common.cuh"
#include "conv_epilogue.cuh"
#include "conv_epilogue_tma.cuh"

namespace {

constexpr int BLOCK_M = 128;            // per CTA; the pair computes 256 rows
constexpr int BLOCK_K = 32;
constexpr int BLOCK_N = 256;
// Two builds of the kernel: the compute-bound one keeps a 5-stage operand ring; the PF one (HBM-bound small-K layers
// whose epilogue reads a residual addend / ReLU gate) trades two stages for the epilogue's asynchronous prefetch slots
// (conv_epilogue.cuh).
constexpr int STAGES_MAIN = 6;          // 6 x 32 KB + swizzled (unpadded) epilogue staging = 226.25 KB
constexpr int STAGES_MAIN_OLD = 5;      // A/B timing (debug knob 10)
constexpr int STAGES_PF = 3;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;            // 16 KB
constexpr int B_STAGE_BYTES = (BLOCK_N / 2) * BLOCK_K * 4;      // 16 KB: this CTA's half of the weight tile
constexpr int EPI_BYTES = epi::BYTES;
// PF build: the region behind the barriers holds EITHER the register epilogue's staging tiles + cp.async slots OR the TMA
// epilogue's per-warp tiles (1024 B aligned: they start 1 KB behind the operand ring) + tables
constexpr int PF_REGION = EPI_BYTES + epi::PF_BYTES > 768 + epi2::BYTES ? EPI_BYTES + epi::PF_BYTES : 768 + epi2::BYTES;
constexpr int smem_bytes(int stages, bool pf) {
  return stages * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256 + (pf ? PF_REGION : EPI_BYTES);
}
static_assert(smem_bytes(STAGES_MAIN, false) <= 232448 && smem_bytes(STAGES_PF, true) <= 232448, "shared memory budget");
constexpr int MAX_TAPS = 16;
constexpr int NUM_THREADS = 384;      // 4 control warps + 8 epilogue warps
constexpr int TMEM_COLS = 512;

struct Conv2KArgs {
  int n, ih, iw, k;
  int nb;
  int oh, ow;
  int fh, fw, ldd, ostride, ooh, oow;
  int istride;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n;
  int num_pairs;              // (pairs of M tiles) x (N tiles)
  int n_taps;
  short dh[MAX_TAPS], dw[MAX_TAPS], btap[MAX_TAPS];
  int kblocks;
  int n_pass;
  int tap_outer;             // 0: K-block outer / tap inner (default, see the producer); 1: tap outer (debug knob 7)
  int reverse;               // 1: walk the tiles from the last to the first (alternating launch directions, conv_gemm.cu)
  tc::FastDiv fd_w, fd_h, fd_nn;   // division by tiles_w, tiles_h, n_tiles_n (tc_common.cuh "cheap tile decoding")
  tc::TapTables tt;
  int wide_pf;               // 1: L2 prefetch of the A stream in 256-channel boxes (1x1 layers with K >= 512, see the producer)
  epi::Params ep;             // epilogue parameter block (conv_epilogue.cuh)
  epi2::Geo tma;              // TMA epilogue (conv_epilogue_tma.cuh; PF build only)
  long long* trace;   // diagnostics (b2_debug_trace): clock64 stamps of CTA 0's pipeline roles, 4 x 512 slots
};

struct TileInfo {
  int n_idx, w0, h0, n0;
  int m_idx;                  // index of this CTA's M tile (row of the fused column statistics)
  uint32_t tap_mask;
};

__host__ __device__ __forceinline__ uint32_t tile_tap_mask(const Conv2KArgs& a, int m) {
  const int q = (int)tc::fast_div((uint32_t)m, a.fd_w);
  const int wt = m - q * a.tiles_w;
  const int nt = (int)tc::fast_div((uint32_t)q, a.fd_h);
  const int ht = q - nt * a.tiles_h;
  if (nt >= a.tiles_n) return 0;       // phantom tile of an odd tile count
  if (a.tt.on) return (uint32_t)a.tt.rows[ht] & (uint32_t)a.tt.cols[wt];
  const int w0 = wt * a.bw, h0 = ht * a.bh;
  uint32_t mask = 0;
  for (int i = 0; i < a.n_taps; ++i) {
    const int lo_h = h0 * a.istride + a.dh[i], hi_h = lo_h + (a.bh - 1) * a.istride;
    const int lo_w = w0 * a.istride + a.dw[i], hi_w = lo_w + (a.bw - 1) * a.istride;
    if (hi_h >= 0 && lo_h < a.ih && hi_w >= 0 && lo_w < a.iw) mask |= 1u << i;
  }
  return mask;
}

// `pair` indexes (pair of adjacent M tiles, N tile), N fastest; `half` selects this CTA's M tile.  The tap mask is the
// union over both halves: the two producers and the single MMA issuer must walk the same K sequence.
__host__ __device__ __forceinline__ TileInfo decode_tile(const Conv2KArgs& a, int pair, int half) {
  TileInfo t;
  if (a.reverse) pair = a.num_pairs - 1 - pair;
  const int mp = (int)tc::fast_div((uint32_t)pair, a.fd_nn);
  t.n_idx = pair - mp * a.n_tiles_n;
  const uint32_t both = tile_tap_mask(a, mp * 2) | tile_tap_mask(a, mp * 2 + 1);
  const int m = mp * 2 + half;
  t.m_idx = m;
  const int q = (int)tc::fast_div((uint32_t)m, a.fd_w);
  const int wt = m - q * a.tiles_w;
  const int nt = (int)tc::fast_div((uint32_t)q, a.fd_h);      // == tiles_n for the phantom tile: every pixel out of range
  const int ht = q - nt * a.tiles_h;
  t.w0 = wt * a.bw; t.h0 = ht * a.bh; t.n0 = nt * a.bn;
  t.tap_mask = both ? both : 1u;
  return t;
}

// ---- cta_group::2 / cluster PTX ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared-window address -> leader CTA
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs retired) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_mcast(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}
// arrive on the leader CTA's copy of `bar` (local when this CTA is the leader).  Default (.release.cta) semantics, as
// CUTLASS' ClusterBarrier::arrive(cta_id): the TMEM reads it orders are fenced by tcgen05.fence::before_thread_sync.
// (A `.release.cluster` arrive compiles to MEMBAR.ALL.GPU, which stalls the warp until every earlier output store has
// been acknowledged by L2 -- ~8k cycles per tile in the pipeline trace of the 1x1 layers.)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(tc::smem_u32(bar)));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int STAGES, bool PF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                  const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAdd,
                  const __grid_constant__ CUtensorMap tmGate, const __grid_constant__ CUtensorMap tmPf,
                  const __grid_constant__ Conv2KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                        // [STAGES]  (leader's copies are the live ones)
  uint64_t* empty_bar = bars + STAGES;              // [STAGES]  per CTA
  uint64_t* tfull_bar = bars + 2 * STAGES;          // [2]       per CTA
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;     // [2]       leader's copies are the live ones
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = (int)tc::uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;     // uniform: role branches converge
  const uint32_t rank = tc::uniform(cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmB);
    if (a.n_pass > 1) { tc::tma_prefetch_desc(&tmAlo); tc::tma_prefetch_desc(&tmBlo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 16); }   // 8 warps x 2 CTAs
    if (PF && a.tma.on) {
      uint8_t* aux = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + epi2::NUM_WARPS * epi2::WARP_BYTES;
      for (int i = 0; i < epi2::NUM_WARPS; ++i)
        for (int k = 0; k < 2; ++k)               // one operand barrier per slot (conv_epilogue_tma.cuh, WarpState)
          tc::mbar_init(reinterpret_cast<uint64_t*>(aux + i * epi2::WARP_AUX_BYTES + 384 + 8 * k), 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tmem2_alloc(tmem_ptr, TMEM_COLS);
    tmem2_relinquish();
  }
  tc::tc_fence_before();
  cluster_sync_all();                               // barriers of BOTH CTAs initialised, TMEM allocated
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc::pdl_wait();                 // everything above overlaps the predecessor's tail (tc_common.cuh)
  tc::pdl_launch_dependents();

  const uint32_t rows_a = a.bw * a.bh * a.bn;
  const uint32_t stage_tx = 2u * (rows_a * 128u + (uint32_t)(BLOCK_N / 2) * 128u);     // both CTAs' bytes

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, one elected lane issues) =====================
    const uint32_t el = tc::elect_one();
    int stage = 0; uint32_t phase = 0;
    int tr_n = 0;
    // Loop order: K-block OUTER, tap INNER.  All CTAs of the grid walk the K blocks roughly in step, so at any moment the
    // chip's working set is ONE 32-channel slab of the few images in flight (a few MB): the 9 shifted boxes of that slab
    // (and the neighbouring tiles' halos) are served from L2.  With the taps outside, every tap streamed all channels of
    // those images (~150 MB at 2048 channels, more than the 126 MB L2) before the next tap could reuse anything: the ASPP
    // fprop read 2.49 GB from DRAM for 0.62 GB of tensors (profiles/r01_v11_ncu_full_aspp_*).  The MMA issuer only counts
    // stages, so the order is the producer's alone.  (Debug knob 7 = 1 restores the old order for A/B timing.)
    const int n_outer = a.tap_outer ? a.n_taps : a.kblocks;
    const int n_inner = a.tap_outer ? a.kblocks : a.n_taps;
    for (int pair = cluster_id; pair < a.num_pairs; pair += num_clusters) {
      const TileInfo t = decode_tile(a, pair, (int)rank);
      const int w_base = t.w0 * a.istride, h_base = t.h0 * a.istride;
      for (int o = 0; o < n_outer; ++o) {
        if (a.wide_pf && (o & 7) == 4) {
          // 1x1 layer: the operand boxes are 128 B per pixel row at a pitch of K * 4 bytes -- 128 B DRAM bursts scattered over as many
          // pages as the tile has pixels.  Four stages before a group of 8 K blocks is needed, ONE prefetch pulls the group's 1 KB per
          // pixel row into L2 (the first group of the next tile while the last group of this one is being loaded).
          const int g = (o >> 3) + 1;
          if (g * 8 < a.kblocks) {
            if (el) tc::tma_prefetch_l2_4d(&tmPf, g * 256, w_base, h_base, t.n0);
          } else if (pair + num_clusters < a.num_pairs) {
            const TileInfo tn = decode_tile(a, pair + num_clusters, (int)rank);
            if (el) tc::tma_prefetch_l2_4d(&tmPf, 0, tn.w0 * a.istride, tn.h0 * a.istride, tn.n0);
          }
        }
        for (int i = 0; i < n_inner; ++i) {
          const int tap = a.tap_outer ? o : i;
          const int kb = a.tap_outer ? i : o;
          if (!(t.tap_mask >> tap & 1)) continue;
          const int cw = w_base + a.dw[tap];
          const int ch = h_base + a.dh[tap];
          const int bt = a.btap[tap];
          for (int p = 0; p < a.n_pass; ++p) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (a.trace && blockIdx.x == 0 && tr_n < 512) { if (el) a.trace[tr_n] = clock64(); ++tr_n; }
            if (el) {
              if (a.ep.dbg & 8) {               // timing experiment: no operand traffic (stale smem contents)
                if (leader) tc::mbar_arrive(&full_bar[stage]);
              } else {
                if (leader) tc::mbar_expect_tx(&full_bar[stage], stage_tx);
                tma2_load_4d(smem_a + stage * A_STAGE_BYTES, (p & 1) ? &tmAlo : &tmA, &full_bar[stage],
                             kb * BLOCK_K, cw, ch, t.n0);
                tma2_load_3d(smem_b + stage * B_STAGE_BYTES, (p & 2) ? &tmBlo : &tmB, &full_bar[stage],
                             kb * BLOCK_K, bt, t.n_idx * BLOCK_N + (int)rank * (BLOCK_N / 2));
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; converged warp, one elected lane issues) =====================
    if (leader) {
      const uint32_t el = tc::elect_one();
      const uint32_t idesc = tc::make_idesc_tf32(2 * BLOCK_M, BLOCK_N, 0, 0);
      // K-major SWIZZLE_128B descriptors: low word = start address >> 4 | LBO (16 B) << 16, high word constant
      // (SBO = 1024 B, version 1, swizzle mode 2).  One K step of 8 tf32 = 32 B = +2 in the low word.
      const uint32_t desc_hi = (uint32_t)(tc::make_smem_desc_sw128(0, 16, 1024) >> 32);
      const uint32_t a_lo0 = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(smem_a), 16, 1024);
      const uint32_t b_lo0 = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(smem_b), 16, 1024);
      const bool no_mma = a.ep.dbg & 16;      // (16: timing experiment without tensor-core work)
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int tr_n = 0, tr_t = 0;
      for (int pair = cluster_id; pair < a.num_pairs; pair += num_clusters) {
        const TileInfo t = decode_tile(a, pair, 0);
        if (a.trace && blockIdx.x == 0 && tr_t < 510) { if (el) a.trace[1024 + tr_t] = clock64(); ++tr_t; }
        tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        if (a.trace && blockIdx.x == 0 && tr_t < 510) { if (el) a.trace[1024 + tr_t] = clock64(); ++tr_t; }
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accum = 0;
        const int iters = __popc(t.tap_mask) * a.kblocks * a.n_pass;
        for (int it = 0; it < iters; ++it) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          if (a.trace && blockIdx.x == 0 && tr_n < 512) { if (el) a.trace[512 + tr_n] = clock64(); ++tr_n; }
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (A_STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (B_STAGE_BYTES >> 4);
          if (el) {
            if (!no_mma) {
#pragma unroll
              for (int ks = 0; ks < BLOCK_K / 8; ++ks)
                mma2_tf32(tmem_d, ((uint64_t)desc_hi << 32) | (a_lo + 2 * ks), ((uint64_t)desc_hi << 32) | (b_lo + 2 * ks), idesc,
                          ks == 0 ? accum : 1u);
            }
            mma2_commit_mcast(&empty_bar[stage]);   // frees this stage in both CTAs when the MMAs retire
          }
          accum = 1;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (el) mma2_commit_mcast(&tfull_bar[acc]);   // accumulator complete -> both epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && PF && a.tma.on) {
    // ===================== TMA epilogue (both CTAs, own 128 rows; conv_epilogue_tma.cuh) =====================
    const int ew = (warp - 4) & 3;        // TMEM lane quarter
    const int eh = (warp - 4) >> 2;       // even / odd chunks
    const int ewi = warp - 4;
    const epi::Params& ep = a.ep;
    uint8_t* tiles = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + ewi * epi2::WARP_BYTES;
    uint8_t* aux = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + epi2::NUM_WARPS * epi2::WARP_BYTES + ewi * epi2::WARP_AUX_BYTES;
    epi2::WarpState ws;
    ws.stage = tc::smem_u32(tiles); ws.slot_add = ws.stage + epi2::TILE_BYTES; ws.slot_gate = ws.stage + 2 * epi2::TILE_BYTES;
    ws.tab = tc::smem_u32(aux); ws.bar = ws.tab + 384; ws.phase = 0; ws.ph0 = ws.ph1 = 0; ws.head = 0; ws.requested = 0;
    const int lin = ew * 32;              // first accumulator row of this warp's quarter
    const int bwbh = a.bw * a.bh;
    const int q_dn = lin / bwbh, q_dh = (lin - q_dn * bwbh) / a.bw, q_dw = lin - q_dn * bwbh - q_dh * a.bw;
    auto quarter = [&](const TileInfo& ti) -> epi2::Quarter {
      epi2::Quarter q;
      q.x = ti.w0 + q_dw; q.y = ti.h0 + q_dh; q.z = ti.n0 + q_dn;
      q.active = ((uint32_t)lin < rows_a && q.x < a.ow && q.y < a.oh && q.z < a.n) ? 1 : 0;
      return q;
    };
    int acc = 0; uint32_t acc_phase = 0;
    for (int pair = cluster_id; pair < a.num_pairs; pair += num_clusters) {
      const TileInfo t = decode_tile(a, pair, (int)rank);
      const epi2::Quarter q = quarter(t);
      const bool have_next = pair + num_clusters < a.num_pairs;
      TileInfo tn = t;
      if (have_next) tn = decode_tile(a, pair + num_clusters, (int)rank);
      const epi2::Quarter qn = quarter(tn);
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
      const int m_idx = t.m_idx;
      const int stat_row = m_idx < a.tiles_w * a.tiles_h * a.tiles_n ? m_idx * 4 + ew : -1;   // phantom tile: none
      epi2::drain_tile(ep, &tmD, &tmAdd, &tmGate, ws, taddr, BLOCK_N, t.n_idx * BLOCK_N, q, have_next, tn.n_idx * BLOCK_N, qn,
                       lane, eh, stat_row, [&]() {
        tc::tc_fence_before();           // accumulator fully read: hand the TMEM stage back to the MMA warp
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) epi2::bulk_wait0();    // the last stores have left shared memory (and completed) before the CTA exits
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows; conv_epilogue.cuh) =====================
    const int ew = (warp - 4) & 3;        // TMEM lane quarter
    const int eh = (warp - 4) >> 2;       // even / odd chunks
    const int ewi = warp - 4;
    const int row = ew * 32 + lane;
    float* stg = reinterpret_cast<float*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256) + ewi * (32 * epi::ROW_FLOATS);
    int* rowpix = reinterpret_cast<int*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + epi::NUM_WARPS * epi::WARP_BYTES) + ewi * 32;
    const epi::Params& ep = a.ep;     // stays in the kernel's constant parameter space
    const uint32_t pf_slot = PF ? tc::smem_u32(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + EPI_BYTES) + ewi * epi::PF_WARP_BYTES : 0u;
    int acc = 0; uint32_t acc_phase = 0;
    const int bwbh = a.bw * a.bh;
    int tr_e = 0;
    const bool tracer = a.trace && blockIdx.x == 0 && warp == 4 && lane == 0;
    const int dn = row / bwbh;
    const int rem = row - dn * bwbh;
    const int dhh = rem / a.bw;
    const int dww = rem - dhh * a.bw;
    auto row_pixel = [&](const TileInfo& ti) -> int {       // output pixel index of this thread's accumulator row
      const int pn = ti.n0 + dn, ph = ti.h0 + dhh, pw = ti.w0 + dww;
      const bool valid = (uint32_t)row < rows_a && pn < a.n && ph < a.oh && pw < a.ow;
      const long long pix = ((long long)pn * a.fh + (long long)ph * a.ostride + a.ooh) * a.fw + (long long)pw * a.ostride + a.oow;
      return valid ? (int)pix : -1;
    };
    const bool has_reads = (ep.dbg & 4) && (ep.addend || ep.gate || ep.accumulate);
    for (int pair = cluster_id; pair < a.num_pairs; pair += num_clusters) {
      const TileInfo t = decode_tile(a, pair, (int)rank);
      __syncwarp();
      rowpix[lane] = row_pixel(t);
      __syncwarp();
      if (has_reads && pair + num_clusters < a.num_pairs) {     // next tile's epilogue operands -> L2 (see prefetch_row)
        const TileInfo tn = decode_tile(a, pair + num_clusters, (int)rank);
        epi::prefetch_row(ep, row_pixel(tn), tn.n_idx * BLOCK_N + eh * (BLOCK_N / 2), BLOCK_N / 2);
      }
      if (PF) epi::pf_prologue(ep, BLOCK_N, t.n_idx * BLOCK_N, rowpix, lane, eh, pf_slot);   // first chunk's operands
      if (tracer && tr_e < 510) a.trace[1536 + tr_e++] = clock64();
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      if (tracer && tr_e < 510) a.trace[1536 + tr_e++] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
      const int m_idx = t.m_idx;
      const int stat_row = m_idx < a.tiles_w * a.tiles_h * a.tiles_n ? m_idx * 4 + ew : -1;   // phantom tile: none
      epi::drain_tile(ep, taddr, BLOCK_N, t.n_idx * BLOCK_N, stg, rowpix, lane, eh, stat_row, pf_slot, [&]() {
        tc::tc_fence_before();           // accumulator fully read: hand the TMEM stage back to the MMA warp
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  cluster_sync_all();                               // nobody leaves (or frees TMEM) while the pair is still working
  if (warp == 2) tmem2_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// choose_box is defined in conv_gemm.cu
extern int g_conv_epi_debug;
extern int g_conv_tap_outer;
extern int g_conv_next_reverse;
extern int g_conv_tap_tables;
extern int g_conv_wide_pf;
// PF build is used when the epilogue reads an addend / gate and K * taps <= this (0 = never).  Measured on B200
// (profiles/r01_v7_pf_microbench.log): faster up to K = 512 (HBM-bound 1x1 layers, 0.231 -> 0.163 ms for 256 -> 1024 with
// addend + gate), slower from K = 1024 on where the 3-stage operand ring starves the MMA pipe.
// Override: environment B200SEG_PF_MAX_K or b2_debug_set(4, v).
int g_conv_pf_max_k = -1;
// TMA epilogue of the PF build (conv_epilogue_tma.cuh): 1 = on where usable (default), 0 = register epilogue everywhere.
// Override: environment B200SEG_TMA_EPI or b2_debug_set(8, v).
int g_conv_tma_epi = -1;
int g_conv_main_stages = STAGES_MAIN;   // b2_debug_set(10, 5): the 5-stage build of the compute-bound kernel (A/B timing)
static long long* g_conv_trace = nullptr;
extern "C" void b2_debug_trace(void* buf) { g_conv_trace = static_cast<long long*>(buf); }
void b2_choose_box(int ow, int oh, int n, int max_rows, int istride, int* bw_o, int* bh_o, int* bn_o);

// Tile geometry of a launch (everything the tile walk of the three pipeline roles depends on).
static int conv2_geometry(const b2_conv_params* p, Conv2KArgs& a) {
  memset(&a, 0, sizeof(a));
  a.n = p->n; a.ih = p->ih; a.iw = p->iw; a.k = p->k; a.nb = p->nb; a.oh = p->oh; a.ow = p->ow;
  a.fh = p->fh; a.fw = p->fw; a.ldd = p->ldd; a.ostride = p->ostride; a.ooh = p->ooh; a.oow = p->oow;
  a.istride = p->istride;
  b2_choose_box(p->ow, p->oh, p->n, BLOCK_M, p->istride, &a.bw, &a.bh, &a.bn);
  a.tiles_w = (p->ow + a.bw - 1) / a.bw; a.tiles_h = (p->oh + a.bh - 1) / a.bh; a.tiles_n = (p->n + a.bn - 1) / a.bn;
  a.n_tiles_n = (p->nb + BLOCK_N - 1) / BLOCK_N;
  const int64_t m_tiles = (int64_t)a.tiles_w * a.tiles_h * a.tiles_n;
  const int64_t num_pairs = ((m_tiles + 1) / 2) * a.n_tiles_n;
  B2_REQUIRE(num_pairs < (1ll << 30), "b2_conv_gemm: too many tiles");
  a.num_pairs = (int)num_pairs;
  a.n_taps = p->n_taps;
  for (int i = 0; i < p->n_taps; ++i) {
    a.dh[i] = (short)p->taps[i * 3 + 0]; a.dw[i] = (short)p->taps[i * 3 + 1]; a.btap[i] = (short)p->taps[i * 3 + 2];
  }
  a.kblocks = (p->k + BLOCK_K - 1) / BLOCK_K;
  a.n_pass = p->n_split;
  a.fd_w = tc::make_fastdiv((uint32_t)a.tiles_w); a.fd_h = tc::make_fastdiv((uint32_t)a.tiles_h);
  a.fd_nn = tc::make_fastdiv((uint32_t)a.n_tiles_n);
  if (a.tiles_w <= tc::TAP_TABLE && a.tiles_h <= tc::TAP_TABLE && g_conv_tap_tables != 0) {
    for (int ht = 0; ht < a.tiles_h; ++ht) {
      uint32_t mk = 0;
      for (int i = 0; i < a.n_taps; ++i) {
        const int lo = ht * a.bh * a.istride + a.dh[i], hi = lo + (a.bh - 1) * a.istride;
        if (hi >= 0 && lo < a.ih) mk |= 1u << i;
      }
      a.tt.rows[ht] = (uint16_t)mk;
    }
    for (int wt = 0; wt < a.tiles_w; ++wt) {
      uint32_t mk = 0;
      for (int i = 0; i < a.n_taps; ++i) {
        const int lo = wt * a.bw * a.istride + a.dw[i], hi = lo + (a.bw - 1) * a.istride;
        if (hi >= 0 && lo < a.iw) mk |= 1u << i;
      }
      a.tt.cols[wt] = (uint16_t)mk;
    }
    a.tt.on = 1;
  }
  return B2_OK;
}

// Host-only description of the launch the CTA-pair kernel would run for `p` on `sms` SMs (no GPU, pointers are not dereferenced):
// out = int64[8] {bw, bh, bn, tile pairs, pipeline stages of the whole launch (padding-only taps skipped), stages of the busiest
// CTA pair, CTA pairs, K steps of 8 per stage}.  One stage = BLOCK_K / 8 tcgen05.mma M256 x N256 x K8 instructions of 128 tensor-pipe
// cycles each (2048 tf32 MAC per clock and SM), so   tensor-pipe busy cycles per SM = stages of its CTA pair x 512:
// divided by a profiler's sm__cycles_elapsed this is the kernel's tensor-pipe occupancy (tools/tensor_busy.py).
extern "C" int b2_conv_gemm_plan(const b2_conv_params* p, int sms, int64_t* out) {
  B2_REQUIRE(p && out && p->taps && p->n_taps >= 1 && p->n_taps <= MAX_TAPS && p->n > 0 && p->oh > 0 && p->ow > 0 && p->k > 0 &&
             p->nb > 0 && p->istride >= 1 && sms >= 2, "b2_conv_gemm_plan: bad args");
  Conv2KArgs a;
  int rc = conv2_geometry(p, a); if (rc) return rc;
  int clusters = sms / 2;
  if (clusters > a.num_pairs) clusters = a.num_pairs;
  int64_t total = 0, worst = 0;
  for (int c = 0; c < clusters; ++c) {
    int64_t mine = 0;
    for (int pair = c; pair < a.num_pairs; pair += clusters) {
      const TileInfo t = decode_tile(a, pair, 0);
      int taps = 0;
      for (int i = 0; i < a.n_taps; ++i) taps += (int)(t.tap_mask >> i & 1u);
      mine += (int64_t)taps * a.kblocks * (a.n_pass > 0 ? a.n_pass : 1);
    }
    total += mine;
    if (mine > worst) worst = mine;
  }
  out[0] = a.bw; out[1] = a.bh; out[2] = a.bn; out[3] = a.num_pairs; out[4] = total; out[5] = worst; out[6] = clusters;
  out[7] = BLOCK_K / 8;
  return B2_OK;
}

int b2_conv_gemm_2cta(const b2_conv_params* p, void* stream) {
  Conv2KArgs a;
  { int rc = conv2_geometry(p, a); if (rc) return rc; }
  a.tap_outer = g_conv_tap_outer;
  a.reverse = g_conv_next_reverse;
  a.ep.d = p->d; a.ep.ldd = p->ldd; a.ep.nb = p->nb;
  a.ep.scale = p->scale; a.ep.shift = p->shift; a.ep.addend = p->addend; a.ep.gate = p->gate; a.ep.scale2 = p->scale2;
  a.ep.ld_add = p->ld_add; a.ep.ld_gate = p->ld_gate; a.ep.relu = p->relu; a.ep.accumulate = p->accumulate;
  a.ep.stats = p->stats; a.ep.ld_stats = p->ld_stats; a.ep.sub = p->stats_sub; a.ep.ld_sub = p->ld_stats_sub;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  a.ep.dbg = g_conv_epi_debug; a.trace = g_conv_trace;
  a.ep.vec_ok = (p->ldd % 4 == 0) && al16(p->d) && (!p->addend || (p->ld_add % 4 == 0 && al16(p->addend))) &&
             (!p->gate || (p->ld_gate % 4 == 0 && al16(p->gate)));
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  {
    const uint64_t dims[4] = {(uint64_t)p->k, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n};
    const uint64_t strides[3] = {(uint64_t)p->lda * 4, (uint64_t)p->iw * p->lda * 4, (uint64_t)p->ih * p->iw * p->lda * 4};
    const uint32_t box[4] = {BLOCK_K, (uint32_t)(a.bw * p->istride), (uint32_t)(a.bh * p->istride), (uint32_t)a.bn};
    const uint32_t es[4] = {1, (uint32_t)p->istride, (uint32_t)p->istride, 1};
    int rc = tc::make_tmap_f32(&tmA, p->a, 4, dims, strides, box, es); if (rc) return rc;
    rc = tc::make_tmap_f32(&tmAlo, p->a_lo ? p->a_lo : p->a, 4, dims, strides, box, es); if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)p->k, (uint64_t)p->tb, (uint64_t)p->nb};
    const uint64_t strides[2] = {(uint64_t)p->ldb * 4, (uint64_t)p->tb * p->ldb * 4};
    const uint32_t box[3] = {BLOCK_K, 1, (uint32_t)(BLOCK_N / 2)};
    const uint32_t es[3] = {1, 1, 1};
    int rc = tc::make_tmap_f32(&tmB, p->b, 3, dims, strides, box, es); if (rc) return rc;
    rc = tc::make_tmap_f32(&tmBlo, p->b_lo ? p->b_lo : p->b, 3, dims, strides, box, es); if (rc) return rc;
  }
  CUtensorMap tmPf = tmA;                                  // placeholder unless the wide prefetch is on
  if (g_conv_wide_pf < 0) { const char* e = getenv("B200SEG_WIDE_PF"); g_conv_wide_pf = e ? atoi(e) : 0; }
  if (g_conv_wide_pf > 0 && p->n_taps == 1 && p->istride == 1 && p->k >= 512 && p->k % 256 == 0 && !g_conv_tap_outer &&
      p->taps[0] == 0 && p->taps[1] == 0) {
    const uint64_t dims[4] = {(uint64_t)p->k, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n};
    const uint64_t strides[3] = {(uint64_t)p->lda * 4, (uint64_t)p->iw * p->lda * 4, (uint64_t)p->ih * p->iw * p->lda * 4};
    const uint32_t box[4] = {256, (uint32_t)a.bw, (uint32_t)a.bh, (uint32_t)a.bn};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = tc::make_tmap_f32_linear(&tmPf, p->a, 4, dims, strides, box, es); if (rc) return rc;
    a.wide_pf = 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<STAGES_MAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(STAGES_MAIN, false)));
    B2_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<STAGES_MAIN_OLD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(STAGES_MAIN_OLD, false)));
    B2_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<STAGES_PF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(STAGES_PF, true)));
    attr_set = true;
  }
  if (g_conv_pf_max_k < 0) {
    const char* e = getenv("B200SEG_PF_MAX_K");
    g_conv_pf_max_k = e ? atoi(e) : 512;
  }
  if (g_conv_tma_epi < 0) {
    const char* e = getenv("B200SEG_TMA_EPI");
    g_conv_tma_epi = e ? atoi(e) : 1;
  }
  const bool small_k = g_conv_pf_max_k > 0 && (int64_t)p->k * p->n_taps <= g_conv_pf_max_k;
  // TMA epilogue (conv_epilogue_tma.cuh): the 32 accumulator rows of a TMEM lane quarter must be one rectangular pixel box
  // of the output tensor, written densely (unit output stride, no phase offset) through 16 B aligned pointers / pitches.
  CUtensorMap tmD = tmA, tmAdd = tmA, tmGate = tmA;        // placeholders unless the TMA epilogue is on
  {
    int ex = 0, ey = 0, ez = 0;
    const int rows_a = a.bw * a.bh * a.bn;
    if (a.bw % 32 == 0) { ex = 32; ey = 1; ez = 1; }
    else if (32 % a.bw == 0 && a.bh % (32 / a.bw) == 0) { ex = a.bw; ey = 32 / a.bw; ez = 1; }
    else if (32 % (a.bw * a.bh) == 0) { ex = a.bw; ey = a.bh; ez = 32 / (a.bw * a.bh); }
    // plain outputs (no residual / gate operand, no accumulation) stay on the deeper operand ring of the main build with the
    // register epilogue: 256->1024 at 32 images 0.138 vs 0.167 ms, 512->2048 0.409 vs 0.474 ms (profiles/r02_v25_*)
    const bool has_epilogue_traffic = p->addend || p->gate || p->accumulate;
    const bool tma_ok = g_conv_tma_epi > 0 && small_k && has_epilogue_traffic && ex > 0 && rows_a % 32 == 0 && a.ep.vec_ok && p->ostride == 1 &&
                        p->ooh == 0 && p->oow == 0 && p->oh == p->fh && p->ow == p->fw && !p->stats_sub &&
                        !(g_conv_epi_debug & 3);
    if (tma_ok) {
      const uint32_t box[4] = {32, (uint32_t)ex, (uint32_t)ey, (uint32_t)ez};
      const uint32_t es[4] = {1, 1, 1, 1};
      auto out_map = [&](CUtensorMap* m, const void* ptr, int ld) -> int {
        const uint64_t dims[4] = {(uint64_t)p->nb, (uint64_t)p->ow, (uint64_t)p->oh, (uint64_t)p->n};
        const uint64_t strides[3] = {(uint64_t)ld * 4, (uint64_t)p->ow * ld * 4, (uint64_t)p->oh * p->ow * ld * 4};
        return tc::make_tmap_f32(m, ptr, 4, dims, strides, box, es);
      };
      int rc = out_map(&tmD, p->d, p->ldd); if (rc) return rc;
      if (p->addend) { rc = out_map(&tmAdd, p->addend, p->ld_add); if (rc) return rc; }
      if (p->gate) { rc = out_map(&tmGate, p->gate, p->ld_gate); if (rc) return rc; }
      a.tma.on = 1; a.tma.ex = ex; a.tma.ey = ey; a.tma.ez = ez;
    }
  }
  const bool use_pf = a.tma.on || (small_k && (p->addend || p->gate) && a.ep.vec_ok && p->n_split == 1);
  int sms = b2_sm_count_cached();
  if (sms <= 0) return b2_fail(B2_ERR_CUDA, "b2_conv_gemm: no CUDA device");
  if (p->max_ctas > 0 && p->max_ctas < sms) sms = p->max_ctas;
  int clusters = sms / 2;
  if (clusters < 1) clusters = 1;
  if (clusters > a.num_pairs) clusters = a.num_pairs;
  if (use_pf)
    tc::launch(conv_gemm2_kernel<STAGES_PF, true>, clusters * 2, NUM_THREADS, smem_bytes(STAGES_PF, true), (cudaStream_t)stream, tmA, tmAlo, tmB, tmBlo, tmD, tmAdd, tmGate, tmPf, a);
  else if (g_conv_main_stages == STAGES_MAIN_OLD)
    tc::launch(conv_gemm2_kernel<STAGES_MAIN_OLD, false>, clusters * 2, NUM_THREADS, smem_bytes(STAGES_MAIN_OLD, false), (cudaStream_t)stream, tmA, tmAlo, tmB, tmBlo, tmD, tmAdd, tmGate, tmPf, a);
  else
    tc::launch(conv_gemm2_kernel<STAGES_MAIN, false>, clusters * 2, NUM_THREADS, smem_bytes(STAGES_MAIN, false), (cudaStream_t)stream, tmA, tmAlo, tmB, tmBlo, tmD, tmAdd, tmGate, tmPf, a);
  B2_LAUNCH_CHECK("conv_gemm2_kernel");
  return B2_OK;
}
